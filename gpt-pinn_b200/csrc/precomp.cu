// precomp.cu -- forward-mode propagation of (value, d/dx, d/dt, q, r) through a small MLP.
//
// Replaces the autograd calls of the reference's precompute (Klein-Gordon/KG_precomp.py:7-21,
// Burgers/B_precomp.py:6-16, Allen-Cahn/AC_precompute.py:7-18).  The reference's "second
// derivatives" are grad(sum(u_x + u_t)): q = u_xx + u_xt and r = u_tt + u_xt;
// the closed form below propagates exactly those mixed channels:
//   a = s(z):  a_x = s' z_x,  a_t = s' z_t,
//              q_a = s'' z_x (z_x + z_t) + s' q_z,   r_a = s'' z_t (z_t + z_x) + s' r_z
// One thread per collocation point; activations of the current layer live in a per-thread
// array (registers / L1-resident local memory), weights are staged in shared memory and read
// as warp-wide broadcasts; four output neurons are register-blocked per pass over the inputs.
#include "precomp.cuh"

namespace gp {

template <int CH> struct ChanCfg;            // which channels are carried
template <> struct ChanCfg<1> { static constexpr bool X = false, T = false, Q = false, R = false; };
template <> struct ChanCfg<3> { static constexpr bool X = true, T = true, Q = false, R = false; };
template <> struct ChanCfg<5> { static constexpr bool X = true, T = true, Q = true, R = true; };

template <int WMAX, int CH>
__global__ void __launch_bounds__(128) mlp_channels_kernel(const MlpDesc d, const float *__restrict__ xt, long long N, const MlpOut o)
{
    extern __shared__ float sw[];           // all weights then all biases, layer after layer
    int off = 0;
    int woff[MLP_MAX_LAYERS], boff[MLP_MAX_LAYERS];
    for (int l = 0; l < d.nlayers - 1; ++l) {
        const int cnt = d.layers[l] * d.layers[l + 1];
        woff[l] = off;
        for (int i = threadIdx.x; i < cnt; i += blockDim.x) sw[off + i] = d.W[l][i];
        off += cnt;
        boff[l] = off;
        for (int i = threadIdx.x; i < d.layers[l + 1]; i += blockDim.x) sw[off + i] = d.b[l][i];
        off += d.layers[l + 1];
    }
    __syncthreads();
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;

    float a[CH][WMAX], z[CH][WMAX];
    a[0][0] = xt[2 * p]; a[0][1] = xt[2 * p + 1];
    if constexpr (CH >= 3) { a[1][0] = 1.f; a[1][1] = 0.f; a[2][0] = 0.f; a[2][1] = 1.f; }
    if constexpr (CH >= 5) { a[3][0] = a[3][1] = 0.f; a[4][0] = a[4][1] = 0.f; }

    for (int l = 0; l < d.nlayers - 1; ++l) {
        const int wi = d.layers[l], wo = d.layers[l + 1];
        const float *W = sw + woff[l];
        const float *B = sw + boff[l];
        for (int j = 0; j < wo; ++j) {
            float acc[CH];
            acc[0] = B[j];
#pragma unroll
            for (int c = 1; c < CH; ++c) acc[c] = 0.f;
            const float *w = W + j * wi;
            for (int i = 0; i < wi; ++i) {
                const float wv = w[i];
#pragma unroll
                for (int c = 0; c < CH; ++c) acc[c] = fmaf(wv, a[c][i], acc[c]);
            }
#pragma unroll
            for (int c = 0; c < CH; ++c) z[c][j] = acc[c];
        }
        const bool last = (l == d.nlayers - 2);
        for (int j = 0; j < wo; ++j) {
            if (last) {
#pragma unroll
                for (int c = 0; c < CH; ++c) a[c][j] = z[c][j];
            } else {
                float s, d1, d2;
                if (d.act == 0) { float sn, cs; sincosf(z[0][j], &sn, &cs); s = cs; d1 = -sn; d2 = -cs; }
                else { s = tanhf(z[0][j]); d1 = 1.f - s * s; d2 = -2.f * s * d1; }
                a[0][j] = s;
                if constexpr (CH >= 3) {
                    const float zx = z[1][j], zt = z[2][j];
                    a[1][j] = d1 * zx;
                    a[2][j] = d1 * zt;
                    if constexpr (CH >= 5) {
                        a[3][j] = fmaf(d2 * zx, zx + zt, d1 * z[3][j]);
                        a[4][j] = fmaf(d2 * zt, zt + zx, d1 * z[4][j]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < CH; ++c)
        if (o.ptr[c]) o.ptr[c][p * o.ld[c]] = a[c][0];
}

template <int WMAX>
static cudaError_t launch_w(const MlpDesc &d, const float *xt, long long N, const MlpOut &o, int ch, size_t smem, cudaStream_t st)
{
    const int threads = 128;
    const unsigned blocks = (unsigned)((N + threads - 1) / threads);
    cudaError_t e;
    if (ch == 1) {
        auto k = mlp_channels_kernel<WMAX, 1>;
        if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k<<<blocks, threads, smem, st>>>(d, xt, N, o);
    } else if (ch == 3) {
        auto k = mlp_channels_kernel<WMAX, 3>;
        if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k<<<blocks, threads, smem, st>>>(d, xt, N, o);
    } else {
        auto k = mlp_channels_kernel<WMAX, 5>;
        if ((e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        k<<<blocks, threads, smem, st>>>(d, xt, N, o);
    }
    return cudaGetLastError();
}

cudaError_t launch_mlp_channels(const MlpDesc &d, const float *xt, long long N, const MlpOut &o, cudaStream_t st)
{
    if (N <= 0) return cudaSuccess;
    int maxw = 0;
    size_t fl = 0;
    for (int l = 0; l < d.nlayers; ++l) maxw = d.layers[l] > maxw ? d.layers[l] : maxw;
    for (int l = 0; l < d.nlayers - 1; ++l) fl += (size_t)d.layers[l] * d.layers[l + 1] + d.layers[l + 1];
    const size_t smem = fl * sizeof(float);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    const int ch = (o.ptr[3] || o.ptr[4]) ? 5 : ((o.ptr[1] || o.ptr[2]) ? 3 : 1);
    if (maxw <= 20) return launch_w<20>(d, xt, N, o, ch, smem, st);
    if (maxw <= 40) return launch_w<40>(d, xt, N, o, ch, smem, st);
    if (maxw <= 64) return launch_w<64>(d, xt, N, o, ch, smem, st);
    if (maxw <= 128) return launch_w<128>(d, xt, N, o, ch, smem, st);
    return cudaErrorInvalidValue;
}

}  // namespace gp
