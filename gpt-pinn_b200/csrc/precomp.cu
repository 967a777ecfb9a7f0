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


// ---------------------------------------------------------------------------------------------------
// Tiled variant for the narrow PINNs (hidden width <= 40: Klein-Gordon [2,40,40,1], Burgers [2,20,20,20,20,1]).
// Block = 4 warps = 4 neuron groups, tile = 32 points (lane = point).  A thread keeps the channels of its own
// neurons in registers; the previous layer's activations of all neurons are staged in shared memory and each
// layer is a small GEMM whose weights are read as warp-wide float4 broadcasts (same structure as the forward
// half of pinn_train.cu).  Persistent over tiles; weights are loaded into shared memory once per block.
// ---------------------------------------------------------------------------------------------------
constexpr int TNG = 4, TPT = 32;

template <int CHN>
__device__ __forceinline__ void tile_act(int act, const float (&z)[CHN], float (&a)[CHN])
{
    float s0, s1, s2;
    if (act == 0) { float sn, cs; sincosf(z[0], &sn, &cs); s0 = cs; s1 = -sn; s2 = -cs; }
    else { s0 = tanhf(z[0]); s1 = 1.f - s0 * s0; s2 = -2.f * s0 * s1; }
    a[0] = s0;
    if constexpr (CHN >= 3) { a[1] = s1 * z[1]; a[2] = s1 * z[2]; }
    if constexpr (CHN >= 5) {
        a[3] = fmaf(s2 * z[1], z[1] + z[2], s1 * z[3]);
        a[4] = fmaf(s2 * z[2], z[2] + z[1], s1 * z[4]);
    }
}

template <int WMAX, int NH, int CHN>
__global__ void __launch_bounds__(TNG * 32, 4) mlp_channels_tile_kernel(const MlpDesc d, const float *__restrict__ xt, long long N, const MlpOut o)
{
    constexpr int GJ = WMAX / TNG, GJP = (GJ + 3) & ~3;
    constexpr int NS = CHN * TPT + 1;
    extern __shared__ __align__(16) float sm[];
    int woff[NH + 1], boff[NH + 1], P = 0;
#pragma unroll
    for (int l = 0; l <= NH; ++l) { woff[l] = P; P += d.layers[l] * d.layers[l + 1]; boff[l] = P; P += d.layers[l + 1]; }
    const int P4 = (P + 3) & ~3;
    float *w_s = sm;
    float *wt_s = w_s + P4;                                   // [NH-1][WMAX][TNG][GJP]
    float *A_s = wt_s + (NH - 1) * WMAX * TNG * GJP;          // [WMAX][NS]
    float *U_s = A_s + WMAX * NS;                             // [TNG][CHN][TPT]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, j0 = warp * GJ;
#pragma unroll
    for (int l = 0; l <= NH; ++l) {
        const int nw = d.layers[l] * d.layers[l + 1];
        for (int i = threadIdx.x; i < nw; i += blockDim.x) w_s[woff[l] + i] = d.W[l][i];
        for (int i = threadIdx.x; i < d.layers[l + 1]; i += blockDim.x) w_s[boff[l] + i] = d.b[l][i];
    }
    __syncthreads();
    for (int l = 1; l < NH; ++l) {
        const float *W = w_s + woff[l];
        const int wi = d.layers[l], wo = d.layers[l + 1];
        float *wt = wt_s + (l - 1) * WMAX * TNG * GJP;
        for (int e = threadIdx.x; e < WMAX * TNG * GJP; e += blockDim.x) {
            const int a = e / (TNG * GJP), g = (e / GJP) % TNG, s = e % GJP, b = g * GJ + s;
            wt[e] = (s < GJ && b < wo && a < wi) ? W[b * wi + a] : 0.f;
        }
    }
    __syncthreads();
    const long long ntiles = (N + TPT - 1) / TPT;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long p = tile * TPT + lane;
        const bool live = p < N;
        const float x = live ? xt[2 * p] : 0.f, t = live ? xt[2 * p + 1] : 0.f;
        float z[CHN][GJ];
        {
            const float *W = w_s + woff[0], *Bv = w_s + boff[0];
#pragma unroll
            for (int jj = 0; jj < GJ; ++jj) {
                const int j = j0 + jj;
                const bool ok = j < d.layers[1];
                const float wx = ok ? W[j * 2] : 0.f, w1 = ok ? W[j * 2 + 1] : 0.f;
                z[0][jj] = ok ? fmaf(wx, x, fmaf(w1, t, Bv[j])) : 0.f;
                if constexpr (CHN >= 3) { z[1][jj] = wx; z[2][jj] = w1; }
                if constexpr (CHN >= 5) { z[3][jj] = 0.f; z[4][jj] = 0.f; }
            }
        }
#pragma unroll
        for (int l = 0; l < NH; ++l) {
#pragma unroll
            for (int jj = 0; jj < GJ; ++jj) {
                float zz[CHN], aa[CHN];
#pragma unroll
                for (int c = 0; c < CHN; ++c) zz[c] = z[c][jj];
                tile_act<CHN>(d.act, zz, aa);
                const bool ok = (j0 + jj) < d.layers[l + 1];
#pragma unroll
                for (int c = 0; c < CHN; ++c) A_s[(j0 + jj) * NS + c * TPT + lane] = ok ? aa[c] : 0.f;
            }
            __syncthreads();
            if (l < NH - 1) {
                const float *wt = wt_s + l * WMAX * TNG * GJP + warp * GJP;
                const float *Bv = w_s + boff[l + 1];
#pragma unroll
                for (int jj = 0; jj < GJ; ++jj) {
                    z[0][jj] = (j0 + jj) < d.layers[l + 2] ? Bv[j0 + jj] : 0.f;
#pragma unroll
                    for (int c = 1; c < CHN; ++c) z[c][jj] = 0.f;
                }
                const int wi = d.layers[l + 1];
                for (int i = 0; i < wi; ++i) {
                    float av[CHN], wv[GJP];
#pragma unroll
                    for (int c = 0; c < CHN; ++c) av[c] = A_s[i * NS + c * TPT + lane];
#pragma unroll
                    for (int s = 0; s < GJP; s += 4) {
                        const float4 w4 = *reinterpret_cast<const float4 *>(wt + i * TNG * GJP + s);
                        wv[s] = w4.x; wv[s + 1] = w4.y; wv[s + 2] = w4.z; wv[s + 3] = w4.w;
                    }
#pragma unroll
                    for (int jj = 0; jj < GJ; ++jj)
#pragma unroll
                        for (int c = 0; c < CHN; ++c) z[c][jj] = fmaf(wv[jj], av[c], z[c][jj]);
                }
                __syncthreads();
            }
        }
        {
            const float *W = w_s + woff[NH];
            float pu[CHN];
#pragma unroll
            for (int c = 0; c < CHN; ++c) pu[c] = 0.f;
#pragma unroll
            for (int jj = 0; jj < GJ; ++jj) {
                const float wv = (j0 + jj) < d.layers[NH] ? W[j0 + jj] : 0.f;
#pragma unroll
                for (int c = 0; c < CHN; ++c) pu[c] = fmaf(wv, A_s[(j0 + jj) * NS + c * TPT + lane], pu[c]);
            }
#pragma unroll
            for (int c = 0; c < CHN; ++c) U_s[(warp * CHN + c) * TPT + lane] = pu[c];
            __syncthreads();
            if (warp == 0 && live) {
#pragma unroll
                for (int c = 0; c < CHN; ++c) {
                    float s = c == 0 ? w_s[boff[NH]] : 0.f;
#pragma unroll
                    for (int g = 0; g < TNG; ++g) s += U_s[(g * CHN + c) * TPT + lane];
                    if (o.ptr[c]) o.ptr[c][p * o.ld[c]] = s;
                }
            }
            __syncthreads();
        }
    }
}

template <int WMAX, int NH>
static cudaError_t launch_tile(const MlpDesc &d, const float *xt, long long N, const MlpOut &o, int ch, cudaStream_t st)
{
    constexpr int GJ = WMAX / TNG, GJP = (GJ + 3) & ~3;
    int P = 0;
    for (int l = 0; l <= NH; ++l) P += d.layers[l] * d.layers[l + 1] + d.layers[l + 1];
    const size_t smem = ((size_t)((P + 3) & ~3) + (size_t)(NH - 1) * WMAX * TNG * GJP + (size_t)WMAX * (ch * TPT + 1) + (size_t)TNG * ch * TPT + 8) * sizeof(float);
    void *kern = ch == 1 ? (void *)mlp_channels_tile_kernel<WMAX, NH, 1> : (ch == 3 ? (void *)mlp_channels_tile_kernel<WMAX, NH, 3> : (void *)mlp_channels_tile_kernel<WMAX, NH, 5>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, nsm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const long long ntiles = (N + TPT - 1) / TPT;
    const unsigned blocks = (unsigned)(ntiles < 4LL * nsm ? ntiles : 4LL * nsm);
    MlpDesc dd = d; MlpOut oo = o; const float *x2 = xt; long long n2 = N;
    void *args[] = {&dd, &x2, &n2, &oo};
    return cudaLaunchKernel(kern, dim3(blocks), dim3(TNG * 32), args, smem, st);
}


// ---------------------------------------------------------------------------------------------------
// Wide networks (hidden width 41..128: Allen-Cahn's [2,128,128,128,128,1] tanh shell, AC_models.py:56-78).
// A CTA of 256 threads takes a tile of 64 points.  The activations of all 128 (padded) neurons, all channels, live in
// shared memory as A[neuron][channel][point]; every hidden->hidden layer is a [128 x 128] x [128 x CH*64] GEMM done
// in registers: thread (jg, cg) owns 8 output neurons x (CH channels x 4 points) and walks the 128 inputs four at a
// time -- per 4 inputs it reads 8 float4 of the layer's weight matrix (row-major as torch stores it; the 16 threads
// that share jg broadcast) and 4*CH float4 of activations (conflict-free: a warp reads 64 consecutive points), i.e.
// 32*CH FMAs per (2 + CH) LDS.128.  The activation function then maps the channels of the thread's own (neuron, point)
// pairs in registers and writes them back in place.  Weights of one layer at a time are staged in shared memory
// (64 KB, L2-resident source).  Channels: CH = 1 (v), 3 (v, x, t) or 4 (v, x, t, q = u_xx + u_xt); the r channel of a
// wide network goes through the one-thread-per-point kernel above (no shipped problem needs it).
// ---------------------------------------------------------------------------------------------------
constexpr int WW = 128, WPTS = 64, WTHREADS = 256;

template <int CH>
__global__ void __launch_bounds__(WTHREADS, 1) mlp_channels_wide_kernel(const MlpDesc d, const float *__restrict__ xt, long long N, const MlpOut o)
{
    extern __shared__ __align__(16) float smw[];
    float *A_s = smw;                                   // [WW][CH][WPTS]
    float *W_s = A_s + WW * CH * WPTS;                  // [WW][WW] row-major [out j][in i], zero padded; later the output partials
    const int tid = threadIdx.x, jg = tid >> 4, cg = tid & 15, j0 = 8 * jg, p0 = 4 * cg;
    const int NH = d.nlayers - 2;
    const long long ntiles = (N + WPTS - 1) / WPTS;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        float acc[8][CH][4];
        // ---- layer 0: inputs (x, t); channels of z known in closed form
        {
            float xs[4], ts[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const long long p = tile * WPTS + p0 + k;
                xs[k] = p < N ? xt[2 * p] : 0.f;
                ts[k] = p < N ? xt[2 * p + 1] : 0.f;
            }
            const int w1 = d.layers[1];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int j = j0 + jj;
                const bool ok = j < w1;
                const float wx = ok ? d.W[0][2 * j] : 0.f, wt = ok ? d.W[0][2 * j + 1] : 0.f, bb = ok ? d.b[0][j] : 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    acc[jj][0][k] = fmaf(wx, xs[k], fmaf(wt, ts[k], bb));
                    if constexpr (CH >= 3) { acc[jj][1][k] = wx; acc[jj][2][k] = wt; }
                    if constexpr (CH >= 4) acc[jj][3][k] = 0.f;
                }
            }
        }
        for (int l = 0; l < NH; ++l) {
            // ---- activation of hidden layer l on the thread's own (neuron, point) pairs, written to A in place
            const int wl = d.layers[l + 1];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const bool ok = (j0 + jj) < wl;
                float a[CH][4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float z0 = acc[jj][0][k];
                    float s0, s1, s2;
                    if (d.act == 0) { float sn, cs; sincosf(z0, &sn, &cs); s0 = cs; s1 = -sn; s2 = -cs; }
                    else { s0 = tanhf(z0); s1 = 1.f - s0 * s0; s2 = -2.f * s0 * s1; }
                    a[0][k] = ok ? s0 : 0.f;
                    if constexpr (CH >= 3) {
                        const float zx = acc[jj][1][k], zt = acc[jj][2][k];
                        a[1][k] = ok ? s1 * zx : 0.f;
                        a[2][k] = ok ? s1 * zt : 0.f;
                        if constexpr (CH >= 4) a[3][k] = ok ? fmaf(s2 * zx, zx + zt, s1 * acc[jj][3][k]) : 0.f;
                    }
                }
#pragma unroll
                for (int c = 0; c < CH; ++c)
                    *reinterpret_cast<float4 *>(A_s + ((j0 + jj) * CH + c) * WPTS + p0) = make_float4(a[c][0], a[c][1], a[c][2], a[c][3]);
            }
            if (l == NH - 1) break;
            // ---- stage the weights of the map hidden l -> hidden l+1 (torch layout [out][in]), zero padded to 128 x 128
            {
                const int wi = d.layers[l + 1], wo = d.layers[l + 2];
                const float *Wg = d.W[l + 1];
                for (int e = tid; e < WW * WW; e += WTHREADS) {
                    const int j = e >> 7, i = e & 127;
                    W_s[e] = (j < wo && i < wi) ? Wg[j * wi + i] : 0.f;
                }
            }
            __syncthreads();                              // A (all neurons) and W are in place
            {
                const int wo = d.layers[l + 2];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const float bb = (j0 + jj) < wo ? d.b[l + 1][j0 + jj] : 0.f;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        acc[jj][0][k] = bb;
#pragma unroll
                        for (int c = 1; c < CH; ++c) acc[jj][c][k] = 0.f;
                    }
                }
            }
#pragma unroll 1
            for (int i4 = 0; i4 < WW; i4 += 4) {
                float4 wq[8];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) wq[jj] = *reinterpret_cast<const float4 *>(W_s + (j0 + jj) * WW + i4);
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    float4 av[CH];
#pragma unroll
                    for (int c = 0; c < CH; ++c) av[c] = *reinterpret_cast<const float4 *>(A_s + ((i4 + ii) * CH + c) * WPTS + p0);
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const float w = ii == 0 ? wq[jj].x : (ii == 1 ? wq[jj].y : (ii == 2 ? wq[jj].z : wq[jj].w));
#pragma unroll
                        for (int c = 0; c < CH; ++c) {
                            acc[jj][c][0] = fmaf(w, av[c].x, acc[jj][c][0]);
                            acc[jj][c][1] = fmaf(w, av[c].y, acc[jj][c][1]);
                            acc[jj][c][2] = fmaf(w, av[c].z, acc[jj][c][2]);
                            acc[jj][c][3] = fmaf(w, av[c].w, acc[jj][c][3]);
                        }
                    }
                }
            }
            __syncthreads();                              // every read of A and W is done: both may be overwritten
        }
        __syncthreads();                                  // A holds the activations of the last hidden layer
        // ---- output unit: partial over the thread's 8 neurons, then over the 16 neuron groups through shared memory
        {
            const int wl = d.layers[NH];
            float pu[CH][4];
#pragma unroll
            for (int c = 0; c < CH; ++c)
#pragma unroll
                for (int k = 0; k < 4; ++k) pu[c][k] = 0.f;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const float w = (j0 + jj) < wl ? d.W[NH][j0 + jj] : 0.f;
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const float4 a = *reinterpret_cast<const float4 *>(A_s + ((j0 + jj) * CH + c) * WPTS + p0);
                    pu[c][0] = fmaf(w, a.x, pu[c][0]); pu[c][1] = fmaf(w, a.y, pu[c][1]);
                    pu[c][2] = fmaf(w, a.z, pu[c][2]); pu[c][3] = fmaf(w, a.w, pu[c][3]);
                }
            }
            float *R = W_s;                               // [16 jg][CH][WPTS]
#pragma unroll
            for (int c = 0; c < CH; ++c)
                *reinterpret_cast<float4 *>(R + (jg * CH + c) * WPTS + p0) = make_float4(pu[c][0], pu[c][1], pu[c][2], pu[c][3]);
            __syncthreads();
            for (int e = tid; e < CH * WPTS; e += WTHREADS) {
                const int c = e / WPTS, k = e % WPTS;
                const long long p = tile * WPTS + k;
                float s = c == 0 ? d.b[NH][0] : 0.f;
#pragma unroll
                for (int g = 0; g < 16; ++g) s += R[(g * CH + c) * WPTS + k];
                if (p < N && o.ptr[c]) o.ptr[c][p * o.ld[c]] = s;
            }
            __syncthreads();                              // R (= W_s) and A are rewritten by the next tile
        }
    }
}

template <int CH>
static cudaError_t launch_wide(const MlpDesc &d, const float *xt, long long N, const MlpOut &o, cudaStream_t st)
{
    const size_t smem = ((size_t)WW * CH * WPTS + (size_t)WW * WW) * sizeof(float);
    auto kern = mlp_channels_wide_kernel<CH>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, nsm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const long long ntiles = (N + WPTS - 1) / WPTS;
    const unsigned blocks = (unsigned)(ntiles < nsm ? ntiles : nsm);
    kern<<<blocks, WTHREADS, smem, st>>>(d, xt, N, o);
    return cudaGetLastError();
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
    const int nh = d.nlayers - 2;
    int maxh = 0;
    for (int l = 1; l < d.nlayers - 1; ++l) maxh = d.layers[l] > maxh ? d.layers[l] : maxh;
    if (nh >= 1 && maxh <= 20 && maxh % 4 == 0) {
        if (nh == 1) return launch_tile<20, 1>(d, xt, N, o, ch, st);
        if (nh == 2) return launch_tile<20, 2>(d, xt, N, o, ch, st);
        if (nh == 3) return launch_tile<20, 3>(d, xt, N, o, ch, st);
        if (nh == 4) return launch_tile<20, 4>(d, xt, N, o, ch, st);
    } else if (nh >= 1 && maxh <= 40 && nh <= 3) {
        if (nh == 1) return launch_tile<40, 1>(d, xt, N, o, ch, st);
        if (nh == 2) return launch_tile<40, 2>(d, xt, N, o, ch, st);
        if (nh == 3) return launch_tile<40, 3>(d, xt, N, o, ch, st);
    }
    if (nh >= 1 && maxh > 40 && maxh <= WW && !o.ptr[4]) {       // wide hidden layers (Allen-Cahn): tiled register GEMM
        if (o.ptr[3]) return launch_wide<4>(d, xt, N, o, st);
        if (o.ptr[1] || o.ptr[2]) return launch_wide<3>(d, xt, N, o, st);
        return launch_wide<1>(d, xt, N, o, st);
    }
    if (maxw <= 20) return launch_w<20>(d, xt, N, o, ch, smem, st);
    if (maxw <= 40) return launch_w<40>(d, xt, N, o, ch, smem, st);
    if (maxw <= 64) return launch_w<64>(d, xt, N, o, ch, smem, st);
    if (maxw <= 128) return launch_w<128>(d, xt, N, o, ch, smem, st);
    return cudaErrorInvalidValue;
}

}  // namespace gp
