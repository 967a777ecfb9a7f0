// pinn_train.cu -- fused full-PINN training (forward + derivative residual + backward + Adam) in ONE
// persistent cooperative kernel for all epochs.
//
// Replaces the reference's pinn_train loops (Klein-Gordon/KG_train.py:48-59 with KG_models.py:84-129,
// Burgers/B_train.py:76-90 with B_models.py:70-105): per epoch the reference builds a triple-nested
// autograd graph (grad of grad of grad) with ~100 kernel launches and an optimizer step; here
//   phase 1  every warp takes 32 collocation / IC / BC points, one point per lane:
//            forward-mode propagation of (u, u_x, u_t, q, r) through the MLP (q = u_xx+u_xt,
//            r = u_tt+u_xt: the reference's mixed second derivatives), the PDE
//            residual and its adjoints, then the hand-derived reverse pass through the same
//            five channels (needs sigma', sigma'', sigma'''); weight gradients are formed per
//            warp as small outer-product GEMMs staged through shared memory and accumulated in a
//            per-warp shared buffer (each weight owned by exactly one lane: no atomics);
//   grid.sync
//   phase 2  a deterministic reduction of the per-warp partial gradients and losses, the
//            tolerance test of B_train.py:84 (checked BEFORE the step) and torch.optim.Adam's
//            update (beta = (0.9, 0.999), eps = 1e-8, bias correction in double);
//   grid.sync
// so the weights never leave the device and the host launches once per training run.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "pinn_train.cuh"

namespace cg = cooperative_groups;

namespace gp {

constexpr int CH = 5;                 // v, x, t, q, r
constexpr int WARPS = 8;              // warps per block
constexpr int JG = 4, IG = 8;         // lane tiling of a weight matrix: 4 output groups x 8 input groups

template <int WMAX>
struct Lane {
    float z[PINN_MAX_HIDDEN][CH][WMAX];      // pre-activations of every hidden layer (this lane's point)
};

__device__ __forceinline__ void act_derivs(int act, float z, float &s0, float &s1, float &s2, float &s3)
{
    if (act == 0) {                   // cos
        float sn, cs;
        sincosf(z, &sn, &cs);
        s0 = cs; s1 = -sn; s2 = -cs; s3 = sn;
    } else {                          // tanh
        const float s = tanhf(z);
        s0 = s; s1 = 1.f - s * s; s2 = -2.f * s * s1; s3 = s1 * (6.f * s * s - 2.f);
    }
}

// activations of hidden layer `l` (l >= 1) or the network inputs (l == 0) for channel set of this lane
template <int WMAX>
__device__ __forceinline__ void activations(const PinnDesc &d, const Lane<WMAX> &L, int l, float x, float t, int j, float (&a)[CH])
{
    if (l == 0) {
        a[0] = j == 0 ? x : t; a[1] = j == 0 ? 1.f : 0.f; a[2] = j == 0 ? 0.f : 1.f; a[3] = 0.f; a[4] = 0.f;
        return;
    }
    const float zv = L.z[l - 1][0][j], zx = L.z[l - 1][1][j], zt = L.z[l - 1][2][j];
    float s0, s1, s2, s3;
    act_derivs(d.act, zv, s0, s1, s2, s3);
    a[0] = s0;
    a[1] = s1 * zx;
    a[2] = s1 * zt;
    a[3] = fmaf(s2 * zx, zx + zt, s1 * L.z[l - 1][3][j]);
    a[4] = fmaf(s2 * zt, zt + zx, s1 * L.z[l - 1][4][j]);
}

template <int WMAX>
__global__ void __launch_bounds__(WARPS * 32, 1) pinn_train_kernel(const PinnDesc d)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ float smem[];
    const int nl = d.nlayers;                       // entries of layers[]; nl-1 linear maps, nl-2 hidden layers
    const int P = d.n_params;
    float *w_s = smem;                              // [P] weights + biases, packed layer after layer (W then b)
    float *warp_base = smem + ((P + 3) & ~3);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_warp = ((P + 3) & ~3) + 2 * 32 * WMAX + 8;
    float *g_s = warp_base + (size_t)warp * per_warp;            // [P] this warp's gradient partial
    float *zb_s = g_s + ((P + 3) & ~3);                          // [32][WMAX] staged z-bar of one channel
    float *ap_s = zb_s + 32 * WMAX;                              // [32][WMAX] staged previous activations of one channel
    const int gwarp = blockIdx.x * WARPS + warp;
    const int nwarps = gridDim.x * WARPS;
    const int jg = lane / IG, ig = lane % IG;

    int woff[PINN_MAX_HIDDEN + 1], boff[PINN_MAX_HIDDEN + 1];
    {
        int off = 0;
        for (int l = 0; l < nl - 1; ++l) {
            woff[l] = off; off += d.layers[l] * d.layers[l + 1];
            boff[l] = off; off += d.layers[l + 1];
        }
    }
    const int chunksR = (d.NR + 31) / 32, chunksI = (d.NI + 31) / 32, chunksB = (d.NB + 31) / 32;
    const int nchunks = chunksR + chunksI + chunksB;

    constexpr int TJ = (WMAX + JG - 1) / JG, TI = (WMAX + IG - 1) / IG;
    Lane<WMAX> L;
    for (int epoch = 1; epoch <= d.epochs; ++epoch) {
        // -------------------------------------------------------------- phase 1
        for (int i = threadIdx.x; i < P; i += blockDim.x) w_s[i] = d.params[i];
        for (int i = lane; i < P; i += 32) g_s[i] = 0.f;
        __syncthreads();
        float lossR = 0.f, lossI1 = 0.f, lossI2 = 0.f, lossB = 0.f;

        for (int chunk = gwarp; chunk < nchunks; chunk += nwarps) {
            int kind, p;                              // 0 residual, 1 IC, 2 BC
            if (chunk < chunksR) { kind = 0; p = chunk * 32 + lane; }
            else if (chunk < chunksR + chunksI) { kind = 1; p = (chunk - chunksR) * 32 + lane; }
            else { kind = 2; p = (chunk - chunksR - chunksI) * 32 + lane; }
            const int np = kind == 0 ? d.NR : (kind == 1 ? d.NI : d.NB);
            const bool live = p < np;
            const float *xt = kind == 0 ? d.xt_resid : (kind == 1 ? d.IC_xt : d.BC_xt);
            const float x = live ? xt[2 * p] : 0.f, t = live ? xt[2 * p + 1] : 0.f;

            // ---- forward: five channels through every layer
            float u[CH];
            float a[CH][WMAX];                               // input activations of the current linear map
            a[0][0] = x; a[1][0] = 1.f; a[2][0] = 0.f; a[3][0] = 0.f; a[4][0] = 0.f;
            a[0][1] = t; a[1][1] = 0.f; a[2][1] = 1.f; a[3][1] = 0.f; a[4][1] = 0.f;
            for (int l = 0; l < nl - 1; ++l) {
                const int wi = d.layers[l], wo = d.layers[l + 1];
                const float *W = w_s + woff[l], *B = w_s + boff[l];
                for (int j = 0; j < wo; ++j) {
                    float acc[CH] = {B[j], 0.f, 0.f, 0.f, 0.f};
                    for (int i = 0; i < wi; ++i) {
                        const float wv = W[j * wi + i];
#pragma unroll
                        for (int c = 0; c < CH; ++c) acc[c] = fmaf(wv, a[c][i], acc[c]);
                    }
                    if (l < nl - 2) {
#pragma unroll
                        for (int c = 0; c < CH; ++c) L.z[l][c][j] = acc[c];
                    } else {
#pragma unroll
                        for (int c = 0; c < CH; ++c) u[c] = acc[c];
                    }
                }
                if (l < nl - 2)
                    for (int j = 0; j < wo; ++j) {
                        float av[CH];
                        activations<WMAX>(d, L, l + 1, x, t, j, av);
#pragma unroll
                        for (int c = 0; c < CH; ++c) a[c][j] = av[c];
                    }
            }

            // ---- loss terms and output adjoints
            float ub[CH] = {0.f, 0.f, 0.f, 0.f, 0.f};
            if (live) {
                if (kind == 0) {
                    const float fh = d.f_hat ? d.f_hat[p] : 0.f;
                    float f;
                    if (d.pde == 0) f = u[4] + d.pa * u[3] + d.pb * u[0] + d.pc * u[0] * u[0] + d.xcos[p];   // KG_models.py:104-107
                    else f = u[2] + (u[0] * u[1] + d.pa * u[3]);                                          // B_models.py:91-94, pa = -nu
                    const float e = f - fh;
                    lossR = fmaf(e, e, lossR);
                    const float eb = 2.f * e * d.invR;
                    if (d.pde == 0) { ub[4] = eb; ub[3] = d.pa * eb; ub[0] = (d.pb + 2.f * d.pc * u[0]) * eb; }
                    else { ub[2] = eb; ub[0] = u[1] * eb; ub[1] = u[0] * eb; ub[3] = d.pa * eb; }
                } else if (kind == 1) {
                    const float e1 = u[0] - d.IC_u1[p];
                    lossI1 = fmaf(e1, e1, lossI1);
                    ub[0] = 2.f * e1 * d.invI;
                    if (d.pde == 0) {
                        const float e2 = u[2] - (d.IC_u2 ? d.IC_u2[p] : 0.f);
                        lossI2 = fmaf(e2, e2, lossI2);
                        ub[2] = 2.f * e2 * d.invI;
                    }
                } else {
                    const float e = u[0] - d.BC_u[p];
                    lossB = fmaf(e, e, lossB);
                    ub[0] = 2.f * e * d.invB;
                }
            }

            // ---- reverse pass.  zbar = adjoint of the pre-activations of the layer being processed
            float zbar[CH][WMAX];
#pragma unroll
            for (int c = 0; c < CH; ++c) zbar[c][0] = ub[c];           // output layer: one unit, linear
            for (int l = nl - 2; l >= 0; --l) {
                const int wi = d.layers[l], wo = d.layers[l + 1];
                const float *W = w_s + woff[l];
                // (1) weight / bias gradients of layer l:  Wbar[j][i] += sum_{c, points} zbar_c[j] * a_c^{(l)}[i]
                //     a small GEMM per warp: lane (jg, ig) owns a TJ x TI block of the weight matrix
                for (int i = 0; i < wi; ++i) {
                    float av[CH];
                    activations<WMAX>(d, L, l, x, t, i, av);
#pragma unroll
                    for (int c = 0; c < CH; ++c) a[c][i] = av[c];
                }
                float acc[TJ][TI];
#pragma unroll
                for (int tj = 0; tj < TJ; ++tj)
#pragma unroll
                    for (int ti = 0; ti < TI; ++ti) acc[tj][ti] = 0.f;
                for (int c = 0; c < CH; ++c) {
                    for (int j = 0; j < wo; ++j) zb_s[lane * WMAX + j] = zbar[c][j];
                    for (int j = wo; j < WMAX; ++j) zb_s[lane * WMAX + j] = 0.f;
                    for (int i = 0; i < wi; ++i) ap_s[lane * WMAX + i] = a[c][i];
                    for (int i = wi; i < WMAX; ++i) ap_s[lane * WMAX + i] = 0.f;
                    __syncwarp();
                    for (int q = 0; q < 32; ++q) {
                        float zv[TJ], av[TI];
#pragma unroll
                        for (int tj = 0; tj < TJ; ++tj) zv[tj] = zb_s[q * WMAX + jg * TJ + tj];
#pragma unroll
                        for (int ti = 0; ti < TI; ++ti) { const int i = ig * TI + ti; av[ti] = i < WMAX ? ap_s[q * WMAX + i] : 0.f; }
#pragma unroll
                        for (int tj = 0; tj < TJ; ++tj)
#pragma unroll
                            for (int ti = 0; ti < TI; ++ti) acc[tj][ti] = fmaf(zv[tj], av[ti], acc[tj][ti]);
                    }
                    if (c == 0)                                       // bias: only the value channel carries b
                        for (int j = lane; j < wo; j += 32) {
                            float bs = 0.f;
                            for (int q = 0; q < 32; ++q) bs += zb_s[q * WMAX + j];
                            g_s[boff[l] + j] += bs;
                        }
                    __syncwarp();
                }
#pragma unroll
                for (int tj = 0; tj < TJ; ++tj) {
                    const int j = jg * TJ + tj;
#pragma unroll
                    for (int ti = 0; ti < TI; ++ti) {
                        const int i = ig * TI + ti;
                        if (j < wo && i < wi) g_s[woff[l] + j * wi + i] += acc[tj][ti];
                    }
                }
                if (l == 0) break;
                // (2) adjoint of the activations of hidden layer l (index l-1 in L.z): abar_c[i] = sum_j W[j][i] zbar_c[j]
                //     then through the activation to the pre-activations of that hidden layer
                float nz[CH][WMAX];
                for (int i = 0; i < wi; ++i) {
                    float ab[CH] = {0.f, 0.f, 0.f, 0.f, 0.f};
                    for (int j = 0; j < wo; ++j) {
                        const float wv = W[j * wi + i];
#pragma unroll
                        for (int c = 0; c < CH; ++c) ab[c] = fmaf(wv, zbar[c][j], ab[c]);
                    }
                    const float zv = L.z[l - 1][0][i], zx = L.z[l - 1][1][i], zt = L.z[l - 1][2][i];
                    const float zq = L.z[l - 1][3][i], zr = L.z[l - 1][4][i];
                    float s0, s1, s2, s3;
                    act_derivs(d.act, zv, s0, s1, s2, s3);
                    nz[3][i] = s1 * ab[3];
                    nz[4][i] = s1 * ab[4];
                    nz[1][i] = fmaf(s1, ab[1], s2 * fmaf(ab[3], 2.f * zx + zt, ab[4] * zt));
                    nz[2][i] = fmaf(s1, ab[2], s2 * fmaf(ab[3], zx, ab[4] * (2.f * zt + zx)));
                    nz[0][i] = s1 * ab[0] + s2 * (zx * ab[1] + zt * ab[2] + zq * ab[3] + zr * ab[4])
                             + s3 * (ab[3] * zx * (zx + zt) + ab[4] * zt * (zt + zx));
                }
                for (int i = 0; i < wi; ++i)
#pragma unroll
                    for (int c = 0; c < CH; ++c) zbar[c][i] = nz[c][i];
            }
        }

        // ---- publish this warp's partials
        for (int off = 16; off; off >>= 1) {
            lossR += __shfl_xor_sync(0xffffffffu, lossR, off);
            lossI1 += __shfl_xor_sync(0xffffffffu, lossI1, off);
            lossI2 += __shfl_xor_sync(0xffffffffu, lossI2, off);
            lossB += __shfl_xor_sync(0xffffffffu, lossB, off);
        }
        __syncwarp();
        if (gwarp < d.n_partials) {
            float *dst = d.partials + (size_t)gwarp * d.partial_stride;
            for (int i = lane; i < P; i += 32) dst[i] = g_s[i];
            if (lane == 0) { dst[P] = lossR; dst[P + 1] = lossI1; dst[P + 2] = lossI2; dst[P + 3] = lossB; }
        }
        grid.sync();

        // -------------------------------------------------------------- phase 2: reduce, tolerance, Adam
        const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
        if (gtid == 0) {
            float s[4] = {0.f, 0.f, 0.f, 0.f};
            for (int w = 0; w < d.n_partials; ++w)
                for (int k = 0; k < 4; ++k) s[k] += d.partials[(size_t)w * d.partial_stride + P + k];
            float loss = s[0] * d.invR + s[1] * d.invI;
            if (d.pde == 0) loss += s[2] * d.invI;
            loss += s[3] * d.invB;
            d.status[0] = loss;                                       // loss of the weights BEFORE this epoch's step
            d.status[1] = __int_as_float(epoch);
            if (d.loss_trace && d.trace_every > 0 && (epoch - 1) % d.trace_every == 0) d.loss_trace[(epoch - 1) / d.trace_every] = loss;
            d.stop[0] = (d.tol >= 0.f && loss < d.tol) ? 1 : 0;       // B_train.py:84: break before the step
        }
        grid.sync();
        const bool stop = d.stop[0] != 0;
        if (!stop) {
            const double bc1 = 1.0 - pow((double)d.beta1, (double)epoch);
            const double bc2 = 1.0 - pow((double)d.beta2, (double)epoch);
            const float step_size = (float)((double)d.lr / bc1);
            const float bc2_sqrt = (float)sqrt(bc2);
            for (int i = gtid; i < P; i += gridDim.x * blockDim.x) {
                float gsum = 0.f;
                for (int w = 0; w < d.n_partials; ++w) gsum += d.partials[(size_t)w * d.partial_stride + i];
                if (d.grad_out) d.grad_out[i] = gsum;
                // torch.optim.Adam (single-tensor semantics)
                const float m = d.adam_m[i] = d.adam_m[i] + (gsum - d.adam_m[i]) * d.omb1;     // exp_avg.lerp_(grad, 1 - beta1)
                const float v = d.adam_v[i] = fmaf(d.omb2 * gsum, gsum, d.beta2 * d.adam_v[i]);   // mul_(beta2).addcmul_(g, g, 1 - beta2)
                const float denom = sqrtf(v) / bc2_sqrt + d.eps;
                d.params[i] = d.params[i] - step_size * (m / denom);
            }
        }
        grid.sync();
        if (stop) break;
    }
}

static void *pick_kernel(const PinnDesc &d, int *wmax)
{
    int maxw = 0;
    for (int l = 0; l < d.nlayers; ++l) maxw = d.layers[l] > maxw ? d.layers[l] : maxw;
    if (maxw <= 20) { *wmax = 20; return (void *)pinn_train_kernel<20>; }
    if (maxw <= 40) { *wmax = 40; return (void *)pinn_train_kernel<40>; }
    return nullptr;
}

cudaError_t plan_pinn_train(const PinnDesc &d, int n_sm, int *blocks_out, size_t *smem_out, int *n_partials)
{
    int WMAX = 0;
    void *kern = pick_kernel(d, &WMAX);
    if (!kern) return cudaErrorInvalidValue;
    const int P4 = (d.n_params + 3) & ~3;
    const size_t smem = ((size_t)P4 + (size_t)WARPS * (P4 + 2 * 32 * WMAX + 8)) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    const int chunks = (d.NR + 31) / 32 + (d.NI + 31) / 32 + (d.NB + 31) / 32;
    int blocks = (chunks + WARPS - 1) / WARPS;
    if (blocks > n_sm * per_sm) blocks = n_sm * per_sm;
    if (blocks < 1) blocks = 1;
    *blocks_out = blocks;
    *smem_out = smem;
    *n_partials = blocks * WARPS < chunks ? blocks * WARPS : chunks;
    return cudaSuccess;
}

cudaError_t launch_pinn_train(const PinnDesc &d, int blocks, size_t smem, cudaStream_t st)
{
    int WMAX = 0;
    void *kern = pick_kernel(d, &WMAX);
    if (!kern) return cudaErrorInvalidValue;
    PinnDesc dd = d;
    void *args[] = {&dd};
    return cudaLaunchCooperativeKernel(kern, dim3(blocks), dim3(WARPS * 32), args, smem, st);
}

}  // namespace gp
