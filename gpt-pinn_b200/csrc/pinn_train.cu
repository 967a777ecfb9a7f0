// pinn_train.cu -- fused full-PINN training (forward + derivative residual + backward + Adam) in ONE
// persistent cooperative kernel for all epochs.
//
// Replaces the reference's pinn_train loops (Klein-Gordon/KG_train.py:48-59 with KG_models.py:84-129,
// Burgers/B_train.py:76-90 with B_models.py:70-105): per epoch the reference builds a triple-nested
// autograd graph (grad of grad of grad) with ~100 kernel launches and an optimizer step.  Here one
// epoch is
//   phase 1  a block of 4 warps takes a tile of 32 collocation / IC / BC points.  Thread = (point,
//            neuron group): lane = point, warp = one quarter of every hidden layer's neurons.  The five
//            derivative channels (u, u_x, u_t, q = u_xx+u_xt, r = u_tt+u_xt: the reference's mixed second
//            derivatives) of the thread's own neurons stay in registers for the whole
//            forward + reverse sweep; only the operands every thread of a point needs (the previous
//            layer's activations, or the next layer's adjoints) are staged through shared memory, so each
//            layer is a small GEMM  [neurons] x [neurons] x [5 channels x 32 points]  with weights read
//            as warp-wide broadcasts.  The reverse pass is hand-derived (uses sigma', sigma'', sigma''');
//            weight gradients are outer-product GEMMs over (channel, point) with a register tile per
//            lane, K split over the four warps, combined in shared memory (no atomics, fixed order);
//   grid.sync
//   phase 2  deterministic reduction of the per-block partial gradients and losses, the tolerance test
//            of B_train.py:84 (checked BEFORE the step) and torch.optim.Adam's update (bias correction
//            in double);
//   grid.sync
// so the weights never leave the device and the host launches once per training run.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "pinn_train.cuh"

namespace cg = cooperative_groups;

#ifndef GP_PINN_UNROLL
#define GP_PINN_UNROLL 4               // rows of a layer GEMM in flight together (measured on B200: 1 -> 65.7, 2 -> 64.0, 4 -> 62.5 us per KG epoch)
#endif
#ifndef GP_PINN_MINB
#define GP_PINN_MINB 2                // resident CTAs per SM the register allocation aims at
#endif

namespace gp {

constexpr int CH = 5;                 // v, x, t, q, r
constexpr int NG = 4;                 // neuron groups = warps per block
constexpr int PT = 32;                // points per tile = lanes
constexpr int TJG = 4, TIG = 8;       // lane tiling of a weight-gradient matrix: 4 output groups x 8 input groups

// The transcendental part of the activation, evaluated ONCE per (layer, neuron, point) in the forward sweep and kept in a
// thread-private shared-memory slot for the reverse sweep (the precise sincosf / tanhf with their range reduction were a
// quarter of the kernel's instructions when every use recomputed them): cos -> (cos z, sin z), tanh -> (tanh z, -).
#ifndef GP_PINN_ACT_INLINE
#define GP_PINN_ACT_INLINE 1          // 0: the precise sincosf / tanhf live in ONE out-of-line copy instead of 20 inlined ones
#endif
#if GP_PINN_ACT_INLINE
__device__ __forceinline__
#else
__device__ __noinline__
#endif
void act_eval(int act, float z, float &p0, float &p1)
{
    if (act == 0) { float sn, cs; sincosf(z, &sn, &cs); p0 = cs; p1 = sn; }
    else { p0 = tanhf(z); p1 = 0.f; }
}
// the activation and its first three derivatives from the cached pair
__device__ __forceinline__ void act_derivs(int act, float p0, float p1, float &s0, float &s1, float &s2, float &s3)
{
    if (act == 0) {                   // cos
        s0 = p0; s1 = -p1; s2 = -p0; s3 = p1;
    } else {                          // tanh
        const float s = p0;
        s0 = s; s1 = 1.f - s * s; s2 = -2.f * s * s1; s3 = s1 * (6.f * s * s - 2.f);
    }
}

// five channels of a = sigma(z) from the five channels of z
__device__ __forceinline__ void act_forward(int act, float p0, float p1, const float (&z)[CH], float (&a)[CH])
{
    float s0, s1, s2, s3;
    act_derivs(act, p0, p1, s0, s1, s2, s3);
    a[0] = s0;
    a[1] = s1 * z[1];
    a[2] = s1 * z[2];
    a[3] = fmaf(s2 * z[1], z[1] + z[2], s1 * z[3]);
    a[4] = fmaf(s2 * z[2], z[2] + z[1], s1 * z[4]);
}

// adjoint of the five channels of z given the adjoint of the five channels of a = sigma(z)
__device__ __forceinline__ void act_backward(int act, float p0, float p1, const float (&z)[CH], const float (&ab)[CH], float (&zb)[CH])
{
    float s0, s1, s2, s3;
    act_derivs(act, p0, p1, s0, s1, s2, s3);
    const float zx = z[1], zt = z[2];
    zb[3] = s1 * ab[3];
    zb[4] = s1 * ab[4];
    zb[1] = fmaf(s1, ab[1], s2 * fmaf(ab[3], 2.f * zx + zt, ab[4] * zt));
    zb[2] = fmaf(s1, ab[2], s2 * fmaf(ab[3], zx, ab[4] * (2.f * zt + zx)));
    zb[0] = s1 * ab[0] + s2 * (zx * ab[1] + zt * ab[2] + z[3] * ab[3] + z[4] * ab[4])
          + s3 * (ab[3] * zx * (zx + zt) + ab[4] * zt * (zt + zx));
}

// WMAX: widest hidden layer (all hidden layers are padded to it), NH: hidden layers (KG 2, Burgers 4)
// ACT: the activation (0 cos, 1 tanh) as a compile-time constant -- with a run-time switch every one of the 20 activation
// sites of a tile carried both sincosf and tanhf, a third of the kernel's 140 KB of code (more than the instruction cache)
template <int WMAX, int NH, int ACT>
__global__ void __launch_bounds__(NG * 32, GP_PINN_MINB) pinn_train_kernel(const PinnDesc d, const PinnBatch bt)
{
    // this CTA's training: group `tr` of the grid, local CTA index `lb`; CTAs beyond n * group only keep the barriers
    const int gsz = bt.group;
    const int tr = (int)blockIdx.x / gsz, lb = (int)blockIdx.x % gsz;
    const bool member = tr < bt.n;
    const PinnTrainee &T = bt.t[member ? tr : 0];
    bool done = !member;                                     // this training has met its tolerance (or the CTA is a spare)
    constexpr int GJ = WMAX / NG;                            // neurons per warp (10 / 5)
    constexpr int GJP = (GJ + 3) & ~3;                       // padded to a float4 multiple (12 / 8)
    constexpr int TJ = (WMAX + TJG - 1) / TJG, TI = (WMAX + TIG - 1) / TIG;
    constexpr int NS = CH * PT + 1;                          // neuron stride of a staging buffer (odd: conflict-free column reads)
    constexpr int STG = WMAX * NS;                           // one staging buffer [neuron][channel][point]
    constexpr int NHH = NH > 1 ? NH - 1 : 1;
    constexpr int UNR = GP_PINN_UNROLL;                      // rows of a layer GEMM in flight together
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) float smem[];
    const int P = d.n_params;
    // Shared copies hold everything EXCEPT the hidden->hidden matrices: those live in the padded wt / wp copies (weights) and
    // in the wacc register tiles (gradients, combined through the staging buffer after the tile loop).  This keeps a CTA at
    // ~71 KB for the KG network, so three CTAs fit an SM (361 point tiles of the default KG problem in ONE round of 444).
    int PC = 0;
    for (int l = 0; l <= NH; ++l) PC += ((l == 0 || l == NH) ? d.layers[l] * d.layers[l + 1] : 0) + d.layers[l + 1];
    const int PC4 = (PC + 3) & ~3;
    float *w_s = smem;                                       // [PC]  first / last matrix and all biases, compact (cw / cb offsets)
    float *g_s = w_s + PC4;                                  // [PC]  this block's gradient partial of the same entries
    float *wt_s = g_s + PC4;                                 // [NH-1][WMAX][NG][GJP]  hidden->hidden weights, transposed + padded:  wt[l][i][jg][jj] = W[jg*GJ+jj][i]
    float *wp_s = wt_s + (NH - 1) * WMAX * NG * GJP;         // [NH-1][WMAX][NG][GJP]  same weights, row-major + padded:          wp[l][j][ig][ii] = W[j][ig*GJ+ii]
    float *A_s = wp_s + (NH - 1) * WMAX * NG * GJP;          // [WMAX][CH][PT]
    float *B_s = A_s + STG;                                  // [WMAX][CH][PT]
    float *U_s = B_s + STG;                                  // [NG][CH][PT] partial network outputs
    float *S_s = U_s + NG * CH * PT + 8;                     // [NH][GJ][2][threads] cached (cos, sin) / (tanh, -): every slot is read back by the thread that wrote it
    auto sslot = [&](int l, int jj, int comp) -> float & { return S_s[((l * GJ + jj) * 2 + comp) * (NG * 32) + threadIdx.x]; };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int j0 = warp * GJ;                                // first neuron owned by this thread in every hidden layer
    const int tjg = lane / TIG, tig = lane % TIG;

    int woff[NH + 1], boff[NH + 1];                          // offsets in the packed parameter vector
    int cw[NH + 1], cb[NH + 1];                              // offsets in the compact shared copies (cw of a hidden->hidden map unused)
    {
        int off = 0, coff = 0;
#pragma unroll
        for (int l = 0; l <= NH; ++l) {
            woff[l] = off; off += d.layers[l] * d.layers[l + 1];
            boff[l] = off; off += d.layers[l + 1];
            cw[l] = coff; if (l == 0 || l == NH) coff += d.layers[l] * d.layers[l + 1];
            cb[l] = coff; coff += d.layers[l + 1];
        }
    }
    const int tilesR = (d.NR + PT - 1) / PT, tilesI = (d.NI + PT - 1) / PT, tilesB = (d.NB + PT - 1) / PT;
    const int ntiles = tilesR + tilesI + tilesB;

    for (int epoch = 1; epoch <= d.epochs; ++epoch) {
      if (!done) {
        // -------------------------------------------------------------- phase 1
#pragma unroll
        for (int l = 0; l <= NH; ++l) {
            const int nw = (l == 0 || l == NH) ? d.layers[l] * d.layers[l + 1] : 0, nb = d.layers[l + 1];
            for (int i = threadIdx.x; i < nw; i += blockDim.x) w_s[cw[l] + i] = T.params[woff[l] + i];
            for (int i = threadIdx.x; i < nb; i += blockDim.x) w_s[cb[l] + i] = T.params[boff[l] + i];
        }
        for (int i = threadIdx.x; i < PC; i += blockDim.x) g_s[i] = 0.f;
        for (int l = 1; l < NH; ++l) {                       // padded copies of the hidden->hidden matrices (from global / L2)
            const float *W = T.params + woff[l];
            const int wi = d.layers[l], wo = d.layers[l + 1];
            float *wt = wt_s + (l - 1) * WMAX * NG * GJP, *wp = wp_s + (l - 1) * WMAX * NG * GJP;
            for (int e = threadIdx.x; e < WMAX * NG * GJP; e += blockDim.x) {
                const int a = e / (NG * GJP), g = (e / GJP) % NG, s = e % GJP;
                const int b = g * GJ + s;
                wt[e] = (s < GJ && b < wo && a < wi) ? W[b * wi + a] : 0.f;      // wt[i=a][g][s] = W[j=b][i=a]
                wp[e] = (s < GJ && a < wo && b < wi) ? W[a * wi + b] : 0.f;      // wp[j=a][g][s] = W[j=a][i=b]
            }
        }
        __syncthreads();
        float lossR = 0.f, lossI1 = 0.f, lossI2 = 0.f, lossB = 0.f;          // accumulated by warp 0 only
        float wacc[NHH][TJ][TI];                                            // weight-gradient tiles of the hidden->hidden layers
#pragma unroll
        for (int l = 0; l < NHH; ++l)
#pragma unroll
            for (int a = 0; a < TJ; ++a)
#pragma unroll
                for (int b = 0; b < TI; ++b) wacc[l][a][b] = 0.f;

        for (int tile = lb; tile < ntiles; tile += gsz) {
            int kind, p;                              // 0 residual, 1 IC, 2 BC
            if (tile < tilesR) { kind = 0; p = tile * PT + lane; }
            else if (tile < tilesR + tilesI) { kind = 1; p = (tile - tilesR) * PT + lane; }
            else { kind = 2; p = (tile - tilesR - tilesI) * PT + lane; }
            const int np = kind == 0 ? d.NR : (kind == 1 ? d.NI : d.NB);
            const bool live = p < np;
            const float *xt = kind == 0 ? d.xt_resid : (kind == 1 ? d.IC_xt : d.BC_xt);
            const float x = live ? xt[2 * p] : 0.f, t = live ? xt[2 * p + 1] : 0.f;

            // ================================ forward ================================
            float z[NH][CH][GJ];                      // pre-activations of this thread's neurons, every hidden layer
            {
                const float *W = w_s + cw[0], *Bv = w_s + cb[0];       // layer 0: inputs (x, t), channels known in closed form
#pragma unroll
                for (int jj = 0; jj < GJ; ++jj) {
                    const int j = j0 + jj;
                    const bool ok = j < d.layers[1];
                    const float wx = ok ? W[j * 2] : 0.f, wt0 = ok ? W[j * 2 + 1] : 0.f;
                    z[0][0][jj] = ok ? fmaf(wx, x, fmaf(wt0, t, Bv[j])) : 0.f;
                    z[0][1][jj] = wx; z[0][2][jj] = wt0; z[0][3][jj] = 0.f; z[0][4][jj] = 0.f;
                }
            }
#pragma unroll
            for (int l = 0; l < NH; ++l) {
                // activations of hidden layer l for the thread's own neurons -> staging buffer A
#pragma unroll
                for (int jj = 0; jj < GJ; ++jj) {
                    float zz[CH], aa[CH];
#pragma unroll
                    for (int c = 0; c < CH; ++c) zz[c] = z[l][c][jj];
                    float p0, p1;
                    act_eval(ACT, zz[0], p0, p1);
                    sslot(l, jj, 0) = p0; sslot(l, jj, 1) = p1;
                    act_forward(ACT, p0, p1, zz, aa);
                    const bool ok = (j0 + jj) < d.layers[l + 1];
#pragma unroll
                    for (int c = 0; c < CH; ++c) A_s[(j0 + jj) * NS + c * PT + lane] = ok ? aa[c] : 0.f;
                }
                __syncthreads();
                if (l < NH - 1) {
                    // z_{l+1}[c][j] = b[j] + sum_i W[j][i] a_l[c][i]   (weights as warp-wide broadcasts)
                    const float *wt = wt_s + l * WMAX * NG * GJP + warp * GJP;
                    const float *Bv = w_s + cb[l + 1];
#pragma unroll
                    for (int jj = 0; jj < GJ; ++jj) {
                        z[l + 1][0][jj] = (j0 + jj) < d.layers[l + 2] ? Bv[j0 + jj] : 0.f;
#pragma unroll
                        for (int c = 1; c < CH; ++c) z[l + 1][c][jj] = 0.f;
                    }
                    // all WMAX rows: rows past the layer width are zero in A and in wt (compile-time trip count: the loads
                    // of the next rows are issued under the FMAs of the current ones)
#pragma unroll UNR
                    for (int i = 0; i < WMAX; ++i) {
                        float av[CH], wv[GJP];
#pragma unroll
                        for (int c = 0; c < CH; ++c) av[c] = A_s[i * NS + c * PT + lane];
#pragma unroll
                        for (int s = 0; s < GJP; s += 4) {
                            const float4 w4 = *reinterpret_cast<const float4 *>(wt + i * NG * GJP + s);
                            wv[s] = w4.x; wv[s + 1] = w4.y; wv[s + 2] = w4.z; wv[s + 3] = w4.w;
                        }
#pragma unroll
                        for (int jj = 0; jj < GJ; ++jj)
#pragma unroll
                            for (int c = 0; c < CH; ++c) z[l + 1][c][jj] = fmaf(wv[jj], av[c], z[l + 1][c][jj]);
                    }
                    __syncthreads();                  // A is rewritten by the next layer
                }
            }
            // output layer (one linear unit): partial over the thread's neurons, then over the four warps
            float u[CH];
            {
                const float *W = w_s + cw[NH];
                float pu[CH] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int jj = 0; jj < GJ; ++jj) {
                    const float wv = (j0 + jj) < d.layers[NH] ? W[j0 + jj] : 0.f;
#pragma unroll
                    for (int c = 0; c < CH; ++c) pu[c] = fmaf(wv, A_s[(j0 + jj) * NS + c * PT + lane], pu[c]);
                }
#pragma unroll
                for (int c = 0; c < CH; ++c) U_s[(warp * CH + c) * PT + lane] = pu[c];
                __syncthreads();
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    float s = c == 0 ? w_s[cb[NH]] : 0.f;
#pragma unroll
                    for (int g = 0; g < NG; ++g) s += U_s[(g * CH + c) * PT + lane];
                    u[c] = s;
                }
            }

            // ================================ loss terms and output adjoints ================================
            float ub[CH] = {0.f, 0.f, 0.f, 0.f, 0.f};
            if (live) {
                if (kind == 0) {
                    const float fh = d.f_hat ? d.f_hat[p] : 0.f;
                    float f;
                    if (d.pde == 0) f = u[4] + T.pa * u[3] + T.pb * u[0] + T.pc * u[0] * u[0] + d.xcos[p];   // KG_models.py:104-107
                    else f = u[2] + (u[0] * u[1] + T.pa * u[3]);                                          // B_models.py:91-94, pa = -nu
                    const float e = f - fh;
                    lossR = fmaf(e, e, lossR);
                    const float eb = 2.f * e * d.invR;
                    if (d.pde == 0) { ub[4] = eb; ub[3] = T.pa * eb; ub[0] = (T.pb + 2.f * T.pc * u[0]) * eb; }
                    else { ub[2] = eb; ub[0] = u[1] * eb; ub[1] = u[0] * eb; ub[3] = T.pa * eb; }
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

            // ================================ reverse ================================
            // output layer: Wbar_out[j] += sum_{c,p} ub_c a_{NH-1}[c][j];  bbar_out += sum_p ub_v;  abar[c][j] = W_out[j] ub_c
            float ab[CH][GJ];                         // adjoint of the activations, then of the pre-activations (in place)
            {
                const float *W = w_s + cw[NH];
#pragma unroll
                for (int jj = 0; jj < GJ; ++jj) {
                    const bool ok = (j0 + jj) < d.layers[NH];
                    float s = 0.f;
#pragma unroll
                    for (int c = 0; c < CH; ++c) s = fmaf(ub[c], A_s[(j0 + jj) * NS + c * PT + lane], s);
                    for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                    if (lane == 0 && ok) g_s[cw[NH] + j0 + jj] += s;
                    const float wv = ok ? W[j0 + jj] : 0.f;
#pragma unroll
                    for (int c = 0; c < CH; ++c) ab[c][jj] = wv * ub[c];
                }
                if (warp == 0) {
                    float s = ub[0];
                    for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                    if (lane == 0) g_s[cb[NH]] += s;
                }
            }
            __syncthreads();                          // all reads of A (activations of the last hidden layer) are done
#pragma unroll
            for (int l = NH - 1; l >= 0; --l) {
                // through the activation of hidden layer l: ab <- zbar
#pragma unroll
                for (int jj = 0; jj < GJ; ++jj) {
                    float zz[CH], aa[CH], zb[CH];
#pragma unroll
                    for (int c = 0; c < CH; ++c) { zz[c] = z[l][c][jj]; aa[c] = ab[c][jj]; }
                    act_backward(ACT, sslot(l, jj, 0), sslot(l, jj, 1), zz, aa, zb);
#pragma unroll
                    for (int c = 0; c < CH; ++c) ab[c][jj] = zb[c];
                }
                // bias gradient of the linear map feeding hidden layer l: bbar[j] += sum_p zbar_v[j]
#pragma unroll
                for (int jj = 0; jj < GJ; ++jj) {
                    float s = ab[0][jj];
                    for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                    if (lane == 0 && (j0 + jj) < d.layers[l + 1]) g_s[cb[l] + j0 + jj] += s;
                }
                if (l == 0) {
                    // first linear map: inputs (x, t) with channels (x,1,0,0,0) and (t,0,1,0,0)
#pragma unroll
                    for (int jj = 0; jj < GJ; ++jj) {
                        float sx = fmaf(ab[0][jj], x, ab[1][jj]), st = fmaf(ab[0][jj], t, ab[2][jj]);
                        for (int off = 16; off; off >>= 1) {
                            sx += __shfl_xor_sync(0xffffffffu, sx, off);
                            st += __shfl_xor_sync(0xffffffffu, st, off);
                        }
                        if (lane == 0 && (j0 + jj) < d.layers[1]) {
                            g_s[cw[0] + (j0 + jj) * 2] += sx;
                            g_s[cw[0] + (j0 + jj) * 2 + 1] += st;
                        }
                    }
                } else {
                    // stage zbar (B) and the activations of hidden layer l-1 (A) for the two GEMMs
#pragma unroll
                    for (int jj = 0; jj < GJ; ++jj) {
                        float zz[CH], aa[CH];
#pragma unroll
                        for (int c = 0; c < CH; ++c) zz[c] = z[l - 1][c][jj];
                        act_forward(ACT, sslot(l > 0 ? l - 1 : 0, jj, 0), sslot(l > 0 ? l - 1 : 0, jj, 1), zz, aa);
                        const bool okA = (j0 + jj) < d.layers[l], okB = (j0 + jj) < d.layers[l + 1];
#pragma unroll
                        for (int c = 0; c < CH; ++c) {
                            A_s[(j0 + jj) * NS + c * PT + lane] = okA ? aa[c] : 0.f;
                            B_s[(j0 + jj) * NS + c * PT + lane] = okB ? ab[c][jj] : 0.f;
                        }
                    }
                    __syncthreads();
                    // weight gradient of map l (hidden l-1 -> hidden l): Wbar[j][i] += sum_{c,p} zbar[c][j][p] a[c][i][p];
                    // lane (tjg, tig) owns a TJ x TI tile, warp w takes points 8w .. 8w+7 of every channel
                    for (int c = 0; c < CH; ++c) {
#pragma unroll
                        for (int q8 = 0; q8 < PT / NG; ++q8) {
                            const int q = warp * (PT / NG) + q8;
                            float zv[TJ], av[TI];
#pragma unroll
                            for (int a = 0; a < TJ; ++a) { const int j = tjg * TJ + a; zv[a] = j < WMAX ? B_s[j * NS + c * PT + q] : 0.f; }
#pragma unroll
                            for (int b = 0; b < TI; ++b) { const int i = tig * TI + b; av[b] = i < WMAX ? A_s[i * NS + c * PT + q] : 0.f; }
#pragma unroll
                            for (int a = 0; a < TJ; ++a)
#pragma unroll
                                for (int b = 0; b < TI; ++b) wacc[l > 0 ? l - 1 : 0][a][b] = fmaf(zv[a], av[b], wacc[l > 0 ? l - 1 : 0][a][b]);
                        }
                    }
                    // adjoint of the activations of hidden layer l-1: abar[c][i] = sum_j W[j][i] zbar[c][j]
                    const float *wp = wp_s + (l > 0 ? l - 1 : 0) * WMAX * NG * GJP + warp * GJP;
#pragma unroll
                    for (int jj = 0; jj < GJ; ++jj)
#pragma unroll
                        for (int c = 0; c < CH; ++c) ab[c][jj] = 0.f;
#pragma unroll UNR
                    for (int j = 0; j < WMAX; ++j) {             // rows past the layer width are zero in B and in wp
                        float zv[CH], wv[GJP];
#pragma unroll
                        for (int c = 0; c < CH; ++c) zv[c] = B_s[j * NS + c * PT + lane];
#pragma unroll
                        for (int s = 0; s < GJP; s += 4) {
                            const float4 w4 = *reinterpret_cast<const float4 *>(wp + j * NG * GJP + s);
                            wv[s] = w4.x; wv[s + 1] = w4.y; wv[s + 2] = w4.z; wv[s + 3] = w4.w;
                        }
#pragma unroll
                        for (int jj = 0; jj < GJ; ++jj)
#pragma unroll
                            for (int c = 0; c < CH; ++c) ab[c][jj] = fmaf(wv[jj], zv[c], ab[c][jj]);
                    }
                    __syncthreads();                  // A / B are rewritten by the next (lower) layer
                }
            }
        }   // tiles

        // ---- combine the four warps' weight-gradient tiles (K split), fixed order, in the staging buffer A (free after
        //      the tile loop), layer l at A_s[l * WMAX * WMAX + j * wi + i]
        __syncthreads();
        for (int i = threadIdx.x; i < (NH - 1) * WMAX * WMAX; i += blockDim.x) A_s[i] = 0.f;
        __syncthreads();
        for (int w = 0; w < NG; ++w) {
            if (warp == w) {
#pragma unroll
                for (int l = 0; l < NH - 1; ++l) {
                    const int wi = d.layers[l + 1], wo = d.layers[l + 2];
#pragma unroll
                    for (int a = 0; a < TJ; ++a)
#pragma unroll
                        for (int b = 0; b < TI; ++b) {
                            const int j = tjg * TJ + a, i = tig * TI + b;
                            if (j < wo && i < wi) A_s[l * WMAX * WMAX + j * wi + i] += wacc[l][a][b];
                        }
                }
            }
            __syncthreads();
        }
        if (warp == 0) {
            for (int off = 16; off; off >>= 1) {
                lossR += __shfl_xor_sync(0xffffffffu, lossR, off);
                lossI1 += __shfl_xor_sync(0xffffffffu, lossI1, off);
                lossI2 += __shfl_xor_sync(0xffffffffu, lossI2, off);
                lossB += __shfl_xor_sync(0xffffffffu, lossB, off);
            }
        }
        {
            float *dst = T.partials + (size_t)lb * d.partial_stride;
#pragma unroll
            for (int l = 0; l <= NH; ++l) {
                const int nw = d.layers[l] * d.layers[l + 1], nb = d.layers[l + 1];
                if (l == 0 || l == NH) { for (int i = threadIdx.x; i < nw; i += blockDim.x) dst[woff[l] + i] = g_s[cw[l] + i]; }
                else { for (int i = threadIdx.x; i < nw; i += blockDim.x) dst[woff[l] + i] = A_s[(l - 1) * WMAX * WMAX + i]; }
                for (int i = threadIdx.x; i < nb; i += blockDim.x) dst[boff[l] + i] = g_s[cb[l] + i];
            }
            if (threadIdx.x == 0) { dst[P] = lossR; dst[P + 1] = lossI1; dst[P + 2] = lossI2; dst[P + 3] = lossB; }
        }
      }   // !done
        grid.sync();
      if (!done) {

        // -------------------------------------------------------------- phase 2: reduce, tolerance, Adam
        // every block reduces the four loss sums itself (same order everywhere => the same stop decision
        // on every block, no extra grid barrier)
        if (warp == 0) {
            float s[4] = {0.f, 0.f, 0.f, 0.f};
            for (int w = lane; w < gsz; w += 32)
#pragma unroll
                for (int k = 0; k < 4; ++k) s[k] += T.partials[(size_t)w * d.partial_stride + P + k];
            for (int off = 16; off; off >>= 1)
#pragma unroll
                for (int k = 0; k < 4; ++k) s[k] += __shfl_xor_sync(0xffffffffu, s[k], off);
            if (lane == 0) {
                float loss = s[0] * d.invR + s[1] * d.invI;
                if (d.pde == 0) loss += s[2] * d.invI;
                loss += s[3] * d.invB;
                U_s[0] = loss;
                if (lb == 0) {
                    T.status[0] = loss;                                   // loss of the weights BEFORE this epoch's step
                    T.status[1] = __int_as_float(epoch);
                    if (T.loss_trace && d.trace_every > 0 && (epoch - 1) % d.trace_every == 0) T.loss_trace[(epoch - 1) / d.trace_every] = loss;
                }
            }
        }
        __syncthreads();
        const bool stop = d.tol >= 0.f && U_s[0] < d.tol;                 // B_train.py:84: break before the step
        if (!stop) {
            const double bc1 = 1.0 - pow(d.beta1_d, (double)epoch);       // torch.optim.Adam: bias corrections in double from the
            const double bc2 = 1.0 - pow(d.beta2_d, (double)epoch);       // Python floats (0.9, 0.999), not from their fp32 roundings
            const float step_size = (float)(d.lr_d / bc1);
            const float bc2_sqrt = (float)sqrt(bc2);
            // gradient reduction: 8 lanes per parameter, fixed order (deterministic), then torch.optim.Adam
            const int gthreads = gsz * blockDim.x;
            const int sub = threadIdx.x & 7;
            const int gid = (lb * blockDim.x + threadIdx.x) >> 3, ngroups = gthreads >> 3;
            for (int i0 = 0; i0 < P; i0 += ngroups) {                     // same trip count on every thread (shuffles below)
                const int i = i0 + gid;
                float gs = 0.f;
                if (i < P)
                    for (int w = sub; w < gsz; w += 8) gs += T.partials[(size_t)w * d.partial_stride + i];
                gs += __shfl_xor_sync(0xffffffffu, gs, 1);
                gs += __shfl_xor_sync(0xffffffffu, gs, 2);
                gs += __shfl_xor_sync(0xffffffffu, gs, 4);
                if (sub == 0 && i < P) {
                    if (T.grad_out) T.grad_out[i] = gs;
                    const float m = T.adam_m[i] = T.adam_m[i] + (gs - T.adam_m[i]) * d.omb1;          // exp_avg.lerp_(grad, 1 - beta1)
                    const float v = T.adam_v[i] = fmaf(d.omb2 * gs, gs, d.beta2 * T.adam_v[i]);        // mul_(beta2).addcmul_(g, g, 1 - beta2)
                    const float denom = sqrtf(v) / bc2_sqrt + d.eps;
                    T.params[i] = T.params[i] - step_size * (m / denom);
                }
            }
        }
        if (stop) {                                                       // this training is finished; the launch ends when all are
            done = true;
            if (lb == 0 && threadIdx.x == 0) atomicAdd(d.stop, 1);
        }
      }   // !done
        grid.sync();
        if (*reinterpret_cast<volatile int *>(d.stop) >= bt.n) break;     // same value on every CTA: all increments precede the barrier
    }
}

struct KernelPick { void *fn; int wmax, nh; };

template <int ACT>
static KernelPick pick_kernel_act(int maxw, int nh)
{
    if (maxw <= 20) {
        if (nh == 1) return {(void *)pinn_train_kernel<20, 1, ACT>, 20, 1};
        if (nh == 2) return {(void *)pinn_train_kernel<20, 2, ACT>, 20, 2};
        if (nh == 3) return {(void *)pinn_train_kernel<20, 3, ACT>, 20, 3};
        if (nh == 4) return {(void *)pinn_train_kernel<20, 4, ACT>, 20, 4};
    } else if (maxw <= 40) {
        if (nh == 1) return {(void *)pinn_train_kernel<40, 1, ACT>, 40, 1};
        if (nh == 2) return {(void *)pinn_train_kernel<40, 2, ACT>, 40, 2};
        if (nh == 3) return {(void *)pinn_train_kernel<40, 3, ACT>, 40, 3};
    }
    return {nullptr, 0, 0};
}

static KernelPick pick_kernel(const PinnDesc &d)
{
    const int nh = d.nlayers - 2;
    int maxw = 0;
    for (int l = 1; l < d.nlayers - 1; ++l) maxw = d.layers[l] > maxw ? d.layers[l] : maxw;
    return d.act == 0 ? pick_kernel_act<0>(maxw, nh) : pick_kernel_act<1>(maxw, nh);
}

static size_t smem_bytes(const PinnDesc &d, const KernelPick &k)
{
    int PC = 0;                                              // first / last matrix + all biases (the kernel's compact copies)
    for (int l = 0; l + 1 < d.nlayers; ++l) PC += ((l == 0 || l == d.nlayers - 2) ? d.layers[l] * d.layers[l + 1] : 0) + d.layers[l + 1];
    const int P4 = (PC + 3) & ~3;
    const int GJ = k.wmax / NG, GJP = (GJ + 3) & ~3;
    const size_t fl = (size_t)2 * P4 + (size_t)2 * (k.nh - 1) * k.wmax * NG * GJP + (size_t)2 * k.wmax * (CH * PT + 1) + (size_t)NG * CH * PT + 8
                      + (size_t)k.nh * GJ * 2 * NG * 32;                          // cached activation pairs
    return fl * sizeof(float);
}

cudaError_t plan_pinn_train(const PinnDesc &d, int n_sm, int nbatch, int *blocks_out, size_t *smem_out, int *group_out)
{
    if (nbatch < 1 || nbatch > PINN_MAX_BATCH) return cudaErrorInvalidValue;
    const KernelPick k = pick_kernel(d);
    if (!k.fn) return cudaErrorInvalidValue;
    const size_t smem = smem_bytes(d, k);
    cudaError_t e = cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k.fn, NG * 32, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    const int tiles = (d.NR + PT - 1) / PT + (d.NI + PT - 1) / PT + (d.NB + PT - 1) / PT;
    int group = (n_sm * per_sm) / nbatch;                 // CTAs per training: an equal share of the resident CTAs,
    if (group > tiles) group = tiles;                     // never more than one CTA per tile,
    if (group < 1) return cudaErrorInvalidConfiguration;
    const int rounds = (tiles + group - 1) / group;       // and no more than the tile rounds need
    group = (tiles + rounds - 1) / rounds;
    if (const char *ev = getenv("GPTPINN_PINN_GROUP")) {  // experiments: CTAs per training
        const int g = atoi(ev);
        if (g >= 1 && g <= (n_sm * per_sm) / nbatch && g <= tiles) group = g;
    }
    *blocks_out = group * nbatch;
    *smem_out = smem;
    *group_out = group;
    return cudaSuccess;
}

cudaError_t launch_pinn_train(const PinnDesc &d, const PinnBatch &b, int blocks, size_t smem, cudaStream_t st)
{
    const KernelPick k = pick_kernel(d);
    if (!k.fn) return cudaErrorInvalidValue;
    PinnDesc dd = d;
    PinnBatch bb = b;
    void *args[] = {&dd, &bb};
    return cudaLaunchCooperativeKernel(k.fn, dim3(blocks), dim3(NG * 32), args, smem, st);
}

}  // namespace gp
