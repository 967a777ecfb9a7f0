// sweep_kernel.cuh -- the fused, persistent coefficient-sweep kernel (sm_100a, FP32 CUDA cores).
//
// One launch runs ALL epochs of the reduced-model gradient descent for every parameter of the
// grid (reference: the per-parameter body of offline_generation, Klein-Gordon/KG_train.py:14-37,
// i.e. KG_optimizer.py:34-68 + KG_models.py:28-50 per epoch; Burgers / Allen-Cahn likewise).
//
// Mapping
//   warp  = one work item = up to SL*M parameters that share the same linear operator
//           T (KG: same (alpha,beta); their gammas differ).  Lanes = SL sibling-lanes x RL
//           row-slices (SL*RL = 32); each thread owns M parameters: c[M][N], grad[M][N] and
//           the loss sums live in registers for the whole run.
//   CTA   = W warps that consume the same tile stream: a 4-deep shared-memory ring filled by
//           TMA bulk copies (cp.async.bulk + mbarrier complete_tx), full/empty mbarriers.
//   T     = per-warp [NQT][R_TILE][4] block in shared memory, rebuilt from the staged
//           P-derivative tiles for every residual tile (2 FMA/element amortised over the
//           siblings) so the inner loop spends the reference's 4n FMA per row and parameter.
//   epoch = one pass over the tile stream (all residual, IC, BC tiles); at its end the RL
//           row-slices are summed with xor-shuffles (fixed order => bitwise deterministic),
//           the loss of the *current* c falls out of the same pass, then c <- c - lr*grad.
#pragma once
#include "common.cuh"

namespace gp {

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }

template <int NQX>
__device__ __forceinline__ void load_row(const float *blk, int r, float (&v)[4 * NQX])
{
#pragma unroll
    for (int q = 0; q < NQX; ++q) {
        const float4 x = ld4(blk + q * QSTRIDE + r * 4);
        v[4 * q + 0] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
    }
}

template <int N>
__device__ __forceinline__ float dotn(const float *a, const float (&c)[N], float init)
{
    float s = init;
#pragma unroll
    for (int k = 0; k < N; ++k) s = fmaf(a[k], c[k], s);
    return s;
}

__device__ __forceinline__ void set_comp(float4 &t, int i, float v)
{
    if (i == 0) t.x = v; else if (i == 1) t.y = v; else if (i == 2) t.z = v; else t.w = v;
}

template <int PDE, int N, int M, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 1) sweep_kernel(const SweepParams p)
{
    using TR = PdeTraits<PDE>;
    constexpr int NQ = (N + 3) / 4;          // column quads of a P block that hold data
    constexpr int NQT = N / 4 + 1;           // quads of T: N columns + the forcing column
    constexpr int SLOT = SlotFloats<PDE>::value;
    constexpr int TT = NQT * QSTRIDE;
    constexpr int VEC0 = TR::NARR * ARR_FLTS, VEC1 = VEC0 + R_TILE;
    constexpr unsigned FULLM = 0xffffffffu;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *stage = reinterpret_cast<float *>(smem_raw);
    float *Tall = stage + NSTAGE * SLOT;
    uint64_t *full = reinterpret_cast<uint64_t *>(Tall + p.W * TT);
    uint64_t *empty = full + NSTAGE;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *Tw = Tall + warp * TT;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], p.W); }
        mbar_fence_init();
    }
    __syncthreads();

    const uint32_t tiles_pp = (uint32_t)p.tiles_per_pass;
    const uint32_t total = (uint32_t)p.rounds * (uint32_t)(p.epochs + 1) * tiles_pp;
    const bool producer = (threadIdx.x == 0);
    uint32_t gt = 0;                         // tiles consumed so far by this warp

    auto issue = [&](uint32_t lt) {
        const uint32_t s = lt % NSTAGE, use = lt / NSTAGE;
        if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
        mbar_arrive_expect_tx(&full[s], SLOT * 4);
        bulk_g2s(stage + s * SLOT, p.packed + (size_t)(lt % tiles_pp) * SLOT, SLOT * 4, &full[s]);
    };
    if (producer)
        for (uint32_t lt = 0; lt < (uint32_t)LOOKAHEAD && lt < total; ++lt) issue(lt);

    auto acquire = [&]() -> const float * {
        if (producer) { const uint32_t lt = gt + LOOKAHEAD; if (lt < total) issue(lt); }
        __syncwarp();
        const uint32_t s = gt % NSTAGE;
        mbar_wait(&full[s], (gt / NSTAGE) & 1);
        return stage + s * SLOT;
    };
    auto release = [&]() {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[gt % NSTAGE]);
        ++gt;
    };

    const float fac1 = p.fac1, fac2 = p.fac2, fac3 = p.fac3;
    const bool has_fhat = p.has_fhat != 0;

    for (int round = 0; round < p.rounds; ++round) {
        const int item = round * (int)gridDim.x * p.W + warp * (int)gridDim.x + (int)blockIdx.x;
        const bool active = item < p.n_items;
        ItemDesc it;
        it.tp0 = 0.f; it.tp1 = 0.f; it.sib_start = 0; it.sib_count = 0; it.SL = 32;
        if (active) it = p.items[item];
        const int SL = it.SL, RL = 32 / SL;
        const int sl = lane & (SL - 1), slice = lane / SL;

        float c[M][N], g[M][N];
        float sp0[M], sp1[M], mn[M], lastloss[M];
        int oidx[M];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const int s = sl * M + m;
            const bool valid = active && s < it.sib_count;
            sp0[m] = 0.f; sp1[m] = 0.f; oidx[m] = -1;
            if (valid) {
                const SibDesc sd = p.sibs[it.sib_start + s];
                sp0[m] = sd.sp0; sp1[m] = sd.sp1; oidx[m] = sd.out_idx;
            }
#pragma unroll
            for (int k = 0; k < N; ++k)
                c[m][k] = valid ? (p.c0_per_param ? p.c0[(size_t)oidx[m] * N + k] : p.c0[k]) : 0.f;
            mn[m] = __int_as_float(0x7f800000);
            lastloss[m] = 0.f;
        }

        for (int j = 0; j <= p.epochs; ++j) {
            float lA[M], lB[M], lC[M], lD[M];      // per-segment squared-residual sums
#pragma unroll
            for (int m = 0; m < M; ++m) {
                lA[m] = lB[m] = lC[m] = lD[m] = 0.f;
#pragma unroll
                for (int k = 0; k < N; ++k) g[m][k] = 0.f;
            }

            // ------------------------------ segment 0: PDE residual rows ------------------
            for (int t = 0; t < p.ntile[0]; ++t) {
                const float *slot = acquire();
                if (active) {
                    // ---- per-warp T block (bit-identical to the reference's T tensor) -----
                    for (int e = lane; e < NQT * R_TILE; e += 32) {
                        const int q = e / R_TILE, r = e % R_TILE;
                        const int o = q * QSTRIDE + r * 4;
                        float4 tv;
                        if constexpr (PDE == PDE_KG) {          // (Ptt + a*Pxx) + b*P   KG_precomp.py:23-26
                            const float4 a = ld4(slot + o), b = ld4(slot + ARR_FLTS + o), d = ld4(slot + 2 * ARR_FLTS + o);
                            tv.x = __fadd_rn(__fadd_rn(a.x, __fmul_rn(it.tp0, b.x)), __fmul_rn(it.tp1, d.x));
                            tv.y = __fadd_rn(__fadd_rn(a.y, __fmul_rn(it.tp0, b.y)), __fmul_rn(it.tp1, d.y));
                            tv.z = __fadd_rn(__fadd_rn(a.z, __fmul_rn(it.tp0, b.z)), __fmul_rn(it.tp1, d.z));
                            tv.w = __fadd_rn(__fadd_rn(a.w, __fmul_rn(it.tp0, b.w)), __fmul_rn(it.tp1, d.w));
                        } else if constexpr (PDE == PDE_B) {    // (-nu)*Pxx + Pt        B_precomp.py:18-20
                            const float4 a = ld4(slot + o), b = ld4(slot + ARR_FLTS + o);
                            tv.x = __fadd_rn(__fmul_rn(it.tp0, b.x), a.x);
                            tv.y = __fadd_rn(__fmul_rn(it.tp0, b.y), a.y);
                            tv.z = __fadd_rn(__fmul_rn(it.tp0, b.z), a.z);
                            tv.w = __fadd_rn(__fmul_rn(it.tp0, b.w), a.w);
                        } else {                                // (Pt + (-l)*Pxx) + (-e)*P  AC_precompute.py:20-24
                            const float4 a = ld4(slot + o), b = ld4(slot + ARR_FLTS + o), d = ld4(slot + 2 * ARR_FLTS + o);
                            tv.x = __fadd_rn(__fadd_rn(a.x, __fmul_rn(it.tp0, b.x)), __fmul_rn(it.tp1, d.x));
                            tv.y = __fadd_rn(__fadd_rn(a.y, __fmul_rn(it.tp0, b.y)), __fmul_rn(it.tp1, d.y));
                            tv.z = __fadd_rn(__fadd_rn(a.z, __fmul_rn(it.tp0, b.z)), __fmul_rn(it.tp1, d.z));
                            tv.w = __fadd_rn(__fadd_rn(a.w, __fmul_rn(it.tp0, b.w)), __fmul_rn(it.tp1, d.w));
                        }
                        if (q == N / 4) {       // forcing column: (xcos) - fhat, multiplied by an implicit c = 1
                            float fc = 0.f;
                            if constexpr (PDE == PDE_KG) {
                                fc = slot[VEC0 + r];
                                if (has_fhat) fc = __fsub_rn(fc, slot[VEC1 + r]);
                            } else {
                                if (has_fhat) fc = -slot[VEC0 + r];
                            }
                            set_comp(tv, N % 4, fc);
                        }
                        st4(Tw + o, tv);
                    }
                    __syncwarp();

                    for (int r = slice; r < R_TILE; r += RL) {
                        float Tv[4 * NQT], Pv[4 * NQ];
                        load_row<NQT>(Tw, r, Tv);
                        float fh = 0.f;
                        if constexpr (PDE == PDE_KG) {
                            load_row<NQ>(slot + 2 * ARR_FLTS, r, Pv);
                            if (has_fhat) fh = slot[VEC1 + r];
#pragma unroll
                            for (int m = 0; m < M; ++m) {
                                const float u = dotn<N>(Pv, c[m], 0.f);
                                const float t1 = dotn<N>(Tv, c[m], Tv[N]);
                                const float d = fmaf(sp0[m], u * u, t1);       // residual - fhat
                                lA[m] = fmaf(d, d, lA[m]);
                                float first = d;
                                if (has_fhat) first = d + fh;                    // update ignores fhat (N3)
                                const float w2 = first * (sp1[m] * u);
#pragma unroll
                                for (int k = 0; k < N; ++k) {
                                    g[m][k] = fmaf(first, Tv[k], g[m][k]);
                                    g[m][k] = fmaf(w2, Pv[k], g[m][k]);
                                }
                            }
                        } else if constexpr (PDE == PDE_B) {
                            float Xv[4 * NQ];
                            load_row<NQ>(slot + 2 * ARR_FLTS, r, Xv);
                            load_row<NQ>(slot + 3 * ARR_FLTS, r, Pv);
                            if (has_fhat) fh = slot[VEC0 + r];
#pragma unroll
                            for (int m = 0; m < M; ++m) {
                                const float u = dotn<N>(Pv, c[m], 0.f);
                                const float ux = dotn<N>(Xv, c[m], 0.f);
                                const float t1 = dotn<N>(Tv, c[m], Tv[N]);
                                const float d = fmaf(u, ux, t1);
                                lA[m] = fmaf(d, d, lA[m]);
                                float first = d;
                                if (has_fhat) first = d + fh;
                                const float wx = first * u, wp = first * ux;
#pragma unroll
                                for (int k = 0; k < N; ++k) {
                                    g[m][k] = fmaf(first, Tv[k], g[m][k]);
                                    g[m][k] = fmaf(wx, Xv[k], g[m][k]);
                                    g[m][k] = fmaf(wp, Pv[k], g[m][k]);
                                }
                            }
                        } else {
                            load_row<NQ>(slot + 2 * ARR_FLTS, r, Pv);
                            if (has_fhat) fh = slot[VEC0 + r];
#pragma unroll
                            for (int m = 0; m < M; ++m) {
                                const float u = dotn<N>(Pv, c[m], 0.f);
                                const float t1 = dotn<N>(Tv, c[m], Tv[N]);
                                const float u2 = u * u;
                                const float d = fmaf(sp0[m], u2 * u, t1);
                                lA[m] = fmaf(d, d, lA[m]);
                                float first = d;
                                if (has_fhat) first = d + fh;
                                const float w2 = first * (sp1[m] * u2);
#pragma unroll
                                for (int k = 0; k < N; ++k) {
                                    g[m][k] = fmaf(first, Tv[k], g[m][k]);
                                    g[m][k] = fmaf(w2, Pv[k], g[m][k]);
                                }
                            }
                        }
                    }
                }
                release();
            }
#pragma unroll
            for (int m = 0; m < M; ++m)
#pragma unroll
                for (int k = 0; k < N; ++k) g[m][k] *= fac1;

            // ------------------------------ segment 1: initial-condition rows -------------
            for (int t = 0; t < p.ntile[1]; ++t) {
                const float *slot = acquire();
                if (active) {
                    for (int r = slice; r < R_TILE; r += RL) {
                        float Av[4 * NQ];
                        load_row<NQ>(slot, r, Av);
                        const float tg = slot[VEC0 + r];
                        if constexpr (PDE == PDE_KG) {
                            float Bv[4 * NQ];
                            load_row<NQ>(slot + ARR_FLTS, r, Bv);
                            const float tg2 = slot[VEC1 + r];
#pragma unroll
                            for (int m = 0; m < M; ++m) {
                                const float d1 = dotn<N>(Av, c[m], 0.f) - tg;
                                const float d2 = dotn<N>(Bv, c[m], 0.f);
                                const float e2 = d2 - tg2;
                                lB[m] = fmaf(d1, d1, lB[m]);
                                lC[m] = fmaf(e2, e2, lC[m]);
                                const float w1 = fac3 * d1, w2 = fac3 * d2;
#pragma unroll
                                for (int k = 0; k < N; ++k) {
                                    g[m][k] = fmaf(w1, Av[k], g[m][k]);
                                    g[m][k] = fmaf(w2, Bv[k], g[m][k]);
                                }
                            }
                        } else {
#pragma unroll
                            for (int m = 0; m < M; ++m) {
                                const float d1 = dotn<N>(Av, c[m], 0.f) - tg;
                                lB[m] = fmaf(d1, d1, lB[m]);
                                const float w1 = fac3 * d1;
#pragma unroll
                                for (int k = 0; k < N; ++k) g[m][k] = fmaf(w1, Av[k], g[m][k]);
                            }
                        }
                    }
                }
                release();
            }

            // ------------------------------ segment 2 (3): boundary rows ------------------
            for (int seg = 2; seg < TR::NSEG; ++seg) {
                for (int t = 0; t < p.ntile[seg]; ++t) {
                    const float *slot = acquire();
                    if (active) {
                        for (int r = slice; r < R_TILE; r += RL) {
                            float Av[4 * NQ];
                            load_row<NQ>(slot, r, Av);
                            if constexpr (PDE == PDE_AC) {     // periodic: (P_ub c - P_lb c) * P_diff
                                float Bv[4 * NQ], Dv[4 * NQ];
                                load_row<NQ>(slot + ARR_FLTS, r, Bv);
                                load_row<NQ>(slot + 2 * ARR_FLTS, r, Dv);
#pragma unroll
                                for (int m = 0; m < M; ++m) {
                                    const float d = dotn<N>(Av, c[m], 0.f) - dotn<N>(Bv, c[m], 0.f);
                                    if (seg == 2) lC[m] = fmaf(d, d, lC[m]); else lD[m] = fmaf(d, d, lD[m]);
                                    const float w = fac2 * d;
#pragma unroll
                                    for (int k = 0; k < N; ++k) g[m][k] = fmaf(w, Dv[k], g[m][k]);
                                }
                            } else {
                                const float tg = slot[VEC0 + r];
#pragma unroll
                                for (int m = 0; m < M; ++m) {
                                    const float dv = dotn<N>(Av, c[m], 0.f);
                                    const float e = dv - tg;
                                    float w;
                                    if constexpr (PDE == PDE_KG) { lD[m] = fmaf(e, e, lD[m]); w = fac2 * e; }
                                    else { lC[m] = fmaf(e, e, lC[m]); w = fac2 * dv; }   // Burgers: BC_u absent from update (N3)
#pragma unroll
                                    for (int k = 0; k < N; ++k) g[m][k] = fmaf(w, Av[k], g[m][k]);
                                }
                            }
                        }
                    }
                    release();
                }
            }

            // ------------------------------ end of pass: reduce, loss, update -------------
            for (int off = SL; off < 32; off <<= 1) {
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    lA[m] += __shfl_xor_sync(FULLM, lA[m], off);
                    lB[m] += __shfl_xor_sync(FULLM, lB[m], off);
                    lC[m] += __shfl_xor_sync(FULLM, lC[m], off);
                    lD[m] += __shfl_xor_sync(FULLM, lD[m], off);
#pragma unroll
                    for (int k = 0; k < N; ++k) g[m][k] += __shfl_xor_sync(FULLM, g[m][k], off);
                }
            }
#pragma unroll
            for (int m = 0; m < M; ++m) {
                float loss;
                if constexpr (PDE == PDE_KG)        // ((R + IC1) + IC2) + BC      KG_models.py:45-50
                    loss = __fadd_rn(__fadd_rn(__fadd_rn(__fdiv_rn(lA[m], p.nR), __fdiv_rn(lB[m], p.nI)),
                                               __fdiv_rn(lC[m], p.nI)), __fdiv_rn(lD[m], p.nB));
                else if constexpr (PDE == PDE_B)    // (R + IC) + BC               B_models.py:42-46
                    loss = __fadd_rn(__fadd_rn(__fdiv_rn(lA[m], p.nR), __fdiv_rn(lB[m], p.nI)), __fdiv_rn(lC[m], p.nB));
                else                                // ((R + IC) + BC) + BCx       AC_models.py:49-54
                    loss = __fadd_rn(__fadd_rn(__fadd_rn(__fdiv_rn(lA[m], p.nR), __fdiv_rn(lB[m], p.nI)),
                                               __fdiv_rn(lC[m], p.nB)), __fdiv_rn(lD[m], p.nB));
                lastloss[m] = loss;
                if (p.trace != nullptr && p.rec_every > 0 && (j % p.rec_every) == 0 && slice == 0 && oidx[m] >= 0)
                    p.trace[(size_t)oidx[m] * p.nrec + j / p.rec_every] = loss;
                if (j < p.epochs) {
                    if (loss < mn[m]) mn[m] = loss;
#pragma unroll
                    for (int k = 0; k < N; ++k) c[m][k] = __fsub_rn(c[m][k], __fmul_rn(p.lr, g[m][k]));
                }
            }
        }   // epochs

        if (slice == 0) {
#pragma unroll
            for (int m = 0; m < M; ++m) {
                if (oidx[m] >= 0) {
                    p.f_out[oidx[m]] = lastloss[m];
                    p.m_out[oidx[m]] = mn[m];
                    if (p.c_out != nullptr) {
#pragma unroll
                        for (int k = 0; k < N; ++k) p.c_out[(size_t)oidx[m] * N + k] = c[m][k];
                    }
                }
            }
        }
    }   // rounds
}

// host-side launcher signature shared by the per-(PDE, n) translation units
typedef cudaError_t (*sweep_launch_fn)(const SweepParams &p, int M, int ctas, size_t smem, cudaStream_t st);

template <int PDE, int N, int M, int MAXW>
cudaError_t launch_one(const SweepParams &p, int ctas, size_t smem, cudaStream_t st)
{
    auto kern = sweep_kernel<PDE, N, M, MAXW>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<ctas, p.W * 32, smem, st>>>(p);
    return cudaGetLastError();
}

// warps-per-CTA ceiling for each M (register budget: 64K regs / (32 * W) per thread)
constexpr int maxw_for_m(int M) { return M >= 4 ? 8 : (M == 3 ? 12 : 16); }

template <int PDE, int N>
cudaError_t launch_n(const SweepParams &p, int M, int ctas, size_t smem, cudaStream_t st)
{
    switch (M) {
        case 1: return launch_one<PDE, N, 1, maxw_for_m(1)>(p, ctas, smem, st);
        case 2: return launch_one<PDE, N, 2, maxw_for_m(2)>(p, ctas, smem, st);
        case 4: return launch_one<PDE, N, 4, maxw_for_m(4)>(p, ctas, smem, st);
        case 5: return launch_one<PDE, N, 5, maxw_for_m(5)>(p, ctas, smem, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace gp
