// sweep_rows.cuh -- row-partitioned variant of the coefficient sweep ("grid teams"), for grids with FEW parameters.
//
// Same arithmetic as sweep_kernel.cuh (reference: KG_optimizer.py:34-68 + KG_models.py:28-50 per epoch, Burgers / Allen-Cahn
// likewise), different decomposition.  The main kernel gives every warp (or team of <= 8 warps of ONE CTA) a set of
// parameters and streams ALL rows past it every epoch; with ~100..1000 parameters that leaves SMs idle (100 work items on
// 148 SMs), makes every CTA pull the whole stream through L2 once per epoch (2.8 MB x 128 CTAs x 2000 epochs for the
// default Burgers grid) and, at 10^6 rows, turns the sweep into an L2 -> SM streaming problem.  Here
//   * the ROWS are split across all CTAs of a cooperative grid (a balanced, contiguous share of every segment per CTA); if a
//     CTA's share fits in shared memory it is loaded ONCE and stays resident for all epochs, otherwise the residual rows are
//     streamed chunk by chunk, each row exactly once per epoch chip-wide;
//   * EVERY parameter is processed by EVERY CTA: thread = (parameter set of M parameters, row slice); c[M][N], grad[M][N] in
//     registers; rows are read from shared memory as LDS.128, broadcast across the parameter sets of a warp;
//   * the linear operator is never materialised ("T-free"): t = P_a.c + tp0 (P_b.c) + tp1 (P.c) and
//     grad += first P_a + (tp0 first) P_b + (...) P -- 6n (KG, AC) / 8n (Burgers) FMAs per row and parameter instead of
//     4n / 6n, in exchange for which any M parameters can share a thread (no sibling structure needed: Burgers and
//     Allen-Cahn grids have none);
//   * per epoch the partial gradients / losses go through global memory: CTA partials -> grid barrier -> every
//     (parameter, component) pair is summed over the CTAs in fixed order by ONE thread of the grid, which also applies
//     c <- c - lr g and keeps the loss bookkeeping (final, min-prefix, trace) -> grid barrier -> threads reload their c.
// Results differ from the main kernel only by summation order and the fused T (within the 1e-5 parity tolerance).
//
// Two flavours of the per-epoch exchange (template flag CL):
//   CL = false: one cooperative grid; partials and coefficients go through global memory, two grid barriers per epoch.  Any
//               row count (the residual share is streamed when it does not fit) -- the 10^6-row test sweep runs here.
//   CL = true : thread-block CLUSTERS.  The parameters are split over the clusters, the rows over the S CTAs of a cluster
//               (S <= 16, every share resident in shared memory); per epoch the CTA partials stay in shared memory, one
//               cluster barrier, every CTA sums its slice of the (parameter, component) pairs over the S partials through
//               distributed shared memory (fixed order) and PUSHES the updated coefficient into the coefficient table of
//               all S CTAs, a second cluster barrier.  No global-memory round trip and no grid-wide barrier on the epoch's
//               critical path: this is what the default-size grids (100..1000 parameters, 2000 epochs of a few us) need.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace gp {

namespace cgx = cooperative_groups;

constexpr int ROWS_THREADS = 256;
constexpr int ROWS_M = 4;                      // parameters per thread
constexpr int ROWS_MAX_PARAMS = ROWS_THREADS * ROWS_M;
constexpr int ROWS_KC = 4;                     // components per pass of the row-slice reduction
static_assert(ROWS_M == 4, "the coefficient table is read as one float4 per thread and component");

struct RowsParams {
    const float *packed;                       // tile stream of the main kernel (RT_PRIV-row tiles, SlotFloats per tile)
    int   nseg;
    int   seg_rows[MAX_SEG];                   // rows per segment
    int   seg_tile0[MAX_SEG];                  // first tile of each segment in the stream
    int   slot;                                // floats per tile slot
    int   chunk_rows;                          // residual rows that fit in the streaming buffer (multiple of 4)
    float invR, invI, invB, fac1, fac2, fac3, lr;
    int   epochs, rec_every, nrec, has_fhat;
    int   Np, npsp, rs;                        // parameters, parameter sets per CTA (<= 256), row slices = 256 / npsp (threads past npsp * rs idle)
    int   S, C;                                // cluster flavour: CTAs per cluster, clusters (parameters split evenly over them)
    int   tpp;                                 // lanes that share one (parameter, component) pair in the owner phase (power of two <= 32)
    const float4 *pp;                          // [Np] (tp0, tp1, sp0, sp1)
    const float *c0; int c0_per_param;
    float *c_out, *f_out, *m_out, *trace;
    unsigned long long *key_out; const int *key_index;      // fused arg-max key (see SweepParams)
    float *cglob;                              // [N][npsp * M]  current coefficients, component-major (one float4 = the four parameters of a thread)
    float *partial;                            // [grid][Np][N+1]
};

// rows [lo, hi) of `total` owned by CTA b of G (balanced to one row)
__device__ __forceinline__ void rows_share(int total, int b, int G, int &lo, int &hi)
{
    lo = (int)(((long long)total * b) / G);
    hi = (int)(((long long)total * (b + 1)) / G);
}

// copy rows [r0, r0 + nr) of one segment from the tile stream into a shared-memory region laid out
// [block][quad][row][4] then [vec][row]; NQX quads per block are kept
template <int NARR, int NVEC, int NQX>
__device__ __forceinline__ void rows_load(const RowsParams &p, int seg, int r0, int nr, int cap, float *dst)
{
    constexpr int RT = RT_PRIV;
    const int per_row = NARR * NQX;
    for (int e = threadIdx.x; e < nr * per_row; e += ROWS_THREADS) {
        const int r = e % nr, bq = e / nr, blk = bq / NQX, q = bq % NQX;
        const int row = r0 + r;
        const float *src = p.packed + (size_t)(p.seg_tile0[seg] + row / RT) * p.slot + blk * 16 * RT + q * 4 * RT + (row % RT) * 4;
        *reinterpret_cast<float4 *>(dst + ((blk * NQX + q) * cap + r) * 4) = *reinterpret_cast<const float4 *>(src);
    }
    for (int e = threadIdx.x; e < nr * NVEC; e += ROWS_THREADS) {
        const int r = e % nr, v = e / nr, row = r0 + r;
        dst[NARR * NQX * cap * 4 + v * cap + r] = p.packed[(size_t)(p.seg_tile0[seg] + row / RT) * p.slot + NARR * 16 * RT + v * RT + (row % RT)];
    }
}

// asynchronous variant (cp.async, 16-byte and 4-byte copies) for the streamed residual chunks: issue, commit, wait later
template <int NARR, int NVEC, int NQX>
__device__ __forceinline__ void rows_load_async(const RowsParams &p, int seg, int r0, int nr, int cap, float *dst)
{
    constexpr int RT = RT_PRIV;
    const int per_row = NARR * NQX;
    for (int e = threadIdx.x; e < nr * per_row; e += ROWS_THREADS) {
        const int r = e % nr, bq = e / nr, blk = bq / NQX, q = bq % NQX;
        const int row = r0 + r;
        const float *src = p.packed + (size_t)(p.seg_tile0[seg] + row / RT) * p.slot + blk * 16 * RT + q * 4 * RT + (row % RT) * 4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + ((blk * NQX + q) * cap + r) * 4)), "l"(src) : "memory");
    }
    for (int e = threadIdx.x; e < nr * NVEC; e += ROWS_THREADS) {
        const int r = e % nr, v = e / nr, row = r0 + r;
        const float *src = p.packed + (size_t)(p.seg_tile0[seg] + row / RT) * p.slot + NARR * 16 * RT + v * RT + (row % RT);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst + NARR * NQX * cap * 4 + v * cap + r)), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int NQX>
__device__ __forceinline__ void rows_ld(const float *blk, int cap, int r, float (&v)[4 * NQX])
{
#pragma unroll
    for (int q = 0; q < NQX; ++q) {
        const float4 x = *reinterpret_cast<const float4 *>(blk + (q * cap + r) * 4);
        v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
    }
}

template <int PDE, int N, bool CL>
__global__ void __launch_bounds__(ROWS_THREADS, 1) sweep_rows_kernel(const RowsParams p)
{
    using TR = PdeTraits<PDE>;
    constexpr int M = ROWS_M;
    constexpr int NQX = (N + 3) / 4;
    constexpr int NV = N + 1;
    constexpr int NARR = TR::NARR, NVEC = TR::NVEC;
    cgx::grid_group grid = cgx::this_grid();
    cgx::cluster_group cluster = cgx::this_cluster();
    extern __shared__ __align__(16) float rsm[];
    const int tid = threadIdx.x;
    // rows are shared out over the G CTAs of the grid (CL: of the cluster); b = this CTA's place among them
    const int b = CL ? (int)cluster.block_rank() : (int)blockIdx.x, G = CL ? p.S : (int)gridDim.x;
    // parameters [pbase, pbase + npc) belong to this CTA's cluster (CL) / all parameters to every CTA
    const int cid = CL ? (int)blockIdx.x / p.S : 0;
    const int pbase = CL ? (int)(((long long)p.Np * cid) / p.C) : 0;
    const int npc = CL ? (int)(((long long)p.Np * (cid + 1)) / p.C) - pbase : p.Np;
    const int ps = tid % p.npsp, RS = p.rs;
    const bool idle = tid / p.npsp >= RS;                                // threads past npsp * RS take no rows
    const int rs = idle ? 0 : tid / p.npsp;

    // ---- this CTA's share of every segment; shared-memory regions
    int lo[MAX_SEG], hi[MAX_SEG], cap[MAX_SEG];
    float *reg[MAX_SEG];
    int off = 0;
    const int row_floats = NARR * NQX * 4 + NVEC;
    for (int s = 0; s < p.nseg; ++s) {
        rows_share(p.seg_rows[s], b, G, lo[s], hi[s]);
        // the LARGEST share sizes the region: every CTA has the same shared-memory layout (the cluster flavour addresses
        // the other CTAs' buffers by its own offsets)
        int n = (p.seg_rows[s] + G - 1) / G;
        const bool streamed = (s == 0 && n > p.chunk_rows);              // residual rows go through a double buffer of chunks
        if (streamed) n = p.chunk_rows;
        cap[s] = (n + 3) & ~3;
        if (cap[s] < 4) cap[s] = 4;
        reg[s] = rsm + off;
        off += cap[s] * row_floats * (streamed ? 2 : 1);
        off = (off + 3) & ~3;
    }
    float *red = rsm + off;                                            // [RS][npsp * M][ROWS_KC]   (RS > 1 only)
    float *csm = red;                                                  // [N][npsp * M] staging of the new coefficients (after the owner
                                                                       // phase; `red` is dead by then, the host sizes the region for both)
    const int NPP = p.npsp * M;
    // cluster flavour: this CTA's partial sums [NPP][NV] and its copy of the coefficient table [N][NPP], both reachable by
    // the other CTAs of the cluster through distributed shared memory
    float *part = red + (RS > 1 ? (size_t)ROWS_THREADS * M * ROWS_KC : 0);
    float *ctab = part + (((size_t)NPP * NV + 3) & ~(size_t)3);
    if constexpr (CL) csm = ctab;
    const int res_rows = hi[0] - lo[0];
    const bool resident = res_rows <= cap[0];
    for (int s = resident ? 0 : 1; s < p.nseg; ++s) rows_load<NARR, NVEC, NQX>(p, s, lo[s], hi[s] - lo[s], cap[s], reg[s]);
    __syncthreads();

    // ---- parameters of this thread
    float c[M][N], g[M][N];
    float tp0[M], tp1[M], sp0[M], sp1[M];
    bool valid[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const int pi = pbase + ps * M + m;
        valid[m] = !idle && ps * M + m < npc;
        const float4 q = valid[m] ? p.pp[pi] : make_float4(0.f, 0.f, 0.f, 0.f);
        tp0[m] = q.x; tp1[m] = q.y; sp0[m] = q.z; sp1[m] = q.w;
#pragma unroll
        for (int k = 0; k < N; ++k) c[m][k] = valid[m] ? (p.c0_per_param ? p.c0[(size_t)pi * N + k] : p.c0[k]) : 0.f;
    }
    const bool has_fhat = p.has_fhat != 0;
    const int Q = p.Np * NV;
    const int gthreads = G * ROWS_THREADS, gtid = b * ROWS_THREADS + tid;
    if constexpr (CL) {
        for (int e = tid; e < N * NPP; e += ROWS_THREADS) {
            const int v = e / NPP, pl = e % NPP;
            ctab[e] = pl < npc ? (p.c0_per_param ? p.c0[(size_t)(pbase + pl) * N + v] : p.c0[v]) : 0.f;
        }
        __syncthreads();
    }

    for (int j = 0; j <= p.epochs; ++j) {
        float lacc[M], lsum[M];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            lacc[m] = 0.f; lsum[m] = 0.f;
#pragma unroll
            for (int k = 0; k < N; ++k) g[m][k] = 0.f;
        }
        // ------------------------------ segment 0: PDE residual rows (resident, or streamed chunk by chunk)
        const int buf_floats = cap[0] * row_floats;
        if (!resident) rows_load_async<NARR, NVEC, NQX>(p, 0, lo[0], min(cap[0], res_rows), cap[0], reg[0]);
        for (int r0 = 0, ci = 0; r0 < res_rows; r0 += cap[0], ++ci) {
            const int nr = min(cap[0], res_rows - r0);
            const float *base = reg[0];
            if (!resident) {
                // chunk ci is in flight in buffer ci & 1; start chunk ci + 1 into the other buffer (its previous contents were
                // consumed before the barrier that ended the previous iteration), then wait for chunk ci only
                const int rn = r0 + cap[0];
                if (rn < res_rows) {
                    rows_load_async<NARR, NVEC, NQX>(p, 0, lo[0] + rn, min(cap[0], res_rows - rn), cap[0], reg[0] + ((ci + 1) & 1) * buf_floats);
                    asm volatile("cp.async.wait_group 1;" ::: "memory");
                } else {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                }
                __syncthreads();
                base = reg[0] + (ci & 1) * buf_floats;
            }
            const float *A0 = base, *A1 = A0 + NQX * cap[0] * 4, *A2 = A1 + NQX * cap[0] * 4;
            const float *A3 = A2 + NQX * cap[0] * 4;                    // Burgers only
            const float *V0 = base + NARR * NQX * cap[0] * 4, *V1 = V0 + cap[0];
            for (int r = idle ? nr : rs; r < nr; r += RS) {
                float Pa[4 * NQX], Pb[4 * NQX], Pv[4 * NQX];
                rows_ld<NQX>(A0, cap[0], r, Pa);                         // P_tt (KG) / P_t (B, AC)
                rows_ld<NQX>(A1, cap[0], r, Pb);                         // P_xx
                float forcing = 0.f, fh = 0.f;
                if constexpr (PDE == PDE_KG) { forcing = V0[r]; if (has_fhat) fh = V1[r]; }
                else { if (has_fhat) fh = V0[r]; }
                if constexpr (PDE == PDE_B) {
                    float Px[4 * NQX];
                    rows_ld<NQX>(A2, cap[0], r, Px);
                    rows_ld<NQX>(A3, cap[0], r, Pv);
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        float u = 0.f, ux = 0.f, a = 0.f, bb = 0.f;
#pragma unroll
                        for (int k = 0; k < N; ++k) {
                            u = fmaf(Pv[k], c[m][k], u); ux = fmaf(Px[k], c[m][k], ux);
                            a = fmaf(Pa[k], c[m][k], a); bb = fmaf(Pb[k], c[m][k], bb);
                        }
                        const float first = fmaf(u, ux, fmaf(tp0[m], bb, a));        // (P_t - nu P_xx).c + u u_x   B_optimizer.py:36-40
                        const float dl = first - fh;
                        lacc[m] = fmaf(dl, dl, lacc[m]);
                        const float wb = tp0[m] * first, wx = first * u, wp = first * ux;
#pragma unroll
                        for (int k = 0; k < N; ++k)
                            g[m][k] = fmaf(wp, Pv[k], fmaf(wx, Px[k], fmaf(wb, Pb[k], fmaf(first, Pa[k], g[m][k]))));
                    }
                } else {
                    rows_ld<NQX>(A2, cap[0], r, Pv);                     // P
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        float u = 0.f, a = 0.f, bb = 0.f;
#pragma unroll
                        for (int k = 0; k < N; ++k) {
                            u = fmaf(Pv[k], c[m][k], u); a = fmaf(Pa[k], c[m][k], a); bb = fmaf(Pb[k], c[m][k], bb);
                        }
                        const float t = fmaf(tp1[m], u, fmaf(tp0[m], bb, a));         // T.c with T = P_a + tp0 P_b + tp1 P
                        float first, nl;
                        if constexpr (PDE == PDE_KG) { first = fmaf(sp0[m], u * u, t) + forcing; nl = sp1[m] * u; }          // g u^2 ; 2g u
                        else { const float u2 = u * u; first = fmaf(sp0[m], u2 * u, t); nl = sp1[m] * u2; }                  // e u^3 ; 3e u^2
                        const float dl = first - fh;
                        lacc[m] = fmaf(dl, dl, lacc[m]);
                        const float wb = tp0[m] * first, wp = fmaf(first, nl, tp1[m] * first);
#pragma unroll
                        for (int k = 0; k < N; ++k)
                            g[m][k] = fmaf(wp, Pv[k], fmaf(wb, Pb[k], fmaf(first, Pa[k], g[m][k])));
                    }
                }
            }
            if (!resident) __syncthreads();                             // this buffer is refilled two chunks ahead
        }
#pragma unroll
        for (int m = 0; m < M; ++m) {
            lsum[m] = lacc[m] * p.invR; lacc[m] = 0.f;
#pragma unroll
            for (int k = 0; k < N; ++k) g[m][k] *= p.fac1;
        }
        // ------------------------------ segment 1: initial-condition rows
        {
            const int nr = hi[1] - lo[1];
            const float *A0 = reg[1], *A1 = A0 + NQX * cap[1] * 4;
            const float *V0 = reg[1] + NARR * NQX * cap[1] * 4, *V1 = V0 + cap[1];
            for (int r = idle ? nr : rs; r < nr; r += RS) {
                float Av[4 * NQX];
                rows_ld<NQX>(A0, cap[1], r, Av);
                const float tg = V0[r];
                if constexpr (PDE == PDE_KG) {
                    float Bv[4 * NQX];
                    rows_ld<NQX>(A1, cap[1], r, Bv);
                    const float tg2 = V1[r];
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        float d1 = 0.f, d2 = 0.f;
#pragma unroll
                        for (int k = 0; k < N; ++k) { d1 = fmaf(Av[k], c[m][k], d1); d2 = fmaf(Bv[k], c[m][k], d2); }
                        d1 -= tg;
                        const float e2 = d2 - tg2;
                        lacc[m] = fmaf(d1, d1, lacc[m]); lacc[m] = fmaf(e2, e2, lacc[m]);
                        const float w1 = p.fac3 * d1, w2 = p.fac3 * d2;               // IC_u2 absent from the update (N3)
#pragma unroll
                        for (int k = 0; k < N; ++k) g[m][k] = fmaf(w2, Bv[k], fmaf(w1, Av[k], g[m][k]));
                    }
                } else {
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        float d1 = 0.f;
#pragma unroll
                        for (int k = 0; k < N; ++k) d1 = fmaf(Av[k], c[m][k], d1);
                        d1 -= tg;
                        lacc[m] = fmaf(d1, d1, lacc[m]);
                        const float w1 = p.fac3 * d1;
#pragma unroll
                        for (int k = 0; k < N; ++k) g[m][k] = fmaf(w1, Av[k], g[m][k]);
                    }
                }
            }
        }
#pragma unroll
        for (int m = 0; m < M; ++m) { lsum[m] = fmaf(lacc[m], p.invI, lsum[m]); lacc[m] = 0.f; }
        // ------------------------------ segment 2 (3): boundary rows
        for (int s = 2; s < TR::NSEG; ++s) {
            const int nr = hi[s] - lo[s];
            const float *A0 = reg[s], *A1 = A0 + NQX * cap[s] * 4, *A2 = A1 + NQX * cap[s] * 4;
            const float *V0 = reg[s] + NARR * NQX * cap[s] * 4;
            for (int r = idle ? nr : rs; r < nr; r += RS) {
                float Av[4 * NQX];
                rows_ld<NQX>(A0, cap[s], r, Av);
                if constexpr (PDE == PDE_AC) {                       // periodic: (P_ub c - P_lb c) * P_diff
                    float Bv[4 * NQX], Dv[4 * NQX];
                    rows_ld<NQX>(A1, cap[s], r, Bv);
                    rows_ld<NQX>(A2, cap[s], r, Dv);
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        float da = 0.f, db = 0.f;
#pragma unroll
                        for (int k = 0; k < N; ++k) { da = fmaf(Av[k], c[m][k], da); db = fmaf(Bv[k], c[m][k], db); }
                        const float d = da - db;
                        lacc[m] = fmaf(d, d, lacc[m]);
                        const float w = p.fac2 * d;
#pragma unroll
                        for (int k = 0; k < N; ++k) g[m][k] = fmaf(w, Dv[k], g[m][k]);
                    }
                } else {
                    const float tg = V0[r];
#pragma unroll
                    for (int m = 0; m < M; ++m) {
                        float dv = 0.f;
#pragma unroll
                        for (int k = 0; k < N; ++k) dv = fmaf(Av[k], c[m][k], dv);
                        const float e = dv - tg;
                        lacc[m] = fmaf(e, e, lacc[m]);
                        const float w = (PDE == PDE_KG) ? p.fac2 * e : p.fac2 * dv;   // Burgers: BC_u absent from the update (N3)
#pragma unroll
                        for (int k = 0; k < N; ++k) g[m][k] = fmaf(w, Av[k], g[m][k]);
                    }
                }
            }
        }
#pragma unroll
        for (int m = 0; m < M; ++m) lsum[m] = fmaf(lacc[m], p.invB, lsum[m]);

        // ------------------------------ CTA partial: sum over the row slices (fixed order) -> [parameters][NV], to global
        // memory (grid flavour) or to this CTA's shared memory (cluster flavour)
        const int QL = CL ? npc * NV : Q;                                 // (parameter, component) pairs this CTA works on
        float *mine = CL ? part : p.partial + (size_t)b * Q;
        if (RS == 1) {
#pragma unroll
            for (int m = 0; m < M; ++m)
                if (valid[m]) {
                    float *dst = mine + (size_t)(ps * M + m) * NV;
#pragma unroll
                    for (int k = 0; k < N; ++k) dst[k] = g[m][k];
                    dst[N] = lsum[m];
                }
        } else {
            // the RS row slices of a parameter set are summed through shared memory, ROWS_KC components per pass
            // (buffer [RS][npsp * M][ROWS_KC]: 16 KB instead of [RS][npsp * M][NV])
            constexpr int KC = ROWS_KC;
            const int plane = p.npsp * M * KC;
#pragma unroll
            for (int k0 = 0; k0 < NV; k0 += KC) {
                if (!idle) {
                    float *my = red + (size_t)rs * plane + (size_t)ps * M * KC;
#pragma unroll
                    for (int m = 0; m < M; ++m)
#pragma unroll
                        for (int kk = 0; kk < KC; ++kk)
                            if (k0 + kk < NV) my[m * KC + kk] = (k0 + kk < N) ? g[m][(k0 + kk < N) ? k0 + kk : 0] : lsum[m];
                }
                __syncthreads();
                const int kc = (NV - k0 < KC) ? NV - k0 : KC;
                for (int e = tid; e < (QL / NV) * KC; e += ROWS_THREADS) {
                    const int pl = e / KC, kk = e % KC;
                    if (kk < kc) {
                        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                        int q = 0;
                        for (; q + 3 < RS; q += 4) {
                            s0 += red[(size_t)q * plane + e]; s1 += red[(size_t)(q + 1) * plane + e];
                            s2 += red[(size_t)(q + 2) * plane + e]; s3 += red[(size_t)(q + 3) * plane + e];
                        }
                        for (; q < RS; ++q) s0 += red[(size_t)q * plane + e];
                        mine[(size_t)pl * NV + k0 + kk] = (s0 + s1) + (s2 + s3);
                    }
                }
                __syncthreads();
            }
        }
        const bool last = (j == p.epochs);
        const bool rec = p.trace != nullptr && p.rec_every > 0 && (j % p.rec_every) == 0;
        // loss bookkeeping of the owner of pair (pi, N): trace, final loss + arg-max key, running minimum of the earlier losses
        auto keep_loss = [&](int pi, float s) {
            if (rec) p.trace[(size_t)pi * p.nrec + j / p.rec_every] = s;
            if (last) {
                p.f_out[pi] = s;
                if (p.epochs == 0) p.m_out[pi] = __int_as_float(0x7f800000);
                if (p.key_out != nullptr && s > 0.f)
                    atomicMax(p.key_out, ((unsigned long long)__float_as_uint(s) << 32) |
                                             (unsigned long long)(0xFFFFFFFFu - (unsigned)(p.key_index ? p.key_index[pi] : pi)));
            } else {
                const float mprev = (j == 0) ? __int_as_float(0x7f800000) : __ldcg(p.m_out + pi);
                p.m_out[pi] = (s < mprev) ? s : mprev;
            }
        };
        if constexpr (CL) {
            // ------------------------------ cluster exchange: the S partials are summed through distributed shared memory
            // (CTA b owns the pairs e = b (mod S); fixed order q = 0..S-1 => deterministic), the owner pushes the new
            // coefficient into the table of every CTA of the cluster
            cluster.sync();
            const int S = p.S;
            // CTA b owns the pairs e = b (mod S).  One thread per (owned pair, source CTA): S consecutive lanes fetch the S
            // partials of a pair in one round trip, a butterfly over those lanes (fixed tree => deterministic, every lane
            // ends with the same bits) sums them, and lane q of the group pushes the new coefficient to CTA q.
            const int owned = (QL - b + S - 1) / S;                      // pairs b, b + S, ... < QL
            const int tasks = ((owned * S + 31) / 32) * 32;              // warp-uniform trip count (full-mask shuffles)
            for (int t = tid; t < tasks; t += ROWS_THREADS) {
                const int ei = t / S, q = t % S, e = b + S * ei;
                const bool act = ei < owned;
                float sum = act ? cluster.map_shared_rank(part, q)[e] : 0.f;
                for (int off = S >> 1; off; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
                if (act) {
                    const int pl = e / NV, v = e % NV, pi = pbase + pl;
                    if (v < N) {
                        const float cur = ctab[v * NPP + pl];
                        if (last) { if (q == 0 && p.c_out) p.c_out[(size_t)pi * N + v] = cur; }
                        else cluster.map_shared_rank(ctab, q)[v * NPP + pl] = __fsub_rn(cur, __fmul_rn(p.lr, sum));
                    } else if (q == 0) keep_loss(pi, sum);
                }
            }
            cluster.sync();                                              // tables complete; partials free for the next epoch
            if (last) break;
        } else {
        __threadfence();
        grid.sync();

        // ------------------------------ owners: `tpp` lanes of the grid per (parameter, component) pair sum the G CTA
        // partials (strided over the CTAs, four loads in flight per lane, fixed combination order => deterministic)
        {
            const int tpp = p.tpp, ppw = 32 / tpp;                       // pairs per warp and step
            const int lane = tid & 31, sub = lane & (tpp - 1);
            const int gwarp = gtid >> 5, gwarps = gthreads >> 5;
            for (int e0 = gwarp * ppw; e0 < Q; e0 += gwarps * ppw) {     // warp-uniform trip count (full-mask shuffles below)
                const int e = e0 + lane / tpp;
                const bool act = e < Q;
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                if (act) {
                    const float *src = p.partial + e;
                    int q = sub;
                    for (; q + 3 * tpp < G; q += 4 * tpp) {
                        s0 += __ldcg(src + (size_t)q * Q); s1 += __ldcg(src + (size_t)(q + tpp) * Q);
                        s2 += __ldcg(src + (size_t)(q + 2 * tpp) * Q); s3 += __ldcg(src + (size_t)(q + 3 * tpp) * Q);
                    }
                    for (; q < G; q += tpp) s0 += __ldcg(src + (size_t)q * Q);
                }
                float s = (s0 + s1) + (s2 + s3);
                for (int off = tpp >> 1; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                if (act && sub == 0) {
                    const int pi = e / NV, v = e % NV;
                    if (v < N) {
                        const float cur = (j == 0) ? (p.c0_per_param ? p.c0[(size_t)pi * N + v] : p.c0[v]) : __ldcg(p.cglob + (size_t)v * NPP + pi);
                        if (last) { if (p.c_out) p.c_out[(size_t)pi * N + v] = cur; }
                        else p.cglob[(size_t)v * NPP + pi] = __fsub_rn(cur, __fmul_rn(p.lr, s));
                    } else keep_loss(pi, s);
                }
            }
        }
        if (last) break;
        __threadfence();
        grid.sync();
        }
        // every CTA needs every coefficient: one coalesced copy of the table into shared memory (each float requested once per
        // CTA -- 148 x 256 threads reading their own strided scalars made the few L2 lines of the table a hot spot), then one
        // LDS.128 per component gives a thread the values of its four parameters
        if constexpr (!CL) {
            for (int e = tid; e < N * NPP / 4; e += ROWS_THREADS)
                reinterpret_cast<float4 *>(csm)[e] = __ldcg(reinterpret_cast<const float4 *>(p.cglob) + e);
            __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const float4 v = *reinterpret_cast<const float4 *>(csm + k * NPP + ps * M);
            c[0][k] = valid[0] ? v.x : 0.f; c[1][k] = valid[1] ? v.y : 0.f; c[2][k] = valid[2] ? v.z : 0.f; c[3][k] = valid[3] ? v.w : 0.f;
        }
        if constexpr (!CL) __syncthreads();                             // csm aliases red: the next epoch's partial sums overwrite it
    }
}

template <int PDE, int N>
cudaError_t rows_launch(const RowsParams &p, int grid, size_t smem, cudaStream_t st)
{
    RowsParams pp = p;
    if (p.S > 0) {                                                       // cluster flavour: grid = C clusters of S CTAs
        auto kern = sweep_rows_kernel<PDE, N, true>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (p.S > 8) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            if (e != cudaSuccess) return e;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(ROWS_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = p.S; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kern, pp);
    }
    auto kern = sweep_rows_kernel<PDE, N, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    void *args[] = {&pp};
    return cudaLaunchCooperativeKernel((void *)kern, dim3(grid), dim3(ROWS_THREADS), args, smem, st);
}

// how many clusters of S CTAs with `smem` bytes each can be co-resident (0 if the shape cannot launch)
template <int PDE, int N>
int rows_max_clusters(int S, size_t smem)
{
    auto kern = sweep_rows_kernel<PDE, N, true>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    if (S > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(S); cfg.blockDim = dim3(ROWS_THREADS); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = S; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

typedef cudaError_t (*rows_launch_fn)(const RowsParams &p, int grid, size_t smem, cudaStream_t st);
typedef int (*rows_clusters_fn)(int S, size_t smem);

}  // namespace gp
