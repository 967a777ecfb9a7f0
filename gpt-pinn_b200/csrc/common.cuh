// common.cuh -- shared definitions for the sm_100a GPT-PINN kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gp {

// ----------------------------------------------------------------------------------------
// Packed activation stream ("mirror") layout.
//
// The reference keeps P, P_xx, ... as [N, number_of_neurons] row-major buffers and hands the
// sweep column-slice views (row stride 15 floats = 60 B: not 16-B aligned, so not TMA-able,
// SURVEY.md H4).  The pack kernel rewrites them once per sweep call into a tile stream:
//
//   stream = tile[0] tile[1] ... tile[tiles_per_pass-1]        (segment after segment)
//   tile   = NARR matrix blocks, then NVEC row vectors          (SLOT floats, 16-B multiple)
//   matrix block = [4 column-quads][RT rows][4 floats]          (k = 4*quad + lane component)
//   row vector   = [RT]                                          (RT = rows per tile, 32 or 64)
//
// so that (a) one tile is ONE contiguous cp.async.bulk (TMA bulk copy) into shared memory,
// (b) a thread reads 4 neurons of one row with one LDS.128, and (c) the row-slices of a warp
// touch consecutive 16-B chunks (bank-conflict free), while lanes that share a row broadcast.
// Rows past N and columns past n are zero, which contributes nothing to any sum.
// ----------------------------------------------------------------------------------------
// Both pipelines use 64-row tiles.  "Private ring" kernels: every warp (or team of warps) streams its own copy of the
// tiles through a double buffer and builds T in place -- no inter-warp coupling, dynamic work distribution.
// "Shared ring" kernels: all warps of a CTA consume one 4-deep stream (small grids, rows split 32 ways).
constexpr int RT_PRIV   = 64;
constexpr int RT_SHARED = 64;
constexpr int NSTAGE   = 4;                  // shared ring depth
constexpr int LOOKAHEAD = 2;                 // shared ring: tiles in flight ahead of the consumers
constexpr int NST_PRIV = 2;                  // private ring depth (per warp): one tile in use, one in flight
constexpr int MAX_SEG  = 4;
constexpr int RED_STRIDE = 20;               // floats per lane in the team-reduction buffer (16 used; 80-B stride: conflict-free LDS.128)
constexpr int TEAM_BYTES_PER_WARP = 32 * RED_STRIDE * 4 + 4;     // reduction rows + the team's ticket slot

constexpr int PDE_KG = 0, PDE_B = 1, PDE_AC = 2;


template <int PDE> struct PdeTraits;
template <> struct PdeTraits<PDE_KG> {      // residual: Ptt,Pxx,P | xcos,fhat ; IC: PIC,PICt | u1,u2 ; BC: PBC | BCu
    static constexpr int NARR = 3, NVEC = 2, NSEG = 3, TARR = 1;   // TARR: # per-warp derived blocks
};
template <> struct PdeTraits<PDE_B> {       // residual: Pt,Pxx,Px,P | fhat ; IC: PIC | ICu ; BC: PBC | BCu
    static constexpr int NARR = 4, NVEC = 1, NSEG = 3, TARR = 1;
};
template <> struct PdeTraits<PDE_AC> {      // residual: Pt,Pxx,P | fhat ; IC: PIC | ICu ; BC: Pub,Plb,Pdiff ; BCx: Pubx,Plbx,Pdiffx
    static constexpr int NARR = 3, NVEC = 1, NSEG = 4, TARR = 1;
};
template <int PDE, int RT> struct SlotFloats { static constexpr int value = PdeTraits<PDE>::NARR * 16 * RT + PdeTraits<PDE>::NVEC * RT; };

struct ItemDesc {          // one warp work item: a set of parameters that share the same T
    float tp0, tp1;        // KG: alpha, beta    B: -nu, 0      AC: -lambda, -eps
    int   sib_start;       // first sibling in SibDesc[]
    int   sib_count;       // 1 .. SL*M
    int   SL;              // sibling lanes of this item (power of two); row slices RL = 32/SL
    int   t_slot;          // private-ring kernels: index of the group's precomputed T array in SweepParams::tglob, or -1
                           // (build T in shared memory / in registers from the staged P-derivative tiles)
    int   pad[2];
};
struct SibDesc {
    float sp0, sp1;        // KG: gamma, 2*gamma  B: 0,0        AC: eps, 3*eps
    int   out_idx;         // row of the caller's grid
    int   pad;
};

struct SweepParams {
    const float   *packed;           // tile stream
    int            ntile[MAX_SEG];   // tiles per segment
    int            tiles_per_pass;
    float          invR, invI, invB; // 1/N_R, 1/N_IC, 1/N_BC as fp32 (loss = sum of per-segment means)
    float          fac1, fac2, fac3; // (float)(2/NR), (float)(2/NB), (float)(2/NI)
    float          lr;
    int            epochs, rec_every, nrec;
    int            has_fhat;
    const ItemDesc *items;
    int            n_items;
    const SibDesc  *sibs;
    const float    *c0;
    int            c0_per_param;
    float          *c_out, *f_out, *m_out, *trace;
    int            W;                // warps per CTA
    int            rounds;           // shared-ring kernels: static rounds of gridDim.x * W items
    unsigned int  *counter;          // private-ring kernels: global work-item ticket (zeroed before launch)
    const float   *bound;            // private-ring kernels: [Np] per-parameter early-exit bound B_p <= L_seq(p), or nullptr
    unsigned long long *key_out;     // optional: atomicMax of (float_bits(final loss) << 32) | (0xFFFFFFFF - global index) over all rows
    const int     *key_index;        // optional [Np]: global grid index of each row (nullptr: the row number)
    const float   *tglob;            // private-ring kernels: precomputed T arrays, [t_slot][residual tile][NQT][RT][4] (see build_t_kernel)
    long long      t_stride;         // floats per T array (= ntile[0] * NQT * 4 * RT)
    int            K;                // private-ring kernels: warps per team (power of two dividing W); a team shares one
                                     // work item, its warps take every K-th tile and sum their gradients once per epoch
};

__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// ---- PTX wrappers: mbarrier + TMA bulk copy --------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    const uint32_t a = smem_u32(bar);
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    } while (!ok);
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// the same for data that is read once and should not displace the reused stream in L2 (evict-first cache hint)
__device__ __forceinline__ void bulk_g2s_stream(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

}  // namespace gp
