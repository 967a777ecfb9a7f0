// pack.cu -- rewrites the reference's strided [N, ld] activation views into the tile stream
// described in common.cuh (one contiguous, 16-B aligned slot per 64-row tile).
#include "pack.cuh"

namespace gp {

__global__ void pack_kernel(const PackJobs jobs, float *__restrict__ stream)
{
    // blockIdx.y = job, blockIdx.x * blockDim.x + threadIdx.x = element (row * 16 + k) or row
    const int j = blockIdx.y;
    if (j < jobs.nmat) {
        const PackMat &m = jobs.mat[j];
        const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        const long long row = e >> 4;
        const int k = (int)(e & 15);
        if (row >= m.nrows || k >= m.ncols) return;
        const float v = m.src[row * m.ld + k];
        const long long tile = m.tile0 + row / jobs.rt;
        const int r = (int)(row % jobs.rt);
        stream[tile * jobs.slot + (long long)m.arr * 16 * jobs.rt + (k >> 2) * 4 * jobs.rt + r * 4 + (k & 3)] = v;
    } else {
        const PackVec &v = jobs.vec[j - jobs.nmat];
        const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        if (row >= v.nrows || v.src == nullptr) return;
        const long long tile = v.tile0 + row / jobs.rt;
        const int r = (int)(row % jobs.rt);
        stream[tile * jobs.slot + (long long)jobs.narr * 16 * jobs.rt + v.vec * jobs.rt + r] = v.src[row];
    }
}

cudaError_t launch_pack(const PackJobs &jobs, float *stream, long long max_rows, cudaStream_t st)
{
    const int threads = 256;
    const long long elems = max_rows * 16;
    dim3 grid((unsigned)((elems + threads - 1) / threads), (unsigned)(jobs.nmat + jobs.nvec));
    pack_kernel<<<grid, threads, 0, st>>>(jobs, stream);
    return cudaGetLastError();
}

// ---- precomputed T (KG_precomp.py:23-26, B_precomp.py:18-20, AC_precompute.py:20-24) ------------------------------
__global__ void build_t_kernel(const float *__restrict__ stream, long long slot, int rt, int narr, int pde, int n,
                               const float2 *__restrict__ groups, float *__restrict__ tglob, long long t_stride)
{
    const int tile = blockIdx.x, g = blockIdx.y;
    const int nqt = n / 4 + 1, arr = 16 * rt, qs = 4 * rt;
    const float tp0 = groups[g].x, tp1 = groups[g].y;
    const float *src = stream + (long long)tile * slot;
    float *dst = tglob + (long long)g * t_stride + (long long)tile * nqt * qs;
    for (int e = threadIdx.x; e < nqt * rt; e += blockDim.x) {
        const int q = e / rt, r = e % rt, o = q * qs + r * 4;
        const float4 a = *reinterpret_cast<const float4 *>(src + o), b = *reinterpret_cast<const float4 *>(src + arr + o);
        float4 tv;
        if (pde == PDE_B) {
            tv.x = __fadd_rn(__fmul_rn(tp0, b.x), a.x); tv.y = __fadd_rn(__fmul_rn(tp0, b.y), a.y);
            tv.z = __fadd_rn(__fmul_rn(tp0, b.z), a.z); tv.w = __fadd_rn(__fmul_rn(tp0, b.w), a.w);
        } else {
            const float4 d = *reinterpret_cast<const float4 *>(src + 2 * arr + o);
            tv.x = __fadd_rn(__fadd_rn(a.x, __fmul_rn(tp0, b.x)), __fmul_rn(tp1, d.x));
            tv.y = __fadd_rn(__fadd_rn(a.y, __fmul_rn(tp0, b.y)), __fmul_rn(tp1, d.y));
            tv.z = __fadd_rn(__fadd_rn(a.z, __fmul_rn(tp0, b.z)), __fmul_rn(tp1, d.z));
            tv.w = __fadd_rn(__fadd_rn(a.w, __fmul_rn(tp0, b.w)), __fmul_rn(tp1, d.w));
        }
        if (q == n / 4) {                       // forcing column (x cos t - x^2 cos^2 t for KG; f_hat == 0 on this path)
            const float fc = pde == PDE_KG ? src[(long long)narr * arr + r] : 0.f;
            const int cpt = n % 4;
            if (cpt == 0) tv.x = fc; else if (cpt == 1) tv.y = fc; else if (cpt == 2) tv.z = fc; else tv.w = fc;
        }
        *reinterpret_cast<float4 *>(dst + o) = tv;
    }
}

cudaError_t launch_build_t(const float *stream, long long slot, int rt, int narr, int pde, int n, int ntile0, const float2 *groups_dev,
                           int ngroups, float *tglob, long long t_stride, cudaStream_t st)
{
    if (ngroups <= 0 || ntile0 <= 0) return cudaSuccess;
    for (int g0 = 0; g0 < ngroups; g0 += 65535) {           // gridDim.y limit
        const int ng = ngroups - g0 < 65535 ? ngroups - g0 : 65535;
        build_t_kernel<<<dim3((unsigned)ntile0, (unsigned)ng), 256, 0, st>>>(stream, slot, rt, narr, pde, n, groups_dev + g0,
                                                                              tglob + (long long)g0 * t_stride, t_stride);
    }
    return cudaGetLastError();
}

// ---- exact arg-max scan ---------------------------------------------------
// Sequential semantics: L = 0, idx = -1; for p in grid order: if !(m_p < L) && f_p > L: L = f_p, idx = p.
//
// Fast path (parallel).  Let w be the FIRST index of F = max f (NaN never wins) and X = max(0, max_{p<w} f_p).
// Every earlier f_p is strictly below F and the running maximum when the scan reaches w is at most X, so
// !(m_w < X) implies that w is accepted there, and nothing after w can exceed F: the answer is (F, w).  If all
// f_p <= 0 nothing is ever accepted: (0, -1).  Only when m_w < X (a winner whose loss once dipped below an earlier
// final loss) the answer depends on the order of acceptance, and the block falls back to the sequential scan.
//
// Fallback (sequential semantics, parallel filter): L only grows, so a 1024-entry chunk in which no entry
// satisfies  !(m < L) && f > L  for the L at chunk entry cannot change L and is skipped.
__device__ __forceinline__ void best_of(float &bf, int &bi, float of, int oi)
{
    if (of > bf || (of == bf && oi < bi)) { bf = of; bi = oi; }
}

__global__ void argmax_scan_kernel(const float *__restrict__ f, const float *__restrict__ m, int Np,
                                   float *L_out, int *idx_out)
{
    __shared__ float sL, sF[32], sX[32];
    __shared__ int sIdx, sI[32], sW, sDone;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const float NEG = __int_as_float(0xff800000);
    // ---- fast path, step 1: first arg-max of f
    float bf = NEG; int bi = 0x7fffffff;
    for (int p = threadIdx.x; p < Np; p += blockDim.x) { const float v = f[p]; if (v > bf) { bf = v; bi = p; } }
    for (int off = 16; off; off >>= 1) best_of(bf, bi, __shfl_xor_sync(0xffffffffu, bf, off), __shfl_xor_sync(0xffffffffu, bi, off));
    if (lane == 0) { sF[warp] = bf; sI[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
        bf = lane < nwarp ? sF[lane] : NEG; bi = lane < nwarp ? sI[lane] : 0x7fffffff;
        for (int off = 16; off; off >>= 1) best_of(bf, bi, __shfl_xor_sync(0xffffffffu, bf, off), __shfl_xor_sync(0xffffffffu, bi, off));
        if (lane == 0) { sF[0] = bf; sW = bi; }
    }
    __syncthreads();
    const float F = sF[0];
    const int w = sW;
    // ---- step 2: X = max(0, max_{p < w} f_p)
    float bx = 0.f;
    if (F > 0.f)
        for (int p = threadIdx.x; p < w; p += blockDim.x) { const float v = f[p]; if (v > bx) bx = v; }
    for (int off = 16; off; off >>= 1) bx = fmaxf(bx, __shfl_xor_sync(0xffffffffu, bx, off));
    if (lane == 0) sX[warp] = bx;
    __syncthreads();
    if (threadIdx.x == 0) {
        float X = 0.f;
        for (int k = 0; k < nwarp; ++k) X = fmaxf(X, sX[k]);
        sDone = 0;
        if (!(F > 0.f)) { *L_out = 0.f; *idx_out = -1; sDone = 1; }
        else if (!(m[w] < X)) { *L_out = F; *idx_out = w; sDone = 1; }
        sL = 0.f; sIdx = -1;
    }
    __syncthreads();
    if (sDone) return;
    // ---- fallback: the literal scan with a parallel chunk filter
    for (int base = 0; base < Np; base += blockDim.x) {
        const int p = base + threadIdx.x;
        const float L = sL;
        bool cand = false;
        if (p < Np) cand = !(m[p] < L) && (f[p] > L);
        if (__syncthreads_or(cand)) {
            if (threadIdx.x == 0) {
                float Lc = sL; int ic = sIdx;
                const int end = min(base + (int)blockDim.x, Np);
                for (int q = base; q < end; ++q)
                    if (!(m[q] < Lc) && (f[q] > Lc)) { Lc = f[q]; ic = q; }
                sL = Lc; sIdx = ic;
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) { *L_out = sL; *idx_out = sIdx; }
}

cudaError_t launch_argmax_scan(const float *f, const float *m, int Np, float *L_out, int *idx_out, cudaStream_t st)
{
    argmax_scan_kernel<<<1, 1024, 0, st>>>(f, m, Np, L_out, idx_out);
    return cudaGetLastError();
}

}  // namespace gp
