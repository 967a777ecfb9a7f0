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

// ---- exact arg-max scan ---------------------------------------------------
// Sequential semantics, parallel filter: L only grows, so a 1024-entry chunk in which no entry
// satisfies  !(m < L) && f > L  for the L at chunk entry cannot change L and is skipped.
__global__ void argmax_scan_kernel(const float *__restrict__ f, const float *__restrict__ m, int Np,
                                   float *L_out, int *idx_out)
{
    __shared__ float sL;
    __shared__ int sIdx;
    if (threadIdx.x == 0) { sL = 0.f; sIdx = -1; }
    __syncthreads();
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
