// pack.cuh -- host-visible descriptors of the pack and scan kernels.
#pragma once
#include "common.cuh"

namespace gp {

struct PackMat {
    const float *src;      // device, row-major, row stride ld
    long long    ld;
    long long    nrows;
    int          ncols;    // n
    int          arr;      // matrix block index inside the tile slot
    long long    tile0;    // first tile of the segment this matrix belongs to
};
struct PackVec {
    const float *src;      // device [nrows] or nullptr (leave zeros)
    long long    nrows;
    int          vec;      // row-vector index inside the tile slot
    long long    tile0;
};
struct PackJobs {
    PackMat mat[12];
    PackVec vec[8];
    int     nmat, nvec;
    int     narr;          // matrix blocks per slot
    int     rt;            // rows per tile
    long long slot;        // floats per slot
};

cudaError_t launch_pack(const PackJobs &jobs, float *stream, long long max_rows, cudaStream_t st);
// T arrays of the (alpha, beta) groups, once per sweep launch: tglob[g][residual tile][n/4+1 quads][rt rows][4], built from the
// packed stream with the reference's rounding (the same expression the sweep kernel uses when it builds T itself)
cudaError_t launch_build_t(const float *stream, long long slot, int rt, int narr, int pde, int n, int ntile0, const float2 *groups_dev,
                           int ngroups, float *tglob, long long t_stride, cudaStream_t st);
cudaError_t launch_argmax_scan(const float *f, const float *m, int Np, float *L_out, int *idx_out, cudaStream_t st);

}  // namespace gp
