// sweep_inst.cu -- one translation unit per (PDE, n): compiled with -DGP_PDE=<0|1|2> -DGP_N=<n>.
#include "sweep_kernel.cuh"
#include "sweep_rows.cuh"

#define GP_CAT2(a, b, c) a##b##_##c
#define GP_CAT(a, b, c) GP_CAT2(a, b, c)

namespace gp {
cudaError_t GP_CAT(sweep_launch_, GP_PDE, GP_N)(const SweepParams &p, int M, int priv, int ctas, size_t smem, cudaStream_t st)
{
    return launch_n<GP_PDE, GP_N>(p, M, priv, ctas, smem, st);
}
cudaError_t GP_CAT(sweep_rows_launch_, GP_PDE, GP_N)(const RowsParams &p, int grid, size_t smem, cudaStream_t st)
{
    return rows_launch<GP_PDE, GP_N>(p, grid, smem, st);
}
int GP_CAT(sweep_rows_clusters_, GP_PDE, GP_N)(int S, size_t smem)
{
    return rows_max_clusters<GP_PDE, GP_N>(S, smem);
}
}  // namespace gp
