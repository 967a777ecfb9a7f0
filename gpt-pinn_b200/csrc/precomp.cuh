// precomp.cuh -- closed-form MLP value/derivative channels (replaces torch.autograd in inputs()).
#pragma once
#include <cuda_runtime.h>

namespace gp {

constexpr int MLP_MAX_LAYERS = 8;          // entries in `layers` (so <= 7 linear maps)

struct MlpDesc {
    int nlayers;
    int layers[MLP_MAX_LAYERS];
    const float *W[MLP_MAX_LAYERS - 1];    // device, [layers[l+1], layers[l]] row-major (torch nn.Linear.weight)
    const float *b[MLP_MAX_LAYERS - 1];    // device, [layers[l+1]]
    int act;                               // 0 = cos, 1 = tanh
};
struct MlpOut {
    float *ptr[5];                         // v, x, t, q, r  (nullptr = not wanted)
    long long ld[5];                       // element stride between consecutive points
};

cudaError_t launch_mlp_channels(const MlpDesc &d, const float *xt, long long N, const MlpOut &o, cudaStream_t st);

}  // namespace gp
