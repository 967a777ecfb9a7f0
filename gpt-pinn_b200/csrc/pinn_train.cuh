// pinn_train.cuh -- descriptor of the fused full-PINN training kernel (pinn_train.cu).
#pragma once
#include <cuda_runtime.h>

namespace gp {

constexpr int PINN_MAX_HIDDEN = 6;            // hidden layers (KG: 2, Burgers: 4)

struct PinnDesc {
    int nlayers;                              // entries of layers[] (inputs, hidden..., output)
    int layers[PINN_MAX_HIDDEN + 2];
    int act;                                  // 0 cos, 1 tanh
    int pde;                                  // 0 Klein-Gordon, 1 Burgers
    float pa, pb, pc;                         // KG: alpha, beta, gamma      Burgers: -nu, 0, 0
    int n_params;                             // total weights + biases
    float *params;                            // device [n_params]: per layer W (row-major [out][in]) then b; updated in place
    float *adam_m, *adam_v;                   // device [n_params]
    float *grad_out;                          // device [n_params] or nullptr: gradient of the last executed step
    int NR, NI, NB;
    const float *xt_resid, *IC_xt, *BC_xt;    // [N,2]
    const float *f_hat, *xcos;                // [NR] (f_hat may be null; xcos KG only)
    const float *IC_u1, *IC_u2, *BC_u;        // [NI], [NI] (KG only, may be null), [NB]
    float invR, invI, invB;
    int epochs;
    float lr, beta1, beta2, omb1, omb2, eps;  // omb = (float)(1 - beta) as torch forms it in double
    double beta1_d, beta2_d, lr_d;            // the Python doubles torch.optim.Adam forms its bias corrections from
    float tol;                                // < 0: no tolerance stop
    float *partials; int n_partials; int partial_stride;   // [n_partials][partial_stride >= n_params + 4]
    float *status;                            // [2]: last evaluated loss, epochs entered (raw int32 bits)
    int *stop;                                // [1]
    float *loss_trace; int trace_every;       // optional: loss before the step of epochs 1, 1+k, ...
};

// Several independent trainings (same network shape, point sets and schedule; own weights and PDE parameters) can
// share one cooperative launch: the grid is cut into `n` equal groups of CTAs, one training per group, and the two
// grid-wide barriers of an epoch are shared.  One training leaves the last tile round of an epoch mostly empty
// (361 tiles on 296 resident CTAs at the KG sizes); a batch fills it.
constexpr int PINN_MAX_BATCH = 8;
struct PinnTrainee {
    float pa, pb, pc;
    float *params, *adam_m, *adam_v, *grad_out;
    float *partials;                          // [group CTAs][partial_stride]
    float *status;                            // [2]
    float *loss_trace;
};
struct PinnBatch {
    int n;                                    // trainings in this launch (1..PINN_MAX_BATCH)
    int group;                                // CTAs per training
    PinnTrainee t[PINN_MAX_BATCH];
};

// blocks = CTAs of the whole launch; group = CTAs per training (= partial rows per training)
cudaError_t plan_pinn_train(const PinnDesc &d, int n_sm, int nbatch, int *blocks, size_t *smem, int *group);
cudaError_t launch_pinn_train(const PinnDesc &d, const PinnBatch &b, int blocks, size_t smem, cudaStream_t st);

}  // namespace gp
