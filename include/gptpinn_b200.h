/*
 * gptpinn_b200.h -- C ABI of libgptpinn_b200.so: the B200 (sm_100a) implementation of the
 * GPT-PINN data-parallel hot path (greedy offline coefficient sweep, its feeder precompute
 * and the grid arg-max).
 *
 * The reference (skoohy/GPT-PINN) is pure Python on PyTorch and has no FFI layer; its
 * boundary for this path is the set of Python functions that *_main.py imports.  Each entry
 * point below names the reference interface it replaces (file:line relative to the
 * reference root).  The Python drop-in modules (gpt-pinn_b200/dropin/KG_train.py ...) bind
 * these symbols with ctypes; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - all `*_dev` / gptpinn_mat pointers are DEVICE pointers on the current CUDA device,
 *     fp32; all `*_host` pointers are host pointers; nothing is retained after return.
 *   - matrices are row-major with an explicit row stride `ld` (floats): the reference hands
 *     over column-slice views `out[:, 0:n]` of [N, number_of_neurons] buffers
 *     (Klein-Gordon/KG_precomp.py:51-52), so ld != n in general.  Vectors ([N,1] tensors in
 *     the reference) are contiguous.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  Calls enqueue
 *     work and return; they synchronise only where stated.
 *   - a handle owns one workspace (packed tile stream, work-item tables, trainer state): calls on the same handle
 *     must be issued from one host thread at a time and on one stream (or on streams the caller orders); use one
 *     handle per stream / per device for concurrent work.  The handle belongs to the device current at creation.
 *   - every function returns GPTPINN_OK (0) or a negative GPTPINN_E_* code and never throws;
 *     gptpinn_last_error() gives the message of the last failure on the calling thread.
 *   - scalars that the reference multiplies into fp32 tensors (alpha, beta, gamma, 2*gamma,
 *     nu, lambda, eps, 3*eps, 2/N, lr) are taken as double and rounded to fp32 at the same
 *     points the reference rounds them.
 */
#ifndef GPTPINN_B200_H
#define GPTPINN_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define GPTPINN_OK              0
#define GPTPINN_E_INVALID      -1   /* bad argument (null pointer, n out of range, ...)        */
#define GPTPINN_E_CUDA         -2   /* a CUDA runtime call failed                               */
#define GPTPINN_E_UNSUPPORTED  -3   /* device is not sm_100 / configuration not compiled        */
#define GPTPINN_E_NOMEM        -4

#define GPTPINN_MAX_NEURONS    15   /* KG uses 15, Burgers / Allen-Cahn 9                       */

typedef struct gptpinn_handle gptpinn_handle;   /* per-device workspace owner (opaque)        */

typedef struct {
    const float *ptr;      /* device, fp32                                                     */
    long long    ld;       /* row stride in floats                                             */
} gptpinn_mat;

/* Klein-Gordon reduced problem: everything KG_optimizer.grad_descent.__init__
 * (Klein-Gordon/KG_optimizer.py:4-32) and KG_models.GPT.__init__ (KG_models.py:8-26) keep.    */
typedef struct {
    int n;                               /* neurons in use, 1..15                              */
    int NR, NI, NB;                      /* xt_size, IC_size, BC_size                          */
    gptpinn_mat P, Pxx, Ptt;             /* out, out_xx, out_tt           [NR, n]              */
    gptpinn_mat PIC, PICt;               /* out_IC, out_IC_t              [NI, n]              */
    gptpinn_mat PBC;                     /* out_BC                        [NB, n]              */
    const float *xcos;                   /* xcos_x2cos2                   [NR]                 */
    const float *fhat;                   /* f_hat [NR], or NULL == all zero                    */
    const float *ICu1, *ICu2;            /* IC_u1, IC_u2 [NI]; ICu2 NULL == all zero           */
    const float *BCu;                    /* BC_u  [NB]                                         */
} gptpinn_kg_problem;

/* Burgers: B_optimizer.py:4-26, B_models.py:12-26                                             */
typedef struct {
    int n;
    int NR, NI, NB;
    gptpinn_mat P, Px, Pt, Pxx;          /* out, out_x, out_t, out_xx     [NR, n]              */
    gptpinn_mat PIC;                     /* out_IC                        [NI, n]              */
    gptpinn_mat PBC;                     /* out_BC                        [NB, n]              */
    const float *fhat;                   /* [NR] or NULL                                       */
    const float *ICu;                    /* [NI]                                               */
    const float *BCu;                    /* [NB] or NULL (used by the loss only, N3)           */
} gptpinn_b_problem;

/* Allen-Cahn: AC_optimizer.py:4-34, AC_models.py:13-29                                        */
typedef struct {
    int n;
    int NR, NI, NB;
    gptpinn_mat P, Pt, Pxx;              /* out, out_t, out_xx            [NR, n]              */
    gptpinn_mat PIC;                     /* out_IC                        [NI, n]              */
    gptpinn_mat Pub, Plb, Pdiff;         /* out_BC_ub/lb/diff             [NB, n]              */
    gptpinn_mat Pubx, Plbx, Pdiffx;      /* out_BC_ub_x/lb_x/diff_x       [NB, n]              */
    const float *fhat;                   /* [NR] or NULL                                       */
    const float *ICu;                    /* [NI]                                               */
    int t_given;                         /* != 0: `Pt` already holds the materialised operator
                                          * T = P_t - lambda P_xx - eps P (the Pt_lPxx_eP_term that
                                          * AC_optimizer.grad_descent / AC_models.GPT receive,
                                          * AC_optimizer.py:4-20): it is used as is (bit-exact), `Pxx`
                                          * is ignored and only eps of each grid row is used          */
} gptpinn_ac_problem;

/* What one sweep call does for each of the Np grid rows (parameters):
 *     c <- c0;  for j = 0..epochs:  loss_j = GPT.loss(c);  if j < epochs: c <- update(c)
 * i.e. the loop of offline_generation (Klein-Gordon/KG_train.py:27-37, Burgers/B_train.py:60-67,
 * Allen-Cahn/AC_train.py:24-31) without its early exit, which is also the loop of gpt_test /
 * gpt_test_loss (KG_test.py:30-32, 67-72).  Per parameter it returns
 *     c_out[p,:]   = c after `epochs` updates                           (may be NULL)
 *     f_out[p]     = loss_epochs                                        (the "final loss")
 *     m_out[p]     = min_{0 <= j < epochs} loss_j                       (early-exit witness)
 *     trace[p, j/rec_every] = loss_j for j % rec_every == 0             (NULL or rec_every<=0: off)
 * from which gptpinn_argmax_scan reproduces offline_generation's (largest_loss, largest_case)
 * exactly.                                                                     */
typedef struct {
    const double *grid_host;   /* [Np, 3] (KG: alpha,beta,gamma) / [Np] (B: nu) / [Np,2] (AC: lambda,eps) */
    int           Np;
    const float  *c0_dev;      /* [n] shared by all parameters, or [Np, n] when c0_per_param   */
    int           c0_per_param;
    int           epochs;
    double        lr;
    int           rec_every;
    float        *c_out_dev;   /* [Np, n] or NULL                                              */
    float        *f_out_dev;   /* [Np]                                                         */
    float        *m_out_dev;   /* [Np]                                                         */
    float        *trace_dev;   /* [Np, epochs/rec_every + 1] or NULL                           */
    /* tuning override, 0 = let the library choose: parameters per thread (M), sibling lanes
     * per warp (SL, power of two <= 32), warps per CTA (W).                                   */
    int           cfg_M, cfg_SL, cfg_W;
    int           cfg_mode;    /* 0 = auto, 1 = shared-ring kernel, 2 = private-ring kernel, 3 = row-partitioned kernel,
                                * grid flavour (few parameters: rows split over all CTAs of a cooperative grid, every CTA
                                * processes every parameter, per-epoch reduction through global memory; needs Np <= 1024,
                                * no bound_dev; cfg_W > 0 caps the residual rows per streaming chunk), 4 = row-partitioned
                                * kernel, cluster flavour (parameters split over thread-block clusters, rows over the CTAs
                                * of a cluster, per-epoch reduction through distributed shared memory; row shares must fit
                                * in shared memory, no bound_dev; cfg_K = CTAs per cluster, 0 = best)                     */
    int           cfg_K;       /* 0 = auto; private-ring kernel: warps that team up on one work item
                                * (1, 2, 4 or 8): they split the tiles of a pass and sum their gradients;
                                * cfg_mode 4: CTAs per cluster (1, 2, 4, 8 or 16)                           */
    /* Optional early exit with the reference's semantics (KG_train.py:31-33 `if loss < largest_loss: break`):
     * bound_dev[p] must be <= the running maximum the reference's sequential loop holds when it reaches grid
     * row p (e.g. the max of min(f, m) over already finished rows p' < p).  Row p then stops at the
     * first j < epochs with loss_j < bound_dev[p] and reports f_out[p] = m_out[p] = loss_j, which
     * gptpinn_argmax_scan rejects exactly like the reference's break; rows that never fall below their bound
     * run all epochs and report exact values.  NULL = no early exit (fixed work).  Forces the private-ring kernel. */
    const float  *bound_dev;   /* [Np] or NULL                                                  */
    /* Optional fused arg-max epilogue for sharded grids: the kernel atomically maximises, over the rows of
     * this call, the 64-bit key (float_bits(f_out[p]) << 32) | (0xFFFFFFFF - key_index_dev[p]) into *key_out_dev (which
     * the caller zeroes; rows with NaN or non-positive final loss do not enter).  The maximum over all ranks (one 8-byte
     * MAX all-reduce) is the first arg-max of the final losses = the reference's selection whenever the winner's
     * min-prefix loss is not below the largest earlier final loss (N1 guard, checked by the caller).                      */
    unsigned long long *key_out_dev;   /* [1] or NULL                                            */
    const int    *key_index_dev;       /* [Np] global grid index of each row, or NULL (= row number) */
} gptpinn_sweep_args;

/* filled by the sweep calls when non-NULL: what was launched (for bench.py / tests)           */
typedef struct {
    int   M, SL, W;            /* chosen configuration                                         */
    int   items;               /* warp work items                                              */
    int   groups;              /* distinct T-parameter groups found in the grid                */
    int   ctas, rounds;
    int   kernels_launched;    /* sweep + pack kernels enqueued by this call                   */
    int   smem_bytes;
    int   tiles_per_pass;
    int   priv;                /* 0 = shared ring, 1 = private-ring kernel (dynamic tickets), 2 = row-partitioned kernel
                                * (grid flavour), 3 = row-partitioned kernel (cluster flavour)    */
    int   K;                   /* warps per team (private-ring kernel); row slices per CTA (priv 2); CTAs per cluster (priv 3) */
} gptpinn_sweep_info;

int  gptpinn_version(void);
const char *gptpinn_last_error(void);

/* handle = workspace owner for one device (packed activation mirror, work-item tables)        */
int  gptpinn_create(gptpinn_handle **out);
int  gptpinn_destroy(gptpinn_handle *h);

/* replaces the per-parameter body of offline_generation / gpt_test / gpt_test_loss:
 * KG_train.py:14-37 + KG_optimizer.py:34-68 + KG_models.py:28-50 + KG_precomp.py:23-29        */
int  gptpinn_kg_sweep(gptpinn_handle *h, const gptpinn_kg_problem *prob,
                      const gptpinn_sweep_args *args, gptpinn_sweep_info *info, void *stream);
/* B_train.py:49-67 + B_optimizer.py:28-58 + B_models.py:28-46 + B_precomp.py:18-20            */
int  gptpinn_b_sweep(gptpinn_handle *h, const gptpinn_b_problem *prob,
                     const gptpinn_sweep_args *args, gptpinn_sweep_info *info, void *stream);
/* AC_train.py:12-31 + AC_optimizer.py:36-66 + AC_models.py:31-54 + AC_precompute.py:20-24     */
int  gptpinn_ac_sweep(gptpinn_handle *h, const gptpinn_ac_problem *prob,
                      const gptpinn_sweep_args *args, gptpinn_sweep_info *info, void *stream);

/* Exact arg-max of offline_generation (KG_train.py:12,31-41; B_train.py:18,62-70;
 * AC_train.py:10,25-38) from the two per-parameter scalars: visits p = 0..Np-1 in order with
 * running maximum L (initially 0) and sets (L, idx) <- (f[p], p) iff !(m[p] < L) && f[p] > L.
 * Writes L to L_out_dev[0] and the winning index (-1 if none) to idx_out_dev[0].              */
int  gptpinn_argmax_scan(const float *f_dev, const float *m_dev, int Np,
                         float *L_out_dev, int *idx_out_dev, void *stream);

/* Pre-trained full-PINN activation network (reference classes KG_models.NN, B_models.NN,
 * AC_models.P): layers[0] = 2 inputs (x,t), tanh or cos hidden activations, linear output.    */
#define GPTPINN_ACT_COS   0     /* Klein-Gordon/KG_models.py:56-60                              */
#define GPTPINN_ACT_TANH  1     /* Burgers/B_models.py:69, Allen-Cahn/AC_models.py:68           */
typedef struct {
    int          nlayers;       /* entries of `layers`, <= 8                                     */
    int          layers[8];
    const float *W[7];          /* device, nn.Linear.weight  [layers[l+1], layers[l]] row-major  */
    const float *b[7];          /* device, nn.Linear.bias    [layers[l+1]]                       */
    int          act;
} gptpinn_mlp;

/* Closed-form replacement of the autograd calls in inputs() (KG_precomp.py:7-21 and 34-52,
 * B_precomp.py:6-16, AC_precompute.py:7-18): evaluates, at N points xt_dev[N,2], the channels
 *   out[0] = u, out[1] = u_x, out[2] = u_t, out[3] = "P_xx" = u_xx + u_xt, out[4] = "P_tt" =
 *   u_tt + u_xt  (the mixed-derivative semantics of the reference)
 * and writes channel c of point p to out_dev[c][p * ld[c]]; out_dev[c] == NULL skips it.  With
 * ld = number_of_neurons and out_dev[c] = &buffer[0][i] this fills column i in place.          */
int  gptpinn_mlp_channels(const gptpinn_mlp *mlp, const float *xt_dev, long long N,
                          float *const out_dev[5], const long long ld[5], void *stream);

/* ---- fused full-PINN training ------------------------------------------------------------------
 * Replaces the Adam loops of pinn_train (Klein-Gordon/KG_train.py:48-59 on KG_models.NN.loss,
 * KG_models.py:92-129; Burgers/B_train.py:76-90 on B_models.NN.loss, B_models.py:79-105): all epochs run
 * in one persistent cooperative kernel (forward-mode derivative channels, PDE residual, hand-derived
 * reverse pass, deterministic gradient reduction, torch.optim.Adam update with default betas/eps).  */
#define GPTPINN_PDE_KG       0
#define GPTPINN_PDE_BURGERS  1
typedef struct {
    int          pde;
    double       p0, p1, p2;        /* KG: alpha, beta, gamma            Burgers: nu                 */
    int          NR, NI, NB;
    const float *xt_resid;          /* [NR, 2] device, contiguous                                     */
    const float *f_hat;             /* [NR] or NULL (zero)                                            */
    const float *xcos;              /* [NR]  KG forcing xcos_x2cos2; NULL for Burgers                 */
    const float *IC_xt;             /* [NI, 2]                                                        */
    const float *IC_u1;             /* [NI]  KG IC_u1 / Burgers IC_u                                  */
    const float *IC_u2;             /* [NI]  KG IC_u2 or NULL (zero); ignored for Burgers             */
    const float *BC_xt;             /* [NB, 2]                                                        */
    const float *BC_u;              /* [NB]                                                           */
} gptpinn_pinn_problem;

typedef struct {
    int    epochs;
    double lr, beta1, beta2, eps;   /* torch.optim.Adam defaults: 0.9, 0.999, 1e-8                    */
    double tol;                     /* < 0: none; else stop BEFORE the step once loss < tol (B_train.py:84) */
    int    trace_every;             /* > 0: loss_trace_dev[k] = loss before the step of epoch 1 + k*trace_every */
    float *loss_trace_dev;          /* [(epochs-1)/trace_every + 1] or NULL                           */
    float *grad_out_dev;            /* [n_params] or NULL: gradient of the last executed step, packed
                                       per layer as weight (row-major [out][in]) then bias             */
} gptpinn_train_args;

/* `net->W[l]` / `net->b[l]` are updated IN PLACE (contiguous fp32 device tensors).  status_dev[0] = the last
 * evaluated loss (what the reference prints as "PINN Final Loss"), status_dev[1] = epochs entered, stored as the raw bits of an int32 (read it with memcpy / *(int *)&status[1]). */
int  gptpinn_pinn_train(gptpinn_handle *h, const gptpinn_pinn_problem *prob, const gptpinn_mlp *net,
                        const gptpinn_train_args *args, float *status_dev, void *stream);

/* `nbatch` (1..8) independent trainings in ONE launch -- the per-test-parameter loop of pinn_test / pinn_test_loss
 * (KG_test.py:81-140, B_test.py:116-181).  probs[t] may differ only in the PDE parameters p0..p2, nets[t] only in
 * the weight tensors, args[t] only in loss_trace_dev / grad_out_dev; everything else must be equal (checked).
 * status_dev is [nbatch][2].  A training that meets `tol` stops stepping; the launch ends when all have.          */
int  gptpinn_pinn_train_batch(gptpinn_handle *h, int nbatch, const gptpinn_pinn_problem *probs, const gptpinn_mlp *nets,
                              const gptpinn_train_args *args, float *status_dev, void *stream);

/* ---- self-adaptive PINN for Allen-Cahn: device kernels of the TensorFlow-free trainer ------------------------------
 * Replaces the TensorFlow graph of Allen-Cahn/AC_main.py:204-339 (= AC_SA_test.py:109-245): neural_net / loss / f_model /
 * u_x_model / grad and the two Adam optimizers of fit().  The four channels (u, u_x, u_t, u_xx -- the TRUE second
 * x-derivative, AC_main.py:235-244) of a tanh MLP are propagated in closed form and the reverse pass is hand-derived; all
 * points (collocation NR, initial NI, lower boundary NB, upper boundary NB, in that order) run as one batch of
 * N = NR + NI + 2 NB points.  Activations, pre-activations and adjoints are [4][N][W] fp32 device tensors (channel-major,
 * W = hidden width), so a hidden->hidden map of all channels is ONE [4N x W] x [W x W] GEMM, which the host side
 * (gptpinn/sa_pinn.py) issues as a plain library GEMM; every other step is one of the calls below.                      */
/* Z[c][p][j]: pre-activations of the first hidden layer from the points xt[N][2] (W0 is nn.Linear layout [W][2])          */
int  gptpinn_sa_layer0(const float *xt_dev, long long N, const float *W0_dev, const float *b0_dev, int W, float *Z_dev, void *stream);
/* Z[0] += bias (NULL: none; stored back), A = tanh channels of Z                                                            */
int  gptpinn_sa_act_forward(float *Z_dev, const float *bias_dev, float *A_dev, long long N, int W, void *stream);
/* G: adjoint of the activations on entry, adjoint of the pre-activations on return (in place)                             */
int  gptpinn_sa_act_backward(const float *Z_dev, float *G_dev, long long N, int W, void *stream);
/* U[c][p] = A[c][p][:] . Wout (+ bout on channel 0)                                                                         */
int  gptpinn_sa_output_forward(const float *A_dev, const float *Wout_dev, const float *bout_dev, long long N, int W, float *U_dev, void *stream);
/* loss of AC_main.py:215-232 from U[4][N]: loss4 = (total, mse_u, mse_b, mse_f); Ubar = dL/dU; gwf / gwu = dL/d(col_weights),
 * dL/d(u_weights) (the ascent directions of AC_main.py:298).  workspace: gptpinn_sa_loss_workspace_floats() floats.          */
int  gptpinn_sa_loss_workspace_floats(void);
int  gptpinn_sa_loss(const float *U_dev, int NR, int NI, int NB, const float *u0_dev, const float *wf_dev, const float *wu_dev,
                     double lambda, double eps, float *Ubar_dev, float *gwf_dev, float *gwu_dev, float *loss4_dev,
                     float *workspace_dev, void *stream);
/* G[c][p][j] = Ubar[c][p] * Wout[j]: adjoint of the last hidden layer's activations                                         */
int  gptpinn_sa_output_backward(const float *Ubar_dev, const float *Wout_dev, long long N, int W, float *G_dev, void *stream);
/* *step_dev += 1 (the optimizer's iteration counter lives on the device so that a whole Adam iteration is one CUDA graph)  */
int  gptpinn_sa_tick(int *step_dev, void *stream);
/* tf.keras.optimizers.Adam step (TF 2.10 semantics, AC_main.py:282-283: lr 5e-3, beta_1 0.99, epsilon 1e-7) on a flat
 * vector; ascent != 0 applies it to -g (gradient ASCENT on the self-adaptive weights, AC_main.py:298).                      */
int  gptpinn_sa_adam(float *p_dev, float *m_dev, float *v_dev, const float *g_dev, long long n, const int *step_dev,
                     double lr, double beta1, double beta2, double eps, int ascent, void *stream);

/* Weight-gradient contraction of a hidden->hidden map over all channels and points (the tape's dW of AC_main.py:254-265):
 * out[j][i] = sum_r G[r][j] A[r][i], G / A row-major [R][W] (R = 4 N rows), out row-major [W][W]; W <= 128.  Row chunks per CTA,
 * partial products in workspace_dev (gptpinn_sa_workspace_floats() floats), summed in a fixed order.                                */
long long gptpinn_sa_workspace_floats(void);
int  gptpinn_sa_wgrad(const float *G_dev, const float *A_dev, long long R, int W, float *out_dev, float *workspace_dev, void *stream);
/* Bias gradient of a map: gb[c] = sum_r G[0][r][c] over the N points (G is [4][N][W]); W <= 256.                                     */
int  gptpinn_sa_bias_grad(const float *G_dev, long long N, int W, float *gb_dev, float *workspace_dev, void *stream);
/* Gradients of the first map (inputs (x, t) with channels (x,1,0,0), (t,0,1,0)): gb0[c] = sum_r G[0][r][c],
 * gW0[c][0] = sum_r x_r G[0][r][c] + G[1][r][c], gW0[c][1] = sum_r t_r G[0][r][c] + G[2][r][c]; xt_dev is [N][2]; W <= 256.           */
int  gptpinn_sa_layer0_grad(const float *G_dev, const float *xt_dev, long long N, int W, float *gW0_dev, float *gb0_dev,
                            float *workspace_dev, void *stream);

/* Hidden->hidden map of all four channels with the activation in its epilogue; tensors are [4][N][W], W_dev is the nn.Linear matrix
 * [out][in], W must be 128 (GPTPINN_E_INVALID otherwise: compose a GEMM with gptpinn_sa_act_forward / _backward), pointers 16-byte
 * aligned.  forward: Z_out = A W^T (+ bias on channel 0), A_out = channels of tanh(Z_out).  adjoint: G_out = act_backward(Zprev, G W)
 * (the adjoint of the previous layer's pre-activations; G_out must not alias G).                                                      */
int  gptpinn_sa_map_forward(const float *A_dev, const float *W_dev, const float *bias_dev, long long N, int W, float *Z_out_dev,
                            float *A_out_dev, void *stream);
int  gptpinn_sa_map_adjoint(const float *G_dev, const float *W_dev, const float *Zprev_dev, long long N, int W, float *G_out_dev, void *stream);

/* L-BFGS search direction d = -H g from the k most recent (s, y) pairs (the two-loop recursion of eager_lbfgs.py:80-106) in one
 * launch.  S_dev / Y_dev are rings of `hist` rows of `ld` floats: pair i (i = 0 oldest ... k-1 newest) sits in row (start + i) % hist,
 * ro_dev[row] = 1 / <y, s> of the pair in that row, Hdiag = <y, s> / <y, y> of the newest pair.  k = 0: d = -Hdiag g.  The vector must
 * fit one CTA's shared memory: P <= gptpinn_sa_lbfgs_max_params().  16-byte aligned rows (ld % 4 == 0) take the float4 path.      */
int  gptpinn_sa_lbfgs_max_params(void);
int  gptpinn_sa_lbfgs_direction(const float *S_dev, const float *Y_dev, const float *ro_dev, int k, int start, int hist, long long ld,
                                long long P, double Hdiag, const float *g_dev, float *d_dev, void *stream);

/* Host-only (no CUDA call): the work-item plan gptpinn_kg_sweep would build for this grid on a device
 * with n_sm SMs.  sl_hist[l] = number of items with 2^l sibling lanes; *covered = parameters assigned
 * (must equal Np).  Used by tests and for tuning; no reference counterpart.                            */
int  gptpinn_kg_plan(const double *grid_host, int Np, int n, int NR, int NI, int NB, int n_sm,
                     const gptpinn_sweep_args *cfg, gptpinn_sweep_info *info, int sl_hist[6], int *covered);

/* FFMA-only microbenchmark on the current device: best-of-`reps` FP32 CUDA-core throughput in
 * TFLOP/s (2 flop per FMA).  It is the measured denominator of the sweep kernel's roofline;
 * it has no counterpart in the reference.  Synchronises the stream.                            */
int  gptpinn_ffma_peak(double *tflops_out, int reps, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* GPTPINN_B200_H */
