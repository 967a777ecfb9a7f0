/*
 * gptpinn_oracle.c -- CPU restatement of the GPT-PINN offline hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (gpt-pinn_b200/) may link, import or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg do.
 *
 * Parity pin: every function below is checked (tests/test_oracle_golden.py) against fixtures
 * under tests/golden/ that were produced by running the UNMODIFIED reference on CPU
 * (oracle/gen_golden.py).  The reference itself ships no tests / golden vectors.
 *
 * Each PDE has two builds of the same code: REAL=float ("_f32": fp32 storage and elementwise
 * arithmetic exactly where the reference rounds, row sums accumulated in double then rounded,
 * which is what ATen's cascade sum approximates) and REAL=double ("_f64": the adjudicator).
 * Scalars (alpha, beta, gamma, 2*gamma, nu, -nu, lambda, eps, 3*eps, 2/N, lr) are Python
 * doubles in the reference and are rounded to fp32 when they meet an fp32 tensor
 * (SURVEY.md N5); both builds use those rounded values.
 *
 * Reference lines followed:
 *   KG : update  Klein-Gordon/KG_optimizer.py:34-68   loss Klein-Gordon/KG_models.py:28-50
 *        T, G2   Klein-Gordon/KG_precomp.py:23-29      loop Klein-Gordon/KG_train.py:12-43
 *   B  : update  Burgers/B_optimizer.py:28-58          loss Burgers/B_models.py:28-46
 *        T       Burgers/B_precomp.py:18-20             loop Burgers/B_train.py:49-70
 *   AC : update  Allen-Cahn/AC_optimizer.py:36-66      loss Allen-Cahn/AC_models.py:31-54
 *        T       Allen-Cahn/AC_precompute.py:20-24      loop Allen-Cahn/AC_train.py:11-40
 *   MLP value/derivative channels: KG_precomp.py:7-21, B_precomp.py:6-16,
 *        AC_precompute.py:7-18 with the mixed-derivative semantics of SURVEY.md N2.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXN 32

typedef struct {
    int n, ld;                 /* neurons used, row stride (floats) of every P array        */
    int NR, NI, NB;
    const float *P, *Pxx, *Ptt;    /* [NR, ld]                                              */
    const float *PIC, *PICt;       /* [NI, ld]                                              */
    const float *PBC;              /* [NB, ld]                                              */
    const float *xcos, *fhat;      /* [NR]                                                  */
    const float *ICu1, *ICu2;      /* [NI]                                                  */
    const float *BCu;              /* [NB]                                                  */
} kg_problem_t;

typedef struct {
    int n, ld;
    int NR, NI, NB;
    const float *P, *Px, *Pt, *Pxx;  /* [NR, ld] */
    const float *PIC;                /* [NI, ld] */
    const float *PBC;                /* [NB, ld] */
    const float *fhat;               /* [NR]     */
    const float *ICu;                /* [NI]     */
    const float *BCu;                /* [NB]     */
} b_problem_t;

typedef struct {
    int n, ld;
    int NR, NI, NB;
    const float *P, *Pt, *Pxx;                 /* [NR, ld] */
    const float *PIC;                          /* [NI, ld] */
    const float *Pub, *Plb, *Pdiff;            /* [NB, ld] */
    const float *Pubx, *Plbx, *Pdiffx;         /* [NB, ld] */
    const float *fhat;                         /* [NR]     */
    const float *ICu;                          /* [NI]     */
} ac_problem_t;

/* ---- the two precisions ------------------------------------------------------------- */
#define REAL float
#define SUF(x) x##_f32
#include "gptpinn_oracle_impl.h"
#undef REAL
#undef SUF
#define REAL double
#define SUF(x) x##_f64
#include "gptpinn_oracle_impl.h"
#undef REAL
#undef SUF

/* ---- exact arg-max semantics of offline_generation from two scalars per parameter ---
 * SURVEY.md N1 (derived from Klein-Gordon/KG_train.py:27-41): parameter p becomes the new
 * maximum iff  not (m_p < L)  and  f_p > L, where m_p = min_{0<=j<E} loss_j, f_p = loss_E,
 * L = running maximum (starts at integer 0).  First index wins ties; NaN never wins.       */
int oracle_argmax_scan(const float *f, const float *m, int Np, float *L_out, int *idx_out)
{
    float L = 0.0f; int idx = -1;
    for (int p = 0; p < Np; ++p) {
        if (!(m[p] < L) && (f[p] > L)) { L = f[p]; idx = p; }
    }
    *L_out = L; *idx_out = idx;
    return 0;
}

/* ---- forward-mode MLP value + derivative channels (closed form of the autograd calls) --
 * channels: v, x, t, q = "P_xx" = u_xx+u_xt, r = "P_tt" = u_tt+u_xt.
 * act: 0 = cos (KG_models.py:56-60), 1 = tanh (B_models.py:69, AC_models.py:68).
 * W[l] is [layers[l+1], layers[l]] row-major (torch nn.Linear), b[l] is [layers[l+1]].
 * Outputs may be NULL.  Computed in double from fp32 weights, rounded on store.             */
int oracle_mlp_channels(const int *layers, int nlayers, const float *const *W, const float *const *b,
                        int act, const float *xt, int N,
                        float *ov, float *ox, float *ot, float *oq, float *orr)
{
    int maxw = 0;
    for (int l = 0; l < nlayers; ++l) if (layers[l] > maxw) maxw = layers[l];
    double *A = (double *)malloc(sizeof(double) * 10 * maxw);
    if (!A) return -1;
    double *av = A, *ax = A + maxw, *at = A + 2 * maxw, *aq = A + 3 * maxw, *ar = A + 4 * maxw;
    double *zv = A + 5 * maxw, *zx = A + 6 * maxw, *zt = A + 7 * maxw, *zq = A + 8 * maxw, *zr = A + 9 * maxw;
    for (int p = 0; p < N; ++p) {
        av[0] = xt[2 * p]; av[1] = xt[2 * p + 1];
        ax[0] = 1; ax[1] = 0; at[0] = 0; at[1] = 1; aq[0] = aq[1] = 0; ar[0] = ar[1] = 0;
        for (int l = 0; l < nlayers - 1; ++l) {
            int wi = layers[l], wo = layers[l + 1];
            for (int j = 0; j < wo; ++j) {
                double sv = b[l][j], sx = 0, st = 0, sq = 0, sr = 0;
                const float *w = W[l] + (size_t)j * wi;
                for (int i = 0; i < wi; ++i) {
                    sv += w[i] * av[i]; sx += w[i] * ax[i]; st += w[i] * at[i];
                    sq += w[i] * aq[i]; sr += w[i] * ar[i];
                }
                zv[j] = sv; zx[j] = sx; zt[j] = st; zq[j] = sq; zr[j] = sr;
            }
            if (l == nlayers - 2) {            /* last layer is linear */
                memcpy(av, zv, sizeof(double) * wo); memcpy(ax, zx, sizeof(double) * wo);
                memcpy(at, zt, sizeof(double) * wo); memcpy(aq, zq, sizeof(double) * wo);
                memcpy(ar, zr, sizeof(double) * wo);
            } else {
                for (int j = 0; j < wo; ++j) {
                    double s, d1, d2;
                    if (act == 0) { s = cos(zv[j]); d1 = -sin(zv[j]); d2 = -s; }
                    else { s = tanh(zv[j]); d1 = 1 - s * s; d2 = -2 * s * d1; }
                    av[j] = s;
                    ax[j] = d1 * zx[j];
                    at[j] = d1 * zt[j];
                    aq[j] = d2 * zx[j] * (zx[j] + zt[j]) + d1 * zq[j];
                    ar[j] = d2 * zt[j] * (zt[j] + zx[j]) + d1 * zr[j];
                }
            }
        }
        if (ov) ov[p] = (float)av[0];
        if (ox) ox[p] = (float)ax[0];
        if (ot) ot[p] = (float)at[0];
        if (oq) oq[p] = (float)aq[0];
        if (orr) orr[p] = (float)ar[0];
    }
    free(A);
    return 0;
}
