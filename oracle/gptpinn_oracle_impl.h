/* gptpinn_oracle_impl.h -- included twice by gptpinn_oracle.c (REAL = float, double).
 * TEST INFRASTRUCTURE ONLY (see gptpinn_oracle.c header).                                   */

static inline REAL SUF(dotc)(const float *row, const REAL *c, int n)
{
    double s = 0.0;                       /* matmul: accumulate wide, round once            */
    for (int k = 0; k < n; ++k) s += (double)((REAL)row[k] * c[k]);
    return (REAL)s;
}
static inline REAL SUF(dotT)(const REAL *row, const REAL *c, int n)
{
    double s = 0.0;
    for (int k = 0; k < n; ++k) s += (double)(row[k] * c[k]);
    return (REAL)s;
}

/* =============================== Klein-Gordon ========================================== */
/* T = (P_tt + alpha*P_xx) + beta*P          KG_precomp.py:23-26                             */
static void SUF(kg_build_T)(const kg_problem_t *q, float alpha, float beta, REAL *T)
{
    for (int r = 0; r < q->NR; ++r)
        for (int k = 0; k < q->n; ++k) {
            size_t i = (size_t)r * q->ld + k;
            REAL t1 = (REAL)alpha * (REAL)q->Pxx[i];
            REAL t2 = (REAL)beta * (REAL)q->P[i];
            T[(size_t)r * q->n + k] = ((REAL)q->Ptt[i] + t1) + t2;
        }
}

/* loss(c) (KG_models.py:28-50) and, when grad != NULL, the update direction
 * (KG_optimizer.py:34-65; note IC_u2 and f_hat are absent from it -- SURVEY.md N3).          */
static REAL SUF(kg_eval)(const kg_problem_t *q, float gamma, float gamma2, const REAL *T,
                         const REAL *c, REAL *grad)
{
    const int n = q->n;
    const REAL fac1 = (REAL)(float)(2.0 / q->NR), fac2 = (REAL)(float)(2.0 / q->NB),
               fac3 = (REAL)(float)(2.0 / q->NI);
    double gR[MAXN] = {0}, gB[MAXN] = {0}, gI1[MAXN] = {0}, gI2[MAXN] = {0};
    double lR = 0, lB = 0, lI1 = 0, lI2 = 0;
    for (int r = 0; r < q->NR; ++r) {
        const float *Pr = q->P + (size_t)r * q->ld;
        const REAL *Tr = T + (size_t)r * n;
        REAL u = SUF(dotc)(Pr, c, n);
        REAL term1 = SUF(dotT)(Tr, c, n);
        REAL term2 = term1 + (REAL)gamma * (u * u);
        REAL first = term2 + (REAL)q->xcos[r];
        REAL d = first - (REAL)q->fhat[r];
        lR += (double)(d * d);
        if (grad)
            for (int k = 0; k < n; ++k) {
                REAL term3 = ((REAL)gamma2 * (REAL)Pr[k]) * u;
                REAL second = Tr[k] + term3;
                gR[k] += (double)(first * second);
            }
    }
    for (int r = 0; r < q->NB; ++r) {
        const float *Br = q->PBC + (size_t)r * q->ld;
        REAL bt = SUF(dotc)(Br, c, n) - (REAL)q->BCu[r];
        lB += (double)(bt * bt);
        if (grad) for (int k = 0; k < n; ++k) gB[k] += (double)(bt * (REAL)Br[k]);
    }
    for (int r = 0; r < q->NI; ++r) {
        const float *Ir = q->PIC + (size_t)r * q->ld;
        const float *Jr = q->PICt + (size_t)r * q->ld;
        REAL i1 = SUF(dotc)(Ir, c, n) - (REAL)q->ICu1[r];
        REAL i2 = SUF(dotc)(Jr, c, n);
        REAL d2 = i2 - (REAL)q->ICu2[r];
        lI1 += (double)(i1 * i1);
        lI2 += (double)(d2 * d2);
        if (grad)
            for (int k = 0; k < n; ++k) {
                gI1[k] += (double)(i1 * (REAL)Ir[k]);
                gI2[k] += (double)(i2 * (REAL)Jr[k]);
            }
    }
    if (grad)
        for (int k = 0; k < n; ++k) {
            REAL g = fac1 * (REAL)gR[k];
            g += fac2 * (REAL)gB[k];
            g += fac3 * (REAL)gI1[k];
            g += fac3 * (REAL)gI2[k];
            grad[k] = g;
        }
    REAL LR = (REAL)(lR / q->NR), LI1 = (REAL)(lI1 / q->NI), LI2 = (REAL)(lI2 / q->NI),
         LB = (REAL)(lB / q->NB);
    return ((LR + LI1) + LI2) + LB;           /* KG_models.py:50 */
}

/* =============================== Burgers =============================================== */
/* T = (-nu)*P_xx + P_t                      B_precomp.py:18-20                              */
static void SUF(b_build_T)(const b_problem_t *q, float mnu, REAL *T)
{
    for (int r = 0; r < q->NR; ++r)
        for (int k = 0; k < q->n; ++k) {
            size_t i = (size_t)r * q->ld + k;
            T[(size_t)r * q->n + k] = (REAL)mnu * (REAL)q->Pxx[i] + (REAL)q->Pt[i];
        }
}
/* B_models.py:28-46 and B_optimizer.py:28-55 (BC_u absent from the update, N3)              */
static REAL SUF(b_eval)(const b_problem_t *q, const REAL *T, const REAL *c, REAL *grad)
{
    const int n = q->n;
    const REAL fac1 = (REAL)(float)(2.0 / q->NR), fac2 = (REAL)(float)(2.0 / q->NB),
               fac3 = (REAL)(float)(2.0 / q->NI);
    double gR[MAXN] = {0}, gB[MAXN] = {0}, gI[MAXN] = {0};
    double lR = 0, lB = 0, lI = 0;
    for (int r = 0; r < q->NR; ++r) {
        const float *Pr = q->P + (size_t)r * q->ld;
        const float *Xr = q->Px + (size_t)r * q->ld;
        const REAL *Tr = T + (size_t)r * n;
        REAL u = SUF(dotc)(Pr, c, n);
        REAL ux = SUF(dotc)(Xr, c, n);
        REAL u_ux = u * ux;
        REAL first = SUF(dotT)(Tr, c, n) + u_ux;
        REAL d = first - (REAL)q->fhat[r];
        lR += (double)(d * d);
        if (grad)
            for (int k = 0; k < n; ++k) {
                REAL term1 = u * (REAL)Xr[k] + (REAL)Pr[k] * ux;
                REAL second = term1 + Tr[k];
                gR[k] += (double)(first * second);
            }
    }
    for (int r = 0; r < q->NB; ++r) {
        const float *Br = q->PBC + (size_t)r * q->ld;
        REAL bt = SUF(dotc)(Br, c, n);
        REAL d = bt - (REAL)q->BCu[r];
        lB += (double)(d * d);
        if (grad) for (int k = 0; k < n; ++k) gB[k] += (double)(bt * (REAL)Br[k]);
    }
    for (int r = 0; r < q->NI; ++r) {
        const float *Ir = q->PIC + (size_t)r * q->ld;
        REAL it = SUF(dotc)(Ir, c, n) - (REAL)q->ICu[r];
        lI += (double)(it * it);
        if (grad) for (int k = 0; k < n; ++k) gI[k] += (double)(it * (REAL)Ir[k]);
    }
    if (grad)
        for (int k = 0; k < n; ++k) {
            REAL g = fac1 * (REAL)gR[k];
            g += fac2 * (REAL)gB[k];
            g += fac3 * (REAL)gI[k];
            grad[k] = g;
        }
    REAL LR = (REAL)(lR / q->NR), LI = (REAL)(lI / q->NI), LB = (REAL)(lB / q->NB);
    return (LR + LI) + LB;                    /* B_models.py:42-46 */
}

/* =============================== Allen-Cahn ============================================ */
/* T = (P_t + (-lambda)*P_xx) + (-eps)*P     AC_precompute.py:20-24                          */
static void SUF(ac_build_T)(const ac_problem_t *q, float mlam, float meps, REAL *T)
{
    for (int r = 0; r < q->NR; ++r)
        for (int k = 0; k < q->n; ++k) {
            size_t i = (size_t)r * q->ld + k;
            REAL eP = (REAL)meps * (REAL)q->P[i];
            REAL lP = (REAL)mlam * (REAL)q->Pxx[i];
            T[(size_t)r * q->n + k] = ((REAL)q->Pt[i] + lP) + eP;
        }
}
/* AC_models.py:31-54 and AC_optimizer.py:36-64                                              */
static REAL SUF(ac_eval)(const ac_problem_t *q, float eps, float eps3, const REAL *T,
                         const REAL *c, REAL *grad)
{
    const int n = q->n;
    const REAL fac1 = (REAL)(float)(2.0 / q->NR), fac2 = (REAL)(float)(2.0 / q->NB),
               fac3 = (REAL)(float)(2.0 / q->NI);
    double gR[MAXN] = {0}, gB[MAXN] = {0}, gBx[MAXN] = {0}, gI[MAXN] = {0};
    double lR = 0, lB = 0, lBx = 0, lI = 0;
    for (int r = 0; r < q->NR; ++r) {
        const float *Pr = q->P + (size_t)r * q->ld;
        const REAL *Tr = T + (size_t)r * n;
        REAL u = SUF(dotc)(Pr, c, n);
        REAL first = SUF(dotT)(Tr, c, n) + (REAL)eps * (u * u * u);
        REAL d = first - (REAL)q->fhat[r];
        lR += (double)(d * d);
        if (grad) {
            REAL u2 = u * u;
            for (int k = 0; k < n; ++k) {
                REAL term1 = (REAL)eps3 * (u2 * (REAL)Pr[k]);
                REAL second = Tr[k] + term1;
                gR[k] += (double)(first * second);
            }
        }
    }
    for (int r = 0; r < q->NI; ++r) {
        const float *Ir = q->PIC + (size_t)r * q->ld;
        REAL it = SUF(dotc)(Ir, c, n) - (REAL)q->ICu[r];
        lI += (double)(it * it);
        if (grad) for (int k = 0; k < n; ++k) gI[k] += (double)(it * (REAL)Ir[k]);
    }
    for (int r = 0; r < q->NB; ++r) {
        size_t o = (size_t)r * q->ld;
        REAL b = SUF(dotc)(q->Pub + o, c, n) - SUF(dotc)(q->Plb + o, c, n);
        REAL bx = SUF(dotc)(q->Pubx + o, c, n) - SUF(dotc)(q->Plbx + o, c, n);
        lB += (double)(b * b);
        lBx += (double)(bx * bx);
        if (grad)
            for (int k = 0; k < n; ++k) {
                gB[k] += (double)(b * (REAL)q->Pdiff[o + k]);
                gBx[k] += (double)(bx * (REAL)q->Pdiffx[o + k]);
            }
    }
    if (grad)
        for (int k = 0; k < n; ++k) {
            REAL g = fac1 * (REAL)gR[k];
            g += fac3 * (REAL)gI[k];
            g += fac2 * (REAL)gB[k];
            g += fac2 * (REAL)gBx[k];
            grad[k] = g;
        }
    REAL LR = (REAL)(lR / q->NR), LI = (REAL)(lI / q->NI), LB = (REAL)(lB / q->NB),
         LBx = (REAL)(lBx / q->NB);
    return ((LR + LI) + LB) + LBx;            /* AC_models.py:49-54 */
}

/* =============================== drivers =============================================== */
/* One parameter, E epochs from c0: the loop of KG_train.py:27-37 without the early exit.
 * Writes c_E, f = loss_E, m = min_{j<E} loss_j and loss_j for j % rec_every == 0.           */
#define ORACLE_TRAJ(EVAL_CALL)                                                              \
    REAL c[MAXN], g[MAXN];                                                                  \
    for (int k = 0; k < n; ++k) c[k] = (REAL)c0[k];                                         \
    REAL mn = (REAL)INFINITY, loss = 0;                                                     \
    for (int j = 0; j <= epochs; ++j) {                                                     \
        REAL *grad = (j < epochs) ? g : NULL;                                               \
        loss = EVAL_CALL;                                                                   \
        if (trace && rec_every > 0 && j % rec_every == 0) trace[j / rec_every] = (float)loss; \
        if (j < epochs) {                                                                   \
            if (loss < mn) mn = loss;                                                       \
            for (int k = 0; k < n; ++k) c[k] = c[k] - (REAL)lrf * g[k];                     \
        }                                                                                   \
    }                                                                                       \
    for (int k = 0; k < n; ++k) c_out[k] = (float)c[k];                                     \
    *f_out = (float)loss; *m_out = (float)mn;

/* grid [Np,3] doubles (alpha,beta,gamma); c0 [Np,n]; c_out [Np,n]; f,m [Np];
 * trace [Np, epochs/rec_every+1] or NULL.                                                   */
int SUF(oracle_kg_sweep)(const kg_problem_t *q, const double *grid, int Np, const float *c0all,
                         int epochs, double lr, int rec_every, float *c_all, float *f_all,
                         float *m_all, float *trace_all)
{
    const int n = q->n;
    const float lrf = (float)lr;
    const int nrec = rec_every > 0 ? epochs / rec_every + 1 : 0;
    if (n > MAXN) return -1;
    int err = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int p = 0; p < Np; ++p) {
        REAL *T = (REAL *)malloc(sizeof(REAL) * (size_t)q->NR * n);
        if (!T) { err = -2; continue; }
        float alpha = (float)grid[3 * p], beta = (float)grid[3 * p + 1], gamma = (float)grid[3 * p + 2];
        float gamma2 = (float)(2.0 * grid[3 * p + 2]);         /* KG_precomp.py:29 */
        SUF(kg_build_T)(q, alpha, beta, T);
        const float *c0 = c0all + (size_t)p * n;
        float *c_out = c_all + (size_t)p * n, *f_out = f_all + p, *m_out = m_all + p;
        float *trace = trace_all ? trace_all + (size_t)p * nrec : NULL;
        ORACLE_TRAJ(SUF(kg_eval)(q, gamma, gamma2, T, c, grad))
        free(T);
    }
    return err;
}

int SUF(oracle_b_sweep)(const b_problem_t *q, const double *grid, int Np, const float *c0all,
                        int epochs, double lr, int rec_every, float *c_all, float *f_all,
                        float *m_all, float *trace_all)
{
    const int n = q->n;
    const float lrf = (float)lr;
    const int nrec = rec_every > 0 ? epochs / rec_every + 1 : 0;
    if (n > MAXN) return -1;
    int err = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int p = 0; p < Np; ++p) {
        REAL *T = (REAL *)malloc(sizeof(REAL) * (size_t)q->NR * n);
        if (!T) { err = -2; continue; }
        SUF(b_build_T)(q, (float)(-grid[p]), T);               /* B_precomp.py:19 */
        const float *c0 = c0all + (size_t)p * n;
        float *c_out = c_all + (size_t)p * n, *f_out = f_all + p, *m_out = m_all + p;
        float *trace = trace_all ? trace_all + (size_t)p * nrec : NULL;
        ORACLE_TRAJ(SUF(b_eval)(q, T, c, grad))
        free(T);
    }
    return err;
}

int SUF(oracle_ac_sweep)(const ac_problem_t *q, const double *grid, int Np, const float *c0all,
                         int epochs, double lr, int rec_every, float *c_all, float *f_all,
                         float *m_all, float *trace_all)
{
    const int n = q->n;
    const float lrf = (float)lr;
    const int nrec = rec_every > 0 ? epochs / rec_every + 1 : 0;
    if (n > MAXN) return -1;
    int err = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int p = 0; p < Np; ++p) {
        REAL *T = (REAL *)malloc(sizeof(REAL) * (size_t)q->NR * n);
        if (!T) { err = -2; continue; }
        double lam = grid[2 * p], eps = grid[2 * p + 1];
        SUF(ac_build_T)(q, (float)(-lam), (float)(-eps), T);
        const float *c0 = c0all + (size_t)p * n;
        float *c_out = c_all + (size_t)p * n, *f_out = f_all + p, *m_out = m_all + p;
        float *trace = trace_all ? trace_all + (size_t)p * nrec : NULL;
        ORACLE_TRAJ(SUF(ac_eval)(q, (float)eps, (float)(3.0 * eps), T, c, grad))
        free(T);
    }
    return err;
}

/* Literal sequential offline_generation (KG_train.py:12-43) WITH the early exit and the
 * running maximum -- used to prove the two-scalar scan equivalent.  Returns index or -1.    */
int SUF(oracle_kg_offline)(const kg_problem_t *q, const double *grid, int Np, const float *c0all,
                           int epochs, double lr, float *L_out, int *idx_out)
{
    const int n = q->n;
    const float lrf = (float)lr;
    REAL *T = (REAL *)malloc(sizeof(REAL) * (size_t)q->NR * n);
    if (!T) return -2;
    REAL L = 0; int idx = -1;
    for (int p = 0; p < Np; ++p) {
        float alpha = (float)grid[3 * p], beta = (float)grid[3 * p + 1], gamma = (float)grid[3 * p + 2];
        float gamma2 = (float)(2.0 * grid[3 * p + 2]);
        SUF(kg_build_T)(q, alpha, beta, T);
        REAL c[MAXN], g[MAXN];
        for (int k = 0; k < n; ++k) c[k] = (REAL)c0all[(size_t)p * n + k];
        REAL loss = SUF(kg_eval)(q, gamma, gamma2, T, c, g);
        for (int i = 1; i <= epochs; ++i) {
            if (loss < L) break;
            for (int k = 0; k < n; ++k) c[k] = c[k] - (REAL)lrf * g[k];
            loss = SUF(kg_eval)(q, gamma, gamma2, T, c, g);
        }
        if (loss > L) { L = loss; idx = p; }
    }
    free(T);
    *L_out = (float)L; *idx_out = idx;
    return 0;
}
