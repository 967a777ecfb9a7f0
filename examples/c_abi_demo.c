/* Minimal C caller of libgptpinn_b200.so -- no Python, no PyTorch: the C ABI of include/gptpinn_b200.h is the
 * drop-in boundary.  It runs the per-parameter body of KG_train.offline_generation (reference
 * Klein-Gordon/KG_train.py:14-37) for a small parameter grid on synthetic activation matrices and prints the
 * arg-max parameter.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/c_abi_demo.c -o c_abi_demo \
 *       -L gpt-pinn_b200/gptpinn/lib -lgptpinn_b200 -L /usr/local/cuda/lib64 -lcudart -lm \
 *       -Wl,-rpath,$PWD/gpt-pinn_b200/gptpinn/lib
 */
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <cuda_runtime.h>

#include "gptpinn_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
#define GP(x) do { int rc_ = (x); if (rc_ != GPTPINN_OK) { fprintf(stderr, "%s: %d %s\n", #x, rc_, gptpinn_last_error()); return 1; } } while (0)

static float *upload_random(size_t count, float scale, unsigned *seed)
{
    float *h = (float *)malloc(count * sizeof(float)), *d = NULL;
    for (size_t i = 0; i < count; ++i) {
        *seed = *seed * 1664525u + 1013904223u;
        h[i] = scale * ((float)(*seed >> 8) / 8388608.0f - 1.0f);
    }
    if (cudaMalloc((void **)&d, count * sizeof(float)) != cudaSuccess) { free(h); return NULL; }
    cudaMemcpy(d, h, count * sizeof(float), cudaMemcpyHostToDevice);
    free(h);
    return d;
}

int main(void)
{
    const int n = 7, NN = 15 /* row stride of the caller's [N, 15] buffers */, NR = 4096, NI = 256, NB = 512;
    const int na = 4, nb = 4, ng = 10, Np = na * nb * ng, epochs = 200;
    unsigned seed = 1234u;
    gptpinn_handle *h = NULL;
    GP(gptpinn_create(&h));

    gptpinn_kg_problem p;
    p.n = n; p.NR = NR; p.NI = NI; p.NB = NB;
    float *bufs[6];
    const int rows[6] = {NR, NR, NR, NI, NI, NB};
    for (int i = 0; i < 6; ++i) if (!(bufs[i] = upload_random((size_t)rows[i] * NN, 0.3f, &seed))) return 1;
    p.P.ptr = bufs[0];    p.P.ld = NN;            /* column-slice views out[:, 0:n] of [N, 15] buffers */
    p.Pxx.ptr = bufs[1];  p.Pxx.ld = NN;
    p.Ptt.ptr = bufs[2];  p.Ptt.ld = NN;
    p.PIC.ptr = bufs[3];  p.PIC.ld = NN;
    p.PICt.ptr = bufs[4]; p.PICt.ld = NN;
    p.PBC.ptr = bufs[5];  p.PBC.ld = NN;
    p.xcos = upload_random(NR, 0.3f, &seed);
    p.fhat = NULL;                                 /* zero targets are passed as NULL */
    p.ICu1 = upload_random(NI, 1.0f, &seed);
    p.ICu2 = NULL;
    p.BCu = upload_random(NB, 1.0f, &seed);

    double *grid = (double *)malloc((size_t)Np * 3 * sizeof(double));
    for (int g = 0, r = 0; g < ng; ++g)            /* KG_main.py:49 ordering: gamma slowest, beta fastest */
        for (int a = 0; a < na; ++a)
            for (int b = 0; b < nb; ++b, ++r) {
                grid[3 * r] = -2.0 + a / (double)(na - 1);
                grid[3 * r + 1] = b / (double)(nb - 1);
                grid[3 * r + 2] = g / (double)(ng - 1);
            }
    float c0h[15], *c0 = NULL, *f = NULL, *m = NULL, *L = NULL;
    int *idx = NULL;
    for (int k = 0; k < n; ++k) c0h[k] = 1.0f / n;
    CK(cudaMalloc((void **)&c0, n * sizeof(float)));
    CK(cudaMemcpy(c0, c0h, n * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMalloc((void **)&f, Np * sizeof(float)));
    CK(cudaMalloc((void **)&m, Np * sizeof(float)));
    CK(cudaMalloc((void **)&L, sizeof(float)));
    CK(cudaMalloc((void **)&idx, sizeof(int)));

    gptpinn_sweep_args a;
    memset(&a, 0, sizeof(a));            /* optional fields (tuning overrides, early-exit bound, fused arg-max key): off */
    a.grid_host = grid; a.Np = Np; a.c0_dev = c0; a.c0_per_param = 0; a.epochs = epochs; a.lr = 0.025; a.rec_every = 0;
    a.c_out_dev = NULL; a.f_out_dev = f; a.m_out_dev = m; a.trace_dev = NULL;
    a.cfg_M = a.cfg_SL = a.cfg_W = a.cfg_mode = a.cfg_K = 0;
    gptpinn_sweep_info info;
    GP(gptpinn_kg_sweep(h, &p, &a, &info, NULL));
    GP(gptpinn_argmax_scan(f, m, Np, L, idx, NULL));
    float Lh = 0.f;
    int ih = -1;
    CK(cudaMemcpy(&Lh, L, sizeof(float), cudaMemcpyDeviceToHost));     /* the only synchronisation */
    CK(cudaMemcpy(&ih, idx, sizeof(int), cudaMemcpyDeviceToHost));
    printf("library version %d; %d parameters x %d epochs, n = %d: M=%d SL=%d W=%d K=%d items=%d ctas=%d\n", gptpinn_version(), Np,
           epochs, n, info.M, info.SL, info.W, info.K, info.items, info.ctas);
    if (ih >= 0)
        printf("largest loss %.7f at parameter %d = (%.4f, %.4f, %.4f)\n", Lh, ih, grid[3 * ih], grid[3 * ih + 1], grid[3 * ih + 2]);
    GP(gptpinn_destroy(h));
    return 0;
}
