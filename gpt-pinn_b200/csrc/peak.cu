// peak.cu -- FFMA-only microbenchmark: the measured FP32 CUDA-core roofline denominator
// (MEASURED_PEAKS.json carries HBM and bf16 tensor peaks only; SURVEY.md 8d asks for this one).
#include <cuda_runtime.h>
#include "../../include/gptpinn_b200.h"

namespace gp {
__global__ void __launch_bounds__(256) ffma_peak_kernel(float *out, int iters, float a, float b)
{
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f,
          x6 = x0 + 6.f, x7 = x0 + 7.f, x8 = x0 + 8.f, x9 = x0 + 9.f, xa = x0 + 10.f, xb = x0 + 11.f,
          xc = x0 + 12.f, xd = x0 + 13.f, xe = x0 + 14.f, xf = x0 + 15.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
            x8 = fmaf(x8, a, b); x9 = fmaf(x9, a, b); xa = fmaf(xa, a, b); xb = fmaf(xb, a, b);
            xc = fmaf(xc, a, b); xd = fmaf(xd, a, b); xe = fmaf(xe, a, b); xf = fmaf(xf, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] =
        x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + x8 + x9 + xa + xb + xc + xd + xe + xf;
}
}  // namespace gp

// Returns measured TFLOP/s (2 flop per FMA) in *tflops_out; runs `reps` timed launches of about
// `ms_target` ms each on `stream` and keeps the best.  Synchronises.
extern "C" int gptpinn_ffma_peak(double *tflops_out, int reps, void *stream)
{
    if (!tflops_out) return GPTPINN_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, nsm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return GPTPINN_E_CUDA;
    if (cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return GPTPINN_E_CUDA;
    const int blocks = nsm * 8, threads = 256, iters = 20000;
    float *out = nullptr;
    if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(float)) != cudaSuccess) return GPTPINN_E_NOMEM;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int r = 0; r < reps + 1; ++r) {
        cudaEventRecord(e0, st);
        gp::ffma_peak_kernel<<<blocks, threads, 0, st>>>(out, iters, 0.999f, 0.001f);
        cudaEventRecord(e1, st);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return GPTPINN_E_CUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * 16 * 8 * (double)iters * blocks * threads;
        const double tf = flop / (ms * 1e-3) / 1e12;
        if (r > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    *tflops_out = best;
    return GPTPINN_OK;
}
