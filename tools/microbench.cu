// microbench.cu -- FFMA issue-rate probes for the sweep kernel's two inner patterns (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu && ./microbench
// Patterns (per thread, M parameters, K = 16 operand registers per block):
//   bwd   : g[m][k] = fma(w[m], t[k], g[m][k])                 (gradient accumulation)
//   fwd_km: u[m]    = fma(p[k], c[m][k], u[m])   k outer, m inner
//   fwd_mk: same, m outer, k inner (one dependent chain per m)
//   fwd2  : packed f32x2 over parameter pairs: (u[2j],u[2j+1]) = fma2((p[k],p[k]), (c[2j][k],c[2j+1][k]), ...)
//   bwd2  : packed: (g[2j][k],g[2j+1][k]) = fma2((w[2j],w[2j+1]), (t[k],t[k]), ...)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define K 16

__device__ __forceinline__ unsigned long long pack2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <int M>
__global__ void __launch_bounds__(256, 1) k_bwd(float *out, const float *in, int iters)
{
    float g[M][K], t[K], w[M];
    for (int k = 0; k < K; ++k) t[k] = in[k + threadIdx.x % 7];
    for (int m = 0; m < M; ++m) { w[m] = in[32 + m]; for (int k = 0; k < K; ++k) g[m][k] = 0.f; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int m = 0; m < M; ++m)
#pragma unroll
            for (int k = 0; k < K; ++k) g[m][k] = fmaf(w[m], t[k], g[m][k]);
#pragma unroll
        for (int m = 0; m < M; ++m) w[m] += 1e-9f;        // keeps the loop from being hoisted; M extra FADD
    }
    float s = 0.f;
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) s += g[m][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int M, bool KOUTER>
__global__ void __launch_bounds__(256, 1) k_fwd(float *out, const float *in, int iters)
{
    float c[M][K], p[K], u[M];
    for (int k = 0; k < K; ++k) p[k] = in[k + threadIdx.x % 7];
    for (int m = 0; m < M; ++m) { u[m] = 0.f; for (int k = 0; k < K; ++k) c[m][k] = in[64 + m * K + k]; }
    for (int it = 0; it < iters; ++it) {
        if (KOUTER) {
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int m = 0; m < M; ++m) u[m] = fmaf(p[k], c[m][k], u[m]);
        } else {
#pragma unroll
            for (int m = 0; m < M; ++m)
#pragma unroll
                for (int k = 0; k < K; ++k) u[m] = fmaf(p[k], c[m][k], u[m]);
        }
#pragma unroll
        for (int k = 0; k < K; k += 4) p[k] += 1e-9f;
    }
    float s = 0.f;
    for (int m = 0; m < M; ++m) s += u[m];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int M>      // M even
__global__ void __launch_bounds__(256, 1) k_fwd2(float *out, const float *in, int iters)
{
    unsigned long long c2[M / 2][K], p2[K], u2[M / 2];
    for (int k = 0; k < K; ++k) { float v = in[k + threadIdx.x % 7]; p2[k] = pack2(v, v); }
    for (int j = 0; j < M / 2; ++j) { u2[j] = pack2(0.f, 0.f); for (int k = 0; k < K; ++k) c2[j][k] = pack2(in[64 + j * K + k], in[65 + j * K + k]); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int j = 0; j < M / 2; ++j) u2[j] = fma2(p2[k], c2[j][k], u2[j]);
#pragma unroll
        for (int k = 0; k < K; k += 4) p2[k] += 1;
    }
    float s = 0.f;
    for (int j = 0; j < M / 2; ++j) { float a, b; unpack2(u2[j], a, b); s += a + b; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int M>
__global__ void __launch_bounds__(256, 1) k_bwd2(float *out, const float *in, int iters)
{
    unsigned long long g2[M / 2][K], t2[K], w2[M / 2];
    for (int k = 0; k < K; ++k) { float v = in[k + threadIdx.x % 7]; t2[k] = pack2(v, v); }
    for (int j = 0; j < M / 2; ++j) { w2[j] = pack2(in[32 + j], in[40 + j]); for (int k = 0; k < K; ++k) g2[j][k] = pack2(0.f, 0.f); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < M / 2; ++j)
#pragma unroll
            for (int k = 0; k < K; ++k) g2[j][k] = fma2(w2[j], t2[k], g2[j][k]);
#pragma unroll
        for (int j = 0; j < M / 2; ++j) w2[j] += 1;
    }
    float s = 0.f;
    for (int j = 0; j < M / 2; ++j) for (int k = 0; k < K; ++k) { float a, b; unpack2(g2[j][k], a, b); s += a + b; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// forward contraction with the two dot products packed in one FFMA2: (u[m], t[m]) += (P[k], T[k]) * (c[m][k], c[m][k])
template <int M>
__global__ void __launch_bounds__(256, 1) k_fwd_pt2(float *out, const float *in, int iters)
{
    unsigned long long cc[M][K], pt[K], acc[M];
    for (int k = 0; k < K; ++k) pt[k] = pack2(in[k + threadIdx.x % 7], in[k + 1 + threadIdx.x % 5]);
    for (int m = 0; m < M; ++m) { acc[m] = pack2(0.f, 0.f); for (int k = 0; k < K; ++k) { float v = in[64 + m * K + k]; cc[m][k] = pack2(v, v); } }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int m = 0; m < M; ++m) acc[m] = fma2(pt[k], cc[m][k], acc[m]);
#pragma unroll
        for (int k = 0; k < K; k += 4) pt[k] += 1;
    }
    float s = 0.f;
    for (int m = 0; m < M; ++m) { float a, b; unpack2(acc[m], a, b); s += a + b; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// whole row step of the proposed M = 4 kernel: packed forward (FFMA2) + plain gradient FFMAs reading the halves of the pairs
template <int M>
__global__ void __launch_bounds__(256, 1) k_row_pt2(float *out, const float *in, int iters)
{
    unsigned long long cc[M][K], pt[K];
    float g[M][K];
    for (int k = 0; k < K; ++k) pt[k] = pack2(in[k + threadIdx.x % 7], in[k + 1 + threadIdx.x % 5]);
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) { float v = in[64 + m * K + k]; cc[m][k] = pack2(v, v); g[m][k] = 0.f; }
    for (int it = 0; it < iters; ++it) {
        unsigned long long acc[M];
#pragma unroll
        for (int m = 0; m < M; ++m) acc[m] = pack2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int m = 0; m < M; ++m) acc[m] = fma2(pt[k], cc[m][k], acc[m]);
        float first[M], w2[M];
#pragma unroll
        for (int m = 0; m < M; ++m) { float u, t; unpack2(acc[m], u, t); first[m] = fmaf(0.3f, u * u, t); w2[m] = first[m] * (0.6f * u); }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float pk, tk;
            unpack2(pt[k], pk, tk);
#pragma unroll
            for (int m = 0; m < M; ++m) { g[m][k] = fmaf(first[m], tk, g[m][k]); g[m][k] = fmaf(w2[m], pk, g[m][k]); }
        }
#pragma unroll
        for (int k = 0; k < K; k += 4) pt[k] += 1;
    }
    float s = 0.f;
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) s += g[m][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the same row step with plain FFMAs (what the shipped kernel does), for comparison
template <int M>
__global__ void __launch_bounds__(256, 1) k_row_plain(float *out, const float *in, int iters)
{
    float c[M][K], p[K], t[K], g[M][K];
    for (int k = 0; k < K; ++k) { p[k] = in[k + threadIdx.x % 7]; t[k] = in[k + 1 + threadIdx.x % 5]; }
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) { c[m][k] = in[64 + m * K + k]; g[m][k] = 0.f; }
    for (int it = 0; it < iters; ++it) {
        float first[M], w2[M];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            float u = 0.f, tt = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) { u = fmaf(p[k], c[m][k], u); tt = fmaf(t[k], c[m][k], tt); }
            first[m] = fmaf(0.3f, u * u, tt); w2[m] = first[m] * (0.6f * u);
        }
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int m = 0; m < M; ++m) { g[m][k] = fmaf(first[m], t[k], g[m][k]); g[m][k] = fmaf(w2[m], p[k], g[m][k]); }
#pragma unroll
        for (int k = 0; k < K; k += 4) { p[k] += 1e-9f; t[k] += 1e-9f; }
    }
    float s = 0.f;
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) s += g[m][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static void run(const char *name, F launch, double fma_per_thread_iter, int iters, int threads)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(iters / 10);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        launch(iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double fmas = fma_per_thread_iter * iters * 148.0 * threads;
    const double tf = 2 * fmas / (best * 1e-3) / 1e12;
    printf("%-22s threads/SM %4d  %8.3f ms  %7.2f TFLOP/s  %6.2f FMA/clk/SM (at %.0f MHz nominal)\n", name, threads, best, tf,
           fmas / (best * 1e-3) / (clk * 1e3) / 148.0, clk / 1e3);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
}

int main()
{
    float *out, *in;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&in, 4096 * sizeof(float));
    cudaMemset(in, 0, 4096 * sizeof(float));
    const int iters = 20000;
    for (int threads : {256}) {
        run("bwd M=5", [&](int it) { k_bwd<5><<<148, threads>>>(out, in, it); }, 5 * K, iters, threads);
        run("bwd M=4", [&](int it) { k_bwd<4><<<148, threads>>>(out, in, it); }, 4 * K, iters, threads);
        run("fwd k-outer M=5", [&](int it) { k_fwd<5, true><<<148, threads>>>(out, in, it); }, 5 * K, iters, threads);
        run("fwd k-outer M=4", [&](int it) { k_fwd<4, true><<<148, threads>>>(out, in, it); }, 4 * K, iters, threads);
        run("fwd m-outer M=5", [&](int it) { k_fwd<5, false><<<148, threads>>>(out, in, it); }, 5 * K, iters, threads);
        run("fwd2 (f32x2) M=4", [&](int it) { k_fwd2<4><<<148, threads>>>(out, in, it); }, 4 * K, iters, threads);
        run("bwd2 (f32x2) M=4", [&](int it) { k_bwd2<4><<<148, threads>>>(out, in, it); }, 4 * K, iters, threads);
        run("fwd (P,T)x(c,c) f32x2 M=4", [&](int it) { k_fwd_pt2<4><<<148, threads>>>(out, in, it); }, 2 * 4 * K, iters, threads);
        run("row step f32x2 fwd M=4", [&](int it) { k_row_pt2<4><<<148, threads>>>(out, in, it); }, 4 * 4 * K, iters / 2, threads);
        run("row step plain M=4", [&](int it) { k_row_plain<4><<<148, threads>>>(out, in, it); }, 4 * 4 * K, iters / 2, threads);
        run("row step plain M=5", [&](int it) { k_row_plain<5><<<148, threads>>>(out, in, it); }, 4 * 5 * K, iters / 2, threads);
    }
    return 0;
}
