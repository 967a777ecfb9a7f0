// Legacy warp-level tensor-core rate on sm_100a: mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32
// (round-2 feasibility of a 3xTF32 forward/backward for the sweep kernel).  Build: nvcc -arch=sm_100a -O3 mma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int CH>
__global__ void k_mma(float *out, int iters)
{
    float d[CH][4];
    unsigned a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
    for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + threadIdx.x * 1e-3f + i);
    for (int c = 0; c < CH; ++c) for (int i = 0; i < 4; ++i) d[c][i] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) mma_tf32(d[c], a, b);
    }
    float s = 0.f;
    for (int c = 0; c < CH; ++c) for (int i = 0; i < 4; ++i) s += d[c][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFMA stream running next to the MMAs (does the tensor pipe overlap with the FMA pipe?)
template <int CH>
__global__ void k_mix(float *out, int iters)
{
    float d[CH][4], f[16];
    unsigned a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
    for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + threadIdx.x * 1e-3f + i);
    for (int c = 0; c < CH; ++c) for (int i = 0; i < 4; ++i) d[c][i] = 0.f;
    for (int i = 0; i < 16; ++i) f[i] = threadIdx.x * 1e-4f + i;
    const float x = 1.0001f + threadIdx.x * 1e-7f, y = 0.999f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            mma_tf32(d[c], a, b);
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = fmaf(f[i], x, y);      // 16 FFMA per MMA
        }
    }
    float s = 0.f;
    for (int c = 0; c < CH; ++c) for (int i = 0; i < 4; ++i) s += d[c][i];
    for (int i = 0; i < 16; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static void run(const char *name, F launch, double mma_per_warp_iter, double ffma_per_thread_iter, int iters, int threads)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(iters / 10);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); launch(iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cycles = best * 1e-3 * clk * 1e3;
    const double mma_sm = mma_per_warp_iter * iters * (threads / 32);
    printf("%-28s warps/SM %2d  %8.3f ms  cycles/MMA/SM %6.2f  (%.0f TF32-FMA/clk/SM, %.1f TFLOP/s dense tf32)", name, threads / 32, best,
           cycles / mma_sm, mma_sm * 1024 / cycles, 2.0 * mma_sm * 1024 * 148 / (best * 1e-3) / 1e12);
    if (ffma_per_thread_iter > 0) printf("  + FFMA %.1f /clk/SM", ffma_per_thread_iter * iters * threads / cycles);
    printf("\n");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
}

int main()
{
    float *out;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    const int iters = 20000;
    for (int threads : {128, 256, 512}) {
        run("mma chains=4", [&](int it) { k_mma<4><<<148, threads>>>(out, it); }, 4, 0, iters, threads);
        run("mma chains=8", [&](int it) { k_mma<8><<<148, threads>>>(out, it); }, 8, 0, iters, threads);
        run("mma chains=8 + 16 FFMA each", [&](int it) { k_mix<8><<<148, threads>>>(out, it); }, 8, 8 * 16, iters, threads);
    }
    return 0;
}
