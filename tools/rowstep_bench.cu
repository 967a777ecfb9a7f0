// rowstep_bench.cu -- register/shared-memory model of the sweep kernel's residual row loop (KG, n = 15), used to
// compare formulations of the inner loop before they go into sweep_kernel.cuh:
//   plain<M>   : what the shipped kernel does: 4n scalar FFMA per row and parameter, rows via LDS.128
//   pairk<M>   : the same arithmetic with the contraction index packed in pairs: fma.rn.f32x2 on
//                (P[2j],P[2j+1]) x (c[m][2j],c[m][2j+1]) and (g[m][2j],g[m][2j+1]) += (d,d) x (T[2j],T[2j+1])
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rowstep_bench rowstep_bench.cu && ./rowstep_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

typedef unsigned long long u64;
constexpr int RT = 64;          // rows per tile
constexpr int NQ = 4;           // column quads (16 columns: 15 neurons + forcing)

__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// tile layout as in the kernel: block = [quad][row][4 floats]; T block then P block
template <int M>
__global__ void __launch_bounds__(256, 1) k_plain(float *out, const float *in, int tiles)
{
    __shared__ __align__(16) float sm[2 * NQ * RT * 4];
    for (int i = threadIdx.x; i < 2 * NQ * RT * 4; i += blockDim.x) sm[i] = in[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, sl = lane & 7, slice = lane >> 3;
    float c[M][15], g[M][15], sp0[M], sp1[M], lacc[M];
    for (int m = 0; m < M; ++m) {
        sp0[m] = 0.1f * (sl * M + m); sp1[m] = 2.f * sp0[m]; lacc[m] = 0.f;
        for (int k = 0; k < 15; ++k) { c[m][k] = in[4096 + (sl * M + m) * 16 + k]; g[m][k] = 0.f; }
    }
    const float *Tb = sm, *Pb = sm + NQ * RT * 4;
    for (int t = 0; t < tiles; ++t) {
        int r = slice;
        float Tv[16], Pv[16];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const float4 a = *reinterpret_cast<const float4 *>(Tb + q * RT * 4 + r * 4), b = *reinterpret_cast<const float4 *>(Pb + q * RT * 4 + r * 4);
            Tv[4 * q] = a.x; Tv[4 * q + 1] = a.y; Tv[4 * q + 2] = a.z; Tv[4 * q + 3] = a.w;
            Pv[4 * q] = b.x; Pv[4 * q + 1] = b.y; Pv[4 * q + 2] = b.z; Pv[4 * q + 3] = b.w;
        }
        for (; r < RT; r += 4) {
            const int rn = min(r + 4, RT - 1);
            float first[M], w2[M], uu[M];
#pragma unroll
            for (int m = 0; m < M; ++m) { float s = 0.f;
#pragma unroll
                for (int k = 0; k < 15; ++k) s = fmaf(Pv[k], c[m][k], s);
                uu[m] = s; }
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const float u = uu[m];
                float d = fmaf(sp0[m], u * u, Tv[15]);
#pragma unroll
                for (int k = 0; k < 15; ++k) d = fmaf(Tv[k], c[m][k], d);
                lacc[m] = fmaf(d, d, lacc[m]);
                first[m] = d; w2[m] = d * (sp1[m] * u);
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
#pragma unroll
                for (int m = 0; m < M; ++m)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const int k = 4 * q + kk;
                        if (k < 15) { g[m][k] = fmaf(first[m], Tv[k], g[m][k]); g[m][k] = fmaf(w2[m], Pv[k], g[m][k]); }
                    }
                const float4 a = *reinterpret_cast<const float4 *>(Tb + q * RT * 4 + rn * 4), b = *reinterpret_cast<const float4 *>(Pb + q * RT * 4 + rn * 4);
                Tv[4 * q] = a.x; Tv[4 * q + 1] = a.y; Tv[4 * q + 2] = a.z; Tv[4 * q + 3] = a.w;
                Pv[4 * q] = b.x; Pv[4 * q + 1] = b.y; Pv[4 * q + 2] = b.z; Pv[4 * q + 3] = b.w;
            }
        }
    }
    float s = 0.f;
    for (int m = 0; m < M; ++m) { s += lacc[m]; for (int k = 0; k < 15; ++k) s += g[m][k]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// packed over contraction-index pairs.  c2[m][7] = (c[m][14], 1): the forcing column of T rides along; P's column 15 is 0.
template <int M, int ORDER>
__global__ void __launch_bounds__(256, 1) k_pairk(float *out, const float *in, int tiles)
{
    __shared__ __align__(16) float sm[2 * NQ * RT * 4];
    for (int i = threadIdx.x; i < 2 * NQ * RT * 4; i += blockDim.x) sm[i] = in[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, sl = lane & 7, slice = lane >> 3;
    u64 c2[M][8], g2[M][8];
    float sp0[M], sp1[M], lacc[M];
    for (int m = 0; m < M; ++m) {
        sp0[m] = 0.1f * (sl * M + m); sp1[m] = 2.f * sp0[m]; lacc[m] = 0.f;
        for (int j = 0; j < 8; ++j) {
            const float a = in[4096 + (sl * M + m) * 16 + 2 * j], b = (j == 7) ? 1.f : in[4096 + (sl * M + m) * 16 + 2 * j + 1];
            c2[m][j] = pack2(a, b); g2[m][j] = pack2(0.f, 0.f);
        }
    }
    const float *Tb = sm, *Pb = sm + NQ * RT * 4;
    for (int t = 0; t < tiles; ++t) {
        int r = slice;
        u64 T2[8], P2[8];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(Tb + q * RT * 4 + r * 4), b = *reinterpret_cast<const ulonglong2 *>(Pb + q * RT * 4 + r * 4);
            T2[2 * q] = a.x; T2[2 * q + 1] = a.y; P2[2 * q] = b.x; P2[2 * q + 1] = b.y;
        }
        for (; r < RT; r += 4) {
            const int rn = min(r + 4, RT - 1);
            u64 dd[M], ww[M];
            float uu[M];
            if (ORDER == 0) {       // j outer, m inner: the row pair is the reused operand
                u64 ua[M];
#pragma unroll
                for (int m = 0; m < M; ++m) ua[m] = 0ull;
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int m = 0; m < M; ++m) ua[m] = fma2(P2[j], c2[m][j], ua[m]);
#pragma unroll
                for (int m = 0; m < M; ++m) { float a, b; unpack2(ua[m], a, b); uu[m] = a + b; }
                u64 da[M];
#pragma unroll
                for (int m = 0; m < M; ++m) da[m] = pack2(sp0[m] * (uu[m] * uu[m]), 0.f);
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int m = 0; m < M; ++m) da[m] = fma2(T2[j], c2[m][j], da[m]);
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    float a, b; unpack2(da[m], a, b);
                    const float d = a + b;
                    lacc[m] = fmaf(d, d, lacc[m]);
                    const float w = d * (sp1[m] * uu[m]);
                    dd[m] = pack2(d, d); ww[m] = pack2(w, w);
                }
            } else {                // m outer, j inner
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    u64 ua = 0ull;
#pragma unroll
                    for (int j = 0; j < 8; ++j) ua = fma2(P2[j], c2[m][j], ua);
                    float a, b; unpack2(ua, a, b); uu[m] = a + b;
                }
#pragma unroll
                for (int m = 0; m < M; ++m) {
                    u64 da = pack2(sp0[m] * (uu[m] * uu[m]), 0.f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) da = fma2(T2[j], c2[m][j], da);
                    float a, b; unpack2(da, a, b);
                    const float d = a + b;
                    lacc[m] = fmaf(d, d, lacc[m]);
                    const float w = d * (sp1[m] * uu[m]);
                    dd[m] = pack2(d, d); ww[m] = pack2(w, w);
                }
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
#pragma unroll
                for (int m = 0; m < M; ++m)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int j = 2 * q + h;
                        g2[m][j] = fma2(dd[m], T2[j], g2[m][j]);
                        g2[m][j] = fma2(ww[m], P2[j], g2[m][j]);
                    }
                const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(Tb + q * RT * 4 + rn * 4), b = *reinterpret_cast<const ulonglong2 *>(Pb + q * RT * 4 + rn * 4);
                T2[2 * q] = a.x; T2[2 * q + 1] = a.y; P2[2 * q] = b.x; P2[2 * q + 1] = b.y;
            }
        }
    }
    float s = 0.f;
    for (int m = 0; m < M; ++m) {
        s += lacc[m];
        for (int j = 0; j < 8; ++j) { float a, b; unpack2(g2[m][j], a, b); s += a; if (j < 7) s += b; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// gradient-only and contraction-only packed loops (upper bounds of the two halves)
template <int M>
__global__ void __launch_bounds__(256, 1) k_pair_grad(float *out, const float *in, int iters)
{
    u64 g2[M][8], T2[8], dd[M];
    for (int j = 0; j < 8; ++j) T2[j] = pack2(in[j + threadIdx.x % 7], in[j + 1]);
    for (int m = 0; m < M; ++m) { dd[m] = pack2(in[32 + m], in[32 + m]); for (int j = 0; j < 8; ++j) g2[m][j] = 0ull; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int m = 0; m < M; ++m)
#pragma unroll
            for (int j = 0; j < 8; ++j) g2[m][j] = fma2(dd[m], T2[j], g2[m][j]);
#pragma unroll
        for (int m = 0; m < M; ++m) dd[m] += 1;
    }
    float s = 0.f;
    for (int m = 0; m < M; ++m) for (int j = 0; j < 8; ++j) { float a, b; unpack2(g2[m][j], a, b); s += a + b; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int M>
__global__ void __launch_bounds__(256, 1) k_pair_dot(float *out, const float *in, int iters)
{
    u64 c2[M][8], P2[8], ua[M];
    for (int j = 0; j < 8; ++j) P2[j] = pack2(in[j + threadIdx.x % 7], in[j + 1]);
    for (int m = 0; m < M; ++m) { ua[m] = 0ull; for (int j = 0; j < 8; ++j) c2[m][j] = pack2(in[64 + m * 16 + 2 * j], in[65 + m * 16 + 2 * j]); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int m = 0; m < M; ++m) ua[m] = fma2(P2[j], c2[m][j], ua[m]);
#pragma unroll
        for (int j = 0; j < 8; j += 4) P2[j] += 1;
    }
    float s = 0.f;
    for (int m = 0; m < M; ++m) { float a, b; unpack2(ua[m], a, b); s += a + b; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static void run(const char *name, F launch, double alg_fma_per_thread_unit, int units, float *out)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(units / 10 + 1);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        launch(units);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    float h[4];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    const double fmas = alg_fma_per_thread_unit * units * 148.0 * 256;
    printf("%-34s %8.3f ms  %7.2f alg TFLOP/s  %6.2f alg FMA/clk/SM @1965 MHz (%.3f of 128)  chk %.6e %.6e\n", name, best, 2 * fmas / (best * 1e-3) / 1e12,
           fmas / (best * 1e-3) / 1.965e9 / 148.0, fmas / (best * 1e-3) / 1.965e9 / 148.0 / 128.0, h[0], h[9]);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
}

int main()
{
    float *out, *in;
    cudaMalloc(&out, 148 * 256 * sizeof(float));
    cudaMalloc(&in, 8192 * sizeof(float));
    float *h = (float *)malloc(8192 * sizeof(float));
    srand(1);
    for (int i = 0; i < 8192; ++i) h[i] = 0.02f * ((rand() % 2001) / 1000.0f - 1.0f);
    // P block column 15 = 0 (as in the packed stream: columns past n are zero)
    for (int r = 0; r < RT; ++r) h[NQ * RT * 4 + 3 * RT * 4 + r * 4 + 3] = 0.f;
    cudaMemcpy(in, h, 8192 * sizeof(float), cudaMemcpyHostToDevice);
    const int tiles = 4000;
    const double per_tile = 16.0 * 60.0;   // 16 row steps per lane and tile, 4n = 60 algorithmic FMAs per row and parameter
    run("plain  M=5", [&](int t) { k_plain<5><<<148, 256>>>(out, in, t); }, per_tile * 5, tiles, out);
    run("plain  M=4", [&](int t) { k_plain<4><<<148, 256>>>(out, in, t); }, per_tile * 4, tiles, out);
    run("pair-k M=5 j-outer", [&](int t) { k_pairk<5, 0><<<148, 256>>>(out, in, t); }, per_tile * 5, tiles, out);
    run("pair-k M=5 m-outer", [&](int t) { k_pairk<5, 1><<<148, 256>>>(out, in, t); }, per_tile * 5, tiles, out);
    run("pair-k M=4 j-outer", [&](int t) { k_pairk<4, 0><<<148, 256>>>(out, in, t); }, per_tile * 4, tiles, out);
    run("pair-k M=4 m-outer", [&](int t) { k_pairk<4, 1><<<148, 256>>>(out, in, t); }, per_tile * 4, tiles, out);
    run("pair-k M=6 j-outer", [&](int t) { k_pairk<6, 0><<<148, 256>>>(out, in, t); }, per_tile * 6, tiles, out);
    run("pair-k M=6 m-outer", [&](int t) { k_pairk<6, 1><<<148, 256>>>(out, in, t); }, per_tile * 6, tiles, out);
    run("pair-k M=3 j-outer (384 thr)", [&](int t) { k_pairk<3, 0><<<148, 256>>>(out, in, t); }, per_tile * 3, tiles, out);
    // halves: executed FMAs (16 columns)
    run("packed gradient only M=5", [&](int t) { k_pair_grad<5><<<148, 256>>>(out, in, t); }, 5 * 16.0, 40000, out);
    run("packed contraction only M=5", [&](int t) { k_pair_dot<5><<<148, 256>>>(out, in, t); }, 5 * 16.0, 40000, out);
    run("packed gradient only M=6", [&](int t) { k_pair_grad<6><<<148, 256>>>(out, in, t); }, 6 * 16.0, 40000, out);
    run("packed contraction only M=6", [&](int t) { k_pair_dot<6><<<148, 256>>>(out, in, t); }, 6 * 16.0, 40000, out);
    return 0;
}
