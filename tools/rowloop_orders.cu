// rowloop.cu -- source-order variants of the KG residual row loop (n = 15), for the SASS register-file model and the
// B200 microbenchmark.  -DVAR=<id> -DMM=<M>.  All variants compute the same quantities (summation order may differ).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#ifndef CV
#define CV 0
#endif
#ifndef GV
#define GV 0
#endif
#ifndef PIN
#define PIN 1
#endif
#ifndef MM
#define MM 5
#endif
#ifndef NN
#define NN 15
#endif
constexpr int RT = 64, NQ = 4;
constexpr int M = MM, N = NN;

__device__ __forceinline__ float ffma(float a, float b, float c)
{
#if PIN
    float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
#else
    return fmaf(a, b, c);
#endif
}
__device__ __forceinline__ void ldq(const float *blk, int q, int r, float *v)
{
    const float4 a = *reinterpret_cast<const float4 *>(blk + q * RT * 4 + r * 4);
    v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
}


__global__ void __launch_bounds__(256, 1) k_row(float *out, const float *in, int tiles)
{
    __shared__ __align__(16) float sm[2 * NQ * RT * 4];
    for (int i = threadIdx.x; i < 2 * NQ * RT * 4; i += blockDim.x) sm[i] = in[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, sl = lane & 7, slice = lane >> 3;
    float c[M][N], g[M][N], sp0[M], sp1[M], lacc[M];
    for (int m = 0; m < M; ++m) {
        sp0[m] = 0.1f * (sl * M + m); sp1[m] = 2.f * sp0[m]; lacc[m] = 0.f;
        for (int k = 0; k < N; ++k) { c[m][k] = in[4096 + (sl * M + m) * 16 + k]; g[m][k] = 0.f; }
    }
    const float *Tb = sm, *Pb = sm + NQ * RT * 4;
    for (int t = 0; t < tiles; ++t) {
        int r = slice;
        float Tv[16], Pv[16];
#pragma unroll
        for (int q = 0; q < NQ; ++q) { ldq(Tb, q, r, Tv); ldq(Pb, q, r, Pv); }
        for (; r < RT; r += 4) {
            const int rn = min(r + 4, RT - 1);
            float first[M], w2[M], uu[M], tt[M];
#pragma unroll
            for (int m = 0; m < M; ++m) { uu[m] = 0.f; tt[m] = 0.f; }
#if CV == 0      // k outer, m inner; P then T
#pragma unroll
            for (int k = 0; k < N; ++k)
#pragma unroll
                for (int m = 0; m < M; ++m) uu[m] = ffma(Pv[k], c[m][k], uu[m]);
#pragma unroll
            for (int k = 0; k < N; ++k)
#pragma unroll
                for (int m = 0; m < M; ++m) tt[m] = ffma(Tv[k], c[m][k], tt[m]);
#elif CV == 1    // m outer, P and T chains interleaved
#pragma unroll
            for (int m = 0; m < M; ++m)
#pragma unroll
                for (int k = 0; k < N; ++k) { uu[m] = ffma(Pv[k], c[m][k], uu[m]); tt[m] = ffma(Tv[k], c[m][k], tt[m]); }
#elif CV == 2    // k outer: all m for P[k], then all m for T[k]
#pragma unroll
            for (int k = 0; k < N; ++k) {
#pragma unroll
                for (int m = 0; m < M; ++m) uu[m] = ffma(Pv[k], c[m][k], uu[m]);
#pragma unroll
                for (int m = 0; m < M; ++m) tt[m] = ffma(Tv[k], c[m][k], tt[m]);
            }
#elif CV == 3    // k outer, m inner, (P, T) pair per (k, m): c[m][k] shared by the pair
#pragma unroll
            for (int k = 0; k < N; ++k)
#pragma unroll
                for (int m = 0; m < M; ++m) { uu[m] = ffma(c[m][k], Pv[k], uu[m]); tt[m] = ffma(c[m][k], Tv[k], tt[m]); }
#elif CV == 4    // m outer, P chain then T chain per m
#pragma unroll
            for (int m = 0; m < M; ++m) {
#pragma unroll
                for (int k = 0; k < N; ++k) uu[m] = ffma(Pv[k], c[m][k], uu[m]);
#pragma unroll
                for (int k = 0; k < N; ++k) tt[m] = ffma(Tv[k], c[m][k], tt[m]);
            }
#endif
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const float u = uu[m];
                const float d = tt[m] + fmaf(sp0[m], u * u, Tv[15]);
                lacc[m] = fmaf(d, d, lacc[m]);
                first[m] = d; w2[m] = d * (sp1[m] * u);
            }
#if GV == 0      // m outer, k inner; P gradient, reload P, T gradient, reload T
#pragma unroll
            for (int m = 0; m < M; ++m)
#pragma unroll
                for (int k = 0; k < N; ++k) g[m][k] = ffma(w2[m], Pv[k], g[m][k]);
#pragma unroll
            for (int q = 0; q < NQ; ++q) ldq(Pb, q, rn, Pv);
#pragma unroll
            for (int m = 0; m < M; ++m)
#pragma unroll
                for (int k = 0; k < N; ++k) g[m][k] = ffma(first[m], Tv[k], g[m][k]);
#pragma unroll
            for (int q = 0; q < NQ; ++q) ldq(Tb, q, rn, Tv);
#elif GV == 1    // k outer, m inner, quad by quad with reloads; P then T
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) { const int k = 4 * q + kk; if (k < N) {
#pragma unroll
                    for (int m = 0; m < M; ++m) g[m][k] = ffma(w2[m], Pv[k], g[m][k]); } }
                ldq(Pb, q, rn, Pv);
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) { const int k = 4 * q + kk; if (k < N) {
#pragma unroll
                    for (int m = 0; m < M; ++m) g[m][k] = ffma(first[m], Tv[k], g[m][k]); } }
                ldq(Tb, q, rn, Tv);
            }
#elif GV == 2    // shipped: quad outer, m, k inner, T and P updates of one g back to back
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
#pragma unroll
                for (int m = 0; m < M; ++m)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) { const int k = 4 * q + kk; if (k < N) { g[m][k] = ffma(first[m], Tv[k], g[m][k]); g[m][k] = ffma(w2[m], Pv[k], g[m][k]); } }
                ldq(Tb, q, rn, Tv); ldq(Pb, q, rn, Pv);
            }
#elif GV == 3    // quad outer; per quad: m outer k inner for T, then for P
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
#pragma unroll
                for (int m = 0; m < M; ++m)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) { const int k = 4 * q + kk; if (k < N) g[m][k] = ffma(first[m], Tv[k], g[m][k]); }
                ldq(Tb, q, rn, Tv);
#pragma unroll
                for (int m = 0; m < M; ++m)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) { const int k = 4 * q + kk; if (k < N) g[m][k] = ffma(w2[m], Pv[k], g[m][k]); }
                ldq(Pb, q, rn, Pv);
            }
#elif GV == 4    // m outer: for each m the T then the P gradient over all k; reload at the end
#pragma unroll
            for (int m = 0; m < M; ++m) {
#pragma unroll
                for (int k = 0; k < N; ++k) g[m][k] = ffma(first[m], Tv[k], g[m][k]);
#pragma unroll
                for (int k = 0; k < N; ++k) g[m][k] = ffma(w2[m], Pv[k], g[m][k]);
            }
#pragma unroll
            for (int q = 0; q < NQ; ++q) { ldq(Tb, q, rn, Tv); ldq(Pb, q, rn, Pv); }
#endif
        }
    }
    float s = 0.f;
    for (int m = 0; m < M; ++m) { s += lacc[m]; for (int k = 0; k < N; ++k) s += g[m][k]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main(int argc, char **argv)
{
    float *out, *in;
    cudaMalloc(&out, 148 * 256 * sizeof(float));
    cudaMalloc(&in, 8192 * sizeof(float));
    float *h = (float *)malloc(8192 * sizeof(float));
    srand(1);
    for (int i = 0; i < 8192; ++i) h[i] = 0.02f * ((rand() % 2001) / 1000.0f - 1.0f);
    cudaMemcpy(in, h, 8192 * sizeof(float), cudaMemcpyHostToDevice);
    const int tiles = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_row<<<148, 256>>>(out, in, tiles / 10); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k_row<<<148, 256>>>(out, in, tiles); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    float hs[2]; cudaMemcpy(hs, out, sizeof(hs), cudaMemcpyDeviceToHost);
    const double fmas = 16.0 * 4 * N * M * tiles * 148.0 * 256;
    printf("CV %d GV %d PIN %d M %d N %2d: %8.3f ms  alg %.3f of 128 FMA/clk/SM  chk %.6e\n", CV, GV, PIN, M, N, best, fmas / (best * 1e-3) / 1.965e9 / 148.0 / 128.0, hs[0]);
    return 0;
}
