// Pipe-throughput microbenchmark: FFMA, packed FFMA2, MUFU.EX2 and a mixed scan-like loop.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
    float a[16], b = seed + threadIdx.x * 1e-3f;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed * i;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], b, 1.0f);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = ex2(a[i]);
        } else if (MODE == 2) {   // scan-like: 1 ex2 + 4 fma-pipe ops per state
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float e = ex2(b * a[i] * 1e-3f);
                a[i] = fmaf(e, a[i], b * 0.5f);
                b = fmaf(a[i], 1e-6f, b);
            }
        } else if (MODE == 3) {   // packed f32x2 fma
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long x, y, z;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(a[i]), "f"(a[i + 1]));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(y) : "f"(b));
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(z) : "f"(1.0f));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x) : "l"(y), "l"(z));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(x));
            }
        }
    }
    float s = b;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double ops_per_iter, int warps_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, sizeof(float) * sms * warps_per_sm * 32);
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms, warps_per_sm * 32>>>(out, 100, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<sms, warps_per_sm * 32>>>(out, iters, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double total = ops_per_iter * iters * (double)sms * warps_per_sm * 32;
    printf("%-10s warps/SM=%2d  %.3f ms  %.1f Gops/s  = %.1f ops/clk/SM at %d MHz nominal\n", name, warps_per_sm, ms,
           total / ms / 1e6, total / ms / 1e6 / sms / (clk / 1e6) * 1e0 / 1e0 * 1e0, clk / 1000);
    cudaFree(out);
}

int main() {
    for (int w : {4, 8, 16, 32}) {
        run<0>("ffma", 16, w);
        run<3>("ffma2", 16, w);
        run<1>("ex2", 16, w);
        run<2>("scanlike", 16, w);
    }
    return 0;
}
