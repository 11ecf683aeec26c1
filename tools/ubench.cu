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
        } else if (MODE == 4) {   // packed half exps: ex2.approx.f16x2 (two results per MUFU instruction?)
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned int* w = reinterpret_cast<unsigned int*>(&a[i]);
                asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(*w));
                w = reinterpret_cast<unsigned int*>(&a[i + 1]);
                asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(*w));
            }
        } else if (MODE == 5) {   // ex2.approx.ftz.bf16x2
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                unsigned int* w = reinterpret_cast<unsigned int*>(&a[i]);
                asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(*w));
            }
        } else if (MODE == 6) {   // f32 pair -> pack f16x2 -> ex2 -> unpack to f32 (what a scan step would do)
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned int w;
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(a[i + 1]), "f"(a[i]));
                asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(w));
                asm volatile("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}"
                             : "=f"(a[i]), "=f"(a[i + 1]) : "r"(w));
                a[i] -= 1.5f; a[i + 1] -= 1.5f;
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

// Warp -> SM sub-partition mapping: MUFU-only CTAs of `tpb` threads, `cps` CTAs per SM; prints the exp rate and
// the histogram of %warpid % 4 on SM 0 (does a 3-warp CTA leave a sub-partition idle?).
__global__ void kmap(float* out, int* ids, int iters, float seed) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed * i + threadIdx.x * 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = ex2(a[i]);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if ((threadIdx.x & 31) == 0) {
        unsigned sm, wid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
        const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
        ids[2 * w] = sm; ids[2 * w + 1] = wid;
    }
}

void run_map(int tpb, int cps) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * cps, nw = grid * tpb / 32;
    float* out; cudaMalloc(&out, sizeof(float) * grid * tpb);
    int* ids; cudaMalloc(&ids, sizeof(int) * 2 * nw);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    kmap<<<grid, tpb>>>(out, ids, 100, 0.5f);
    cudaEventRecord(e0);
    kmap<<<grid, tpb>>>(out, ids, iters, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int* h = new int[2 * nw];
    cudaMemcpy(h, ids, sizeof(int) * 2 * nw, cudaMemcpyDeviceToHost);
    int hist[4] = {0, 0, 0, 0}, on0 = 0;
    for (int w = 0; w < nw; ++w) if (h[2 * w] == 0) { hist[h[2 * w + 1] & 3]++; on0++; }
    const double total = 16.0 * iters * (double)grid * tpb;
    printf("map tpb=%3d ctas/SM=%2d  %.3f ms  %.2f exp/clk/SM (1965 MHz)  SM0: %d warps, warpid%%4 histogram %d %d %d %d\n",
           tpb, cps, ms, total / ms / 1e6 / sms / 1.965e3, on0, hist[0], hist[1], hist[2], hist[3]);
    delete[] h; cudaFree(out); cudaFree(ids);
}

int main() {
    run_map(96, 5); run_map(96, 2); run_map(96, 7); run_map(128, 4); run_map(32, 15); run_map(32, 16);
    run_map(192, 2); run_map(64, 8); run_map(256, 2);
    for (int w : {4, 8, 16, 32}) {
        run<0>("ffma", 16, w);
        run<3>("ffma2", 16, w);
        run<1>("ex2", 16, w);
        run<2>("scanlike", 16, w);
        run<4>("ex2.f16x2", 32, w);
        run<5>("ex2.bf16x2", 32, w);
        run<6>("pack+ex2h2+unpack", 16, w);
    }
    return 0;
}
