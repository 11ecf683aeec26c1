// Shared device helpers for the SS2D hot-path kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/medmamba_b200.h"

namespace mmb {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kMaxState = 16;   // d_state handled per row (MedMamba uses 16 everywhere, MedMamba.py:127)

// One MUFU.EX2 / MUFU.LG2 each, no denormal fix-up code around them.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Packed fp32 pairs (FFMA2 / FMUL2 on sm_100a): two FMAs per issue slot.  The scan's inner loop is
// bound by issue slots and MUFU, not by FMA-pipe lanes, so halving the instruction count of the
// state update pays directly.  Operands are passed as scalars; ptxas allocates aligned pairs.
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void mul2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "mul.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

// softplus(x) = log1p(exp(x)), identity above 20 (torch F.softplus defaults; temp.py:63-64).
// v = e^x; below 1/8 the alternating series (rel. error < 1e-7), above it lg2(1+v) whose
// absolute error (2^-22) is then small against the result.
__device__ __forceinline__ float softplus_f(float x) {
    const float v = ex2_approx(x * kLog2e);
    const float big = lg2_approx(1.0f + v) * kLn2;
    float s = fmaf(v, -0.125f, 1.0f / 7.0f);
    s = fmaf(v, s, -1.0f / 6.0f);
    s = fmaf(v, s, 0.2f);
    s = fmaf(v, s, -0.25f);
    s = fmaf(v, s, 1.0f / 3.0f);
    s = fmaf(v, s, -0.5f);
    s = fmaf(v, s, 1.0f);
    const float r = v < 0.125f ? v * s : big;
    return x > 20.0f ? x : r;
}

// Two softplus evaluations with the series and the fix-up arithmetic issued as packed pairs.
__device__ __forceinline__ void softplus2_f(float& r0, float& r1, float x0, float x1) {
    const float v0 = ex2_approx(x0 * kLog2e), v1 = ex2_approx(x1 * kLog2e);
    const float big0 = lg2_approx(1.0f + v0) * kLn2, big1 = lg2_approx(1.0f + v1) * kLn2;
    float s0, s1;
    fma2(s0, s1, v0, v1, -0.125f, -0.125f, 1.0f / 7.0f, 1.0f / 7.0f);
    fma2(s0, s1, v0, v1, s0, s1, -1.0f / 6.0f, -1.0f / 6.0f);
    fma2(s0, s1, v0, v1, s0, s1, 0.2f, 0.2f);
    fma2(s0, s1, v0, v1, s0, s1, -0.25f, -0.25f);
    fma2(s0, s1, v0, v1, s0, s1, 1.0f / 3.0f, 1.0f / 3.0f);
    fma2(s0, s1, v0, v1, s0, s1, -0.5f, -0.5f);
    fma2(s0, s1, v0, v1, s0, s1, 1.0f, 1.0f);
    mul2(s0, s1, v0, v1, s0, s1);
    const float q0 = v0 < 0.125f ? s0 : big0, q1 = v1 < 0.125f ? s1 : big1;
    r0 = x0 > 20.0f ? x0 : q0;
    r1 = x1 > 20.0f ? x1 : q1;
}

// Four softplus evaluations for one warp-uniform group of steps.  The lg2 branch of softplus_f is needed only when
// e^x >= 1/8; with MedMamba's dt range (softplus^-1 of 1e-3 .. 1e-1 plus a small projection) that is rare, so the
// warp votes once per group and skips the four MUFU.LG2 (the scan kernel is bound by the MUFU rate) when no lane
// needs them.  Bit-identical to softplus_f in both branches.
__device__ __forceinline__ void softplus4_vote(float (&r)[4], const float (&x)[4]) {
    float v[4], s[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = ex2_approx(x[i] * kLog2e);
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
        float s0, s1;
        fma2(s0, s1, v[i], v[i + 1], -0.125f, -0.125f, 1.0f / 7.0f, 1.0f / 7.0f);
        fma2(s0, s1, v[i], v[i + 1], s0, s1, -1.0f / 6.0f, -1.0f / 6.0f);
        fma2(s0, s1, v[i], v[i + 1], s0, s1, 0.2f, 0.2f);
        fma2(s0, s1, v[i], v[i + 1], s0, s1, -0.25f, -0.25f);
        fma2(s0, s1, v[i], v[i + 1], s0, s1, 1.0f / 3.0f, 1.0f / 3.0f);
        fma2(s0, s1, v[i], v[i + 1], s0, s1, -0.5f, -0.5f);
        fma2(s0, s1, v[i], v[i + 1], s0, s1, 1.0f, 1.0f);
        mul2(s[i], s[i + 1], v[i], v[i + 1], s0, s1);
    }
    const bool need = !(v[0] < 0.125f) || !(v[1] < 0.125f) || !(v[2] < 0.125f) || !(v[3] < 0.125f);   // NaN -> slow path
    if (__any_sync(0xffffffffu, need)) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float big = lg2_approx(1.0f + v[i]) * kLn2;
            const float q = v[i] < 0.125f ? s[i] : big;
            r[i] = x[i] > 20.0f ? x[i] : q;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = s[i];
    }
}

// softplus(x) and its derivative sigmoid(x) = e^x / (1 + e^x) from one e^x (three MUFU operations instead of four).
__device__ __forceinline__ void softplus_sigmoid_f(float x, float& sp, float& sg) {
    const float v = ex2_approx(x * kLog2e);
    const float big = lg2_approx(1.0f + v) * kLn2;
    float s = fmaf(v, -0.125f, 1.0f / 7.0f);
    s = fmaf(v, s, -1.0f / 6.0f);
    s = fmaf(v, s, 0.2f);
    s = fmaf(v, s, -0.25f);
    s = fmaf(v, s, 1.0f / 3.0f);
    s = fmaf(v, s, -0.5f);
    s = fmaf(v, s, 1.0f);
    const float r = v < 0.125f ? v * s : big;
    sp = x > 20.0f ? x : r;
    sg = x > 20.0f ? 1.0f : v * rcp_approx(1.0f + v);
}

// d softplus / dx = sigmoid(x) (1 above the threshold, as torch's backward does).
__device__ __forceinline__ float sigmoid_f(float x) {
    return rcp_approx(1.0f + ex2_approx(-x * kLog2e));
}
__device__ __forceinline__ float silu_f(float x) { return x * sigmoid_f(x); }
// (silu(x) = h + h tanh(h), h = x / 2, with MUFU.TANH was tried for bf16 outputs and rejected: for x << 0 the result
// is h (1 + tanh h) with 1 + tanh h ~ 1e-3 and tanh.approx's absolute error 5e-4 -- the relative error explodes.)

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

// 4 consecutive elements -> float4 (16 B for fp32, 8 B for 16-bit types); p must be aligned to that.
template <typename T> __device__ __forceinline__ float4 load4(const T* p);
template <> __device__ __forceinline__ float4 load4<float>(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}
template <> __device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
template <> __device__ __forceinline__ float4 load4<__half>(const __half* p) {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T> __device__ __forceinline__ void store4(T* p, float4 v);
template <> __device__ __forceinline__ void store4<float>(float* p, float4 v) {
    *reinterpret_cast<float4*>(p) = v;
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
}
template <> __device__ __forceinline__ void store4<__half>(__half* p, float4 v) {
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 r;
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = r;
}

template <typename T> __host__ __device__ constexpr int vec4_align() { return 4 * (int)sizeof(T); }

inline int cuda_status(cudaError_t e) { return e == cudaSuccess ? MMB_OK : MMB_ERR_CUDA_BASE - (int)e; }
inline int launch_status() { return cuda_status(cudaGetLastError()); }

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace mmb
