// selective_scan_fn forward at the interface layout (batch, dim, seqlen) -- the drop-in for
// mamba_ssm's selective_scan_cuda.fwd reached from MedMamba.py:273-279 (semantics temp.py:57-139).
//
// Mapping.  One CTA of 128 threads owns RT = 128/S channel rows of one (batch, group) and walks
// the sequence in chunks of T steps.  A row is owned by S adjacent lanes; lane q keeps the
// states n = q, q+S, q+2S, ... in registers for the whole sequence (fp32), so the recurrence
//     h_n <- exp2(delta * A_n*log2e) * h_n + (delta*u) * B_n ;   y += C_n * h_n
// costs one MUFU.EX2 and four FMA-pipe ops per (row, step, state) and nothing is re-computed.
// S is picked on the host so that the launch has enough warps to fill 148 SMs: big batches of
// wide stages run one thread per row (S = 1), the 96-channel stage-1 shape splits the 16 states
// over 4 lanes.  The partial y of the S lanes are combined with xor-shuffles.
//
// Staging.  Per chunk the CTA loads the u and delta tiles (RT x T, coalesced 128-bit loads along
// L; delta gets + bias and softplus here, once per element) and the B / C tiles (16 x T, any
// element strides: the reference passes views with stride(-1) = dt_rank + 2*d_state) into
// shared memory with a row pitch of T+4 floats, which makes the per-thread LDS.128 reads of
// four consecutive steps bank-conflict free.  y overwrites u in place and is stored back
// coalesced (with the optional silu(z) gate) while the next chunk is loaded.
#include <stdlib.h>

#include "common.cuh"

namespace mmb {

struct ScanFwdParams {
    const void* u; const void* delta; const void* Bm; const void* Cm; const void* z; void* out;
    const float* A; const float* Dv; const float* bias; float* last_state; float* chunk_state;
    int batch, dim, L, N, G, H, nchunks, softplus;
    int flags;       // async path: bit 0 = u / delta rows take 16-byte pieces, bit 1 = B / C rows do
    int64_t u_bs, u_ds, d_bs, d_ds, z_bs, z_ds, o_bs, o_ds;
    int64_t B_bs, B_gs, B_ns, B_ls, C_bs, C_gs, C_ns, C_ls;
};

template <typename T>
__device__ __forceinline__ float4 load_row4(const T* row, int t, int len, bool vec) {
    // 4 consecutive steps t..t+3 of one row, zero beyond len
    if (vec && t + 4 <= len) return load4<T>(row + t);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t + 0 < len) v.x = to_f<T>(row[t + 0]);
    if (t + 1 < len) v.y = to_f<T>(row[t + 1]);
    if (t + 2 < len) v.z = to_f<T>(row[t + 2]);
    if (t + 3 < len) v.w = to_f<T>(row[t + 3]);
    return v;
}

template <typename T>
__device__ __forceinline__ void store_row4(T* row, int t, int len, bool vec, float4 v) {
    if (vec && t + 4 <= len) { store4<T>(row + t, v); return; }
    if (t + 0 < len) row[t + 0] = from_f<T>(v.x);
    if (t + 1 < len) row[t + 1] = from_f<T>(v.y);
    if (t + 2 < len) row[t + 2] = from_f<T>(v.z);
    if (t + 3 < len) row[t + 3] = from_f<T>(v.w);
}

template <int S, int T, typename io_t, typename bc_t>
__global__ void __launch_bounds__(128) scan_fwd_kernel(const ScanFwdParams p) {
    constexpr int NT = 128;
    constexpr int RT = NT / S;           // rows per CTA
    constexpr int NS = kMaxState / S;    // states per lane
    constexpr int JB = NS < 8 ? NS : 8;  // states per register block
    constexpr int TP = T + 4;            // smem row pitch (floats)
    constexpr int T4 = T / 4;

    extern __shared__ __align__(16) float smem[];
    float* su = smem;                    // [RT][TP]  u, then y in place
    float* sd = su + RT * TP;            // [RT][TP]  softplus(delta + bias)
    float* sB = sd + RT * TP;            // [16][TP]
    float* sC = sB + kMaxState * TP;     // [16][TP]

    const int tid = threadIdx.x;
    const int r = tid / S, q = tid % S;
    const int b = blockIdx.z, g = blockIdx.y;
    const int row0 = blockIdx.x * RT;                 // first row of the tile inside the group
    const int rows_here = min(RT, p.H - row0);
    const bool valid = r < rows_here;
    const int d = g * p.H + row0 + (valid ? r : 0);   // channel row of this thread

    const io_t* ub = reinterpret_cast<const io_t*>(p.u) + (int64_t)b * p.u_bs;
    const io_t* db = reinterpret_cast<const io_t*>(p.delta) + (int64_t)b * p.d_bs;
    const io_t* zb = p.z ? reinterpret_cast<const io_t*>(p.z) + (int64_t)b * p.z_bs : nullptr;
    io_t* ob = reinterpret_cast<io_t*>(p.out) + (int64_t)b * p.o_bs;
    const bc_t* Bb = reinterpret_cast<const bc_t*>(p.Bm) + (int64_t)b * p.B_bs + (int64_t)g * p.B_gs;
    const bc_t* Cb = reinterpret_cast<const bc_t*>(p.Cm) + (int64_t)b * p.C_bs + (int64_t)g * p.C_gs;

    constexpr int VA = vec4_align<io_t>();
    const bool vec_u = ((reinterpret_cast<uintptr_t>(ub) % VA) == 0) && (p.u_ds % 4 == 0);
    const bool vec_d = ((reinterpret_cast<uintptr_t>(db) % VA) == 0) && (p.d_ds % 4 == 0);
    const bool vec_z = zb && ((reinterpret_cast<uintptr_t>(zb) % VA) == 0) && (p.z_ds % 4 == 0);
    const bool vec_o = ((reinterpret_cast<uintptr_t>(ob) % VA) == 0) && (p.o_ds % 4 == 0);

    float Ap[NS], h[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        const int n = q + S * j;
        Ap[j] = (valid && n < p.N) ? p.A[(int64_t)d * p.N + n] * kLog2e : 0.f;
        h[j] = 0.f;
    }
    const float Dd = (valid && p.Dv) ? p.Dv[d] : 0.f;

    for (int c = 0; c < p.nchunks; ++c) {
        const int t0 = c * T;
        const int len = min(T, p.L - t0);
        // ---- stage: store y of the previous chunk, then load u / delta of this one -----------
        for (int idx = tid; idx < RT * T4; idx += NT) {
            const int rr = idx / T4, tt = (idx % T4) * 4;
            float4* us = reinterpret_cast<float4*>(su + rr * TP + tt);
            float4* ds = reinterpret_cast<float4*>(sd + rr * TP + tt);
            if (rr < rows_here) {
                const int64_t dd = g * p.H + row0 + rr;
                if (c > 0) {
                    float4 y = *us;
                    const int tp = t0 - T;   // previous chunk is always full
                    if (zb) {
                        const float4 zz = load_row4<io_t>(zb + dd * p.z_ds + tp, tt, T, vec_z);
                        y.x *= silu_f(zz.x); y.y *= silu_f(zz.y); y.z *= silu_f(zz.z); y.w *= silu_f(zz.w);
                    }
                    store_row4<io_t>(ob + dd * p.o_ds + tp, tt, T, vec_o, y);
                }
                *us = load_row4<io_t>(ub + dd * p.u_ds + t0, tt, len, vec_u);
                float4 dv = load_row4<io_t>(db + dd * p.d_ds + t0, tt, len, vec_d);
                const float bs = p.bias ? p.bias[dd] : 0.f;
                dv.x += bs; dv.y += bs; dv.z += bs; dv.w += bs;
                if (p.softplus) {
                    dv.x = softplus_f(dv.x); dv.y = softplus_f(dv.y);
                    dv.z = softplus_f(dv.z); dv.w = softplus_f(dv.w);
                }
                // steps beyond the sequence: delta = 0, u = 0  =>  a = 1, b = 0, state untouched
                if (tt + 0 >= len) dv.x = 0.f;
                if (tt + 1 >= len) dv.y = 0.f;
                if (tt + 2 >= len) dv.z = 0.f;
                if (tt + 3 >= len) dv.w = 0.f;
                *ds = dv;
            } else {
                *us = make_float4(0.f, 0.f, 0.f, 0.f);
                *ds = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        // B / C tiles -> [n][TP]; iterate with the unit-stride index fastest
        if (p.B_ls == 1 || p.B_ns != 1) {
            for (int idx = tid; idx < kMaxState * T; idx += NT) {
                const int n = idx / T, tt = idx % T;
                sB[n * TP + tt] = (n < p.N && tt < len) ? to_f<bc_t>(Bb[(int64_t)n * p.B_ns + (int64_t)(t0 + tt) * p.B_ls]) : 0.f;
            }
        } else {
            for (int idx = tid; idx < kMaxState * T; idx += NT) {
                const int n = idx % kMaxState, tt = idx / kMaxState;
                sB[n * TP + tt] = (n < p.N && tt < len) ? to_f<bc_t>(Bb[(int64_t)n + (int64_t)(t0 + tt) * p.B_ls]) : 0.f;
            }
        }
        if (p.C_ls == 1 || p.C_ns != 1) {
            for (int idx = tid; idx < kMaxState * T; idx += NT) {
                const int n = idx / T, tt = idx % T;
                sC[n * TP + tt] = (n < p.N && tt < len) ? to_f<bc_t>(Cb[(int64_t)n * p.C_ns + (int64_t)(t0 + tt) * p.C_ls]) : 0.f;
            }
        } else {
            for (int idx = tid; idx < kMaxState * T; idx += NT) {
                const int n = idx % kMaxState, tt = idx / kMaxState;
                sC[n * TP + tt] = (n < p.N && tt < len) ? to_f<bc_t>(Cb[(int64_t)n + (int64_t)(t0 + tt) * p.C_ls]) : 0.f;
            }
        }
        __syncthreads();

        // ---- recurrence over the chunk, four steps per iteration --------------------------------
        const int steps = (len + 3) & ~3;
        for (int tt = 0; tt < steps; tt += 4) {
            const float4 u4 = *reinterpret_cast<const float4*>(su + r * TP + tt);
            const float4 d4 = *reinterpret_cast<const float4*>(sd + r * TP + tt);
            const float dl[4] = {d4.x, d4.y, d4.z, d4.w};
            const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
            float du[4], y[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) { du[s] = dl[s] * uu[s]; y[s] = 0.f; }
#pragma unroll
            for (int jb = 0; jb < NS; jb += JB) {
                float4 Bv[JB], Cv[JB];
#pragma unroll
                for (int j = 0; j < JB; ++j) {
                    const int n = q + S * (jb + j);
                    Bv[j] = *reinterpret_cast<const float4*>(sB + n * TP + tt);
                    Cv[j] = *reinterpret_cast<const float4*>(sC + n * TP + tt);
                }
#pragma unroll
                for (int s = 0; s < 4; ++s) {
#pragma unroll
                    for (int j = 0; j < JB; ++j) {
                        const float bb = s == 0 ? Bv[j].x : s == 1 ? Bv[j].y : s == 2 ? Bv[j].z : Bv[j].w;
                        const float cc = s == 0 ? Cv[j].x : s == 1 ? Cv[j].y : s == 2 ? Cv[j].z : Cv[j].w;
                        const float a = ex2_approx(dl[s] * Ap[jb + j]);
                        h[jb + j] = fmaf(a, h[jb + j], du[s] * bb);
                        y[s] = fmaf(h[jb + j], cc, y[s]);
                    }
                }
            }
#pragma unroll
            for (int off = S / 2; off > 0; off >>= 1) {
#pragma unroll
                for (int s = 0; s < 4; ++s) y[s] += __shfl_xor_sync(0xffffffffu, y[s], off);
            }
            if (q == 0) {
                float4 yo;
                yo.x = fmaf(Dd, uu[0], y[0]); yo.y = fmaf(Dd, uu[1], y[1]);
                yo.z = fmaf(Dd, uu[2], y[2]); yo.w = fmaf(Dd, uu[3], y[3]);
                *reinterpret_cast<float4*>(su + r * TP + tt) = yo;
            }
        }
        if (p.chunk_state && valid) {
            float* cs = p.chunk_state + (((int64_t)b * p.dim + d) * p.nchunks + c) * p.N;
#pragma unroll
            for (int j = 0; j < NS; ++j) { const int n = q + S * j; if (n < p.N) cs[n] = h[j]; }
        }
        __syncthreads();
    }
    // ---- tail: y of the last chunk -----------------------------------------------------------
    {
        const int t0 = (p.nchunks - 1) * T;
        const int len = p.L - t0;
        for (int idx = tid; idx < RT * T4; idx += NT) {
            const int rr = idx / T4, tt = (idx % T4) * 4;
            if (rr < rows_here && tt < len) {
                const int64_t dd = g * p.H + row0 + rr;
                float4 y = *reinterpret_cast<const float4*>(su + rr * TP + tt);
                if (zb) {
                    const float4 zz = load_row4<io_t>(zb + dd * p.z_ds + t0, tt, len, vec_z);
                    y.x *= silu_f(zz.x); y.y *= silu_f(zz.y); y.z *= silu_f(zz.z); y.w *= silu_f(zz.w);
                }
                store_row4<io_t>(ob + dd * p.o_ds + t0, tt, len, vec_o, y);
            }
        }
    }
    if (p.last_state && valid) {
        float* ls = p.last_state + ((int64_t)b * p.dim + d) * p.N;
#pragma unroll
        for (int j = 0; j < NS; ++j) { const int n = q + S * j; if (n < p.N) ls[n] = h[j]; }
    }
}

// ------------------------------------------------------------------------------------------------
// fp32, L-contiguous fast path: same ownership and recurrence, but the tiles of chunk c+1 travel global -> shared
// memory with cp.async (16-byte pieces, zero-filled past the sequence end) while chunk c is computed, into a
// second set of buffers.  The synchronous kernel above spends its time in the staging phase (ncu, batch 64 stage 1:
// long-scoreboard stalls 3.2 per issued instruction, 30 % of the stall samples on the STS that parks a loaded
// value, XU 44 %); here the only exposed global latency is the first chunk's.
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(src_bytes) : "memory");
}
// 4-byte pieces for rows that are only element-aligned (L = 49) and for B / C views with an arbitrary step stride
// (the reference's call site passes stride(-1) = dt_rank + 2 * d_state, MedMamba.py:261, 267-268)
__device__ __forceinline__ void cp_async4(void* dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// four consecutive elements of a shared-memory tile -> float4
template <typename T> __device__ __forceinline__ float4 load4_smem(const T* p);
template <> __device__ __forceinline__ float4 load4_smem<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 load4_smem<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
template <> __device__ __forceinline__ float4 load4_smem<__half>(const __half* p) {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

// The recurrence over one staged chunk (fp32 tiles with a row pitch of T + 4 floats), four steps per iteration; y
// overwrites u in place.  Shared by the fp32 and the 16-bit cp.async kernels.
template <int S, int T>
__device__ __forceinline__ void scan_chunk_steps(float* __restrict__ su, const float* __restrict__ sd, const float* __restrict__ sB,
                                                 const float* __restrict__ sC, const int r, const int q, const int steps,
                                                 const float (&Ap)[kMaxState / S], float (&h)[kMaxState / S], const float Dd) {
    constexpr int NS = kMaxState / S, JB = NS < 8 ? NS : 8, TP = T + 4;
    for (int tt = 0; tt < steps; tt += 4) {
        const float4 u4 = *reinterpret_cast<const float4*>(su + r * TP + tt);
        const float4 d4 = *reinterpret_cast<const float4*>(sd + r * TP + tt);
        const float dl[4] = {d4.x, d4.y, d4.z, d4.w};
        const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
        float du[4], y[4];
#pragma unroll
        for (int s = 0; s < 4; ++s) { du[s] = dl[s] * uu[s]; y[s] = 0.f; }
        if constexpr (NS >= 2) {
            // two states per packed FMUL2 / FFMA2 (one issue slot for two fp32 operations), two y accumulators
            float y2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int jb = 0; jb < NS; jb += JB) {
                float4 Bv[JB], Cv[JB];
#pragma unroll
                for (int j = 0; j < JB; ++j) {
                    const int n = q + S * (jb + j);
                    Bv[j] = *reinterpret_cast<const float4*>(sB + n * TP + tt);
                    Cv[j] = *reinterpret_cast<const float4*>(sC + n * TP + tt);
                }
#pragma unroll
                for (int s = 0; s < 4; ++s) {
#pragma unroll
                    for (int j = 0; j < JB; j += 2) {
                        const float b0 = s == 0 ? Bv[j].x : s == 1 ? Bv[j].y : s == 2 ? Bv[j].z : Bv[j].w;
                        const float b1 = s == 0 ? Bv[j + 1].x : s == 1 ? Bv[j + 1].y : s == 2 ? Bv[j + 1].z : Bv[j + 1].w;
                        const float c0 = s == 0 ? Cv[j].x : s == 1 ? Cv[j].y : s == 2 ? Cv[j].z : Cv[j].w;
                        const float c1 = s == 0 ? Cv[j + 1].x : s == 1 ? Cv[j + 1].y : s == 2 ? Cv[j + 1].z : Cv[j + 1].w;
                        float x0, x1, w0, w1;
                        mul2(x0, x1, dl[s], dl[s], Ap[jb + j], Ap[jb + j + 1]);
                        mul2(w0, w1, du[s], du[s], b0, b1);
                        const float a0 = ex2_approx(x0), a1 = ex2_approx(x1);
                        fma2(h[jb + j], h[jb + j + 1], a0, a1, h[jb + j], h[jb + j + 1], w0, w1);
                        fma2(y[s], y2[s], h[jb + j], h[jb + j + 1], c0, c1, y[s], y2[s]);
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < 4; ++s) y[s] += y2[s];
        } else {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const float bb = sB[q * TP + tt + s], cc = sC[q * TP + tt + s];
                const float a = ex2_approx(dl[s] * Ap[0]);
                h[0] = fmaf(a, h[0], du[s] * bb);
                y[s] = fmaf(h[0], cc, y[s]);
            }
        }
#pragma unroll
        for (int off = S / 2; off > 0; off >>= 1) {
#pragma unroll
            for (int s = 0; s < 4; ++s) y[s] += __shfl_xor_sync(0xffffffffu, y[s], off);
        }
        if (q == 0) {
            float4 yo;
            yo.x = fmaf(Dd, uu[0], y[0]); yo.y = fmaf(Dd, uu[1], y[1]);
            yo.z = fmaf(Dd, uu[2], y[2]); yo.w = fmaf(Dd, uu[3], y[3]);
            *reinterpret_cast<float4*>(su + r * TP + tt) = yo;
        }
    }
}

template <int S, int T>
__global__ void __launch_bounds__(128) scan_fwd_async_kernel(const ScanFwdParams p) {
    constexpr int NT = 128, RT = NT / S, NS = kMaxState / S, JB = NS < 8 ? NS : 8, TP = T + 4, T4 = T / 4;
    constexpr int BUF = (2 * RT + 2 * kMaxState) * TP;      // floats per buffer set
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int r = tid / S, q = tid % S;
    const int b = blockIdx.z, g = blockIdx.y;
    const int row0 = blockIdx.x * RT;
    const int rows_here = min(RT, p.H - row0);
    const bool valid = r < rows_here;
    const int d = g * p.H + row0 + (valid ? r : 0);
    const float* ub = reinterpret_cast<const float*>(p.u) + (int64_t)b * p.u_bs;
    const float* db = reinterpret_cast<const float*>(p.delta) + (int64_t)b * p.d_bs;
    const float* zb = p.z ? reinterpret_cast<const float*>(p.z) + (int64_t)b * p.z_bs : nullptr;
    float* ob = reinterpret_cast<float*>(p.out) + (int64_t)b * p.o_bs;
    const float* Bb = reinterpret_cast<const float*>(p.Bm) + (int64_t)b * p.B_bs + (int64_t)g * p.B_gs;
    const float* Cb = reinterpret_cast<const float*>(p.Cm) + (int64_t)b * p.C_bs + (int64_t)g * p.C_gs;
    const bool vec_z = zb && ((reinterpret_cast<uintptr_t>(zb) % 16) == 0) && (p.z_ds % 4 == 0);
    const bool vec_o = ((reinterpret_cast<uintptr_t>(ob) % 16) == 0) && (p.o_ds % 4 == 0);

    float Ap[NS], h[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        const int n = q + S * j;
        Ap[j] = (valid && n < p.N) ? p.A[(int64_t)d * p.N + n] * kLog2e : 0.f;
        h[j] = 0.f;
    }
    const float Dd = (valid && p.Dv) ? p.Dv[d] : 0.f;

    // 16-byte pieces when every row starts 16-byte aligned and runs along L; 4-byte pieces otherwise
    const bool vec_ud = (p.flags & 1) != 0, vec_bc = (p.flags & 2) != 0;
    auto prefetch = [&](int c) {
        float* base = smem + (c & 1) * BUF;
        float* su = base, *sd = su + RT * TP, *sB = sd + RT * TP, *sC = sB + kMaxState * TP;
        const int t0 = c * T, len = min(T, p.L - t0);
        if (vec_ud) {
            for (int idx = tid; idx < RT * T4; idx += NT) {
                const int rr = idx / T4, tt = (idx % T4) * 4;
                const bool rok = rr < rows_here;
                const int64_t dd = g * p.H + row0 + (rok ? rr : 0);
                const int nb = rok ? max(0, min(4, len - tt)) * 4 : 0;
                const int ts = nb ? t0 + tt : 0;                       // keep the (unused) source address inside the row
                cp_async16(su + rr * TP + tt, ub + dd * p.u_ds + ts, nb);
                cp_async16(sd + rr * TP + tt, db + dd * p.d_ds + ts, nb);
            }
        } else {
            for (int idx = tid; idx < RT * T; idx += NT) {
                const int rr = idx / T, tt = idx % T;
                const bool ok = rr < rows_here && tt < len;
                const int64_t dd = g * p.H + row0 + (ok ? rr : 0);
                const int ts = ok ? t0 + tt : 0;
                cp_async4(su + rr * TP + tt, ub + dd * p.u_ds + ts, ok ? 4 : 0);
                cp_async4(sd + rr * TP + tt, db + dd * p.d_ds + ts, ok ? 4 : 0);
            }
        }
        if (vec_bc) {
            for (int idx = tid; idx < kMaxState * T4; idx += NT) {
                const int n = idx / T4, tt = (idx % T4) * 4;
                const int nb = n < p.N ? max(0, min(4, len - tt)) * 4 : 0;
                const int ts = nb ? t0 + tt : 0;
                const int nn = n < p.N ? n : 0;
                cp_async16(sB + n * TP + tt, Bb + (int64_t)nn * p.B_ns + ts, nb);
                cp_async16(sC + n * TP + tt, Cb + (int64_t)nn * p.C_ns + ts, nb);
            }
        } else {
            // any element strides; the unit-stride index runs fastest over the threads (n for the call-site views)
            const bool n_fast = p.B_ns == 1 && p.B_ls != 1;
            for (int idx = tid; idx < kMaxState * T; idx += NT) {
                const int n = n_fast ? idx % kMaxState : idx / T, tt = n_fast ? idx / kMaxState : idx % T;
                const bool ok = n < p.N && tt < len;
                const int64_t off_b = ok ? (int64_t)n * p.B_ns + (int64_t)(t0 + tt) * p.B_ls : 0;
                const int64_t off_c = ok ? (int64_t)n * p.C_ns + (int64_t)(t0 + tt) * p.C_ls : 0;
                cp_async4(sB + n * TP + tt, Bb + off_b, ok ? 4 : 0);
                cp_async4(sC + n * TP + tt, Cb + off_c, ok ? 4 : 0);
            }
        }
        cp_async_commit();
    };

    prefetch(0);
    for (int c = 0; c < p.nchunks; ++c) {
        float* base = smem + (c & 1) * BUF;
        float* su = base, *sd = su + RT * TP, *sB = sd + RT * TP, *sC = sB + kMaxState * TP;
        const int t0 = c * T, len = min(T, p.L - t0);
        if (c + 1 < p.nchunks) { prefetch(c + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        // delta tile in place: + bias, softplus, 0 beyond the sequence (a = 1, b = 0: state untouched)
        for (int idx = tid; idx < RT * T4; idx += NT) {
            const int rr = idx / T4, tt = (idx % T4) * 4;
            if (rr < rows_here) {
                float4* ds = reinterpret_cast<float4*>(sd + rr * TP + tt);
                float4 dv = *ds;
                const float bs = p.bias ? p.bias[g * p.H + row0 + rr] : 0.f;
                dv.x += bs; dv.y += bs; dv.z += bs; dv.w += bs;
                if (p.softplus) { dv.x = softplus_f(dv.x); dv.y = softplus_f(dv.y); dv.z = softplus_f(dv.z); dv.w = softplus_f(dv.w); }
                if (tt + 0 >= len) dv.x = 0.f;
                if (tt + 1 >= len) dv.y = 0.f;
                if (tt + 2 >= len) dv.z = 0.f;
                if (tt + 3 >= len) dv.w = 0.f;
                *ds = dv;
            }
        }
        __syncthreads();
        scan_chunk_steps<S, T>(su, sd, sB, sC, r, q, (len + 3) & ~3, Ap, h, Dd);
        if (p.chunk_state && valid) {      // training: the state after every chunk (what mmb_scan_bwd recomputes from)
            float* cs = p.chunk_state + (((int64_t)b * p.dim + d) * p.nchunks + c) * p.N;
#pragma unroll
            for (int j = 0; j < NS; ++j) { const int n = q + S * j; if (n < p.N) cs[n] = h[j]; }
        }
        __syncthreads();
        // y of this chunk: shared -> global (coalesced), optional silu(z) gate
        for (int idx = tid; idx < RT * T4; idx += NT) {
            const int rr = idx / T4, tt = (idx % T4) * 4;
            if (rr < rows_here && tt < len) {
                const int64_t dd = g * p.H + row0 + rr;
                float4 y = *reinterpret_cast<const float4*>(su + rr * TP + tt);
                if (zb) {
                    const float4 zz = load_row4<float>(zb + dd * p.z_ds + t0, tt, len, vec_z);
                    y.x *= silu_f(zz.x); y.y *= silu_f(zz.y); y.z *= silu_f(zz.z); y.w *= silu_f(zz.w);
                }
                store_row4<float>(ob + dd * p.o_ds + t0, tt, len, vec_o, y);
            }
        }
        __syncthreads();      // the buffer is refilled by the prefetch of chunk c + 2 at the top of the next iteration
    }
    if (p.last_state && valid) {
        float* ls = p.last_state + ((int64_t)b * p.dim + d) * p.N;
#pragma unroll
        for (int j = 0; j < NS; ++j) { const int n = q + S * j; if (n < p.N) ls[n] = h[j]; }
    }
}

// ------------------------------------------------------------------------------------------------
// 16-bit I/O (bf16 / fp16 u, delta, z, out; B / C in the same type or fp32) on the cp.async path: the raw tiles of chunk
// c + 1 travel into a second staging set while chunk c is computed; a conversion pass (where the fp32 kernel runs its
// softplus pass) expands them into the fp32 tiles the recurrence reads.  Pieces of 16 bytes (8 elements) where rows are
// 16-byte aligned (L % 8 == 0), of 4 bytes (2 elements) for even row strides.
template <int S, int T, typename io_t, typename bc_t>
__global__ void __launch_bounds__(128) scan_fwd_async16_kernel(const ScanFwdParams p) {
    constexpr int NT = 128, RT = NT / S, NS = kMaxState / S, TP = T + 4, T4 = T / 4;
    constexpr int F32 = (2 * RT + 2 * kMaxState) * TP;                  // floats: su, sd, sB, sC
    constexpr int RAW_IO = 2 * RT * T;                                   // io_t elements per staging set: u, delta
    constexpr int RAW_BC = 2 * kMaxState * T;                            // bc_t elements per staging set: B, C
    constexpr int RAW_BYTES = RAW_IO * (int)sizeof(io_t) + RAW_BC * (int)sizeof(bc_t);
    constexpr int EB = 16 / (int)sizeof(bc_t);                           // B / C elements per 16-byte piece
    extern __shared__ __align__(16) float smem[];
    float* su = smem, *sd = su + RT * TP, *sB = sd + RT * TP, *sC = sB + kMaxState * TP;
    uint8_t* raw0 = reinterpret_cast<uint8_t*>(smem + F32);
    const int tid = threadIdx.x;
    const int r = tid / S, q = tid % S;
    const int b = blockIdx.z, g = blockIdx.y;
    const int row0 = blockIdx.x * RT;
    const int rows_here = min(RT, p.H - row0);
    const bool valid = r < rows_here;
    const int d = g * p.H + row0 + (valid ? r : 0);
    const io_t* ub = reinterpret_cast<const io_t*>(p.u) + (int64_t)b * p.u_bs;
    const io_t* db = reinterpret_cast<const io_t*>(p.delta) + (int64_t)b * p.d_bs;
    const io_t* zb = p.z ? reinterpret_cast<const io_t*>(p.z) + (int64_t)b * p.z_bs : nullptr;
    io_t* ob = reinterpret_cast<io_t*>(p.out) + (int64_t)b * p.o_bs;
    const bc_t* Bb = reinterpret_cast<const bc_t*>(p.Bm) + (int64_t)b * p.B_bs + (int64_t)g * p.B_gs;
    const bc_t* Cb = reinterpret_cast<const bc_t*>(p.Cm) + (int64_t)b * p.C_bs + (int64_t)g * p.C_gs;
    constexpr int VA = vec4_align<io_t>();
    const bool vec_z = zb && ((reinterpret_cast<uintptr_t>(zb) % VA) == 0) && (p.z_ds % 4 == 0);
    const bool vec_o = ((reinterpret_cast<uintptr_t>(ob) % VA) == 0) && (p.o_ds % 4 == 0);

    float Ap[NS], h[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        const int n = q + S * j;
        Ap[j] = (valid && n < p.N) ? p.A[(int64_t)d * p.N + n] * kLog2e : 0.f;
        h[j] = 0.f;
    }
    const float Dd = (valid && p.Dv) ? p.Dv[d] : 0.f;
    const bool vec_ud = (p.flags & 1) != 0, vec_bc = (p.flags & 2) != 0;

    auto prefetch = [&](int c) {
        io_t* ru = reinterpret_cast<io_t*>(raw0 + (c & 1) * RAW_BYTES);
        io_t* rd = ru + RT * T;
        bc_t* rB = reinterpret_cast<bc_t*>(rd + RT * T);
        bc_t* rC = rB + kMaxState * T;
        const int t0 = c * T, len = min(T, p.L - t0);
        if (vec_ud) {
            for (int idx = tid; idx < RT * (T / 8); idx += NT) {
                const int rr = idx / (T / 8), tt = (idx % (T / 8)) * 8;
                const bool rok = rr < rows_here;
                const int64_t dd = g * p.H + row0 + (rok ? rr : 0);
                const int nb = rok ? max(0, min(8, len - tt)) * 2 : 0;
                const int ts = nb ? t0 + tt : 0;
                cp_async16(ru + rr * T + tt, ub + dd * p.u_ds + ts, nb);
                cp_async16(rd + rr * T + tt, db + dd * p.d_ds + ts, nb);
            }
        } else {
            for (int idx = tid; idx < RT * (T / 2); idx += NT) {
                const int rr = idx / (T / 2), tt = (idx % (T / 2)) * 2;
                const bool rok = rr < rows_here;
                const int64_t dd = g * p.H + row0 + (rok ? rr : 0);
                const int nb = rok ? max(0, min(2, len - tt)) * 2 : 0;
                const int ts = nb ? t0 + tt : 0;
                cp_async4(ru + rr * T + tt, ub + dd * p.u_ds + ts, nb);
                cp_async4(rd + rr * T + tt, db + dd * p.d_ds + ts, nb);
            }
        }
        if (vec_bc) {
            for (int idx = tid; idx < kMaxState * (T / EB); idx += NT) {
                const int n = idx / (T / EB), tt = (idx % (T / EB)) * EB;
                const int nb = n < p.N ? max(0, min(EB, len - tt)) * (int)sizeof(bc_t) : 0;
                const int ts = nb ? t0 + tt : 0;
                const int nn = n < p.N ? n : 0;
                cp_async16(rB + n * T + tt, Bb + (int64_t)nn * p.B_ns + ts, nb);
                cp_async16(rC + n * T + tt, Cb + (int64_t)nn * p.C_ns + ts, nb);
            }
        } else {
            // 4-byte pieces: one fp32 element with any strides, or two 16-bit elements along L
            constexpr int E4 = 4 / (int)sizeof(bc_t);
            const bool n_fast = E4 == 1 && p.B_ns == 1 && p.B_ls != 1;
            for (int idx = tid; idx < kMaxState * (T / E4); idx += NT) {
                const int n = n_fast ? idx % kMaxState : idx / (T / E4);
                const int tt = (n_fast ? idx / kMaxState : idx % (T / E4)) * E4;
                const bool ok = n < p.N && tt < len;
                const int nb = ok ? max(0, min(E4, len - tt)) * (int)sizeof(bc_t) : 0;
                const int64_t off_b = ok ? (int64_t)n * p.B_ns + (int64_t)(t0 + tt) * p.B_ls : 0;
                const int64_t off_c = ok ? (int64_t)n * p.C_ns + (int64_t)(t0 + tt) * p.C_ls : 0;
                cp_async4(rB + n * T + tt, Bb + off_b, nb);
                cp_async4(rC + n * T + tt, Cb + off_c, nb);
            }
        }
        cp_async_commit();
    };

    prefetch(0);
    for (int c = 0; c < p.nchunks; ++c) {
        const io_t* ru = reinterpret_cast<const io_t*>(raw0 + (c & 1) * RAW_BYTES);
        const io_t* rd = ru + RT * T;
        const bc_t* rB = reinterpret_cast<const bc_t*>(rd + RT * T);
        const bc_t* rC = rB + kMaxState * T;
        const int t0 = c * T, len = min(T, p.L - t0);
        if (c + 1 < p.nchunks) { prefetch(c + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        // expand the raw tiles: u as is; delta + bias, softplus, 0 beyond the sequence (a = 1, b = 0: state untouched)
        for (int idx = tid; idx < RT * T4; idx += NT) {
            const int rr = idx / T4, tt = (idx % T4) * 4;
            float4 uv = make_float4(0.f, 0.f, 0.f, 0.f), dv = uv;
            if (rr < rows_here) {
                uv = load4_smem<io_t>(ru + rr * T + tt);
                dv = load4_smem<io_t>(rd + rr * T + tt);
                const float bs = p.bias ? p.bias[g * p.H + row0 + rr] : 0.f;
                dv.x += bs; dv.y += bs; dv.z += bs; dv.w += bs;
                if (p.softplus) { dv.x = softplus_f(dv.x); dv.y = softplus_f(dv.y); dv.z = softplus_f(dv.z); dv.w = softplus_f(dv.w); }
                if (tt + 0 >= len) dv.x = 0.f;
                if (tt + 1 >= len) dv.y = 0.f;
                if (tt + 2 >= len) dv.z = 0.f;
                if (tt + 3 >= len) dv.w = 0.f;
            }
            *reinterpret_cast<float4*>(su + rr * TP + tt) = uv;
            *reinterpret_cast<float4*>(sd + rr * TP + tt) = dv;
        }
        for (int idx = tid; idx < kMaxState * T4; idx += NT) {
            const int n = idx / T4, tt = (idx % T4) * 4;
            *reinterpret_cast<float4*>(sB + n * TP + tt) = load4_smem<bc_t>(rB + n * T + tt);
            *reinterpret_cast<float4*>(sC + n * TP + tt) = load4_smem<bc_t>(rC + n * T + tt);
        }
        __syncthreads();
        scan_chunk_steps<S, T>(su, sd, sB, sC, r, q, (len + 3) & ~3, Ap, h, Dd);
        if (p.chunk_state && valid) {
            float* cs = p.chunk_state + (((int64_t)b * p.dim + d) * p.nchunks + c) * p.N;
#pragma unroll
            for (int j = 0; j < NS; ++j) { const int n = q + S * j; if (n < p.N) cs[n] = h[j]; }
        }
        __syncthreads();
        for (int idx = tid; idx < RT * T4; idx += NT) {
            const int rr = idx / T4, tt = (idx % T4) * 4;
            if (rr < rows_here && tt < len) {
                const int64_t dd = g * p.H + row0 + rr;
                float4 y = *reinterpret_cast<const float4*>(su + rr * TP + tt);
                if (zb) {
                    const float4 zz = load_row4<io_t>(zb + dd * p.z_ds + t0, tt, len, vec_z);
                    y.x *= silu_f(zz.x); y.y *= silu_f(zz.y); y.z *= silu_f(zz.z); y.w *= silu_f(zz.w);
                }
                store_row4<io_t>(ob + dd * p.o_ds + t0, tt, len, vec_o, y);
            }
        }
        __syncthreads();      // su / sd are rewritten by the next chunk's expansion; its raw set by the prefetch of c + 2
    }
    if (p.last_state && valid) {
        float* ls = p.last_state + ((int64_t)b * p.dim + d) * p.N;
#pragma unroll
        for (int j = 0; j < NS; ++j) { const int n = q + S * j; if (n < p.N) ls[n] = h[j]; }
    }
}

template <int S, int T, typename io_t, typename bc_t>
static int launch_scan_fwd_async16(const ScanFwdParams& p, cudaStream_t stream) {
    constexpr size_t smem = sizeof(float) * (size_t)(2 * (128 / S) + 2 * kMaxState) * (T + 4) +
                            2 * ((size_t)2 * (128 / S) * T * sizeof(io_t) + (size_t)2 * kMaxState * T * sizeof(bc_t));
    auto kern = scan_fwd_async16_kernel<S, T, io_t, bc_t>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e);
    }
    const int RT = 128 / S;
    dim3 grid((p.H + RT - 1) / RT, p.G, p.batch);
    kern<<<grid, 128, smem, stream>>>(p);
    return launch_status();
}

template <typename io_t, typename bc_t>
static int dispatch_async16(const ScanFwdParams& p, int S, cudaStream_t st) {
    switch (S) {
        case 0: return launch_scan_fwd_async16<4, 16, io_t, bc_t>(p, st);      // checkpoint geometry
        case 1: return launch_scan_fwd_async16<1, 16, io_t, bc_t>(p, st);
        case 2: return launch_scan_fwd_async16<2, 32, io_t, bc_t>(p, st);
        case 4: return launch_scan_fwd_async16<4, 32, io_t, bc_t>(p, st);
        case 8: return launch_scan_fwd_async16<8, 32, io_t, bc_t>(p, st);
        default: return launch_scan_fwd_async16<16, 32, io_t, bc_t>(p, st);
    }
}

// 16-bit u / delta rows with an even stride (4-byte pieces) or 16-byte aligned (8-element pieces); B / C fp32 in any
// layout the fp32 path takes, or 16-bit and contiguous along L with the same alignment rules.
static bool async16_path_ok(ScanFwdParams& p, int io_dtype, int bc_dtype) {
    if ((io_dtype != MMB_BF16 && io_dtype != MMB_F16) || (bc_dtype != MMB_F32 && bc_dtype != io_dtype)) return false;
    if (getenv("MMB_SCAN_SYNC")) return false;
    const auto al = [](const void* q, int a) { return reinterpret_cast<uintptr_t>(q) % a == 0; };
    if (!al(p.u, 4) || !al(p.delta, 4) || p.u_ds % 2 || p.u_bs % 2 || p.d_ds % 2 || p.d_bs % 2) return false;
    p.flags = 0;
    if (al(p.u, 16) && al(p.delta, 16) && p.u_ds % 8 == 0 && p.u_bs % 8 == 0 && p.d_ds % 8 == 0 && p.d_bs % 8 == 0) p.flags |= 1;
    if (bc_dtype == MMB_F32) {
        if (!al(p.Bm, 4) || !al(p.Cm, 4)) return false;
        if (al(p.Bm, 16) && al(p.Cm, 16) && p.B_ls == 1 && p.C_ls == 1 && p.B_ns % 4 == 0 && p.C_ns % 4 == 0 && p.B_bs % 4 == 0 &&
            p.C_bs % 4 == 0 && p.B_gs % 4 == 0 && p.C_gs % 4 == 0) p.flags |= 2;
    } else {
        if (p.B_ls != 1 || p.C_ls != 1 || !al(p.Bm, 4) || !al(p.Cm, 4)) return false;
        if (p.B_ns % 2 || p.C_ns % 2 || p.B_bs % 2 || p.C_bs % 2 || p.B_gs % 2 || p.C_gs % 2) return false;
        if (al(p.Bm, 16) && al(p.Cm, 16) && p.B_ns % 8 == 0 && p.C_ns % 8 == 0 && p.B_bs % 8 == 0 && p.C_bs % 8 == 0 &&
            p.B_gs % 8 == 0 && p.C_gs % 8 == 0) p.flags |= 2;
    }
    return true;
}

template <int S, int T>
static int launch_scan_fwd_async(const ScanFwdParams& p, cudaStream_t stream) {
    constexpr size_t smem = 2 * sizeof(float) * (size_t)(2 * (128 / S) + 2 * kMaxState) * (T + 4);
    auto kern = scan_fwd_async_kernel<S, T>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e);
    }
    const int RT = 128 / S;
    dim3 grid((p.H + RT - 1) / RT, p.G, p.batch);
    kern<<<grid, 128, smem, stream>>>(p);
    return launch_status();
}

// The cp.async path takes any fp32 layout the operator accepts (unit stride along L for u / delta, any strides for
// B / C): 16-byte pieces where rows are 16-byte aligned and L-contiguous, 4-byte pieces otherwise.
static bool async_path_ok(ScanFwdParams& p, int io_dtype, int bc_dtype) {
    if (io_dtype != MMB_F32 || bc_dtype != MMB_F32) return false;
    if (getenv("MMB_SCAN_SYNC")) return false;
    const auto al = [](const void* q, int a) { return reinterpret_cast<uintptr_t>(q) % a == 0; };
    if (!al(p.u, 4) || !al(p.delta, 4) || !al(p.Bm, 4) || !al(p.Cm, 4)) return false;
    p.flags = 0;
    if (al(p.u, 16) && al(p.delta, 16) && p.u_ds % 4 == 0 && p.u_bs % 4 == 0 && p.d_ds % 4 == 0 && p.d_bs % 4 == 0) p.flags |= 1;
    if (al(p.Bm, 16) && al(p.Cm, 16) && p.B_ls == 1 && p.C_ls == 1 && p.B_ns % 4 == 0 && p.C_ns % 4 == 0 && p.B_bs % 4 == 0 &&
        p.C_bs % 4 == 0 && p.B_gs % 4 == 0 && p.C_gs % 4 == 0) p.flags |= 2;
    return true;
}

template <int S, int T> constexpr size_t scan_fwd_smem() {
    return sizeof(float) * (size_t)(2 * (128 / S) + 2 * kMaxState) * (T + 4);
}

// lanes per row: the smallest split that gives the launch ~2 waves of 16 warps per SM
static int pick_split(int batch, int dim) {
    const long rows = (long)batch * dim;
    const long want_threads = 32L * 16 * num_sms();
    int S = 1;
    while (S < 16 && rows * S < want_threads) S *= 2;
    return S;
}
static int chunk_for_split(int S) { return S == 1 ? 32 : 64; }

template <int S, int T, typename io_t, typename bc_t>
static int launch_scan_fwd(const ScanFwdParams& p, cudaStream_t stream) {
    constexpr size_t smem = scan_fwd_smem<S, T>();
    auto kern = scan_fwd_kernel<S, T, io_t, bc_t>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e);
    }
    const int RT = 128 / S;
    dim3 grid((p.H + RT - 1) / RT, p.G, p.batch);
    kern<<<grid, 128, smem, stream>>>(p);
    return launch_status();
}

template <typename io_t, typename bc_t>
static int dispatch_split(const ScanFwdParams& p, int S, cudaStream_t stream) {
    switch (S) {
        case 0: return launch_scan_fwd<4, 16, io_t, bc_t>(p, stream);
        case 1: return launch_scan_fwd<1, 32, io_t, bc_t>(p, stream);
        case 2: return launch_scan_fwd<2, 64, io_t, bc_t>(p, stream);
        case 4: return launch_scan_fwd<4, 64, io_t, bc_t>(p, stream);
        case 8: return launch_scan_fwd<8, 64, io_t, bc_t>(p, stream);
        default: return launch_scan_fwd<16, 64, io_t, bc_t>(p, stream);
    }
}

template <typename io_t>
static int dispatch_bc(const ScanFwdParams& p, int io_dtype, int bc_dtype, int S, cudaStream_t stream) {
    // B / C come either in the dtype of u (what mamba_ssm does) or in fp32 (what MedMamba passes)
    if (bc_dtype == MMB_F32) return dispatch_split<io_t, float>(p, S, stream);
    if (bc_dtype == io_dtype) return dispatch_split<io_t, io_t>(p, S, stream);
    return MMB_ERR_UNSUPPORTED;
}

}  // namespace mmb

// spacing of the state checkpoints written when chunk_state != NULL (= the backward's chunk)
extern "C" int mmb_scan_chunk_len(void) { return 16; }

extern "C" int mmb_scan_fwd(const void* u, const void* delta, const float* A, const void* Bm, const void* Cm,
                            const float* Dv, const void* z, const float* delta_bias, void* out,
                            float* last_state, float* chunk_state,
                            int batch, int dim, int seqlen, int dstate, int ngroups,
                            int64_t u_bs, int64_t u_ds, int64_t delta_bs, int64_t delta_ds,
                            int64_t z_bs, int64_t z_ds, int64_t out_bs, int64_t out_ds,
                            int64_t B_bs, int64_t B_gs, int64_t B_ns, int64_t B_ls,
                            int64_t C_bs, int64_t C_gs, int64_t C_ns, int64_t C_ls,
                            int delta_softplus, int io_dtype, int bc_dtype, void* stream) {
    using namespace mmb;
    if (!u || !delta || !A || !Bm || !Cm || !out) return MMB_ERR_INVALID_ARG;
    if (batch < 0 || dim <= 0 || seqlen < 0 || dstate <= 0 || ngroups <= 0) return MMB_ERR_INVALID_ARG;
    if (dim % ngroups != 0) return MMB_ERR_INVALID_ARG;
    if (dstate > kMaxState) return MMB_ERR_UNSUPPORTED;
    if (batch > 65535 || ngroups > 65535) return MMB_ERR_UNSUPPORTED;
    if (batch == 0 || seqlen == 0) return MMB_OK;
    ScanFwdParams p;
    p.u = u; p.delta = delta; p.Bm = Bm; p.Cm = Cm; p.z = z; p.out = out;
    p.A = A; p.Dv = Dv; p.bias = delta_bias; p.last_state = last_state; p.chunk_state = chunk_state;
    p.batch = batch; p.dim = dim; p.L = seqlen; p.N = dstate; p.G = ngroups; p.H = dim / ngroups;
    p.softplus = delta_softplus; p.flags = 0;
    p.u_bs = u_bs; p.u_ds = u_ds; p.d_bs = delta_bs; p.d_ds = delta_ds;
    p.z_bs = z_bs; p.z_ds = z_ds; p.o_bs = out_bs; p.o_ds = out_ds;
    p.B_bs = B_bs; p.B_gs = B_gs; p.B_ns = B_ns; p.B_ls = B_ls;
    p.C_bs = C_bs; p.C_gs = C_gs; p.C_ns = C_ns; p.C_ls = C_ls;
    // with checkpoints: 4 lanes per row and 16-step chunks, the geometry mmb_scan_bwd recomputes from
    const int S = chunk_state ? 0 : pick_split(batch, dim);
    const int T = chunk_state ? 16 : chunk_for_split(S);
    p.nchunks = (seqlen + T - 1) / T;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (async_path_ok(p, io_dtype, bc_dtype)) {
        // chunk length: 16 steps for one lane per row (128 rows per CTA: two buffer sets of 46 KB, 4 CTAs per SM)
        if (chunk_state) return launch_scan_fwd_async<4, 16>(p, st);      // the checkpoint geometry: nchunks set above
        const int Ta = S == 1 ? 16 : 32;
        p.nchunks = (seqlen + Ta - 1) / Ta;
        switch (S) {
            case 1: return launch_scan_fwd_async<1, 16>(p, st);
            case 2: return launch_scan_fwd_async<2, 32>(p, st);
            case 4: return launch_scan_fwd_async<4, 32>(p, st);
            case 8: return launch_scan_fwd_async<8, 32>(p, st);
            default: return launch_scan_fwd_async<16, 32>(p, st);
        }
    }
    if (async16_path_ok(p, io_dtype, bc_dtype)) {
        const int Ta = (S == 0 || S == 1) ? 16 : 32;
        p.nchunks = (seqlen + Ta - 1) / Ta;
        if (io_dtype == MMB_BF16)
            return bc_dtype == MMB_F32 ? dispatch_async16<__nv_bfloat16, float>(p, S, st)
                                       : dispatch_async16<__nv_bfloat16, __nv_bfloat16>(p, S, st);
        return bc_dtype == MMB_F32 ? dispatch_async16<__half, float>(p, S, st) : dispatch_async16<__half, __half>(p, S, st);
    }
    switch (io_dtype) {
        case MMB_F32: return dispatch_bc<float>(p, io_dtype, bc_dtype, S, st);
        case MMB_BF16: return dispatch_bc<__nv_bfloat16>(p, io_dtype, bc_dtype, S, st);
        case MMB_F16: return dispatch_bc<__half>(p, io_dtype, bc_dtype, S, st);
        default: return MMB_ERR_INVALID_ARG;
    }
}
