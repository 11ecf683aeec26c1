// The steps between stages, channels-last, inference path:
//   patch_embed_ln : Conv2d(3 -> E, kernel 4, stride 4) + NCHW->NHWC permute + LayerNorm(E)   MedMamba.py:54-76
//   patch_merge_ln : 2x2 neighbourhood gather + concat + LayerNorm(4C)                        MedMamba.py:93-117
// The reference runs them as cuDNN conv / 4 strided slices + cat, a permute copy and a LayerNorm: 3 to 6 passes
// over the largest activations of the network.  Here each is one kernel that reads its input once and writes
// the normalised tokens once.
#include "common.cuh"

namespace mmb {

// ------------------------------------------------------------------------------------------------
// Patch embedding.  A CTA of 4 warps owns 56 tokens of one output row (b, i): the 3 x 4 input row pieces it
// needs are staged in shared memory with coalesced 128-bit loads; the weights (E x 48, row pitch 52 floats:
// conflict-free LDS.128 of 4 consecutive taps) are staged once per CTA, CTAs are persistent over the units.
// A warp takes TT = 14 tokens; lane l owns output channels l, l+32, .. (M = E/32 of them) of all 14: per 4 taps
// M weight LDS.128 + 14 broadcast input LDS.128 feed 56 M FMAs, issued as packed FFMA2 over token pairs (the
// strip is stored pair-interleaved for that).  Measured at batch 1024: 1.43 ms for the first version, 1.01 ms
// now; 8 tokens per warp at 28 warps per SM is no faster (1.06 ms) -- see profiles/README.md.
// LayerNorm of a token is a warp reduction over the lanes' M channels (two-pass, like layernorm_fwd_kernel).
constexpr int kPeTaps = 48;      // 3 channels x 4 x 4
constexpr int kPePitch = 52;
constexpr int kPeTT = 14;
constexpr int kPeWarps = 4;
constexpr int kPeThreads = 32 * kPeWarps;
constexpr int kPeCtasPerSm = 4;
constexpr int kPeChunk = kPeWarps * kPeTT;   // tokens per work unit: one group of kPeTT per warp

template <int M, typename in_t>
__global__ void __launch_bounds__(kPeThreads, kPeCtasPerSm)
patch_embed_ln_kernel(const in_t* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ cbias,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ out,
                      int B, int Hin, int Win, float eps) {
    constexpr int E = 32 * M;
    extern __shared__ __align__(16) float smem[];
    float* sw = smem;                       // [E][kPePitch]
    float* sx = smem + E * kPePitch;        // [12][kPeChunk] float4
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Ho = Hin / 4, Wo = Win / 4;
    for (int i = tid; i < E * kPeTaps; i += kPeThreads) sw[(i / kPeTaps) * kPePitch + i % kPeTaps] = __ldg(wgt + i);
    float cb[M], gm[M], bt[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const int oc = lane + 32 * m;
        cb[m] = cbias ? __ldg(cbias + oc) : 0.f;
        gm[m] = __ldg(gamma + oc);
        bt[m] = __ldg(beta + oc);
    }
    // Work unit: (image b, output row i, chunk of kPeChunk = 56 tokens = 224 input columns): 4 warps x 14 tokens.
    // The strip of the next unit travels global -> registers while this one is computed, registers -> shared
    // memory afterwards, so the DRAM latency hides behind the FMAs (a strip is 12 x 56 float4: 6 per thread).
    constexpr int kPeStage = (12 * kPeChunk + kPeThreads - 1) / kPeThreads;
    const int nchunks = (Wo + kPeChunk - 1) / kPeChunk;
    const int nunits = B * Ho * nchunks;
    float4 nxt[kPeStage];
    auto fetch = [&](int unit) {
        const int ch = unit % nchunks, row = unit / nchunks;
        const int b = row / Ho, i = row - b * Ho;
        const int w4n = min(kPeChunk, Wo - ch * kPeChunk);          // float4 (= tokens) per strip row
#pragma unroll
        for (int u = 0; u < kPeStage; ++u) {
            const int idx = tid + u * kPeThreads;
            const int rr = idx / kPeChunk, w4 = idx - rr * kPeChunk;      // rr = c * 4 + r
            if (rr < 12 && w4 < w4n) {
                const in_t* src = x + (((int64_t)b * 3 + (rr >> 2)) * Hin + 4 * i + (rr & 3)) * Win + 4 * (ch * kPeChunk + w4);
                nxt[u] = load4<in_t>(src);
            } else {
                nxt[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };
    auto park = [&]() {
#pragma unroll
        for (int u = 0; u < kPeStage; ++u) {
            const int idx = tid + u * kPeThreads;
            const int rr = idx / kPeChunk, w4 = idx - rr * kPeChunk;
            if (rr < 12) {
                // pair layout: tokens (2p, 2p+1) interleaved per tap, so one LDS.128 yields two packed FMA operands
                float* d = sx + ((rr * (kPeChunk / 2) + (w4 >> 1)) * 4) * 2 + (w4 & 1);
                d[0] = nxt[u].x; d[2] = nxt[u].y; d[4] = nxt[u].z; d[6] = nxt[u].w;
            }
        }
    };
    if ((int)blockIdx.x < nunits) fetch(blockIdx.x);
    for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
        const int ch = unit % nchunks, row = unit / nchunks;
        const int b = row / Ho, i = row - b * Ho;
        const int ntok = min(kPeChunk, Wo - ch * kPeChunk);
        __syncthreads();                     // previous unit fully consumed (and the weights staged)
        park();
        __syncthreads();
        if (unit + (int)gridDim.x < nunits) fetch(unit + gridDim.x);
        const int t0 = warp * kPeTT;         // first token of this warp inside the chunk
        if (t0 < ntok) {
            float acc[kPeTT][M];
#pragma unroll
            for (int t = 0; t < kPeTT; ++t)
#pragma unroll
                for (int m = 0; m < M; ++m) acc[t][m] = cb[m];
#pragma unroll 1
            for (int rr = 0; rr < 12; ++rr) {
                float4 wv[M];
#pragma unroll
                for (int m = 0; m < M; ++m)
                    wv[m] = *reinterpret_cast<const float4*>(sw + (lane + 32 * m) * kPePitch + 4 * rr);
                // tokens past the end of a ragged chunk read stale strip data; their results are never stored.
                // Two tokens per packed FMA (FFMA2: the kernel is issue-bound with scalar FFMA), tap-major.
                const float4* sp = reinterpret_cast<const float4*>(sx) + (rr * (kPeChunk / 2) + (t0 >> 1)) * 2;
#pragma unroll
                for (int t = 0; t < kPeTT; t += 2) {
                    const float4 q0 = sp[t], q1 = sp[t + 1];      // (a.x b.x a.y b.y), (a.z b.z a.w b.w)
#pragma unroll
                    for (int m = 0; m < M; ++m) fma2(acc[t][m], acc[t + 1][m], wv[m].x, wv[m].x, q0.x, q0.y, acc[t][m], acc[t + 1][m]);
#pragma unroll
                    for (int m = 0; m < M; ++m) fma2(acc[t][m], acc[t + 1][m], wv[m].y, wv[m].y, q0.z, q0.w, acc[t][m], acc[t + 1][m]);
#pragma unroll
                    for (int m = 0; m < M; ++m) fma2(acc[t][m], acc[t + 1][m], wv[m].z, wv[m].z, q1.x, q1.y, acc[t][m], acc[t + 1][m]);
#pragma unroll
                    for (int m = 0; m < M; ++m) fma2(acc[t][m], acc[t + 1][m], wv[m].w, wv[m].w, q1.z, q1.w, acc[t][m], acc[t + 1][m]);
                }
            }
            float sum[kPeTT], sq[kPeTT];
#pragma unroll
            for (int t = 0; t < kPeTT; ++t) {
                sum[t] = 0.f;
#pragma unroll
                for (int m = 0; m < M; ++m) sum[t] += acc[t][m];
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1)
#pragma unroll
                for (int t = 0; t < kPeTT; ++t) sum[t] += __shfl_xor_sync(0xffffffffu, sum[t], off);
#pragma unroll
            for (int t = 0; t < kPeTT; ++t) {
                const float mean = sum[t] / (float)E;
                sq[t] = 0.f;
#pragma unroll
                for (int m = 0; m < M; ++m) { acc[t][m] -= mean; sq[t] = fmaf(acc[t][m], acc[t][m], sq[t]); }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1)
#pragma unroll
                for (int t = 0; t < kPeTT; ++t) sq[t] += __shfl_xor_sync(0xffffffffu, sq[t], off);
#pragma unroll
            for (int t = 0; t < kPeTT; ++t) {
                if (t0 + t < ntok) {
                    const float rstd = rsqrtf(sq[t] / (float)E + eps);
                    float* o = out + (((int64_t)b * Ho + i) * Wo + ch * kPeChunk + t0 + t) * E + lane;
#pragma unroll
                    for (int m = 0; m < M; ++m) o[32 * m] = fmaf(acc[t][m] * rstd, gm[m], bt[m]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Patch merging: token (b, i, j) of the half-resolution grid is the concatenation of the pixels
// (2i, 2j), (2i+1, 2j), (2i, 2j+1), (2i+1, 2j+1) -- x0, x1, x2, x3 of MedMamba.py:100-113 -- followed by
// LayerNorm over the 4C channels.  Warp per token; odd trailing rows / columns are dropped like the reference.
template <int V, typename in_t, typename out_t>
__global__ void __launch_bounds__(256)
patch_merge_ln_kernel(const in_t* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                      out_t* __restrict__ out, int B, int H, int W, int C, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int h2 = H / 2, w2 = W / 2, D = 4 * C, C4 = C / 4, D4 = C;
    const int64_t tokens = (int64_t)B * h2 * w2;
    for (int64_t tok = warp; tok < tokens; tok += nwarps) {
        const int j = (int)(tok % w2);
        const int64_t r = tok / w2;
        const int i = (int)(r % h2), b = (int)(r / h2);
        const in_t* base = x + (((int64_t)b * H + 2 * i) * W + 2 * j) * C;
        float4 v[V];
        float sum = 0.f;
#pragma unroll
        for (int u = 0; u < V; ++u) {
            const int d4 = lane + 32 * u;
            if (d4 < D4) {
                const int quad = d4 / C4, c4 = d4 - quad * C4;
                v[u] = load4<in_t>(base + ((int64_t)(quad & 1) * W + (quad >> 1)) * C + 4 * c4);
            } else {
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            sum += (v[u].x + v[u].y) + (v[u].z + v[u].w);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        const float mean = sum / (float)D;
        float sq = 0.f;
#pragma unroll
        for (int u = 0; u < V; ++u) {
            if (lane + 32 * u < D4) {
                v[u].x -= mean; v[u].y -= mean; v[u].z -= mean; v[u].w -= mean;
                sq += (v[u].x * v[u].x + v[u].y * v[u].y) + (v[u].z * v[u].z + v[u].w * v[u].w);
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
        const float rstd = rsqrtf(sq / (float)D + eps);
#pragma unroll
        for (int u = 0; u < V; ++u) {
            const int d4 = lane + 32 * u;
            if (d4 < D4) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + d4);
                const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + d4);
                float4 o;
                o.x = fmaf(v[u].x * rstd, g.x, bt.x); o.y = fmaf(v[u].y * rstd, g.y, bt.y);
                o.z = fmaf(v[u].z * rstd, g.z, bt.z); o.w = fmaf(v[u].w * rstd, g.w, bt.w);
                store4<out_t>(out + tok * D + 4 * d4, o);
            }
        }
    }
}

template <typename T> static bool aligned4(const void* p) { return reinterpret_cast<uintptr_t>(p) % vec4_align<T>() == 0; }

}  // namespace mmb

int mmb_patch_embed_ln_mma(const float* x, const float* weight, const float* conv_bias, const float* gamma, const float* beta,
                           float* out, int batch, int Hin, int Win, int embed_dim, float eps, cudaStream_t st);   // stem_mma.cu

extern "C" int mmb_patch_embed_ln_fwd(const void* x, const float* weight, const float* conv_bias, const float* gamma,
                                      const float* beta, float* out, int batch, int Hin, int Win, int embed_dim,
                                      float eps, int in_dtype, int math_mode, void* stream) {
    using namespace mmb;
    if (!x || !weight || !gamma || !beta || !out) return MMB_ERR_INVALID_ARG;
    if (batch < 0 || Hin <= 0 || Win <= 0 || embed_dim <= 0) return MMB_ERR_INVALID_ARG;
    if (Hin % 4 != 0 || Win % 4 != 0 || embed_dim % 32 != 0 || embed_dim > 128) return MMB_ERR_UNSUPPORTED;
    if (reinterpret_cast<uintptr_t>(out) % 4 != 0) return MMB_ERR_UNSUPPORTED;
    if (math_mode != 0 && math_mode != 1) return MMB_ERR_INVALID_ARG;
    if (batch == 0) return MMB_OK;
    if (math_mode == 1 && in_dtype == MMB_F32 && Win / 4 <= 128 && reinterpret_cast<uintptr_t>(x) % 16 == 0 &&
        reinterpret_cast<uintptr_t>(out) % 16 == 0)
        return mmb_patch_embed_ln_mma(reinterpret_cast<const float*>(x), weight, conv_bias, gamma, beta, out, batch, Hin, Win,
                                      embed_dim, eps, reinterpret_cast<cudaStream_t>(stream));
    const size_t smem = sizeof(float) * ((size_t)embed_dim * kPePitch + 12 * 4 * (size_t)kPeChunk);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long units = (long)batch * (Hin / 4) * ((Win / 4 + kPeChunk - 1) / kPeChunk);
    const int grid = (int)(units < (long)kPeCtasPerSm * num_sms() ? units : (long)kPeCtasPerSm * num_sms());
#define MMB_PE(M, TI)                                                                                             \
    do {                                                                                                          \
        if (!aligned4<TI>(x)) return MMB_ERR_UNSUPPORTED;                                                         \
        auto kern = patch_embed_ln_kernel<M, TI>;                                                                 \
        if (smem > 48 * 1024) {                                                                                   \
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
            if (e != cudaSuccess) return cuda_status(e);                                                          \
        }                                                                                                         \
        kern<<<grid, kPeThreads, smem, st>>>(reinterpret_cast<const TI*>(x), weight, conv_bias, gamma, beta, out, batch, \
                                      Hin, Win, eps);                                                             \
        return launch_status();                                                                                   \
    } while (0)
#define MMB_PE_M(TI)                                                                                              \
    do {                                                                                                          \
        switch (embed_dim / 32) {                                                                                 \
            case 1: MMB_PE(1, TI);                                                                                \
            case 2: MMB_PE(2, TI);                                                                                \
            case 3: MMB_PE(3, TI);                                                                                \
            default: MMB_PE(4, TI);                                                                               \
        }                                                                                                         \
    } while (0)
    if (in_dtype == MMB_F32) MMB_PE_M(float);
    if (in_dtype == MMB_BF16) MMB_PE_M(__nv_bfloat16);
#undef MMB_PE_M
#undef MMB_PE
    return MMB_ERR_UNSUPPORTED;
}

extern "C" int mmb_patch_merge_ln_fwd(const void* x, const float* gamma, const float* beta, void* out, int batch, int H,
                                      int W, int C, float eps, int in_dtype, int out_dtype, void* stream) {
    using namespace mmb;
    if (!x || !gamma || !beta || !out) return MMB_ERR_INVALID_ARG;
    if (batch < 0 || H <= 0 || W <= 0 || C <= 0) return MMB_ERR_INVALID_ARG;
    if (C % 4 != 0 || C > 512) return MMB_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) % 16 != 0) return MMB_ERR_UNSUPPORTED;
    const int64_t tokens = (int64_t)batch * (H / 2) * (W / 2);
    if (tokens == 0) return MMB_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t want = (tokens * 32 + 255) / 256;
    const int grid = (int)(want < 16L * num_sms() ? want : 16L * num_sms());
#define MMB_PM(V, TI, TO)                                                                                         \
    do {                                                                                                          \
        if (!aligned4<TI>(x) || !aligned4<TO>(out)) return MMB_ERR_UNSUPPORTED;                                   \
        patch_merge_ln_kernel<V, TI, TO><<<grid, 256, 0, st>>>(reinterpret_cast<const TI*>(x), gamma, beta,       \
            reinterpret_cast<TO*>(out), batch, H, W, C, eps);                                                     \
        return launch_status();                                                                                   \
    } while (0)
#define MMB_PM_V(TI, TO)                                                                                          \
    do {                                                                                                          \
        if (C <= 128) MMB_PM(4, TI, TO);                                                                          \
        if (C <= 256) MMB_PM(8, TI, TO);                                                                          \
        MMB_PM(16, TI, TO);                                                                                       \
    } while (0)
    if (in_dtype == MMB_F32 && out_dtype == MMB_F32) MMB_PM_V(float, float);
    if (in_dtype == MMB_F32 && out_dtype == MMB_BF16) MMB_PM_V(float, __nv_bfloat16);
    if (in_dtype == MMB_BF16 && out_dtype == MMB_BF16) MMB_PM_V(__nv_bfloat16, __nv_bfloat16);
    if (in_dtype == MMB_BF16 && out_dtype == MMB_F32) MMB_PM_V(__nv_bfloat16, float);
#undef MMB_PM_V
#undef MMB_PM
    return MMB_ERR_UNSUPPORTED;
}
