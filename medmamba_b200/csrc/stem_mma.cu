// Patch embedding on tensor cores (autocast path): Conv2d(3 -> E, kernel 4, stride 4) + bias + NCHW->NHWC + LayerNorm(E)
// in one pass (MedMamba.py:54-76).  Per token the convolution is a 48-term contraction (3 channels x 4 x 4 taps)
// against E output channels: a GEMM with K = 48, N = E and one row per token.  Round 1 ran it on the FP32 pipe
// (1.0 ms at batch 1024 against a 0.27 ms HBM floor, issue-bound); under bf16 autocast the reference's convolution runs
// in bf16 with fp32 accumulation anyway, so here the operands are rounded to bf16 and multiplied with
// mma.sync.m16n8k16 (HMMA), fp32 accumulators, the bias + LayerNorm as the epilogue on the accumulator fragments.
//
// The A operand needs no im2col: for a fixed (channel c, tap row r) the 4 taps of consecutive tokens of one token row
// are one contiguous image row.  A strip of TR token rows is staged as 12 chunks [chunk = c*4 + r][token][4 taps] --
// byte for byte the image rows, copied with 1-D bulk copies (UBLKCP) into a double buffer -- and an A fragment
// (row = token, k = 4*chunk + tap) is two adjacent floats of a chunk row: conflict-free 64-bit shared loads.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tma.cuh"

namespace mmb {

constexpr int kPmChunks = 12;           // 3 input channels x 4 tap rows
constexpr int kPmMaxTok = 128;          // tokens per strip (8 warps x 16)

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);      // .x = lo (low half), .y = hi
    return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// NT = E / 8 column tiles.  One warp owns 16 consecutive tokens of the strip and all E channels.
template <int NT>
__global__ void __launch_bounds__(256)
patch_embed_ln_mma_kernel(const float* __restrict__ x, const float* __restrict__ weight, const float* __restrict__ cbias,
                          const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ out,
                          int B, int Hin, int Win, float eps, int TR, int strips_per_img, long total_strips) {
    constexpr int E = 8 * NT;
    constexpr int EP = E + 8;            // pitch of the packed weight rows: (t * EP + g) hits 32 distinct banks
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int Wt = Win >> 2, Ht = Hin >> 2;
    const int ntok = TR * Wt;                                   // tokens of a full strip (<= kPmMaxTok)
    const int cstride = ntok * 4 + 16;                          // floats per chunk: +16 keeps a quad's two chunks on disjoint banks
    const int buf_floats = kPmChunks * cstride;
    float* xs = reinterpret_cast<float*>(smem_raw);             // [2][12][ntok * 4 + 16]
    uint32_t* wsm = reinterpret_cast<uint32_t*>(xs + 2 * buf_floats);      // [24][EP]: bf16 pairs (k even, k odd)
    float* prm = reinterpret_cast<float*>(wsm + 24 * EP);       // [3][E]: conv bias, gamma, beta
    uint64_t* full = reinterpret_cast<uint64_t*>(prm + 3 * E);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;

    // strip st -> image b, first token row i0; its image rows -> buffer `buf`
    auto issue = [&](long st, int buf) {
        const int b = (int)(st / strips_per_img), i0 = (int)(st % strips_per_img) * TR;
        const int nr = min(TR, Ht - i0);
        float* dst = xs + buf * buf_floats;
        mbar_expect_tx(&full[buf], (uint32_t)(nr * kPmChunks * Win * 4));
        for (int tr = 0; tr < nr; ++tr)
            for (int ch = 0; ch < kPmChunks; ++ch) {
                const int c = ch >> 2, r = ch & 3;
                const float* src = x + (((long)b * 3 + c) * Hin + 4 * (i0 + tr) + r) * Win;
                bulk_load_1d(dst + ch * cstride + tr * Wt * 4, src, (uint32_t)(Win * 4), &full[buf]);
            }
    };

    if (tid == 0) {
        mbar_init(&full[0], 1); mbar_init(&full[1], 1);
        mbar_fence_init();
    }
    // weights -> bf16 pairs in shared memory: wsm[kk][e] = (W[e][2kk], W[e][2kk+1]), K index = c*16 + r*4 + s
    for (int i = tid; i < 24 * E; i += blockDim.x) {
        const int kk = i / E, e = i - kk * E;
        wsm[kk * EP + e] = pack_bf16x2(weight[e * 48 + 2 * kk], weight[e * 48 + 2 * kk + 1]);
    }
    for (int i = tid; i < E; i += blockDim.x) {
        prm[i] = cbias ? cbias[i] : 0.f; prm[E + i] = gamma[i]; prm[2 * E + i] = beta[i];
    }
    __syncthreads();
    long st = blockIdx.x;
    if (tid == 0 && st < total_strips) {
        issue(st, 0);
        if (st + gridDim.x < total_strips) issue(st + gridDim.x, 1);
    }

    for (int it = 0; st < total_strips; st += gridDim.x, ++it) {
        const int buf = it & 1, ph = (it >> 1) & 1;
        const int b = (int)(st / strips_per_img), i0 = (int)(st % strips_per_img) * TR;
        const int nvalid = min(TR, Ht - i0) * Wt;               // tokens of this strip
        const float* xb = xs + buf * buf_floats;
        mbar_wait(&full[buf], ph);
        const int m0 = warp * 16;
        if (m0 < nvalid) {
            float acc[NT][4];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
            const int tok0 = min(m0 + g, ntok - 1), tok1 = min(m0 + g + 8, ntok - 1);     // rows past the strip: clamped, discarded
#pragma unroll
            for (int ks = 0; ks < 3; ++ks) {
                // k = 16 ks + 2t (+1): chunk 4 ks + t / 2, taps 2 (t % 2) (+1);  k + 8: chunk + 2
                const float* c0p = xb + (4 * ks + (t >> 1)) * cstride + 2 * (t & 1);
                const float* c1p = c0p + 2 * cstride;
                const float2 v0 = *reinterpret_cast<const float2*>(c0p + tok0 * 4);
                const float2 v1 = *reinterpret_cast<const float2*>(c0p + tok1 * 4);
                const float2 v2 = *reinterpret_cast<const float2*>(c1p + tok0 * 4);
                const float2 v3 = *reinterpret_cast<const float2*>(c1p + tok1 * 4);
                const uint32_t a[4] = {pack_bf16x2(v0.x, v0.y), pack_bf16x2(v1.x, v1.y), pack_bf16x2(v2.x, v2.y),
                                       pack_bf16x2(v3.x, v3.y)};
                const uint32_t* w0 = wsm + (8 * ks + t) * EP + g;
                const uint32_t* w1 = w0 + 4 * EP;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) mma_bf16_16816(acc[nt], a, w0[8 * nt], w1[8 * nt]);
            }
            // epilogue on the fragments: + conv bias, LayerNorm over the E channels of a token (rows g and g + 8 of the
            // tile; a row's E values live in the 4 lanes of a quad), affine, store
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const float2 cb = *reinterpret_cast<const float2*>(prm + 8 * nt + 2 * t);
                acc[nt][0] += cb.x; acc[nt][1] += cb.y; acc[nt][2] += cb.x; acc[nt][3] += cb.y;
                s0 += acc[nt][0] + acc[nt][1]; s1 += acc[nt][2] + acc[nt][3];
            }
            s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
            const float mean0 = s0 * (1.f / E), mean1 = s1 * (1.f / E);
            float q0 = 0.f, q1 = 0.f;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                acc[nt][0] -= mean0; acc[nt][1] -= mean0; acc[nt][2] -= mean1; acc[nt][3] -= mean1;
                q0 = fmaf(acc[nt][0], acc[nt][0], fmaf(acc[nt][1], acc[nt][1], q0));
                q1 = fmaf(acc[nt][2], acc[nt][2], fmaf(acc[nt][3], acc[nt][3], q1));
            }
            q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
            q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
            const float r0 = rsqrtf(q0 * (1.f / E) + eps), r1 = rsqrtf(q1 * (1.f / E) + eps);
            // token `tok` of the strip is token (i0 + tok / Wt, tok % Wt) of image b: consecutive in the NHWC output
            const long obase = ((long)b * Ht + i0) * Wt;
            float* o0 = out + (obase + m0 + g) * E + 2 * t;
            float* o1 = o0 + 8 * E;
            const bool ok0 = m0 + g < nvalid, ok1 = m0 + g + 8 < nvalid;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const float2 gm = *reinterpret_cast<const float2*>(prm + E + 8 * nt + 2 * t);
                const float2 bt = *reinterpret_cast<const float2*>(prm + 2 * E + 8 * nt + 2 * t);
                if (ok0) *reinterpret_cast<float2*>(o0 + 8 * nt) = make_float2(fmaf(acc[nt][0] * r0, gm.x, bt.x), fmaf(acc[nt][1] * r0, gm.y, bt.y));
                if (ok1) *reinterpret_cast<float2*>(o1 + 8 * nt) = make_float2(fmaf(acc[nt][2] * r1, gm.x, bt.x), fmaf(acc[nt][3] * r1, gm.y, bt.y));
            }
        }
        __syncthreads();                                        // every warp is done with this buffer
        if (tid == 0 && st + 2L * gridDim.x < total_strips) issue(st + 2L * gridDim.x, buf);
    }
}

}  // namespace mmb

// Tensor-core path of mmb_patch_embed_ln_fwd (math_mode 1, fp32 NCHW images): see include/medmamba_b200.h.
int mmb_patch_embed_ln_mma(const float* x, const float* weight, const float* conv_bias, const float* gamma, const float* beta,
                           float* out, int batch, int Hin, int Win, int embed_dim, float eps, cudaStream_t st) {
    using namespace mmb;
    const int Wt = Win / 4, Ht = Hin / 4;
    if (Wt > kPmMaxTok || Win % 4 != 0 || embed_dim % 32 != 0 || embed_dim > 128) return MMB_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) % 16 != 0) return MMB_ERR_UNSUPPORTED;
    int TR = kPmMaxTok / Wt;
    if (TR > Ht) TR = Ht;
    if (TR > 4) TR = 4;
    const int ntok = TR * Wt;
    const int warps = (ntok + 15) / 16;
    const int strips_per_img = (Ht + TR - 1) / TR;
    const long total = (long)batch * strips_per_img;
    const size_t smem = (size_t)2 * kPmChunks * (ntok * 16 + 64) + (size_t)24 * (embed_dim + 8) * 4 + (size_t)3 * embed_dim * 4 + 16;
    int grid = 3 * num_sms();
    if (total < grid) grid = (int)total;
#define MMB_PM(NTV)                                                                                                \
    do {                                                                                                           \
        auto kern = patch_embed_ln_mma_kernel<NTV>;                                                                \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
        if (e != cudaSuccess) return cuda_status(e);                                                               \
        kern<<<grid, warps * 32, smem, st>>>(x, weight, conv_bias, gamma, beta, out, batch, Hin, Win, eps, TR,     \
                                             strips_per_img, total);                                              \
        return launch_status();                                                                                    \
    } while (0)
    switch (embed_dim / 32) {
        case 1: MMB_PM(4);
        case 2: MMB_PM(8);
        case 3: MMB_PM(12);
        default: MMB_PM(16);
    }
#undef MMB_PM
}
