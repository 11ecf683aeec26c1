// The bandwidth-bound kernels either side of the scan, all channels-last (B, H, W, C):
//   dwconv3x3 + bias + SiLU        replaces MedMamba.py:294-295 (permute copy, cuDNN depthwise conv, SiLU)
//   (y0+y2+y1+y3) -> LayerNorm -> * SiLU(z)   replaces MedMamba.py:298-301 (3 adds, transpose copy, LN, gate)
//   cat + channel_shuffle(2) + residual       replaces MedMamba.py:355-357 and :308-320
// Each reads its inputs once with 128-bit loads and writes its output once.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace mmb {

// ------------------------------------------------------------------------------------------------
// Depthwise 3x3, padding 1, + bias, SiLU.  x is a channels-last view with an arbitrary pixel pitch
// (it is the first half of the in_proj output, pitch 2*D); out is dense (B, H, W, D).
// A thread owns 4 channels and a strip of WS output pixels along w: 3 x (WS + 2) float4 loads.
template <int WS, typename in_t, typename out_t, bool ACT = true, bool FLIP = false>
__global__ void __launch_bounds__(256)
dwconv3x3_silu_kernel(const in_t* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ bias,
                      out_t* __restrict__ out, int B, int H, int W, int D, int64_t x_pix, int64_t x_batch,
                      int64_t o_pix) {
    // weights transposed into shared memory once per CTA: swt[tap][D], so a thread's 4 channels are one LDS.128
    extern __shared__ __align__(16) float swt[];
    for (int i = threadIdx.x; i < D * 9; i += blockDim.x) {
        const int ch = i / 9, tp = i % 9;
        swt[(FLIP ? 8 - tp : tp) * D + ch] = __ldg(wgt + i);
    }
    __syncthreads();
    const int C4 = D / 4;
    const int strips = (W + WS - 1) / WS;
    // 32-bit index arithmetic: the host rejects problems with more than 2^31 work items
    const uint32_t total = (uint32_t)B * H * strips * C4;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % (uint32_t)C4);
        uint32_t r = idx / (uint32_t)C4;
        const int st = (int)(r % (uint32_t)strips); r /= (uint32_t)strips;
        const int h = (int)(r % (uint32_t)H);
        const int b = (int)(r / (uint32_t)H);
        const int c = c4 * 4, w0 = st * WS;
        float wk[9][4];
#pragma unroll
        for (int tp = 0; tp < 9; ++tp) {
            const float4 wv = *reinterpret_cast<const float4*>(swt + tp * D + c);
            wk[tp][0] = wv.x; wk[tp][1] = wv.y; wk[tp][2] = wv.z; wk[tp][3] = wv.w;
        }
        float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias) bs = __ldg(reinterpret_cast<const float4*>(bias + c));
        float acc[WS][4];
#pragma unroll
        for (int i = 0; i < WS; ++i) { acc[i][0] = bs.x; acc[i][1] = bs.y; acc[i][2] = bs.z; acc[i][3] = bs.w; }
        const in_t* xb = x + (int64_t)b * x_batch + c;
        // All 3 x (WS + 2) neighbour loads are issued unconditionally from clamped coordinates (branches around
        // them would serialise their latencies); out-of-image taps are zeroed afterwards.
        float4 nb[3][WS + 2];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int hy = min(max(h + dy - 1, 0), H - 1);
#pragma unroll
            for (int j = 0; j < WS + 2; ++j) {
                const int wx = min(max(w0 + j - 1, 0), W - 1);
                nb[dy][j] = load4<in_t>(xb + ((int64_t)hy * W + wx) * x_pix);
            }
        }
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const bool hok = (h + dy - 1 >= 0) && (h + dy - 1 < H);
#pragma unroll
            for (int j = 0; j < WS + 2; ++j) {
                const bool ok = hok && (w0 + j - 1 >= 0) && (w0 + j - 1 < W);
                const float vv[4] = {ok ? nb[dy][j].x : 0.f, ok ? nb[dy][j].y : 0.f, ok ? nb[dy][j].z : 0.f,
                                     ok ? nb[dy][j].w : 0.f};
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int i = j - dx;          // output pixel this input contributes to with tap (dy, dx)
                    if (i >= 0 && i < WS) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[i][e] = fmaf(wk[dy * 3 + dx][e], vv[e], acc[i][e]);
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < WS; ++i) {
            const int wx = w0 + i;
            if (wx < W) {
                float4 o;
                if (ACT) { o.x = silu_f(acc[i][0]); o.y = silu_f(acc[i][1]); o.z = silu_f(acc[i][2]); o.w = silu_f(acc[i][3]); }
                else { o.x = acc[i][0]; o.y = acc[i][1]; o.z = acc[i][2]; o.w = acc[i][3]; }
                store4<out_t>(out + (((int64_t)b * H + h) * W + wx) * o_pix + c, o);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// bf16 in / bf16 out build of the same convolution (the autocast layout): a thread owns 8 channels -- one 128-bit
// load per neighbour instead of a 64-bit one -- and a strip of WS output pixels; weights come from shared memory
// one row of taps at a time (72 of them would not fit in registers beside the 8 x WS accumulators).
template <int WS>
__global__ void __launch_bounds__(256, 2)
dwconv3x3_silu_bf16x8_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ wgt,
                             const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int B, int H, int W, int D,
                             int64_t x_pix, int64_t x_batch) {
    extern __shared__ __align__(16) float swt[];        // [tap][D]
    for (int i = threadIdx.x; i < D * 9; i += blockDim.x) swt[(i % 9) * D + i / 9] = __ldg(wgt + i);
    __syncthreads();
    const int C8 = D / 8;
    const int strips = (W + WS - 1) / WS;
    const uint32_t total = (uint32_t)B * H * strips * C8;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int c8 = (int)(idx % (uint32_t)C8);
        uint32_t r = idx / (uint32_t)C8;
        const int st = (int)(r % (uint32_t)strips); r /= (uint32_t)strips;
        const int h = (int)(r % (uint32_t)H);
        const int b = (int)(r / (uint32_t)H);
        const int c = c8 * 8, w0 = st * WS;
        float acc[WS][8];
        {
            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
            if (bias) { b0 = __ldg(reinterpret_cast<const float4*>(bias + c)); b1 = __ldg(reinterpret_cast<const float4*>(bias + c + 4)); }
#pragma unroll
            for (int i = 0; i < WS; ++i) {
                acc[i][0] = b0.x; acc[i][1] = b0.y; acc[i][2] = b0.z; acc[i][3] = b0.w;
                acc[i][4] = b1.x; acc[i][5] = b1.y; acc[i][6] = b1.z; acc[i][7] = b1.w;
            }
        }
        const __nv_bfloat16* xb = x + (int64_t)b * x_batch + c;
        // all 3 x (WS + 2) neighbour loads issued up front from clamped coordinates, masked afterwards
        uint4 nb[3][WS + 2];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int hy = min(max(h + dy - 1, 0), H - 1);
#pragma unroll
            for (int j = 0; j < WS + 2; ++j) {
                const int wx = min(max(w0 + j - 1, 0), W - 1);
                nb[dy][j] = __ldg(reinterpret_cast<const uint4*>(xb + ((int64_t)hy * W + wx) * x_pix));
            }
        }
        // strips that touch no image border (most of them) skip the zero masks of the out-of-image taps
        const bool interior = h >= 1 && h + 1 < H && w0 >= 1 && w0 + WS < W;
        auto taps = [&](auto interior_tag) {
            constexpr bool INTERIOR = decltype(interior_tag)::value;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const bool hok = INTERIOR || ((h + dy - 1 >= 0) && (h + dy - 1 < H));
                float wk[3][8];
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const float4 w0v = *reinterpret_cast<const float4*>(swt + (dy * 3 + dx) * D + c);
                    const float4 w1v = *reinterpret_cast<const float4*>(swt + (dy * 3 + dx) * D + c + 4);
                    wk[dx][0] = w0v.x; wk[dx][1] = w0v.y; wk[dx][2] = w0v.z; wk[dx][3] = w0v.w;
                    wk[dx][4] = w1v.x; wk[dx][5] = w1v.y; wk[dx][6] = w1v.z; wk[dx][7] = w1v.w;
                }
#pragma unroll
                for (int j = 0; j < WS + 2; ++j) {
                    uint32_t wds[4] = {nb[dy][j].x, nb[dy][j].y, nb[dy][j].z, nb[dy][j].w};
                    if (!INTERIOR) {
                        const bool ok = hok && (w0 + j - 1 >= 0) && (w0 + j - 1 < W);
                        const uint32_t mk = ok ? 0xffffffffu : 0u;    // out-of-image taps: zero the packed words (4 ops, not 8)
#pragma unroll
                        for (int e = 0; e < 4; ++e) wds[e] &= mk;
                    }
                    float vv[8];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        // bf16 -> fp32 is a shift: low half is element 2e, high half element 2e + 1
                        vv[2 * e] = __uint_as_float(wds[e] << 16);
                        vv[2 * e + 1] = __uint_as_float(wds[e] & 0xffff0000u);
                    }
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const int i = j - dx;
                        if (i >= 0 && i < WS) {
#pragma unroll
                            for (int e = 0; e < 8; e += 2)       // two channels per packed FFMA2
                                fma2(acc[i][e], acc[i][e + 1], wk[dx][e], wk[dx][e + 1], vv[e], vv[e + 1], acc[i][e], acc[i][e + 1]);
                        }
                    }
                }
            }
        };
        if (interior) taps(std::true_type{}); else taps(std::false_type{});
#pragma unroll
        for (int i = 0; i < WS; ++i) {
            const int wx = w0 + i;
            if (wx < W) {
                uint4 o;
                uint32_t* ow = &o.x;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const __nv_bfloat162 pr = __floats2bfloat162_rn(silu_f(acc[i][2 * e]), silu_f(acc[i][2 * e + 1]));
                    ow[e] = *reinterpret_cast<const uint32_t*>(&pr);
                }
                *reinterpret_cast<uint4*>(out + (((int64_t)b * H + h) * W + wx) * D + c) = o;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Two output rows per thread (bf16 in / bf16 out).  ncu of the bf16x8 kernel above (profiles/r2s3_dwconv_fwd_s1_metrics.txt):
// 30.6 thread instructions per output, 45 % of them integer / logic -- the bf16 -> fp32 shifts of 18 neighbour vectors, their
// clamped 64-bit addresses and three run-time divisions per thread -- at 54 % issue-slot utilisation: the kernel is bound by
// instructions per output.  Here a thread owns 4 channels x 4 pixels x 2 rows: the two output rows share two of their three
// input rows, so 24 neighbour vectors (4 rows x 6 columns) serve 32 outputs instead of 36 -- a third fewer loads, unpack shifts
// and addresses per output; the index is decomposed with multiply-high by host-computed reciprocals.  Per output the taps are
// accumulated in the order of the kernel above (tap row, then column, then tap column), so results are bit-identical to it.
__global__ void __launch_bounds__(256, 2)
dwconv3x3_silu_bf16_rows_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ wgt,
                                const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int H, int W, int D,
                                int64_t x_pix, int64_t x_batch, uint32_t magic_c4, uint32_t magic_strips) {
    constexpr int WS = 4, RS = 2;
    extern __shared__ __align__(16) float swt[];        // [tap][D]
    for (int i = threadIdx.x; i < D * 9; i += blockDim.x) swt[(i % 9) * D + i / 9] = __ldg(wgt + i);
    __syncthreads();
    const int C4 = D / 4;
    const int strips = (W + WS - 1) / WS;
    const uint32_t per_img = (uint32_t)((H + RS - 1) / RS) * strips * C4;
    const int b = blockIdx.y;
    const __nv_bfloat16* xi = x + (int64_t)b * x_batch;
    __nv_bfloat16* oi = out + (int64_t)b * H * W * D;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < per_img; idx += gridDim.x * blockDim.x) {
        const uint32_t t = C4 == 1 ? idx : __umulhi(idx, magic_c4);
        const int c = (int)(idx - t * (uint32_t)C4) * 4;
        const uint32_t rp = strips == 1 ? t : __umulhi(t, magic_strips);     // 2^32 / 1 does not fit the 32-bit reciprocal
        const int w0 = (int)(t - rp * (uint32_t)strips) * WS, h0 = (int)rp * RS;
        float acc[RS][WS][4];
        {
            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bias) b0 = __ldg(reinterpret_cast<const float4*>(bias + c));
#pragma unroll
            for (int r = 0; r < RS; ++r)
#pragma unroll
                for (int i = 0; i < WS; ++i) { acc[r][i][0] = b0.x; acc[r][i][1] = b0.y; acc[r][i][2] = b0.z; acc[r][i][3] = b0.w; }
        }
        // all (RS + 2) x (WS + 2) neighbour loads issued up front from clamped coordinates, masked afterwards
        uint2 nb[RS + 2][WS + 2];
        int coff[WS + 2];
#pragma unroll
        for (int j = 0; j < WS + 2; ++j) coff[j] = min(max(w0 + j - 1, 0), W - 1) * (int)x_pix;
#pragma unroll
        for (int ir = 0; ir < RS + 2; ++ir) {
            const int hy = min(max(h0 + ir - 1, 0), H - 1);
            const __nv_bfloat16* rowp = xi + (int64_t)hy * W * x_pix + c;
#pragma unroll
            for (int j = 0; j < WS + 2; ++j) nb[ir][j] = __ldg(reinterpret_cast<const uint2*>(rowp + coff[j]));
        }
        const bool interior = h0 >= 1 && h0 + RS < H && w0 >= 1 && w0 + WS < W;
        auto taps = [&](auto interior_tag) {
            constexpr bool INTERIOR = decltype(interior_tag)::value;
#pragma unroll
            for (int ir = 0; ir < RS + 2; ++ir) {
                const bool hok = INTERIOR || ((h0 + ir - 1 >= 0) && (h0 + ir - 1 < H));
                float vv[WS + 2][4];
#pragma unroll
                for (int j = 0; j < WS + 2; ++j) {
                    uint32_t lo = nb[ir][j].x, hi = nb[ir][j].y;
                    if (!INTERIOR) {
                        const bool ok = hok && (w0 + j - 1 >= 0) && (w0 + j - 1 < W);
                        const uint32_t mk = ok ? 0xffffffffu : 0u;
                        lo &= mk; hi &= mk;
                    }
                    vv[j][0] = __uint_as_float(lo << 16); vv[j][1] = __uint_as_float(lo & 0xffff0000u);
                    vv[j][2] = __uint_as_float(hi << 16); vv[j][3] = __uint_as_float(hi & 0xffff0000u);
                }
#pragma unroll
                for (int r = 0; r < RS; ++r) {
                    const int dy = ir - r;                   // tap row through which input row ir reaches output row r
                    if (dy < 0 || dy > 2) continue;
                    float4 wk[3];
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) wk[dx] = *reinterpret_cast<const float4*>(swt + (dy * 3 + dx) * D + c);
#pragma unroll
                    for (int j = 0; j < WS + 2; ++j) {
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const int i = j - dx;
                            if (i >= 0 && i < WS) {
                                fma2(acc[r][i][0], acc[r][i][1], wk[dx].x, wk[dx].y, vv[j][0], vv[j][1], acc[r][i][0], acc[r][i][1]);
                                fma2(acc[r][i][2], acc[r][i][3], wk[dx].z, wk[dx].w, vv[j][2], vv[j][3], acc[r][i][2], acc[r][i][3]);
                            }
                        }
                    }
                }
            }
        };
        if (interior) taps(std::true_type{}); else taps(std::false_type{});
#pragma unroll
        for (int r = 0; r < RS; ++r) {
            const int hh = h0 + r;
#pragma unroll
            for (int i = 0; i < WS; ++i) {
                const int wx = w0 + i;
                if (hh < H && wx < W) {
                    const __nv_bfloat162 p0 = __floats2bfloat162_rn(silu_f(acc[r][i][0]), silu_f(acc[r][i][1]));
                    const __nv_bfloat162 p1 = __floats2bfloat162_rn(silu_f(acc[r][i][2]), silu_f(acc[r][i][3]));
                    uint2 o;
                    o.x = *reinterpret_cast<const uint32_t*>(&p0);
                    o.y = *reinterpret_cast<const uint32_t*>(&p1);
                    *reinterpret_cast<uint2*>(oi + ((int64_t)hh * W + wx) * D + c) = o;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// One warp per token: y = ((y0 + y2) + y1) + y3 over the four direction slices of ydir (B, L, 4, D)
// -- the order of `y1 + y2 + y3 + y4` at MedMamba.py:298, whose operands are out0, flipped out2,
// transposed out1, flipped-transposed out3 (MedMamba.py:286) -- then LayerNorm over D and * SiLU(z).
// y_t = float: the slices hold y_k = <C, h> + D_k u.  y_t = bf16 (autocast path): the slices hold the state terms only
// and the skip term u * sum_k D_k is added here in fp32 from xc.
template <int V, typename z_t, typename out_t, typename y_t>   // V float4 per lane: D <= 128 * V
__global__ void __launch_bounds__(256)
outnorm_gate_kernel(const y_t* __restrict__ ydir, const z_t* __restrict__ z, const float* __restrict__ gamma,
                    const float* __restrict__ beta, out_t* __restrict__ out, float* __restrict__ ymerged,
                    const __nv_bfloat16* __restrict__ xc, const float* __restrict__ Dsum,
                    int64_t tokens, int D, int64_t z_pix, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int C4 = D / 4;
    for (int64_t tok = warp; tok < tokens; tok += nwarps) {
        const y_t* y0 = ydir + tok * 4 * D;
        float4 v[V];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int c4 = lane + 32 * i;
            if (c4 < C4) {
                const float4 a = load4<y_t>(y0 + 4 * c4), bq = load4<y_t>(y0 + D + 4 * c4), cq = load4<y_t>(y0 + 2 * D + 4 * c4),
                             dq = load4<y_t>(y0 + 3 * D + 4 * c4);
                v[i].x = ((a.x + cq.x) + bq.x) + dq.x; v[i].y = ((a.y + cq.y) + bq.y) + dq.y;
                v[i].z = ((a.z + cq.z) + bq.z) + dq.z; v[i].w = ((a.w + cq.w) + bq.w) + dq.w;
                if constexpr (!std::is_same<y_t, float>::value) {
                    const float4 u = load4<__nv_bfloat16>(xc + tok * D + 4 * c4);
                    const float4 ds = __ldg(reinterpret_cast<const float4*>(Dsum) + c4);
                    v[i].x = fmaf(u.x, ds.x, v[i].x); v[i].y = fmaf(u.y, ds.y, v[i].y);
                    v[i].z = fmaf(u.z, ds.z, v[i].z); v[i].w = fmaf(u.w, ds.w, v[i].w);
                }
                sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
                if (ymerged) reinterpret_cast<float4*>(ymerged + tok * D)[c4] = v[i];
            } else {
                v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        const float mean = sum / (float)D;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            if (lane + 32 * i < C4) {
                const float a = v[i].x - mean, bq = v[i].y - mean, cq = v[i].z - mean, dq = v[i].w - mean;
                sq += (a * a + bq * bq) + (cq * cq + dq * dq);
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
        const float rstd = rsqrtf(sq / (float)D + eps);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int c4 = lane + 32 * i;
            if (c4 < C4) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
                const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + c4);
                const float4 zz = load4<z_t>(z + tok * z_pix + 4 * c4);
                float4 o;
                o.x = fmaf((v[i].x - mean) * rstd, g.x, bt.x) * silu_f(zz.x);
                o.y = fmaf((v[i].y - mean) * rstd, g.y, bt.y) * silu_f(zz.y);
                o.z = fmaf((v[i].z - mean) * rstd, g.z, bt.z) * silu_f(zz.z);
                o.w = fmaf((v[i].w - mean) * rstd, g.w, bt.w) * silu_f(zz.w);
                store4<out_t>(out + tok * D + 4 * c4, o);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The autocast layout of the same operation: bf16 direction slices holding the state terms, bf16 xc / z / out.
// Eight channels (one 128-bit load) per lane and vector, G lanes per token so that narrow rows keep the lanes busy
// (D = 96: 12 vectors -> 16-lane groups, two tokens per warp).  y = ((s0 + s2) + s1) + s3 + u * sum_k D_k in fp32.
struct F8 { float v[8]; };
__device__ __forceinline__ F8 ld_bf16x8(const __nv_bfloat16* p) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    F8 o;
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
        o.v[2 * i] = f.x; o.v[2 * i + 1] = f.y;
    }
    return o;
}
__device__ __forceinline__ void st_bf16x8(__nv_bfloat16* p, const F8& a) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(a.v[2 * i], a.v[2 * i + 1]);
        w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ F8 ld_f32x8(const float* p) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    F8 o;
    o.v[0] = a.x; o.v[1] = a.y; o.v[2] = a.z; o.v[3] = a.w; o.v[4] = b.x; o.v[5] = b.y; o.v[6] = b.z; o.v[7] = b.w;
    return o;
}

template <int V, int G>
__global__ void __launch_bounds__(256, 3)
outnorm_gate_bf16x8_kernel(const __nv_bfloat16* __restrict__ ydir, const __nv_bfloat16* __restrict__ z,
                           const float* __restrict__ gamma, const float* __restrict__ beta,
                           __nv_bfloat16* __restrict__ out, float* __restrict__ ymerged,
                           const __nv_bfloat16* __restrict__ xc, const float* __restrict__ Dsum,
                           int64_t tokens, int D, int64_t z_pix, float eps) {
    constexpr int TPW = 32 / G;
    constexpr bool HOIST = false;                    // per-channel constants stay in L1: registers buy occupancy here
    const int lane = threadIdx.x & 31, gl = lane % G;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int C8 = D / 8;
    const int64_t ngroups = (tokens + TPW - 1) / TPW;
    F8 gm[HOIST ? V : 1], bt[HOIST ? V : 1], dsm[HOIST ? V : 1];
    if (HOIST) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int c8 = gl + G * i;
            if (c8 < C8) { gm[i] = ld_f32x8(gamma + 8 * c8); bt[i] = ld_f32x8(beta + 8 * c8); dsm[i] = ld_f32x8(Dsum + 8 * c8); }
        }
    }
    for (int64_t grp = warp; grp < ngroups; grp += nwarps) {
        const int64_t tok = grp * TPW + lane / G;
        const bool tvalid = tok < tokens;
        const __nv_bfloat16* y0 = ydir + tok * 4 * D;
        F8 v[V];
        uint4 zraw[V];                               // the gate travels with the slices: every load of a token in flight at once
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int c8 = gl + G * i;
            zraw[i] = (tvalid && c8 < C8) ? __ldg(reinterpret_cast<const uint4*>(z + tok * z_pix + 8 * c8)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int c8 = gl + G * i;
            if (tvalid && c8 < C8) {
                const F8 a = ld_bf16x8(y0 + 8 * c8), bq = ld_bf16x8(y0 + D + 8 * c8), cq = ld_bf16x8(y0 + 2 * D + 8 * c8),
                         dq = ld_bf16x8(y0 + 3 * D + 8 * c8), u = ld_bf16x8(xc + tok * D + 8 * c8);
                const F8 ds = HOIST ? dsm[i] : ld_f32x8(Dsum + 8 * c8);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    v[i].v[e] = fmaf(u.v[e], ds.v[e], ((a.v[e] + cq.v[e]) + bq.v[e]) + dq.v[e]);
                    sum += v[i].v[e];
                }
                if (ymerged) {
                    float4* m = reinterpret_cast<float4*>(ymerged + tok * D + 8 * c8);
                    m[0] = make_float4(v[i].v[0], v[i].v[1], v[i].v[2], v[i].v[3]);
                    m[1] = make_float4(v[i].v[4], v[i].v[5], v[i].v[6], v[i].v[7]);
                }
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[i].v[e] = 0.f;
            }
        }
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        const float mean = sum / (float)D;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            if (gl + G * i < C8) {
#pragma unroll
                for (int e = 0; e < 8; ++e) { v[i].v[e] -= mean; sq = fmaf(v[i].v[e], v[i].v[e], sq); }
            }
        }
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
        const float rstd = rsqrtf(sq / (float)D + eps);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int c8 = gl + G * i;
            if (tvalid && c8 < C8) {
                const F8 g = HOIST ? gm[i] : ld_f32x8(gamma + 8 * c8);
                const F8 b = HOIST ? bt[i] : ld_f32x8(beta + 8 * c8);
                F8 zz;
                {
                    const uint32_t w[4] = {zraw[i].x, zraw[i].y, zraw[i].z, zraw[i].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
                        zz.v[2 * e] = f.x; zz.v[2 * e + 1] = f.y;
                    }
                }
                F8 o;
#pragma unroll
                for (int e = 0; e < 8; ++e) o.v[e] = fmaf(v[i].v[e] * rstd, g.v[e], b.v[e]) * silu_f(zz.v[e]);
                st_bf16x8(out + tok * D + 8 * c8, o);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Plain LayerNorm over the channels of a channels-last token matrix (warp per token).  Used for ln_1 on the
// strided right half of the residual stream (MedMamba.py:351), the patch-embed norm (:75) and the patch-merging
// norm (:116): rows of 48..1536 channels, for which a block-per-row LayerNorm leaves most threads idle.
// G lanes per token (32, or 16 / 8 for narrow rows so that few lanes idle: ln_1 of stage 1 has 12 float4 per token).
template <int V, typename in_t, typename out_t, int G = 32>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const in_t* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                     out_t* __restrict__ out, int64_t tokens, int D, int64_t x_pix, float eps) {
    constexpr int TPW = 32 / G;                      // tokens per warp
    const int lane = threadIdx.x & 31, gl = lane % G;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int C4 = D / 4;
    const int64_t ngroups = (tokens + TPW - 1) / TPW;
    for (int64_t grp = warp; grp < ngroups; grp += nwarps) {
        const int64_t tok = grp * TPW + lane / G;
        const bool tvalid = tok < tokens;
        float4 v[V];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int c4 = gl + G * i;
            v[i] = (tvalid && c4 < C4) ? load4<in_t>(x + tok * x_pix + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        const float mean = sum / (float)D;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            if (gl + G * i < C4) {
                v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
                sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
            }
        }
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
        const float rstd = rsqrtf(sq / (float)D + eps);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int c4 = gl + G * i;
            if (tvalid && c4 < C4) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
                const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + c4);
                float4 o;
                o.x = fmaf(v[i].x * rstd, g.x, bt.x); o.y = fmaf(v[i].y * rstd, g.y, bt.y);
                o.z = fmaf(v[i].z * rstd, g.z, bt.z); o.w = fmaf(v[i].w * rstd, g.w, bt.w);
                store4<out_t>(out + tok * D + 4 * c4, o);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// out[tok, c] = x[tok, c] * scale[c] + shift[c] on a strided channels-last view -> dense: the eval-mode BatchNorm
// that opens the CNN branch (MedMamba.py:338), fused with the gather of the left half of the residual stream
// and the cast to the convolution's dtype.
template <typename in_t, typename out_t>
__global__ void __launch_bounds__(256)
affine_cast_kernel(const in_t* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                   out_t* __restrict__ out, int64_t tokens, int C, int64_t x_pix) {
    const int C4 = C / 4;
    const uint32_t total = (uint32_t)tokens * C4;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % (uint32_t)C4);
        const int64_t tok = idx / (uint32_t)C4;
        const float4 v = load4<in_t>(x + tok * x_pix + 4 * c4);
        const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + c4);
        const float4 sh = __ldg(reinterpret_cast<const float4*>(shift) + c4);
        store4<out_t>(out + tok * C + 4 * c4,
                      make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w)));
    }
}

// ------------------------------------------------------------------------------------------------
// out[..., 2j] = left[..., j] + inp[..., 2j];  out[..., 2j+1] = ssm[..., j] + inp[..., 2j+1]
// (torch.cat + channel_shuffle(groups=2) + residual).  A thread produces 8 output channels.
template <typename TB, typename T>   // TB: branch dtype (left, ssm); T: residual stream dtype (inp, out)
__global__ void __launch_bounds__(256)
shuffle_cat_residual_kernel(const TB* __restrict__ left, const TB* __restrict__ ssm, const T* __restrict__ inp,
                            T* __restrict__ out, int64_t tokens, int c, int64_t left_pix, int64_t ssm_pix,
                            int64_t inp_pix) {
    const int C4 = c / 4;
    const uint32_t total = (uint32_t)tokens * C4;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % (uint32_t)C4);
        const int64_t tok = idx / (uint32_t)C4;
        const float4 l = load4<TB>(left + tok * left_pix + 4 * c4);
        const float4 s = load4<TB>(ssm + tok * ssm_pix + 4 * c4);
        const float4 i0 = load4<T>(inp + tok * inp_pix + 8 * c4);
        const float4 i1 = load4<T>(inp + tok * inp_pix + 8 * c4 + 4);
        float4 o0, o1;
        o0.x = l.x + i0.x; o0.y = s.x + i0.y; o0.z = l.y + i0.z; o0.w = s.y + i0.w;
        o1.x = l.z + i1.x; o1.y = s.z + i1.y; o1.z = l.w + i1.z; o1.w = s.w + i1.w;
        store4<T>(out + tok * 2 * c + 8 * c4, o0);
        store4<T>(out + tok * 2 * c + 8 * c4 + 4, o1);
    }
}

constexpr int kDwconvRowsDefault = 1;     // MMB_DWCONV_ROWS: 1 = two output rows per thread

static int grid_for(int64_t work_items, int threads) {
    int64_t blocks = (work_items + threads - 1) / threads;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <typename T> static bool aligned_for4(const void* p) { return reinterpret_cast<uintptr_t>(p) % vec4_align<T>() == 0; }

}  // namespace mmb

extern "C" int mmb_dwconv3x3_silu_fwd(const void* x, const float* weight, const float* bias, void* out,
                                      int batch, int H, int W, int D, int64_t x_pixel_stride, int64_t x_batch_stride,
                                      int in_dtype, int out_dtype, void* stream) {
    using namespace mmb;
    if (!x || !weight || !out) return MMB_ERR_INVALID_ARG;
    if (batch < 0 || H <= 0 || W <= 0 || D <= 0) return MMB_ERR_INVALID_ARG;
    if (D % 4 != 0 || D > 1280 || x_pixel_stride % 4 != 0 || x_batch_stride % 4 != 0) return MMB_ERR_UNSUPPORTED;
    if (bias && reinterpret_cast<uintptr_t>(bias) % 16 != 0) return MMB_ERR_UNSUPPORTED;
    if (batch == 0) return MMB_OK;
    constexpr int WS = 4;
    const int64_t items = (int64_t)batch * H * ((W + WS - 1) / WS) * (D / 4);
    if (items >= (1LL << 31)) return MMB_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = grid_for(items, 256);
#define MMB_DW(IN, OUT)                                                                                          \
    do {                                                                                                         \
        if (!aligned_for4<IN>(x) || !aligned_for4<OUT>(out)) return MMB_ERR_UNSUPPORTED;                         \
        dwconv3x3_silu_kernel<WS, IN, OUT><<<grid, 256, (size_t)D * 36, st>>>(reinterpret_cast<const IN*>(x), weight, bias,   \
            reinterpret_cast<OUT*>(out), batch, H, W, D, x_pixel_stride, x_batch_stride, (int64_t)D);                        \
        return launch_status();                                                                                  \
    } while (0)
    if (in_dtype == MMB_BF16 && out_dtype == MMB_BF16 && D % 8 == 0 && x_pixel_stride % 8 == 0 && x_batch_stride % 8 == 0 &&
        reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0 &&
        (!bias || reinterpret_cast<uintptr_t>(bias) % 16 == 0)) {
        const char* rows_env = getenv("MMB_DWCONV_ROWS");
        const uint64_t per_img = (uint64_t)((H + 1) / 2) * ((W + WS - 1) / WS) * (D / 4);
        if ((rows_env ? atoi(rows_env) != 0 : kDwconvRowsDefault) && batch <= 65535 && per_img * (D / 4) < (1ull << 31) &&
            (int64_t)(W - 1) * x_pixel_stride < (1LL << 31)) {
            // multiply-high reciprocals: exact for n * d < 2^32 (n < per_img)
            const uint32_t c4 = (uint32_t)(D / 4), strips = (uint32_t)((W + WS - 1) / WS);
            const uint32_t m_c4 = (uint32_t)(((1ull << 32) + c4 - 1) / c4), m_st = (uint32_t)(((1ull << 32) + strips - 1) / strips);
            int gx = (int)((per_img + 255) / 256);
            const int capx = (num_sms() * 16 + batch - 1) / batch;
            if (gx > capx) gx = capx < 1 ? 1 : capx;
            dim3 grid(gx, batch);
            dwconv3x3_silu_bf16_rows_kernel<<<grid, 256, (size_t)D * 36, st>>>(
                reinterpret_cast<const __nv_bfloat16*>(x), weight, bias, reinterpret_cast<__nv_bfloat16*>(out), H, W, D,
                x_pixel_stride, x_batch_stride, m_c4, m_st);
            return launch_status();
        }
        const int64_t items8 = (int64_t)batch * H * ((W + WS - 1) / WS) * (D / 8);
        dwconv3x3_silu_bf16x8_kernel<WS><<<grid_for(items8, 256), 256, (size_t)D * 36, st>>>(
            reinterpret_cast<const __nv_bfloat16*>(x), weight, bias, reinterpret_cast<__nv_bfloat16*>(out), batch, H, W, D,
            x_pixel_stride, x_batch_stride);
        return launch_status();
    }
    if (in_dtype == MMB_F32 && out_dtype == MMB_F32) MMB_DW(float, float);
    if (in_dtype == MMB_BF16 && out_dtype == MMB_F32) MMB_DW(__nv_bfloat16, float);
    if (in_dtype == MMB_F16 && out_dtype == MMB_F32) MMB_DW(__half, float);
    if (in_dtype == MMB_BF16 && out_dtype == MMB_BF16) MMB_DW(__nv_bfloat16, __nv_bfloat16);
    if (in_dtype == MMB_F32 && out_dtype == MMB_BF16) MMB_DW(float, __nv_bfloat16);
#undef MMB_DW
    return MMB_ERR_UNSUPPORTED;
}

extern "C" int mmb_dwconv3x3_bwd_dx(const float* ds, const float* weight, void* dx, int batch, int H, int W, int D,
                                    int64_t dx_pixel_stride, int out_dtype, void* stream) {
    using namespace mmb;
    if (!ds || !weight || !dx) return MMB_ERR_INVALID_ARG;
    if (batch < 0 || H <= 0 || W <= 0 || D <= 0) return MMB_ERR_INVALID_ARG;
    if (D % 4 != 0 || D > 1280 || !aligned_for4<float>(ds)) return MMB_ERR_UNSUPPORTED;
    if (dx_pixel_stride < D || dx_pixel_stride % 4 != 0) return MMB_ERR_UNSUPPORTED;
    if (batch == 0) return MMB_OK;
    constexpr int WS = 4;
    const int64_t items = (int64_t)batch * H * ((W + WS - 1) / WS) * (D / 4);
    if (items >= (1LL << 31)) return MMB_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = grid_for(items, 256);
    if (out_dtype == MMB_F32) {
        if (!aligned_for4<float>(dx)) return MMB_ERR_UNSUPPORTED;
        dwconv3x3_silu_kernel<WS, float, float, false, true><<<grid, 256, (size_t)D * 36, st>>>(
            ds, weight, nullptr, reinterpret_cast<float*>(dx), batch, H, W, D, (int64_t)D, (int64_t)H * W * D, dx_pixel_stride);
        return launch_status();
    }
    if (out_dtype == MMB_BF16) {
        if (!aligned_for4<__nv_bfloat16>(dx)) return MMB_ERR_UNSUPPORTED;
        dwconv3x3_silu_kernel<WS, float, __nv_bfloat16, false, true><<<grid, 256, (size_t)D * 36, st>>>(
            ds, weight, nullptr, reinterpret_cast<__nv_bfloat16*>(dx), batch, H, W, D, (int64_t)D, (int64_t)H * W * D,
            dx_pixel_stride);
        return launch_status();
    }
    return MMB_ERR_UNSUPPORTED;
}

extern "C" int mmb_outnorm_gate_fwd(const void* ydir, const void* z, const float* gamma, const float* beta,
                                    void* out, float* ymerged, const void* xc, const float* Dsum, int64_t tokens, int D,
                                    int64_t z_pixel_stride, float eps, int ydir_dtype, int z_dtype, int out_dtype,
                                    void* stream) {
    using namespace mmb;
    if (!ydir || !z || !gamma || !beta || !out) return MMB_ERR_INVALID_ARG;
    if (tokens < 0 || D <= 0) return MMB_ERR_INVALID_ARG;
    if (D % 4 != 0 || D > 1024 || z_pixel_stride % 4 != 0 || z_dtype != out_dtype) return MMB_ERR_UNSUPPORTED;
    if (ydir_dtype != MMB_F32 && ydir_dtype != MMB_BF16) return MMB_ERR_UNSUPPORTED;
    if (ydir_dtype == MMB_BF16 && (!xc || !Dsum)) return MMB_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(ydir) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) |
         reinterpret_cast<uintptr_t>(ymerged) | reinterpret_cast<uintptr_t>(xc) | reinterpret_cast<uintptr_t>(Dsum)) % 16 != 0)
        return MMB_ERR_UNSUPPORTED;
    if (tokens == 0) return MMB_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (ydir_dtype == MMB_BF16 && z_dtype == MMB_BF16 && D % 8 == 0 && z_pixel_stride % 8 == 0 &&
        (reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(out)) % 16 == 0) {
        auto yb = reinterpret_cast<const __nv_bfloat16*>(ydir);
        auto zb = reinterpret_cast<const __nv_bfloat16*>(z);
        auto xb = reinterpret_cast<const __nv_bfloat16*>(xc);
        auto ob = reinterpret_cast<__nv_bfloat16*>(out);
        const int C8 = D / 8;
#define MMB_X8(V, G)                                                                                             \
        do {                                                                                                     \
            const int g8 = grid_for(tokens * G, 256);                                                            \
            outnorm_gate_bf16x8_kernel<V, G><<<g8, 256, 0, st>>>(yb, zb, gamma, beta, ob, ymerged, xb, Dsum,     \
                                                                tokens, D, z_pixel_stride, eps);                 \
            return launch_status();                                                                              \
        } while (0)
        if (C8 <= 8) MMB_X8(1, 8);
        if (C8 <= 16) MMB_X8(1, 16);
        if (C8 <= 32) MMB_X8(1, 32);
        if (C8 <= 64) MMB_X8(2, 32);
        if (C8 <= 96) MMB_X8(3, 32);
        MMB_X8(4, 32);
#undef MMB_X8
    }
    const int grid = grid_for(tokens * 32, 256);
#define MMB_ON(V, T, Y)                                                                                          \
    do {                                                                                                         \
        if (!aligned_for4<T>(z) || !aligned_for4<T>(out)) return MMB_ERR_UNSUPPORTED;                            \
        outnorm_gate_kernel<V, T, T, Y><<<grid, 256, 0, st>>>(reinterpret_cast<const Y*>(ydir),                  \
            reinterpret_cast<const T*>(z), gamma, beta, reinterpret_cast<T*>(out), ymerged,                     \
            reinterpret_cast<const __nv_bfloat16*>(xc), Dsum, tokens, D, z_pixel_stride, eps);                   \
        return launch_status();                                                                                  \
    } while (0)
#define MMB_ON_V(T, Y)                                                                                           \
    do {                                                                                                         \
        if (D <= 128) MMB_ON(1, T, Y);                                                                           \
        if (D <= 256) MMB_ON(2, T, Y);                                                                           \
        if (D <= 512) MMB_ON(4, T, Y);                                                                           \
        MMB_ON(8, T, Y);                                                                                         \
    } while (0)
    if (ydir_dtype == MMB_F32) {
        if (z_dtype == MMB_F32) MMB_ON_V(float, float);
        if (z_dtype == MMB_BF16) MMB_ON_V(__nv_bfloat16, float);
        if (z_dtype == MMB_F16) MMB_ON_V(__half, float);
    } else {
        if (z_dtype == MMB_F32) MMB_ON_V(float, __nv_bfloat16);
        if (z_dtype == MMB_BF16) MMB_ON_V(__nv_bfloat16, __nv_bfloat16);
    }
#undef MMB_ON_V
#undef MMB_ON
    return MMB_ERR_UNSUPPORTED;
}

extern "C" int mmb_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* out, int64_t tokens, int D,
                                 int64_t x_pixel_stride, float eps, int in_dtype, int out_dtype, void* stream) {
    using namespace mmb;
    if (!x || !gamma || !beta || !out) return MMB_ERR_INVALID_ARG;
    if (tokens < 0 || D <= 0) return MMB_ERR_INVALID_ARG;
    if (D % 4 != 0 || D > 2048 || x_pixel_stride % 4 != 0) return MMB_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) % 16 != 0) return MMB_ERR_UNSUPPORTED;
    if (tokens == 0) return MMB_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    #define MMB_LN(V, TI, TO, G)                                                                                     \
    do {                                                                                                         \
        if (!aligned_for4<TI>(x) || !aligned_for4<TO>(out)) return MMB_ERR_UNSUPPORTED;                          \
        layernorm_fwd_kernel<V, TI, TO, G><<<grid_for(tokens * G, 256), 256, 0, st>>>(                           \
            reinterpret_cast<const TI*>(x), gamma, beta, reinterpret_cast<TO*>(out), tokens, D, x_pixel_stride,  \
            eps);                                                                                                \
        return launch_status();                                                                                  \
    } while (0)
#define MMB_LN_V(TI, TO)                                                                                         \
    do {                                                                                                         \
        if (D <= 32) MMB_LN(1, TI, TO, 8);                                                                       \
        if (D <= 64) MMB_LN(1, TI, TO, 16);                                                                      \
        if (D <= 128) MMB_LN(1, TI, TO, 32);                                                                         \
        if (D <= 256) MMB_LN(2, TI, TO, 32);                                                                         \
        if (D <= 512) MMB_LN(4, TI, TO, 32);                                                                         \
        if (D <= 1024) MMB_LN(8, TI, TO, 32);                                                                        \
        MMB_LN(16, TI, TO, 32);                                                                                    \
    } while (0)
    if (in_dtype == MMB_F32 && out_dtype == MMB_F32) MMB_LN_V(float, float);
    if (in_dtype == MMB_F32 && out_dtype == MMB_BF16) MMB_LN_V(float, __nv_bfloat16);
    if (in_dtype == MMB_BF16 && out_dtype == MMB_BF16) MMB_LN_V(__nv_bfloat16, __nv_bfloat16);
    if (in_dtype == MMB_BF16 && out_dtype == MMB_F32) MMB_LN_V(__nv_bfloat16, float);
#undef MMB_LN_V
#undef MMB_LN
    return MMB_ERR_UNSUPPORTED;
}

extern "C" int mmb_affine_cast_fwd(const void* x, const float* scale, const float* shift, void* out, int64_t tokens, int C,
                                   int64_t x_pixel_stride, int in_dtype, int out_dtype, void* stream) {
    using namespace mmb;
    if (!x || !scale || !shift || !out) return MMB_ERR_INVALID_ARG;
    if (tokens < 0 || C <= 0) return MMB_ERR_INVALID_ARG;
    if (C % 4 != 0 || x_pixel_stride % 4 != 0) return MMB_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(shift)) % 16 != 0) return MMB_ERR_UNSUPPORTED;
    if (tokens == 0) return MMB_OK;
    if (tokens * (C / 4) >= (1LL << 31)) return MMB_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = grid_for(tokens * (C / 4), 256);
#define MMB_AC(TI, TO)                                                                                           \
    do {                                                                                                         \
        if (!aligned_for4<TI>(x) || !aligned_for4<TO>(out)) return MMB_ERR_UNSUPPORTED;                          \
        affine_cast_kernel<TI, TO><<<grid, 256, 0, st>>>(reinterpret_cast<const TI*>(x), scale, shift,           \
                                                        reinterpret_cast<TO*>(out), tokens, C, x_pixel_stride);  \
        return launch_status();                                                                                  \
    } while (0)
    if (in_dtype == MMB_F32 && out_dtype == MMB_F32) MMB_AC(float, float);
    if (in_dtype == MMB_F32 && out_dtype == MMB_BF16) MMB_AC(float, __nv_bfloat16);
    if (in_dtype == MMB_BF16 && out_dtype == MMB_BF16) MMB_AC(__nv_bfloat16, __nv_bfloat16);
    if (in_dtype == MMB_BF16 && out_dtype == MMB_F32) MMB_AC(__nv_bfloat16, float);
#undef MMB_AC
    return MMB_ERR_UNSUPPORTED;
}

extern "C" int mmb_shuffle_cat_residual_fwd(const void* left, const void* ssm, const void* inp, void* out,
                                            int64_t tokens, int c, int64_t left_pixel_stride,
                                            int64_t ssm_pixel_stride, int64_t inp_pixel_stride, int branch_dtype,
                                            int res_dtype, void* stream) {
    using namespace mmb;
    if (!left || !ssm || !inp || !out) return MMB_ERR_INVALID_ARG;
    if (tokens < 0 || c <= 0) return MMB_ERR_INVALID_ARG;
    if (c % 4 != 0 || left_pixel_stride % 4 != 0 || ssm_pixel_stride % 4 != 0 || inp_pixel_stride % 4 != 0)
        return MMB_ERR_UNSUPPORTED;
    if (tokens == 0) return MMB_OK;
    if (tokens * (c / 4) >= (1LL << 31)) return MMB_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = grid_for(tokens * (c / 4), 256);
#define MMB_SH(TB, T)                                                                                            \
    do {                                                                                                         \
        if (!aligned_for4<TB>(left) || !aligned_for4<TB>(ssm) || !aligned_for4<T>(inp) || !aligned_for4<T>(out)) \
            return MMB_ERR_UNSUPPORTED;                                                                          \
        shuffle_cat_residual_kernel<TB, T><<<grid, 256, 0, st>>>(reinterpret_cast<const TB*>(left),              \
            reinterpret_cast<const TB*>(ssm), reinterpret_cast<const T*>(inp), reinterpret_cast<T*>(out), tokens,\
            c, left_pixel_stride, ssm_pixel_stride, inp_pixel_stride);                                           \
        return launch_status();                                                                                  \
    } while (0)
    if (branch_dtype == MMB_F32 && res_dtype == MMB_F32) MMB_SH(float, float);
    if (branch_dtype == MMB_BF16 && res_dtype == MMB_BF16) MMB_SH(__nv_bfloat16, __nv_bfloat16);
    if (branch_dtype == MMB_F16 && res_dtype == MMB_F16) MMB_SH(__half, __half);
    if (branch_dtype == MMB_BF16 && res_dtype == MMB_F32) MMB_SH(__nv_bfloat16, float);
    if (branch_dtype == MMB_F16 && res_dtype == MMB_F32) MMB_SH(__half, float);
#undef MMB_SH
    return MMB_ERR_UNSUPPORTED;
}
