// Backward of the bandwidth kernels of csrc/pointwise.cu (training path of the fused SS2D block).
// Parameter gradients are reduced deterministically: per-CTA partials in HBM, summed by the host.
#include "common.cuh"

namespace mmb {

constexpr int kPartBlocks = 592;     // CTAs of the LayerNorm-type reducing kernels = rows of their partial buffers (4 per SM)
constexpr int kConvPartBlocks = 296; // the same for dwconv3x3_silu_bwd_ds (128 registers by __launch_bounds__(256, 2) -- left to
                                     // itself ptxas took 157 and ONE CTA fitted an SM, ncu r2s3 -- 40 KB of partials per
                                     // CTA: 2 per SM; 592 CTAs ran 2.4x slower)

// Per-CTA partial of two per-channel sums (dgamma / dbeta): every lane holds the sums of channel slot (gl + G i) over
// the tokens it saw.  Warps are added in fixed order through shared memory, then the 32 / G token sub-groups of a
// warp with xor-shuffles: bit-reproducible.  part: (gridDim.x, 2, D).
template <int V, int G>
__device__ __forceinline__ void store_gb_partial(const float4 (&dg)[V], const float4 (&db)[V], float* __restrict__ part, int D) {
    __shared__ float4 sred[8][2][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, gl = lane % G;
    const int C4 = D / 4;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        sred[wib][0][lane] = dg[i]; sred[wib][1][lane] = db[i];
        __syncthreads();
        if (wib < 2) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int w = 0; w < 8; ++w) {
                const float4 q4 = sred[w][wib][lane];
                a.x += q4.x; a.y += q4.y; a.z += q4.z; a.w += q4.w;
            }
#pragma unroll
            for (int off = G; off < 32; off <<= 1) {
                a.x += __shfl_xor_sync(0xffffffffu, a.x, off); a.y += __shfl_xor_sync(0xffffffffu, a.y, off);
                a.z += __shfl_xor_sync(0xffffffffu, a.z, off); a.w += __shfl_xor_sync(0xffffffffu, a.w, off);
            }
            const int c4 = gl + G * i;
            if (lane < G && c4 < C4) reinterpret_cast<float4*>(part + ((int64_t)blockIdx.x * 2 + wib) * D)[c4] = a;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// out = LayerNorm(y) * silu(z):  given dout -> dy (fp32), dz, and per-CTA partials of dgamma / dbeta.
// G lanes per token (all 32, or 16 / 8 for narrow rows), every global load of a token issued before the first
// reduction: the first version (one warp per token, z / dout loaded after the statistics, 2 CTAs per SM) ran at a
// sixth of the HBM roofline at batch 128.
template <int V, typename z_t, int G>
__global__ void __launch_bounds__(256)
outnorm_gate_bwd_kernel(const z_t* __restrict__ dout, const float* __restrict__ ymerged, const z_t* __restrict__ z,
                        const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ dy,
                        z_t* __restrict__ dz, float* __restrict__ part, int64_t tokens, int D, int64_t z_pix,
                        int64_t dz_pix, float eps) {
    constexpr int TPW = 32 / G;
    constexpr bool EARLY = V <= 4;      // wide rows (D > 512): z / dout loaded after the statistics, or 8 more float4 spill
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, gl = lane % G;
    const int64_t warp = (int64_t)blockIdx.x * 8 + wib, nwarps = (int64_t)gridDim.x * 8;
    const int C4 = D / 4;
    const int64_t ngroups = (tokens + TPW - 1) / TPW;
    float4 g[V], bt[V], dg[V], db[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int c4 = gl + G * i;
        g[i] = c4 < C4 ? __ldg(reinterpret_cast<const float4*>(gamma) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        bt[i] = c4 < C4 ? __ldg(reinterpret_cast<const float4*>(beta) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        dg[i] = make_float4(0.f, 0.f, 0.f, 0.f); db[i] = dg[i];
    }
    for (int64_t grp = warp; grp < ngroups; grp += nwarps) {
        const int64_t tok = grp * TPW + lane / G;
        const bool tvalid = tok < tokens;
        float4 v[V], zz[EARLY ? V : 1], go[EARLY ? V : 1];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int c4 = gl + G * i;
            const bool ok = tvalid && c4 < C4;
            v[i] = ok ? __ldg(reinterpret_cast<const float4*>(ymerged + tok * D) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (EARLY) {
                zz[i] = ok ? load4<z_t>(z + tok * z_pix + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
                go[i] = ok ? load4<z_t>(dout + tok * D + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
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
        float4 t[V];
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int c4 = gl + G * i;
            t[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (tvalid && c4 < C4) {
                const float xh[4] = {v[i].x * rstd, v[i].y * rstd, v[i].z * rstd, v[i].w * rstd};
                const float gg[4] = {g[i].x, g[i].y, g[i].z, g[i].w}, bb[4] = {bt[i].x, bt[i].y, bt[i].z, bt[i].w};
                const float4 z4 = EARLY ? zz[EARLY ? i : 0] : load4<z_t>(z + tok * z_pix + 4 * c4);
                const float4 g4 = EARLY ? go[EARLY ? i : 0] : load4<z_t>(dout + tok * D + 4 * c4);
                const float zq[4] = {z4.x, z4.y, z4.z, z4.w}, gq[4] = {g4.x, g4.y, g4.z, g4.w};
                float dzz[4], tt[4], dgg[4], dbb[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float sgz = sigmoid_f(zq[e]);
                    const float n = fmaf(xh[e], gg[e], bb[e]);
                    const float dn = gq[e] * zq[e] * sgz;
                    dzz[e] = gq[e] * n * sgz * (1.f + zq[e] * (1.f - sgz));
                    dgg[e] = dn * xh[e]; dbb[e] = dn;
                    tt[e] = dn * gg[e];
                    m1 += tt[e]; m2 = fmaf(tt[e], xh[e], m2);
                }
                store4<z_t>(dz + tok * dz_pix + 4 * c4, make_float4(dzz[0], dzz[1], dzz[2], dzz[3]));
                dg[i].x += dgg[0]; dg[i].y += dgg[1]; dg[i].z += dgg[2]; dg[i].w += dgg[3];
                db[i].x += dbb[0]; db[i].y += dbb[1]; db[i].z += dbb[2]; db[i].w += dbb[3];
                t[i] = make_float4(tt[0], tt[1], tt[2], tt[3]);
                v[i] = make_float4(xh[0], xh[1], xh[2], xh[3]);
            }
        }
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) {
            m1 += __shfl_xor_sync(0xffffffffu, m1, off);
            m2 += __shfl_xor_sync(0xffffffffu, m2, off);
        }
        m1 /= (float)D; m2 /= (float)D;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int c4 = gl + G * i;
            if (tvalid && c4 < C4) {
                float4 o;
                o.x = rstd * (t[i].x - m1 - v[i].x * m2); o.y = rstd * (t[i].y - m1 - v[i].y * m2);
                o.z = rstd * (t[i].z - m1 - v[i].z * m2); o.w = rstd * (t[i].w - m1 - v[i].w * m2);
                reinterpret_cast<float4*>(dy + tok * D)[c4] = o;
            }
        }
    }
    store_gb_partial<V, G>(dg, db, part, D);
}

// ------------------------------------------------------------------------------------------------
// Backward of the plain LayerNorm (layernorm_fwd_kernel): ln_1, the patch-embed and patch-merging norms in
// training.  Statistics are recomputed from x (the row is in registers anyway), so the forward saves nothing:
//   xhat = (x - mean) rstd,  t = dy gamma,  dx = rstd (t - mean(t) - xhat mean(t xhat)),
//   dgamma += dy xhat,  dbeta += dy   (per-CTA partials, warps added in fixed order: bit-reproducible).
template <int V, typename x_t, typename dy_t, int G>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const x_t* __restrict__ x, const dy_t* __restrict__ dy, const float* __restrict__ gamma,
                     x_t* __restrict__ dx, float* __restrict__ part, int64_t tokens, int D, int64_t x_pix, float eps) {
    constexpr int TPW = 32 / G;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, gl = lane % G;
    const int64_t warp = (int64_t)blockIdx.x * 8 + wib, nwarps = (int64_t)gridDim.x * 8;
    const int C4 = D / 4;
    const int64_t ngroups = (tokens + TPW - 1) / TPW;
    float4 g[V], dg[V], db[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int c4 = gl + G * i;
        g[i] = c4 < C4 ? __ldg(reinterpret_cast<const float4*>(gamma) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        dg[i] = make_float4(0.f, 0.f, 0.f, 0.f); db[i] = dg[i];
    }
    for (int64_t grp = warp; grp < ngroups; grp += nwarps) {
        const int64_t tok = grp * TPW + lane / G;
        const bool tvalid = tok < tokens;
        float4 v[V], t[V];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int c4 = gl + G * i;
            const bool ok = tvalid && c4 < C4;
            v[i] = ok ? load4<x_t>(x + tok * x_pix + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            t[i] = ok ? load4<dy_t>(dy + tok * D + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
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
        float m1 = 0.f, m2 = 0.f;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            if (tvalid && gl + G * i < C4) {
                v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;          // xhat
                dg[i].x = fmaf(t[i].x, v[i].x, dg[i].x); dg[i].y = fmaf(t[i].y, v[i].y, dg[i].y);
                dg[i].z = fmaf(t[i].z, v[i].z, dg[i].z); dg[i].w = fmaf(t[i].w, v[i].w, dg[i].w);
                db[i].x += t[i].x; db[i].y += t[i].y; db[i].z += t[i].z; db[i].w += t[i].w;
                t[i].x *= g[i].x; t[i].y *= g[i].y; t[i].z *= g[i].z; t[i].w *= g[i].w;  // dy gamma
                m1 += (t[i].x + t[i].y) + (t[i].z + t[i].w);
                m2 += (t[i].x * v[i].x + t[i].y * v[i].y) + (t[i].z * v[i].z + t[i].w * v[i].w);
            }
        }
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) {
            m1 += __shfl_xor_sync(0xffffffffu, m1, off);
            m2 += __shfl_xor_sync(0xffffffffu, m2, off);
        }
        m1 /= (float)D; m2 /= (float)D;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const int c4 = gl + G * i;
            if (tvalid && c4 < C4) {
                float4 o;
                o.x = rstd * (t[i].x - m1 - v[i].x * m2); o.y = rstd * (t[i].y - m1 - v[i].y * m2);
                o.z = rstd * (t[i].z - m1 - v[i].z * m2); o.w = rstd * (t[i].w - m1 - v[i].w * m2);
                store4<x_t>(dx + tok * D + 4 * c4, o);
            }
        }
    }
    store_gb_partial<V, G>(dg, db, part, D);
}

// ------------------------------------------------------------------------------------------------
// xc = silu(s), s = dwconv(x) + bias.  K1: ds = dxc * silu'(s) (s recomputed), with per-CTA partials of
// dweight (D, 9) and dbias (D).  The input gradient is then the flipped-kernel convolution of ds
// (dwconv3x3_silu_kernel<..., ACT=false, FLIP=true> in pointwise.cu).
// The upstream gradient is the sum of up to three addends, added here instead of by a chain of elementwise kernels:
// dxc (tokens, D) fp32, the four direction slices of the core backward's dudir (tokens, 4, D) fp32, and the x_proj
// GEMM's input gradient dxe (tokens, D) in the activation dtype.
template <typename in_t>
__global__ void __launch_bounds__(256, 2)
dwconv3x3_silu_bwd_ds_kernel(const in_t* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ bias,
                             const float* __restrict__ dxc, const float* __restrict__ dudir, const in_t* __restrict__ dxe,
                             float* __restrict__ ds, float* __restrict__ part,
                             int B, int H, int W, int D, int64_t x_pix, int64_t x_batch) {
    extern __shared__ float sacc[];              // [ty][C4][40]
    const int C4 = D / 4;
    const int tx = threadIdx.x % C4, ty = threadIdx.x / C4, TY = blockDim.x / C4;
    const bool active = ty < TY;
    const int c = tx * 4;
    float wk[9][4], acc[10][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
#pragma unroll
        for (int tp = 0; tp < 9; ++tp) wk[tp][e] = __ldg(wgt + (int64_t)(c + e) * 9 + tp);
#pragma unroll
        for (int tp = 0; tp < 10; ++tp) acc[tp][e] = 0.f;
    }
    float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) bs = __ldg(reinterpret_cast<const float4*>(bias + c));
    const int pixels = B * H * W;                 // < 2^31 (checked by the entry point): 32-bit index arithmetic
    if (active) {
        for (int pix32 = blockIdx.x * TY + ty; pix32 < pixels; pix32 += gridDim.x * TY) {
            const int64_t pix = pix32;
            const int row = pix32 / W;
            const int w = pix32 - row * W;
            const int b = row / H;
            const int h = row - b * H;
            const in_t* xb = x + (int64_t)b * x_batch + c;
            float s[4] = {bs.x, bs.y, bs.z, bs.w};
            float xn[9][4];
            // unconditional loads from clamped coordinates (branches would serialise the nine latencies)
            float4 nb[9];
#pragma unroll
            for (int tp = 0; tp < 9; ++tp) {
                const int hy = min(max(h + tp / 3 - 1, 0), H - 1), wx = min(max(w + tp % 3 - 1, 0), W - 1);
                nb[tp] = load4<in_t>(xb + ((int64_t)hy * W + wx) * x_pix);
            }
#pragma unroll
            for (int tp = 0; tp < 9; ++tp) {
                const int hy = h + tp / 3 - 1, wx = w + tp % 3 - 1;
                const bool ok = hy >= 0 && hy < H && wx >= 0 && wx < W;
                xn[tp][0] = ok ? nb[tp].x : 0.f; xn[tp][1] = ok ? nb[tp].y : 0.f;
                xn[tp][2] = ok ? nb[tp].z : 0.f; xn[tp][3] = ok ? nb[tp].w : 0.f;
#pragma unroll
                for (int e = 0; e < 4; ++e) s[e] = fmaf(wk[tp][e], xn[tp][e], s[e]);
            }
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            if (dxc) g = __ldg(reinterpret_cast<const float4*>(dxc + pix * D + c));
            if (dudir) {
                const float4* dd = reinterpret_cast<const float4*>(dudir + pix * 4 * D + c);
                const float4 a0 = __ldg(dd), a1 = __ldg(dd + C4), a2 = __ldg(dd + 2 * C4), a3 = __ldg(dd + 3 * C4);
                g.x += (a0.x + a1.x) + (a2.x + a3.x); g.y += (a0.y + a1.y) + (a2.y + a3.y);
                g.z += (a0.z + a1.z) + (a2.z + a3.z); g.w += (a0.w + a1.w) + (a2.w + a3.w);
            }
            if (dxe) {
                const float4 e4 = load4<in_t>(dxe + pix * D + c);
                g.x += e4.x; g.y += e4.y; g.z += e4.z; g.w += e4.w;
            }
            const float gg[4] = {g.x, g.y, g.z, g.w};
            float d[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float sg = sigmoid_f(s[e]);
                d[e] = gg[e] * sg * (1.f + s[e] * (1.f - sg));
                acc[9][e] += d[e];
#pragma unroll
                for (int tp = 0; tp < 9; ++tp) acc[tp][e] = fmaf(d[e], xn[tp][e], acc[tp][e]);
            }
            *reinterpret_cast<float4*>(ds + pix * D + c) = make_float4(d[0], d[1], d[2], d[3]);
        }
#pragma unroll
        for (int tp = 0; tp < 10; ++tp)
#pragma unroll
            for (int e = 0; e < 4; ++e) sacc[((ty * C4 + tx) * 10 + tp) * 4 + e] = acc[tp][e];
    }
    __syncthreads();
    // part[block][c][10]: rows of TY added in fixed order
    for (int idx = threadIdx.x; idx < C4 * 40; idx += blockDim.x) {
        const int cx = idx / 40, r = idx % 40, tp = r / 4, e = r % 4;
        float a = 0.f;
        for (int y = 0; y < TY; ++y) a += sacc[((y * C4 + cx) * 10 + tp) * 4 + e];
        part[((int64_t)blockIdx.x * D + cx * 4 + e) * 10 + tp] = a;
    }
}

// de-interleave: dleft[..., j] = dout[..., 2j], dssm[..., j] = dout[..., 2j + 1]
template <typename T, typename TB>
__global__ void __launch_bounds__(256)
shuffle_cat_residual_bwd_kernel(const T* __restrict__ dout, TB* __restrict__ dleft, TB* __restrict__ dssm,
                                int64_t tokens, int c) {
    const int C4 = c / 4;
    const int64_t total = tokens * C4;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % C4);
        const int64_t tok = idx / C4;
        const float4 a = load4<T>(dout + tok * 2 * c + 8 * c4), b = load4<T>(dout + tok * 2 * c + 8 * c4 + 4);
        store4<TB>(dleft + tok * c + 4 * c4, make_float4(a.x, a.z, b.x, b.z));
        store4<TB>(dssm + tok * c + 4 * c4, make_float4(a.y, a.w, b.y, b.w));
    }
}

template <typename T> static bool al4(const void* p) { return reinterpret_cast<uintptr_t>(p) % vec4_align<T>() == 0; }

}  // namespace mmb

extern "C" int mmb_partial_blocks(void) { return mmb::kPartBlocks; }
extern "C" int mmb_dwconv_partial_blocks(void) { return mmb::kConvPartBlocks; }

extern "C" int mmb_outnorm_gate_bwd(const void* dout, const float* ymerged, const void* z, const float* gamma,
                                    const float* beta, float* dy, void* dz, float* dgb_part, int64_t tokens, int D,
                                    int64_t z_pixel_stride, int64_t dz_pixel_stride, float eps, int z_dtype, void* stream) {
    using namespace mmb;
    if (!dout || !ymerged || !z || !gamma || !beta || !dy || !dz || !dgb_part) return MMB_ERR_INVALID_ARG;
    if (tokens < 0 || D <= 0) return MMB_ERR_INVALID_ARG;
    if (D % 4 != 0 || D > 1024 || z_pixel_stride % 4 != 0 || dz_pixel_stride % 4 != 0 || dz_pixel_stride < D) return MMB_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(ymerged) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) |
         reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dgb_part)) % 16 != 0) return MMB_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define MMB_OB(V, T, G)                                                                                          \
    do {                                                                                                         \
        if (!al4<T>(z) || !al4<T>(dout) || !al4<T>(dz)) return MMB_ERR_UNSUPPORTED;                              \
        outnorm_gate_bwd_kernel<V, T, G><<<kPartBlocks, 256, 0, st>>>(reinterpret_cast<const T*>(dout), ymerged, \
            reinterpret_cast<const T*>(z), gamma, beta, dy, reinterpret_cast<T*>(dz), dgb_part, tokens, D,       \
            z_pixel_stride, dz_pixel_stride, eps);                                                               \
        return launch_status();                                                                                  \
    } while (0)
#define MMB_OB_V(T)                                                                                              \
    do {                                                                                                         \
        if (D <= 32) MMB_OB(1, T, 8);                                                                            \
        if (D <= 64) MMB_OB(1, T, 16);                                                                           \
        if (D <= 128) MMB_OB(1, T, 32);                                                                          \
        if (D <= 256) MMB_OB(2, T, 32);                                                                          \
        if (D <= 512) MMB_OB(4, T, 32);                                                                          \
        MMB_OB(8, T, 32);                                                                                        \
    } while (0)
    if (z_dtype == MMB_F32) MMB_OB_V(float);
    if (z_dtype == MMB_BF16) MMB_OB_V(__nv_bfloat16);
#undef MMB_OB_V
#undef MMB_OB
    return MMB_ERR_UNSUPPORTED;
}

extern "C" int mmb_layernorm_bwd(const void* x, const void* dy, const float* gamma, void* dx, float* dgb_part,
                                 int64_t tokens, int D, int64_t x_pixel_stride, float eps, int x_dtype, int dy_dtype,
                                 void* stream) {
    using namespace mmb;
    if (!x || !dy || !gamma || !dx || !dgb_part) return MMB_ERR_INVALID_ARG;
    if (tokens < 0 || D <= 0) return MMB_ERR_INVALID_ARG;
    if (D % 4 != 0 || D > 512 || x_pixel_stride % 4 != 0) return MMB_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(dgb_part)) % 16 != 0) return MMB_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define MMB_LB(V, TX, TY, G)                                                                                     \
    do {                                                                                                         \
        if (!al4<TX>(x) || !al4<TX>(dx) || !al4<TY>(dy)) return MMB_ERR_UNSUPPORTED;                             \
        layernorm_bwd_kernel<V, TX, TY, G><<<kPartBlocks, 256, 0, st>>>(reinterpret_cast<const TX*>(x),          \
            reinterpret_cast<const TY*>(dy), gamma, reinterpret_cast<TX*>(dx), dgb_part, tokens, D,              \
            x_pixel_stride, eps);                                                                                \
        return launch_status();                                                                                  \
    } while (0)
#define MMB_LB_V(TX, TY)                                                                                         \
    do {                                                                                                         \
        if (D <= 32) MMB_LB(1, TX, TY, 8);                                                                       \
        if (D <= 64) MMB_LB(1, TX, TY, 16);                                                                      \
        if (D <= 128) MMB_LB(1, TX, TY, 32);                                                                     \
        if (D <= 256) MMB_LB(2, TX, TY, 32);                                                                     \
        MMB_LB(4, TX, TY, 32);                                                                                   \
    } while (0)
    if (x_dtype == MMB_F32 && dy_dtype == MMB_F32) MMB_LB_V(float, float);
    if (x_dtype == MMB_F32 && dy_dtype == MMB_BF16) MMB_LB_V(float, __nv_bfloat16);
    if (x_dtype == MMB_BF16 && dy_dtype == MMB_BF16) MMB_LB_V(__nv_bfloat16, __nv_bfloat16);
    if (x_dtype == MMB_BF16 && dy_dtype == MMB_F32) MMB_LB_V(__nv_bfloat16, float);
#undef MMB_LB_V
#undef MMB_LB
    return MMB_ERR_UNSUPPORTED;
}

extern "C" int mmb_dwconv3x3_silu_bwd_ds(const void* x, const float* weight, const float* bias, const float* dxc,
                                         const float* dudir, const void* dxc_extra, float* ds, float* dwb_part, int batch,
                                         int H, int W, int D, int64_t x_pixel_stride, int64_t x_batch_stride, int in_dtype,
                                         void* stream) {
    using namespace mmb;
    if (!x || !weight || (!dxc && !dudir && !dxc_extra) || !ds || !dwb_part) return MMB_ERR_INVALID_ARG;
    if (batch < 0 || H <= 0 || W <= 0 || D <= 0) return MMB_ERR_INVALID_ARG;
    if (D % 4 != 0 || D / 4 > 256 || x_pixel_stride % 4 != 0 || x_batch_stride % 4 != 0) return MMB_ERR_UNSUPPORTED;
    if ((int64_t)batch * H * W > 0x7ffffff0LL - (int64_t)kConvPartBlocks * 256) return MMB_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(dxc) | reinterpret_cast<uintptr_t>(ds) | reinterpret_cast<uintptr_t>(bias) |
         reinterpret_cast<uintptr_t>(dudir) | reinterpret_cast<uintptr_t>(dxc_extra)) % 16 != 0)
        return MMB_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int C4 = D / 4;
    const int TY = 256 / C4;
    const int threads = 256;
    const size_t smem = sizeof(float) * (size_t)TY * C4 * 40;
#define MMB_DB(T)                                                                                                \
    do {                                                                                                         \
        if (!al4<T>(x)) return MMB_ERR_UNSUPPORTED;                                                              \
        auto kern = dwconv3x3_silu_bwd_ds_kernel<T>;                                                             \
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);\
        kern<<<kConvPartBlocks, threads, smem, st>>>(reinterpret_cast<const T*>(x), weight, bias, dxc, dudir,    \
                                                 reinterpret_cast<const T*>(dxc_extra), ds, dwb_part,            \
                                                 batch, H, W, D, x_pixel_stride, x_batch_stride);                \
        return launch_status();                                                                                  \
    } while (0)
    if (in_dtype == MMB_F32) MMB_DB(float);
    if (in_dtype == MMB_BF16) MMB_DB(__nv_bfloat16);
#undef MMB_DB
    return MMB_ERR_UNSUPPORTED;
}

extern "C" int mmb_shuffle_cat_residual_bwd(const void* dout, void* dleft, void* dssm, int64_t tokens, int c,
                                            int res_dtype, int branch_dtype, void* stream) {
    using namespace mmb;
    if (!dout || !dleft || !dssm) return MMB_ERR_INVALID_ARG;
    if (tokens < 0 || c <= 0) return MMB_ERR_INVALID_ARG;
    if (c % 4 != 0) return MMB_ERR_UNSUPPORTED;
    if (tokens == 0) return MMB_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int64_t blocks = (tokens * (c / 4) + 255) / 256;
    if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
#define MMB_SB(T, TB)                                                                                            \
    do {                                                                                                         \
        if (!al4<T>(dout) || !al4<TB>(dleft) || !al4<TB>(dssm)) return MMB_ERR_UNSUPPORTED;                      \
        shuffle_cat_residual_bwd_kernel<T, TB><<<(int)blocks, 256, 0, st>>>(reinterpret_cast<const T*>(dout),    \
            reinterpret_cast<TB*>(dleft), reinterpret_cast<TB*>(dssm), tokens, c);                               \
        return launch_status();                                                                                  \
    } while (0)
    if (res_dtype == MMB_F32 && branch_dtype == MMB_F32) MMB_SB(float, float);
    if (res_dtype == MMB_F32 && branch_dtype == MMB_BF16) MMB_SB(float, __nv_bfloat16);
    if (res_dtype == MMB_BF16 && branch_dtype == MMB_BF16) MMB_SB(__nv_bfloat16, __nv_bfloat16);
#undef MMB_SB
    return MMB_ERR_UNSUPPORTED;
}
