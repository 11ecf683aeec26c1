// Fused SS2D core, forward: cross-scan + dt_proj + softplus + selective scan + cross-merge addressing.
// Replaces MedMamba.py:256-257 (cross-scan), :262,:266 (dt_proj einsum + copy), :273-279
// (selective_scan_fn) and :282-286 (flip / transpose back) for all four directions in one call.
//
// Layout (channels-last, everything indexed by the token position p = h*W + w, never by the
// per-direction sequence index):
//   xc   (B, H, W, D)        conv+SiLU output = u for every direction             [fp32 | bf16]
//   proj (B, H, W, 4, CP)    x_proj of the un-permuted tokens, per direction k:
//                            [ B_n (16) | C_n (16) | dt_r (RP, zero padded) ]      [fp32]
//   ydir (B, H, W, 4, D)     scan output of direction k written at the position it belongs to;
//                            fp32 xc: y_k = <C, h> + D_k u in fp32 (the out_norm kernel adds the four in the
//                            reference's order); bf16 xc: only the state term <C, h> in bf16 -- the out_norm kernel
//                            adds u * sum_k D_k in fp32, so the slices cost half the bytes without rounding the
//                            large u * D term.
// Direction k walks the positions in the order of SURVEY.md Appendix A (row-major, column-major
// and their reverses), so no flipped or transposed copy of anything is ever materialised.
//
// Work decomposition.  An ITEM is (batch b, direction k, tile of CT channels): one sequence of L steps, cut into
// blocks of <= 32 consecutive steps that stream through a 4-stage TMA ring.  One CTA runs one item -- or, for launches
// with fewer sequences than the machine has CTA slots (small batches, the 512x512 config), one SEGMENT of an item:
//   * L-parallel mode (passes 1 + 2): every sequence is cut into G segments of whole blocks.  Pass 1 runs segments
//     0..G-2 from a zero state and keeps only the segment summaries: the end state and the sum of delta, which gives
//     the segment's decay product exp(A * sum delta) without a second exp per step.  Pass 2 runs all G segments in
//     parallel, each starting from the prefix carried over its predecessors' summaries
//     (h <- h * exp(A * sum delta_j) + h_end_j, j = 0..g-1).  Exact up to fp32 rounding of the decay products.
// A persistent grid (one CTA per resident slot walking a list of items with one continuous ring, static round-robin
// or balanced shares with an exact state hand-off) was built and measured in round 2 and is NOT used: with the same
// inner loop it was 11-43 % slower than letting the hardware hand fresh CTAs to whichever SM frees a slot
// (profiles/README.md, round 2: CTAs that start together stay in lock step, so their per-item parameter loads and
// block boundaries hit L2 and the MUFU pipe in bursts, and a static share is as slow as its slowest CTA).
// Every warp holds one channel per S lanes with the 16/S states of the lane in registers.  Per step and state: one
// MUFU.EX2 and four FMA-pipe operations; the kernel is bound by the 16 exp/clk/SM MUFU rate, not by HBM.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "core_geom.cuh"
#include "tma.cuh"

namespace mmb {

constexpr int kCoreStages = 4;
constexpr bool kVoteSoftplus = true;     // one warp vote per group of 4 steps skips the lg2 half of softplus
constexpr bool kPackSoftplus = true;    // A/B on B200: packed pairs 0.9% faster over the four stage shapes
constexpr int kCoreStageBytes = 16 * 1024;
constexpr int kMaxSegs = 32;

struct CoreFwdParams {
    void* ydir;
    const float* Wdt;     // (4, D, R)
    const float* bias;    // (4, D)
    const float* A;       // (4*D, N)
    const float* Ds;      // (4*D)
    int B, H, W, L, D, N, R, CT, tiles;
    int T_row, NB_row;                 // row view: NB_row blocks of T_row consecutive positions
    int nw, T_col, NI_col, NO_col;     // column view: NO_col column groups x NI_col row blocks
    int cap;                           // steps a block can hold
    int sub;                           // blocks per ring stage: 1, or 4 in the training forward, whose 8-step blocks (the
                                       // checkpoint spacing the backward recomputes from) would otherwise cost a barrier
                                       // round trip and a TMA issue per 8 steps (blocks shorter than 32 steps: up to -16 %)
    float* hsave;                      // NULL, or (B, 4, NBmax, D, 16): state after every block (training)
    int NBmax;
    int segs, bps_row, bps_col;        // L-parallel: segments per sequence, blocks per segment in either view
    float* seg_h;                      // (B, 4, segs - 1, D, 16): end state of a segment started from zero
    float* seg_dsum;                   // (B, 4, segs - 1, D):     sum of delta over the segment
};

template <int S, int RP, typename xc_t, typename y_t, int PASS, int MB>
__global__ void __launch_bounds__(S == 1 ? 256 : 384, MB)
ss2d_core_fwd_kernel(const __grid_constant__ CUtensorMap tmx_row, const __grid_constant__ CUtensorMap tmx_col,
                     const __grid_constant__ CUtensorMap tmp_row, const __grid_constant__ CUtensorMap tmp_col,
                     const CoreFwdParams p) {
    constexpr int NS = kMaxState / S;
    constexpr int CP = 32 + RP;
    constexpr int OWN = 4 / S;          // delta evaluations per lane per group of 4 steps
    constexpr bool SUMMARY = PASS == 1; // segment summaries only: no C, no y
    constexpr bool YSPLIT = !std::is_same<y_t, float>::value;    // bf16 slices hold the state term only

    extern __shared__ __align__(128) uint8_t smem_raw[];
    constexpr int XE = (int)sizeof(xc_t);
    const int xsub = (p.cap * p.CT * XE + 127) & ~127, psub = (p.cap * CP * 4 + 127) & ~127;     // one block
    const int xpad = p.sub * xsub, ppad = p.sub * psub;                                           // one ring stage
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kCoreStages * (xpad + ppad));
    uint64_t* empty = full + kCoreStages;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nwarps = blockDim.x >> 5;
    // grid: x = tile + tiles * segment, y = direction, z = batch
    const int k = blockIdx.y, b = blockIdx.z;
    const int tile = PASS == 0 ? blockIdx.x : blockIdx.x % p.tiles;
    const int seg = PASS == 0 ? 0 : blockIdx.x / p.tiles;
    const int c0 = tile * p.CT;
    const bool colview = (k & 1) != 0, rev = k >= 2;
    const int NB = colview ? p.NO_col * p.NI_col : p.NB_row;        // blocks of the sequence
    const int NST = (NB + p.sub - 1) / p.sub;                        // ring stages of the sequence (sub blocks each)
    // stages [jb0, jb1) of the sequence, in time order: the whole sequence, or this CTA's segment (sub == 1 there)
    int jb0 = 0, jb1 = NST;
    if (PASS != 0) {
        const int bps = colview ? p.bps_col : p.bps_row;
        jb0 = seg * bps;
        jb1 = jb0 + bps < NB ? jb0 + bps : NB;
    }

    // Stage js (time order) -> TMA loads of its blocks into ring slot (js - jb0) % kCoreStages; issued by thread 0 only.
    auto issue = [&](int js) {
        const int s = (js - jb0) % kCoreStages;
        const int nblk = min(p.sub, NB - js * p.sub);
        uint8_t* xs = smem_raw + s * (xpad + ppad);
        uint8_t* ps = xs + xpad;
        mbar_expect_tx(&full[s], nblk * (colview ? p.nw * p.T_col : p.T_row) * (p.CT * XE + CP * 4));
        for (int sb = 0; sb < nblk; ++sb) {
            const int jb = js * p.sub + sb;
            const int blk = rev ? NB - 1 - jb : jb;
            if (!colview) {
                tma_load_3d(xs + sb * xsub, &tmx_row, &full[s], c0, blk * p.T_row, b);
                tma_load_4d(ps + sb * psub, &tmp_row, &full[s], 0, k, blk * p.T_row, b);
            } else {
                const int o = blk / p.NI_col, i = blk % p.NI_col;
                tma_load_4d(xs + sb * xsub, &tmx_col, &full[s], c0, o * p.nw, i * p.T_col, b);
                tma_load_5d(ps + sb * psub, &tmp_col, &full[s], 0, k, o * p.nw, i * p.T_col, b);
            }
        }
    };

    if (tid == 0) {
        prefetch_tmap(colview ? &tmx_col : &tmx_row);
        prefetch_tmap(colview ? &tmp_col : &tmp_row);
        for (int s = 0; s < kCoreStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], nwarps); }
        mbar_fence_init();
        for (int jb = jb0; jb < jb0 + kCoreStages - 1 && jb < jb1; ++jb) issue(jb);
    }
    __syncthreads();
    int jn = jb0 + kCoreStages - 1 < jb1 ? jb0 + kCoreStages - 1 : jb1;     // next stage to load (thread 0)

    const int t = warp * 32 + lane;
    const int cl = t / S, q = t % S;            // channel inside the tile, state split index
    const int c = c0 + cl;
    const bool cvalid = c < p.D;
    const int row = k * p.D + (cvalid ? c : 0);
    const int lane_base = lane & ~(S - 1);

    // Per-channel parameters.  Rows of A (N floats) and Wdt (R floats) are contiguous per channel: 128-bit loads
    // when the row is 16-byte aligned (one L2 sector pair per lane instead of one sector per scalar: at the 7x7 stage
    // these loads are a sixth of a CTA's lifetime in L2 traffic otherwise).
    float Ap[NS], h[NS], Wd[RP];
    if (p.N == kMaxState) {
        const float4* ar = reinterpret_cast<const float4*>(p.A + (int64_t)row * kMaxState) + q;
#pragma unroll
        for (int j4 = 0; j4 < NS / 4; ++j4) {
            const float4 v = cvalid ? __ldg(ar + j4 * S) : make_float4(0.f, 0.f, 0.f, 0.f);
            Ap[4 * j4 + 0] = v.x * kLog2e; Ap[4 * j4 + 1] = v.y * kLog2e; Ap[4 * j4 + 2] = v.z * kLog2e; Ap[4 * j4 + 3] = v.w * kLog2e;
        }
    } else {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            const int n = (j & 3) + 4 * q + 4 * S * (j >> 2);
            Ap[j] = (cvalid && n < p.N) ? p.A[(int64_t)row * p.N + n] * kLog2e : 0.f;
        }
    }
#pragma unroll
    for (int j = 0; j < NS; ++j) h[j] = 0.f;
    if (p.R == RP) {                                   // dt_rank a multiple of 4 (12, 24, ...): aligned rows
        const float4* wr = reinterpret_cast<const float4*>(p.Wdt + (int64_t)row * RP);
#pragma unroll
        for (int r4 = 0; r4 < RP / 4; ++r4) {
            const float4 v = cvalid ? __ldg(wr + r4) : make_float4(0.f, 0.f, 0.f, 0.f);
            Wd[4 * r4 + 0] = v.x; Wd[4 * r4 + 1] = v.y; Wd[4 * r4 + 2] = v.z; Wd[4 * r4 + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int r = 0; r < RP; ++r) Wd[r] = (cvalid && r < p.R) ? __ldg(p.Wdt + (int64_t)row * p.R + r) : 0.f;
    }
    const float bias = cvalid ? __ldg(p.bias + row) : 0.f;
    const float Dd = cvalid ? __ldg(p.Ds + row) : 0.f;
    y_t* yb = reinterpret_cast<y_t*>(p.ydir) + ((int64_t)b * p.L * 4 + k) * p.D + c;   // + pos * 4 * D
    const int64_t ystride = 4 * (int64_t)p.D;
    float dsum = 0.f;

    if (PASS == 2 && seg > 0 && cvalid) {
        // prefix carried over the summaries of segments 0 .. seg-1: h <- h * exp(A * sum delta) + h_end
        const int64_t sb = ((int64_t)(b * 4 + k) * (p.segs - 1)) * p.D + c;
        for (int sj = 0; sj < seg; ++sj) {
            const float ds = p.seg_dsum[sb + (int64_t)sj * p.D];
            const float* hs = p.seg_h + (sb + (int64_t)sj * p.D) * kMaxState + 4 * q;
#pragma unroll
            for (int j4 = 0; j4 < NS / 4; ++j4) {
                const float4 e = *reinterpret_cast<const float4*>(hs + 4 * S * j4);
                h[4 * j4 + 0] = fmaf(h[4 * j4 + 0], ex2_approx(Ap[4 * j4 + 0] * ds), e.x);
                h[4 * j4 + 1] = fmaf(h[4 * j4 + 1], ex2_approx(Ap[4 * j4 + 1] * ds), e.y);
                h[4 * j4 + 2] = fmaf(h[4 * j4 + 2], ex2_approx(Ap[4 * j4 + 2] * ds), e.z);
                h[4 * j4 + 3] = fmaf(h[4 * j4 + 3], ex2_approx(Ap[4 * j4 + 3] * ds), e.w);
            }
        }
    }

    for (int js = jb0; js < jb1; ++js) {
        const int jl = js - jb0;
        const int s = jl % kCoreStages, ph = (jl / kCoreStages) & 1;
        // Refill the ring (inline producer: one thread).  The slot of stage jn was last used by stage jn - kCoreStages and is
        // free once every warp released that one.  Thread 0 only TESTS for it here and blocks only when the stage this warp
        // is about to compute has not been requested yet: waiting would tie warp 0 to the slowest warp of the CTA at every
        // block, and a refill that is one block late still has kCoreStages - 2 blocks of lead (-1 % at stage 1).
        if (tid == 0) {
            while (jn < jb1 && jn - js < kCoreStages) {
                const int jprev = jn - kCoreStages;
                if (jprev >= jb0) {
                    uint64_t* eb = &empty[(jn - jb0) % kCoreStages];
                    const uint32_t par = ((jprev - jb0) / kCoreStages) & 1;
                    if (!mbar_test_wait(eb, par)) {
                        if (jn > js) break;
                        mbar_wait(eb, par);
                    }
                }
                issue(jn);
                ++jn;
            }
        }
        __syncwarp();
        mbar_wait(&full[s], ph);
      for (int sb = 0; sb < p.sub; ++sb) {
        const int jb = js * p.sub + sb;
        if (jb >= NB) break;
        const int blk = rev ? NB - 1 - jb : jb;
        // geometry: nrows x ncols positions, column-major in time; lane ti owns the slot / position
        // of the block's ti-th step in forward order (a block never has more than 32 steps)
        int nrows, ncols, nwbox, psh, pbase;
        if (!colview) {
            pbase = blk * p.T_row; nrows = min(p.T_row, p.L - pbase); ncols = 1; nwbox = 1; psh = 1;
        } else {
            const int o = blk / p.NI_col, i = blk % p.NI_col;
            const int w0 = o * p.nw, h0 = i * p.T_col;
            nrows = min(p.T_col, p.H - h0); ncols = min(p.nw, p.W - w0); nwbox = p.nw; psh = p.W;
            pbase = h0 * p.W + w0;
        }
        const int nsteps = nrows * ncols;
        int slot_l = 0, pos_l = 0;
        if (lane < nsteps) {
            const int ww = lane / nrows, hh = lane - ww * nrows;
            slot_l = hh * nwbox + ww;
            pos_l = pbase + hh * psh + ww;
        }
        const xc_t* xs = reinterpret_cast<const xc_t*>(smem_raw + s * (xpad + ppad) + sb * xsub) + cl;
        const float* ps = reinterpret_cast<const float*>(smem_raw + s * (xpad + ppad) + xpad + sb * psub);

        const bool single_col = nwbox == 1;     // slot / position are then affine in the step index: no shuffles
        // One group = four consecutive steps.  MODE 0: ragged last group of a block (per-step validity predicates);
        // MODE 3: full group, run-time direction, lane-table shuffles when the block spans several columns;
        // MODE 1 / 2: full group of a single-column block walked forwards / backwards: slots and positions are
        // base + i * constant, no predicates, no selects on the direction.
        auto group = [&](auto mode_tag, const int g0) {
            constexpr int MODE = decltype(mode_tag)::value;
            constexpr bool FULL = MODE != 0;
            int slot[4], pos[4];
            bool ok[4];
            if (MODE == 0 || MODE == 3) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int tl = g0 + i;
                    ok[i] = FULL || tl < nsteps;
                    const int ti = ok[i] ? (rev ? nsteps - 1 - tl : tl) : 0;
                    if (single_col) { slot[i] = ti; pos[i] = pbase + ti * psh; }
                    else { slot[i] = __shfl_sync(0xffffffffu, slot_l, ti); pos[i] = __shfl_sync(0xffffffffu, pos_l, ti); }
                }
            } else {
                const int t0 = MODE == 1 ? g0 : nsteps - 1 - g0;
                const int p0 = pbase + t0 * psh;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    ok[i] = true;
                    slot[i] = MODE == 1 ? t0 + i : t0 - i;
                    pos[i] = MODE == 1 ? p0 + i * psh : p0 - i * psh;
                }
            }
            float uu[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) uu[i] = ok[i] ? to_f<xc_t>(xs[slot[i] * p.CT]) : 0.f;
            // delta = softplus(Wdt . dt_r + bias): each lane evaluates OWN of the four steps
            float down[OWN];
            float raw[OWN];
            bool okd[OWN];
#pragma unroll
            for (int m = 0; m < OWN; ++m) {
                int sl;
                if (S == 1) { sl = slot[m]; okd[m] = ok[m]; }
                else {
                    const int tl = g0 + q + S * m;
                    okd[m] = FULL || tl < nsteps;
                    const int ti = okd[m] ? (rev ? nsteps - 1 - tl : tl) : 0;
                    sl = single_col ? ti : __shfl_sync(0xffffffffu, slot_l, ti);
                }
                const float4* dtp = reinterpret_cast<const float4*>(ps + sl * CP + 32);
                float acc0 = bias, acc1 = 0.f;
#pragma unroll
                for (int r4 = 0; r4 < RP / 4; ++r4) {
                    const float4 v = dtp[r4];
                    fma2(acc0, acc1, Wd[4 * r4 + 0], Wd[4 * r4 + 1], v.x, v.y, acc0, acc1);
                    fma2(acc0, acc1, Wd[4 * r4 + 2], Wd[4 * r4 + 3], v.z, v.w, acc0, acc1);
                }
                raw[m] = acc0 + acc1;
            }
            if constexpr (OWN == 4 && kVoteSoftplus) {
                softplus4_vote(down, raw);
            } else if (OWN % 2 == 0 && kPackSoftplus) {
#pragma unroll
                for (int m = 0; m < OWN; m += 2) softplus2_f(down[m], down[m + 1], raw[m], raw[m + 1]);
            } else {
#pragma unroll
                for (int m = 0; m < OWN; ++m) down[m] = softplus_f(raw[m]);
            }
            if (!FULL) {
#pragma unroll
                for (int m = 0; m < OWN; ++m) down[m] = okd[m] ? down[m] : 0.f;
            }
            float dl[4], du[4], y[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                dl[i] = S == 1 ? down[i / S] : __shfl_sync(0xffffffffu, down[i / S], lane_base + (i % S));
                du[i] = dl[i] * uu[i];
            }
            if (SUMMARY) dsum += (dl[0] + dl[1]) + (dl[2] + dl[3]);
            // Per step, in explicit phases so that the shared-memory latency is paid once per step and not once
            // per state pair: (1) all B / C vectors of the step, (2) all exponents and inputs, (3) all exps,
            // (4) state update, (5) y with two independent accumulator pairs.
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4* bp = reinterpret_cast<const float4*>(ps + slot[i] * CP) + q;
                float4 bv[NS / 4], cv[NS / 4];
#pragma unroll
                for (int j4 = 0; j4 < NS / 4; ++j4) {
                    bv[j4] = bp[j4 * S];
                    if (!SUMMARY) cv[j4] = bp[4 + j4 * S];
                }
                float x[NS], w[NS];
#pragma unroll
                for (int j4 = 0; j4 < NS / 4; ++j4) {
                    const int j = j4 * 4;
                    mul2(x[j + 0], x[j + 1], dl[i], dl[i], Ap[j + 0], Ap[j + 1]);
                    mul2(x[j + 2], x[j + 3], dl[i], dl[i], Ap[j + 2], Ap[j + 3]);
                    mul2(w[j + 0], w[j + 1], du[i], du[i], bv[j4].x, bv[j4].y);
                    mul2(w[j + 2], w[j + 3], du[i], du[i], bv[j4].z, bv[j4].w);
                }
#pragma unroll
                for (int j = 0; j < NS; ++j) x[j] = ex2_approx(x[j]);
                float ye[2] = {0.f, 0.f}, yo[2] = {0.f, 0.f};
#pragma unroll
                for (int j4 = 0; j4 < NS / 4; ++j4) {
                    const int j = j4 * 4;
                    fma2(h[j + 0], h[j + 1], x[j + 0], x[j + 1], h[j + 0], h[j + 1], w[j + 0], w[j + 1]);
                    fma2(h[j + 2], h[j + 3], x[j + 2], x[j + 3], h[j + 2], h[j + 3], w[j + 2], w[j + 3]);
                    if (!SUMMARY) {
                        fma2(ye[0], yo[0], h[j + 0], h[j + 1], cv[j4].x, cv[j4].y, ye[0], yo[0]);
                        fma2(ye[1], yo[1], h[j + 2], h[j + 3], cv[j4].z, cv[j4].w, ye[1], yo[1]);
                    }
                }
                y[i] = (ye[0] + yo[0]) + (ye[1] + yo[1]);
            }
            if (!SUMMARY) {
#pragma unroll
                for (int off = S / 2; off > 0; off >>= 1) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) y[i] += __shfl_xor_sync(0xffffffffu, y[i], off);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (ok[i] && cvalid && q == (i % S))
                        yb[pos[i] * ystride] = from_f<y_t>(YSPLIT ? y[i] : fmaf(Dd, uu[i], y[i]));
                }
            }
        };
        const int nfull = nsteps & ~3;
        if (!single_col || S != 1) { for (int g0 = 0; g0 < nfull; g0 += 4) group(std::integral_constant<int, 3>{}, g0); }
        else if (!rev) { for (int g0 = 0; g0 < nfull; g0 += 4) group(std::integral_constant<int, 1>{}, g0); }
        else { for (int g0 = 0; g0 < nfull; g0 += 4) group(std::integral_constant<int, 2>{}, g0); }
        if (nfull < nsteps) group(std::integral_constant<int, 0>{}, nfull);
        if (PASS != 1 && p.hsave && cvalid) {
            float* hs = p.hsave + ((((int64_t)b * 4 + k) * p.NBmax + jb) * p.D + c) * kMaxState + 4 * q;
#pragma unroll
            for (int j4 = 0; j4 < NS / 4; ++j4)
                *reinterpret_cast<float4*>(hs + 4 * S * j4) = make_float4(h[4 * j4], h[4 * j4 + 1], h[4 * j4 + 2], h[4 * j4 + 3]);
        }
      }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }

    if (PASS == 1 && cvalid) {
        const int64_t sb = ((int64_t)(b * 4 + k) * (p.segs - 1) + seg) * p.D + c;
        float* hs = p.seg_h + sb * kMaxState + 4 * q;
#pragma unroll
        for (int j4 = 0; j4 < NS / 4; ++j4)
            *reinterpret_cast<float4*>(hs + 4 * S * j4) = make_float4(h[4 * j4], h[4 * j4 + 1], h[4 * j4 + 2], h[4 * j4 + 3]);
        if (q == 0) p.seg_dsum[sb] = dsum;
    }
}

struct CorePlan {
    int S, CT, tiles, T_row, NB_row, nw, T_col, NI_col, NO_col, cap, threads;
    int segs, bps_row, bps_col, ctas_per_sm, sub;
    size_t smem;
};

// Resident CTAs per SM of the one-lane build for a tile of `wpc` warps (128 registers: 16 warps per SM at most).
static inline int one_lane_ctas_per_sm(int wpc) { return 16 / wpc; }

// Channels per CTA for the one-lane-per-channel build.  The grid is persistent, so what matters is the steady-state
// cost of a resident set: with w warps per SM a step costs about 320 + 23 w cycles (measured at w = 6, 12, 15, 16:
// 460, 584, 684, 692; profiles/README.md) and CTA shapes other than 3..5 warps carry the measured penalty (1- and
// 2-warp CTAs repeat the proj loads and the per-block bookkeeping, 8-warp CTAs wait on their slowest warp before a
// stage is refilled).
static void plan_one_lane_tiles(int B, int D, int L, int& CT, int& tiles) {
    static const double shape_penalty[9] = {0, 1.19, 1.13, 1.02, 1.0, 1.03, 1.05, 1.08, 1.12};
    const int sms = num_sms();
    double best = 1e300;
    CT = 0; tiles = 0;
    for (int nt = 1; nt <= (D + 31) / 32; ++nt) {
        const int ct = ((D + nt - 1) / nt + 31) / 32 * 32;
        if (ct > 256 || (D + ct - 1) / ct != nt) continue;
        const int wpc = ct / 32;
        const int c = one_lane_ctas_per_sm(wpc);
        const long n = 4L * B * nt, per_round = (long)sms * c;
        const long full = n / per_round, rest = n % per_round;
        auto round_cost = [](int warps) { return 320.0 + 23.0 * warps; };
        double cost = full * round_cost(c * wpc);
        if (rest) cost += round_cost((int)((rest + sms - 1) / sms) * wpc);
        // lanes of the last tile that hold no channel still issue every instruction
        cost *= shape_penalty[wpc] * (L + 24) * ((double)nt * ct / D);
        if (cost < best * 0.999) { best = cost; CT = ct; tiles = nt; }
    }
}

static int env_int(const char* name, int lo, int hi, int dflt) {
    if (const char* e = getenv(name)) { const int v = atoi(e); if (v >= lo && v <= hi) return v; }
    return dflt;
}

// Lanes per channel (S) and channels per CTA (CT): enough warps to fill the machine, CTAs of at
// most 384 threads, TMA boxes of at most 256 channels.
static bool plan_core_tiles(int B, int D, int L, bool train, CorePlan& pl) {
    // One lane per channel whenever the sequences can be cut into enough segments to fill the machine (the L-parallel
    // passes); four lanes per channel only for launches that stay small even then (a handful of short sequences).
    const long warps1 = 4L * B * ((D + 31) / 32);                 // warps of the one-lane build, whole sequences
    const long want = 12L * num_sms();
    const long max_segs = train ? 1 : (L / 64 > kMaxSegs ? kMaxSegs : (L / 64 < 1 ? 1 : L / 64));
    int S = warps1 * max_segs >= want / 8 ? 1 : 4;
    S = env_int("MMB_CORE_S", 1, 4, S);
    if (S == 2 || S == 3) S = 4;
    const int gran = 32 / S;                          // channels per warp
    int CT, tiles;
    if (S == 1) {
        plan_one_lane_tiles(B, D, L, CT, tiles);
        const int v = env_int("MMB_CORE_CT", 32, 256, 0);
        if (v && v % 32 == 0) { CT = v; tiles = (D + CT - 1) / CT; }
    } else {
        int capc = 256 < 384 / S ? 256 : 384 / S;
        capc -= capc % gran;
        tiles = (D + capc - 1) / capc;
        CT = (D + tiles - 1) / tiles;
        CT = (CT + gran - 1) / gran * gran;
        tiles = (D + CT - 1) / CT;
    }
    if (CT > 256) return false;
    pl.S = S; pl.CT = CT; pl.tiles = tiles; pl.threads = CT * S;
    return true;
}

// Steps per stage (cap) and the block geometry of both views.  The ring is sized so that shared
// memory allows as many CTAs per SM as the register file does (regs = registers per thread of the
// instantiated kernel): the kernel lives on warps in flight, not on deep prefetch.
static bool plan_core_blocks(int H, int W, int RP, int regs, int XE, bool train, CorePlan& pl) {
    const int CP = 32 + RP;
    const int step_bytes = pl.CT * XE + CP * 4;
    const int regs_alloc = (regs + 7) / 8 * 8;
    int ctas = 65536 / (pl.threads * regs_alloc);
    if (ctas < 1) ctas = 1;
    if (ctas > 16) ctas = 16;
    int stage_bytes = (220 * 1024 / ctas - 1024) / kCoreStages;
    if (stage_bytes > kCoreStageBytes) stage_bytes = kCoreStageBytes;
    int cap = stage_bytes / step_bytes;
    if (cap > 32) cap = 32;
    if (cap < 8) cap = 8;
    cap = env_int("MMB_CORE_CAP", 4, 32, cap);
    pl.sub = 1;
    if (train) {                     // checkpoint spacing is part of the forward/backward contract: 8-step blocks, four per stage
        cap = kTrainCap;
        pl.sub = env_int("MMB_CORE_TRAIN_SUB", 1, 8, 4);
    }
    CoreGeom g;
    if (!core_geometry(H, W, cap, g)) return false;
    pl.T_row = g.T_row; pl.NB_row = g.NB_row; pl.nw = g.nw; pl.T_col = g.T_col; pl.NI_col = g.NI_col; pl.NO_col = g.NO_col;
    pl.cap = g.cap;
    const size_t xpad = ((size_t)pl.cap * pl.CT * XE + 127) & ~(size_t)127, ppad = ((size_t)pl.cap * CP * 4 + 127) & ~(size_t)127;
    if (pl.sub > 1) {                // as many checkpoint blocks per stage as the stage budget of this occupancy holds
        const int fit = (int)((size_t)stage_bytes / (xpad + ppad));
        if (pl.sub > fit) pl.sub = fit < 1 ? 1 : fit;
    }
    pl.smem = kCoreStages * pl.sub * (xpad + ppad) + 2 * kCoreStages * sizeof(uint64_t);
    pl.ctas_per_sm = ctas;
    return pl.smem <= 200 * 1024;
}

// Cycles of one pass over `units` CTAs of `steps` steps each: rounds of resident CTAs, a round
// with w warps per SM costing (320 + 23 w) cycles per step (the measured round model, profiles/README.md).
static double pass_cost(long units, double steps, int wpc, int occ, double factor) {
    const long sms = num_sms(), slots = sms * occ;
    double cost = 0;
    long left = units;
    while (left > 0) {
        const long now = left < slots ? left : slots;
        const int per_sm = (int)((now + sms - 1) / sms);
        cost += steps * factor * (320.0 + 23.0 * per_sm * wpc);
        left -= now;
        if (units > 8 * slots) { cost *= (double)units / (units - left); break; }     // many rounds: extrapolate
    }
    return cost;
}

// Segments per sequence for the L-parallel passes, from the round-cost model: pass 1 (summaries of G-1 segments,
// ~0.8 of a scan step: no C, no y) plus pass 2 (all G segments) against the whole sequences in one pass.  Chosen only
// when it wins by a margin -- the passes spend ~1.8x the exps, so they pay when the machine is mostly idle otherwise.
static void plan_segments(int B, int L, bool train, int occ, CorePlan& pl) {
    const int items = 4 * B * pl.tiles;
    const int NBc = pl.NO_col * pl.NI_col;
    const int NBmin = pl.NB_row < NBc ? pl.NB_row : NBc;
    const int wpc = pl.threads / 32;
    int G = 1;
    if (!train && pl.S == 1) {
        double best = pass_cost(items, L, wpc, occ, 1.0) * 0.8;
        const int gmax = NBmin / 2 < kMaxSegs ? NBmin / 2 : kMaxSegs;
        for (int g = 2; g <= gmax; ++g) {
            const double steps = (double)L / g;
            const double c = pass_cost((long)items * (g - 1), steps, wpc, occ, 0.8) + pass_cost((long)items * g, steps, wpc, occ, 1.0) +
                             25000.0 + 600.0 * g;       // second launch (~10 us of launch, prologue and parameter loads), carry prologue
            if (c < best) { best = c; G = g; }
        }
    }
    G = env_int("MMB_CORE_SEGS", 1, kMaxSegs, G);
    if (G > NBmin) G = NBmin;
    if (train || G < 1 || pl.S != 1) G = 1;
    pl.segs = G;
    pl.bps_row = (pl.NB_row + G - 1) / G;
    pl.bps_col = (NBc + G - 1) / G;
    // every segment must own at least one block in either view (an empty segment would leave its summary unwritten)
    while (pl.segs > 1 && ((pl.segs - 1) * pl.bps_row >= pl.NB_row || (pl.segs - 1) * pl.bps_col >= NBc)) {
        --pl.segs;
        pl.bps_row = (pl.NB_row + pl.segs - 1) / pl.segs;
        pl.bps_col = (NBc + pl.segs - 1) / pl.segs;
    }
}

struct CoreWorkspace {      // carve-up of the caller's workspace buffer
    size_t seg_h, seg_dsum, total;
};

static CoreWorkspace core_workspace(int B, int D, const CorePlan& pl) {
    CoreWorkspace w;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t off = 0;
    const size_t nseg = pl.segs > 1 ? (size_t)pl.segs - 1 : 0;
    w.seg_h = off; off += up((size_t)B * 4 * nseg * D * kMaxState * sizeof(float));
    w.seg_dsum = off; off += up((size_t)B * 4 * nseg * D * sizeof(float));
    w.total = off + 256;
    return w;
}

template <int S, int RP, typename xc_t, typename y_t, int PASS, int MB>
static int launch_pass(const CorePlan& pl, const CoreFwdParams& p, const CUtensorMap* tm, int segs_in_grid, cudaStream_t st) {
    auto kern = ss2d_core_fwd_kernel<S, RP, xc_t, y_t, PASS, MB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return cuda_status(e);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    dim3 grid(pl.tiles * segs_in_grid, 4, p.B);
    kern<<<grid, pl.threads, pl.smem, st>>>(tm[0], tm[1], tm[2], tm[3], p);
    return launch_status();
}

template <int S, int RP, typename xc_t, int MB>
static int regs_of() {
    static int regs = 0;
    if (regs == 0) {
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, ss2d_core_fwd_kernel<S, RP, xc_t, xc_t, 0, MB>) != cudaSuccess) {
            cudaGetLastError();
            return S == 1 ? 128 : 152;      // no device (host-only planning): the budgets the kernels are compiled for
        }
        regs = fa.numRegs;
    }
    return regs;
}

template <int S, int RP, typename xc_t, typename y_t, int MB = 1>
static int run_core(CorePlan& pl, CoreFwdParams& p, const void* xc, const float* proj, void* workspace,
                    int64_t workspace_bytes, bool plan_only, cudaStream_t st) {
    constexpr int CP = 32 + RP;
    constexpr uint64_t XE = sizeof(xc_t);
    const bool train = p.hsave != nullptr;
    const int regs = regs_of<S, RP, xc_t, MB>();
    if (regs < 0) return MMB_ERR_UNSUPPORTED;
    if (!plan_core_blocks(p.H, p.W, RP, regs, (int)XE, train, pl)) return MMB_ERR_UNSUPPORTED;
    // resident CTAs per SM of this kernel shape (what the segment planner's round model needs)
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ss2d_core_fwd_kernel<S, RP, xc_t, xc_t, 0, MB>, pl.threads, pl.smem) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        occ = pl.ctas_per_sm > 0 ? pl.ctas_per_sm : 1;
    }
    plan_segments(p.B, p.L, train, occ, pl);
    const CoreWorkspace ws = core_workspace(p.B, p.D, pl);
    if (plan_only) return MMB_OK;
    if (pl.segs > 1 && ((int64_t)ws.total > workspace_bytes || !workspace)) return MMB_ERR_INVALID_ARG;
    if (train) { CoreGeom g; core_geometry(p.H, p.W, kTrainCap, g); p.NBmax = g.nblocks_max(); }
    p.T_row = pl.T_row; p.NB_row = pl.NB_row; p.nw = pl.nw; p.T_col = pl.T_col; p.NI_col = pl.NI_col; p.NO_col = pl.NO_col;
    p.cap = pl.cap; p.CT = pl.CT; p.tiles = pl.tiles; p.sub = pl.sub;
    p.segs = pl.segs; p.bps_row = pl.bps_row; p.bps_col = pl.bps_col;
    uint8_t* wsb = reinterpret_cast<uint8_t*>(workspace);
    p.seg_h = pl.segs > 1 ? reinterpret_cast<float*>(wsb + ws.seg_h) : nullptr;
    p.seg_dsum = pl.segs > 1 ? reinterpret_cast<float*>(wsb + ws.seg_dsum) : nullptr;
    const CUtensorMapDataType xdt = XE == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUtensorMap tm[4];
    const uint64_t B = p.B, H = p.H, W = p.W, L = p.L, D = p.D;
    {
        const uint64_t dims[3] = {D, L, B}, str[2] = {D * XE, L * D * XE};
        const uint32_t box[3] = {(uint32_t)pl.CT, (uint32_t)pl.T_row, 1};
        if (!make_tmap(&tm[0], xdt, 3, xc, dims, str, box)) return MMB_ERR_UNSUPPORTED;
    }
    {
        const uint64_t dims[4] = {D, W, H, B}, str[3] = {D * XE, W * D * XE, L * D * XE};
        const uint32_t box[4] = {(uint32_t)pl.CT, (uint32_t)pl.nw, (uint32_t)pl.T_col, 1};
        if (!make_tmap(&tm[1], xdt, 4, xc, dims, str, box)) return MMB_ERR_UNSUPPORTED;
    }
    {
        const uint64_t dims[4] = {CP, 4, L, B}, str[3] = {CP * 4, 4 * CP * 4, L * 4 * CP * 4};
        const uint32_t box[4] = {CP, 1, (uint32_t)pl.T_row, 1};
        if (!make_tmap(&tm[2], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, proj, dims, str, box)) return MMB_ERR_UNSUPPORTED;
    }
    {
        const uint64_t dims[5] = {CP, 4, W, H, B}, str[4] = {CP * 4, 4 * CP * 4, W * 4 * CP * 4, L * 4 * CP * 4};
        const uint32_t box[5] = {CP, 1, (uint32_t)pl.nw, (uint32_t)pl.T_col, 1};
        if (!make_tmap(&tm[3], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, proj, dims, str, box)) return MMB_ERR_UNSUPPORTED;
    }
    if (pl.segs == 1) return launch_pass<S, RP, xc_t, y_t, 0, MB>(pl, p, tm, 1, st);
    if constexpr (S == 1) {
        int rc = launch_pass<S, RP, xc_t, y_t, 1, MB>(pl, p, tm, pl.segs - 1, st);
        if (rc != MMB_OK) return rc;
        return launch_pass<S, RP, xc_t, y_t, 2, MB>(pl, p, tm, pl.segs, st);
    } else {
        return MMB_ERR_UNSUPPORTED;
    }
}

template <int RP, typename xc_t, typename y_t>
static int dispatch_core_s(CorePlan& pl, CoreFwdParams& p, const void* xc, const float* proj, void* ws, int64_t wsb,
                           bool plan_only, cudaStream_t st) {
    switch (pl.S) {
        // one lane per channel: 128-register budget (2 CTAs of 256 max; the register sweep is in profiles/README.md)
        case 1: return run_core<1, RP, xc_t, y_t, 2>(pl, p, xc, proj, ws, wsb, plan_only, st);
        default: return run_core<4, RP, xc_t, y_t>(pl, p, xc, proj, ws, wsb, plan_only, st);
    }
}

template <typename xc_t, typename y_t>
static int dispatch_core_rp(int dt_pad, CorePlan& pl, CoreFwdParams& p, const void* xc, const float* proj, void* ws,
                            int64_t wsb, bool plan_only, cudaStream_t st) {
    switch (dt_pad) {
        case 4: return dispatch_core_s<4, xc_t, y_t>(pl, p, xc, proj, ws, wsb, plan_only, st);
        case 8: return dispatch_core_s<8, xc_t, y_t>(pl, p, xc, proj, ws, wsb, plan_only, st);
        case 12: return dispatch_core_s<12, xc_t, y_t>(pl, p, xc, proj, ws, wsb, plan_only, st);
        case 16: return dispatch_core_s<16, xc_t, y_t>(pl, p, xc, proj, ws, wsb, plan_only, st);
        case 24: return dispatch_core_s<24, xc_t, y_t>(pl, p, xc, proj, ws, wsb, plan_only, st);
        case 32: return dispatch_core_s<32, xc_t, y_t>(pl, p, xc, proj, ws, wsb, plan_only, st);
        default: return MMB_ERR_UNSUPPORTED;
    }
}

static int core_entry(const void* xc, const float* proj, const float* Wdt, const float* dt_bias, const float* A,
                      const float* Ds, void* ydir, float* hsave, void* workspace, int64_t workspace_bytes, int batch,
                      int H, int W, int D, int dstate, int dt_rank, int dt_pad, int xc_dtype, int y_dtype, bool train,
                      bool plan_only, CorePlan& pl, cudaStream_t st) {
    if (batch < 0 || H <= 0 || W <= 0 || D <= 0 || dstate <= 0 || dt_rank <= 0) return MMB_ERR_INVALID_ARG;
    if (dstate > kMaxState || dt_pad != mmb_ss2d_core_dt_pad(dt_rank)) return MMB_ERR_UNSUPPORTED;
    if (xc_dtype != MMB_F32 && xc_dtype != MMB_BF16) return MMB_ERR_UNSUPPORTED;
    if (D % (xc_dtype == MMB_F32 ? 4 : 8) != 0 || batch > 65535) return MMB_ERR_UNSUPPORTED;
    if (batch == 0) { pl = CorePlan{}; return MMB_OK; }
    if (!plan_core_tiles(batch, D, H * W, train, pl)) return MMB_ERR_UNSUPPORTED;
    CoreFwdParams p{};
    p.ydir = ydir; p.Wdt = Wdt; p.bias = dt_bias; p.A = A; p.Ds = Ds; p.hsave = train ? hsave : nullptr; p.NBmax = 0;
    if (plan_only && train) p.hsave = reinterpret_cast<float*>(1);     // geometry only; never dereferenced
    p.B = batch; p.H = H; p.W = W; p.L = H * W; p.D = D; p.N = dstate; p.R = dt_rank; p.CT = pl.CT;
    if (y_dtype != MMB_F32 && !(y_dtype == MMB_BF16 && xc_dtype == MMB_BF16)) return MMB_ERR_UNSUPPORTED;
    if (xc_dtype == MMB_F32) return dispatch_core_rp<float, float>(dt_pad, pl, p, xc, proj, workspace, workspace_bytes, plan_only, st);
    if (y_dtype == MMB_F32) return dispatch_core_rp<__nv_bfloat16, float>(dt_pad, pl, p, xc, proj, workspace, workspace_bytes, plan_only, st);
    return dispatch_core_rp<__nv_bfloat16, __nv_bfloat16>(dt_pad, pl, p, xc, proj, workspace, workspace_bytes, plan_only, st);
}

}  // namespace mmb

extern "C" int mmb_ss2d_core_dt_pad(int dt_rank) {
    const int sup[6] = {4, 8, 12, 16, 24, 32};
    for (int i = 0; i < 6; ++i) if (dt_rank <= sup[i]) return sup[i];
    return MMB_ERR_UNSUPPORTED;
}

extern "C" int mmb_ss2d_core_train_blocks(int H, int W) {
    if (H <= 0 || W <= 0) return MMB_ERR_INVALID_ARG;
    mmb::CoreGeom g;
    if (!mmb::core_geometry(H, W, mmb::kTrainCap, g)) return MMB_ERR_UNSUPPORTED;
    return g.nblocks_max();
}

extern "C" int mmb_ss2d_core_plan(int batch, int H, int W, int D, int* lanes_per_channel, int* channels_per_cta,
                                  int* channel_tiles) {
    using namespace mmb;
    if (batch <= 0 || H <= 0 || W <= 0 || D <= 0 || !lanes_per_channel || !channels_per_cta || !channel_tiles)
        return MMB_ERR_INVALID_ARG;
    CorePlan pl;
    if (!plan_core_tiles(batch, D, H * W, false, pl)) return MMB_ERR_UNSUPPORTED;
    *lanes_per_channel = pl.S; *channels_per_cta = pl.CT; *channel_tiles = pl.tiles;
    return MMB_OK;
}

extern "C" int64_t mmb_ss2d_core_fwd_workspace_bytes(int batch, int H, int W, int D, int dstate, int dt_rank, int xc_dtype,
                                                     int save_states, int* segments, int* ctas_per_sm) {
    using namespace mmb;
    CorePlan pl{};
    const int rp = mmb_ss2d_core_dt_pad(dt_rank);
    if (rp < 0) return rp;
    const int rc = core_entry(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, batch, H, W, D,
                              dstate, dt_rank, rp, xc_dtype, xc_dtype, save_states != 0, true, pl, nullptr);
    if (rc != MMB_OK) return rc;
    if (segments) *segments = batch ? pl.segs : 1;
    if (ctas_per_sm) *ctas_per_sm = batch ? pl.ctas_per_sm : 0;
    if (batch == 0) return 256;
    return (int64_t)core_workspace(batch, D, pl).total;
}

extern "C" int mmb_ss2d_core_fwd(const void* xc, const float* proj, const float* Wdt, const float* dt_bias,
                                 const float* A, const float* Ds, void* ydir, float* hsave, void* workspace,
                                 int64_t workspace_bytes, int batch, int H, int W, int D, int dstate, int dt_rank,
                                 int dt_pad, int xc_dtype, int ydir_dtype, void* stream) {
    using namespace mmb;
    if (!xc || !proj || !Wdt || !dt_bias || !A || !Ds || !ydir) return MMB_ERR_INVALID_ARG;
    if ((reinterpret_cast<uintptr_t>(xc) | reinterpret_cast<uintptr_t>(proj) | reinterpret_cast<uintptr_t>(workspace) |
         reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Wdt)) % 16 != 0)
        return MMB_ERR_UNSUPPORTED;
    CorePlan pl{};
    return core_entry(xc, proj, Wdt, dt_bias, A, Ds, ydir, hsave, workspace, workspace_bytes, batch, H, W, D, dstate,
                      dt_rank, dt_pad, xc_dtype, ydir_dtype, hsave != nullptr, false, pl, reinterpret_cast<cudaStream_t>(stream));
}
