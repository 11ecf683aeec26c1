// Fused SS2D core, forward: cross-scan + dt_proj + softplus + selective scan + cross-merge addressing.
// Replaces MedMamba.py:256-257 (cross-scan), :262,:266 (dt_proj einsum + copy), :273-279
// (selective_scan_fn) and :282-286 (flip / transpose back) for all four directions in one launch.
//
// Layout (channels-last, everything indexed by the token position p = h*W + w, never by the
// per-direction sequence index):
//   xc   (B, H, W, D)        conv+SiLU output = u for every direction             [fp32]
//   proj (B, H, W, 4, CP)    x_proj of the un-permuted tokens, per direction k:
//                            [ B_n (16) | C_n (16) | dt_r (RP, zero padded) ]      [fp32]
//   ydir (B, H, W, 4, D)     scan output of direction k written at the position it belongs to;
//                            the out_norm kernel adds the four in the reference's order.
// Direction k walks the positions in the order of SURVEY.md Appendix A (row-major, column-major
// and their reverses), so no flipped or transposed copy of anything is ever materialised.
//
// One CTA owns (batch b, direction k, CT channels) for the whole sequence.  Blocks of <= 32
// consecutive steps of xc and proj are streamed into a shared-memory ring with TMA tensor copies
// (a row block for k = 0, 2; a column block for k = 1, 3) signalled through mbarriers; thread 0
// re-arms a stage as soon as every warp has released it (no dedicated producer warp: its
// registers would be a quarter of the register file at 3 warps per CTA).  Every warp holds one
// channel per S lanes with the 16/S states of the lane in registers.
// Per step and state: one MUFU.EX2 and four FMA-pipe operations; the kernel is bound by the
// 16 exp/clk/SM MUFU rate, not by HBM.
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

struct CoreFwdParams {
    float* ydir;
    const float* Wdt;     // (4, D, R)
    const float* bias;    // (4, D)
    const float* A;       // (4*D, N)
    const float* Ds;      // (4*D)
    int B, H, W, L, D, N, R, CT;
    int T_row, NB_row;                 // row view: NB_row blocks of T_row consecutive positions
    int nw, T_col, NI_col, NO_col;     // column view: NO_col column groups x NI_col row blocks
    int cap;                           // steps a stage can hold
    int kmask;                         // debug: directions to run (bit k); 15 in production
    int dbg;                           // debug flags: 1 = try_wait instead of polling
    float* hsave;                      // NULL, or (B, 4, NBmax, D, 16): state after every block (training)
    int NBmax;
};

template <int S, int RP, typename xc_t, int MB>
__global__ void __launch_bounds__(S == 1 ? 256 : 384, MB)
ss2d_core_fwd_kernel(const __grid_constant__ CUtensorMap tmx_row, const __grid_constant__ CUtensorMap tmx_col,
                     const __grid_constant__ CUtensorMap tmp_row, const __grid_constant__ CUtensorMap tmp_col,
                     const CoreFwdParams p) {
    constexpr int NS = kMaxState / S;
    constexpr int CP = 32 + RP;
    constexpr int OWN = 4 / S;          // delta evaluations per lane per group of 4 steps

    extern __shared__ __align__(128) uint8_t smem_raw[];
    constexpr int XE = (int)sizeof(xc_t);
    const int xbytes = p.cap * p.CT * XE, pbytes = p.cap * CP * 4;
    const int xpad = (xbytes + 127) & ~127, ppad = (pbytes + 127) & ~127;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kCoreStages * (xpad + ppad));
    uint64_t* empty = full + kCoreStages;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nwarps = blockDim.x >> 5;
    const int k = blockIdx.y, b = blockIdx.z, c0 = blockIdx.x * p.CT;
    const bool colview = (k & 1) != 0, rev = k >= 2;
    const int NB = colview ? p.NO_col * p.NI_col : p.NB_row;
    if (!((p.kmask >> k) & 1)) return;

    // Block jb (time order) -> TMA loads into stage jb % kCoreStages; issued by thread 0 only.
    auto issue = [&](int jb) {
        const int s = jb % kCoreStages;
        const int blk = rev ? NB - 1 - jb : jb;
        uint8_t* xs = smem_raw + s * (xpad + ppad);
        uint8_t* ps = xs + xpad;
        if (!colview) {
            mbar_expect_tx(&full[s], p.T_row * (p.CT * XE + CP * 4));
            tma_load_3d(xs, &tmx_row, &full[s], c0, blk * p.T_row, b);
            tma_load_4d(ps, &tmp_row, &full[s], 0, k, blk * p.T_row, b);
        } else {
            const int o = blk / p.NI_col, i = blk % p.NI_col;
            mbar_expect_tx(&full[s], p.nw * p.T_col * (p.CT * XE + CP * 4));
            tma_load_4d(xs, &tmx_col, &full[s], c0, o * p.nw, i * p.T_col, b);
            tma_load_5d(ps, &tmp_col, &full[s], 0, k, o * p.nw, i * p.T_col, b);
        }
    };

    if (tid == 0) {
        prefetch_tmap(colview ? &tmx_col : &tmx_row);
        prefetch_tmap(colview ? &tmp_col : &tmp_row);
        for (int s = 0; s < kCoreStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], nwarps); }
        mbar_fence_init();
        for (int jb = 0; jb < kCoreStages - 1 && jb < NB; ++jb) issue(jb);
    }
    __syncthreads();

    const int t = warp * 32 + lane;
    const int cl = t / S, q = t % S;            // channel inside the tile, state split index
    const int c = c0 + cl;
    const bool cvalid = c < p.D;
    const int row = k * p.D + (cvalid ? c : 0);
    const int lane_base = lane & ~(S - 1);

    float Ap[NS], h[NS], Wd[RP];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        const int n = (j & 3) + 4 * q + 4 * S * (j >> 2);
        Ap[j] = (cvalid && n < p.N) ? p.A[(int64_t)row * p.N + n] * kLog2e : 0.f;
        h[j] = 0.f;
    }
#pragma unroll
    for (int r = 0; r < RP; ++r) Wd[r] = (cvalid && r < p.R) ? p.Wdt[(int64_t)row * p.R + r] : 0.f;
    const float bias = cvalid ? p.bias[row] : 0.f;
    const float Dd = cvalid ? p.Ds[row] : 0.f;
    float* yb = p.ydir + ((int64_t)b * p.L * 4 + k) * p.D + c;   // + pos * 4 * D
    const int64_t ystride = 4 * (int64_t)p.D;

    for (int jb = 0; jb < NB; ++jb) {
        const int s = jb % kCoreStages, ph = (jb / kCoreStages) & 1;
        // refill the stage block jb-1 has just left (inline producer: one thread)
        if (tid == 0 && jb + kCoreStages - 1 < NB) {
            const int jn = jb + kCoreStages - 1;
            if (jb > 0) mbar_wait(&empty[jn % kCoreStages], ((jb - 1) / kCoreStages) & 1);
            issue(jn);
        }
        __syncwarp();
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
        const xc_t* xs = reinterpret_cast<const xc_t*>(smem_raw + s * (xpad + ppad)) + cl;
        const float* ps = reinterpret_cast<const float*>(smem_raw + s * (xpad + ppad) + xpad);
        if (p.dbg & 1) { while (!mbar_try_wait(&full[s], ph)) {} } else mbar_wait(&full[s], ph);

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
            // Per step, in explicit phases so that the shared-memory latency is paid once per step and not once
            // per state pair: (1) all B / C vectors of the step, (2) all exponents and inputs, (3) all exps,
            // (4) state update, (5) y with two independent accumulator pairs.
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4* bp = reinterpret_cast<const float4*>(ps + slot[i] * CP) + q;
                float4 bv[NS / 4], cv[NS / 4];
#pragma unroll
                for (int j4 = 0; j4 < NS / 4; ++j4) { bv[j4] = bp[j4 * S]; cv[j4] = bp[4 + j4 * S]; }
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
                    fma2(ye[0], yo[0], h[j + 0], h[j + 1], cv[j4].x, cv[j4].y, ye[0], yo[0]);
                    fma2(ye[1], yo[1], h[j + 2], h[j + 3], cv[j4].z, cv[j4].w, ye[1], yo[1]);
                }
                y[i] = (ye[0] + yo[0]) + (ye[1] + yo[1]);
            }
#pragma unroll
            for (int off = S / 2; off > 0; off >>= 1) {
#pragma unroll
                for (int i = 0; i < 4; ++i) y[i] += __shfl_xor_sync(0xffffffffu, y[i], off);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (ok[i] && cvalid && q == (i % S)) yb[pos[i] * ystride] = fmaf(Dd, uu[i], y[i]);
            }
        };
        const int nfull = nsteps & ~3;
        if (!single_col || S != 1) { for (int g0 = 0; g0 < nfull; g0 += 4) group(std::integral_constant<int, 3>{}, g0); }
        else if (!rev) { for (int g0 = 0; g0 < nfull; g0 += 4) group(std::integral_constant<int, 1>{}, g0); }
        else { for (int g0 = 0; g0 < nfull; g0 += 4) group(std::integral_constant<int, 2>{}, g0); }
        if (nfull < nsteps) group(std::integral_constant<int, 0>{}, nfull);
        if (p.hsave && cvalid) {
            float* hs = p.hsave + ((((int64_t)b * 4 + k) * p.NBmax + jb) * p.D + c) * kMaxState + 4 * q;
#pragma unroll
            for (int j4 = 0; j4 < NS / 4; ++j4)
                *reinterpret_cast<float4*>(hs + 4 * S * j4) = make_float4(h[4 * j4], h[4 * j4 + 1], h[4 * j4 + 2], h[4 * j4 + 3]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
}

struct CorePlan {
    int S, CT, tiles, T_row, NB_row, nw, T_col, NI_col, NO_col, cap, threads;
    size_t smem;
};

// Channels per CTA for the one-lane-per-channel build (128 registers: 16 warps per SM at most).
// Cost model fitted to the CT sweeps in profiles/README.md (round 1, session 2): a round of resident CTAs with
// w warps per SM costs about 320 + 23 w cycles per step (measured at w = 6, 12, 15, 16: 460, 584, 684, 692; the
// MUFU floor is 36 w: 18 MUFU x 8 cycles per warp-step on 4 sub-partitions), the CTAs left over after the full
// rounds spread over the SMs, and
// CTA shapes other than 3..5 warps carry the measured penalty (1- and 2-warp CTAs repeat the proj loads and
// the per-block bookkeeping, 8-warp CTAs wait on their slowest warp before a stage is refilled).
static void plan_one_lane_tiles(int B, int D, int L, int& CT, int& tiles) {
    static const double shape_penalty[9] = {0, 1.19, 1.13, 1.02, 1.0, 1.03, 1.05, 1.08, 1.12};
    const int sms = num_sms();
    double best = 1e300;
    CT = 0; tiles = 0;
    for (int nt = 1; nt <= (D + 31) / 32; ++nt) {
        const int ct = ((D + nt - 1) / nt + 31) / 32 * 32;
        if (ct > 256 || (D + ct - 1) / ct != nt) continue;
        const int wpc = ct / 32;
        const int c = 16 / wpc;
        const long n = 4L * B * nt, per_round = (long)sms * c;
        const long full = n / per_round, rest = n % per_round;
        auto round_cost = [](int warps) { return 320.0 + 23.0 * warps; };
        double cost = full * round_cost(c * wpc);
        if (rest) cost += round_cost((int)((rest + sms - 1) / sms) * wpc);
        cost *= shape_penalty[wpc] * (L + 24);
        if (cost < best * 0.999) { best = cost; CT = ct; tiles = nt; }
    }
}

// Lanes per channel (S) and channels per CTA (CT): enough warps to fill the machine, CTAs of at
// most 384 threads, TMA boxes of at most 256 channels.
static bool plan_core_tiles(int B, int D, int L, CorePlan& pl) {
    // Measured (profiles/README.md, S sweep): one lane per channel wins as soon as there are ~3 warps of rows
    // per SM; below that the launch is latency-bound and splitting the 16 states over 4 lanes helps.
    const long rows = 4L * B * D;
    int S = rows >= 32L * 3 * num_sms() ? 1 : 4;
    if (const char* e = getenv("MMB_CORE_S")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) S = v; }
    const int gran = 32 / S;                          // channels per warp
    int CT, tiles;
    if (S == 1) {
        plan_one_lane_tiles(B, D, L, CT, tiles);
        if (const char* e = getenv("MMB_CORE_CT")) { const int v = atoi(e); if (v >= 32 && v <= 256 && v % 32 == 0) { CT = v; tiles = (D + CT - 1) / CT; } }
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
    if (const char* e = getenv("MMB_CORE_CAP")) { const int v = atoi(e); if (v >= 4 && v <= 32) cap = v; }
    if (train) cap = kTrainCap;      // checkpoint spacing is part of the forward/backward contract
    CoreGeom g;
    if (!core_geometry(H, W, cap, g)) return false;
    pl.T_row = g.T_row; pl.NB_row = g.NB_row; pl.nw = g.nw; pl.T_col = g.T_col; pl.NI_col = g.NI_col; pl.NO_col = g.NO_col;
    pl.cap = g.cap;
    const size_t xpad = ((size_t)pl.cap * pl.CT * XE + 127) & ~(size_t)127, ppad = ((size_t)pl.cap * CP * 4 + 127) & ~(size_t)127;
    pl.smem = kCoreStages * (xpad + ppad) + 2 * kCoreStages * sizeof(uint64_t);
    return pl.smem <= 200 * 1024;
}

template <int S, int RP, typename xc_t, int MB = 1>
static int launch_core(CorePlan& pl, CoreFwdParams& p, const void* xc, const float* proj, cudaStream_t st) {
    constexpr int CP = 32 + RP;
    constexpr uint64_t XE = sizeof(xc_t);
    const CUtensorMapDataType xdt = XE == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    auto kern = ss2d_core_fwd_kernel<S, RP, xc_t, MB>;
    static int regs = 0;
    if (regs == 0) {
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, kern);
        if (e != cudaSuccess) return cuda_status(e);
        regs = fa.numRegs;
    }
    if (!plan_core_blocks(p.H, p.W, RP, regs, (int)XE, p.hsave != nullptr, pl)) return MMB_ERR_UNSUPPORTED;
    if (p.hsave) { CoreGeom g; core_geometry(p.H, p.W, kTrainCap, g); p.NBmax = g.nblocks_max(); }
    p.T_row = pl.T_row; p.NB_row = pl.NB_row; p.nw = pl.nw; p.T_col = pl.T_col; p.NI_col = pl.NI_col; p.NO_col = pl.NO_col;
    p.cap = pl.cap;
    CUtensorMap tmx_row, tmx_col, tmp_row, tmp_col;
    const uint64_t B = p.B, H = p.H, W = p.W, L = p.L, D = p.D;
    {
        const uint64_t dims[3] = {D, L, B}, str[2] = {D * XE, L * D * XE};
        const uint32_t box[3] = {(uint32_t)pl.CT, (uint32_t)pl.T_row, 1};
        if (!make_tmap(&tmx_row, xdt, 3, xc, dims, str, box)) return MMB_ERR_UNSUPPORTED;
    }
    {
        const uint64_t dims[4] = {D, W, H, B}, str[3] = {D * XE, W * D * XE, L * D * XE};
        const uint32_t box[4] = {(uint32_t)pl.CT, (uint32_t)pl.nw, (uint32_t)pl.T_col, 1};
        if (!make_tmap(&tmx_col, xdt, 4, xc, dims, str, box)) return MMB_ERR_UNSUPPORTED;
    }
    {
        const uint64_t dims[4] = {CP, 4, L, B}, str[3] = {CP * 4, 4 * CP * 4, L * 4 * CP * 4};
        const uint32_t box[4] = {CP, 1, (uint32_t)pl.T_row, 1};
        if (!make_tmap(&tmp_row, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, proj, dims, str, box)) return MMB_ERR_UNSUPPORTED;
    }
    {
        const uint64_t dims[5] = {CP, 4, W, H, B}, str[4] = {CP * 4, 4 * CP * 4, W * 4 * CP * 4, L * 4 * CP * 4};
        const uint32_t box[5] = {CP, 1, (uint32_t)pl.nw, (uint32_t)pl.T_col, 1};
        if (!make_tmap(&tmp_col, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, proj, dims, str, box)) return MMB_ERR_UNSUPPORTED;
    }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
    if (e != cudaSuccess) return cuda_status(e);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    dim3 grid(pl.tiles, 4, p.B);
    kern<<<grid, pl.threads, pl.smem, st>>>(tmx_row, tmx_col, tmp_row, tmp_col, p);
    return launch_status();
}

template <int RP, typename xc_t>
static int dispatch_core_s(CorePlan& pl, CoreFwdParams& p, const void* xc, const float* proj, cudaStream_t st) {
    switch (pl.S) {
        // one lane per channel: 128-register budget (2 CTAs of 256 threads; the register sweep is in profiles/README.md)
        case 1: return launch_core<1, RP, xc_t, 2>(pl, p, xc, proj, st);
        case 2: return launch_core<2, RP, xc_t>(pl, p, xc, proj, st);
        default: return launch_core<4, RP, xc_t>(pl, p, xc, proj, st);
    }
}

template <typename xc_t>
static int dispatch_core_rp(int dt_pad, CorePlan& pl, CoreFwdParams& p, const void* xc, const float* proj, cudaStream_t st) {
    switch (dt_pad) {
        case 4: return dispatch_core_s<4, xc_t>(pl, p, xc, proj, st);
        case 8: return dispatch_core_s<8, xc_t>(pl, p, xc, proj, st);
        case 12: return dispatch_core_s<12, xc_t>(pl, p, xc, proj, st);
        case 16: return dispatch_core_s<16, xc_t>(pl, p, xc, proj, st);
        case 24: return dispatch_core_s<24, xc_t>(pl, p, xc, proj, st);
        case 32: return dispatch_core_s<32, xc_t>(pl, p, xc, proj, st);
        default: return MMB_ERR_UNSUPPORTED;
    }
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
    if (!plan_core_tiles(batch, D, H * W, pl)) return MMB_ERR_UNSUPPORTED;
    *lanes_per_channel = pl.S; *channels_per_cta = pl.CT; *channel_tiles = pl.tiles;
    return MMB_OK;
}

extern "C" int mmb_ss2d_core_fwd(const void* xc, const float* proj, const float* Wdt, const float* dt_bias,
                                 const float* A, const float* Ds, float* ydir, float* hsave,
                                 int batch, int H, int W, int D, int dstate, int dt_rank, int dt_pad, int xc_dtype, void* stream) {
    using namespace mmb;
    if (!xc || !proj || !Wdt || !dt_bias || !A || !Ds || !ydir) return MMB_ERR_INVALID_ARG;
    if (batch < 0 || H <= 0 || W <= 0 || D <= 0 || dstate <= 0 || dt_rank <= 0) return MMB_ERR_INVALID_ARG;
    if (dstate > kMaxState || dt_pad != mmb_ss2d_core_dt_pad(dt_rank)) return MMB_ERR_UNSUPPORTED;
    if (xc_dtype != MMB_F32 && xc_dtype != MMB_BF16) return MMB_ERR_UNSUPPORTED;
    if (D % (xc_dtype == MMB_F32 ? 4 : 8) != 0 || batch > 65535) return MMB_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(xc) | reinterpret_cast<uintptr_t>(proj)) % 16 != 0) return MMB_ERR_UNSUPPORTED;
    if (batch == 0) return MMB_OK;
    CorePlan pl;
    if (!plan_core_tiles(batch, D, H * W, pl)) return MMB_ERR_UNSUPPORTED;
    CoreFwdParams p;
    p.ydir = ydir; p.Wdt = Wdt; p.bias = dt_bias; p.A = A; p.Ds = Ds; p.hsave = hsave; p.NBmax = 0;
    p.B = batch; p.H = H; p.W = W; p.L = H * W; p.D = D; p.N = dstate; p.R = dt_rank; p.CT = pl.CT;
    p.kmask = 15;
    if (const char* e = getenv("MMB_CORE_KMASK")) p.kmask = atoi(e);
    p.dbg = 0;
    if (const char* e = getenv("MMB_CORE_DBG")) p.dbg = atoi(e);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (xc_dtype == MMB_F32) return dispatch_core_rp<float>(dt_pad, pl, p, xc, proj, st);
    return dispatch_core_rp<__nv_bfloat16>(dt_pad, pl, p, xc, proj, st);
}
