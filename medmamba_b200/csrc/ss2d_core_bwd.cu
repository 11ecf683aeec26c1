// Fused SS2D core, backward.  Gradient of mmb_ss2d_core_fwd w.r.t. xc (per direction), proj (dB_n, dC_n,
// d dt_r), Wdt, dt_bias, A and D.  Formulas: SURVEY.md Appendix B; the cross-scan / cross-merge index maps are
// the forward's (Appendix A) -- the backward of a gather by index is a scatter to the same index,
// so "cross-merge of the per-direction du" is again just the store address.
//
// Same ownership and TMA ring as the forward: one lane per channel with the 16 states in registers, a CTA of up
// to three warps per (batch, direction, 96 channels), warps independent of each other (no CTA barrier in the
// loop).  The sequence is walked BACKWARDS in the blocks of kTrainCap = 8 steps whose end states the forward
// checkpointed (hsave).  A block is processed as two half blocks of 4 steps, second half first:
//   pass 1   from the checkpoint, advance over the first half to get the mid-block state (no history kept);
//   phase A  re-run the half block forward, parking h_{t-1} of every step in the warp's shared-memory history;
//   phase B  the reverse recurrence g_t = dy_t C_t + a_{t+1} g_{t+1} and all gradients (a_t recomputed).
// That is 2.5 forward passes of exps; the MUFU pipe has room (the cost of a backward is instructions).
// dB_n / dC_n / d dt_r need a sum over the channels of the (batch, direction).  Each warp reduces its lanes' 32 + RP
// values per step with a shuffle butterfly (a reduce-scatter: at every stage a lane keeps one half of its values and
// sends the other; the pairing order -- dB_n with dC_n first, then adjacent states -- lets every stage run as soon as
// its two operands exist, so the values never pile up in registers) and lane v stores the warp's sum of value v: one
// partial row per warp-sized channel group to HBM, the host adds the groups.  Round 1 transposed the values through
// shared memory instead: 36-56 STS + 8-16 LDS.128 per step against 31-37 SHFL now, and the 5 KB tile per warp was what
// kept a fourth CTA off the SM.  dA, dD, dWdt, d dt_bias are per-thread accumulators, one partial per batch element.
// No float atomics: bit-reproducible.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "core_geom.cuh"
#include "tma.cuh"

namespace mmb {

constexpr int kBwdStages = 3;
constexpr int kHalf = kTrainCap / 2;      // steps per half block

struct CoreBwdParams {
    const float* Wdt; const float* bias; const float* A; const float* Ds; const float* hsave;
    float* dudir;                   // (B, L, 4, D)
    float* dproj;                   // (groups, B, L, 4, CP): [dB_n | dC_n | d dt_r], groups = ceil(D / 32)
    float* dA_part;                 // (B, 4D, N)
    float* dW_part;                 // (B, 4D, RP)
    float* dD_part; float* db_part; // (B, 4D)
    int B, H, W, L, D, N, R, CT, tiles, NBmax;
    int T_row, NB_row, nw, T_col, NI_col, NO_col, cap;
};

// keep `hi ? b : a` and exchange the other one with the lane `mask` away: one stage of the reduce-scatter butterfly
__device__ __forceinline__ float bfly(const float a, const float b, const bool hi, const int mask) {
    const float send = hi ? a : b, keep = hi ? b : a;
    return keep + __shfl_xor_sync(0xffffffffu, send, mask);
}

template <int S, int RP, typename xc_t>
__global__ void __launch_bounds__(96, S == 1 ? 4 : 5)
ss2d_core_bwd_kernel(const __grid_constant__ CUtensorMap tmx_row, const __grid_constant__ CUtensorMap tmx_col,
                     const __grid_constant__ CUtensorMap tmd_row, const __grid_constant__ CUtensorMap tmd_col,
                     const __grid_constant__ CUtensorMap tmp_row, const __grid_constant__ CUtensorMap tmp_col,
                     const CoreBwdParams p) {
    constexpr int NS = kMaxState / S, CP = 32 + RP, TB = kTrainCap, XE = (int)sizeof(xc_t);
    constexpr int CW = 32 / S;                 // channels per warp
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int xpad = (p.cap * p.CT * XE + 127) & ~127, dpad = (p.cap * p.CT * 4 + 127) & ~127,
              ppad = (p.cap * CP * 4 + 127) & ~127;
    const int stage_bytes = xpad + dpad + ppad;
    const int nwarps = blockDim.x >> 5;
    float* shist = reinterpret_cast<float*>(smem_raw + kBwdStages * stage_bytes);     // [nwarps][kHalf][16][32]
    float* sstp = shist + nwarps * kHalf * NS * 32;                                   // [nwarps][TB][3][32]
    uint64_t* full = reinterpret_cast<uint64_t*>(sstp + nwarps * TB * 3 * 32);
    uint64_t* empty = full + kBwdStages;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k = blockIdx.y, b = blockIdx.z, tile = blockIdx.x, c0 = tile * p.CT;
    const bool colview = (k & 1) != 0, rev = k >= 2;
    const int NB = colview ? p.NO_col * p.NI_col : p.NB_row;

    // iteration it handles the block with time-order index jb = NB-1-it
    auto issue = [&](int it) {
        const int s = it % kBwdStages;
        const int jb = NB - 1 - it;
        const int blk = rev ? NB - 1 - jb : jb;
        uint8_t* xs = smem_raw + s * stage_bytes;
        uint8_t* ds = xs + xpad;
        uint8_t* ps = ds + dpad;
        if (!colview) {
            mbar_expect_tx(&full[s], p.T_row * (p.CT * (XE + 4) + CP * 4));
            tma_load_3d(xs, &tmx_row, &full[s], c0, blk * p.T_row, b);
            tma_load_3d(ds, &tmd_row, &full[s], c0, blk * p.T_row, b);
            tma_load_4d(ps, &tmp_row, &full[s], 0, k, blk * p.T_row, b);
        } else {
            const int o = blk / p.NI_col, i = blk % p.NI_col;
            mbar_expect_tx(&full[s], p.nw * p.T_col * (p.CT * (XE + 4) + CP * 4));
            tma_load_4d(xs, &tmx_col, &full[s], c0, o * p.nw, i * p.T_col, b);
            tma_load_4d(ds, &tmd_col, &full[s], c0, o * p.nw, i * p.T_col, b);
            tma_load_5d(ps, &tmp_col, &full[s], 0, k, o * p.nw, i * p.T_col, b);
        }
    };

    if (tid == 0) {
        for (int s = 0; s < kBwdStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], nwarps); }
        mbar_fence_init();
        for (int it = 0; it < kBwdStages - 1 && it < NB; ++it) issue(it);
    }
    __syncthreads();
    int itn = kBwdStages - 1 < NB ? kBwdStages - 1 : NB;          // next block to load (thread 0)

    const int cl = tid / S, q = tid % S;      // channel inside the tile; which 16/S states this lane owns
    const int col = lane / S;                 // channel inside the warp
    const int c = c0 + cl;
    const bool cvalid = c < p.D;
    const int row = k * p.D + (cvalid ? c : 0);
    const int group = c0 / CW + warp;         // channel group of this warp: index of its dproj partial
    const bool gvalid = group * CW < p.D;

    float Ap[NS], gcar[NS], dA[NS], Wd[RP], dWd[RP];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        Ap[j] = (cvalid && q * NS + j < p.N) ? p.A[(int64_t)row * p.N + q * NS + j] * kLog2e : 0.f;
        gcar[j] = 0.f; dA[j] = 0.f;
    }
#pragma unroll
    for (int r = 0; r < RP; ++r) { Wd[r] = (cvalid && r < p.R) ? p.Wdt[(int64_t)row * p.R + r] : 0.f; dWd[r] = 0.f; }
    const float bias = cvalid ? p.bias[row] : 0.f;
    const float Dd = cvalid ? p.Ds[row] : 0.f;
    float dD_acc = 0.f, db_acc = 0.f;
    const int64_t gstride = 4 * (int64_t)p.D;
    float* dub = p.dudir + ((int64_t)b * p.L * 4 + k) * p.D + c;
    // value this lane owns after the butterflies: [0,16) dB_n, [16,32) dC_n with the state bits spread over the lane
    // bits from high to low (lane bit 4 = dB / dC, bit 3 = n bit 0, bit 2 = n bit 1, bit 1 = n bit 2, bit 0 = n bit 3;
    // with two lanes per channel bit 0 is the state half q instead)
    const int vown = ((lane >> 4) & 1) * 16 + ((lane >> 3) & 1) + 2 * ((lane >> 2) & 1) + 4 * ((lane >> 1) & 1) + 8 * (lane & 1);
    constexpr int RP2 = RP <= 4 ? 4 : (RP <= 8 ? 8 : (RP <= 16 ? 16 : 32));      // dt_r values padded to a power of two
    constexpr int RLG = RP2 == 4 ? 2 : (RP2 == 8 ? 3 : (RP2 == 16 ? 4 : 5));
    int rown = 0;                                                               // dt_r index this lane owns
#pragma unroll
    for (int i = 0; i < RLG; ++i) rown |= ((lane >> (4 - i)) & 1) << i;
    const bool rwriter = (lane & ((32 >> RLG) - 1)) == 0 && rown < RP;
    float* dpb = p.dproj + (((int64_t)(gvalid ? group : 0) * p.B + b) * p.L * 4 + k) * CP;
    float* hist = shist + warp * kHalf * NS * 32 + lane;        // [tl][j] at (tl * 16 + j) * 32
    float* stp = sstp + warp * TB * 3 * 32 + lane;              // per-step delta / d softplus / u of this lane
    const float* hck = p.hsave + (((int64_t)b * 4 + k) * p.NBmax * p.D + (cvalid ? c : 0)) * kMaxState + q * NS;   // + (jb-1) * D * 16

    for (int it = 0; it < NB; ++it) {
        const int s = it % kBwdStages, ph = (it / kBwdStages) & 1;
        // refill: thread 0 tests whether the slot's previous block was released by every warp and blocks only if the
        // block it is about to compute was never requested (see ss2d_core_fwd.cu: waiting here ties warp 0 to the
        // slowest warp of the CTA at every block)
        if (tid == 0) {
            while (itn < NB && itn - it < kBwdStages) {
                const int iprev = itn - kBwdStages;
                if (iprev >= 0) {
                    uint64_t* eb = &empty[itn % kBwdStages];
                    const uint32_t par = (iprev / kBwdStages) & 1;
                    if (!mbar_test_wait(eb, par)) {
                        if (itn > it) break;
                        mbar_wait(eb, par);
                    }
                }
                issue(itn);
                ++itn;
            }
        }
        __syncwarp();
        const int jb = NB - 1 - it;
        const int blk = rev ? NB - 1 - jb : jb;
        int nrows, ncols, nwbox, psh, pbase;
        if (!colview) {
            pbase = blk * p.T_row; nrows = min(p.T_row, p.L - pbase); ncols = 1; nwbox = 1; psh = 1;
        } else {
            const int o = blk / p.NI_col, i = blk % p.NI_col;
            const int w0 = o * p.nw, h0i = i * p.T_col;
            nrows = min(p.T_col, p.H - h0i); ncols = min(p.nw, p.W - w0); nwbox = p.nw; psh = p.W;
            pbase = h0i * p.W + w0;
        }
        const int nsteps = nrows * ncols;          // <= TB
        const bool single_col = nwbox == 1;
        int slot_l = 0, pos_l = 0;
        if (lane < nsteps) {
            const int ww = lane / nrows, hh = lane - ww * nrows;
            slot_l = hh * nwbox + ww;
            pos_l = pbase + hh * psh + ww;
        }
        const xc_t* xs = reinterpret_cast<const xc_t*>(smem_raw + s * stage_bytes) + cl;
        const float* dys = reinterpret_cast<const float*>(smem_raw + s * stage_bytes + xpad) + cl;
        const float* ps = reinterpret_cast<const float*>(smem_raw + s * stage_bytes + xpad + dpad);
        mbar_wait(&full[s], ph);

        auto slot_of = [&](const int tl) {
            const int ti = tl < nsteps ? (rev ? nsteps - 1 - tl : tl) : 0;
            return single_col ? ti : __shfl_sync(0xffffffffu, slot_l, ti);
        };
        auto pos_of = [&](const int tl) {
            const int ti = tl < nsteps ? (rev ? nsteps - 1 - tl : tl) : 0;
            return single_col ? pbase + ti * psh : __shfl_sync(0xffffffffu, pos_l, ti);
        };
        // state at the start of the block: the forward's checkpoint after block jb-1
        auto load_checkpoint = [&](float (&h)[NS]) {
            if (jb > 0 && cvalid) {
                const float4* hp4 = reinterpret_cast<const float4*>(hck + (int64_t)(jb - 1) * p.D * kMaxState);
#pragma unroll
                for (int j4 = 0; j4 < NS / 4; ++j4) {
                    const float4 v = __ldg(hp4 + j4);
                    h[4 * j4] = v.x; h[4 * j4 + 1] = v.y; h[4 * j4 + 2] = v.z; h[4 * j4 + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < NS; ++j) h[j] = 0.f;
            }
        };
        // per-step scalars of the block -> the warp's step table: delta, d softplus, u
        for (int tl = 0; tl < TB; ++tl) {
            const bool ok = tl < nsteps;
            const int sl = slot_of(tl);
            const float4* dtp = reinterpret_cast<const float4*>(ps + sl * CP + 32);
            float acc0 = bias, acc1 = 0.f;
#pragma unroll
            for (int r4 = 0; r4 < RP / 4; ++r4) {
                const float4 v = dtp[r4];
                fma2(acc0, acc1, Wd[4 * r4 + 0], Wd[4 * r4 + 1], v.x, v.y, acc0, acc1);
                fma2(acc0, acc1, Wd[4 * r4 + 2], Wd[4 * r4 + 3], v.z, v.w, acc0, acc1);
            }
            const float raw = acc0 + acc1;
            float sp, sg;
            softplus_sigmoid_f(raw, sp, sg);
            stp[(tl * 3 + 0) * 32] = ok ? sp : 0.f;
            stp[(tl * 3 + 1) * 32] = ok ? sg : 0.f;
            stp[(tl * 3 + 2) * 32] = ok ? to_f<xc_t>(xs[sl * p.CT]) : 0.f;
        }
        float h[NS];
        // one forward step of all 16 states; `park` stores h_{t-1} into history slab `hslot`
        auto advance = [&](float (&h)[NS], const int tl, const bool park, const int hslot) {
            const float4* bp = reinterpret_cast<const float4*>(ps + slot_of(tl) * CP) + q * (NS / 4);
            const float dli = stp[(tl * 3 + 0) * 32];
            const float dlu = dli * stp[(tl * 3 + 2) * 32];
#pragma unroll
            for (int j4 = 0; j4 < NS / 4; ++j4) {
                const float4 bv = bp[j4];
                const int j = 4 * j4;
                if (park) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) hist[(hslot * NS + j + e) * 32] = h[j + e];
                }
                float x0, x1, x2, x3, w0, w1, w2, w3;
                mul2(x0, x1, dli, dli, Ap[j], Ap[j + 1]);
                mul2(x2, x3, dli, dli, Ap[j + 2], Ap[j + 3]);
                mul2(w0, w1, dlu, dlu, bv.x, bv.y);
                mul2(w2, w3, dlu, dlu, bv.z, bv.w);
                const float a0 = ex2_approx(x0), a1 = ex2_approx(x1), a2 = ex2_approx(x2), a3 = ex2_approx(x3);
                fma2(h[j], h[j + 1], a0, a1, h[j], h[j + 1], w0, w1);
                fma2(h[j + 2], h[j + 3], a2, a3, h[j + 2], h[j + 3], w2, w3);
            }
        };
        // reverse step tl, history slab hslot
        // FULLB: the block has all TB steps, so no step is masked (the selects on `ok` fold away)
        auto reverse = [&](auto full_tag, const int tl, const int hslot) {
            const bool ok = decltype(full_tag)::value || tl < nsteps;
            const int sl = slot_of(tl);
            const float4* bp = reinterpret_cast<const float4*>(ps + sl * CP) + q * (NS / 4);
            const float dy = ok ? dys[sl * p.CT] : 0.f;
            const float dli = stp[(tl * 3 + 0) * 32], sgi = stp[(tl * 3 + 1) * 32], u = stp[(tl * 3 + 2) * 32];
            const float dlu = dli * u;
            float adu[2] = {0.f, 0.f}, adA[2] = {0.f, 0.f};
            float yg[NS / 4];                     // per group of 4 states: the butterfly's partial after three stages
#pragma unroll
            for (int j4 = 0; j4 < NS / 4; ++j4) {
                const float4 bv = bp[j4], cv = bp[4 + j4];
                const float bb[4] = {bv.x, bv.y, bv.z, bv.w}, cc[4] = {cv.x, cv.y, cv.z, cv.w};
                const int j = 4 * j4;
                float hp[4], a[4], x[4], xp[2];
                mul2(x[0], x[1], dli, dli, Ap[j], Ap[j + 1]);
                mul2(x[2], x[3], dli, dli, Ap[j + 2], Ap[j + 3]);
#pragma unroll
                for (int e = 0; e < 4; ++e) { hp[e] = hist[(hslot * NS + j + e) * 32]; a[e] = ex2_approx(x[e]); }
#pragma unroll
                for (int e = 0; e < 4; e += 2) {
                    // h_t is not recomputed: it is the h_{t-1} the previous reverse iteration (step t+1) loaded, or the
                    // state the forward re-run ended in (h); a g h_{t-1} is formed from the carried a g
                    float gt0, gt1, vb0, vb1, vc0, vc1, t0, t1, ga0, ga1;
                    fma2(gt0, gt1, dy, dy, cc[e], cc[e + 1], gcar[j + e], gcar[j + e + 1]);    // g_t
                    mul2(ga0, ga1, a[e], a[e + 1], gt0, gt1);                                  // a_t g_t: carried to step t-1
                    mul2(vb0, vb1, gt0, gt1, dlu, dlu);                                        // dB_n of this channel
                    mul2(vc0, vc1, dy, dy, h[j + e], h[j + e + 1]);                          // dC_n of this channel
                    fma2(adu[0], adu[1], gt0, gt1, bb[e], bb[e + 1], adu[0], adu[1]);          // sum_n g B
                    mul2(t0, t1, ga0, ga1, hp[e], hp[e + 1]);                                  // g h_{t-1} a
                    fma2(adA[0], adA[1], t0, t1, Ap[j + e], Ap[j + e + 1], adA[0], adA[1]);
                    fma2(dA[j + e], dA[j + e + 1], t0, t1, dli, dli, dA[j + e], dA[j + e + 1]);
                    h[j + e] = hp[e]; h[j + e + 1] = hp[e + 1];
                    gcar[j + e] = ok ? ga0 : gcar[j + e];
                    gcar[j + e + 1] = ok ? ga1 : gcar[j + e + 1];
                    // butterfly stages 1 (dB_n | dC_n over lane bit 4) and 2 (the two states of the pair over bit 3)
                    const float u0 = bfly(vb0, vc0, (lane & 16) != 0, 16), u1 = bfly(vb1, vc1, (lane & 16) != 0, 16);
                    xp[e >> 1] = bfly(u0, u1, (lane & 8) != 0, 8);
                }
                yg[j4] = bfly(xp[0], xp[1], (lane & 4) != 0, 4);                               // stage 3: state bit 1
            }
            // stages 4 and 5: state bits 2 and 3 (two lanes per channel: bit 3 is the lane's own half q, no exchange)
            float vsum;
            if constexpr (S == 1) {
                const float z0 = bfly(yg[0], yg[1], (lane & 2) != 0, 2), z1 = bfly(yg[2], yg[3], (lane & 2) != 0, 2);
                vsum = bfly(z0, z1, (lane & 1) != 0, 1);
            } else {
                vsum = bfly(yg[0], yg[1], (lane & 2) != 0, 2);
            }
            float su = adu[0] + adu[1], sa = adA[0] + adA[1];
#pragma unroll
            for (int off = S / 2; off > 0; off >>= 1) {      // the S lanes of a channel hold disjoint states
                su += __shfl_xor_sync(0xffffffffu, su, off);
                sa += __shfl_xor_sync(0xffffffffu, sa, off);
            }
            // d delta = sum_n g (B u + h_{t-1} a A_n);  Ap = A log2(e), so the A part is scaled back by ln 2
            const float ddl = fmaf(su, u, sa * kLn2);
            const float ddr = ddl * sgi;                                  // 0 on masked steps (sg = 0)
            const int pos = pos_of(tl);
            if (ok && cvalid && q == 0) {
                dub[pos * gstride] = fmaf(Dd, dy, dli * su);
                dD_acc = fmaf(dy, u, dD_acc);
            }
            db_acc += ddr;
            // d dt_r = sum over the channels of Wdt[c][r] * d delta_raw[c]: the same butterfly over the RP2 padded values
            // (adjacent r first), then plain exchanges over the lane bits that are left
            const float ddq = q == 0 ? ddr : 0.f;                          // one contribution per channel
            const float4* dtp = reinterpret_cast<const float4*>(ps + sl * CP + 32);
            float tr[RP2];
#pragma unroll
            for (int r4 = 0; r4 < RP / 4; ++r4) {
                const float4 v = dtp[r4];
                const float dt4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    dWd[4 * r4 + e] = fmaf(ddr, dt4[e], dWd[4 * r4 + e]);
                    tr[4 * r4 + e] = Wd[4 * r4 + e] * ddq;
                }
            }
#pragma unroll
            for (int r = RP; r < RP2; ++r) tr[r] = 0.f;
#pragma unroll
            for (int lv = 0; lv < RLG; ++lv) {
                const int mask = 16 >> lv;
#pragma unroll
                for (int i = 0; i < (RP2 >> (lv + 1)); ++i) tr[i] = bfly(tr[2 * i], tr[2 * i + 1], (lane & mask) != 0, mask);
            }
            float rsum = tr[0];
#pragma unroll
            for (int mask = 16 >> RLG; mask > 0; mask >>= 1) rsum += __shfl_xor_sync(0xffffffffu, rsum, mask);
            if (ok && gvalid) {
                float* dpr = dpb + (int64_t)pos * 4 * CP;
                if (S == 1) dpr[vown] = vsum;
                else dpr[((lane >> 4) & 1) * 16 + q * NS + ((lane >> 3) & 1) + 2 * ((lane >> 2) & 1) + 4 * ((lane >> 1) & 1)] = vsum;
                if (rwriter) dpr[32 + rown] = rsum;
            }
        };

        __syncwarp();
        if (nsteps > kHalf) {
            // second half first: pass 1 over the first half, then phase A / B on steps kHalf .. TB-1
            load_checkpoint(h);
            for (int tl = 0; tl < kHalf; ++tl) advance(h, tl, false, 0);
            for (int tl = kHalf; tl < TB; ++tl) advance(h, tl, true, tl - kHalf);
            __syncwarp();
            if (nsteps == TB) { for (int tl = TB - 1; tl >= kHalf; --tl) reverse(std::true_type{}, tl, tl - kHalf); }
            else { for (int tl = TB - 1; tl >= kHalf; --tl) reverse(std::false_type{}, tl, tl - kHalf); }
        }
        // first half
        load_checkpoint(h);
        for (int tl = 0; tl < kHalf; ++tl) advance(h, tl, true, tl);
        __syncwarp();
        if (nsteps >= kHalf) { for (int tl = kHalf - 1; tl >= 0; --tl) reverse(std::true_type{}, tl, tl); }
        else { for (int tl = kHalf - 1; tl >= 0; --tl) reverse(std::false_type{}, tl, tl); }

        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
    }
    if (cvalid) {
#pragma unroll
        for (int j = 0; j < NS; ++j)
            if (q * NS + j < p.N) p.dA_part[((int64_t)b * 4 * p.D + row) * p.N + q * NS + j] = dA[j];
        if (q == 0) {
#pragma unroll
            for (int r = 0; r < RP; ++r) p.dW_part[((int64_t)b * 4 * p.D + row) * RP + r] = dWd[r];
            p.dD_part[(int64_t)b * 4 * p.D + row] = dD_acc;
            p.db_part[(int64_t)b * 4 * p.D + row] = db_acc;
        }
    }
}

// lanes per channel: one (the fewest instructions; four CTAs of three warps per SM) as soon as the launch fills most of
// one round of resident CTAs, two for smaller launches (thin warps hide the latency of the reverse step better)
static int core_bwd_split(int B, int D) {
    if (const char* e = getenv("MMB_BWD_S")) { const int v = atoi(e); if (v == 1 || v == 2) return v; }
    const long ctas1 = 4L * B * ((D + 95) / 96);
    return ctas1 * 10 >= 6L * 4 * num_sms() ? 1 : 2;
}
// channels per CTA: one warp per 32/S-channel group, up to three warps, chosen to divide the group count
static int core_bwd_ct(int D, int S) {
    const int cw = 32 / S;
    const int groups = (D + cw - 1) / cw;
    const int per = groups % 3 == 0 ? 3 : (groups % 2 == 0 ? 2 : (groups <= 3 ? groups : 1));
    return per * cw;
}

template <int S, int RP, typename xc_t>
static int launch_core_bwd(CoreBwdParams& p, const void* xc, const float* dY, const float* proj, cudaStream_t st) {
    constexpr int CP = 32 + RP;
    constexpr uint64_t XE = sizeof(xc_t);
    const CUtensorMapDataType xdt = XE == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUtensorMap tmx_row, tmx_col, tmd_row, tmd_col, tmp_row, tmp_col;
    const uint64_t B = p.B, H = p.H, W = p.W, L = p.L, D = p.D;
    {
        const uint64_t dims[3] = {D, L, B};
        const uint32_t box[3] = {(uint32_t)p.CT, (uint32_t)p.T_row, 1};
        const uint64_t sx[2] = {D * XE, L * D * XE}, sd[2] = {D * 4, L * D * 4};
        if (!make_tmap(&tmx_row, xdt, 3, xc, dims, sx, box)) return MMB_ERR_UNSUPPORTED;
        if (!make_tmap(&tmd_row, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, dY, dims, sd, box)) return MMB_ERR_UNSUPPORTED;
    }
    {
        const uint64_t dims[4] = {D, W, H, B};
        const uint32_t box[4] = {(uint32_t)p.CT, (uint32_t)p.nw, (uint32_t)p.T_col, 1};
        const uint64_t sx[3] = {D * XE, W * D * XE, L * D * XE}, sd[3] = {D * 4, W * D * 4, L * D * 4};
        if (!make_tmap(&tmx_col, xdt, 4, xc, dims, sx, box)) return MMB_ERR_UNSUPPORTED;
        if (!make_tmap(&tmd_col, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dY, dims, sd, box)) return MMB_ERR_UNSUPPORTED;
    }
    {
        const uint64_t dims[4] = {CP, 4, L, B}, str[3] = {CP * 4, 4 * CP * 4, L * 4 * CP * 4};
        const uint32_t box[4] = {CP, 1, (uint32_t)p.T_row, 1};
        if (!make_tmap(&tmp_row, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, proj, dims, str, box)) return MMB_ERR_UNSUPPORTED;
    }
    {
        const uint64_t dims[5] = {CP, 4, W, H, B}, str[4] = {CP * 4, 4 * CP * 4, W * 4 * CP * 4, L * 4 * CP * 4};
        const uint32_t box[5] = {CP, 1, (uint32_t)p.nw, (uint32_t)p.T_col, 1};
        if (!make_tmap(&tmp_col, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, proj, dims, str, box)) return MMB_ERR_UNSUPPORTED;
    }
    const int threads = p.CT * S, nwarps = threads / 32;
    const size_t xpad = ((size_t)p.cap * p.CT * XE + 127) & ~(size_t)127, dpad = ((size_t)p.cap * p.CT * 4 + 127) & ~(size_t)127,
                 ppad = ((size_t)p.cap * CP * 4 + 127) & ~(size_t)127;
    const size_t smem = kBwdStages * (xpad + dpad + ppad) + (size_t)nwarps * (kHalf * (kMaxState / S) * 32 + kTrainCap * 3 * 32) * 4 +
                        2 * kBwdStages * sizeof(uint64_t);
    auto kern = ss2d_core_bwd_kernel<S, RP, xc_t>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    dim3 grid(p.tiles, 4, p.B);
    kern<<<grid, threads, smem, st>>>(tmx_row, tmx_col, tmd_row, tmd_col, tmp_row, tmp_col, p);
    return launch_status();
}

template <typename xc_t>
static int dispatch_core_bwd(int S, int dt_pad, CoreBwdParams& p, const void* xc, const float* dY, const float* proj, cudaStream_t st) {
#define MMB_BWD_CASE(RPV)                                                                       \
    case RPV: return S == 1 ? launch_core_bwd<1, RPV, xc_t>(p, xc, dY, proj, st)                    \
                            : launch_core_bwd<2, RPV, xc_t>(p, xc, dY, proj, st);
    switch (dt_pad) {
        MMB_BWD_CASE(4) MMB_BWD_CASE(8) MMB_BWD_CASE(12) MMB_BWD_CASE(16) MMB_BWD_CASE(24) MMB_BWD_CASE(32)
        default: return MMB_ERR_UNSUPPORTED;
    }
#undef MMB_BWD_CASE
}

}  // namespace mmb

extern "C" int mmb_ss2d_core_bwd_tiles(int batch, int D) {
    if (D <= 0 || batch < 0) return MMB_ERR_INVALID_ARG;
    const int cw = 32 / mmb::core_bwd_split(batch > 0 ? batch : 1, D);       // channels per warp = per dproj partial
    return (D + cw - 1) / cw;
}

extern "C" int mmb_ss2d_core_bwd(const void* xc, const float* proj, const float* dY, const float* Wdt,
                                 const float* dt_bias, const float* A, const float* Ds, const float* hsave,
                                 float* dudir, float* dproj_part, float* dA_part, float* dW_part, float* dD_part,
                                 float* db_part, int batch, int H, int W, int D, int dstate, int dt_rank, int dt_pad, int xc_dtype,
                                 void* stream) {
    using namespace mmb;
    if (!xc || !proj || !dY || !Wdt || !dt_bias || !A || !Ds || !hsave || !dudir || !dproj_part || !dA_part || !dW_part ||
        !dD_part || !db_part) return MMB_ERR_INVALID_ARG;
    if (batch < 0 || H <= 0 || W <= 0 || D <= 0 || dstate <= 0 || dt_rank <= 0) return MMB_ERR_INVALID_ARG;
    if (dstate > kMaxState || dt_pad != mmb_ss2d_core_dt_pad(dt_rank)) return MMB_ERR_UNSUPPORTED;
    if (xc_dtype != MMB_F32 && xc_dtype != MMB_BF16) return MMB_ERR_UNSUPPORTED;
    if (D % (xc_dtype == MMB_F32 ? 4 : 8) != 0 || batch > 65535) return MMB_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(xc) | reinterpret_cast<uintptr_t>(proj) | reinterpret_cast<uintptr_t>(dY) |
         reinterpret_cast<uintptr_t>(hsave)) % 16 != 0) return MMB_ERR_UNSUPPORTED;
    if (batch == 0) return MMB_OK;
    CoreGeom g;
    if (!core_geometry(H, W, kTrainCap, g)) return MMB_ERR_UNSUPPORTED;
    CoreBwdParams p;
    p.Wdt = Wdt; p.bias = dt_bias; p.A = A; p.Ds = Ds; p.hsave = hsave;
    p.dudir = dudir; p.dproj = dproj_part; p.dA_part = dA_part; p.dW_part = dW_part; p.dD_part = dD_part; p.db_part = db_part;
    p.B = batch; p.H = H; p.W = W; p.L = H * W; p.D = D; p.N = dstate; p.R = dt_rank;
    const int S = core_bwd_split(batch, D);
    p.CT = core_bwd_ct(D, S); p.tiles = (D + p.CT - 1) / p.CT; p.NBmax = g.nblocks_max();
    p.T_row = g.T_row; p.NB_row = g.NB_row; p.nw = g.nw; p.T_col = g.T_col; p.NI_col = g.NI_col; p.NO_col = g.NO_col;
    p.cap = g.cap;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (xc_dtype == MMB_F32) return dispatch_core_bwd<float>(S, dt_pad, p, xc, dY, proj, st);
    return dispatch_core_bwd<__nv_bfloat16>(S, dt_pad, p, xc, dY, proj, st);
}
