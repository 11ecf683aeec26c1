// Fused SS2D core, backward.  Gradient of mmb_ss2d_core_fwd w.r.t. xc (per direction), proj (dB_n, dC_n,
// d dt_r), Wdt, dt_bias, A and D.  Formulas: SURVEY.md Appendix B; the cross-scan / cross-merge index maps are
// the forward's (Appendix A) -- the backward of a gather by index is a scatter to the same index,
// so "cross-merge of the per-direction du" is again just the store address.
//
// Same ownership and TMA ring as the forward, with the sequence walked BACKWARDS in the blocks of
// kTrainCap = 8 steps whose end states the forward checkpointed (hsave).  Four lanes per channel,
// four states per lane, so that a whole block of history lives in REGISTERS:
//   phase A  from the checkpoint, re-run the 8 steps forward keeping h_{t-1} and a_t of every step
//            (64 registers) -- the only exps of the backward;
//   phase B  the reverse recurrence g_t = dy_t C_t + a_{t+1} g_{t+1} and all gradients, no exp.
// dB_n / dC_n need a sum over the channels of the (batch, direction): an 8-lane transposing shuffle
// reduction inside each warp, per-warp tiles in shared memory added in fixed order, one partial per
// channel tile in HBM.  dA / dD are per-thread accumulators, one partial per batch element.
// No float atomics anywhere: results are bit-reproducible.
#include <type_traits>

#include "common.cuh"
#include "core_geom.cuh"
#include "tma.cuh"

namespace mmb {

constexpr int kBwdStages = 3;

struct CoreBwdParams {
    const float* Wdt; const float* bias; const float* A; const float* Ds; const float* hsave;
    float* dudir;                   // (B, L, 4, D)
    float* dproj;                   // (tiles, B, L, 4, CP): [dB_n | dC_n | d dt_r]
    float* dA_part;                 // (B, 4D, N)
    float* dW_part;                 // (B, 4D, RP)
    float* dD_part; float* db_part; // (B, 4D)
    int B, H, W, L, D, N, R, CT, tiles, NBmax;
    int T_row, NB_row, nw, T_col, NI_col, NO_col, cap;
};

template <int RP, typename xc_t>
__global__ void __launch_bounds__(192, 2)
ss2d_core_bwd_kernel(const __grid_constant__ CUtensorMap tmx_row, const __grid_constant__ CUtensorMap tmx_col,
                     const __grid_constant__ CUtensorMap tmd_row, const __grid_constant__ CUtensorMap tmd_col,
                     const __grid_constant__ CUtensorMap tmp_row, const __grid_constant__ CUtensorMap tmp_col,
                     const CoreBwdParams p) {
    constexpr int S = 4, NS = 4, CP = 32 + RP, TB = kTrainCap, XE = (int)sizeof(xc_t);
    extern __shared__ __align__(128) uint8_t smem_raw[];
    constexpr int RW = CP;                      // floats per step in the cross-channel reduction tiles
    const int xpad = (p.cap * p.CT * XE + 127) & ~127, dpad = (p.cap * p.CT * 4 + 127) & ~127,
              ppad = (p.cap * CP * 4 + 127) & ~127, hpad = p.CT * kMaxState * 4;
    const int stage_bytes = xpad + dpad + ppad + hpad;
    const int nwarps = blockDim.x >> 5;
    float* swred = reinterpret_cast<float*>(smem_raw + kBwdStages * stage_bytes);     // [2][nwarps][TB][RW]
    int* spos = reinterpret_cast<int*>(swred + 2 * nwarps * TB * RW);                 // [2][TB]
    uint64_t* full = reinterpret_cast<uint64_t*>(spos + 2 * TB);
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
        uint8_t* hs = ps + ppad;
        // checkpoint after block jb-1 = state at the start of block jb: CT x 16 floats, contiguous
        const int hbytes = jb > 0 ? min(p.CT, p.D - c0) * kMaxState * 4 : 0;
        if (!colview) {
            mbar_expect_tx(&full[s], p.T_row * (p.CT * (XE + 4) + CP * 4) + hbytes);
            tma_load_3d(xs, &tmx_row, &full[s], c0, blk * p.T_row, b);
            tma_load_3d(ds, &tmd_row, &full[s], c0, blk * p.T_row, b);
            tma_load_4d(ps, &tmp_row, &full[s], 0, k, blk * p.T_row, b);
        } else {
            const int o = blk / p.NI_col, i = blk % p.NI_col;
            mbar_expect_tx(&full[s], p.nw * p.T_col * (p.CT * (XE + 4) + CP * 4) + hbytes);
            tma_load_4d(xs, &tmx_col, &full[s], c0, o * p.nw, i * p.T_col, b);
            tma_load_4d(ds, &tmd_col, &full[s], c0, o * p.nw, i * p.T_col, b);
            tma_load_5d(ps, &tmp_col, &full[s], 0, k, o * p.nw, i * p.T_col, b);
        }
        if (hbytes)
            bulk_load_1d(hs, p.hsave + ((((int64_t)b * 4 + k) * p.NBmax + (jb - 1)) * p.D + c0) * kMaxState, hbytes, &full[s]);
    };

    if (tid == 0) {
        for (int s = 0; s < kBwdStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], nwarps); }
        mbar_fence_init();
        for (int it = 0; it < kBwdStages - 1 && it < NB; ++it) issue(it);
    }
    __syncthreads();

    const int cl = tid / S, q = tid % S;
    const int c = c0 + cl;
    const bool cvalid = c < p.D;
    const int row = k * p.D + (cvalid ? c : 0);
    const int lane_base = lane & ~(S - 1);

    float Ap[NS], Araw[NS], gcar[NS], dA[NS], Wd[RP], Wq[RP / 4], dWq[RP / 4];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        const int n = 4 * q + j;
        Araw[j] = (cvalid && n < p.N) ? p.A[(int64_t)row * p.N + n] : 0.f;
        Ap[j] = Araw[j] * kLog2e;
        gcar[j] = 0.f; dA[j] = 0.f;
    }
#pragma unroll
    for (int r = 0; r < RP; ++r) Wd[r] = (cvalid && r < p.R) ? p.Wdt[(int64_t)row * p.R + r] : 0.f;
    // lane q owns the dt ranks r = q, q+4, ... of its channel for dWdt and d dt_r
#pragma unroll
    for (int i = 0; i < RP / 4; ++i) {
        Wq[i] = q == 0 ? Wd[4 * i] : q == 1 ? Wd[4 * i + 1] : q == 2 ? Wd[4 * i + 2] : Wd[4 * i + 3];
        dWq[i] = 0.f;
    }
    const float bias = cvalid ? p.bias[row] : 0.f;
    const float Dd = cvalid ? p.Ds[row] : 0.f;
    float dD_acc = 0.f, db_acc = 0.f;
    const int64_t gstride = 4 * (int64_t)p.D;
    float* dub = p.dudir + ((int64_t)b * p.L * 4 + k) * p.D + c;

    for (int it = 0; it < NB; ++it) {
        const int s = it % kBwdStages, ph = (it / kBwdStages) & 1;
        if (tid == 0 && it + kBwdStages - 1 < NB) {
            const int itn = it + kBwdStages - 1;
            if (it > 0) mbar_wait(&empty[itn % kBwdStages], ((it - 1) / kBwdStages) & 1);
            issue(itn);
        }
        __syncwarp();
        const int jb = NB - 1 - it;
        const int blk = rev ? NB - 1 - jb : jb;
        int nrows, ncols, nwbox, psh, pbase;
        if (!colview) {
            pbase = blk * p.T_row; nrows = min(p.T_row, p.L - pbase); ncols = 1; nwbox = 1; psh = 1;
        } else {
            const int o = blk / p.NI_col, i = blk % p.NI_col;
            const int w0 = o * p.nw, h0 = i * p.T_col;
            nrows = min(p.T_col, p.H - h0); ncols = min(p.nw, p.W - w0); nwbox = p.nw; psh = p.W;
            pbase = h0 * p.W + w0;
        }
        const int nsteps = nrows * ncols;          // <= TB
        const bool single_col = nwbox == 1;        // slot / position affine in the step index
        int slot_l = 0, pos_l = 0;
        if (lane < nsteps) {
            const int ww = lane / nrows, hh = lane - ww * nrows;
            slot_l = hh * nwbox + ww;
            pos_l = pbase + hh * psh + ww;
        }
        const xc_t* xs = reinterpret_cast<const xc_t*>(smem_raw + s * stage_bytes) + cl;
        const float* dys = reinterpret_cast<const float*>(smem_raw + s * stage_bytes + xpad) + cl;
        const float* ps = reinterpret_cast<const float*>(smem_raw + s * stage_bytes + xpad + dpad);
        const float* hs = reinterpret_cast<const float*>(smem_raw + s * stage_bytes + xpad + dpad + ppad);
        float* wred = swred + (it & 1) * nwarps * TB * RW + warp * TB * RW;
        if (warp == 0 && lane < TB) {
            const int ti = lane < nsteps ? (rev ? nsteps - 1 - lane : lane) : 0;
            spos[(it & 1) * TB + lane] = __shfl_sync(0xffu, pos_l, ti);
        }
        mbar_wait(&full[s], ph);
        // state at the start of the block: the forward's checkpoint after block jb-1 (staged with the tiles)
        float h[NS];
        if (jb > 0 && cvalid) {
            const float4 v = *reinterpret_cast<const float4*>(hs + cl * kMaxState + 4 * q);
            h[0] = v.x; h[1] = v.y; h[2] = v.z; h[3] = v.w;
        } else {
            h[0] = h[1] = h[2] = h[3] = 0.f;
        }

        // ---- phase A: forward through the block (time order), history in registers ---------------------
        int slot[TB];
        float dl[TB], sg[TB], hist[TB][NS], aa[TB][NS];
        {
            float own_dl[2], own_sg[2];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const int tl = q + S * m;
                const bool okk = tl < nsteps;
                const int ti = okk ? (rev ? nsteps - 1 - tl : tl) : 0;
                const int sl = single_col ? ti : __shfl_sync(0xffffffffu, slot_l, ti);
                const float4* dtp = reinterpret_cast<const float4*>(ps + sl * CP + 32);
                float acc0 = bias, acc1 = 0.f;
#pragma unroll
                for (int r4 = 0; r4 < RP / 4; ++r4) {
                    const float4 v = dtp[r4];
                    fma2(acc0, acc1, Wd[4 * r4 + 0], Wd[4 * r4 + 1], v.x, v.y, acc0, acc1);
                    fma2(acc0, acc1, Wd[4 * r4 + 2], Wd[4 * r4 + 3], v.z, v.w, acc0, acc1);
                }
                const float raw = acc0 + acc1;
                own_dl[m] = okk ? softplus_f(raw) : 0.f;
                own_sg[m] = okk ? (raw > 20.f ? 1.f : sigmoid_f(raw)) : 0.f;
            }
#pragma unroll
            for (int tl = 0; tl < TB; ++tl) {
                dl[tl] = __shfl_sync(0xffffffffu, own_dl[tl / S], lane_base + (tl % S));
                sg[tl] = __shfl_sync(0xffffffffu, own_sg[tl / S], lane_base + (tl % S));
                const bool ok = tl < nsteps;
                const int ti = ok ? (rev ? nsteps - 1 - tl : tl) : 0;
                slot[tl] = single_col ? ti : __shfl_sync(0xffffffffu, slot_l, ti);
            }
        }
#pragma unroll
        for (int tl = 0; tl < TB; ++tl) {
            const float uu = tl < nsteps ? to_f<xc_t>(xs[slot[tl] * p.CT]) : 0.f;
            const float dlu = dl[tl] * uu;
            const float4 bv = reinterpret_cast<const float4*>(ps + slot[tl] * CP)[q];
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
            float x0, x1, x2, x3, w0, w1, w2, w3;
            mul2(x0, x1, dl[tl], dl[tl], Ap[0], Ap[1]);
            mul2(x2, x3, dl[tl], dl[tl], Ap[2], Ap[3]);
            mul2(w0, w1, dlu, dlu, bb[0], bb[1]);
            mul2(w2, w3, dlu, dlu, bb[2], bb[3]);
#pragma unroll
            for (int j = 0; j < NS; ++j) hist[tl][j] = h[j];
            aa[tl][0] = ex2_approx(x0); aa[tl][1] = ex2_approx(x1); aa[tl][2] = ex2_approx(x2); aa[tl][3] = ex2_approx(x3);
            fma2(h[0], h[1], aa[tl][0], aa[tl][1], h[0], h[1], w0, w1);
            fma2(h[2], h[3], aa[tl][2], aa[tl][3], h[2], h[3], w2, w3);
        }
        // ---- phase B: reverse recurrence --------------------------------------------------------------------
#pragma unroll
        for (int tl = TB - 1; tl >= 0; --tl) {
            const bool ok = tl < nsteps;
            const float uu = ok ? to_f<xc_t>(xs[slot[tl] * p.CT]) : 0.f;
            const float dy = ok ? dys[slot[tl] * p.CT] : 0.f;
            const float dlu = dl[tl] * uu;
            const float4 bv = reinterpret_cast<const float4*>(ps + slot[tl] * CP)[q];
            const float4 cv = reinterpret_cast<const float4*>(ps + slot[tl] * CP)[4 + q];
            const float bb[4] = {bv.x, bv.y, bv.z, bv.w}, cc[4] = {cv.x, cv.y, cv.z, cv.w};
            float adu, adl, v[8];
            {
                // the four states of the lane as two packed pairs (FFMA2 / FMUL2)
                float w[4], ht[4], gt[4], hpa[4], t0[4], t1[4], ga[4], acc_u[2] = {0.f, 0.f}, acc_l[2] = {0.f, 0.f};
#pragma unroll
                for (int j = 0; j < NS; j += 2) {
                    mul2(w[j], w[j + 1], dlu, dlu, bb[j], bb[j + 1]);
                    fma2(ht[j], ht[j + 1], aa[tl][j], aa[tl][j + 1], hist[tl][j], hist[tl][j + 1], w[j], w[j + 1]);
                    fma2(gt[j], gt[j + 1], dy, dy, cc[j], cc[j + 1], gcar[j], gcar[j + 1]);
                    mul2(hpa[j], hpa[j + 1], hist[tl][j], hist[tl][j + 1], aa[tl][j], aa[tl][j + 1]);
                    mul2(v[j], v[j + 1], gt[j], gt[j + 1], dlu, dlu);
                    mul2(v[NS + j], v[NS + j + 1], dy, dy, ht[j], ht[j + 1]);
                    fma2(acc_u[0], acc_u[1], gt[j], gt[j + 1], bb[j], bb[j + 1], acc_u[0], acc_u[1]);
                    mul2(t0[j], t0[j + 1], hpa[j], hpa[j + 1], Araw[j], Araw[j + 1]);
                    fma2(t1[j], t1[j + 1], bb[j], bb[j + 1], uu, uu, t0[j], t0[j + 1]);
                    fma2(acc_l[0], acc_l[1], gt[j], gt[j + 1], t1[j], t1[j + 1], acc_l[0], acc_l[1]);
                    mul2(t0[j], t0[j + 1], gt[j], gt[j + 1], hpa[j], hpa[j + 1]);
                    fma2(dA[j], dA[j + 1], t0[j], t0[j + 1], dl[tl], dl[tl], dA[j], dA[j + 1]);
                    mul2(ga[j], ga[j + 1], aa[tl][j], aa[tl][j + 1], gt[j], gt[j + 1]);
                    gcar[j] = ok ? ga[j] : gcar[j];
                    gcar[j + 1] = ok ? ga[j + 1] : gcar[j + 1];
                }
                adu = acc_u[0] + acc_u[1];
                adl = acc_l[0] + acc_l[1];
            }
            adu += __shfl_xor_sync(0xffffffffu, adu, 1); adu += __shfl_xor_sync(0xffffffffu, adu, 2);
            adl += __shfl_xor_sync(0xffffffffu, adl, 1); adl += __shfl_xor_sync(0xffffffffu, adl, 2);
            const int tip = ok ? (rev ? nsteps - 1 - tl : tl) : 0;
            const int pos = single_col ? pbase + tip * psh : __shfl_sync(0xffffffffu, pos_l, tip);
            const float ddr = adl * sg[tl];                  // d(Wdt.dt_r + bias); sg = 0 on masked steps
            if (ok && cvalid && q == 0) {
                dub[pos * gstride] = fmaf(Dd, dy, dl[tl] * adu);
                dD_acc = fmaf(dy, uu, dD_acc);
            }
            db_acc += ddr;
            // dWdt[c][r] += ddr * dt_r and d dt_r += Wdt[c][r] * ddr (summed over the warp's 8 channels)
#pragma unroll
            for (int i = 0; i < RP / 4; ++i) {
                const float dtv = ps[slot[tl] * CP + 32 + q + 4 * i];
                dWq[i] = fmaf(ddr, dtv, dWq[i]);
                float w = Wq[i] * ddr;
                w += __shfl_xor_sync(0xffffffffu, w, 4);
                w += __shfl_xor_sync(0xffffffffu, w, 8);
                w += __shfl_xor_sync(0xffffffffu, w, 16);
                if (lane < 4) wred[tl * RW + 32 + q + 4 * i] = w;
            }
            // sum v[0..7] over the 8 channels of this warp (lanes with equal q)
#pragma unroll
            for (int half = 4, off = 16; half >= 1; half >>= 1, off >>= 1) {
                const bool hi = (lane & off) != 0;
#pragma unroll
                for (int i = 0; i < half; ++i) {
                    const float send = hi ? v[i] : v[i + half];
                    const float keep = hi ? v[i + half] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            {
                const int idx = ((lane & 16) ? 4 : 0) + ((lane & 8) ? 2 : 0) + ((lane & 4) ? 1 : 0);
                const int n = 4 * q + (idx & 3);
                wred[tl * RW + (idx < 4 ? n : 16 + n)] = v[0];
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);     // the stage's tiles are no longer needed
        __syncthreads();
        // ---- this channel tile's partial dB / dC of the block, warps added in fixed order ---------------
        {
            const float* wr = swred + (it & 1) * nwarps * TB * RW;
            float* out = p.dproj + (((int64_t)tile * p.B + b) * p.L * 4 + k) * RW;
            for (int idx = tid; idx < nsteps * RW; idx += blockDim.x) {
                const int tl = idx / RW, vv = idx - tl * RW;
                float sum = 0.f;
                for (int w = 0; w < nwarps; ++w) sum += wr[(w * TB + tl) * RW + vv];
                out[(int64_t)spos[(it & 1) * TB + tl] * 4 * RW + vv] = sum;
            }
        }
    }
    if (cvalid) {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            const int n = 4 * q + j;
            if (n < p.N) p.dA_part[((int64_t)b * 4 * p.D + row) * p.N + n] = dA[j];
        }
#pragma unroll
        for (int i = 0; i < RP / 4; ++i) p.dW_part[((int64_t)b * 4 * p.D + row) * RP + q + 4 * i] = dWq[i];
        if (q == 0) {
            p.dD_part[(int64_t)b * 4 * p.D + row] = dD_acc;
            p.db_part[(int64_t)b * 4 * p.D + row] = db_acc;
        }
    }
}

template <int RP, typename xc_t>
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
    const int threads = p.CT * 4, nwarps = threads / 32;
    const size_t xpad = ((size_t)p.cap * p.CT * XE + 127) & ~(size_t)127, dpad = ((size_t)p.cap * p.CT * 4 + 127) & ~(size_t)127,
                 ppad = ((size_t)p.cap * CP * 4 + 127) & ~(size_t)127, hpad = (size_t)p.CT * kMaxState * 4;
    const size_t smem = kBwdStages * (xpad + dpad + ppad + hpad) + (size_t)2 * nwarps * kTrainCap * CP * 4 + 2 * kTrainCap * 4 +
                        2 * kBwdStages * sizeof(uint64_t);
    auto kern = ss2d_core_bwd_kernel<RP, xc_t>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_status(e);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    dim3 grid(p.tiles, 4, p.B);
    kern<<<grid, threads, smem, st>>>(tmx_row, tmx_col, tmd_row, tmd_col, tmp_row, tmp_col, p);
    return launch_status();
}

template <typename xc_t>
static int dispatch_core_bwd(int dt_pad, CoreBwdParams& p, const void* xc, const float* dY, const float* proj, cudaStream_t st) {
    switch (dt_pad) {
        case 4: return launch_core_bwd<4, xc_t>(p, xc, dY, proj, st);
        case 8: return launch_core_bwd<8, xc_t>(p, xc, dY, proj, st);
        case 12: return launch_core_bwd<12, xc_t>(p, xc, dY, proj, st);
        case 16: return launch_core_bwd<16, xc_t>(p, xc, dY, proj, st);
        case 24: return launch_core_bwd<24, xc_t>(p, xc, dY, proj, st);
        case 32: return launch_core_bwd<32, xc_t>(p, xc, dY, proj, st);
        default: return MMB_ERR_UNSUPPORTED;
    }
}

static int core_bwd_ct(int D) {            // channels per CTA: <= 48 (192 threads, two CTAs per SM), a multiple of 8
    const int tiles = (D + 47) / 48;
    int ct = (D + tiles - 1) / tiles;
    return (ct + 7) / 8 * 8;
}

}  // namespace mmb

extern "C" int mmb_ss2d_core_bwd_tiles(int D) {
    if (D <= 0) return MMB_ERR_INVALID_ARG;
    const int ct = mmb::core_bwd_ct(D);
    return (D + ct - 1) / ct;
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
    p.CT = core_bwd_ct(D); p.tiles = (D + p.CT - 1) / p.CT; p.NBmax = g.nblocks_max();
    p.T_row = g.T_row; p.NB_row = g.NB_row; p.nw = g.nw; p.T_col = g.T_col; p.NI_col = g.NI_col; p.NO_col = g.NO_col;
    p.cap = g.cap;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (xc_dtype == MMB_F32) return dispatch_core_bwd<float>(dt_pad, p, xc, dY, proj, st);
    return dispatch_core_bwd<__nv_bfloat16>(dt_pad, p, xc, dY, proj, st);
}
