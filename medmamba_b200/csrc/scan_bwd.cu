// selective_scan_fn backward at the interface layout -- the drop-in for mamba_ssm's
// selective_scan_cuda.bwd (SURVEY.md Appendix B for the formulas; forward semantics temp.py:57-139).
//
// Same ownership as the forward: a CTA of 128 threads owns 32 channel rows of one (batch, group),
// four lanes per row, four states per lane.  The sequence is walked BACKWARDS in chunks of
// kBwdChunk = 16 steps; the forward kernel stored the state at the end of every chunk
// (chunk_state), so each chunk is
//   phase A  re-run forward from its checkpoint, parking h_{t-1} of every step in shared memory,
//   phase B  run the reverse recurrence g_t = dy_t C_t + a_{t+1} g_{t+1} and emit all gradients.
// One exp per (row, step, state) in each phase: the backward costs two forwards of MUFU work.
//
// Reductions are deterministic (no float atomics):
//   dB, dC   sum over the rows of a group: an 8-row transposing shuffle reduction inside the warp,
//            per-warp tiles in shared memory added in fixed order, one partial per row tile in HBM
//            (summed over row tiles on the host side);
//   dA, dD, d(delta_bias)   per-thread register accumulators over the whole sequence, one partial per
//            batch element in HBM (summed over the batch on the host side).
#include "common.cuh"

namespace mmb {

constexpr int kBwdChunk = 16;

struct ScanBwdParams {
    const void* u; const void* delta; const void* Bm; const void* Cm; const void* z; const void* dout;
    const float* A; const float* Dv; const float* bias; const float* chunk_state;
    void* du; void* ddelta; void* dz;
    float* dB_part; float* dC_part;      // (tiles, batch, G, N, L)
    float* dA_part;                      // (batch, dim, N)
    float* dD_part; float* dbias_part;   // (batch, dim)
    int batch, dim, L, N, G, H, nchunks, softplus, tiles;
    int64_t u_bs, u_ds, d_bs, d_ds, z_bs, z_ds, o_bs, o_ds;          // o_*: dout strides
    int64_t B_bs, B_gs, B_ns, B_ls, C_bs, C_gs, C_ns, C_ls;
};

template <typename T>
__device__ __forceinline__ float4 ld_row4(const T* row, int t, int len, bool vec) {
    if (vec && t + 4 <= len) return load4<T>(row + t);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t + 0 < len) v.x = to_f<T>(row[t + 0]);
    if (t + 1 < len) v.y = to_f<T>(row[t + 1]);
    if (t + 2 < len) v.z = to_f<T>(row[t + 2]);
    if (t + 3 < len) v.w = to_f<T>(row[t + 3]);
    return v;
}
template <typename T>
__device__ __forceinline__ void st_row4(T* row, int t, int len, bool vec, float4 v) {
    if (vec && t + 4 <= len) { store4<T>(row + t, v); return; }
    if (t + 0 < len) row[t + 0] = from_f<T>(v.x);
    if (t + 1 < len) row[t + 1] = from_f<T>(v.y);
    if (t + 2 < len) row[t + 2] = from_f<T>(v.z);
    if (t + 3 < len) row[t + 3] = from_f<T>(v.w);
}

template <typename io_t, typename bc_t, bool HAS_Z>
__global__ void __launch_bounds__(128) scan_bwd_kernel(const ScanBwdParams p) {
    constexpr int NT = 128, S = 4, RT = NT / S, NS = kMaxState / S, T = kBwdChunk, TP = T + 4, T4 = T / 4;
    extern __shared__ __align__(16) float smem[];
    float* su = smem;                          // [RT][TP] u
    float* sdl = su + RT * TP;                 // [RT][TP] delta (after softplus); d(delta raw) on the way out
    float* ssg = sdl + RT * TP;                // [RT][TP] d softplus / d raw
    float* sdy = ssg + RT * TP;                // [RT][TP] dy = dout (* silu(z)); du on the way out
    float* sB = sdy + RT * TP;                 // [16][TP]
    float* sC = sB + kMaxState * TP;           // [16][TP]
    float* shist = sC + kMaxState * TP;        // [T][NS][NT] h_{t-1}
    float* swred = shist + T * NS * NT;        // [4 warps][T][32]
    float* sz = swred + 4 * T * 32;            // HAS_Z: [RT][TP] z, [RT][TP] raw dout, [RT][TP] y before the gate
    float* sdr = sz + RT * TP;
    float* sy = sdr + RT * TP;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = tid / S, q = tid % S;
    const int b = blockIdx.z, g = blockIdx.y, tile = blockIdx.x;
    const int row0 = tile * RT;
    const int rows_here = min(RT, p.H - row0);
    const bool valid = r < rows_here;
    const int d = g * p.H + row0 + (valid ? r : 0);

    const io_t* ub = reinterpret_cast<const io_t*>(p.u) + (int64_t)b * p.u_bs;
    const io_t* db = reinterpret_cast<const io_t*>(p.delta) + (int64_t)b * p.d_bs;
    const io_t* zb = HAS_Z ? reinterpret_cast<const io_t*>(p.z) + (int64_t)b * p.z_bs : nullptr;
    const io_t* gob = reinterpret_cast<const io_t*>(p.dout) + (int64_t)b * p.o_bs;
    const bc_t* Bb = reinterpret_cast<const bc_t*>(p.Bm) + (int64_t)b * p.B_bs + (int64_t)g * p.B_gs;
    const bc_t* Cb = reinterpret_cast<const bc_t*>(p.Cm) + (int64_t)b * p.C_bs + (int64_t)g * p.C_gs;
    // gradients of u / delta / z are dense (batch, dim, L)
    io_t* dub = reinterpret_cast<io_t*>(p.du) + (int64_t)b * p.dim * p.L;
    io_t* ddb = reinterpret_cast<io_t*>(p.ddelta) + (int64_t)b * p.dim * p.L;
    io_t* dzb = HAS_Z ? reinterpret_cast<io_t*>(p.dz) + (int64_t)b * p.dim * p.L : nullptr;

    constexpr int VA = vec4_align<io_t>();
    const bool vec_u = ((reinterpret_cast<uintptr_t>(ub) % VA) == 0) && (p.u_ds % 4 == 0);
    const bool vec_d = ((reinterpret_cast<uintptr_t>(db) % VA) == 0) && (p.d_ds % 4 == 0);
    const bool vec_z = HAS_Z && ((reinterpret_cast<uintptr_t>(zb) % VA) == 0) && (p.z_ds % 4 == 0);
    const bool vec_o = ((reinterpret_cast<uintptr_t>(gob) % VA) == 0) && (p.o_ds % 4 == 0);
    const bool vec_g = ((reinterpret_cast<uintptr_t>(dub) % VA) == 0) && (p.L % 4 == 0);

    float Ap[NS], Araw[NS], gcar[NS], dA[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        const int n = q + S * j;
        Araw[j] = (valid && n < p.N) ? p.A[(int64_t)d * p.N + n] : 0.f;
        Ap[j] = Araw[j] * kLog2e;
        gcar[j] = 0.f; dA[j] = 0.f;
    }
    const float Dd = (valid && p.Dv) ? p.Dv[d] : 0.f;
    float dD_acc = 0.f, dbias_acc = 0.f;

    for (int c = p.nchunks - 1; c >= 0; --c) {
        const int t0 = c * T;
        const int len = min(T, p.L - t0);
        // ---- stage -------------------------------------------------------------------------------
        for (int idx = tid; idx < RT * T4; idx += NT) {
            const int rr = idx / T4, tt = (idx % T4) * 4;
            float4 uv = make_float4(0.f, 0.f, 0.f, 0.f), dv = uv, sg = uv, dy = uv, zv = uv, dr = uv;
            if (rr < rows_here) {
                const int64_t dd = g * p.H + row0 + rr;
                uv = ld_row4<io_t>(ub + dd * p.u_ds + t0, tt, len, vec_u);
                dv = ld_row4<io_t>(db + dd * p.d_ds + t0, tt, len, vec_d);
                dy = ld_row4<io_t>(gob + dd * p.o_ds + t0, tt, len, vec_o);
                const float bs = p.bias ? p.bias[dd] : 0.f;
                float raw[4] = {dv.x + bs, dv.y + bs, dv.z + bs, dv.w + bs};
                float dl[4], sgm[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (p.softplus) { dl[i] = softplus_f(raw[i]); sgm[i] = raw[i] > 20.f ? 1.f : sigmoid_f(raw[i]); }
                    else { dl[i] = raw[i]; sgm[i] = 1.f; }
                    if (tt + i >= len) { dl[i] = 0.f; sgm[i] = 0.f; }
                }
                dv = make_float4(dl[0], dl[1], dl[2], dl[3]);
                sg = make_float4(sgm[0], sgm[1], sgm[2], sgm[3]);
                if (HAS_Z) {
                    zv = ld_row4<io_t>(zb + dd * p.z_ds + t0, tt, len, vec_z);
                    dr = dy;
                    dy.x *= silu_f(zv.x); dy.y *= silu_f(zv.y); dy.z *= silu_f(zv.z); dy.w *= silu_f(zv.w);
                }
            }
            *reinterpret_cast<float4*>(su + rr * TP + tt) = uv;
            *reinterpret_cast<float4*>(sdl + rr * TP + tt) = dv;
            *reinterpret_cast<float4*>(ssg + rr * TP + tt) = sg;
            *reinterpret_cast<float4*>(sdy + rr * TP + tt) = dy;
            if (HAS_Z) {
                *reinterpret_cast<float4*>(sz + rr * TP + tt) = zv;
                *reinterpret_cast<float4*>(sdr + rr * TP + tt) = dr;
            }
        }
        for (int idx = tid; idx < kMaxState * T; idx += NT) {
            int n, tt;
            if (p.B_ls == 1 || p.B_ns != 1) { n = idx / T; tt = idx % T; } else { n = idx % kMaxState; tt = idx / kMaxState; }
            sB[n * TP + tt] = (n < p.N && tt < len) ? to_f<bc_t>(Bb[(int64_t)n * p.B_ns + (int64_t)(t0 + tt) * p.B_ls]) : 0.f;
            if (p.C_ls == 1 || p.C_ns != 1) { n = idx / T; tt = idx % T; } else { n = idx % kMaxState; tt = idx / kMaxState; }
            sC[n * TP + tt] = (n < p.N && tt < len) ? to_f<bc_t>(Cb[(int64_t)n * p.C_ns + (int64_t)(t0 + tt) * p.C_ls]) : 0.f;
        }
        // state at the start of the chunk = checkpoint written by the forward at the end of chunk c-1
        float h[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            const int n = q + S * j;
            h[j] = (c > 0 && valid && n < p.N)
                       ? p.chunk_state[(((int64_t)b * p.dim + d) * p.nchunks + (c - 1)) * p.N + n] : 0.f;
        }
        __syncthreads();

        // Both phases walk the chunk four steps at a time: one LDS.128 per row tile (u, delta, dy, d softplus) and per state
        // (B_n, C_n) serves four steps, the state history is one float4 per (step, thread), and the per-state arithmetic runs
        // as packed FMUL2 / FFMA2 pairs.  Steps beyond the sequence have delta = 0, dy = 0: a = 1, every product 0 -- they
        // pass through both phases as no-ops, so there is no per-step validity test.
        const int steps4 = (len + 3) & ~3;
        float4* hist4 = reinterpret_cast<float4*>(shist) + tid;          // [t][NT] float4 = the lane's four states
        // ---- phase A: forward through the chunk, parking h_{t-1} ------------------------------------
        for (int t4 = 0; t4 < steps4; t4 += 4) {
            const float4 d4 = *reinterpret_cast<const float4*>(sdl + r * TP + t4);
            const float4 u4 = *reinterpret_cast<const float4*>(su + r * TP + t4);
            float4 B4[NS], C4[NS];
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                B4[j] = *reinterpret_cast<const float4*>(sB + (q + S * j) * TP + t4);
                if (HAS_Z) C4[j] = *reinterpret_cast<const float4*>(sC + (q + S * j) * TP + t4);
            }
            const float dls[4] = {d4.x, d4.y, d4.z, d4.w}, uus[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) {
                const float dl = dls[s4], dlu = dl * uus[s4];
                hist4[(t4 + s4) * NT] = make_float4(h[0], h[1], h[2], h[3]);
                const float bb[4] = {s4 == 0 ? B4[0].x : s4 == 1 ? B4[0].y : s4 == 2 ? B4[0].z : B4[0].w,
                                     s4 == 0 ? B4[1].x : s4 == 1 ? B4[1].y : s4 == 2 ? B4[1].z : B4[1].w,
                                     s4 == 0 ? B4[2].x : s4 == 1 ? B4[2].y : s4 == 2 ? B4[2].z : B4[2].w,
                                     s4 == 0 ? B4[3].x : s4 == 1 ? B4[3].y : s4 == 2 ? B4[3].z : B4[3].w};
                float x[4], w[4];
                mul2(x[0], x[1], dl, dl, Ap[0], Ap[1]); mul2(x[2], x[3], dl, dl, Ap[2], Ap[3]);
                mul2(w[0], w[1], dlu, dlu, bb[0], bb[1]); mul2(w[2], w[3], dlu, dlu, bb[2], bb[3]);
                const float a0 = ex2_approx(x[0]), a1 = ex2_approx(x[1]), a2 = ex2_approx(x[2]), a3 = ex2_approx(x[3]);
                fma2(h[0], h[1], a0, a1, h[0], h[1], w[0], w[1]);
                fma2(h[2], h[3], a2, a3, h[2], h[3], w[2], w[3]);
                if (HAS_Z) {
                    const float cc[4] = {s4 == 0 ? C4[0].x : s4 == 1 ? C4[0].y : s4 == 2 ? C4[0].z : C4[0].w,
                                         s4 == 0 ? C4[1].x : s4 == 1 ? C4[1].y : s4 == 2 ? C4[1].z : C4[1].w,
                                         s4 == 0 ? C4[2].x : s4 == 1 ? C4[2].y : s4 == 2 ? C4[2].z : C4[2].w,
                                         s4 == 0 ? C4[3].x : s4 == 1 ? C4[3].y : s4 == 2 ? C4[3].z : C4[3].w};
                    float yp = fmaf(h[0], cc[0], fmaf(h[1], cc[1], fmaf(h[2], cc[2], h[3] * cc[3])));
                    yp += __shfl_xor_sync(0xffffffffu, yp, 1);
                    yp += __shfl_xor_sync(0xffffffffu, yp, 2);
                    if (q == 0) sy[r * TP + t4 + s4] = fmaf(Dd, uus[s4], yp);
                }
            }
        }
        // ---- phase B: reverse recurrence ---------------------------------------------------------------
        // h holds h_t of the step being processed: the state phase A ended in, then the h_{t-1} each step loads.
        for (int t4 = steps4 - 4; t4 >= 0; t4 -= 4) {
            const float4 d4 = *reinterpret_cast<const float4*>(sdl + r * TP + t4);
            const float4 u4 = *reinterpret_cast<const float4*>(su + r * TP + t4);
            const float4 y4 = *reinterpret_cast<const float4*>(sdy + r * TP + t4);
            const float4 g4 = *reinterpret_cast<const float4*>(ssg + r * TP + t4);
            float4 B4[NS], C4[NS];
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                B4[j] = *reinterpret_cast<const float4*>(sB + (q + S * j) * TP + t4);
                C4[j] = *reinterpret_cast<const float4*>(sC + (q + S * j) * TP + t4);
            }
            const float dls[4] = {d4.x, d4.y, d4.z, d4.w}, uus[4] = {u4.x, u4.y, u4.z, u4.w};
            const float dys[4] = {y4.x, y4.y, y4.z, y4.w}, sgs[4] = {g4.x, g4.y, g4.z, g4.w};
            float duo[4], ddo[4];
#pragma unroll
            for (int s4 = 3; s4 >= 0; --s4) {
                const int t = t4 + s4;
                const float dl = dls[s4], uu = uus[s4], dy = dys[s4], sg = sgs[s4];
                const float dlu = dl * uu;
                const float4 hp4 = hist4[t * NT];
                const float hp[4] = {hp4.x, hp4.y, hp4.z, hp4.w};
                const float bb[4] = {s4 == 0 ? B4[0].x : s4 == 1 ? B4[0].y : s4 == 2 ? B4[0].z : B4[0].w,
                                     s4 == 0 ? B4[1].x : s4 == 1 ? B4[1].y : s4 == 2 ? B4[1].z : B4[1].w,
                                     s4 == 0 ? B4[2].x : s4 == 1 ? B4[2].y : s4 == 2 ? B4[2].z : B4[2].w,
                                     s4 == 0 ? B4[3].x : s4 == 1 ? B4[3].y : s4 == 2 ? B4[3].z : B4[3].w};
                const float cc[4] = {s4 == 0 ? C4[0].x : s4 == 1 ? C4[0].y : s4 == 2 ? C4[0].z : C4[0].w,
                                     s4 == 0 ? C4[1].x : s4 == 1 ? C4[1].y : s4 == 2 ? C4[1].z : C4[1].w,
                                     s4 == 0 ? C4[2].x : s4 == 1 ? C4[2].y : s4 == 2 ? C4[2].z : C4[2].w,
                                     s4 == 0 ? C4[3].x : s4 == 1 ? C4[3].y : s4 == 2 ? C4[3].z : C4[3].w};
                float x[4], gt[4], ga[4], tt[4], v[8];
                mul2(x[0], x[1], dl, dl, Ap[0], Ap[1]); mul2(x[2], x[3], dl, dl, Ap[2], Ap[3]);
                const float a0 = ex2_approx(x[0]), a1 = ex2_approx(x[1]), a2 = ex2_approx(x[2]), a3 = ex2_approx(x[3]);
                fma2(gt[0], gt[1], dy, dy, cc[0], cc[1], gcar[0], gcar[1]);                 // g_t
                fma2(gt[2], gt[3], dy, dy, cc[2], cc[3], gcar[2], gcar[3]);
                mul2(ga[0], ga[1], a0, a1, gt[0], gt[1]); mul2(ga[2], ga[3], a2, a3, gt[2], gt[3]);   // a_t g_t: the carry
                mul2(tt[0], tt[1], ga[0], ga[1], hp[0], hp[1]); mul2(tt[2], tt[3], ga[2], ga[3], hp[2], hp[3]);   // g h_{t-1} a
                mul2(v[0], v[1], gt[0], gt[1], dlu, dlu); mul2(v[2], v[3], gt[2], gt[3], dlu, dlu);   // dB_n of this row
                mul2(v[4], v[5], dy, dy, h[0], h[1]); mul2(v[6], v[7], dy, dy, h[2], h[3]);           // dC_n of this row
                float adu0, adu1, adA0, adA1;
                mul2(adu0, adu1, gt[0], gt[1], bb[0], bb[1]);
                fma2(adu0, adu1, gt[2], gt[3], bb[2], bb[3], adu0, adu1);                             // sum_n g B
                mul2(adA0, adA1, tt[0], tt[1], Araw[0], Araw[1]);
                fma2(adA0, adA1, tt[2], tt[3], Araw[2], Araw[3], adA0, adA1);                         // sum_n g h a A
                fma2(dA[0], dA[1], tt[0], tt[1], dl, dl, dA[0], dA[1]);
                fma2(dA[2], dA[3], tt[2], tt[3], dl, dl, dA[2], dA[3]);
#pragma unroll
                for (int j = 0; j < NS; ++j) { gcar[j] = ga[j]; h[j] = hp[j]; }
                float adu = adu0 + adu1, adA = adA0 + adA1;
                adu += __shfl_xor_sync(0xffffffffu, adu, 1); adu += __shfl_xor_sync(0xffffffffu, adu, 2);
                adA += __shfl_xor_sync(0xffffffffu, adA, 1); adA += __shfl_xor_sync(0xffffffffu, adA, 2);
                const float ddraw = fmaf(adu, uu, adA) * sg;       // d delta = sum_n g (B u + h_{t-1} a A_n), times d softplus
                duo[s4] = fmaf(Dd, dy, dl * adu);
                ddo[s4] = ddraw;
                if (q == 0) {
                    dD_acc = fmaf(dy, uu, dD_acc);
                    dbias_acc += ddraw;
                }
                // sum the 8 values over the 8 rows of this warp (lanes with equal q): transposing reduction
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
                    const int j = idx & 3, n = q + S * j;
                    swred[(warp * T + t) * 32 + (idx < 4 ? n : 16 + n)] = v[0];
                }
            }
            if (q == 0) {
                *reinterpret_cast<float4*>(sdy + r * TP + t4) = make_float4(duo[0], duo[1], duo[2], duo[3]);   // du
                *reinterpret_cast<float4*>(sdl + r * TP + t4) = make_float4(ddo[0], ddo[1], ddo[2], ddo[3]);   // d(delta raw)
            }
        }
        __syncthreads();
        // ---- write out: du, d(delta), dz, and this row tile's partial dB / dC --------------------------
        for (int idx = tid; idx < RT * T4; idx += NT) {
            const int rr = idx / T4, tt = (idx % T4) * 4;
            if (rr < rows_here && tt < len) {
                const int64_t dd = g * p.H + row0 + rr;
                st_row4<io_t>(dub + dd * p.L + t0, tt, len, vec_g, *reinterpret_cast<const float4*>(sdy + rr * TP + tt));
                st_row4<io_t>(ddb + dd * p.L + t0, tt, len, vec_g, *reinterpret_cast<const float4*>(sdl + rr * TP + tt));
                if (HAS_Z) {
                    const float4 zv = *reinterpret_cast<const float4*>(sz + rr * TP + tt);
                    const float4 dr = *reinterpret_cast<const float4*>(sdr + rr * TP + tt);
                    const float4 yv = *reinterpret_cast<const float4*>(sy + rr * TP + tt);
                    const float zz[4] = {zv.x, zv.y, zv.z, zv.w}, dd4[4] = {dr.x, dr.y, dr.z, dr.w}, yy[4] = {yv.x, yv.y, yv.z, yv.w};
                    float o[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float sgz = sigmoid_f(zz[i]);
                        o[i] = dd4[i] * yy[i] * sgz * (1.f + zz[i] * (1.f - sgz));
                    }
                    st_row4<io_t>(dzb + dd * p.L + t0, tt, len, vec_g, make_float4(o[0], o[1], o[2], o[3]));
                }
            }
        }
        {
            const int64_t base = (((int64_t)tile * p.batch + b) * p.G + g) * p.N * (int64_t)p.L;
            for (int idx = tid; idx < 32 * T; idx += NT) {
                const int vsl = idx / T, tt = idx % T;
                const int n = vsl & 15;
                if (tt < len && n < p.N) {
                    float s = 0.f;
#pragma unroll
                    for (int w = 0; w < 4; ++w) s += swred[(w * T + tt) * 32 + vsl];
                    (vsl < 16 ? p.dB_part : p.dC_part)[base + (int64_t)n * p.L + t0 + tt] = s;
                }
            }
        }
        __syncthreads();
    }
    if (valid) {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            const int n = q + S * j;
            if (n < p.N) p.dA_part[((int64_t)b * p.dim + d) * p.N + n] = dA[j];
        }
        if (q == 0) {
            p.dD_part[(int64_t)b * p.dim + d] = dD_acc;
            p.dbias_part[(int64_t)b * p.dim + d] = dbias_acc;
        }
    }
}

template <bool HAS_Z> constexpr size_t scan_bwd_smem() {
    constexpr int RT = 32, T = kBwdChunk, TP = T + 4;
    return sizeof(float) * (size_t)((4 + (HAS_Z ? 3 : 0)) * RT * TP + 2 * kMaxState * TP + T * 4 * 128 + 4 * T * 32);
}

template <typename io_t, typename bc_t, bool HAS_Z>
static int launch_scan_bwd(const ScanBwdParams& p, cudaStream_t st) {
    constexpr size_t smem = scan_bwd_smem<HAS_Z>();
    auto kern = scan_bwd_kernel<io_t, bc_t, HAS_Z>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_status(e);
    }
    dim3 grid(p.tiles, p.G, p.batch);
    kern<<<grid, 128, smem, st>>>(p);
    return launch_status();
}

template <typename io_t, typename bc_t>
static int dispatch_bwd_z(const ScanBwdParams& p, cudaStream_t st) {
    return p.z ? launch_scan_bwd<io_t, bc_t, true>(p, st) : launch_scan_bwd<io_t, bc_t, false>(p, st);
}

template <typename io_t>
static int dispatch_bwd_bc(const ScanBwdParams& p, int io_dtype, int bc_dtype, cudaStream_t st) {
    if (bc_dtype == MMB_F32) return dispatch_bwd_z<io_t, float>(p, st);
    if (bc_dtype == io_dtype) return dispatch_bwd_z<io_t, io_t>(p, st);
    return MMB_ERR_UNSUPPORTED;
}

}  // namespace mmb

extern "C" int mmb_scan_bwd_row_tiles(int dim, int ngroups) {
    if (dim <= 0 || ngroups <= 0 || dim % ngroups != 0) return MMB_ERR_INVALID_ARG;
    return (dim / ngroups + 31) / 32;
}

extern "C" int mmb_scan_bwd(const void* u, const void* delta, const float* A, const void* Bm, const void* Cm,
                            const float* Dv, const void* z, const float* delta_bias, const void* dout,
                            const float* chunk_state, void* du, void* ddelta, void* dz,
                            float* dB_part, float* dC_part, float* dA_part, float* dD_part, float* dbias_part,
                            int batch, int dim, int seqlen, int dstate, int ngroups,
                            int64_t u_bs, int64_t u_ds, int64_t delta_bs, int64_t delta_ds,
                            int64_t z_bs, int64_t z_ds, int64_t dout_bs, int64_t dout_ds,
                            int64_t B_bs, int64_t B_gs, int64_t B_ns, int64_t B_ls,
                            int64_t C_bs, int64_t C_gs, int64_t C_ns, int64_t C_ls,
                            int delta_softplus, int io_dtype, int bc_dtype, void* stream) {
    using namespace mmb;
    if (!u || !delta || !A || !Bm || !Cm || !dout || !du || !ddelta || !dB_part || !dC_part || !dA_part ||
        !dD_part || !dbias_part) return MMB_ERR_INVALID_ARG;
    if (z && !dz) return MMB_ERR_INVALID_ARG;
    if (batch < 0 || dim <= 0 || seqlen < 0 || dstate <= 0 || ngroups <= 0 || dim % ngroups != 0) return MMB_ERR_INVALID_ARG;
    if (dstate > kMaxState || batch > 65535 || ngroups > 65535) return MMB_ERR_UNSUPPORTED;
    if (batch == 0 || seqlen == 0) return MMB_OK;
    ScanBwdParams p;
    p.u = u; p.delta = delta; p.Bm = Bm; p.Cm = Cm; p.z = z; p.dout = dout;
    p.A = A; p.Dv = Dv; p.bias = delta_bias; p.chunk_state = chunk_state;
    p.du = du; p.ddelta = ddelta; p.dz = dz;
    p.dB_part = dB_part; p.dC_part = dC_part; p.dA_part = dA_part; p.dD_part = dD_part; p.dbias_part = dbias_part;
    p.batch = batch; p.dim = dim; p.L = seqlen; p.N = dstate; p.G = ngroups; p.H = dim / ngroups;
    p.nchunks = (seqlen + kBwdChunk - 1) / kBwdChunk;
    if (p.nchunks > 1 && !chunk_state) return MMB_ERR_INVALID_ARG;
    p.softplus = delta_softplus; p.tiles = (p.H + 31) / 32;
    p.u_bs = u_bs; p.u_ds = u_ds; p.d_bs = delta_bs; p.d_ds = delta_ds; p.z_bs = z_bs; p.z_ds = z_ds;
    p.o_bs = dout_bs; p.o_ds = dout_ds;
    p.B_bs = B_bs; p.B_gs = B_gs; p.B_ns = B_ns; p.B_ls = B_ls; p.C_bs = C_bs; p.C_gs = C_gs; p.C_ns = C_ns; p.C_ls = C_ls;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (io_dtype) {
        case MMB_F32: return dispatch_bwd_bc<float>(p, io_dtype, bc_dtype, st);
        case MMB_BF16: return dispatch_bwd_bc<__nv_bfloat16>(p, io_dtype, bc_dtype, st);
        case MMB_F16: return dispatch_bwd_bc<__half>(p, io_dtype, bc_dtype, st);
        default: return MMB_ERR_INVALID_ARG;
    }
}
