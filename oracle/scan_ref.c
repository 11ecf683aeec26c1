/* Plain-C restatement of selective_scan_ref -- TEST INFRASTRUCTURE (see oracle/__init__.py).
 *
 * Follows the reference's text of the algorithm at temp.py:57-139 (third-party
 * mamba_ssm==1.0.1, README.md:19):
 *   temp.py:61-64    delta = delta + delta_bias; softplus (torch default: beta 1, threshold 20)
 *   temp.py:88-98    deltaA = exp(delta*A); deltaB_u = delta*B*u, B/C groups repeated over channels
 *   temp.py:111-125  x = deltaA*x + deltaB_u; y = <x, C_t>; last_state
 *   temp.py:135-138  out = y + u*D; out *= silu(z)
 * and, for the gradient oracle, SURVEY.md Appendix B (validated against autograd in
 * tests/test_oracle.py through oracle/selective_scan_ref.py).
 *
 * Exists so that the full BASELINE shapes (batch 64, L = 3136) can be checked in seconds: the
 * torch form needs ~10 GB of (B, D, L, N) temporaries there.  One row (b, d) per OpenMP task.
 *
 * Layouts (all contiguous): u, delta, z, out: (B, D, L); A: (D, N); Bm, Cm: (B, G, N, L);
 * Dv, delta_bias: (D); last_state: (B, D, N).  Any of z, Dv, delta_bias, last_state may be NULL.
 * The f32 entry computes in float like the reference (.float() at temp.py:59-60); the f64 entry
 * takes the same float inputs and computes in double (the "truth" the tolerances are set against).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#define MAXN 256

static inline float softplus_f(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
static inline double softplus_d(double x) { return x > 20.0 ? x : log1p(exp(x)); }

int scan_fwd_ref_f32(const float *u, const float *delta, const float *A, const float *Bm,
                     const float *Cm, const float *Dv, const float *z, const float *delta_bias,
                     int delta_softplus, float *out, float *last_state,
                     long batch, long dim, long L, long N, long G)
{
    if (N > MAXN || G <= 0 || dim % G != 0) return -1;
    const long H = dim / G;
#pragma omp parallel for schedule(static)
    for (long row = 0; row < batch * dim; ++row) {
        const long b = row / dim, d = row % dim, g = d / H;
        const float *ur = u + row * L, *dr = delta + row * L;
        const float *Bg = Bm + (b * G + g) * N * L, *Cg = Cm + (b * G + g) * N * L;
        const float *Ar = A + d * N;
        float x[MAXN];
        for (long n = 0; n < N; ++n) x[n] = 0.0f;
        const float bias = delta_bias ? delta_bias[d] : 0.0f;
        for (long t = 0; t < L; ++t) {
            float dl = dr[t] + bias;
            if (delta_softplus) dl = softplus_f(dl);
            const float dlu = dl * ur[t];
            float y = 0.0f;
            for (long n = 0; n < N; ++n) {
                x[n] = expf(dl * Ar[n]) * x[n] + dlu * Bg[n * L + t];
                y += x[n] * Cg[n * L + t];
            }
            float o = Dv ? y + ur[t] * Dv[d] : y;
            if (z) { const float zz = z[row * L + t]; o *= zz / (1.0f + expf(-zz)); }
            out[row * L + t] = o;
        }
        if (last_state) for (long n = 0; n < N; ++n) last_state[row * N + n] = x[n];
    }
    return 0;
}

int scan_fwd_ref_f64(const float *u, const float *delta, const float *A, const float *Bm,
                     const float *Cm, const float *Dv, const float *z, const float *delta_bias,
                     int delta_softplus, double *out, double *last_state,
                     long batch, long dim, long L, long N, long G)
{
    if (N > MAXN || G <= 0 || dim % G != 0) return -1;
    const long H = dim / G;
#pragma omp parallel for schedule(static)
    for (long row = 0; row < batch * dim; ++row) {
        const long b = row / dim, d = row % dim, g = d / H;
        const float *ur = u + row * L, *dr = delta + row * L;
        const float *Bg = Bm + (b * G + g) * N * L, *Cg = Cm + (b * G + g) * N * L;
        const float *Ar = A + d * N;
        double x[MAXN];
        for (long n = 0; n < N; ++n) x[n] = 0.0;
        const double bias = delta_bias ? (double)delta_bias[d] : 0.0;
        for (long t = 0; t < L; ++t) {
            double dl = (double)dr[t] + bias;
            if (delta_softplus) dl = softplus_d(dl);
            const double dlu = dl * (double)ur[t];
            double y = 0.0;
            for (long n = 0; n < N; ++n) {
                x[n] = exp(dl * (double)Ar[n]) * x[n] + dlu * (double)Bg[n * L + t];
                y += x[n] * (double)Cg[n * L + t];
            }
            double o = Dv ? y + (double)ur[t] * (double)Dv[d] : y;
            if (z) { const double zz = z[row * L + t]; o *= zz / (1.0 + exp(-zz)); }
            out[row * L + t] = o;
        }
        if (last_state) for (long n = 0; n < N; ++n) last_state[row * N + n] = x[n];
    }
    return 0;
}

/* Analytic backward in double.  Float inputs, double gradients.
 * du, ddelta, dz: (B, D, L); dA: (D, N); dB, dC: (B, G, N, L); dD, dbias: (D).
 * dz / dD / dbias may be NULL when the matching input is NULL.  Gradient buffers are overwritten.
 * Parallel over (b, g); the cross-batch sums (dA, dD, dbias) go through per-batch partials. */
int scan_bwd_ref_f64(const float *u, const float *delta, const float *A, const float *Bm,
                     const float *Cm, const float *Dv, const float *z, const float *delta_bias,
                     int delta_softplus, const float *dout,
                     double *du, double *ddelta, double *dA, double *dB, double *dC,
                     double *dD, double *dz, double *dbias,
                     long batch, long dim, long L, long N, long G)
{
    if (N > MAXN || G <= 0 || dim % G != 0) return -1;
    const long H = dim / G;
    double *pA = (double *)calloc((size_t)(batch * dim * N), sizeof(double));
    double *pD = (double *)calloc((size_t)(batch * dim), sizeof(double));
    double *pb = (double *)calloc((size_t)(batch * dim), sizeof(double));
    if (!pA || !pD || !pb) { free(pA); free(pD); free(pb); return -2; }
    int fail = 0;
#pragma omp parallel for schedule(dynamic)
    for (long bg = 0; bg < batch * G; ++bg) {
        const long b = bg / G, g = bg % G;
        const float *Bg = Bm + bg * N * L, *Cg = Cm + bg * N * L;
        double *dBg = dB + bg * N * L, *dCg = dC + bg * N * L;
        memset(dBg, 0, sizeof(double) * (size_t)(N * L));
        memset(dCg, 0, sizeof(double) * (size_t)(N * L));
        double *hprev = (double *)malloc(sizeof(double) * (size_t)(L * N));
        double *dls = (double *)malloc(sizeof(double) * (size_t)L);
        double *dys = (double *)malloc(sizeof(double) * (size_t)L);
        if (!hprev || !dls || !dys) { fail = 1; free(hprev); free(dls); free(dys); continue; }
        for (long dd = 0; dd < H; ++dd) {
            const long d = g * H + dd, row = b * dim + d;
            const float *ur = u + row * L, *dr = delta + row * L, *Ar = A + d * N;
            const double bias = delta_bias ? (double)delta_bias[d] : 0.0;
            const double Dd = Dv ? (double)Dv[d] : 0.0;
            double h[MAXN], gcar[MAXN];
            for (long n = 0; n < N; ++n) { h[n] = 0.0; gcar[n] = 0.0; }
            double accD = 0.0;
            for (long t = 0; t < L; ++t) {
                const double xr = (double)dr[t] + bias;
                const double dl = delta_softplus ? softplus_d(xr) : xr;
                dls[t] = dl;
                const double dlu = dl * (double)ur[t];
                double y = 0.0;
                for (long n = 0; n < N; ++n) {
                    hprev[t * N + n] = h[n];
                    h[n] = exp(dl * (double)Ar[n]) * h[n] + dlu * (double)Bg[n * L + t];
                    y += h[n] * (double)Cg[n * L + t];
                }
                const double pre = y + Dd * (double)ur[t];
                double dy = (double)dout[row * L + t];
                if (z) {
                    const double zz = z[row * L + t], sg = 1.0 / (1.0 + exp(-zz));
                    if (dz) dz[row * L + t] = dy * pre * sg * (1.0 + zz * (1.0 - sg));
                    dy *= zz * sg;
                }
                dys[t] = dy;
                accD += dy * (double)ur[t];
            }
            double accb = 0.0;
            for (long t = L - 1; t >= 0; --t) {
                const double dl = dls[t], dy = dys[t], ut = (double)ur[t];
                double s_du = 0.0, s_dl = 0.0;
                for (long n = 0; n < N; ++n) {
                    const double a = exp(dl * (double)Ar[n]);
                    const double hp = hprev[t * N + n];
                    const double Bt = (double)Bg[n * L + t];
                    const double ht = a * hp + dl * ut * Bt;
                    const double gt = dy * (double)Cg[n * L + t] + gcar[n];
                    dCg[n * L + t] += dy * ht;
                    dBg[n * L + t] += gt * dl * ut;
                    s_du += gt * Bt;
                    s_dl += gt * (Bt * ut + hp * a * (double)Ar[n]);
                    pA[row * N + n] += gt * hp * a * dl;
                    gcar[n] = a * gt;
                }
                du[row * L + t] = Dd * dy + dl * s_du;
                double draw = s_dl;
                if (delta_softplus) {
                    const double xr = (double)dr[t] + bias;
                    draw *= xr > 20.0 ? 1.0 : 1.0 / (1.0 + exp(-xr));
                }
                ddelta[row * L + t] = draw;
                accb += draw;
            }
            pD[row] = accD;
            pb[row] = accb;
        }
        free(hprev); free(dls); free(dys);
    }
    for (long d = 0; d < dim; ++d) {
        double sD = 0.0, sb = 0.0;
        for (long n = 0; n < N; ++n) dA[d * N + n] = 0.0;
        for (long b = 0; b < batch; ++b) {
            sD += pD[b * dim + d];
            sb += pb[b * dim + d];
            for (long n = 0; n < N; ++n) dA[d * N + n] += pA[(b * dim + d) * N + n];
        }
        if (dD) dD[d] = sD;
        if (dbias) dbias[d] = sb;
    }
    free(pA); free(pD); free(pb);
    return fail ? -2 : 0;
}
