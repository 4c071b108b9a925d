/* ORACLE (test infrastructure, not product code).
 *
 * Plain-C restatement of the reference's SS2D hot path, sequential in L, so that the parity tests can
 * check the CUDA kernels at the BASELINE.json sizes (L up to 524288) in seconds:
 *
 *   vmasr_ref_scan_fwd   selective_scan_ref, kernels/selective_scan/test_selective_scan.py:287-367
 *                        (same recurrence as selective_scan_fwd_kernel.cuh:113-161)
 *   vmasr_ref_scan_bwd   gradients of selective_scan_bwd_kernel.cuh:125-272 (adjoint recurrence)
 *   vmasr_ref_cross_scan / vmasr_ref_cross_merge   model/vmamba.py:27-73 (index maps, fp32, bit-exact)
 *
 * All accumulation is double; inputs are float (the caller up-casts half types).  Pinned against the
 * golden fixtures by tests/test_oracle_golden.py.  Built by oracle/Makefile into oracle/_build/.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this library.
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

static double softplus_ref(double x) { return x <= 20.0 ? log1p(exp(x)) : x; }

/* u, delta, out: (B, D, L) contiguous; A: (D, N); Bm, Cm: (B, G, N, L); Dv, bias: (D,) or NULL.
 * last: (B, D, N) or NULL.  chunk_state: (B, D, n_chunks, 2N) or NULL -- (cumulative decay, state) at the
 * end of every `chunk` positions, the layout of the reference's `x` (fwd_kernel.cuh:155-158). */
void vmasr_ref_scan_fwd(const float *u, const float *delta, const float *A, const float *Bm,
                        const float *Cm, const float *Dv, const float *bias, int softplus,
                        int B, int D, int L, int N, int G, double *out, double *last,
                        double *chunk_state, int chunk)
{
    const int per = D / G;
    const int n_chunks = chunk > 0 ? (L + chunk - 1) / chunk : 0;
    double *y = (double *)malloc(sizeof(double) * (size_t)L);
    for (int b = 0; b < B; ++b)
        for (int d = 0; d < D; ++d) {
            const float *ur = u + ((size_t)b * D + d) * L;
            const float *dr = delta + ((size_t)b * D + d) * L;
            const int g = d / per;
            for (int l = 0; l < L; ++l) y[l] = Dv ? (double)Dv[d] * ur[l] : 0.0;
            for (int n = 0; n < N; ++n) {
                const float *Br = Bm + (((size_t)b * G + g) * N + n) * L;
                const float *Cr = Cm + (((size_t)b * G + g) * N + n) * L;
                const double a = A[(size_t)d * N + n];
                double h = 0.0, p = 1.0;
                for (int l = 0; l < L; ++l) {
                    double dt = (double)dr[l] + (bias ? (double)bias[d] : 0.0);
                    if (softplus) dt = softplus_ref(dt);
                    const double decay = exp(dt * a);
                    h = decay * h + dt * Br[l] * ur[l];
                    p *= decay;
                    y[l] += Cr[l] * h;
                    if (chunk_state && ((l + 1) % chunk == 0 || l == L - 1)) {
                        double *cs = chunk_state + ((((size_t)b * D + d) * n_chunks + l / chunk) * N + n) * 2;
                        cs[0] = p;
                        cs[1] = h;
                    }
                }
                if (last) last[((size_t)b * D + d) * N + n] = h;
            }
            memcpy(out + ((size_t)b * D + d) * L, y, sizeof(double) * (size_t)L);
        }
    free(y);
}

/* Gradients.  du, ddelta: (B, D, L); dA: (D, N); dB, dC: (B, G, N, L); dD, dbias: (D,) (may be NULL).
 * All outputs are overwritten (zeroed here first). */
void vmasr_ref_scan_bwd(const float *u, const float *delta, const float *A, const float *Bm,
                        const float *Cm, const float *Dv, const float *bias, int softplus,
                        const float *dout, int B, int D, int L, int N, int G,
                        double *du, double *ddelta, double *dA, double *dB, double *dC,
                        double *dD, double *dbias)
{
    const int per = D / G;
    memset(du, 0, sizeof(double) * (size_t)B * D * L);
    memset(ddelta, 0, sizeof(double) * (size_t)B * D * L);
    memset(dA, 0, sizeof(double) * (size_t)D * N);
    memset(dB, 0, sizeof(double) * (size_t)B * G * N * L);
    memset(dC, 0, sizeof(double) * (size_t)B * G * N * L);
    if (dD) memset(dD, 0, sizeof(double) * (size_t)D);
    if (dbias) memset(dbias, 0, sizeof(double) * (size_t)D);
    double *hs = (double *)malloc(sizeof(double) * (size_t)L);
    double *dts = (double *)malloc(sizeof(double) * (size_t)L);
    double *dec = (double *)malloc(sizeof(double) * (size_t)L);
    double *ddt = (double *)malloc(sizeof(double) * (size_t)L);
    for (int b = 0; b < B; ++b)
        for (int d = 0; d < D; ++d) {
            const size_t row = ((size_t)b * D + d) * L;
            const float *ur = u + row, *dr = delta + row, *gr = dout + row;
            const int g = d / per;
            const double bs = bias ? (double)bias[d] : 0.0;
            for (int l = 0; l < L; ++l) {
                double x = (double)dr[l] + bs;
                dts[l] = softplus ? softplus_ref(x) : x;
                ddt[l] = 0.0;
                du[row + l] = Dv ? (double)Dv[d] * gr[l] : 0.0;
                if (dD) dD[d] += (double)gr[l] * ur[l];
            }
            for (int n = 0; n < N; ++n) {
                const size_t brow = (((size_t)b * G + g) * N + n) * L;
                const float *Br = Bm + brow, *Cr = Cm + brow;
                const double a = A[(size_t)d * N + n];
                double h = 0.0;
                for (int l = 0; l < L; ++l) {
                    dec[l] = exp(dts[l] * a);
                    h = dec[l] * h + dts[l] * Br[l] * ur[l];
                    hs[l] = h;
                }
                double adj = 0.0; /* decay[l+1] * g[l+1] */
                for (int l = L - 1; l >= 0; --l) {
                    const double gl = (double)Cr[l] * gr[l] + adj;
                    const double drive = dts[l] * Br[l] * ur[l];
                    const double carried = hs[l] - drive; /* decay[l] * h[l-1] */
                    du[row + l] += gl * dts[l] * Br[l];
                    ddt[l] += gl * ((double)Br[l] * ur[l] + a * carried);
                    dA[(size_t)d * N + n] += gl * dts[l] * carried;
                    dB[brow + l] += gl * dts[l] * ur[l];
                    dC[brow + l] += (double)gr[l] * hs[l];
                    adj = dec[l] * gl;
                }
            }
            for (int l = 0; l < L; ++l) {
                double x = (double)dr[l] + bs;
                double v = ddt[l];
                if (softplus && x <= 20.0) v = v / (1.0 + exp(-x));
                ddelta[row + l] = v;
                if (dbias) dbias[d] += v;
            }
        }
    free(hs); free(dts); free(dec); free(ddt);
}

/* x: (B, C, H, W) -> xs: (B, 4, C, H*W) */
void vmasr_ref_cross_scan(const float *x, float *xs, int B, int C, int H, int W)
{
    const size_t L = (size_t)H * W;
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            const float *src = x + ((size_t)b * C + c) * L;
            float *d0 = xs + (((size_t)b * 4 + 0) * C + c) * L;
            float *d1 = xs + (((size_t)b * 4 + 1) * C + c) * L;
            float *d2 = xs + (((size_t)b * 4 + 2) * C + c) * L;
            float *d3 = xs + (((size_t)b * 4 + 3) * C + c) * L;
            for (int h = 0; h < H; ++h)
                for (int w = 0; w < W; ++w) {
                    const float v = src[(size_t)h * W + w];
                    const size_t lr = (size_t)h * W + w, lc = (size_t)w * H + h;
                    d0[lr] = v; d1[lc] = v; d2[L - 1 - lr] = v; d3[L - 1 - lc] = v;
                }
        }
}

/* ys: (B, 4, C, H*W) -> y: (B, C, H*W); association (ys0+flip ys2) + T(ys1+flip ys3), vmamba.py:55-60 */
void vmasr_ref_cross_merge(const float *ys, float *y, int B, int C, int H, int W)
{
    const size_t L = (size_t)H * W;
    for (int b = 0; b < B; ++b)
        for (int c = 0; c < C; ++c) {
            const float *s0 = ys + (((size_t)b * 4 + 0) * C + c) * L;
            const float *s1 = ys + (((size_t)b * 4 + 1) * C + c) * L;
            const float *s2 = ys + (((size_t)b * 4 + 2) * C + c) * L;
            const float *s3 = ys + (((size_t)b * 4 + 3) * C + c) * L;
            float *dst = y + ((size_t)b * C + c) * L;
            for (int h = 0; h < H; ++h)
                for (int w = 0; w < W; ++w) {
                    const size_t lr = (size_t)h * W + w, lc = (size_t)w * H + h;
                    volatile float rowp = s0[lr] + s2[L - 1 - lr];
                    volatile float colp = s1[lc] + s3[L - 1 - lc];
                    dst[lr] = rowp + colp;
                }
        }
}
