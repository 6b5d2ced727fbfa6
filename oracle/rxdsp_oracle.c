/*
 * ORACLE (test infrastructure, NOT product code) — plain-C restatement of the reference's two
 * serial receiver-DSP loops, double precision, single-threaded.
 *
 *   oracle_core_adapt_eq : optic/dsp/equalization.py:354-516 (coreAdaptEq) with the tap updates
 *                          cmaUp :789-843, rdeUp :847-909, dardeUp :913-973, nlmsUp :520-572,
 *                          ddlmsUp :648-708 and the 'static' branch :505-506
 *   oracle_bps           : optic/dsp/carrierRecovery.py:172-223 (bps)
 *
 * Pinned against the reference's own numba implementation through tests/golden/ (see
 * tests/golden/make_golden.py and tests/test_oracle_golden.py).  Only tests/, smoke() and
 * bench.py's CPU-baseline legs may load this library.
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC; no -ffast-math)
 */
#define _GNU_SOURCE
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef double complex cplx;

enum { ALG_CMA = 0, ALG_RDE = 1, ALG_NLMS = 2, ALG_DDLMS = 3, ALG_DARDE = 4, ALG_STATIC = 5 };

/* x: (nSamp, nModes) row-major; ref: (L, nModes); H, Hwl: (nModes^2, nTaps) in/out;
 * y: (L, nModes); errSq: (nModes, L); Hiter: NULL or (nModes^2, nTaps, L).
 * Tap layout H[(m + n*nModes)*nTaps + t]: tap t, input mode n -> output mode m (:467).
 * Returns 0, or 1 on an unknown algorithm (:507-510). */
int oracle_core_adapt_eq(const cplx* x, const cplx* ref, cplx* H, cplx* Hwl, cplx* y, double* errSq,
                         cplx* Hiter, int64_t L, int nModes, int nTaps, int SpS, int alg, double mu,
                         const cplx* constSymb, int M, int runWL) {
    if (alg < ALG_CMA || alg > ALG_STATIC) return 1;
    /* radii: R_cma = mean|c|^4 / mean|c|^2 (:453-455); R_rde = unique(|c|) ascending (:456) */
    double m4 = 0.0, m2 = 0.0;
    double* radii = (double*)malloc(sizeof(double) * (size_t)(M > 0 ? M : 1));
    int nR = 0;
    for (int c = 0; c < M; ++c) {
        double a = cabs(constSymb[c]);
        m4 += a * a * a * a;
        m2 += a * a;
        int k = 0;
        while (k < nR && radii[k] != a) ++k;
        if (k == nR) radii[nR++] = a;
    }
    for (int i = 1; i < nR; ++i) { /* insertion sort */
        double v = radii[i];
        int j = i - 1;
        while (j >= 0 && radii[j] > v) { radii[j + 1] = radii[j]; --j; }
        radii[j + 1] = v;
    }
    const double Rcma = (M > 0) ? (m4 / M) / (m2 / M) : 0.0;
    cplx* out = (cplx*)malloc(sizeof(cplx) * (size_t)nModes);
    cplx* g = (cplx*)malloc(sizeof(cplx) * (size_t)nModes);

    for (int64_t ind = 0; ind < L; ++ind) {
        const cplx* win = x + ind * SpS * nModes; /* rows ind*SpS .. ind*SpS+nTaps-1 (:461) */
        for (int m = 0; m < nModes; ++m) {
            cplx acc = 0.0;
            for (int n = 0; n < nModes; ++n) {
                const cplx* h = H + (size_t)(m + n * nModes) * nTaps;
                for (int t = 0; t < nTaps; ++t) acc += h[t] * win[t * nModes + n]; /* :466-468 */
                if (runWL) {
                    const cplx* hw = Hwl + (size_t)(m + n * nModes) * nTaps;
                    for (int t = 0; t < nTaps; ++t) acc += hw[t] * conj(win[t * nModes + n]); /* :470 */
                }
            }
            out[m] = acc;
            y[ind * nModes + m] = acc; /* :473 */
        }
        for (int m = 0; m < nModes; ++m) {
            const double a2 = creal(out[m]) * creal(out[m]) + cimag(out[m]) * cimag(out[m]);
            double e;
            switch (alg) {
                case ALG_CMA: /* :826-829 */
                    e = Rcma - a2;
                    g[m] = e * out[m];
                    errSq[(size_t)m * L + ind] = e * e;
                    break;
                case ALG_RDE: { /* :887-894, nearest radius, first index on ties */
                    double r = sqrt(a2), best = fabs(radii[0] - r), Rd = radii[0];
                    for (int i = 1; i < nR; ++i) {
                        double d = fabs(radii[i] - r);
                        if (d < best) { best = d; Rd = radii[i]; }
                    }
                    e = Rd * Rd - a2;
                    g[m] = e * out[m];
                    errSq[(size_t)m * L + ind] = e * e;
                } break;
                case ALG_DARDE: { /* :953-959 */
                    double Rd = cabs(ref[ind * nModes + m]);
                    e = Rd * Rd - a2;
                    g[m] = e * out[m];
                    errSq[(size_t)m * L + ind] = e * e;
                } break;
                case ALG_NLMS: /* :556 */
                    g[m] = ref[ind * nModes + m] - out[m];
                    errSq[(size_t)m * L + ind] = creal(g[m]) * creal(g[m]) + cimag(g[m]) * cimag(g[m]);
                    break;
                case ALG_DDLMS: { /* :688-691 */
                    int bi = 0;
                    double best = cabs(out[m] - constSymb[0]);
                    for (int c = 1; c < M; ++c) {
                        double d = cabs(out[m] - constSymb[c]);
                        if (d < best) { best = d; bi = c; }
                    }
                    g[m] = constSymb[bi] - out[m];
                    errSq[(size_t)m * L + ind] = creal(g[m]) * creal(g[m]) + cimag(g[m]) * cimag(g[m]);
                } break;
                default: /* static (:505-506); first entry undefined in the reference, 0 here */
                    g[m] = 0.0;
                    errSq[(size_t)m * L + ind] = ind > 0 ? errSq[(size_t)m * L + ind - 1] : 0.0;
            }
        }
        if (alg != ALG_STATIC) {
            for (int n = 0; n < nModes; ++n) {
                double inv = 1.0;
                if (alg == ALG_NLMS) { /* :563  x / ||x||^2 */
                    double s = 0.0;
                    for (int t = 0; t < nTaps; ++t) {
                        cplx v = win[t * nModes + n];
                        s += creal(v) * creal(v) + cimag(v) * cimag(v);
                    }
                    inv = 1.0 / s;
                }
                for (int m = 0; m < nModes; ++m) {
                    cplx* h = H + (size_t)(m + n * nModes) * nTaps;
                    const cplx w = mu * g[m];
                    for (int t = 0; t < nTaps; ++t) h[t] += w * conj(win[t * nModes + n] * inv); /* :838-840 */
                    if (runWL) {
                        cplx* hw = Hwl + (size_t)(m + n * nModes) * nTaps;
                        for (int t = 0; t < nTaps; ++t) hw[t] += w * (win[t * nModes + n] * inv); /* :842 */
                    }
                }
            }
        }
        if (Hiter) { /* :511-512 */
            for (int r = 0; r < nModes * nModes; ++r)
                for (int t = 0; t < nTaps; ++t) Hiter[((size_t)r * nTaps + t) * L + ind] = H[(size_t)r * nTaps + t];
        }
    }
    free(out);
    free(g);
    free(radii);
    return 0;
}

/* x: (L, nModes) row-major complex128; idx/phase: (L, nModes).  Zero-pad N symbols at both ends
 * (:203-206); dmin[b][k] = min_c |x_k e^{j phi_b} - c|^2 (:216-217); output k = argmin_b of the
 * centred (2N+1)-window sum (:218-221), first index on ties. */
void oracle_bps(const cplx* x, int64_t L, int nModes, const cplx* constSymb, int M, int B, int N,
                int32_t* idx, double* phase) {
    const int W = 2 * N + 1;
    cplx* rot = (cplx*)malloc(sizeof(cplx) * (size_t)B);
    double* ph = (double*)malloc(sizeof(double) * (size_t)B);
    double* ring = (double*)malloc(sizeof(double) * (size_t)B * W);
    for (int b = 0; b < B; ++b) {
        ph[b] = ((double)b * (M_PI / 2.0)) / (double)B; /* :199 */
        rot[b] = cos(ph[b]) + I * sin(ph[b]);
    }
    for (int n = 0; n < nModes; ++n) {
        memset(ring, 0, sizeof(double) * (size_t)B * W);
        for (int64_t k = 0; k < L + 2 * N; ++k) { /* k indexes the padded sequence */
            const int64_t src = k - N;
            const cplx v = (src >= 0 && src < L) ? x[src * nModes + n] : 0.0;
            const int slot = (int)(k % W);
            for (int b = 0; b < B; ++b) {
                const cplx z = v * rot[b];
                double best = INFINITY;
                for (int c = 0; c < M; ++c) {
                    const double dr = creal(z) - creal(constSymb[c]), di = cimag(z) - cimag(constSymb[c]);
                    const double d = dr * dr + di * di;
                    if (d < best) best = d;
                }
                ring[(size_t)b * W + slot] = best;
            }
            if (k >= 2 * N) {
                int bi = 0;
                double bs = INFINITY;
                for (int b = 0; b < B; ++b) {
                    /* sum the window in time order (oldest first) */
                    double s = 0.0;
                    for (int t = 1; t <= W; ++t) s += ring[(size_t)b * W + (slot + t) % W];
                    if (s < bs) { bs = s; bi = b; }
                }
                idx[(k - 2 * N) * nModes + n] = bi;
                phase[(k - 2 * N) * nModes + n] = ph[bi];
            }
        }
    }
    free(rot);
    free(ph);
    free(ring);
}
