// Fused four-step FFT kernels of the split-step propagators (the sm_100a fast path, N = N1*N2,
// N1 = 32*Q1, N2 = 32*Q2, Q in {8,16,32}).
//
// A length-N transform along a planar row x[n], n = N2*n1 + n2, is split as
//     column pass : N1-point FFT over n1 for every column n2, times W_N^{n2 k1}   -> W[k1][n2]
//     row pass    : N2-point FFT over n2 for every row k1                          -> X[k1 + N1 k2]
// The inverse mirrors it.  Because the linear operator is diagonal in frequency and the Kerr
// rotation is pointwise in time, one half step  ifft(fft(.)*L)  costs THREE passes over the data:
//     k_col (… -> FFT_N1 -> twiddle)   k_row (FFT_N2 · L · IFFT_N2)   k_col (twiddle* -> IFFT_N1 -> …)
// and every pointwise operation of the split-step loop (power, nonlinear phase, rotation,
// convergence sums, max power) rides in the time-domain end of a column kernel.
//
// Reference formulas restated here: optic/models/channels.py:388-390, 406-421, 424, 436, 493,
// 517-519 (manakovSSF), :219-229 (ssfm), optic/dsp/equalization.py:1077, 1129 (DBP signs).
#pragma once
#include "fft_core.cuh"
#include "ssfm_kernels.cuh"

namespace ocb {

enum ColMode { COL_FWD = 0, COL_INV = 1, COL_FIRST = 2, COL_ITER = 3, COL_NLSE = 4 };

struct ColArgs {
    const float2* in;     // COL_FWD: time-domain field ; others: W-domain buffer
    float2* out;          // COL_INV: time-domain field ; others: W-domain buffer
    const float2* aux0;   // COL_FIRST: step-start field Ech ; COL_ITER: previous iterate E_conv
    float2* aux1;         // COL_FIRST: E_hd (written)       ; COL_ITER: new iterate (written)
    const float2* ehd;    // COL_ITER: E_hd (read)
    float* pch;           // COL_FIRST: written ; COL_ITER: read
    const float2* tw;     // [32][Q1]  exp(-2 pi i q ka / N1)
    const float2* tabV;   // [N2][32]  exp(-2 pi i n2 ka / N)
    const float2* tabU;   // [N2][Q1]  exp(-2 pi i n2 32 kq / N)
    double* partials;     // COL_ITER reduction scratch
    double* sums;
    unsigned* ticket;
    int64_t N;            // samples per row
    int N2;               // columns
    float cphi;           // dir*hz*(8/9)γ (FIRST) | dir*hz*(8/9)γ/2 (ITER) | γ hz (NLSE)
    float out_scale;      // COL_INV: extra gain applied to the time-domain output
};

// phase rotation exp(j ph): short polynomial for the small per-step phases (|ph| < 0.5 rad,
// truncation error < 1e-9), libm sincosf otherwise.
__device__ __forceinline__ float2 phase_rot(float ph) {
    float s, c;
    if (fabsf(ph) < 0.5f) {
        const float x2 = ph * ph;
        s = ph * fmaf(x2, fmaf(x2, fmaf(x2, -1.9841270e-4f, 8.3333333e-3f), -1.6666667e-1f), 1.0f);
        c = fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, 2.4801587e-5f, -1.3888889e-3f), 4.1666667e-2f), -0.5f), 1.0f);
    } else {
        sincosf(ph, &s, &c);
    }
    return make_float2(c, s);
}

// ------------------------------------------------------------------------------------------
// Column kernel.  One CTA = one tile of C adjacent columns, both polarisations (NP = 2) or one
// (NP = 1).  Thread (pol, q, c): time-domain rows Q1*a' + q (a' < 32), frequency rows
// (q*G + g) + 32*kq.  A warp touches 32/C adjacent rows x C*8 contiguous bytes per access.
// ------------------------------------------------------------------------------------------
template <int Q1, int C, int NP, int MODE>
__global__ void __launch_bounds__(NP* Q1* C, (NP * Q1 * C <= 256) ? 2 : 1)
k_col(const ColArgs A) {
    using namespace fft;
    constexpr int G = 32 / Q1, STR = Q1 * C + C, NT = NP * Q1 * C;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tw = reinterpret_cast<float2*>(smem_raw);              // [32*Q1]
    float2* Vs = tw + 32 * Q1;                                      // [C][32]
    float2* Us = Vs + C * 32;                                       // [C][Q1]
    float* xbuf = reinterpret_cast<float*>(Us + C * Q1);            // [NP][2][32*STR]

    const int tid = threadIdx.x;
    const int pol = tid / (Q1 * C);
    const int q = (tid % (Q1 * C)) / C;
    const int c = tid % C;
    const int col0 = blockIdx.x * C;
    const int N2 = A.N2;
    const int64_t rowoff = (int64_t)pol * A.N + col0 + c;
    float* xr = xbuf + (size_t)pol * 2 * 32 * STR;
    float* xi = xr + 32 * STR;

    for (int i = tid; i < 32 * Q1; i += NT) tw[i] = A.tw[i];
    for (int i = tid; i < C * 32; i += NT) Vs[i] = A.tabV[(int64_t)(col0 + i / 32) * 32 + (i % 32)];
    for (int i = tid; i < C * Q1; i += NT) Us[i] = A.tabU[(int64_t)(col0 + i / Q1) * Q1 + (i % Q1)];
    __syncthreads();
    auto bsync = [] { __syncthreads(); };

    float2 v[32];

    // ---- enter: either the time-domain field (FWD) or the W-domain buffer (inverse first) ----
    if constexpr (MODE == COL_FWD) {
        const float2* src = A.in + rowoff;
#pragma unroll
        for (int a = 0; a < 32; ++a) v[a] = __ldg(src + (int64_t)(Q1 * a + q) * N2);
    } else {
        const float2* src = A.in + rowoff;
        static_for<0, G>([&](auto gg) {
            constexpr int GI = decltype(gg)::value;
            static_for<0, Q1>([&](auto kk) {
                constexpr int KQ = decltype(kk)::value, SLOT = GI * Q1 + brev<Q1>(KQ);
                v[SLOT] = __ldg(src + (int64_t)((q * G + GI) + 32 * KQ) * N2);
            });
        });
        static_for<0, G>([&](auto gg) {
            constexpr int GI = decltype(gg)::value;
            const float2 wv = Vs[c * 32 + q * G + GI];
            static_for<0, Q1>([&](auto kk) {
                constexpr int KQ = decltype(kk)::value, SLOT = GI * Q1 + brev<Q1>(KQ);
                const float2 w = cmul(wv, Us[c * Q1 + KQ]);
                v[SLOT] = cmul_conj(v[SLOT], w);  // conj twiddle W_N^{-n2 k1}
            });
        });
        coop_fft_inverse<Q1, C, C>(v, xr, xi, tw, q, c, bsync);  // v[a'] = field at row Q1*a' + q
    }

    // ---- time-domain work -------------------------------------------------------------------
    float s_num = 0.f, s_den = 0.f, s_max = 0.f;
    if constexpr (MODE == COL_INV) {
        float2* dst = A.out + rowoff;
#pragma unroll
        for (int a = 0; a < 32; ++a)
            dst[(int64_t)(Q1 * a + q) * N2] = make_float2(v[a].x * A.out_scale, v[a].y * A.out_scale);
        return;
    }
    if constexpr (MODE == COL_NLSE) {  // channels.py:225
#pragma unroll
        for (int a = 0; a < 32; ++a) v[a] = cmul(v[a], phase_rot(A.cphi * cabs2(v[a])));
        __syncthreads();  // the forward transform below reuses the exchange buffers
    }
    if constexpr (MODE == COL_FIRST || MODE == COL_ITER) {
        static_assert(NP == 2, "Manakov modes need both polarisations in the CTA");
        __syncthreads();  // exchange buffers are free again; reuse xr[pol] to share |E|² across pols
        if constexpr (MODE == COL_FIRST) {
            // v = E_hd (store it); power of the step-start field Ech  (channels.py:388)
            float2* ehd_out = A.aux1 + rowoff;
            const float2* ech = A.aux0 + rowoff;
#pragma unroll
            for (int a = 0; a < 32; ++a) {
                const int64_t o = (int64_t)(Q1 * a + q) * N2;
                ehd_out[o] = v[a];
                xr[a * STR + q * C + c] = cabs2(__ldg(ech + o));
            }
        } else {
            // v = E_fd: convergence sums against the previous iterate, store as the new iterate
            const float2* ec = A.aux0 + rowoff;
            float2* ec_new = A.aux1 + rowoff;
#pragma unroll
            for (int a = 0; a < 32; ++a) {
                const int64_t o = (int64_t)(Q1 * a + q) * N2;
                const float2 e = __ldg(ec + o);
                s_num += cabs2(make_float2(v[a].x - e.x, v[a].y - e.y));  // channels.py:517
                s_den += cabs2(e);
                ec_new[o] = v[a];
                xr[a * STR + q * C + c] = cabs2(v[a]);
            }
        }
        __syncthreads();
        const float* other = xbuf + (size_t)(1 - pol) * 2 * 32 * STR;
        float* pch = A.pch + col0 + c;
        const float2* ehd = (MODE == COL_ITER) ? A.ehd + rowoff : nullptr;
#pragma unroll
        for (int a = 0; a < 32; ++a) {
            const int64_t o = (int64_t)(Q1 * a + q) * N2;
            const float P = xr[a * STR + q * C + c] + other[a * STR + q * C + c];
            float ph;
            if constexpr (MODE == COL_FIRST) {
                if (pol == 0) pch[o] = P;
                ph = A.cphi * P;  // φ = (8/9)γ(P+P)/2, channels.py:390/493 with E_conv == Ech
            } else {
                s_max = fmaxf(s_max, P);
                ph = A.cphi * (__ldg(pch + o) + P);  // channels.py:436
                v[a] = __ldg(ehd + o);
            }
            v[a] = cmul(v[a], phase_rot(ph));  // channels.py:414-417
        }
        __syncthreads();  // before the forward transform reuses the exchange buffers
    }

    // ---- leave: forward column FFT + inter-pass twiddle -> W-domain buffer ----------------------
    coop_fft_forward<Q1, C, C>(v, xr, xi, tw, q, c, bsync);
    {
        float2* dst = A.out + rowoff;
        static_for<0, G>([&](auto gg) {
            constexpr int GI = decltype(gg)::value;
            const float2 wv = Vs[c * 32 + q * G + GI];
            static_for<0, Q1>([&](auto kk) {
                constexpr int KQ = decltype(kk)::value, SLOT = GI * Q1 + brev<Q1>(KQ);
                const float2 w = cmul(wv, Us[c * Q1 + KQ]);
                dst[(int64_t)((q * G + GI) + 32 * KQ) * N2] = cmul(v[SLOT], w);
            });
        });
    }
    if constexpr (MODE == COL_ITER) {
        if (pol == 1) s_max = 0.f;  // both pol threads saw the same total power
        block_reduce3_finalize(s_num, s_den, s_max, A.partials, A.sums, A.ticket);
    }
}

// ------------------------------------------------------------------------------------------
// Row kernel: for every row k1 of the W-domain buffer:  FFT_N2 -> x LP[k1][.] -> IFFT_N2, in place.
// Q2 threads own one row; a 256-thread CTA processes 256/Q2 rows per trip (grid-stride).
// LP is the linear operator in the kernel's own consumption order:
//   LP[k1*N2 + s*Q2 + t] = scale * exp((a + j b ω_k²) h),  k = k1 + N1*(ka + 32 kq),
//   ka = t*G + s/Q2, kq = brev<Q2>(s % Q2)
// ------------------------------------------------------------------------------------------
template <int Q2>
__global__ void __launch_bounds__(256)
k_row(float2* __restrict__ W, const float2* __restrict__ LP, const float2* __restrict__ tw_g, int N1,
      int64_t n_rows) {
    using namespace fft;
    constexpr int S2 = 32 * Q2, STR = Q2 + 1, RPB = 256 / Q2;
    constexpr int GBUF = 32 * STR + (Q2 < 32 ? Q2 : 0);  // stagger groups sharing a warp
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* tw = reinterpret_cast<float2*>(smem_raw);   // [32*Q2]
    float* xbuf = reinterpret_cast<float*>(tw + 32 * Q2);
    const int tid = threadIdx.x, grp = tid / Q2, q = tid % Q2;
    float* xr = xbuf + (size_t)grp * 2 * GBUF;
    float* xi = xr + GBUF;
    for (int i = tid; i < 32 * Q2; i += 256) tw[i] = tw_g[i];
    __syncthreads();
    auto wsync = [] { __syncwarp(); };

    for (int64_t row0 = (int64_t)blockIdx.x * RPB; row0 < n_rows; row0 += (int64_t)gridDim.x * RPB) {
        const int64_t row = row0 + grp;
        if (row >= n_rows) continue;  // n_rows is a multiple of RPB in practice (whole warps stay together)
        float2* p = W + row * S2;
        const float2* lp = LP + (row % N1) * S2 + q;
        float2 v[32];
#pragma unroll
        for (int a = 0; a < 32; ++a) v[a] = p[Q2 * a + q];
        coop_fft_forward<Q2, 1, 1>(v, xr, xi, tw, q, 0, wsync);
#pragma unroll
        for (int s = 0; s < 32; ++s) v[s] = cmul(v[s], __ldg(lp + s * Q2));
        __syncwarp();
        coop_fft_inverse<Q2, 1, 1>(v, xr, xi, tw, q, 0, wsync);
#pragma unroll
        for (int a = 0; a < 32; ++a) p[Q2 * a + q] = v[a];
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------
// Table builders (float64 math, rounded once to float32).
// ------------------------------------------------------------------------------------------
// tw[ka*Q + q] = exp(-2 pi i q ka / (32 Q))
__global__ void k_tab_tw(float2* tw, int Q) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 32 * Q) return;
    const int ka = i / Q, q = i % Q;
    double s, c;
    sincospi(-2.0 * (double)(q * ka) / (double)(32 * Q), &s, &c);
    tw[i] = make_float2((float)c, (float)s);
}
// V[n2*32 + ka] = exp(-2 pi i n2 ka / N) ; U[n2*Q1 + kq] = exp(-2 pi i n2 32 kq / N)
__global__ void k_tab_inter(float2* V, float2* U, int N2, int Q1, int64_t N) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < (int64_t)N2 * 32) {
        const int64_t n2 = i / 32, ka = i % 32;
        double s, c;
        sincospi(-2.0 * (double)((n2 * ka) % N) / (double)N, &s, &c);
        V[i] = make_float2((float)c, (float)s);
    }
    if (i < (int64_t)N2 * Q1) {
        const int64_t n2 = i / Q1, kq = i % Q1;
        double s, c;
        sincospi(-2.0 * (double)((n2 * 32 * kq) % N) / (double)N, &s, &c);
        U[i] = make_float2((float)c, (float)s);
    }
}
// Linear operator in the row kernel's consumption order (see k_row).  a, b, Fs, h, scale as in
// k_linop_table: value = scale * exp((a + j b ω_k²) h), ω_k = 2π Fs fftfreq(N)[k].
template <int Q2>
__global__ void k_tab_linop_perm(float2* __restrict__ LP, int N1, int64_t N, double a, double b, double Fs,
                                 double h, double scale) {
    using namespace fft;
    constexpr int S2 = 32 * Q2, G = 32 / Q2;
    const double two_pi = 6.283185307179586476925286766559;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k1 = i / S2;
        const int r = (int)(i % S2), s = r / Q2, t = r % Q2;
        const int ka = t * G + s / Q2;
        int kq = 0, pp = s % Q2;
        for (int bit = 0; bit < ilog2(Q2); ++bit) kq |= ((pp >> bit) & 1) << (ilog2(Q2) - 1 - bit);
        const int64_t k = k1 + (int64_t)N1 * (ka + 32 * kq);
        const int64_t kk = (k <= (N - 1) / 2) ? k : k - N;
        const double w = two_pi * Fs * ((double)kk / (double)N);
        const double amp = scale * exp(a * h);
        double sn, cs;
        sincos(b * (w * w) * h, &sn, &cs);
        LP[i] = make_float2((float)(amp * cs), (float)(amp * sn));
    }
}

}  // namespace ocb
