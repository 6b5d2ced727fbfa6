// In-register radix-2 FFTs (size 8/16/32, fully unrolled, compile-time twiddles) and the
// "32 x Q" cooperative FFT of size S = 32 Q built from them: every thread owns 32 complex
// samples; Q threads own one length-S transform; one shared-memory exchange per transform.
//
//   forward (DIF):  natural-order input  -> bit-reversed output
//   inverse (DIT):  bit-reversed input   -> natural-order output
// so a forward/inverse pair never needs an explicit permutation.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace ocb {
namespace fft {

template <int I, int N, typename F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

// cos(2 pi k / 32), k = 0..8, correctly rounded to float
__host__ __device__ constexpr float cos32_q(int k) {
    constexpr float t[9] = {1.0f,           0.98078528040323f, 0.92387953251129f,
                            0.83146961230255f, 0.70710678118655f, 0.55557023301960f,
                            0.38268343236509f, 0.19509032201613f, 0.0f};
    return t[k];
}
// cos / sin of 2 pi k / 32 for any integer k (symmetry-reduced to the table above)
__host__ __device__ constexpr float cos32(int k) {
    k = ((k % 32) + 32) % 32;
    if (k <= 8) return cos32_q(k);
    if (k <= 16) return -cos32_q(16 - k);
    if (k <= 24) return -cos32_q(k - 16);
    return cos32_q(32 - k);
}
__host__ __device__ constexpr float sin32(int k) { return cos32(k - 8); }

__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }
template <int R>
__host__ __device__ constexpr int brev(int k) {
    int r = 0;
    for (int b = 0; b < ilog2(R); ++b) r |= ((k >> b) & 1) << (ilog2(R) - 1 - b);
    return r;
}

// t * exp(DIR * 2 pi i * J / LEN), J and LEN compile-time (trivial factors cost no multiply)
template <int LEN, int J, int DIR>
__device__ __forceinline__ float2 twiddle_mul(float2 t) {
    constexpr int K = (J * (32 / LEN)) % 32;  // index on the 32-point circle
    if constexpr (K == 0) {
        return t;
    } else if constexpr (K == 8) {  // exp(DIR i pi/2) = DIR*i
        return DIR > 0 ? make_float2(-t.y, t.x) : make_float2(t.y, -t.x);
    } else if constexpr (K == 16) {
        return make_float2(-t.x, -t.y);
    } else if constexpr (K == 24) {
        return DIR > 0 ? make_float2(t.y, -t.x) : make_float2(-t.y, t.x);
    } else {
        constexpr float c = cos32(K), s = (DIR > 0 ? 1.f : -1.f) * sin32(K);
        return make_float2(fmaf(t.x, c, -t.y * s), fmaf(t.x, s, t.y * c));
    }
}

// decimation in frequency: v natural -> X[k] at v[brev<R>(k)]
template <int R, int DIR>
__device__ __forceinline__ void fft_dif(float2* v) {
    static_for<0, ilog2(R)>([&](auto st) {
        constexpr int LEN = R >> decltype(st)::value, HALF = LEN / 2;
        static_for<0, R / LEN>([&](auto grp) {
            constexpr int BASE = decltype(grp)::value * LEN;
            static_for<0, HALF>([&](auto jj) {
                constexpr int J = decltype(jj)::value;
                const float2 a = v[BASE + J], b = v[BASE + J + HALF];
                v[BASE + J] = make_float2(a.x + b.x, a.y + b.y);
                v[BASE + J + HALF] = twiddle_mul<LEN, J, DIR>(make_float2(a.x - b.x, a.y - b.y));
            });
        });
    });
}

// decimation in time: v[brev<R>(n)] = x[n] on input -> X[k] at v[k]
template <int R, int DIR>
__device__ __forceinline__ void fft_dit(float2* v) {
    static_for<0, ilog2(R)>([&](auto st) {
        constexpr int LEN = 2 << decltype(st)::value, HALF = LEN / 2;
        static_for<0, R / LEN>([&](auto grp) {
            constexpr int BASE = decltype(grp)::value * LEN;
            static_for<0, HALF>([&](auto jj) {
                constexpr int J = decltype(jj)::value;
                const float2 a = v[BASE + J];
                const float2 b = twiddle_mul<LEN, J, DIR>(v[BASE + J + HALF]);
                v[BASE + J] = make_float2(a.x + b.x, a.y + b.y);
                v[BASE + J + HALF] = make_float2(a.x - b.x, a.y - b.y);
            });
        });
    });
}

// ------------------------------------------------------------------------------------------
// Cooperative transform of size S = 32*Q.  Thread q in [0,Q) of a group owns x[Q*a + q], a<32.
// Index algebra (k = ka + 32 kq):
//   X[ka + 32 kq] = sum_q W_Q^{q kq} [ W_S^{q ka} sum_a x[Q a + q] W_32^{a ka} ]
// After the forward transform thread t owns the G = 32/Q values ka = t*G + g (g < G) and all kq:
//   u[g*Q + brev<Q>(kq)] = X[ka + 32 kq].
// The inverse consumes exactly that arrangement and returns x'[Q a' + q'] in v[a'] of thread q'.
//
// Exchange buffer: planar float arrays xr/xi of size 32*(Q*CP + PAD) per group, element
// (ka, q, c) at ka*(Q*CP + PAD) + q*CP + c, where CP = number of column-threads interleaved
// (1 for the row kernel) — conflict-free for both access directions.
// tw: twiddle table TW[ka*Q + q] = exp(-2 pi i q ka / S), followed by its transpose TWT[q*32 + ka]
// (32*Q entries each) so that both transform directions read it lane-contiguously.
// ------------------------------------------------------------------------------------------
template <int Q>
struct Coop {
    static constexpr int S = 32 * Q;
    static constexpr int G = 32 / Q;
};

// HALF = true: the exchange runs in two rounds (real parts, then imaginary parts) through ONE planar array
// (xr; xi unused), which halves the shared-memory footprint at the price of two more barriers per transform.
template <int Q, int CP, int PAD, bool HALF = false, typename SyncF>
__device__ __forceinline__ void coop_fft_forward(float2* v, float* xr, float* xi, const float2* __restrict__ tw,
                                                 int q, int c, SyncF&& sync) {
    constexpr int G = 32 / Q, STR = Q * CP + PAD;
    fft_dif<32, -1>(v);  // v[brev5(ka)] = Z[ka]
    if constexpr (!HALF) {
        static_for<0, 32>([&](auto kk) {
            constexpr int KA = decltype(kk)::value, SLOT = brev<32>(KA);
            float2 z = v[SLOT];
            if constexpr (KA != 0) z = cmul(z, tw[KA * Q + q]);
            xr[KA * STR + q * CP + c] = z.x;
            xi[KA * STR + q * CP + c] = z.y;
        });
        sync();
        // thread t = q now gathers ka = t*G + g, all j: u[g*Q + j] = Z[ka][j]
        static_for<0, G>([&](auto gg) {
            constexpr int GI = decltype(gg)::value;
            const int ka = q * G + GI;
            static_for<0, Q>([&](auto jj) {
                constexpr int J = decltype(jj)::value;
                v[GI * Q + J] = make_float2(xr[ka * STR + J * CP + c], xi[ka * STR + J * CP + c]);
            });
        });
    } else {
        static_for<0, 32>([&](auto kk) {
            constexpr int KA = decltype(kk)::value, SLOT = brev<32>(KA);
            if constexpr (KA != 0) v[SLOT] = cmul(v[SLOT], tw[KA * Q + q]);
            xr[KA * STR + q * CP + c] = v[SLOT].x;
        });
        sync();
        float re[32];
        static_for<0, G>([&](auto gg) {
            constexpr int GI = decltype(gg)::value;
            const int ka = q * G + GI;
            static_for<0, Q>([&](auto jj) { constexpr int J = decltype(jj)::value; re[GI * Q + J] = xr[ka * STR + J * CP + c]; });
        });
        sync();
        static_for<0, 32>([&](auto kk) {
            constexpr int KA = decltype(kk)::value, SLOT = brev<32>(KA);
            xr[KA * STR + q * CP + c] = v[SLOT].y;
        });
        sync();
        static_for<0, G>([&](auto gg) {
            constexpr int GI = decltype(gg)::value;
            const int ka = q * G + GI;
            static_for<0, Q>([&](auto jj) {
                constexpr int J = decltype(jj)::value;
                v[GI * Q + J] = make_float2(re[GI * Q + J], xr[ka * STR + J * CP + c]);
            });
        });
    }
    static_for<0, G>([&](auto gg) { fft_dif<Q, -1>(v + decltype(gg)::value * Q); });
}

// twt: the transposed copy of the table, entry [J][ka] (defaults to tw + 32*Q; for Q = 32 the table is
// symmetric and twt may alias tw).
template <int Q, int CP, int PAD, bool HALF = false, typename SyncF>
__device__ __forceinline__ void coop_fft_inverse(float2* v, float* xr, float* xi, const float2* __restrict__ tw,
                                                 int q, int c, SyncF&& sync, const float2* __restrict__ twt = nullptr) {
    constexpr int G = 32 / Q, STR = Q * CP + PAD;
    if (twt == nullptr) twt = tw + 32 * Q;
    static_for<0, G>([&](auto gg) { fft_dit<Q, +1>(v + decltype(gg)::value * Q); });  // over kq -> q'
    static_for<0, G>([&](auto gg) {
        constexpr int GI = decltype(gg)::value;
        const int ka = q * G + GI;
        static_for<0, Q>([&](auto jj) {
            constexpr int J = decltype(jj)::value;  // q'
            // transposed copy of the table: entry [J][ka], so that lanes (ka) are contiguous
            const float2 w = twt[J * 32 + ka];
            const float2 z = cmul_conj(v[GI * Q + J], w);  // conj twiddle for the inverse
            v[GI * Q + J] = z;
            xr[ka * STR + J * CP + c] = z.x;
            if constexpr (!HALF) xi[ka * STR + J * CP + c] = z.y;
        });
    });
    sync();
    if constexpr (!HALF) {
        static_for<0, 32>([&](auto kk) {
            constexpr int KA = decltype(kk)::value, SLOT = brev<32>(KA);
            v[SLOT] = make_float2(xr[KA * STR + q * CP + c], xi[KA * STR + q * CP + c]);
        });
    } else {
        float re[32];
        static_for<0, 32>([&](auto kk) { constexpr int KA = decltype(kk)::value; re[KA] = xr[KA * STR + q * CP + c]; });
        sync();
        static_for<0, G>([&](auto gg) {
            constexpr int GI = decltype(gg)::value;
            const int ka = q * G + GI;
            static_for<0, Q>([&](auto jj) { constexpr int J = decltype(jj)::value; xr[ka * STR + J * CP + c] = v[GI * Q + J].y; });
        });
        sync();
        static_for<0, 32>([&](auto kk) {
            constexpr int KA = decltype(kk)::value, SLOT = brev<32>(KA);
            v[SLOT] = make_float2(re[KA], xr[KA * STR + q * CP + c]);
        });
    }
    fft_dit<32, +1>(v);  // natural a'
}

}  // namespace fft
}  // namespace ocb
