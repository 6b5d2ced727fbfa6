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

// Twiddle precision (OCB_TW_MODE).  The FIXED rounding error of a float twiddle (~3e-8 relative) acts on nearly the same
// field in every split-step, so it accumulates linearly with the step count (measured 2.6e-7 per step with plain float
// twiddles, tools/drift_experiment.py).
//   0  plain float twiddles (round 1)
//   1  double-single (default): every twiddle is an unevaluated sum hi + lo of two floats (lo = float(w - hi), about
//      2^-24 |w|) and a product x*w is evaluated as x*hi + x*lo inside ONE fused chain, so that the lo term takes part in
//      the rounding of the result: +4 FMA per complex multiply.  The fixed error drops to ~1e-15 and what is left is the
//      rounding of the linear-operator table plus the data-dependent rounding of the float arithmetic: 3.9e-8 per step.
// (A cheaper variant — compensate only the modulus error, z + eps*z after the float product — was measured and does not
// work: eps*z is below half an ulp of z, so the separate correction never changes the rounded result; 1.7e-7 per step.)
#ifndef OCB_TW_MODE
#define OCB_TW_MODE 1
#endif
constexpr bool kDS = (OCB_TW_MODE == 1);
constexpr bool kLO = kDS;                   // a second table entry (the lo part) per twiddle exists

// cos(2 pi k / 32), k = 0..8, to double precision
__host__ __device__ constexpr double cos32_qd(int k) {
    constexpr double t[9] = {1.0,
                             0.98078528040323044912618223613424,
                             0.92387953251128675612818318939679,
                             0.83146961230254523707878837761791,
                             0.70710678118654752440084436210485,
                             0.55557023301960222474283081394853,
                             0.38268343236508977172845998403040,
                             0.19509032201612826784828486847702,
                             0.0};
    return t[k];
}
// cos / sin of 2 pi k / 32 for any integer k (symmetry-reduced to the table above)
__host__ __device__ constexpr double cos32d(int k) {
    k = ((k % 32) + 32) % 32;
    if (k <= 8) return cos32_qd(k);
    if (k <= 16) return -cos32_qd(16 - k);
    if (k <= 24) return -cos32_qd(k - 16);
    return cos32_qd(32 - k);
}
__host__ __device__ constexpr double sin32d(int k) { return cos32d(k - 8); }
__host__ __device__ constexpr float hi_part(double v) { return (float)v; }
__host__ __device__ constexpr float lo_part(double v) { return (float)(v - (double)(float)v); }
__host__ __device__ constexpr float cos32(int k) { return hi_part(cos32d(k)); }
__host__ __device__ constexpr float sin32(int k) { return hi_part(sin32d(k)); }

__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }
template <int R>
__host__ __device__ constexpr int brev(int k) {
    int r = 0;
    for (int b = 0; b < ilog2(R); ++b) r |= ((k >> b) & 1) << (ilog2(R) - 1 - b);
    return r;
}

// Packed FP32 pairs (sm_100a FADD2): a complex add / subtract is ONE instruction on the (re, im) register pair instead
// of two.  The butterflies of the radix-2 stages are 60 % of a transform's additions.  Same IEEE results as two FADDs.
// PK selects it per kernel: it pays in the time pass (instruction-fetch bound: 15 % fewer instructions, -6..-13 % time)
// and costs 4 % in the frequency pass (math-pipe bound: FADD2 occupies the FMA pipe like two FADDs, while scalar FADDs
// also issue to the second pipe) — measured on the B200, profiles/r2_summary.md.
#ifndef OCB_F32X2
#define OCB_F32X2 1
#endif
template <bool PK = true>
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
    if constexpr (!PK) return make_float2(a.x + b.x, a.y + b.y);
#if OCB_F32X2
    unsigned long long ua = *reinterpret_cast<unsigned long long*>(&a), ub = *reinterpret_cast<unsigned long long*>(&b), ud;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(ud) : "l"(ua), "l"(ub));
    return *reinterpret_cast<float2*>(&ud);
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
template <bool PK = true>
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
    if constexpr (!PK) return make_float2(a.x - b.x, a.y - b.y);
#if OCB_F32X2
    unsigned long long ua = *reinterpret_cast<unsigned long long*>(&a), ub = *reinterpret_cast<unsigned long long*>(&b), ud;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(ud) : "l"(ua), "l"(ub));
    return *reinterpret_cast<float2*>(&ud);
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}

// a * (h + l) and a * conj(h + l) for a double-single factor (l is ignored in mode 0)
__device__ __forceinline__ float2 cmul_ds(float2 a, float2 h, float2 l) {
    if constexpr (kDS) {
        return make_float2(fmaf(a.x, h.x, fmaf(-a.y, h.y, fmaf(a.x, l.x, -a.y * l.y))),
                           fmaf(a.x, h.y, fmaf(a.y, h.x, fmaf(a.x, l.y, a.y * l.x))));
    } else {
        return cmul(a, h);
    }
}
__device__ __forceinline__ float2 cmul_conj_ds(float2 a, float2 h, float2 l) {
    if constexpr (kDS) {
        return make_float2(fmaf(a.x, h.x, fmaf(a.y, h.y, fmaf(a.x, l.x, a.y * l.y))),
                           fmaf(a.y, h.x, fmaf(-a.x, h.y, fmaf(a.y, l.x, -a.x * l.y))));
    } else {
        return cmul_conj(a, h);
    }
}
// inter-pass twiddle x * V * U (or x * conj(V U)) from its two table factors: in double-single mode the factors are
// applied one after the other (a float product V*U would carry its own fixed rounding error)
template <bool CONJ>
__device__ __forceinline__ float2 cmul_vu(float2 x, float2 vh, float2 vl, float2 uh, float2 ul) {
    if constexpr (kDS) {
        return CONJ ? cmul_conj_ds(cmul_conj_ds(x, vh, vl), uh, ul) : cmul_ds(cmul_ds(x, vh, vl), uh, ul);
    } else {
        return CONJ ? cmul_conj(x, cmul(vh, uh)) : cmul(x, cmul(vh, uh));
    }
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
    } else if constexpr (kLO && (K == 4 || K == 12)) {
        // exp(DIR i pi/4) = r (1 + DIR i), exp(DIR 3 i pi/4) = r (-1 + DIR i), r = sqrt(1/2) = rh + rl
        constexpr float rh = hi_part(cos32d(4)), rl = lo_part(cos32d(4));
        constexpr float d = (DIR > 0 ? 1.f : -1.f);
        const float a = (K == 4) ? (t.x - d * t.y) : (-t.x - d * t.y);
        const float b = (K == 4) ? (t.y + d * t.x) : (-t.y + d * t.x);
        return make_float2(fmaf(a, rh, a * rl), fmaf(b, rh, b * rl));
    } else if constexpr (kDS) {
        constexpr float d = (DIR > 0 ? 1.f : -1.f);
        constexpr float c = hi_part(cos32d(K)), cl = lo_part(cos32d(K));
        constexpr float s = d * hi_part(sin32d(K)), sl = d * lo_part(sin32d(K));
        return make_float2(fmaf(t.x, c, fmaf(-t.y, s, fmaf(t.x, cl, -t.y * sl))),
                           fmaf(t.x, s, fmaf(t.y, c, fmaf(t.x, sl, t.y * cl))));
    } else {
        constexpr float c = cos32(K), s = (DIR > 0 ? 1.f : -1.f) * sin32(K);
        return make_float2(fmaf(t.x, c, -t.y * s), fmaf(t.x, s, t.y * c));
    }
}

// Radix-4 merge of the two outermost stages of the 32-point transforms (OCB_FFT_R4, default on).  Two radix-2 stages
// apply four twiddles to a group (j, j+8, j+16, j+24): W32^j, W32^(j+8), W16^j twice.  Since W32^8 = DIR*i is a free
// rotation, the same group needs three: W32^j, W32^2j, W32^3j — over the eight groups 16 full + 4 half-trivial
// multiplies instead of 20 + 6 (44 fewer instructions per transform with double-single twiddles), same additions, same
// output positions as the radix-2 stages they replace.
#ifndef OCB_FFT_R4
#define OCB_FFT_R4 1
#endif
template <int DIR>
__device__ __forceinline__ float2 rot_i(float2 t) {  // (DIR * i) * t
    return DIR > 0 ? make_float2(-t.y, t.x) : make_float2(t.y, -t.x);
}
template <int DIR, bool PK>
__device__ __forceinline__ void dif32_first_two_stages(float2* v) {
    static_for<0, 8>([&](auto jj) {
        constexpr int J = decltype(jj)::value;
        const float2 a = v[J], b = v[J + 8], c = v[J + 16], d = v[J + 24];
        const float2 s0 = cadd<PK>(a, c), d0 = csub<PK>(a, c), s1 = cadd<PK>(b, d), r = rot_i<DIR>(csub<PK>(b, d));
        v[J] = cadd<PK>(s0, s1);
        v[J + 8] = twiddle_mul<32, 2 * J, DIR>(csub<PK>(s0, s1));
        v[J + 16] = twiddle_mul<32, J, DIR>(cadd<PK>(d0, r));
        v[J + 24] = twiddle_mul<32, 3 * J, DIR>(csub<PK>(d0, r));
    });
}
template <int DIR, bool PK>
__device__ __forceinline__ void dit32_last_two_stages(float2* v) {
    static_for<0, 8>([&](auto jj) {
        constexpr int J = decltype(jj)::value;
        const float2 a = v[J], b = twiddle_mul<32, 2 * J, DIR>(v[J + 8]);
        const float2 c = twiddle_mul<32, J, DIR>(v[J + 16]), d = twiddle_mul<32, 3 * J, DIR>(v[J + 24]);
        const float2 p0 = cadd<PK>(a, b), p1 = csub<PK>(a, b), q0 = cadd<PK>(c, d), r = rot_i<DIR>(csub<PK>(c, d));
        v[J] = cadd<PK>(p0, q0);
        v[J + 16] = csub<PK>(p0, q0);
        v[J + 8] = cadd<PK>(p1, r);
        v[J + 24] = csub<PK>(p1, r);
    });
}

// decimation in frequency: v natural -> X[k] at v[brev<R>(k)]
template <int R, int DIR, bool PK = true>
__device__ __forceinline__ void fft_dif(float2* v) {
    constexpr int FIRST = (R == 32 && OCB_FFT_R4) ? 2 : 0;
    if constexpr (FIRST == 2) dif32_first_two_stages<DIR, PK>(v);
    static_for<FIRST, ilog2(R)>([&](auto st) {
        constexpr int LEN = R >> decltype(st)::value, HALF = LEN / 2;
        static_for<0, R / LEN>([&](auto grp) {
            constexpr int BASE = decltype(grp)::value * LEN;
            static_for<0, HALF>([&](auto jj) {
                constexpr int J = decltype(jj)::value;
                const float2 a = v[BASE + J], b = v[BASE + J + HALF];
                v[BASE + J] = cadd<PK>(a, b);
                v[BASE + J + HALF] = twiddle_mul<LEN, J, DIR>(csub<PK>(a, b));
            });
        });
    });
}

// decimation in time: v[brev<R>(n)] = x[n] on input -> X[k] at v[k]
template <int R, int DIR, bool PK = true>
__device__ __forceinline__ void fft_dit(float2* v) {
    constexpr int LAST = (R == 32 && OCB_FFT_R4) ? ilog2(R) - 2 : ilog2(R);
    static_for<0, LAST>([&](auto st) {
        constexpr int LEN = 2 << decltype(st)::value, HALF = LEN / 2;
        static_for<0, R / LEN>([&](auto grp) {
            constexpr int BASE = decltype(grp)::value * LEN;
            static_for<0, HALF>([&](auto jj) {
                constexpr int J = decltype(jj)::value;
                const float2 a = v[BASE + J];
                const float2 b = twiddle_mul<LEN, J, DIR>(v[BASE + J + HALF]);
                v[BASE + J] = cadd<PK>(a, b);
                v[BASE + J + HALF] = csub<PK>(a, b);
            });
        });
    });
    if constexpr (LAST != ilog2(R)) dit32_last_two_stages<DIR, PK>(v);
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
// tw_lo: the second table (lo parts or modulus corrections), same layout (read only when OCB_TW_MODE != 0)
template <int Q, int CP, int PAD, bool HALF = false, bool PK = true, typename SyncF>
__device__ __forceinline__ void coop_fft_forward(float2* v, float* xr, float* xi, const float2* __restrict__ tw,
                                                 const float2* __restrict__ tw_lo, int q, int c, SyncF&& sync) {
    constexpr int G = 32 / Q, STR = Q * CP + PAD;
    fft_dif<32, -1, PK>(v);  // v[brev5(ka)] = Z[ka]
    if constexpr (!HALF) {
        static_for<0, 32>([&](auto kk) {
            constexpr int KA = decltype(kk)::value, SLOT = brev<32>(KA);
            float2 z = v[SLOT];
            if constexpr (KA != 0) z = cmul_ds(z, tw[KA * Q + q], kLO ? tw_lo[KA * Q + q] : float2{});
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
            if constexpr (KA != 0) v[SLOT] = cmul_ds(v[SLOT], tw[KA * Q + q], kLO ? tw_lo[KA * Q + q] : float2{});
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
    static_for<0, G>([&](auto gg) { fft_dif<Q, -1, PK>(v + decltype(gg)::value * Q); });
}

// twt / twt_lo: the TRANSPOSED copy of the table (hi and lo parts), entry [J][ka]; for Q = 32 the table is
// symmetric and the transposed copy is the table itself.
template <int Q, int CP, int PAD, bool HALF = false, bool PK = true, typename SyncF>
__device__ __forceinline__ void coop_fft_inverse(float2* v, float* xr, float* xi, const float2* __restrict__ twt,
                                                 const float2* __restrict__ twt_lo, int q, int c, SyncF&& sync) {
    constexpr int G = 32 / Q, STR = Q * CP + PAD;
    static_for<0, G>([&](auto gg) { fft_dit<Q, +1, PK>(v + decltype(gg)::value * Q); });  // over kq -> q'
    static_for<0, G>([&](auto gg) {
        constexpr int GI = decltype(gg)::value;
        const int ka = q * G + GI;
        static_for<0, Q>([&](auto jj) {
            constexpr int J = decltype(jj)::value;  // q'
            // transposed copy of the table: entry [J][ka], so that lanes (ka) are contiguous
            const float2 z = cmul_conj_ds(v[GI * Q + J], twt[J * 32 + ka], kLO ? twt_lo[J * 32 + ka] : float2{});  // conj twiddle
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
    fft_dit<32, +1, PK>(v);  // natural a'
}

}  // namespace fft
}  // namespace ocb
