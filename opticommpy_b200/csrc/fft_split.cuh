// Pair-split variant of the cooperative 1024-point transform (Q = 32): every thread owns 16 complex
// samples instead of 32 and each 32-point register FFT is shared by a lane pair (partner = lane ^ PM):
// its first (DIF) / last (DIT) radix-2 stage is one warp-shuffle exchange, the remaining 16-point FFT
// is private.  Twice the warps per transform, half the registers and half the unrolled code per thread.
//
// Thread coordinates inside a 64-thread transform group: x in [0,32) (position), h in {0,1} (half).
//   forward  in : v[j]            = x_in[32*(16h + j) + x]                       (natural order)
//   forward  out: v[brev16(k)]    = X[k1 = x + 32*(2k + h)]
//   inverse consumes exactly the forward output and returns the forward input arrangement.
#pragma once
#include "fft_core.cuh"

namespace ocb {
namespace fft {

// natural input a = 16h + j  ->  Z[ka = 2k + h] at v[brev16(k)]
template <int DIR, int PM>
__device__ __forceinline__ void split32_dif(float2* v, int h) {
    const bool hi = h != 0;
    static_for<0, 16>([&](auto jj) {
        constexpr int J = decltype(jj)::value;
        const float2 mine = v[J];
        const float2 oth = make_float2(__shfl_xor_sync(0xffffffffu, mine.x, PM), __shfl_xor_sync(0xffffffffu, mine.y, PM));
        // h = 0: x_j + x_{j+16} ; h = 1 (mine = x_{j+16}, oth = x_j): (x_j - x_{j+16}) * W32^j
        float2 r = hi ? make_float2(oth.x - mine.x, oth.y - mine.y) : make_float2(mine.x + oth.x, mine.y + oth.y);
        if constexpr (J != 0) {
            constexpr float c = cos32(J), s = (DIR > 0 ? 1.f : -1.f) * sin32(J);
            const float2 w = make_float2(hi ? c : 1.f, hi ? s : 0.f);
            r = cmul(r, w);
        }
        v[J] = r;
    });
    fft_dif<16, DIR>(v);
}

// Z[ka = 2k + h] at v[brev16(k)]  ->  natural output a = 16h + m at v[m]
template <int DIR, int PM>
__device__ __forceinline__ void split32_dit(float2* v, int h) {
    const bool hi = h != 0;
    fft_dit<16, DIR>(v);  // h = 0: E[m] (even ka) ; h = 1: O[m] (odd ka)
    static_for<0, 16>([&](auto mm) {
        constexpr int M = decltype(mm)::value;
        float2 mine = v[M];
        if constexpr (M != 0) {
            constexpr float c = cos32(M), s = (DIR > 0 ? 1.f : -1.f) * sin32(M);
            const float2 w = make_float2(hi ? c : 1.f, hi ? s : 0.f);
            mine = cmul(mine, w);  // O'[m] = W32^m O[m] on the odd half
        }
        const float2 oth = make_float2(__shfl_xor_sync(0xffffffffu, mine.x, PM), __shfl_xor_sync(0xffffffffu, mine.y, PM));
        // h = 0: E + O' -> a = m ; h = 1 (mine = O', oth = E): E - O' -> a = m + 16
        v[M] = hi ? make_float2(oth.x - mine.x, oth.y - mine.y) : make_float2(mine.x + oth.x, mine.y + oth.y);
    });
}

// Exchange buffer: planar xr/xi, element (ka, q, c) at ka*STR + q*CP + c with STR = 32*CP + PAD.
// tw[ka*32 + q] = exp(-2 pi i q ka / 1024).
template <int CP, int PAD, int PM, typename SyncF>
__device__ __forceinline__ void coop1024s_forward(float2* v, float* xr, float* xi, const float2* __restrict__ tw,
                                                  int x, int h, int c, SyncF&& sync) {
    constexpr int STR = 32 * CP + PAD;
    split32_dif<-1, PM>(v, h);
    static_for<0, 16>([&](auto kk) {
        constexpr int K = decltype(kk)::value, SLOT = brev<16>(K);
        const int ka = 2 * K + h;
        const float2 z = cmul(v[SLOT], __ldg(tw + ka * 32 + x));
        xr[ka * STR + x * CP + c] = z.x;
        xi[ka * STR + x * CP + c] = z.y;
    });
    sync();
    static_for<0, 16>([&](auto jj) {  // now (t = x, h2 = h): Z[t][q = 16h + j]
        constexpr int J = decltype(jj)::value;
        const int q = 16 * h + J;
        v[J] = make_float2(xr[x * STR + q * CP + c], xi[x * STR + q * CP + c]);
    });
    split32_dif<-1, PM>(v, h);
}

template <int CP, int PAD, int PM, typename SyncF>
__device__ __forceinline__ void coop1024s_inverse(float2* v, float* xr, float* xi, const float2* __restrict__ tw,
                                                  int x, int h, int c, SyncF&& sync) {
    constexpr int STR = 32 * CP + PAD;
    split32_dit<+1, PM>(v, h);  // v[m] = y[q = 16h + m] of row ka = x
    static_for<0, 16>([&](auto mm) {
        constexpr int M = decltype(mm)::value;
        const int q = 16 * h + M;
        const float2 z = cmul_conj(v[M], __ldg(tw + q * 32 + x));  // W^{x q} is symmetric: lane-contiguous read
        xr[x * STR + q * CP + c] = z.x;
        xi[x * STR + q * CP + c] = z.y;
    });
    sync();
    static_for<0, 16>([&](auto kk) {
        constexpr int K = decltype(kk)::value, SLOT = brev<16>(K);
        const int ka = 2 * K + h;
        v[SLOT] = make_float2(xr[ka * STR + x * CP + c], xi[ka * STR + x * CP + c]);
    });
    split32_dit<+1, PM>(v, h);
}

}  // namespace fft
}  // namespace ocb
