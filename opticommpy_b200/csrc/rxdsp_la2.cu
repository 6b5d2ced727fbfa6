// Latency-mode adaptive equalizer with a TWO-symbol look-ahead (one warp per output mode of one stream, one tap per lane).
//
// coreAdaptEq (optic/dsp/equalization.py:461-510) is serial:  o_s = H_s . x_s ;  g_s = err(o_s) ;
// H_{s+1} = H_s + mu g_s conj(x_s) [/ ||x_s||^2 for NLMS].  Substituting two updates into the output gives
//     o_{s+2} = H_s . x_{s+2} + mu g_s D_{s,s+2} + mu g_{s+1} D_{s+1,s+2},
//     D_{i,j} = sum_n [ sum_t conj(x_n,i[t]) x_n,j[t] ] / ||x_n,i||^2      (input only, computed per staged chunk)
// so the 32-lane reduction of H_s . x_{s+2} — five shuffle stages, ~140 cycles of latency — has two symbol periods to
// complete and leaves the per-symbol dependency chain as: one complex FMA -> error term -> mu g.  The one-symbol
// look-ahead kernel (rxdsp_eq_la.cuh) still has that reduction inside a two-symbol recurrence (180 cycles per symbol).
// Same arithmetic as the reference up to the rounding of the rearranged sums (tests: relative L2 against the float64
// oracle, identical hard decisions).
//
// Schedule of trip i (symbol i), W[k] = window registers of symbol k:
//   o_i      = Q_i + mu g_{i-1} D_{i-1,i}                      critical path: 1 complex FMA, error term, scale by mu
//   H_i      = H_{i-1} + mu g_{i-1} conj(W[i-1])               (its inputs are known when the trip starts)
//   R_{i+2}  = first two butterfly stages of H_i . W[i+2]
//   Q_{i+1}  = last three stages of R_{i+1} + mu g_{i-1} D_{i-1,i+1}
//   W[i+3] is loaded into the registers W[i-1] occupied.
// Scope: nTaps <= 32 (one tap per lane), no widely-linear taps, no tap history; the dispatcher in rxdsp.cu sends everything
// else to the one-symbol kernel.
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "../../include/opticomm_b200.h"
#include "common.cuh"

using namespace ocb;

namespace {

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <int NM>
__device__ __forceinline__ void lds_window(unsigned addr, float2 (&w)[NM]) {
    if constexpr (NM == 1) {
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(w[0].x), "=f"(w[0].y) : "r"(addr));
    } else {
#pragma unroll
        for (int n = 0; n < NM; n += 2)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(w[n].x), "=f"(w[n].y), "=f"(w[n + 1].x), "=f"(w[n + 1].y)
                         : "r"(addr + 8u * n));
    }
}

#include "rxdsp_eq_la.cuh"  // eq_error_term (the error terms of cmaUp / rdeUp / dardeUp / nlmsUp / ddlmsUp)

constexpr int kLa2Chunk = 128;  // symbols per staged chunk: a multiple of 4 (four window register sets rotate their roles)
constexpr int kLa2Ahead = 3;    // windows of the next chunk's first symbols that must sit in this chunk's buffer

template <int NM>
__global__ void __launch_bounds__(32 * NM)
k_mimo_eq_la2(const float2* __restrict__ X, const float2* __restrict__ REF, float2* __restrict__ Hg,
              float2* __restrict__ Y, float* __restrict__ ERR, int64_t xStride, int64_t refStride, int64_t yStride,
              int64_t errStride, int64_t errModeStride, int64_t L, int nTaps, int SpS, int alg, float mu,
              const float2* __restrict__ constSymb, int M, const float* __restrict__ radii, int nR, float Rcma) {
    extern __shared__ __align__(16) float2 smem_la2[];
    constexpr int NT = 32 * NM;
    const int tid = threadIdx.x, l = tid & 31, m = tid >> 5;
    const int stream = blockIdx.x;
    const int rows_chunk = (kLa2Chunk + kLa2Ahead - 1) * SpS + nTaps;  // windows of local symbols 0 .. chunk + 2
    const int RP = (rows_chunk + SpS - 1) / SpS + 1;                   // rows per sample phase in the de-interleaved arrays
    float2* xbuf = smem_la2;                                  // [2][rows_chunk*NM]
    float2* rbuf = xbuf + 2 * rows_chunk * NM;                // [2][kLa2Chunk*NM]
    float2* c1buf = rbuf + 2 * kLa2Chunk * NM;                // [SpS][RP][NM]  conj(x_n[r]) x_n[r + SpS]
    float2* c2buf = c1buf + SpS * RP * NM;                    // [SpS][RP][NM]  conj(x_n[r]) x_n[r + 2 SpS]
    float2* D1buf = c2buf + SpS * RP * NM;                    // [kLa2Chunk]  D_{i,i+1}
    float2* D2buf = D1buf + kLa2Chunk;                        // [kLa2Chunk]  D_{i,i+2}
    float2* zrow = D2buf + kLa2Chunk;                         // [NM] zeros: window of the lanes beyond nTaps
    float* pbuf = reinterpret_cast<float*>(zrow + NM);        // [SpS][RP][NM]  |x_n[r]|^2   (NLMS)
    float* ibuf = pbuf + SpS * RP * NM;                       // [kLa2Chunk][NM] 1/||x_n window||^2 (NLMS)

    const float2* x = X + (int64_t)stream * xStride;
    const float2* ref = REF ? REF + (int64_t)stream * refStride : nullptr;
    float2* Hs = Hg + (int64_t)stream * NM * NM * nTaps;
    float2* y = Y + (int64_t)stream * yStride;
    float* err = ERR + (int64_t)stream * errStride + (int64_t)m * errModeStride;

    const bool tapv = l < nTaps;
    float2 H[NM];  // rows m + n*NM, n < NM: tap l from input mode n to output mode m
#pragma unroll
    for (int n = 0; n < NM; ++n) H[n] = tapv ? Hs[(m + n * NM) * nTaps + l] : make_float2(0.f, 0.f);
    if (tid < NM) zrow[tid] = make_float2(0.f, 0.f);

    constexpr int kMaxR = 10;
    float rad2[kMaxR], thr2[kMaxR];
#pragma unroll
    for (int i = 0; i < kMaxR; ++i) {
        const float ri = (radii && i < nR) ? radii[i] : 0.f;
        const float rp = (radii && i >= 1 && i < nR) ? radii[i - 1] : 0.f;
        const float mid = 0.5f * (rp + ri);
        rad2[i] = ri * ri;
        thr2[i] = (radii && i >= 1 && i < nR) ? mid * mid : 3.4e38f;
    }

    const unsigned xbuf_s = (unsigned)__cvta_generic_to_shared(xbuf);
    const unsigned zrow_s = (unsigned)__cvta_generic_to_shared(zrow);
    const unsigned sym_stride = tapv ? (unsigned)(SpS * NM) * 8u : 0u;
    const unsigned woff = (unsigned)(l * NM) * 8u;
    const int64_t nchunks = (L + kLa2Chunk - 1) / kLa2Chunk;

    // windows present in the buffer of chunk k: local symbols 0 .. min(chunk + kLa2Ahead, L - s0) - 1
    auto chunk_windows = [&](int64_t k) -> int {
        const int64_t left = L - k * kLa2Chunk;
        return (int)(left < kLa2Chunk + kLa2Ahead ? left : kLa2Chunk + kLa2Ahead);
    };
    auto stage = [&](int64_t k) {
        const int64_t s0 = k * kLa2Chunk;
        if (s0 >= L) return;
        const int nsym = (int)((L - s0) < kLa2Chunk ? (L - s0) : kLa2Chunk);
        const int rows = (chunk_windows(k) - 1) * SpS + nTaps;
        float2* dst = xbuf + (k & 1) * rows_chunk * NM;
        const float2* src = x + s0 * SpS * NM;
        for (int i = tid; i < rows * NM; i += NT) cp_async8(dst + i, src + i);
        if (ref) {
            float2* rd = rbuf + (k & 1) * kLa2Chunk * NM;
            const float2* rs = ref + s0 * NM;
            for (int i = tid; i < nsym * NM; i += NT) cp_async8(rd + i, rs + i);
        }
    };
    // tap sum over the warp, split 2 + 3 butterfly stages
    auto dot_begin = [&](const float2 (&w)[NM]) -> float2 {
        float2 a = make_float2(0.f, 0.f);
#pragma unroll
        for (int n = 0; n < NM; ++n) {
            const float2 pr = cmul(H[n], w[n]);  // equalization.py:464-468
            a.x += pr.x; a.y += pr.y;
        }
#pragma unroll
        for (int st = 0; st < 2; ++st) {
            a.x += __shfl_xor_sync(0xffffffffu, a.x, 16 >> st);
            a.y += __shfl_xor_sync(0xffffffffu, a.y, 16 >> st);
        }
        return a;
    };
    auto dot_end = [&](float2 a) -> float2 {
#pragma unroll
        for (int st = 2; st < 5; ++st) {
            a.x += __shfl_xor_sync(0xffffffffu, a.x, 16 >> st);
            a.y += __shfl_xor_sync(0xffffffffu, a.y, 16 >> st);
        }
        return a;
    };

    auto run = [&](auto algc, auto fewc) {
        constexpr int ALG = decltype(algc)::value;
        constexpr bool FEW = decltype(fewc)::value;
        stage(0);
        cp_async_commit();
        float2 gmu = make_float2(0.f, 0.f);  // mu * g_{i-1}
        float invprev[NM];                   // 1 / ||x_n,{i-1}||^2 (NLMS)
#pragma unroll
        for (int n = 0; n < NM; ++n) invprev[n] = 1.f;
        float2 W0[NM], W1[NM], W2[NM], W3[NM];  // window of the symbol with local index j lives in W[j % 4]
#pragma unroll
        for (int n = 0; n < NM; ++n) W0[n] = W1[n] = W2[n] = W3[n] = make_float2(0.f, 0.f);
        float2 Q = make_float2(0.f, 0.f);    // Q_i: tap sum of symbol i with every correction but the last one
        float2 R = make_float2(0.f, 0.f);    // partially reduced tap sum of symbol i+1
        float prev_err = 0.f;
        const bool wr = l == 0;

        auto update_taps = [&](const float2 (&xp)[NM]) {  // H_i = H_{i-1} + mu g_{i-1} conj(x_{i-1}) [/ ||x_{i-1}||^2]
            if constexpr (ALG != OCB_ALG_STATIC) {
#pragma unroll
                for (int n = 0; n < NM; ++n) {
                    float2 wg = gmu;
                    if constexpr (ALG == OCB_ALG_NLMS) { wg.x *= invprev[n]; wg.y *= invprev[n]; }  // equalization.py:563
                    const float2 u = cmul_conj(wg, xp[n]);
                    H[n].x += u.x; H[n].y += u.y;
                }
            }
        };

        for (int64_t k = 0; k < nchunks; ++k) {
            cp_async_wait_all();
            __syncthreads();  // chunk k landed; chunk k-1 (and its D / inv arrays) fully consumed
            stage(k + 1);     // lands while this chunk is processed
            cp_async_commit();
            const int64_t s0 = k * kLa2Chunk;
            const int nsym = (int)((L - s0) < kLa2Chunk ? (L - s0) : kLa2Chunk);
            const int nwin = chunk_windows(k);
            const int rows = (nwin - 1) * SpS + nTaps;
            const float2* xb = xbuf + (k & 1) * rows_chunk * NM;
            const float2* rb = rbuf + (k & 1) * kLa2Chunk * NM;

            // ---- pre-pass 1: per input row r the lag-1 and lag-2 products and |x|^2, de-interleaved by sample phase
            for (int r = tid; r < rows; r += NT) {
                const int ph = r % SpS, qd = r / SpS;
                const int ci = (ph * RP + qd) * NM;
#pragma unroll
                for (int n = 0; n < NM; ++n) {
                    const float2 a = xb[r * NM + n];
                    if (r + SpS < rows) c1buf[ci + n] = cmul_conj(xb[(r + SpS) * NM + n], a);
                    if (r + 2 * SpS < rows) c2buf[ci + n] = cmul_conj(xb[(r + 2 * SpS) * NM + n], a);
                    if constexpr (ALG == OCB_ALG_NLMS) pbuf[ci + n] = cabs2(a);
                }
            }
            __syncthreads();
            // ---- pre-pass 2: D_{i,i+1} and D_{i,i+2} = sum_n inv_n(i) sum_t c_n[i SpS + t]
            for (int i = tid; i < nsym; i += NT) {
                float2 C1[NM], C2[NM];
                float P[NM];
#pragma unroll
                for (int n = 0; n < NM; ++n) { C1[n] = C2[n] = make_float2(0.f, 0.f); P[n] = 0.f; }
                const bool has1 = i + 1 < nwin, has2 = i + 2 < nwin;
                int ph = 0, qd = 0;
                for (int t = 0; t < nTaps; ++t) {
                    const int ci = (ph * RP + qd + i) * NM;
#pragma unroll
                    for (int n = 0; n < NM; ++n) {
                        if (has1) { const float2 c = c1buf[ci + n]; C1[n].x += c.x; C1[n].y += c.y; }
                        if (has2) { const float2 c = c2buf[ci + n]; C2[n].x += c.x; C2[n].y += c.y; }
                        if constexpr (ALG == OCB_ALG_NLMS) P[n] += pbuf[ci + n];
                    }
                    if (++ph == SpS) { ph = 0; ++qd; }
                }
                float2 D1 = make_float2(0.f, 0.f), D2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int n = 0; n < NM; ++n) {
                    float inv = 1.f;
                    if constexpr (ALG == OCB_ALG_NLMS) { inv = 1.0f / P[n]; ibuf[i * NM + n] = inv; }
                    D1.x = fmaf(C1[n].x, inv, D1.x); D1.y = fmaf(C1[n].y, inv, D1.y);
                    D2.x = fmaf(C2[n].x, inv, D2.x); D2.y = fmaf(C2[n].y, inv, D2.y);
                }
                D1buf[i] = D1;
                D2buf[i] = D2;
            }
            __syncthreads();

            const unsigned wbase = tapv ? xbuf_s + (unsigned)((k & 1) * rows_chunk * NM) * 8u + woff : zrow_s;
            auto load_win = [&](int local_sym, float2 (&w)[NM]) { lds_window<NM>(wbase + (unsigned)local_sym * sym_stride, w); };
            if (k == 0) {
                // o_0 = H_0 . x_0 and the first stages of H_0 . x_1: the two symbols that have no two-step history
                load_win(0, W0);
                Q = dot_end(dot_begin(W0));
                if (1 < nwin) { load_win(1, W1); R = dot_begin(W1); }
                if (2 < nwin) load_win(2, W2);
            }
            float2* yp = y + (s0 * NM + m);
            float* ep = err + s0;

            // XP = W[i-1] (freed after the tap update, refilled with W[i+3]), XN = W[i+2]
            auto trip = [&](const int i, float2 (&XP)[NM], const float2 (&XN)[NM]) {
                // D_{i-1,i} and D_{i-1,i+1} were fetched at the end of the previous trip (Dp1, Dp2 below live in gD1/gD2)
                float2 o;
                o.x = fmaf(gmu.x, Dp1.x, fmaf(-gmu.y, Dp1.y, Q.x));
                o.y = fmaf(gmu.x, Dp1.y, fmaf(gmu.y, Dp1.x, Q.y));
                if (wr) yp[i * NM] = o;  // equalization.py:473
                // taps H_i and the tap sum of symbol i+2 with them
                update_taps(XP);
                float2 Rn = make_float2(0.f, 0.f);
                if (i + 2 < nwin) Rn = dot_begin(XN);
                // finish the tap sum of symbol i+1 and add the correction of g_{i-1}
                const float2 A = dot_end(R);
                Q.x = fmaf(gmu.x, Dp2.x, fmaf(-gmu.y, Dp2.y, A.x));
                Q.y = fmaf(gmu.x, Dp2.y, fmaf(gmu.y, Dp2.x, A.y));
                R = Rn;
                // error term of symbol i
                float2 refsym = make_float2(0.f, 0.f);
                if constexpr (ALG == OCB_ALG_NLMS || ALG == OCB_ALG_DARDE) refsym = rb[i * NM + m];
                float2 g;
                float esq;
                eq_error_term<ALG, FEW>(o, refsym, Rcma, rad2, thr2, nR, radii, constSymb, M, l, prev_err, g, esq);
                prev_err = esq;
                if (wr) ep[i] = esq;
                gmu = make_float2(mu * g.x, mu * g.y);
                Dp1 = D1buf[i];
                Dp2 = D2buf[i];
                if constexpr (ALG == OCB_ALG_NLMS) {
#pragma unroll
                    for (int n = 0; n < NM; ++n) invprev[n] = ibuf[i * NM + n];
                }
                if (i + 3 < nwin) load_win(i + 3, XP);
            };
            int i = 0;
            for (; i + 4 <= nsym; i += 4) {
                trip(i, W3, W2);
                trip(i + 1, W0, W3);
                trip(i + 2, W1, W0);
                trip(i + 3, W2, W1);
            }
            if (i < nsym) { trip(i, W3, W2); ++i; }  // tail of the last chunk (roles need not line up any more)
            if (i < nsym) { trip(i, W0, W3); ++i; }
            if (i < nsym) { trip(i, W1, W0); ++i; }
        }
        // taps after the last symbol L-1: its window sits in W[(local index) % 4]
        {
            const int jl = (int)((L - 1) % kLa2Chunk) % 4;
            if (jl == 0) update_taps(W0);
            else if (jl == 1) update_taps(W1);
            else if (jl == 2) update_taps(W2);
            else update_taps(W3);
        }
    };
    const bool few = nR <= 4;
    using T = std::true_type;
    using F = std::false_type;
    switch (alg) {
        case OCB_ALG_CMA: run(std::integral_constant<int, OCB_ALG_CMA>{}, T{}); break;
        case OCB_ALG_RDE:
            if (few) run(std::integral_constant<int, OCB_ALG_RDE>{}, T{});
            else run(std::integral_constant<int, OCB_ALG_RDE>{}, F{});
            break;
        case OCB_ALG_NLMS: run(std::integral_constant<int, OCB_ALG_NLMS>{}, T{}); break;
        case OCB_ALG_DDLMS: run(std::integral_constant<int, OCB_ALG_DDLMS>{}, T{}); break;
        case OCB_ALG_DARDE: run(std::integral_constant<int, OCB_ALG_DARDE>{}, T{}); break;
        default: run(std::integral_constant<int, OCB_ALG_STATIC>{}, T{}); break;
    }
    cp_async_wait_all();
#pragma unroll
    for (int n = 0; n < NM; ++n)
        if (tapv) Hs[(m + n * NM) * nTaps + l] = H[n];
}

template <int NM>
int launch_la2(cudaStream_t st, const float2* X, const float2* REF, float2* H, float2* Y, float* ERR, int nStreams,
               int64_t xs, int64_t rs, int64_t ys, int64_t es, int64_t ems, int64_t L, int nTaps, int SpS, int alg, float mu,
               const float2* cs, int M, const float* radii, int nR, float Rcma) {
    const int rows_chunk = (kLa2Chunk + kLa2Ahead - 1) * SpS + nTaps;
    const int RP = (rows_chunk + SpS - 1) / SpS + 1;
    const size_t smem = ((size_t)2 * rows_chunk * NM + 2 * kLa2Chunk * NM + 2 * (size_t)SpS * RP * NM + 2 * kLa2Chunk + NM) * sizeof(float2) +
                        ((size_t)SpS * RP * NM + (size_t)kLa2Chunk * NM) * sizeof(float);
    if (smem > 200 * 1024) return -1;  // the caller falls back to the one-symbol kernel
    OCB_CUDA(cudaFuncSetAttribute(k_mimo_eq_la2<NM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    OCB_LAUNCH((k_mimo_eq_la2<NM>), nStreams, 32 * NM, smem, st, X, REF, H, Y, ERR, xs, rs, ys, es, ems, L, nTaps, SpS, alg, mu, cs,
               M, radii, nR, Rcma);
    return 0;
}

}  // namespace

namespace ocb {
// -1: geometry not covered (the caller uses the one-symbol look-ahead kernel); 0: launched; 1: error
int mimo_eq_la2_try(int nModes, cudaStream_t st, const float2* X, const float2* REF, float2* H, float2* Y, float* ERR,
                    int nStreams, int64_t xs, int64_t rs, int64_t ys, int64_t es, int64_t ems, int64_t L, int nTaps, int SpS,
                    int alg, float mu, const float2* cs, int M, const float* radii, int nR, float Rcma) {
    if (nTaps > 32 || L < 1) return -1;
    switch (nModes) {
        case 1: return launch_la2<1>(st, X, REF, H, Y, ERR, nStreams, xs, rs, ys, es, ems, L, nTaps, SpS, alg, mu, cs, M, radii, nR, Rcma);
        case 2: return launch_la2<2>(st, X, REF, H, Y, ERR, nStreams, xs, rs, ys, es, ems, L, nTaps, SpS, alg, mu, cs, M, radii, nR, Rcma);
        case 4: return launch_la2<4>(st, X, REF, H, Y, ERR, nStreams, xs, rs, ys, es, ems, L, nTaps, SpS, alg, mu, cs, M, radii, nR, Rcma);
    }
    return -1;
}
}  // namespace ocb
