// Latency-mode adaptive equalizer (one warp per task) with a one-symbol look-ahead.
//
// The recurrence of coreAdaptEq (optic/dsp/equalization.py:461-510) is strictly serial:
//     o_s = H_s . x_s ;  g_s = err(o_s) ;  H_{s+1} = H_s + mu g_s conj(x_s) [/ ||x_s||^2 for NLMS]
// A lone warp that evaluates it literally pays, per symbol, a 5-stage shuffle reduction inside the
// loop-carried dependency chain (measured 270-360 cycles per symbol).  Substituting the update into the
// next output gives
//     o_{s+1} = H_s . x_{s+1} + mu g_s D_s ,   D_s = sum_n [sum_t conj(x_n,s[t]) x_n,s+1[t]] / ||x_n,s||^2
// where D_s depends on the INPUT only.  D_s is computed for a whole staged chunk in parallel (pre-pass),
// and the reduction of A_{s+1} = H_s . x_{s+1} starts as soon as g_{s-1} is known, so that it overlaps the
// error computation of symbol s; the chain per symbol shrinks to one complex FMA + the error term, with the
// tap-sum reduction pipelined one symbol ahead.  Same arithmetic up to the rounding of the rearranged sum
// (tests: relative L2 <= 1e-4 against the float64 oracle, identical decisions).
//
// Included by rxdsp.cu inside its anonymous namespace (uses cp_async8, lds_window).
#pragma once

// error term g_m and squared error of one output (equalization.py:826-829, 887-894, 953-959, 556, 688-691)
template <int ALG, bool FEWRINGS>
__device__ __forceinline__ void eq_error_term(const float2 o, const float2 refsym, const float Rcma, const float* rad2,
                                              const float* thr2, const int nR, const float* __restrict__ radii,
                                              const float2* __restrict__ constSymb, const int M, const int l,
                                              const float prev_err, float2& g, float& esq) {
    constexpr int kMaxR = 10;
    const float a2 = cabs2(o);
    if constexpr (ALG == OCB_ALG_CMA) {
        const float e = Rcma - a2;
        g = make_float2(e * o.x, e * o.y);
        esq = e * e;
    } else if constexpr (ALG == OCB_ALG_RDE) {
        float Rd2 = rad2[0];
        if constexpr (FEWRINGS) {
#pragma unroll
            for (int i = 1; i < 4; ++i) Rd2 = (a2 > thr2[i]) ? rad2[i] : Rd2;
        } else {
#pragma unroll
            for (int i = 1; i < kMaxR; ++i) Rd2 = (a2 > thr2[i]) ? rad2[i] : Rd2;
            for (int i = kMaxR; i < nR; ++i) {
                const float ri = radii[i], mid = 0.5f * (radii[i - 1] + ri);
                if (a2 > mid * mid) Rd2 = ri * ri;
            }
        }
        const float e = Rd2 - a2;
        g = make_float2(e * o.x, e * o.y);
        esq = e * e;
    } else if constexpr (ALG == OCB_ALG_DARDE) {
        const float Rd = sqrtf(cabs2(refsym));
        const float e = Rd * Rd - a2;
        g = make_float2(e * o.x, e * o.y);
        esq = e * e;
    } else if constexpr (ALG == OCB_ALG_NLMS) {
        g = make_float2(refsym.x - o.x, refsym.y - o.y);
        esq = cabs2(g);
    } else if constexpr (ALG == OCB_ALG_DDLMS) {  // nearest constellation point, first index on ties
        float best = 3.4e38f;
        int bi = 0x7fffffff;
        for (int c = l; c < M; c += 32) {
            const float2 sc = __ldg(constSymb + c);
            const float dd = cabs2(make_float2(o.x - sc.x, o.y - sc.y));
            if (dd < best) { best = dd; bi = c; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        const float2 sc = __ldg(constSymb + bi);
        g = make_float2(sc.x - o.x, sc.y - o.y);
        esq = cabs2(g);
    } else {  // static: no update (equalization.py:505-506)
        g = make_float2(0.f, 0.f);
        esq = prev_err;
    }
}

// Symbols per staged chunk: a multiple of 3, because the symbol loop is unrolled three times over the three
// window register sets (previous / current / next symbol) and their roles must line up at chunk boundaries.
constexpr int kLaChunk = 126;

// CTA = the NM tasks (output modes) of ONE stream, one warp each; lane l owns taps t = l + 32 j, j < TPL.
// HITS: storeCoeff (a per-symbol copy of the taps) compiled in or out.
template <int NM, int TPL, bool WL, bool HITS>
__global__ void __launch_bounds__(32 * NM)
k_mimo_eq_la(const float2* __restrict__ X, const float2* __restrict__ REF, float2* __restrict__ Hg,
             float2* __restrict__ HWg, float2* __restrict__ Y, float* __restrict__ ERR, float2* __restrict__ HIT,
             int64_t xStride, int64_t refStride, int64_t yStride, int64_t errStride, int64_t errModeStride,
             int64_t L, int nTaps, int SpS, int alg, float mu, const float2* __restrict__ constSymb, int M,
             const float* __restrict__ radii, int nR, float Rcma) {
    extern __shared__ __align__(16) float2 smem_la[];
    constexpr int NT = 32 * NM;
    const int tid = threadIdx.x, l = tid & 31, m = tid >> 5;
    const int stream = blockIdx.x;
    const int rows_chunk = kLaChunk * SpS + nTaps;     // incl. the look-ahead window of the chunk's last symbol
    const int RP = (rows_chunk + SpS - 1) / SpS + 1;   // rows per sample phase in the de-interleaved c / p arrays
    float2* xbuf = smem_la;                                   // [2][rows_chunk*NM]
    float2* rbuf = xbuf + 2 * rows_chunk * NM;                // [2][kLaChunk*NM]
    float2* cbuf = rbuf + 2 * kLaChunk * NM;                  // [SpS][RP][NM]  conj(x_n[r]) x_n[r+SpS]
    float2* Dbuf = cbuf + SpS * RP * NM;                      // [kLaChunk]
    float2* zrow = Dbuf + kLaChunk;                           // [NM] zeros: window of the lanes beyond nTaps
    float* pbuf = reinterpret_cast<float*>(zrow + NM);        // [SpS][RP][NM]  |x_n[r]|^2   (NLMS)
    float* ibuf = pbuf + SpS * RP * NM;                       // [kLaChunk][NM] 1/||x_n window||^2 (NLMS)

    const float2* x = X + (int64_t)stream * xStride;
    const float2* ref = REF ? REF + (int64_t)stream * refStride : nullptr;
    float2* Hs = Hg + (int64_t)stream * NM * NM * nTaps;
    float2* HWs = WL ? HWg + (int64_t)stream * NM * NM * nTaps : nullptr;
    float2* y = Y + (int64_t)stream * yStride;
    float* err = ERR + (int64_t)stream * errStride + (int64_t)m * errModeStride;
    float2* hit = HIT ? HIT + (int64_t)stream * L * NM * NM * nTaps : nullptr;

    float2 H[NM][TPL], HW[WL ? NM : 1][TPL];  // rows m + n*NM, n < NM
    bool tapv[TPL];
#pragma unroll
    for (int j = 0; j < TPL; ++j) {
        const int t = l + 32 * j;
        tapv[j] = t < nTaps;
#pragma unroll
        for (int n = 0; n < NM; ++n) {
            H[n][j] = tapv[j] ? Hs[(m + n * NM) * nTaps + t] : make_float2(0.f, 0.f);
            if (WL) HW[n][j] = tapv[j] ? HWs[(m + n * NM) * nTaps + t] : make_float2(0.f, 0.f);
        }
    }
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
    const unsigned sym_stride = (unsigned)(SpS * NM) * 8u;
    const int64_t nchunks = (L + kLaChunk - 1) / kLaChunk;

    // rows staged for chunk k: up to and including the window of the NEXT chunk's first symbol, if there is one
    auto chunk_rows = [&](int64_t k) -> int {
        const int64_t s0 = k * kLaChunk;
        const int nsym = (int)((L - s0) < kLaChunk ? (L - s0) : kLaChunk);
        return (s0 + nsym < L) ? nsym * SpS + nTaps : (nsym - 1) * SpS + nTaps;
    };
    auto stage = [&](int64_t k) {
        const int64_t s0 = k * kLaChunk;
        if (s0 >= L) return;
        const int nsym = (int)((L - s0) < kLaChunk ? (L - s0) : kLaChunk);
        const int rows = chunk_rows(k);
        float2* dst = xbuf + (k & 1) * rows_chunk * NM;
        const float2* src = x + s0 * SpS * NM;
        for (int i = tid; i < rows * NM; i += NT) cp_async8(dst + i, src + i);
        if (ref) {
            float2* rd = rbuf + (k & 1) * kLaChunk * NM;
            const float2* rs = ref + s0 * NM;
            for (int i = tid; i < nsym * NM; i += NT) cp_async8(rd + i, rs + i);
        }
    };
    // Tap-sum reduction over the warp, split in two: the first kSplit butterfly stages run in the trip that
    // forms the products, the rest in the next trip, interleaved with that trip's products — two reductions
    // are in flight at any time and neither sits alone on the per-symbol critical path.
    constexpr int kSplit = 3;
    auto tap_dot_begin = [&](const float2 (&w)[NM][TPL]) -> float2 {
        float2 a = make_float2(0.f, 0.f);
#pragma unroll
        for (int n = 0; n < NM; ++n)
#pragma unroll
            for (int j = 0; j < TPL; ++j) {
                // products accumulated inside the FMA chain (one instruction per real product instead of multiply + add:
                // a lone warp issues in order, so the instruction count per symbol is the time)
                const float2 h = H[n][j], xw = w[n][j];  // equalization.py:464-468 (plain dot)
                a.x = fmaf(h.x, xw.x, fmaf(-h.y, xw.y, a.x));
                a.y = fmaf(h.x, xw.y, fmaf(h.y, xw.x, a.y));
                if (WL) {
                    const float2 hw = HW[n][j];  // H_ . conj(x), :469-471
                    a.x = fmaf(hw.x, xw.x, fmaf(hw.y, xw.y, a.x));
                    a.y = fmaf(hw.y, xw.x, fmaf(-hw.x, xw.y, a.y));
                }
            }
#pragma unroll
        for (int st = 0; st < kSplit; ++st) {
            a.x += __shfl_xor_sync(0xffffffffu, a.x, 16 >> st);
            a.y += __shfl_xor_sync(0xffffffffu, a.y, 16 >> st);
        }
        return a;
    };
    auto tap_dot_end = [&](float2 a) -> float2 {  // all lanes end with the total
#pragma unroll
        for (int st = kSplit; st < 5; ++st) {
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
        float2 gmu = make_float2(0.f, 0.f);    // mu * g_{s-1}
        float2 Dprev = make_float2(0.f, 0.f);  // D_{s-1}
        float invprev[NM];
#pragma unroll
        for (int n = 0; n < NM; ++n) invprev[n] = 1.f;
        // three window register sets; symbol with local index i (mod 3) lives in W[i % 3]
        float2 W0[NM][TPL], W1[NM][TPL], W2[NM][TPL];
#pragma unroll
        for (int n = 0; n < NM; ++n)
#pragma unroll
            for (int j = 0; j < TPL; ++j) W0[n][j] = W1[n][j] = W2[n][j] = make_float2(0.f, 0.f);
        float2 Apart = make_float2(0.f, 0.f);  // partially reduced tap sum of the coming symbol
        float prev_err = 0.f;
        const bool wr = l == 0;

        // H_s = H_{s-1} + mu g_{s-1} conj(x_{s-1}) [/ ||x_{s-1}||^2]   (:838-840 and siblings)
        auto update_taps = [&](const float2 (&xp)[NM][TPL]) {
            if constexpr (ALG != OCB_ALG_STATIC) {
#pragma unroll
                for (int n = 0; n < NM; ++n) {
                    float2 wg = gmu;
                    if constexpr (ALG == OCB_ALG_NLMS) { wg.x *= invprev[n]; wg.y *= invprev[n]; }  // :563
#pragma unroll
                    for (int j = 0; j < TPL; ++j) {
                        const float2 xv = xp[n][j];  // H += wg conj(x), H_ += wg x, accumulated inside the FMA chains
                        H[n][j].x = fmaf(wg.x, xv.x, fmaf(wg.y, xv.y, H[n][j].x));
                        H[n][j].y = fmaf(wg.y, xv.x, fmaf(-wg.x, xv.y, H[n][j].y));
                        if (WL) {
                            HW[n][j].x = fmaf(wg.x, xv.x, fmaf(-wg.y, xv.y, HW[n][j].x));
                            HW[n][j].y = fmaf(wg.x, xv.y, fmaf(wg.y, xv.x, HW[n][j].y));
                        }
                    }
                }
            }
        };
        auto store_hit = [&](int64_t ind) {  // storeCoeff (:511-512): Hiter[:, :, ind] = H after the update of symbol ind
#pragma unroll
            for (int n = 0; n < NM; ++n)
#pragma unroll
                for (int j = 0; j < TPL; ++j)
                    if (tapv[j]) hit[(ind * NM * NM + m + n * NM) * nTaps + l + 32 * j] = H[n][j];
        };
        // per-lane window addressing: base (tap offset, or the zero row beyond nTaps) + symbol * step
        unsigned wstep[TPL], woff[TPL];
#pragma unroll
        for (int j = 0; j < TPL; ++j) {
            wstep[j] = tapv[j] ? sym_stride : 0u;
            woff[j] = (unsigned)((l + 32 * j) * NM) * 8u;
        }

        for (int64_t k = 0; k < nchunks; ++k) {
            cp_async_wait_all();
            __syncthreads();  // chunk k landed; chunk k-1 (and its D / inv arrays) fully consumed
            stage(k + 1);     // lands while this chunk is processed
            cp_async_commit();
            const int64_t s0 = k * kLaChunk;
            const int nsym = (int)((L - s0) < kLaChunk ? (L - s0) : kLaChunk);
            const int rows = chunk_rows(k);
            const bool more = s0 + nsym < L;
            const int nsymD = more ? nsym : nsym - 1;  // symbols of this chunk that have a successor
            const float2* xb = xbuf + (k & 1) * rows_chunk * NM;
            const float2* rb = rbuf + (k & 1) * kLaChunk * NM;

            // ---- pre-pass 1: per input row r, c_n[r] = conj(x_n[r]) x_n[r+SpS] and p_n[r] = |x_n[r]|^2, stored
            //      de-interleaved by sample phase so that pre-pass 2 reads consecutive entries per lane
            for (int r = tid; r < rows; r += NT) {
                const int ph = r % SpS, qd = r / SpS;
                const int ci = (ph * RP + qd) * NM;
#pragma unroll
                for (int n = 0; n < NM; ++n) {
                    const float2 a = xb[r * NM + n];
                    if (r + SpS < rows) cbuf[ci + n] = cmul_conj(xb[(r + SpS) * NM + n], a);
                    if constexpr (ALG == OCB_ALG_NLMS) pbuf[ci + n] = cabs2(a);
                }
            }
            __syncthreads();
            // ---- pre-pass 2: D_i = sum_n inv_n(i) sum_t c_n[i SpS + t]  (WL: + its conjugate)
            for (int i = tid; i < nsym; i += NT) {
                float2 C[NM];
                float P[NM];
#pragma unroll
                for (int n = 0; n < NM; ++n) { C[n] = make_float2(0.f, 0.f); P[n] = 0.f; }
                int ph = 0, qd = 0;
                for (int t = 0; t < nTaps; ++t) {
                    const int ci = (ph * RP + qd + i) * NM;
#pragma unroll
                    for (int n = 0; n < NM; ++n) {
                        if (i < nsymD) { const float2 c = cbuf[ci + n]; C[n].x += c.x; C[n].y += c.y; }
                        if constexpr (ALG == OCB_ALG_NLMS) P[n] += pbuf[ci + n];
                    }
                    if (++ph == SpS) { ph = 0; ++qd; }
                }
                float2 D = make_float2(0.f, 0.f);
#pragma unroll
                for (int n = 0; n < NM; ++n) {
                    float inv = 1.f;
                    if constexpr (ALG == OCB_ALG_NLMS) { inv = 1.0f / P[n]; ibuf[i * NM + n] = inv; }
                    D.x = fmaf(C[n].x, inv, D.x);
                    D.y = fmaf(C[n].y, inv, D.y);
                }
                if (WL) D = make_float2(2.f * D.x, 0.f);  // sum_t [conj(x_s) x_{s+1} + x_s conj(x_{s+1})]
                Dbuf[i] = D;
            }
            __syncthreads();

            const unsigned buf_s = xbuf_s + (unsigned)((k & 1) * rows_chunk * NM) * 8u;
            unsigned wbase[TPL];
#pragma unroll
            for (int j = 0; j < TPL; ++j) wbase[j] = tapv[j] ? buf_s + woff[j] : zrow_s;
            auto load_win = [&](int local_sym, float2 (&w)[NM][TPL]) {
#pragma unroll
                for (int j = 0; j < TPL; ++j) {
                    float2 t[NM];
                    lds_window<NM>(wbase[j] + (unsigned)local_sym * wstep[j], t);
#pragma unroll
                    for (int n = 0; n < NM; ++n) w[n][j] = t[n];
                }
            };
            const int last_win = more ? nsym : nsym - 1;  // highest local window index present in this buffer
            if (k == 0) {  // o_0 = H_0 . x_0 directly
                load_win(0, W0);
                Apart = tap_dot_begin(W0);
            }
            if (1 <= last_win) load_win(1, W1);
            float2* yp = y + (s0 * NM + m);
            float* ep = err + s0;

            // one symbol: XP holds x_{s-1} (freed after the tap update, then refilled with x_{s+2}), WN holds x_{s+1}
            auto trip = [&](const int i, float2 (&XP)[NM][TPL], const float2 (&WN)[NM][TPL]) {
                // (0) finish the tap sum of THIS symbol (its products were formed one trip earlier)
                const float2 A = tap_dot_end(Apart);
                // (2) taps that define o_s, from the error of the previous symbol
                update_taps(XP);
                if constexpr (HITS) { if (s0 + i > 0) store_hit(s0 + i - 1); }
                // (3) tap sum of the NEXT symbol with those taps: products + first reduction stages
                Apart = tap_dot_begin(WN);
                // (1) output of this symbol: o_s = A_s + mu g_{s-1} D_{s-1}
                float2 o;
                o.x = fmaf(gmu.x, Dprev.x, fmaf(-gmu.y, Dprev.y, A.x));
                o.y = fmaf(gmu.x, Dprev.y, fmaf(gmu.y, Dprev.x, A.y));
                if (wr) yp[i * NM] = o;  // equalization.py:473
                // (4) error term
                float2 refsym = make_float2(0.f, 0.f);
                if constexpr (ALG == OCB_ALG_NLMS || ALG == OCB_ALG_DARDE) refsym = rb[i * NM + m];
                float2 g;
                float esq;
                eq_error_term<ALG, FEW>(o, refsym, Rcma, rad2, thr2, nR, radii, constSymb, M, l, prev_err, g, esq);
                prev_err = esq;
                if (wr) ep[i] = esq;
                gmu = make_float2(mu * g.x, mu * g.y);
                Dprev = Dbuf[i];
                if constexpr (ALG == OCB_ALG_NLMS) {
#pragma unroll
                    for (int n = 0; n < NM; ++n) invprev[n] = ibuf[i * NM + n];
                }
                // prefetch the window after next into the register set that has just been consumed
                if (i + 2 <= last_win) load_win(i + 2, XP);
            };
            int i = 0;
            for (; i + 3 <= nsym; i += 3) {
                trip(i, W2, W1);
                trip(i + 1, W0, W2);
                trip(i + 2, W1, W0);
            }
            if (i < nsym) { trip(i, W2, W1); ++i; }      // tail of the last chunk (roles need not line up any more)
            if (i < nsym) { trip(i, W0, W2); ++i; }
        }
        // taps after the last symbol L-1: its window sits in W[(local index) % 3]
        {
            const int jl = (int)((L - 1) % kLaChunk) % 3;
            if (jl == 0) update_taps(W0);
            else if (jl == 1) update_taps(W1);
            else update_taps(W2);
            if constexpr (HITS) store_hit(L - 1);
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
#pragma unroll
        for (int j = 0; j < TPL; ++j) {
            const int t = l + 32 * j;
            if (t < nTaps) {
                Hs[(m + n * NM) * nTaps + t] = H[n][j];
                if (WL) HWs[(m + n * NM) * nTaps + t] = HW[n][j];
            }
        }
}

template <int NM, int TPL>
int launch_mimo_la(bool wl, cudaStream_t st, const float2* X, const float2* REF, float2* H, float2* HW, float2* Y,
                   float* ERR, float2* HIT, int nStreams, int64_t xs, int64_t rs, int64_t ys, int64_t es, int64_t ems,
                   int64_t L, int nTaps, int SpS, int alg, float mu, const float2* cs, int M, const float* radii, int nR,
                   float Rcma) {
    const int rows_chunk = kLaChunk * SpS + nTaps;
    const int RP = (rows_chunk + SpS - 1) / SpS + 1;
    const size_t smem = ((size_t)2 * rows_chunk * NM + 2 * kLaChunk * NM + (size_t)SpS * RP * NM + kLaChunk + NM) * sizeof(float2) +
                        ((size_t)SpS * RP * NM + (size_t)kLaChunk * NM) * sizeof(float);
    OCB_REQUIRE(smem <= 200 * 1024, "mimo_eq_run: SpS/nTaps too large for the staged input chunk");
#define OCB_LA_LAUNCH(WL_, HITS_)                                                                                   \
    do {                                                                                                            \
        OCB_CUDA(cudaFuncSetAttribute(k_mimo_eq_la<NM, TPL, WL_, HITS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        OCB_LAUNCH((k_mimo_eq_la<NM, TPL, WL_, HITS_>), nStreams, 32 * NM, smem, st, X, REF, H, HW, Y, ERR, HIT, xs, rs, ys, es,   \
                   ems, L, nTaps, SpS, alg, mu, cs, M, radii, nR, Rcma);                                            \
    } while (0)
    if (wl) { if (HIT) OCB_LA_LAUNCH(true, true); else OCB_LA_LAUNCH(true, false); }
    else { if (HIT) OCB_LA_LAUNCH(false, true); else OCB_LA_LAUNCH(false, false); }
#undef OCB_LA_LAUNCH
    return 0;
}
