// Time pass of the fused four-step engine for N1 = 1024 (N = 2^19, 2^20): persistent, bulk-copy fed.
//
// Same arithmetic as k_time<32, NP, MODE> (fused_kernels.cuh) — one warp per (time row n2, polarisation), 32 samples
// per lane — but a different data path, built around what bounded the one-wave kernel (ncu, round 1: 0.86 waves,
// long-scoreboard stalls on a chain of four dependent global round trips per row):
//
//  * One 256-thread CTA per SM (8 warps = 4 row pairs at a time for the dual-pol modes), looping over its rows.
//  * Every operand that does NOT depend on the preceding kernel of the stream — the previous iterate E_c, the
//    half-dispersed field E_hd and the power row P_ch (all written two or more kernels earlier) — is fetched by the
//    bulk-copy engine (cp.async.bulk global -> shared, completion on an mbarrier; one instruction per 8 KB row),
//    for the first row BEFORE griddepcontrol.wait, i.e. while the preceding k_freq is still draining, and for every
//    later row while the current row is being transformed.  Only the W row (written by k_freq) is loaded the
//    classic way, straight into registers.
//  * The twiddle table (hi and lo parts, 16 KB) lives in shared memory; the exchange of the cooperative
//    1024-point transform runs in two half-footprint rounds, which leaves room for the landing buffers.
//  * |E|^2 of a row is handed to the other polarisation's warp through the (then idle) exchange buffer.
//  * The warps of a CTA run the row body in lockstep (CTA-wide barrier at every phase boundary): the body is ~130 KB of
//    fully unrolled code, and free-running warps thrash the instruction cache.
//  * Convergence sums are accumulated across the rows of a CTA and posted once per CTA (148 partials).
//
// Shared memory per CTA: 8 x (8 KB E_c + 8 KB E_hd + 4.1 KB exchange) + 4 x 4 KB P_ch + 16 KB table = 193 KB.
//
// Reference formulas: optic/models/channels.py:388-390, 406-421, 424, 436, 493, 517-519 (manakovSSF), :219-229
// (ssfm), optic/dsp/equalization.py:1077, 1129 (DBP signs) — see fused_kernels.cuh.
#pragma once
#include "fused_kernels.cuh"

namespace ocb {

struct TimeBulkCfg {
    static constexpr int N1 = 1024, WARPS = 8;
    static constexpr int TW_ENTRIES = (fft::kLO ? 2 : 1) * 1024;       // symmetric 32 x 32 table: hi (+ lo)
    static constexpr int XBUF_FLOATS = 32 * 33;                        // half-footprint exchange, one plane
    static constexpr int ROW_BYTES = N1 * 8;
    static constexpr int OFF_TW = 0;
    static constexpr int OFF_A = OFF_TW + TW_ENTRIES * 8;              // [WARPS][N1] float2   E_c / Ech landing
    static constexpr int OFF_B = OFF_A + WARPS * ROW_BYTES;            // [WARPS][N1] float2   E_hd landing
    static constexpr int OFF_P = OFF_B + WARPS * ROW_BYTES;            // [WARPS/2][N1] float  P_ch landing
    static constexpr int OFF_X = OFF_P + (WARPS / 2) * N1 * 4;         // [WARPS][XBUF_FLOATS] exchange / |E|^2 hand-over
    static constexpr int OFF_BAR = OFF_X + WARPS * XBUF_FLOATS * 4;    // [WARPS][3] mbarriers
    static constexpr int SMEM_BYTES = OFF_BAR + WARPS * 3 * 8;
};

template <int NP, int MODE>
__global__ void __launch_bounds__(256, 1)
k_time_bulk(const TimeArgs A) {
    using namespace fft;
    using Cfg = TimeBulkCfg;
    constexpr int Q1 = 32, N1 = 1024, S = Cfg::WARPS / NP;  // S row units in flight per CTA
    constexpr bool kManakov = (MODE == TM_FIRST || MODE == TM_ITER || MODE == TM_ITERF || MODE == TM_ROT);
    constexpr bool kSums = (MODE == TM_ITER || MODE == TM_ITERF);
    constexpr bool kUseA = kManakov;                                  // E_c (ITER, ITERF, ROT) or Ech (FIRST)
    constexpr bool kUseB = (MODE == TM_ITER || MODE == TM_ROT);       // E_hd and P_ch
    static_assert(!kManakov || NP == 2, "Manakov modes need both polarisations of a row in the CTA");
    extern __shared__ __align__(128) unsigned char smem[];
    float2* tws = reinterpret_cast<float2*>(smem + Cfg::OFF_TW);
    const float2* tws_lo = tws + 1024;
    const int tid = threadIdx.x, warp = tid >> 5, t = tid & 31;
    const int slot = warp / NP, pol = warp % NP;
    float2* bufA = reinterpret_cast<float2*>(smem + Cfg::OFF_A) + warp * N1;
    float2* bufB = reinterpret_cast<float2*>(smem + Cfg::OFF_B) + warp * N1;
    float* bufP = reinterpret_cast<float*>(smem + Cfg::OFF_P) + (NP == 2 ? slot : 0) * N1;
    float* xr = reinterpret_cast<float*>(smem + Cfg::OFF_X) + warp * Cfg::XBUF_FLOATS;
    const float* xr_other = reinterpret_cast<const float*>(smem + Cfg::OFF_X) + (warp ^ 1) * Cfg::XBUF_FLOATS;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* barA = bars + warp * 3;                                  // own E_c row
    uint64_t* barB = bars + warp * 3 + 1;                              // own E_hd row
    uint64_t* barP = bars + (NP == 2 ? (slot * NP) : warp) * 3 + 2;    // the pair's P_ch row (owned by the pol-0 warp)
    auto wsync = [] { __syncwarp(); };

    // rows of this warp: unit u = blockIdx.x + gridDim.x * (slot + S * k), k = 0, 1, ...
    const int units = A.N2;
    const int first_unit = blockIdx.x + gridDim.x * slot;
    const int unit_stride = gridDim.x * S;

    auto issue_prefetch = [&](int unit) {  // one lane per warp; only reads arrays older than the preceding kernel
        const int64_t base = (int64_t)pol * A.N + (int64_t)unit * N1;
        if constexpr (kUseA) {
            mbar_expect_tx(barA, Cfg::ROW_BYTES);
            bulk_g2s(bufA, A.aux0 + base, Cfg::ROW_BYTES, barA);
        }
        if constexpr (kUseB) {
            mbar_expect_tx(barB, Cfg::ROW_BYTES);
            bulk_g2s(bufB, A.ehd + base, Cfg::ROW_BYTES, barB);
            if (pol == 0) {
                mbar_expect_tx(barP, N1 * 4);
                bulk_g2s(bufP, A.pch + (int64_t)unit * N1, N1 * 4, barP);
            }
        }
    };

    // ---- prologue: nothing here touches data of the preceding kernel ---------------------------------------------
    if (t == 0) {
        mbar_init(barA, 1);
        mbar_init(barB, 1);
        if (pol == 0) mbar_init(barP, 1);
        mbar_fence_init();
        if (kManakov && first_unit < units) issue_prefetch(first_unit);
    }
    for (int i = tid; i < 1024; i += 256) {
        tws[i] = __ldg(A.tw + i);
        if constexpr (kLO) tws[1024 + i] = __ldg(A.tw + 64 * Q1 + i);
    }
    __syncthreads();  // table staged, barriers initialised

    pdl_wait();
    pdl_launch_dependents();
    bool skip = false;
    if (A.need_flag && *reinterpret_cast<const volatile long long*>(A.need_flag) != A.need_id) skip = true;
    if constexpr (kSums) {
        // speculative launch of an iteration whose predecessor already converged: nothing to do
        if (A.ext.mail && *reinterpret_cast<volatile long long*>(A.ext.converged_step) == A.ext.step_id) skip = true;
    }
    if (skip) {  // the bulk copies issued above must land before the CTA (and its shared memory) goes away
        if (kManakov && first_unit < units) {
            if constexpr (kUseA) mbar_wait(barA, 0);
            if constexpr (kUseB) { mbar_wait(barB, 0); mbar_wait(barP, 0); }
        }
        return;
    }

    // All warps of the CTA walk through the (large, fully unrolled) row body in LOCKSTEP: a CTA-wide barrier at every
    // phase boundary keeps them in the same code region, so the instruction cache streams the body once per round
    // instead of once per warp (ncu on the free-running version: no_instruction stalls 2.7 per issue).
    auto phase_sync = [] { __syncthreads(); };
    float s_num = 0.f, s_den = 0.f, s_max = 0.f;
    unsigned parity = 0;
    const int rounds = (units - (int)blockIdx.x + unit_stride - 1) / unit_stride;  // rounds of slot 0 (the longest)
    for (int rnd = 0; rnd < rounds; ++rnd, parity ^= 1u) {
        const int unit = first_unit + rnd * unit_stride;
        phase_sync();
        if (unit >= units) {  // idle slot in the last round: keep the barrier count of the active warps
            if constexpr (MODE != TM_FWD && MODE != TM_ROT) phase_sync();
            if constexpr (kManakov && MODE != TM_ITERF) { phase_sync(); phase_sync(); }
            continue;
        }
        const int row = unit;
        const int64_t base = (int64_t)pol * A.N + (int64_t)row * N1;
        const int next_unit = unit + unit_stride;
        const float2 wV = __ldg(A.tabV + (int64_t)row * 32 + t);
        const float2 wVl = kLO ? __ldg(A.tabV + (int64_t)(A.N2 + row) * 32 + t) : float2{};
        const float2* Urow = A.tabU + (int64_t)row * Q1;
        const float2* Urow_lo = A.tabU + (int64_t)(A.N2 + row) * Q1;
        float2 v[32];

        // ---- enter ----------------------------------------------------------------------------------------------
        if constexpr (MODE == TM_FWD) {
            const float2* src = A.in + base;
#pragma unroll
            for (int a = 0; a < 32; ++a) v[a] = ld_stream(src + Q1 * a + t);
        } else if constexpr (MODE != TM_ROT) {
            const float2* src = A.in + base;
#pragma unroll
            for (int s = 0; s < 32; ++s) v[s] = ld_stream(src + s * Q1 + t);
            static_for<0, Q1>([&](auto kk) {
                constexpr int KQ = decltype(kk)::value, SLOT = brev<Q1>(KQ);
                v[SLOT] = cmul_vu<true>(v[SLOT], wV, wVl, __ldg(Urow + KQ), kLO ? __ldg(Urow_lo + KQ) : float2{});  // W_N^{-n2 k1}
            });
            coop_fft_inverse<Q1, 1, 1, true>(v, xr, nullptr, tws, tws_lo, t, 0, wsync);  // v[a'] = sample n1 = 32 a' + t
            phase_sync();  // exchange buffer free again; lockstep
        }

        // ---- pointwise stage in the time domain -----------------------------------------------------------------
        if constexpr (MODE == TM_INV) {
            float2* dst = A.out + base;
#pragma unroll
            for (int a = 0; a < 32; ++a) st_stream(dst + Q1 * a + t, make_float2(v[a].x * A.out_scale, v[a].y * A.out_scale));
            continue;
        }
        if constexpr (MODE == TM_NLSE) {  // channels.py:225
#pragma unroll
            for (int a = 0; a < 32; ++a) v[a] = cmul(v[a], phase_rot(A.cphi * cabs2(v[a])));
        }
        if constexpr (kManakov) {
            float* pown = xr;                 // |E|^2 of this polarisation, read by the other warp of the row
            const float* poth = xr_other;
            mbar_wait(barA, parity);
            if constexpr (MODE == TM_FIRST) {
                // v = E_hd (store it); power of the step-start field Ech  (channels.py:388)
                float2* ehd_out = A.aux1 + base;
#pragma unroll
                for (int a = 0; a < 32; ++a) {
                    st_stream(ehd_out + Q1 * a + t, v[a]);
                    pown[Q1 * a + t] = cabs2(bufA[Q1 * a + t]);
                }
            } else if constexpr (MODE == TM_ROT) {
#pragma unroll
                for (int a = 0; a < 32; ++a) pown[Q1 * a + t] = cabs2(bufA[Q1 * a + t]);  // the iterate stored by TM_ITERF
            } else {
                // v = E_fd: convergence sums against the previous iterate (channels.py:517), stored as the new iterate
                // IN PLACE (aux1 may equal aux0: the bulk copy of this row has completed)
                float2* ec_new = A.aux1 + base;
#pragma unroll
                for (int a = 0; a < 32; ++a) {
                    const float2 e = bufA[Q1 * a + t];
                    s_num += cabs2(make_float2(v[a].x - e.x, v[a].y - e.y));
                    s_den += cabs2(e);
                    st_stream(ec_new + Q1 * a + t, v[a]);
                    if constexpr (MODE == TM_ITER) pown[Q1 * a + t] = cabs2(v[a]);
                }
            }
            float pc[(MODE == TM_ITER || MODE == TM_ROT) ? 32 : 1];
            if constexpr (kUseB) {
                mbar_wait(barB, parity);
                mbar_wait(barP, parity);
#pragma unroll
                for (int a = 0; a < 32; ++a) v[a] = bufB[Q1 * a + t];
#pragma unroll
                for (int a = 0; a < 32; ++a) pc[a] = bufP[Q1 * a + t];
            }
            if constexpr (MODE != TM_ITERF) {
                phase_sync();  // both polarisations of the row have written their |E|^2
                float* pch = A.pch + (int64_t)row * N1;
#pragma unroll
                for (int a = 0; a < 32; ++a) {
                    const float P = pown[Q1 * a + t] + poth[Q1 * a + t];
                    float ph;
                    if constexpr (MODE == TM_FIRST) {
                        if (pol == 0) st_stream(pch + Q1 * a + t, P);
                        ph = A.cphi * P;  // φ = (8/9)γ(P+P)/2, channels.py:390/493 with E_conv == Ech
                    } else {
                        s_max = fmaxf(s_max, P);
                        ph = A.cphi * (pc[a] + P);  // channels.py:436
                    }
                    v[a] = cmul(v[a], phase_rot(ph));  // channels.py:414-417
                }
                phase_sync();  // the landing buffers and both |E|^2 rows have been consumed
            } else {
                __syncwarp();
            }
            // landing buffers are free: start the next row's copies so that they overlap the forward transform
            if (next_unit < units && t == 0) {
                fence_proxy_async();
                issue_prefetch(next_unit);
            }
        }

        // ---- leave: forward FFT over n1 + inter-pass twiddle -> W row -----------------------------------------------
        coop_fft_forward<Q1, 1, 1, true>(v, xr, nullptr, tws, tws_lo, t, 0, wsync);
        {
            float2* dst = A.out + base;
            static_for<0, Q1>([&](auto kk) {
                constexpr int KQ = decltype(kk)::value, SLOT = brev<Q1>(KQ);
                st_stream(dst + SLOT * Q1 + t, cmul_vu<false>(v[SLOT], wV, wVl, __ldg(Urow + KQ), kLO ? __ldg(Urow_lo + KQ) : float2{}));
            });
        }
    }

    if constexpr (kSums) {
        if (MODE == TM_ITER && pol == 1) s_max = 0.f;  // both polarisation warps saw the same total power
        const unsigned my_ticket = block_reduce3_post(s_num, s_den, s_max, A.partials, A.ticket);
        block_reduce3_final(my_ticket, A.partials, A.sums, A.ticket, &A.ext);
    }
}

}  // namespace ocb
