// Persistent, software-pipelined variants of the fused four-step kernels (N1 = 1024, one pol-pair).
//
// The one-wave kernels in fused_kernels.cuh put every warp of the chip in the same phase (load ->
// transform -> load -> pointwise -> transform -> store), so HBM/L2 idles while the SMs compute and
// the SMs idle while the loads are in flight.  Here a CTA owns a strided set of rows (k_time_p) or
// column tiles (k_freq_p) and, as soon as it has moved a staged stream of the CURRENT task from shared
// memory into registers, it starts the asynchronous copy (cp.async, 16 B per lane, L1-bypassing) of
// the same stream of its NEXT task into the freed buffer.  The copy then has the whole transform +
// pointwise time of the current task to land, and the only exposed memory latency is that of the
// first task of a CTA.  Arithmetic is bit-identical to the one-wave kernels (same operations in the
// same order), which tests/test_gpu_fused_engine.py checks.
//
// cp.async bookkeeping: every thread commits exactly NS groups per task, in consumption order
// (W, C, [P, H]); empty groups are committed where a thread has nothing to copy, so
// `cp.async.wait_group NS-1` always means "the oldest outstanding stream has landed".
#pragma once
#include "fused_kernels.cuh"

namespace ocb {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
}
// one warp copies `bytes` (multiple of 512) contiguous bytes global -> shared
template <int BYTES>
__device__ __forceinline__ void warp_copy_async(void* sdst, const void* gsrc, int lane) {
    static_assert(BYTES % 512 == 0, "whole 16-byte chunks per lane");
#pragma unroll
    for (int i = 0; i < BYTES / 512; ++i)
        cp_async16(reinterpret_cast<char*>(sdst) + (i * 32 + lane) * 16,
                   reinterpret_cast<const char*>(gsrc) + (i * 32 + lane) * 16);
}

// ------------------------------------------------------------------------------------------
// Time kernel, persistent.  CTA = 64 threads = the x and y warps of one time row (n2); rows are
// taken with stride gridDim.x.  Modes TM_FWD, TM_FIRST, TM_ITER of k_time<32, 2, MODE>.
//   staged streams : W row (all modes), C = A.aux0 row (FIRST: step-start field, ITER: previous
//                    iterate), and with SH: P = P_ch row (copied by the x warp, read by both) and
//                    H = E_hd row.  Without SH the latter two are read from L2 (prefetched).
// Shared memory per warp: exchange buffer (aliased by the |E|^2 row between the transforms),
// the U-table row, the staged rows.
// ------------------------------------------------------------------------------------------
template <int MODE, bool SH>
struct TimePipeCfg {
    static constexpr int N1 = 1024, GBUF = 32 * 33;
    static constexpr bool HAS_C = (MODE == TM_FIRST || MODE == TM_ITER);
    static constexpr bool HAS_PH = (MODE == TM_ITER) && SH;
    static constexpr int NS = 1 + (HAS_C ? 1 : 0) + (HAS_PH ? 2 : 0);  // cp.async groups per task
    static constexpr int ROWS = 1 + (HAS_C ? 1 : 0) + (HAS_PH ? 1 : 0);  // staged complex rows per warp
    static constexpr int WARP_FLOATS = 2 * GBUF + 64 + ROWS * 2 * N1;
    static constexpr int SMEM_BYTES = (2 * WARP_FLOATS + (HAS_PH ? N1 : 0)) * 4;
};

template <int MODE, bool SH>
__global__ void __launch_bounds__(64)
k_time_p(const TimeArgs A) {
    using namespace fft;
    using Cfg = TimePipeCfg<MODE, SH>;
    constexpr int Q1 = 32, N1 = Cfg::N1, GBUF = Cfg::GBUF, NS = Cfg::NS;
    static_assert(MODE == TM_FWD || MODE == TM_FIRST || MODE == TM_ITER, "pipelined modes");
    if constexpr (MODE == TM_ITER) {
        if (A.ext.mail && *reinterpret_cast<volatile long long*>(A.ext.converged_step) == A.ext.step_id) return;
    }
    extern __shared__ __align__(16) float smem_p[];
    const int t = threadIdx.x & 31, pol = threadIdx.x >> 5;
    float* mine = smem_p + pol * Cfg::WARP_FLOATS;
    float* xr = mine;
    float* xi = xr + GBUF;
    float2* Us = reinterpret_cast<float2*>(mine + 2 * GBUF);        // U-table row of the current task
    float2* sW = reinterpret_cast<float2*>(mine + 2 * GBUF + 64);   // staged W row
    float2* sC = sW + N1;                                           // staged aux0 row
    float2* sH = sC + N1;                                           // staged E_hd row (SH)
    float* sP = smem_p + 2 * Cfg::WARP_FLOATS;                      // staged P_ch row (SH, shared by both warps)
    float* pown = xr;                                               // |E|^2 row aliases the exchange buffer
    const float* poth = smem_p + (pol ^ 1) * Cfg::WARP_FLOATS;
    auto wsync = [] { __syncwarp(); };
    const float2* tw = A.tw;
    const int N2 = A.N2, stride = gridDim.x;

    auto issue_W = [&](int r) { warp_copy_async<N1 * 8>(sW, A.in + (int64_t)pol * A.N + (int64_t)r * N1, t); };
    auto issue_C = [&](int r) { warp_copy_async<N1 * 8>(sC, A.aux0 + (int64_t)pol * A.N + (int64_t)r * N1, t); };
    auto issue_H = [&](int r) { warp_copy_async<N1 * 8>(sH, A.ehd + (int64_t)pol * A.N + (int64_t)r * N1, t); };
    auto issue_P = [&](int r) { if (pol == 0) warp_copy_async<N1 * 4>(sP, A.pch + (int64_t)r * N1, t); };
    auto prefetch_direct = [&](int r) {  // !SH: E_hd and P_ch rows of task r -> L2
        if constexpr (MODE == TM_ITER && !SH) {
            const int64_t b = (int64_t)pol * A.N + (int64_t)r * N1;
            for (int l = t; l < N1 * 8 / 128; l += 32) prefetch_l2(reinterpret_cast<const char*>(A.ehd + b) + l * 128);
            if (pol == 0) prefetch_l2(reinterpret_cast<const char*>(A.pch + (int64_t)r * N1) + t * 128);
        }
    };

    double d_num = 0.0, d_den = 0.0;  // per-row float sums (as in the one-wave kernel) folded in double
    float s_max = 0.f;
    int row = blockIdx.x;
    if (row < N2) {
        issue_W(row); cp_async_commit();
        if constexpr (Cfg::HAS_C) { issue_C(row); cp_async_commit(); }
        if constexpr (Cfg::HAS_PH) { issue_P(row); cp_async_commit(); issue_H(row); cp_async_commit(); }
        prefetch_direct(row);
        float2 wV = __ldg(A.tabV + (int64_t)row * 32 + t);
        Us[t] = __ldg(A.tabU + (int64_t)row * Q1 + t);
        __syncwarp();

        for (;;) {
            const int nxt = row + stride;
            const bool has_next = nxt < N2;
            float2 wVn = wV, un = make_float2(0.f, 0.f);
            if (has_next) {
                wVn = __ldg(A.tabV + (int64_t)nxt * 32 + t);
                un = __ldg(A.tabU + (int64_t)nxt * Q1 + t);
                prefetch_direct(nxt);
            }
            const int64_t base = (int64_t)pol * A.N + (int64_t)row * N1;

            // ---- enter: staged row -> registers (+ conjugate inter-pass twiddle and inverse transform) ----
            float2 v[32];
            cp_async_wait<NS - 1>();
            __syncwarp();
#pragma unroll
            for (int s = 0; s < 32; ++s) v[s] = sW[s * Q1 + t];
            if constexpr (MODE != TM_FWD) {
                static_for<0, Q1>([&](auto kk) {
                    constexpr int KQ = decltype(kk)::value, SLOT = brev<Q1>(KQ);
                    const float2 w = cmul(wV, Us[KQ]);
                    v[SLOT] = cmul_conj(v[SLOT], w);  // conj twiddle W_N^{-n2 k1}
                });
            }
            __syncwarp();  // every lane has consumed its part of sW
            if (has_next) issue_W(nxt);
            cp_async_commit();
            if constexpr (MODE != TM_FWD) {
                coop_fft_inverse<Q1, 1, 1>(v, xr, xi, tw, t, 0, wsync);  // v[a'] = sample n1 = Q1*a' + t
                __syncwarp();  // exchange buffer free -> may be reused as the |E|^2 row
            }

            // ---- pointwise stage (channels.py:388-390, 414-417, 424, 436, 493, 517-519) ----
            if constexpr (MODE == TM_FIRST) {
                float2* ehd_out = A.aux1 + base;
                cp_async_wait<NS - 1>();
                __syncwarp();
#pragma unroll
                for (int a = 0; a < 32; ++a) {
                    st_stream(ehd_out + Q1 * a + t, v[a]);
                    pown[Q1 * a + t] = cabs2(sC[Q1 * a + t]);
                }
                __syncwarp();
                if (has_next) issue_C(nxt);
                cp_async_commit();
                __syncthreads();
                float* pch = A.pch + (int64_t)row * N1;
#pragma unroll
                for (int a = 0; a < 32; ++a) {
                    const float P = pown[Q1 * a + t] + poth[Q1 * a + t];
                    if (pol == 0) st_stream(pch + Q1 * a + t, P);
                    v[a] = cmul(v[a], phase_rot(A.cphi * P));
                }
                __syncthreads();  // both warps are done with the |E|^2 rows before the transform reuses them
            }
            if constexpr (MODE == TM_ITER) {
                float2* ec_new = A.aux1 + base;
                float s_num = 0.f, s_den = 0.f;
                cp_async_wait<NS - 1>();
                __syncwarp();
#pragma unroll
                for (int a = 0; a < 32; ++a) {
                    const float2 e = sC[Q1 * a + t];
                    s_num += cabs2(make_float2(v[a].x - e.x, v[a].y - e.y));  // channels.py:517
                    s_den += cabs2(e);
                    st_stream(ec_new + Q1 * a + t, v[a]);
                    pown[Q1 * a + t] = cabs2(v[a]);
                }
                d_num += (double)s_num;
                d_den += (double)s_den;
                __syncwarp();
                if (has_next) issue_C(nxt);
                cp_async_commit();
                if constexpr (SH) cp_async_wait<NS - 2>();  // P and H (the two oldest groups) have landed
                __syncthreads();                            // ... and are visible to both warps
                const float* pch = A.pch + (int64_t)row * N1;
                const float2* ehd = A.ehd + base;
#pragma unroll
                for (int a = 0; a < 32; ++a) {
                    const float P = pown[Q1 * a + t] + poth[Q1 * a + t];
                    if (pol == 0) s_max = fmaxf(s_max, P);
                    float pc;
                    float2 h;
                    if constexpr (SH) { pc = sP[Q1 * a + t]; h = sH[Q1 * a + t]; }
                    else { pc = ld_stream(pch + Q1 * a + t); h = ld_stream(ehd + Q1 * a + t); }
                    v[a] = cmul(h, phase_rot(A.cphi * (pc + P)));  // channels.py:436, 414-417
                }
                __syncthreads();
                if constexpr (SH) {
                    if (has_next) issue_P(nxt);
                    cp_async_commit();
                    if (has_next) issue_H(nxt);
                    cp_async_commit();
                }
            }

            // ---- leave: forward FFT over n1 + inter-pass twiddle -> W row ----
            coop_fft_forward<Q1, 1, 1>(v, xr, xi, tw, t, 0, wsync);
            {
                float2* dst = A.out + base;
                static_for<0, Q1>([&](auto kk) {
                    constexpr int KQ = decltype(kk)::value, SLOT = brev<Q1>(KQ);
                    const float2 w = cmul(wV, Us[KQ]);
                    st_stream(dst + SLOT * Q1 + t, cmul(v[SLOT], w));
                });
            }
            if (!has_next) break;
            __syncwarp();
            wV = wVn;
            Us[t] = un;
            __syncwarp();
            row = nxt;
        }
        cp_async_wait<0>();
    }
    if constexpr (MODE == TM_ITER) block_reduce3_finalize(d_num, d_den, s_max, A.partials, A.sums, A.ticket, &A.ext);
}

}  // namespace ocb
