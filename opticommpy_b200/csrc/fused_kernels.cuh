// Fused four-step FFT kernels of the split-step propagators (the sm_100a fast path, N = N1*N2,
// N1 = 32*Q1, N2 = 32*Q2, Q in {8,16,32}).
//
// A length-N transform of x[n], n = N2*n1 + n2, factors as
//     time pass : N1-point FFT over n1 for every n2, times W_N^{n2 k1}      -> W[n2][k1]
//     freq pass : N2-point FFT over n2 for every k1                          -> X[k1 + N1 k2]
// and the inverse mirrors it.  The engine keeps every time-domain field TRANSPOSED in HBM
// (index N1*n2 + n1), so that the time pass — the one that also carries every pointwise operation
// of the split-step loop and therefore streams up to six arrays — works on contiguous rows with one
// warp per row, while the strided pass (k_freq) only streams the W buffer:
//     k_time : [.. -> IFFT_N1 ->] pointwise stage [-> FFT_N1 -> twiddle]      (contiguous rows)
//     k_freq : FFT_N2 -> x linear operator -> IFFT_N2                         (column tiles, in place)
// One half step ifft(fft(.)*L) therefore costs three passes over the data and the Kerr rotation,
// power, convergence sums and max-power reduction ride inside k_time.
//
// W-row layout: position p = slot*Q1 + t holds k1 = (t*G1 + slot/Q1) + 32*brev<Q1>(slot % Q1);
// it is whatever order the cooperative FFT leaves in registers, so loads and stores of W rows are
// lane-contiguous.  k_freq and the operator table only need to agree on that order.
//
// Reference formulas restated here: optic/models/channels.py:388-390, 406-421, 424, 436, 493,
// 517-519 (manakovSSF), :219-229 (ssfm), optic/dsp/equalization.py:1077, 1129 (DBP signs).
#pragma once
#include "fft_core.cuh"
#include "ssfm_kernels.cuh"

namespace ocb {

// TM_ITERF: an iteration that the host predicts to be the last one of its step.  It does everything
//           TM_ITER does up to the convergence sums and the store of the new iterate, but then transforms
//           that iterate itself (the next step starts from it, channels.py:438-439 + :409) instead of the
//           re-rotated E_hd, which saves the separate TM_FWD pass of the next step.
// TM_ROT  : recovery when that prediction was wrong: rotate E_hd with the phase of the stored iterate
//           (channels.py:436, 414-417) and transform it, i.e. the second half of TM_ITER.
enum TimeMode { TM_FWD = 0, TM_INV = 1, TM_FIRST = 2, TM_ITER = 3, TM_NLSE = 4, TM_ITERF = 5, TM_ROT = 6 };

struct TimeArgs {
    const float2* in;     // TM_FWD: time-domain field ; others: W buffer
    float2* out;          // TM_INV: time-domain field ; others: W buffer
    const float2* aux0;   // TM_FIRST: step-start field Ech ; TM_ITER: previous iterate E_conv
    float2* aux1;         // TM_FIRST: E_hd (written)       ; TM_ITER: new iterate (written)
    const float2* ehd;    // TM_ITER: E_hd (read)
    float* pch;           // TM_FIRST: written ; TM_ITER: read
    const float2* tw;     // [32][Q1] exp(-2 pi i q ka / N1): hi, hi transposed, lo, lo transposed (32*Q1 entries each)
    const float2* tabV;   // [N2][32]  exp(-2 pi i n2 ka / N), followed by the lo parts [N2][32]
    const float2* tabU;   // [N2][Q1]  exp(-2 pi i n2 32 kq / N), followed by the lo parts [N2][Q1]
    double* partials;     // TM_ITER reduction scratch
    double* sums;
    unsigned* ticket;
    int64_t N;            // samples per polarisation row
    int N2;               // number of time rows (n2)
    float cphi;           // dir*hz*(8/9)γ (FIRST) | dir*hz*(8/9)γ/2 (ITER) | γ hz (NLSE)
    float out_scale;      // TM_INV: gain applied to the time-domain output
    FinalizeExt ext;      // TM_ITER: host mailbox + convergence flag (ext.mail == nullptr: unused)
    const long long* need_flag;  // speculative launch across a step boundary: run only if *need_flag == need_id
    long long need_id;
    int kernel_choice;    // host side only: 1 selects the bulk-copy-fed kernel where it exists (launch_time)
};

// ------------------------------------------------------------------------------------------
// Time kernel.  64-thread CTA = 64/Q1 row tasks; a task is (time row n2, polarisation) and is owned
// by Q1 threads (one warp when N1 = 1024).  With NP = 2 the x and y tasks of a row sit in the same
// CTA and share |E|^2 through shared memory.  Every global access is lane-contiguous.
// ------------------------------------------------------------------------------------------
template <int Q1, int NP, int MODE>
__global__ void __launch_bounds__(64, 7)
k_time(const TimeArgs A) {
    using namespace fft;
    constexpr int G = 32 / Q1, N1 = 32 * Q1, TASKS = 64 / Q1;
    constexpr int STR = Q1 + 1, GBUF = 32 * STR + (Q1 < 32 ? Q1 : 0);
    __shared__ float xbuf[TASKS * 2 * GBUF];
    constexpr bool kManakov = (MODE == TM_FIRST || MODE == TM_ITER || MODE == TM_ITERF || MODE == TM_ROT);
    constexpr bool kSums = (MODE == TM_ITER || MODE == TM_ITERF);
    __shared__ float pbuf[kManakov ? TASKS * N1 : 1];

    const int tid = threadIdx.x, grp = tid / Q1, t = tid % Q1;
    const int task = blockIdx.x * TASKS + grp;
    const int row = task / NP, pol = task % NP;
    float* xr = xbuf + grp * 2 * GBUF;
    float* xi = xr + GBUF;
    const int64_t base = (int64_t)pol * A.N + (int64_t)row * N1;
    auto wsync = [] { __syncwarp(); };
    const float2* tw = A.tw;  // 8 KB table, L1-resident

    float2 v[32];
    float2 wV[G], wVl[G];  // hi / lo parts (double-single twiddles, fft_core.cuh)
#pragma unroll
    for (int g = 0; g < G; ++g) {
        wV[g] = __ldg(A.tabV + (int64_t)row * 32 + t * G + g);
        wVl[g] = kLO ? __ldg(A.tabV + (int64_t)(A.N2 + row) * 32 + t * G + g) : float2{};
    }
    const float2* Urow = A.tabU + (int64_t)row * Q1;
    const float2* Urow_lo = A.tabU + (int64_t)(A.N2 + row) * Q1;
    // inter-pass twiddle W_N^{n2 k1} = V[n2][ka] U[n2][kq], applied factor by factor
    auto twiddle_fwd = [&](float2 x, int gi, int kq) {
        return cmul_vu<false>(x, wV[gi], wVl[gi], __ldg(Urow + kq), kLO ? __ldg(Urow_lo + kq) : float2{});
    };
    auto twiddle_inv = [&](float2 x, int gi, int kq) {
        return cmul_vu<true>(x, wV[gi], wVl[gi], __ldg(Urow + kq), kLO ? __ldg(Urow_lo + kq) : float2{});
    };

    // Secondary streams of the pointwise stage: start their HBM->L2 fetch now so that it overlaps the
    // W-row load and the inverse transform (one 128-byte line per lane and trip).
    if constexpr (kManakov) {
        constexpr int LINES = N1 * 8 / 128;
        for (int l = t; l < LINES; l += Q1) {
            prefetch_l2(reinterpret_cast<const char*>(A.aux0 + base) + l * 128);
            if constexpr (MODE == TM_ITER || MODE == TM_ROT) prefetch_l2(reinterpret_cast<const char*>(A.ehd + base) + l * 128);
        }
        if constexpr (MODE == TM_ITER || MODE == TM_ROT) {
            if (pol == 0)
                for (int l = t; l < LINES / 2; l += Q1)
                    prefetch_l2(reinterpret_cast<const char*>(A.pch + (int64_t)row * N1) + l * 128);
        }
    }

    // Everything above only touched tables and issued prefetch hints; the field data below may have been
    // written by the preceding kernel of the stream (programmatic dependent launch).
    pdl_wait();
    pdl_launch_dependents();
    if (A.need_flag && *reinterpret_cast<const volatile long long*>(A.need_flag) != A.need_id) return;
    if constexpr (kSums) {
        // speculative launch of an iteration whose predecessor already converged: nothing to do
        if (A.ext.mail && *reinterpret_cast<volatile long long*>(A.ext.converged_step) == A.ext.step_id) return;
    }

    // ---- enter ----------------------------------------------------------------------------------
    if constexpr (MODE == TM_FWD) {
        const float2* src = A.in + base;
#pragma unroll
        for (int a = 0; a < 32; ++a) v[a] = ld_stream(src + Q1 * a + t);
    } else if constexpr (MODE != TM_ROT) {
        const float2* src = A.in + base;
#pragma unroll
        for (int s = 0; s < 32; ++s) v[s] = ld_stream(src + s * Q1 + t);
        static_for<0, G>([&](auto gg) {
            constexpr int GI = decltype(gg)::value;
            static_for<0, Q1>([&](auto kk) {
                constexpr int KQ = decltype(kk)::value, SLOT = GI * Q1 + brev<Q1>(KQ);
                v[SLOT] = twiddle_inv(v[SLOT], GI, KQ);  // conj twiddle W_N^{-n2 k1}
            });
        });
        coop_fft_inverse<Q1, 1, 1>(v, xr, xi, tw + 32 * Q1, tw + 96 * Q1, t, 0, wsync);  // v[a'] = sample n1 = Q1*a' + t
        __syncwarp();
    }

    // ---- pointwise stage in the time domain ---------------------------------------------------------
    float s_num = 0.f, s_den = 0.f, s_max = 0.f;
    if constexpr (MODE == TM_INV) {
        float2* dst = A.out + base;
#pragma unroll
        for (int a = 0; a < 32; ++a) st_stream(dst + Q1 * a + t, make_float2(v[a].x * A.out_scale, v[a].y * A.out_scale));
        return;
    }
    if constexpr (MODE == TM_NLSE) {  // channels.py:225
#pragma unroll
        for (int a = 0; a < 32; ++a) v[a] = cmul(v[a], phase_rot(A.cphi * cabs2(v[a])));
    }
    if constexpr (kManakov) {
        static_assert(NP == 2, "Manakov modes need both polarisations in the CTA");
        float* pown = pbuf + grp * N1;
        const float* poth = pbuf + (grp ^ 1) * N1;  // the other polarisation of the same row
        if constexpr (MODE == TM_FIRST) {
            // v = E_hd (store it); power of the step-start field Ech  (channels.py:388)
            float2* ehd_out = A.aux1 + base;
            const float2* ech = A.aux0 + base;
            // two batches of 16 loads, each issued together ahead of the stores that would otherwise pace them
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
                float2 e[16];
#pragma unroll
                for (int a = 0; a < 16; ++a) e[a] = ld_stream_pinned(ech + Q1 * (16 * hb + a) + t);
#pragma unroll
                for (int a = 0; a < 16; ++a) {
                    st_stream(ehd_out + Q1 * (16 * hb + a) + t, v[16 * hb + a]);
                    pown[Q1 * (16 * hb + a) + t] = cabs2(e[a]);
                }
            }
        } else if constexpr (MODE == TM_ROT) {
            const float2* ec = A.aux0 + base;  // the iterate stored by TM_ITERF
#pragma unroll
            for (int a = 0; a < 32; ++a) pown[Q1 * a + t] = cabs2(ld_stream(ec + Q1 * a + t));
        } else {
            // v = E_fd: convergence sums against the previous iterate, store as the new iterate.  aux0 and
            // aux1 may be the SAME buffer (in-place update keeps the working set L2-sized): every thread
            // reads its own samples before it overwrites them; two batches of 16 keep the loads in flight.
            const float2* ec = A.aux0 + base;
            float2* ec_new = A.aux1 + base;
            constexpr int EB = 16;
#pragma unroll
            for (int hb = 0; hb < 32 / EB; ++hb) {
                float2 e[EB];
#pragma unroll
                for (int a = 0; a < EB; ++a) e[a] = ld_stream_ordered(ec + Q1 * (EB * hb + a) + t);
#pragma unroll
                for (int a = 0; a < EB; ++a) {
                    const int aa = EB * hb + a;
                    s_num += cabs2(make_float2(v[aa].x - e[a].x, v[aa].y - e[a].y));  // channels.py:517
                    s_den += cabs2(e[a]);
                    st_stream(ec_new + Q1 * aa + t, v[aa]);
                    if constexpr (MODE == TM_ITER) pown[Q1 * aa + t] = cabs2(v[aa]);
                }
            }
        }
        float* pch = A.pch + (int64_t)row * N1;
        float pc[(MODE == TM_ITER || MODE == TM_ROT) ? 32 : 1];
        if constexpr (MODE == TM_ITER || MODE == TM_ROT) {
            // E_fd has been stored, so v[] is free: fetch the whole E_hd and P_ch rows in one batch (64
            // loads in flight per thread) before the barrier.  The rotation loop below contains a branch
            // (large-phase path), which would otherwise make every trip wait for its own two loads.
            const float2* ehd = A.ehd + base;
#pragma unroll
            for (int a = 0; a < 32; ++a) v[a] = ld_stream_pinned(ehd + Q1 * a + t);
#pragma unroll
            for (int a = 0; a < 32; ++a) pc[a] = ld_stream_pinned(pch + Q1 * a + t);
        }
        // TM_ITERF: v stays E_fd, the field the next step starts from.  (It runs in fixed-step mode only, where
        // the maximum power — used for the adaptive step size, channels.py:392-397 — is not needed.)
        if constexpr (MODE != TM_ITERF) __syncthreads();
#pragma unroll
        for (int a = 0; a < (MODE == TM_ITERF ? 0 : 32); ++a) {
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
    }

    // The convergence sums are final here: post them now, so that the ticket's fence + atomic round trip
    // overlaps the forward transform and the stores below.
    unsigned my_ticket = 0;
    if constexpr (MODE == TM_ITER) {
        if (pol == 1) s_max = 0.f;  // both polarisation tasks saw the same total power
        my_ticket = block_reduce3_post(s_num, s_den, s_max, A.partials, A.ticket);
    }

    // ---- leave: forward FFT over n1 + inter-pass twiddle -> W row ---------------------------------------
    coop_fft_forward<Q1, 1, 1>(v, xr, xi, tw, tw + 64 * Q1, t, 0, wsync);
    // TM_ITERF has just issued its 32 stores of the new iterate: post after the transform, when they have
    // drained, so that the ticket's fence does not wait for them
    if constexpr (MODE == TM_ITERF) my_ticket = block_reduce3_post(s_num, s_den, 0.f, A.partials, A.ticket);
    {
        float2* dst = A.out + base;
        static_for<0, G>([&](auto gg) {
            constexpr int GI = decltype(gg)::value;
            static_for<0, Q1>([&](auto kk) {
                constexpr int KQ = decltype(kk)::value, SLOT = GI * Q1 + brev<Q1>(KQ);
                st_stream(dst + SLOT * Q1 + t, twiddle_fwd(v[SLOT], GI, KQ));
            });
        });
    }
    if constexpr (kSums) block_reduce3_final(my_ticket, A.partials, A.sums, A.ticket, &A.ext);
}

// ------------------------------------------------------------------------------------------
// Frequency kernel: for a tile of C adjacent W positions p (all N2 rows of one polarisation):
// FFT_N2 over n2 -> x LP -> IFFT_N2, in place.  Thread (q, c) owns rows n2 = Q2*a + q of column
// p0 + c; a warp touches 32/C rows x C*8 contiguous bytes per access.
// LP is stored in consumption order: LP[((tile*32 + slot)*Q2 + q)*C + c], see k_tab_linop_perm.
// ------------------------------------------------------------------------------------------
// Shared memory of one k_freq CTA: the (half-footprint) exchange array, the twiddle table(s) and the tile's
// whole operator slice, which the TMA engine fetches (cp.async.bulk) BEFORE the kernel waits for its predecessor.
template <int Q2, int C>
struct FreqCfg {
    static constexpr int STR = Q2 * C + C;
    static constexpr int TW_HALF = (Q2 == 32 ? 1 : 2) * 32 * Q2;  // a 32 x 32 table is symmetric: one copy
    static constexpr int TW_ENTRIES = (fft::kLO ? 2 : 1) * TW_HALF;   // hi parts, then lo parts
    static constexpr int LP_ENTRIES = 32 * Q2 * C;
    static constexpr int OFF_BAR = 32 * STR * 4 + TW_ENTRIES * 8 + LP_ENTRIES * 8;  // one mbarrier (operator slice landed)
    static constexpr int SMEM_BYTES = OFF_BAR + 16;
};
template <int Q2, int C>
__global__ void __launch_bounds__(Q2* C, (Q2 * C <= 256) ? 2 : 1)
k_freq(float2* __restrict__ W, const float2* __restrict__ LP, const float2* __restrict__ tw, int N1,
       const long long* __restrict__ converged_step, long long step_id,
       const long long* __restrict__ need_flag, long long need_id) {
    using namespace fft;
    using Cfg = FreqCfg<Q2, C>;
    constexpr int STR = Cfg::STR, NT = Q2 * C;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* xr = reinterpret_cast<float*>(smem_raw);  // [32*STR] real parts, then imaginary parts (two rounds)
    // twiddle table staged in shared memory: 63 table reads per thread sit inside the dependent transform
    // chains, and a shared-memory read is both faster than an L1 hit and immune to eviction by the field data
    float2* tws = reinterpret_cast<float2*>(xr + 32 * STR);
    float2* lps = tws + Cfg::TW_ENTRIES;             // operator slice of this tile, consumption order
    const int tid = threadIdx.x, q = tid / C, c = tid % C;
    const int tiles_per_pol = N1 / C;
    const int pol = blockIdx.x / tiles_per_pol, tile = blockIdx.x % tiles_per_pol;
    float2* base = W + ((int64_t)pol * (32 * Q2)) * N1 + tile * C + c;  // row n2 = 0 of this column
    auto bsync = [] { __syncthreads(); };

    // Neither the operator table nor the twiddle table depends on the preceding kernel: fetch them now, so that
    // the copies overlap the predecessor's tail (programmatic dependent launch) and this kernel's own W loads.  The
    // tile's operator slice is one contiguous block (consumption order): ONE bulk copy by the TMA engine (cp.async.bulk,
    // completion on an mbarrier) instead of 16 cp.async per thread.
    uint64_t* lp_bar = reinterpret_cast<uint64_t*>(smem_raw + Cfg::OFF_BAR);
    if (tid == 0) {
        mbar_init(lp_bar, 1);
        mbar_fence_init();
        constexpr unsigned kChunk = 32768;  // keep every copy well inside the per-instruction size limit
        mbar_expect_tx(lp_bar, Cfg::LP_ENTRIES * 8);
        const char* src = reinterpret_cast<const char*>(LP + (int64_t)tile * Cfg::LP_ENTRIES);
        for (unsigned off = 0; off < (unsigned)Cfg::LP_ENTRIES * 8; off += kChunk) {
            const unsigned nb = min(kChunk, (unsigned)Cfg::LP_ENTRIES * 8 - off);
            bulk_g2s(reinterpret_cast<char*>(lps) + off, src + off, nb, lp_bar);
        }
    }
    // global table: hi, hi transposed, lo, lo transposed (32*Q2 entries each)
    for (int i = tid; i < Cfg::TW_HALF; i += NT) {
        tws[i] = __ldg(tw + i);
        if constexpr (kLO) tws[Cfg::TW_HALF + i] = __ldg(tw + 64 * Q2 + i);
    }
    const float2* twt = (Q2 == 32) ? tws : tws + 32 * Q2;
    const float2* tws_lo = tws + Cfg::TW_HALF;
    const float2* twt_lo = (Q2 == 32) ? tws_lo : tws_lo + 32 * Q2;

    pdl_wait();  // W was written by the preceding time pass (programmatic dependent launch)
    pdl_launch_dependents();
    if ((converged_step && *reinterpret_cast<const volatile long long*>(converged_step) == step_id) ||
        (need_flag && *reinterpret_cast<const volatile long long*>(need_flag) != need_id)) {
        __syncthreads();          // barrier initialisation visible
        mbar_wait(lp_bar, 0);     // the bulk copy must land before the shared memory goes away
        return;
    }
    float2 v[32];
#pragma unroll
    for (int a = 0; a < 32; ++a) v[a] = ld_stream(base + (int64_t)(Q2 * a + q) * N1);
    __syncthreads();  // twiddle table staged, barrier initialisation visible
    mbar_wait(lp_bar, 0);  // operator slice landed
    coop_fft_forward<Q2, C, C, true, false>(v, xr, nullptr, tws, tws_lo, q, c, bsync);
#pragma unroll
    for (int s = 0; s < 32; ++s) v[s] = cmul(v[s], lps[s * NT + tid]);
    __syncthreads();
    coop_fft_inverse<Q2, C, C, true, false>(v, xr, nullptr, twt, twt_lo, q, c, bsync);
#pragma unroll
    for (int a = 0; a < 32; ++a) st_stream(base + (int64_t)(Q2 * a + q) * N1, v[a]);
}

// ------------------------------------------------------------------------------------------
// Layout conversion: planar natural order [r][N2*n1 + n2]  <->  engine order [r][N1*n2 + n1]
// (a tiled matrix transpose per row r).  in: (rows_in x cols_in) -> out: (cols_in x rows_in).
// ------------------------------------------------------------------------------------------
__global__ void k_transpose(const float2* __restrict__ in, float2* __restrict__ out, int rows_in, int cols_in,
                            float scale) {
    __shared__ float2 tile[32][33];
    const int64_t plane = (int64_t)rows_in * cols_in * blockIdx.z;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y)
        tile[j][threadIdx.x] = in[plane + (int64_t)(r0 + j) * cols_in + c0 + threadIdx.x];
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        float2 e = tile[threadIdx.x][j];
        out[plane + (int64_t)(c0 + j) * rows_in + r0 + threadIdx.x] = make_float2(e.x * scale, e.y * scale);
    }
}

// ------------------------------------------------------------------------------------------
// Table builders (float64 math, rounded once to float32).
// ------------------------------------------------------------------------------------------
// second table entry of a twiddle whose exact value is (c, s) and whose float value is hi: the lo part (fft_core.cuh)
__device__ __forceinline__ float2 second_entry(float2 hi, double c, double s) {
    return make_float2((float)(c - (double)hi.x), (float)(s - (double)hi.y));
}
// tw[ka*Q + q] = exp(-2 pi i q ka / (32 Q)): hi parts, hi parts transposed [q*32 + ka], lo parts, lo parts transposed
// (32*Q entries each; the lo parts are what float rounding dropped, see fft_core.cuh)
__global__ void k_tab_tw(float2* tw, int Q) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 32 * Q) return;
    const int ka = i / Q, q = i % Q;
    double s, c;
    sincospi(-2.0 * (double)(q * ka) / (double)(32 * Q), &s, &c);
    const float2 hi = make_float2((float)c, (float)s);
    const float2 lo = second_entry(hi, c, s);
    tw[i] = hi;
    tw[32 * Q + q * 32 + ka] = hi;
    tw[64 * Q + i] = lo;
    tw[96 * Q + q * 32 + ka] = lo;
}
// V[n2*32 + ka] = exp(-2 pi i n2 ka / N) ; U[n2*Q1 + kq] = exp(-2 pi i n2 32 kq / N); each followed by its lo parts
__global__ void k_tab_inter(float2* V, float2* U, int N2, int Q1, int64_t N) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < (int64_t)N2 * 32) {
        const int64_t n2 = i / 32, ka = i % 32;
        double s, c;
        sincospi(-2.0 * (double)((n2 * ka) % N) / (double)N, &s, &c);
        const float2 hi = make_float2((float)c, (float)s);
        V[i] = hi;
        V[(int64_t)N2 * 32 + i] = second_entry(hi, c, s);
    }
    if (i < (int64_t)N2 * Q1) {
        const int64_t n2 = i / Q1, kq = i % Q1;
        double s, c;
        sincospi(-2.0 * (double)((n2 * 32 * kq) % N) / (double)N, &s, &c);
        const float2 hi = make_float2((float)c, (float)s);
        U[i] = hi;
        U[(int64_t)N2 * Q1 + i] = second_entry(hi, c, s);
    }
}
__host__ __device__ inline int brev_rt(int v, int bits) {
    int r = 0;
    for (int b = 0; b < bits; ++b) r |= ((v >> b) & 1) << (bits - 1 - b);
    return r;
}
// Linear operator in k_freq's consumption order.  Entry i = ((tile*32 + slot2)*Q2 + t2)*C + c is the
// operator at frequency k = k1(p) + N1*k2 with p = tile*C + c,
//   k1(p) = ((p % Q1)*G1 + (p / Q1) / Q1) + 32*brev<Q1>((p / Q1) % Q1)      (W-row layout, see header)
//   k2    = (t2*G2 + slot2 / Q2) + 32*brev<Q2>(slot2 % Q2)
// value = scale * exp((a + j b ω_k²) h), ω_k = 2π Fs fftfreq(N)[k]   (channels.py:368, 406)
__global__ void k_tab_linop_perm(float2* __restrict__ LP, int Q1, int Q2, int C, int64_t N, double a, double b,
                                 double Fs, double h, double scale) {
    const double two_pi = 6.283185307179586476925286766559;
    const int N1 = 32 * Q1, G1 = 32 / Q1, G2 = 32 / Q2;
    const int lq1 = fft::ilog2(Q1), lq2 = fft::ilog2(Q2);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const int t2 = (int)((i / C) % Q2);
        const int slot2 = (int)((i / ((int64_t)C * Q2)) % 32);
        const int64_t tile = i / ((int64_t)C * Q2 * 32);
        const int p = (int)(tile * C + c);
        const int s1 = p / Q1, t1 = p % Q1;
        const int k1 = (t1 * G1 + s1 / Q1) + 32 * brev_rt(s1 % Q1, lq1);
        const int k2 = (t2 * G2 + slot2 / Q2) + 32 * brev_rt(slot2 % Q2, lq2);
        const int64_t k = k1 + (int64_t)N1 * k2;
        const int64_t kk = (k <= (N - 1) / 2) ? k : k - N;
        const double w = two_pi * Fs * ((double)kk / (double)N);
        const double amp = scale * exp(a * h);
        double sn, cs;
        sincos(b * (w * w) * h, &sn, &cs);
        LP[i] = make_float2((float)(amp * cs), (float)(amp * sn));
    }
}

}  // namespace ocb
