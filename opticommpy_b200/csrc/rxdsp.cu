// Receiver-DSP kernels behind the C-ABI: EDC (overlap-save), N x N adaptive MIMO equalizer,
// blind phase search.  Reference behaviour restated (not code):
//   edc / blockwiseFFTConv : optic/dsp/equalization.py:36-122, optic/dsp/core.py:973-1046
//   coreAdaptEq + *Up      : optic/dsp/equalization.py:354-516, 520-973
//   bps                    : optic/dsp/carrierRecovery.py:172-223
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "../../include/opticomm_b200.h"
#include "common.cuh"
#include "plan_cache.cuh"

using namespace ocb;

// =============================================================================================
// EDC: y = conv(x, h)[D : D+L], D = (K-1)//2, evaluated by overlap-save with the library's own
// block size (the result of an exact linear convolution does not depend on the block size).
// =============================================================================================
namespace {

struct EdcGeom {
    int nfft;        // block FFT size
    int d;           // hop = nfft - K + 1 valid outputs per block
    int64_t nblk;    // blocks per mode
};
EdcGeom edc_geom(int64_t L, int K) {
    EdcGeom g;
    int nfft = 4096;
    while (nfft < 8 * K) nfft *= 2;  // keep the overlap redundancy <= 1/8
    g.nfft = nfft;
    g.d = nfft - K + 1;
    g.nblk = (L + g.d - 1) / g.d;
    return g;
}

// seg[m][b][t] = x[m][b*d + D - (K-1) + t]  (zero outside [0, L))
__global__ void k_edc_gather(const float2* __restrict__ x, float2* __restrict__ seg, int64_t L, int nfft,
                             int d, int64_t nblk, int shift /* D-(K-1) */, int nModes) {
    const int64_t total = (int64_t)nModes * nblk * nfft;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int t = (int)(i % nfft);
        int64_t mb = i / nfft;
        int64_t b = mb % nblk;
        int64_t m = mb / nblk;
        int64_t src = b * d + shift + t;
        float2 v = make_float2(0.f, 0.f);
        if (src >= 0 && src < L) v = __ldg(x + m * L + src);
        seg[i] = v;
    }
}
// seg[m][b][k] *= Hf[k] / nfft
__global__ void k_edc_mul(float2* __restrict__ seg, const float2* __restrict__ Hf, int64_t total, int nfft,
                          float inv_n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        float2 h = __ldg(Hf + (i % nfft));
        float2 v = cmul(seg[i], h);
        seg[i] = make_float2(v.x * inv_n, v.y * inv_n);
    }
}
// y[m][b*d + j] = seg[m][b][K-1+j], j in [0, d)
__global__ void k_edc_scatter(const float2* __restrict__ seg, float2* __restrict__ y, int64_t L, int nfft,
                              int d, int64_t nblk, int K, int nModes) {
    const int64_t total = (int64_t)nModes * nblk * d;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int j = (int)(i % d);
        int64_t mb = i / d;
        int64_t b = mb % nblk;
        int64_t m = mb / nblk;
        int64_t dst = b * d + j;
        if (dst < L) y[m * L + dst] = seg[mb * nfft + (K - 1) + j];
    }
}
__global__ void k_edc_pad_taps(const float2* __restrict__ h, float2* __restrict__ hp, int K, int nfft) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nfft; i += gridDim.x * blockDim.x)
        hp[i] = i < K ? h[i] : make_float2(0.f, 0.f);
}

}  // namespace

extern "C" int64_t ocb_edc_workspace_bytes(int64_t L, int nModes, int K) {
    if (L <= 0 || nModes <= 0 || K <= 0) return -1;
    EdcGeom g = edc_geom(L, K);
    return (int64_t)nModes * g.nblk * g.nfft * 8 + (int64_t)g.nfft * 8 + 512;
}

extern "C" int ocb_edc_run(const void* x_rows, void* y_rows, int64_t L, int nModes, const void* h_taps, int K,
                           void* workspace, int64_t workspace_bytes, void* stream) {
    OCB_REQUIRE(x_rows && y_rows && h_taps && workspace, "edc_run: NULL argument");
    OCB_REQUIRE(L > 0 && nModes > 0 && K > 0, "edc_run: bad sizes");
    OCB_REQUIRE(workspace_bytes >= ocb_edc_workspace_bytes(L, nModes, K), "edc_run: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    EdcGeom g = edc_geom(L, K);
    const int D = (K - 1) / 2;  // core.py:1004
    float2* Hf = (float2*)workspace;
    float2* seg = (float2*)((char*)workspace + (((int64_t)g.nfft * 8 + 255) / 256) * 256);
    const int64_t nseg = (int64_t)nModes * g.nblk;
    OCB_REQUIRE(nseg < (1ll << 31), "edc_run: too many blocks");

    // plans come from the per-process cache (plan_cache.cuh): no allocation, no host synchronisation in this call
    cufftHandle ph, pb;
    OCB_CUFFT(fft_plan_cached(CUFFT_C2C, g.nfft, 1, st, &ph));
    OCB_CUFFT(fft_plan_cached(CUFFT_C2C, g.nfft, (int)nseg, st, &pb));
    OCB_LAUNCH(k_edc_pad_taps, grid_for(g.nfft, 256, 1), 256, 0, st, (const float2*)h_taps, Hf, K, g.nfft);
    OCB_CUFFT(cufftExecC2C(ph, Hf, Hf, CUFFT_FORWARD));  // core.py:1020
    OCB_LAUNCH(k_edc_gather, grid_for(nseg * g.nfft, 256, 4), 256, 0, st, (const float2*)x_rows, seg, L, g.nfft, g.d, g.nblk,
               D - (K - 1), nModes);
    OCB_CUFFT(cufftExecC2C(pb, seg, seg, CUFFT_FORWARD));
    OCB_LAUNCH(k_edc_mul, grid_for(nseg * g.nfft, 256, 4), 256, 0, st, seg, Hf, nseg * g.nfft, g.nfft, 1.0f / (float)g.nfft);
    OCB_CUFFT(cufftExecC2C(pb, seg, seg, CUFFT_INVERSE));
    OCB_LAUNCH(k_edc_scatter, grid_for(nseg * g.d, 256, 4), 256, 0, st, seg, (float2*)y_rows, L, g.nfft, g.d, g.nblk, K, nModes);
    return 0;
}

// =============================================================================================
// Adaptive MIMO equalizer.  The tap recurrences of the NM output modes of a stream are independent
// given the shared input window (every update touches only the rows H[m + n*NM, :] of its own output
// m), so the unit of work is a TASK = (stream, output mode), owned by LPS lanes (32 for a few streams:
// one warp per task, the NM tasks of a stream in one CTA; 8 when there are many streams).  Each lane
// keeps taps t = l + LPS*j (j < TPL) of the NM sub-filters of its output in registers.  The input of a
// chunk of 128 symbols is staged in shared memory by cp.async (double buffered) and shared by the
// tasks of the stream; per symbol: smem window read -> FMAs -> log2(LPS) shuffle stages -> error term
// -> tap update.  Tap layout: H[(m + n*NM), t] = tap t from input mode n to output mode m
// (equalization.py:467).
// =============================================================================================
namespace {

constexpr int kEqChunk = 128;  // symbols per staged chunk

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// NM consecutive float2 from a 32-bit shared-space address (one 128-bit access per mode pair)
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

// CTA = SPB stream slots x NM tasks x LPS lanes ; thread = (slot, m, l)
template <int NM, int TPL, int LPS, bool WL>
__global__ void __launch_bounds__(256)
k_mimo_eq(const float2* __restrict__ X, const float2* __restrict__ REF, float2* __restrict__ Hg,
          float2* __restrict__ HWg, float2* __restrict__ Y, float* __restrict__ ERR, float2* __restrict__ HIT,
          int nStreams, int SPB, int64_t xStride, int64_t refStride, int64_t yStride, int64_t errStride,
          int64_t errModeStride, int64_t L, int nTaps, int SpS, int alg, float mu,
          const float2* __restrict__ constSymb, int M, const float* __restrict__ radii, int nR, float Rcma) {
    extern __shared__ __align__(16) float2 smem_eq[];
    const int tid = threadIdx.x;
    const int l = tid % LPS;
    const int m = (tid / LPS) % NM;          // output mode of this task
    const int slot = tid / (LPS * NM);       // stream slot inside the CTA
    const int tslot = tid % (LPS * NM);      // thread index inside the stream slot (staging)
    const int stream_raw = blockIdx.x * SPB + slot;
    const bool live = stream_raw < nStreams;  // dead slots shadow the last stream, never store
    const int stream = live ? stream_raw : nStreams - 1;
    const int rows_chunk = (kEqChunk - 1) * SpS + nTaps;
    float2* xbuf = smem_eq + (size_t)slot * 2 * (rows_chunk * NM + kEqChunk * NM);
    float2* rbuf = xbuf + 2 * rows_chunk * NM;  // [2][kEqChunk*NM] reference symbols

    const float2* x = X + (int64_t)stream * xStride;
    const float2* ref = REF ? REF + (int64_t)stream * refStride : nullptr;
    float2* Hs = Hg + (int64_t)stream * NM * NM * nTaps;
    float2* HWs = WL ? HWg + (int64_t)stream * NM * NM * nTaps : nullptr;
    float2* y = Y + (int64_t)stream * yStride;
    float* err = ERR + (int64_t)stream * errStride + (int64_t)m * errModeStride;
    float2* hit = HIT ? HIT + (int64_t)stream * L * NM * NM * nTaps : nullptr;

    float2 H[NM][TPL], HW[WL ? NM : 1][TPL];  // rows m + n*NM, n < NM
#pragma unroll
    for (int n = 0; n < NM; ++n)
#pragma unroll
        for (int j = 0; j < TPL; ++j) {
            const int t = l + LPS * j;
            H[n][j] = t < nTaps ? Hs[(m + n * NM) * nTaps + t] : make_float2(0.f, 0.f);
            if (WL) HW[n][j] = t < nTaps ? HWs[(m + n * NM) * nTaps + t] : make_float2(0.f, 0.f);
        }

    // stage chunk k (symbols [k*CH, min(L, (k+1)*CH))) into buffer k&1; all NM*LPS threads of the slot copy
    auto stage = [&](int64_t k) {
        const int64_t s0 = k * kEqChunk;
        if (s0 >= L) return;
        const int nsym = (int)((L - s0) < kEqChunk ? (L - s0) : kEqChunk);
        const int rows = (nsym - 1) * SpS + nTaps;
        float2* dst = xbuf + (k & 1) * rows_chunk * NM;
        const float2* src = x + s0 * SpS * NM;
        for (int i = tslot; i < rows * NM; i += LPS * NM) cp_async8(dst + i, src + i);
        if (ref) {
            float2* rd = rbuf + (k & 1) * kEqChunk * NM;
            const float2* rs = ref + s0 * NM;
            for (int i = tslot; i < nsym * NM; i += LPS * NM) cp_async8(rd + i, rs + i);
        }
    };

    float prev_err = 0.f;
    // RDE ring decision without a square root on the per-symbol critical path: the radii are ascending
    // (np.unique, equalization.py:456), so argmin_i |R_i - |o|| is the last i with |o|² > ((R_{i-1} + R_i)/2)²
    // (a tie keeps the lower ring, like argmin).  Squared radii / thresholds of the first rings in registers
    // (16/64-QAM have 3/9 rings); more rings fall back to memory.
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
    // window addressing: lane tap t = l + LPS*j reads row (s*SpS + t) of the staged chunk; taps beyond nTaps
    // read a clamped (valid) row and are zeroed by a loop-invariant select
    bool tapv[TPL];
    unsigned tapoff[TPL];
#pragma unroll
    for (int j = 0; j < TPL; ++j) {
        const int t = l + LPS * j;
        tapv[j] = t < nTaps;
        tapoff[j] = (unsigned)((tapv[j] ? t : nTaps - 1) * NM) * 8u;
    }
    const unsigned xbuf_s = (unsigned)__cvta_generic_to_shared(xbuf);
    const unsigned sym_stride = (unsigned)(SpS * NM) * 8u;

    // The symbol loop is instantiated once per algorithm (compile-time ALG) so that no algorithm dispatch,
    // constant-bank reload or dead select chain sits on the per-symbol critical path of a lone warp.
    auto run = [&](auto algc) {
    constexpr int ALG = decltype(algc)::value;
    stage(0);
    cp_async_commit();
    const int64_t nchunks = (L + kEqChunk - 1) / kEqChunk;
    for (int64_t k = 0; k < nchunks; ++k) {
        cp_async_wait_all();
        __syncthreads();  // chunk k landed for every task of the CTA; chunk k-1 fully consumed
        stage(k + 1);     // lands while this chunk is processed
        cp_async_commit();
        const float2* rb = rbuf + (k & 1) * kEqChunk * NM;
        const int64_t s0 = k * kEqChunk;
        const int nsym = (int)((L - s0) < kEqChunk ? (L - s0) : kEqChunk);

        // software pipeline: the window of symbol s+1 is read while symbol s goes through its reduction
        unsigned waddr = xbuf_s + (unsigned)((k & 1) * rows_chunk * NM) * 8u;
        float2 wn[TPL][NM];
#pragma unroll
        for (int j = 0; j < TPL; ++j) lds_window<NM>(waddr + tapoff[j], wn[j]);
        float2* yp = y + (s0 * NM + m);
        float* ep = err + s0;
        const bool wr = live && l == 0;
        for (int s = 0; s < nsym; ++s) {
            const int64_t ind = s0 + s;
            float2 w[NM][TPL];
#pragma unroll
            for (int j = 0; j < TPL; ++j)
#pragma unroll
                for (int n = 0; n < NM; ++n) w[n][j] = tapv[j] ? wn[j][n] : make_float2(0.f, 0.f);
            waddr += (s + 1 < nsym) ? sym_stride : 0u;
#pragma unroll
            for (int j = 0; j < TPL; ++j) lds_window<NM>(waddr + tapoff[j], wn[j]);
            // ---- filter: out_m = sum_n H[m + n NM, :] . x_n[window]   (equalization.py:464-471)
            float2 o = make_float2(0.f, 0.f);
            float nrm[NM];
#pragma unroll
            for (int n = 0; n < NM; ++n)
#pragma unroll
                for (int j = 0; j < TPL; ++j) {
                    float2 pr = cmul(H[n][j], w[n][j]);
                    o.x += pr.x; o.y += pr.y;
                    if (WL) {
                        float2 qq = cmul_conj(HW[n][j], w[n][j]);  // H_ . conj(x)
                        o.x += qq.x; o.y += qq.y;
                    }
                }
            if constexpr (ALG == OCB_ALG_NLMS) {
#pragma unroll
                for (int n = 0; n < NM; ++n) {
                    float sacc = 0.f;
#pragma unroll
                    for (int j = 0; j < TPL; ++j) sacc += cabs2(w[n][j]);
                    nrm[n] = sacc;
                }
            }
#pragma unroll
            for (int off = LPS / 2; off > 0; off >>= 1) {
                o.x += __shfl_xor_sync(0xffffffffu, o.x, off);
                o.y += __shfl_xor_sync(0xffffffffu, o.y, off);
                if constexpr (ALG == OCB_ALG_NLMS) {
#pragma unroll
                    for (int n = 0; n < NM; ++n) nrm[n] += __shfl_xor_sync(0xffffffffu, nrm[n], off);
                }
            }
            if (wr) *yp = o;  // equalization.py:473
            yp += NM;

            // ---- error term g and squared error, per algorithm
            float2 g;
            float esq;
            const float a2 = cabs2(o);
            if constexpr (ALG == OCB_ALG_CMA) {  // :826-829
                float e = Rcma - a2;
                g = make_float2(e * o.x, e * o.y);
                esq = e * e;
            } else if constexpr (ALG == OCB_ALG_RDE || ALG == OCB_ALG_DARDE) {  // :887-894, :953-959
                float Rd;
                float Rd2;
                if constexpr (ALG == OCB_ALG_RDE) {
                    Rd2 = rad2[0];
                    if (nR <= 4) {  // uniform branch: QPSK / 16-QAM / 16-APSK need three selects, not nine
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
                } else {
                    Rd = sqrtf(cabs2(rb[s * NM + m]));
                    Rd2 = Rd * Rd;
                }
                float e = Rd2 - a2;
                g = make_float2(e * o.x, e * o.y);
                esq = e * e;
            } else if constexpr (ALG == OCB_ALG_NLMS) {  // :556
                float2 sr = rb[s * NM + m];
                g = make_float2(sr.x - o.x, sr.y - o.y);
                esq = cabs2(g);
            } else if constexpr (ALG == OCB_ALG_DDLMS) {  // :688-691  nearest constellation point, first index on ties
                float best = 3.4e38f;
                int bi = 0x7fffffff;
                for (int c = l; c < M; c += LPS) {
                    float2 sc = __ldg(constSymb + c);
                    float dd = cabs2(make_float2(o.x - sc.x, o.y - sc.y));
                    if (dd < best) { best = dd; bi = c; }
                }
#pragma unroll
                for (int off = LPS / 2; off > 0; off >>= 1) {
                    float ob = __shfl_xor_sync(0xffffffffu, best, off);
                    int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                }
                float2 sc = __ldg(constSymb + bi);
                g = make_float2(sc.x - o.x, sc.y - o.y);
                esq = cabs2(g);
            } else {  // static: no update (:505-506)
                g = make_float2(0.f, 0.f);
                esq = prev_err;
            }
            prev_err = esq;
            if (wr) *ep = esq;
            ++ep;

            // ---- tap update: H[m + n NM, :] += mu g conj(x_n)   (:838-840 and siblings)
            if constexpr (ALG != OCB_ALG_STATIC) {
                const float2 wg = make_float2(mu * g.x, mu * g.y);
#pragma unroll
                for (int n = 0; n < NM; ++n) {
                    const float inv = (ALG == OCB_ALG_NLMS) ? 1.0f / nrm[n] : 1.0f;
#pragma unroll
                    for (int j = 0; j < TPL; ++j) {
                        float2 xin = w[n][j];
                        if constexpr (ALG == OCB_ALG_NLMS) { xin.x *= inv; xin.y *= inv; }  // :563
                        float2 u = cmul_conj(wg, xin);
                        H[n][j].x += u.x; H[n][j].y += u.y;
                        if (WL) {
                            float2 vv = cmul(wg, xin);
                            HW[n][j].x += vv.x; HW[n][j].y += vv.y;
                        }
                    }
                }
            }
            if (hit && live) {  // storeCoeff (:511-512): Hiter[:, :, ind] = H, stored as (L, NM^2, nTaps)
#pragma unroll
                for (int n = 0; n < NM; ++n)
#pragma unroll
                    for (int j = 0; j < TPL; ++j) {
                        const int t = l + LPS * j;
                        if (t < nTaps) hit[(ind * NM * NM + m + n * NM) * nTaps + t] = H[n][j];
                    }
            }
        }
    }
    };
    switch (alg) {
        case OCB_ALG_CMA: run(std::integral_constant<int, OCB_ALG_CMA>{}); break;
        case OCB_ALG_RDE: run(std::integral_constant<int, OCB_ALG_RDE>{}); break;
        case OCB_ALG_NLMS: run(std::integral_constant<int, OCB_ALG_NLMS>{}); break;
        case OCB_ALG_DDLMS: run(std::integral_constant<int, OCB_ALG_DDLMS>{}); break;
        case OCB_ALG_DARDE: run(std::integral_constant<int, OCB_ALG_DARDE>{}); break;
        default: run(std::integral_constant<int, OCB_ALG_STATIC>{}); break;
    }
    cp_async_wait_all();

    if (live) {
#pragma unroll
        for (int n = 0; n < NM; ++n)
#pragma unroll
            for (int j = 0; j < TPL; ++j) {
                const int t = l + LPS * j;
                if (t < nTaps) {
                    Hs[(m + n * NM) * nTaps + t] = H[n][j];
                    if (WL) HWs[(m + n * NM) * nTaps + t] = HW[n][j];
                }
            }
    }
}

template <int NM, int TPL, int LPS>
int launch_mimo(bool wl, cudaStream_t st, const float2* X, const float2* REF, float2* H,
                float2* HW, float2* Y, float* ERR, float2* HIT, int nStreams, int64_t xs, int64_t rs, int64_t ys,
                int64_t es, int64_t ems, int64_t L, int nTaps,
                int SpS, int alg, float mu, const float2* cs, int M, const float* radii, int nR, float Rcma) {
    // stream slots per CTA: one in latency mode (CTA = NM warps), up to 128 threads' worth otherwise
    int spb = (LPS == 32) ? 1 : (128 / (LPS * NM) > 0 ? 128 / (LPS * NM) : 1);
    if (spb > nStreams) spb = nStreams;
    if (spb * LPS * NM < 32) spb = 32 / (LPS * NM);  // never launch a partial warp (dead slots shadow the last stream)
    const int grid = (nStreams + spb - 1) / spb;
    const int block = spb * NM * LPS;
    const int rows_chunk = (kEqChunk - 1) * SpS + nTaps;
    const size_t smem = (size_t)spb * 2 * ((size_t)rows_chunk * NM + kEqChunk * NM) * sizeof(float2);
    OCB_REQUIRE(smem <= 200 * 1024, "mimo_eq_run: SpS/nTaps too large for the staged input chunk");
    if (wl) {
        OCB_CUDA(cudaFuncSetAttribute(k_mimo_eq<NM, TPL, LPS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        OCB_LAUNCH((k_mimo_eq<NM, TPL, LPS, true>), grid, block, smem, st, X, REF, H, HW, Y, ERR, HIT, nStreams, spb, xs, rs, ys, es, ems, L, nTaps, SpS, alg, mu, cs, M, radii, nR, Rcma);
    } else {
        OCB_CUDA(cudaFuncSetAttribute(k_mimo_eq<NM, TPL, LPS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        OCB_LAUNCH((k_mimo_eq<NM, TPL, LPS, false>), grid, block, smem, st, X, REF, H, HW, Y, ERR, HIT, nStreams, spb, xs, rs, ys, es, ems, L, nTaps, SpS, alg, mu, cs, M, radii, nR, Rcma);
    }
    return 0;
}

#include "rxdsp_eq_la.cuh"
#include "rxdsp_rls.cuh"

}  // namespace

extern "C" int64_t ocb_mimo_eq_rls_workspace_bytes(int nStreams, int nModes, int64_t L, int nTaps) {
    if (nStreams <= 0 || nModes <= 0 || L < 0 || nTaps < 1 || nTaps > 64) return -1;
    return (int64_t)nStreams * nModes * L * (nTaps <= 32 ? 32 : 64) * (int64_t)sizeof(float2) + 256;
}

extern "C" int ocb_mimo_eq_rls_run(const void* x, const void* ref, void* H, void* y, void* errSq, void* Hiter,
                                   int nStreams, int64_t nSamp, int64_t x_stream_stride, int64_t ref_stream_stride,
                                   int64_t y_stream_stride, int64_t err_stream_stride, int64_t err_mode_stride,
                                   int64_t L, int nModes, int nTaps, int SpS, int decision_directed, float lambda,
                                   const void* constSymb, int M, void* workspace, int64_t workspace_bytes,
                                   void* stream) {
    OCB_REQUIRE(x && H && y && errSq && workspace, "mimo_eq_rls_run: NULL argument");
    OCB_REQUIRE(nStreams > 0 && L >= 0 && nSamp > 0, "mimo_eq_rls_run: bad sizes");
    OCB_REQUIRE(nModes == 1 || nModes == 2 || nModes == 4, "mimo_eq_rls_run: nModes must be 1, 2 or 4");
    OCB_REQUIRE(nTaps >= 1 && nTaps <= 64, "mimo_eq_rls_run: nTaps must be in [1, 64] (one or two matrix rows per lane)");
    OCB_REQUIRE(SpS >= 1 && lambda > 0.f, "mimo_eq_rls_run: SpS must be >= 1 and lambda > 0");
    OCB_REQUIRE(L == 0 || (L - 1) * SpS + nTaps <= nSamp, "mimo_eq_rls_run: window runs past the end of the input");
    if (decision_directed) OCB_REQUIRE(constSymb && M >= 1, "mimo_eq_rls_run: constellation required for dd-rls");
    else OCB_REQUIRE(ref != nullptr, "mimo_eq_rls_run: reference symbols required for rls");
    OCB_REQUIRE(workspace_bytes >= ocb_mimo_eq_rls_workspace_bytes(nStreams, nModes, L, nTaps), "mimo_eq_rls_run: workspace too small");
    if (L == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
#define OCB_RLS_CASE(NM_)                                                                                          \
    if (nModes == NM_)                                                                                             \
        return launch_rls<NM_>(st, (const float2*)x, (const float2*)ref, (float2*)workspace, (float2*)H, (float2*)y, \
                               (float*)errSq, (float2*)Hiter, nStreams, x_stream_stride, ref_stream_stride,        \
                               y_stream_stride, err_stream_stride, err_mode_stride, L, nTaps, SpS,                 \
                               decision_directed != 0, lambda, (const float2*)constSymb, M);
    OCB_RLS_CASE(1) OCB_RLS_CASE(2) OCB_RLS_CASE(4)
#undef OCB_RLS_CASE
    return fail("mimo_eq_rls_run: unsupported nModes", __FILE__, __LINE__);
}

extern "C" int ocb_mimo_eq_run(const void* x, const void* ref, void* H, void* Hwl, void* y, void* errSq,
                               void* Hiter, int nStreams, int64_t nSamp, int64_t x_stream_stride,
                               int64_t ref_stream_stride, int64_t y_stream_stride, int64_t err_stream_stride,
                               int64_t err_mode_stride, int64_t L, int nModes, int nTaps, int SpS,
                               int alg, float mu, const void* constSymb, int M, const void* radii, int nR,
                               float Rcma, int runWL, void* stream) {
    OCB_REQUIRE(x && H && y && errSq, "mimo_eq_run: NULL argument");
    OCB_REQUIRE(nStreams > 0 && L >= 0 && nSamp > 0, "mimo_eq_run: bad sizes");
    OCB_REQUIRE(nModes == 1 || nModes == 2 || nModes == 4, "mimo_eq_run: nModes must be 1, 2 or 4");
    OCB_REQUIRE(nTaps >= 1 && nTaps <= 128, "mimo_eq_run: nTaps must be in [1, 128]");
    OCB_REQUIRE(SpS >= 1, "mimo_eq_run: SpS must be >= 1");
    OCB_REQUIRE(alg >= OCB_ALG_CMA && alg <= OCB_ALG_STATIC,
                "Equalization algorithm not specified (or incorrectly specified).");  // equalization.py:507-510
    OCB_REQUIRE(L == 0 || (L - 1) * SpS + nTaps <= nSamp, "mimo_eq_run: window runs past the end of the input");
    if (alg == OCB_ALG_NLMS || alg == OCB_ALG_DARDE) OCB_REQUIRE(ref != nullptr, "mimo_eq_run: reference symbols required");
    if (alg == OCB_ALG_RDE) OCB_REQUIRE(radii && nR >= 1, "mimo_eq_run: radii required for RDE");
    if (alg == OCB_ALG_DDLMS) OCB_REQUIRE(constSymb && M >= 1, "mimo_eq_run: constellation required for DD-LMS");
    if (runWL) OCB_REQUIRE(Hwl != nullptr, "mimo_eq_run: runWL needs the augmented taps H_");
    if (L == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    // Lanes per stream.  Few streams (latency mode): a full warp per stream, one tap per lane when
    // nTaps <= 32 — the per-symbol instruction count is what bounds a lone stream.  Many streams
    // (throughput mode): 8 lanes x 4 taps, four streams per warp, 3 shuffle stages instead of 5.
    const bool many = nStreams >= 4 * kNumSMs;
    int lps = (many && nTaps <= 32) ? 8 : ((many && nTaps <= 64) ? 16 : 32);
    if (const char* e = getenv("OCB_EQ_LPS")) {  // tuning override (8, 16 or 32 lanes per task)
        const int v = atoi(e);
        if ((v == 8 || v == 16 || v == 32) && (nTaps + v - 1) / v <= 4 && (v == 32 || (nTaps + v - 1) / v >= (v == 16 ? 2 : 1))) lps = v;
    }
    const int tpl = (nTaps + lps - 1) / lps;
    // latency mode (a warp per task): look-ahead kernel, rxdsp_eq_la.cuh (OCB_EQ_LA=0: plain recurrence, for A/B runs)
    static const bool use_la = !(getenv("OCB_EQ_LA") && atoi(getenv("OCB_EQ_LA")) == 0);
#define OCB_MIMO_LA_CASE(NM_, TPL_)                                                                            \
    if (use_la && lps == 32 && nModes == NM_ && tpl == TPL_)                                                   \
        return launch_mimo_la<NM_, TPL_>(runWL != 0, st, (const float2*)x, (const float2*)ref, (float2*)H,     \
                                         (float2*)Hwl, (float2*)y, (float*)errSq, (float2*)Hiter, nStreams,    \
                                         x_stream_stride, ref_stream_stride, y_stream_stride,                  \
                                         err_stream_stride, err_mode_stride, L, nTaps, SpS, alg, mu,           \
                                         (const float2*)constSymb, M, (const float*)radii, nR, Rcma);
    OCB_MIMO_LA_CASE(1, 1) OCB_MIMO_LA_CASE(1, 2) OCB_MIMO_LA_CASE(1, 3) OCB_MIMO_LA_CASE(1, 4)
    OCB_MIMO_LA_CASE(2, 1) OCB_MIMO_LA_CASE(2, 2) OCB_MIMO_LA_CASE(2, 3) OCB_MIMO_LA_CASE(2, 4)
    OCB_MIMO_LA_CASE(4, 1) OCB_MIMO_LA_CASE(4, 2)
#undef OCB_MIMO_LA_CASE
#define OCB_MIMO_CASE(NM_, TPL_, LPS_)                                                                         \
    if (nModes == NM_ && lps == LPS_ && tpl == TPL_)                                                           \
        return launch_mimo<NM_, TPL_, LPS_>(runWL != 0, st, (const float2*)x, (const float2*)ref,              \
                                         (float2*)H, (float2*)Hwl, (float2*)y, (float*)errSq, (float2*)Hiter,  \
                                         nStreams, x_stream_stride, ref_stream_stride, y_stream_stride,        \
                                         err_stream_stride, err_mode_stride, L, nTaps, SpS, alg, mu,           \
                                         (const float2*)constSymb, M, (const float*)radii, nR, Rcma);
    OCB_MIMO_CASE(1, 1, 32) OCB_MIMO_CASE(1, 2, 32) OCB_MIMO_CASE(1, 3, 32) OCB_MIMO_CASE(1, 4, 32)
    OCB_MIMO_CASE(2, 1, 32) OCB_MIMO_CASE(2, 2, 32) OCB_MIMO_CASE(2, 3, 32) OCB_MIMO_CASE(2, 4, 32)
    OCB_MIMO_CASE(4, 1, 32) OCB_MIMO_CASE(4, 2, 32)
    OCB_MIMO_CASE(1, 1, 8) OCB_MIMO_CASE(1, 2, 8) OCB_MIMO_CASE(1, 3, 8) OCB_MIMO_CASE(1, 4, 8)
    OCB_MIMO_CASE(2, 1, 8) OCB_MIMO_CASE(2, 2, 8) OCB_MIMO_CASE(2, 3, 8) OCB_MIMO_CASE(2, 4, 8)
    OCB_MIMO_CASE(1, 2, 16) OCB_MIMO_CASE(2, 2, 16)
    OCB_MIMO_CASE(1, 3, 16) OCB_MIMO_CASE(1, 4, 16) OCB_MIMO_CASE(2, 3, 16) OCB_MIMO_CASE(2, 4, 16)
#undef OCB_MIMO_CASE
    return fail("mimo_eq_run: unsupported (nModes, nTaps) combination", __FILE__, __LINE__);
}

// =============================================================================================
// Blind phase search (float64, like the reference): per (mode, symbol, test phase) the minimum
// squared distance to the constellation, then a centred (2N+1)-window sum and an argmin over
// the B test phases (first index on ties).  One CTA = one tile of symbols of one mode; the
// dmin tile (B x (TS+2N)) lives in shared memory.
// =============================================================================================
namespace {

__global__ void __launch_bounds__(512)
k_bps(const double2* __restrict__ X, int64_t L, const double2* __restrict__ cs, int M, int B, int Nh, int TS,
      int32_t* __restrict__ idx_out, double* __restrict__ ph_out) {
    extern __shared__ double sm[];
    const int W = TS + 2 * Nh;            // tile width incl. halo
    double* dmin = sm;                    // [B][W]
    double2* rot = (double2*)(dmin + (size_t)B * W);  // [B]
    double2* csm = rot + B;               // [M]
    const int mode = blockIdx.y;
    const int64_t k0 = (int64_t)blockIdx.x * TS;
    const int nModes = gridDim.y;  // samples are interleaved (L, nModes) like the reference arrays

    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        double phi = ((double)b * (M_PI / 2.0)) / (double)B;  // carrierRecovery.py:199
        double s, c;
        sincos(phi, &s, &c);
        rot[b] = make_double2(c, s);
    }
    for (int c = threadIdx.x; c < M; c += blockDim.x) csm[c] = cs[c];
    __syncthreads();

    // phase A: dmin[b][j] for symbol k0 - Nh + j  (zero-padded outside [0, L): :203-206).  The M distances of an item are
    // independent (the minimum is exact in any order): four running minima break the dependent fmin chain.
    for (int i = threadIdx.x; i < B * W; i += blockDim.x) {
        const int j = i % W, b = i / W;
        const int64_t k = k0 - Nh + j;
        double2 v = make_double2(0.0, 0.0);
        if (k >= 0 && k < L) v = X[k * nModes + mode];
        const double2 r = rot[b];
        const double zr = v.x * r.x - v.y * r.y, zi = v.x * r.y + v.y * r.x;
        double m0 = 1.0e300, m1 = 1.0e300, m2 = 1.0e300, m3 = 1.0e300;
        int c = 0;
        for (; c + 4 <= M; c += 4) {
            const double2 c0 = csm[c], c1 = csm[c + 1], c2 = csm[c + 2], c3 = csm[c + 3];
            const double a0 = zr - c0.x, b0 = zi - c0.y, a1 = zr - c1.x, b1 = zi - c1.y;
            const double a2 = zr - c2.x, b2 = zi - c2.y, a3 = zr - c3.x, b3 = zi - c3.y;
            m0 = fmin(m0, a0 * a0 + b0 * b0);  // :216-217 (same expression per point: dr*dr + di*di)
            m1 = fmin(m1, a1 * a1 + b1 * b1);
            m2 = fmin(m2, a2 * a2 + b2 * b2);
            m3 = fmin(m3, a3 * a3 + b3 * b3);
        }
        for (; c < M; ++c) {
            const double dr = zr - csm[c].x, di = zi - csm[c].y;
            m0 = fmin(m0, dr * dr + di * di);
        }
        dmin[(size_t)b * W + j] = fmin(fmin(m0, m1), fmin(m2, m3));
    }
    __syncthreads();

    // phase B: window sums + argmin over b  (:218-221).  Thread (h, j) handles the test phases b = h, h + NH, ... of
    // symbol j; every window sum runs over t = 0 .. 2 Nh in that order (the summation order is part of the bit-exact
    // contract), four sums at a time so that their dependent additions overlap.
    const int NH = blockDim.x / TS;       // threads per symbol (host guarantees blockDim.x = NH * TS)
    const int j = threadIdx.x % TS, h = threadIdx.x / TS;
    const int per = (B + NH - 1) / NH;    // contiguous b range of this thread: [h*per, min(B, (h+1)*per))
    double best = 1.0e300;
    int bi = 0x7fffffff;
    {
        const int b_lo = h * per, b_hi = min(B, (h + 1) * per);
        int b = b_lo;
        for (; b + 4 <= b_hi; b += 4) {
            const double* r0 = dmin + (size_t)b * W + j;
            const double *r1 = r0 + W, *r2 = r1 + W, *r3 = r2 + W;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            for (int t = 0; t <= 2 * Nh; ++t) { s0 += r0[t]; s1 += r1[t]; s2 += r2[t]; s3 += r3[t]; }
            if (s0 < best) { best = s0; bi = b; }
            if (s1 < best) { best = s1; bi = b + 1; }
            if (s2 < best) { best = s2; bi = b + 2; }
            if (s3 < best) { best = s3; bi = b + 3; }
        }
        for (; b < b_hi; ++b) {
            const double* row = dmin + (size_t)b * W + j;
            double s = 0.0;
            for (int t = 0; t <= 2 * Nh; ++t) s += row[t];
            if (s < best) { best = s; bi = b; }
        }
    }
    if (NH > 1) {
        __syncthreads();  // dmin has been read by everyone: reuse its head as the hand-over buffer
        double* hb = sm;                            // [NH][TS] best sums
        int* hi = (int*)(sm + (size_t)NH * TS);     // [NH][TS] their indices
        hb[h * TS + j] = best;
        hi[h * TS + j] = bi;
        __syncthreads();
        if (h == 0) {
            for (int q = 1; q < NH; ++q) {          // ascending b ranges: a later range wins only with a strictly smaller sum
                const double ob = hb[q * TS + j];
                if (ob < best) { best = ob; bi = hi[q * TS + j]; }
            }
        }
    }
    if (h == 0) {
        const int64_t k = k0 + j;
        if (k < L) {
            idx_out[k * nModes + mode] = bi;
            ph_out[k * nModes + mode] = ((double)bi * (M_PI / 2.0)) / (double)B;
        }
    }
}

}  // namespace

extern "C" int ocb_bps_run(const void* x, int64_t L, int nModes, const void* constSymb, int M, int B, int Nhalf,
                           void* idx_out, void* phase_out, void* stream) {
    OCB_REQUIRE(x && constSymb && idx_out && phase_out, "bps_run: NULL argument");
    OCB_REQUIRE(L > 0 && nModes > 0 && M > 0 && B > 0 && Nhalf >= 0, "bps_run: bad sizes");
    OCB_REQUIRE(nModes <= 65535, "bps_run: too many modes");
    cudaStream_t st = (cudaStream_t)stream;
    // Tile of TS symbols per CTA (+ 2*Nhalf halo columns); 512 threads = NH threads per symbol in the window-sum phase.
    // TS = 128: 64 x 152 x 8 B = 78 KB at B = 64, N = 25 -> two CTAs (32 warps) per SM, 19 % halo recomputation.
    int TS = 128;
    auto smem_for = [&](int ts) { return (size_t)B * (ts + 2 * Nhalf) * 8 + (size_t)(B + M) * 16; };
    while (TS > 16 && smem_for(TS) > 100 * 1024) TS /= 2;
    const int threads = 512, NH = threads / TS;
    size_t smem = smem_for(TS);
    if (smem < (size_t)NH * TS * 12 + 16) smem = (size_t)NH * TS * 12 + 16;  // hand-over buffer of the window-sum phase
    OCB_REQUIRE(smem <= 227 * 1024, "bps_run: B*(window) does not fit in shared memory");
    OCB_CUDA(cudaFuncSetAttribute(k_bps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((L + TS - 1) / TS), (unsigned)nModes);
    OCB_LAUNCH(k_bps, grid, threads, smem, st, (const double2*)x, L, (const double2*)constSymb, M, B, Nhalf, TS,
               (int32_t*)idx_out, (double*)phase_out);
    return 0;
}
