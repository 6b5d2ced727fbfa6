// Rx front-end glue on the device: the polarisation-multiplexed coherent front end with
// ideal photodiodes, the channel down-shift, and the correlation machinery of symbolSync.  Reference behaviour restated
// (not code):
//   pdmCoherentReceiver / coherentReceiver / opticalHybrid2x4 / balancedPD / pbs : optic/models/devices.py:574-668,
//                                                 506-571, 447-503, 402-444, 223-262
//   iqMixing (amplitude / phase imbalance)      : optic/dsp/core.py:925-972
//   symbolSync / finddelay                      : optic/dsp/core.py:552-675, 678-698
#include <math.h>

#include "../../include/opticomm_b200.h"
#include "common.cuh"
#include "plan_cache.cuh"

using namespace ocb;

namespace {

__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }

// One pass over the samples: polarisation rotation of the signal (pbs, devices.py:254-259), PDL (:655-657), LO split
// at 45 degrees (:647), 2x4 90-degree hybrid (devices.py:489-503) + balanced ideal photodiodes (R |E|^2, :313-318,
// 437-444), IQ amplitude / phase imbalance (core.py:951-959).  Es: planar rows [2][N]; lo: [N] (the LO field) or
// nullptr for a noiseless CW LO generated on the fly: sqrt(Plo) exp(j 2 pi f n / Fs) — the channel down-shift.
__global__ void k_pdm_frontend(const float2* __restrict__ Es, const float2* __restrict__ lo, float2* __restrict__ S,
                               int64_t N, float cr, float sr, float gx, float gy, float R, double lo_amp,
                               double lo_cycles_per_sample, double2 k1x, double2 k2x, double2 k1y, double2 k2y) {
    const float c45 = 0.70710678118654752f;  // cos(pi/4) = sin(pi/4) rounded like numpy's float64 -> complex64 path
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        const float2 ex = Es[n], ey = Es[N + n];
        float2 l;
        if (lo) {
            l = lo[n];
        } else {
            double s, c;
            double ph = lo_cycles_per_sample * (double)n;
            ph -= floor(ph);
            sincospi(2.0 * ph, &s, &c);
            l = make_float2((float)(lo_amp * c), (float)(lo_amp * s));
        }
        // E @ [[c, -s], [s, c]]
        float2 esx = make_float2(ex.x * cr + ey.x * sr, ex.y * cr + ey.y * sr);
        float2 esy = make_float2(-ex.x * sr + ey.x * cr, -ex.y * sr + ey.y * cr);
        esx = cscale(esx, gx);
        esy = cscale(esy, gy);
        const float2 lx = cscale(l, c45), ly = cscale(l, -c45);
        float2 out[2];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const float2 es = p ? esy : esx, el = p ? ly : lx;
            // Eo = T [Es, 0, 0, Elo]^T
            const float2 e0 = make_float2(0.5f * es.x - 0.5f * el.x, 0.5f * es.y - 0.5f * el.y);
            const float2 e1 = make_float2(-0.5f * es.y - 0.5f * el.y, 0.5f * es.x + 0.5f * el.x);
            const float2 e2 = make_float2(-0.5f * es.y - 0.5f * el.x, 0.5f * es.x - 0.5f * el.y);
            const float2 e3 = make_float2(-0.5f * es.x - 0.5f * el.y, -0.5f * es.y + 0.5f * el.x);
            const float sI = R * cabs2(e1) - R * cabs2(e0);
            const float sQ = R * cabs2(e2) - R * cabs2(e3);
            // sig_ = k1 s + k2 conj(s)   (core.py:959)
            const double2 k1 = p ? k1y : k1x, k2 = p ? k2y : k2x;
            const double re = k1.x * sI - k1.y * sQ + k2.x * sI + k2.y * sQ;
            const double im = k1.x * sQ + k1.y * sI - k2.x * sQ + k2.y * sI;
            out[p] = make_float2((float)re, (float)im);
        }
        S[n] = out[0];
        S[N + n] = out[1];
    }
}

// rows[r][n] *= exp(-j 2 pi f n / Fs)  — stand-alone channel down-shift (phase in double, reduced mod 1 turn)
__global__ void k_freq_shift(float2* __restrict__ rows, int nRows, int64_t N, double cycles_per_sample) {
    const int64_t total = (int64_t)nRows * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = i % N;
        double ph = cycles_per_sample * (double)n;
        ph -= floor(ph);
        double s, c;
        sincospi(-2.0 * ph, &s, &c);
        const float2 v = rows[i];
        rows[i] = make_float2((float)(v.x * c - v.y * s), (float)(v.x * s + v.y * c));
    }
}

// ---- symbolSync: real sequences and cross-correlation peaks -----------------------------------------------------------
// out[r][n] = kind 0: |z| - mean|z| ; 1: Re z ; 2: Im z      (z = column r of an (L, nCols) interleaved complex128 array)
__global__ void k_sync_sequence(const double2* __restrict__ z, int nCols, int64_t L, int kind, double* __restrict__ out) {
    __shared__ double sh[32];
    __shared__ double mean_s;
    const int r = blockIdx.x;
    double acc = 0.0;
    for (int64_t n = threadIdx.x; n < L; n += blockDim.x) {
        const double2 v = z[n * nCols + r];
        const double val = kind == 0 ? hypot(v.x, v.y) : (kind == 1 ? v.x : v.y);
        out[(int64_t)r * L + n] = val;
        acc += val;
    }
    if (kind != 0) return;
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
        mean_s = t / (double)L;
    }
    __syncthreads();
    const double m = mean_s;
    for (int64_t n = threadIdx.x; n < L; n += blockDim.x) out[(int64_t)r * L + n] -= m;
}
__global__ void k_real_to_padded(const double* __restrict__ src, int nRows, int64_t L, int64_t nfft, double2* __restrict__ dst) {
    const int64_t total = (int64_t)nRows * nfft;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = i % nfft, r = i / nfft;
        dst[i] = make_double2(n < L ? src[r * L + n] : 0.0, 0.0);
    }
}
// C[i*nB + j][k] = FA[i][k] conj(FB[j][k])
__global__ void k_xcorr_mul(const double2* __restrict__ FA, const double2* __restrict__ FB, int nA, int nB, int64_t nfft,
                            double2* __restrict__ Cx) {
    const int64_t total = (int64_t)nA * nB * nfft;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = i % nfft, pair = i / nfft;
        const double2 a = FA[(pair / nB) * nfft + k], b = FB[(pair % nB) * nfft + k];
        Cx[i] = make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
    }
}
// full correlation c[k] = r[k - (Lb - 1)], k = 0 .. La + Lb - 2 (scipy.signal.correlate, mode 'full'); peak = first k that
// maximises |c[k]| (numpy argmax), value = c[k]
__global__ void k_xcorr_peak(const double2* __restrict__ Cx, int64_t nfft, int64_t La, int64_t Lb, int64_t* __restrict__ idx,
                             double* __restrict__ val) {
    __shared__ double sv[32], sa[32];
    __shared__ long long si[32];
    const double2* c = Cx + (int64_t)blockIdx.x * nfft;
    const double inv = 1.0 / (double)nfft;
    double best_a = -1.0, best_v = 0.0;
    long long best_k = 0;
    for (int64_t k = threadIdx.x; k < La + Lb - 1; k += blockDim.x) {
        int64_t m = k - (Lb - 1);
        if (m < 0) m += nfft;
        const double v = c[m].x * inv, a = fabs(v);
        if (a > best_a) { best_a = a; best_v = v; best_k = k; }   // per-thread k is increasing: keeps the first maximum
    }
    auto better = [](double a1, long long k1, double a2, long long k2) { return a1 > a2 || (a1 == a2 && k1 < k2); };
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double oa = __shfl_xor_sync(0xffffffffu, best_a, o), ov = __shfl_xor_sync(0xffffffffu, best_v, o);
        const long long ok = __shfl_xor_sync(0xffffffffu, best_k, o);
        if (better(oa, ok, best_a, best_k)) { best_a = oa; best_v = ov; best_k = ok; }
    }
    if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = best_a; sv[threadIdx.x >> 5] = best_v; si[threadIdx.x >> 5] = best_k; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (better(sa[w], si[w], sa[0], si[0])) { sa[0] = sa[w]; sv[0] = sv[w]; si[0] = si[w]; }
        idx[blockIdx.x] = si[0];
        val[blockIdx.x] = sv[0];
    }
}
// out[n][k] = conj?( rot_k * tx[(n + delay_k) mod L][swap_k] )     (core.py:648-666)
__global__ void k_sync_apply(const double2* __restrict__ tx, double2* __restrict__ out, int64_t L, int nCols,
                             const int32_t* __restrict__ swap, const double2* __restrict__ rot,
                             const int32_t* __restrict__ conj_flag, const int64_t* __restrict__ delay) {
    const int64_t total = L * nCols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % nCols);
        const int64_t n = i / nCols;
        int64_t src = (n + delay[k]) % L;
        if (src < 0) src += L;
        const double2 v = tx[src * nCols + swap[k]], r = rot[k];
        double2 o = make_double2(r.x * v.x - r.y * v.y, r.x * v.y + r.y * v.x);
        if (conj_flag[k]) o.y = -o.y;
        out[i] = o;
    }
}

int64_t next_pow2(int64_t n) {
    int64_t p = 1;
    while (p < n) p <<= 1;
    return p;
}

}  // namespace

extern "C" int ocb_pdm_frontend_run(const void* Es_rows, const void* Elo, void* S_rows, int64_t N, double polRotation,
                                    double pdl_dB, double R, double lo_power_w, double lo_freq_shift, double Fs,
                                    const double* iq_k /* k1x, k2x, k1y, k2y as (re, im) pairs */, void* stream) {
    OCB_REQUIRE(Es_rows && S_rows && N > 0 && iq_k, "pdm_frontend_run: bad argument");
    OCB_REQUIRE(R > 0, "pdm_frontend_run: photodiode responsivity must be positive (devices.py:306)");
    OCB_REQUIRE(Elo != nullptr || (lo_power_w > 0 && Fs > 0), "pdm_frontend_run: neither an LO field nor CW LO parameters");
    cudaStream_t st = (cudaStream_t)stream;
    const float gx = (float)pow(10.0, -(pdl_dB / 2.0) / 20.0), gy = (float)pow(10.0, (pdl_dB / 2.0) / 20.0);
    OCB_LAUNCH(k_pdm_frontend, grid_for(N, 256, 2), 256, 0, st, (const float2*)Es_rows, (const float2*)Elo, (float2*)S_rows, N,
               (float)cos(polRotation), (float)sin(polRotation), gx, gy, (float)R, sqrt(lo_power_w > 0 ? lo_power_w : 0.0),
               Fs > 0 ? lo_freq_shift / Fs : 0.0, make_double2(iq_k[0], iq_k[1]), make_double2(iq_k[2], iq_k[3]),
               make_double2(iq_k[4], iq_k[5]), make_double2(iq_k[6], iq_k[7]));
    return 0;
}

extern "C" int ocb_freq_shift_run(void* rows, int nRows, int64_t N, double freq, double Fs, void* stream) {
    OCB_REQUIRE(rows && nRows > 0 && N > 0 && Fs > 0, "freq_shift_run: bad argument");
    OCB_LAUNCH(k_freq_shift, grid_for((int64_t)nRows * N, 256, 2), 256, 0, (cudaStream_t)stream, (float2*)rows, nRows, N, freq / Fs);
    return 0;
}

extern "C" int ocb_sync_sequence_run(const void* z_dev, int nCols, int64_t L, int kind, void* out_rows, void* stream) {
    OCB_REQUIRE(z_dev && out_rows && nCols > 0 && L > 0 && kind >= 0 && kind <= 2, "sync_sequence_run: bad argument");
    OCB_LAUNCH(k_sync_sequence, nCols, 1024, 0, (cudaStream_t)stream, (const double2*)z_dev, nCols, L, kind, (double*)out_rows);
    return 0;
}

extern "C" int64_t ocb_xcorr_workspace_bytes(int nA, int64_t La, int nB, int64_t Lb) {
    if (nA <= 0 || nB <= 0 || La <= 0 || Lb <= 0) return -1;
    const int64_t nfft = next_pow2(La + Lb - 1);
    return ((int64_t)nA + nB + (int64_t)nA * nB) * nfft * 16 + (int64_t)nA * nB * 16 + 1024;
}

extern "C" int ocb_xcorr_peak_run(const void* a_rows, int nA, int64_t La, const void* b_rows, int nB, int64_t Lb,
                                  int64_t* peak_idx_host, double* peak_val_host, void* workspace, int64_t workspace_bytes,
                                  void* stream) {
    OCB_REQUIRE(a_rows && b_rows && peak_idx_host && peak_val_host && workspace, "xcorr_peak_run: NULL argument");
    OCB_REQUIRE(workspace_bytes >= ocb_xcorr_workspace_bytes(nA, La, nB, Lb) && ocb_xcorr_workspace_bytes(nA, La, nB, Lb) > 0,
                "xcorr_peak_run: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nfft = next_pow2(La + Lb - 1);
    OCB_REQUIRE(nfft < (1ll << 31), "xcorr_peak_run: sequences too long");
    const int P = nA * nB;
    double2* FA = (double2*)workspace;
    double2* FB = FA + (int64_t)nA * nfft;
    double2* Cx = FB + (int64_t)nB * nfft;
    int64_t* d_idx = (int64_t*)(Cx + (int64_t)P * nfft);
    double* d_val = (double*)(d_idx + P);
    OCB_LAUNCH(k_real_to_padded, grid_for((int64_t)nA * nfft, 256, 2), 256, 0, st, (const double*)a_rows, nA, La, nfft, FA);
    OCB_LAUNCH(k_real_to_padded, grid_for((int64_t)nB * nfft, 256, 2), 256, 0, st, (const double*)b_rows, nB, Lb, nfft, FB);
    cufftHandle pa, pb, pc;
    OCB_CUFFT(fft_plan_cached(CUFFT_Z2Z, (int)nfft, nA, st, &pa));
    OCB_CUFFT(cufftExecZ2Z(pa, (cufftDoubleComplex*)FA, (cufftDoubleComplex*)FA, CUFFT_FORWARD));
    OCB_CUFFT(fft_plan_cached(CUFFT_Z2Z, (int)nfft, nB, st, &pb));
    OCB_CUFFT(cufftExecZ2Z(pb, (cufftDoubleComplex*)FB, (cufftDoubleComplex*)FB, CUFFT_FORWARD));
    OCB_LAUNCH(k_xcorr_mul, grid_for((int64_t)P * nfft, 256, 2), 256, 0, st, FA, FB, nA, nB, nfft, Cx);
    OCB_CUFFT(fft_plan_cached(CUFFT_Z2Z, (int)nfft, P, st, &pc));
    OCB_CUFFT(cufftExecZ2Z(pc, (cufftDoubleComplex*)Cx, (cufftDoubleComplex*)Cx, CUFFT_INVERSE));
    OCB_LAUNCH(k_xcorr_peak, P, 1024, 0, st, Cx, nfft, La, Lb, d_idx, d_val);
    OCB_CUDA(cudaMemcpyAsync(peak_idx_host, d_idx, P * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    OCB_CUDA(cudaMemcpyAsync(peak_val_host, d_val, P * sizeof(double), cudaMemcpyDeviceToHost, st));
    OCB_CUDA(cudaStreamSynchronize(st));  // the decisions of symbolSync are taken on the host from these scalars
    return 0;
}

extern "C" int ocb_sync_apply_run(const void* tx_dev, void* out_dev, int64_t L, int nCols, const int32_t* swap_dev,
                                  const void* rot_dev, const int32_t* conj_dev, const int64_t* delay_dev, void* stream) {
    OCB_REQUIRE(tx_dev && out_dev && swap_dev && rot_dev && conj_dev && delay_dev && L > 0 && nCols > 0, "sync_apply_run: bad argument");
    OCB_LAUNCH(k_sync_apply, grid_for(L * nCols, 256, 2), 256, 0, (cudaStream_t)stream, (const double2*)tx_dev, (double2*)out_dev, L,
               nCols, swap_dev, (const double2*)rot_dev, conj_dev, delay_dev);
    return 0;
}
