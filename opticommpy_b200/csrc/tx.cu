// WDM transmitter on the device (the input generator of the fiber model at scale).
// Reference behaviour restated (not code):
//   simpleWDMTx                      : optic/models/tx.py:42-228   (per channel and mode: upsample -> pulse shaping ->
//                                      amplitude normalisation -> IQ modulator -> power normalisation -> frequency
//                                      shift -> accumulate on the mode's WDM field)
//   upsample                         : optic/dsp/core.py:395-432
//   iqm / mzm / pm, calcMZM / calcPM : optic/models/devices.py:94-220, optic/dsp/core.py:1075-1130
//   pnorm / signalPower / freqShift  : optic/dsp/core.py:702-717, 65-84, 1050-1072
// The pulse-shaping convolution itself is ocb_edc_run (firFilter = 'same' linear convolution, rxdsp.cu).
#include <math.h>

#include <algorithm>
#include <vector>

#include "../../include/opticomm_b200.h"
#include "common.cuh"

using namespace ocb;

namespace {

// rows[r][SpS*k] = sym[r][k], zero elsewhere  (core.py:425-432)
__global__ void k_upsample(const float2* __restrict__ sym, int nRows, int64_t nSym, int SpS, float2* __restrict__ rows) {
    const int64_t N = nSym * SpS, total = (int64_t)nRows * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = i % N, r = i / N;
        rows[i] = (n % SpS == 0) ? sym[r * nSym + n / SpS] : make_float2(0.f, 0.f);
    }
}

// amax2[r] = max_n |rows[r][n]|^2 (tx.py:196 divides by max|.|); non-negative floats order like their bit patterns
__global__ void k_row_absmax2(const float2* __restrict__ rows, int64_t N, unsigned* __restrict__ amax2) {
    __shared__ float sh[32];
    const int r = blockIdx.y;
    const float2* x = rows + (int64_t)r * N;
    float m = 0.f;
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, cabs2(x[n]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
        m = warp_max(m);
        if (threadIdx.x == 0) atomicMax(amax2 + r, __float_as_uint(m));
    }
}

struct IqmConst {  // host-evaluated constants of the two MZMs and the phase modulator
    double aI, bI, aQ, bQ;   // (sqrt(1+g) + sqrt(1-g)) / (2 sqrt 2) and (sqrt(1+g) - sqrt(1-g)) / (2 sqrt 2) per arm
    double kI, oI, kQ, oQ;   // phase of an arm: k * Re/Im(u) + o = pi (u + Vb) / (2 Vpi), u = mzmScale * s / max|s|
    double rx, ry;           // exp(j pi Vphi / Vpi): rotation of the Q arm (pm, devices.py:214)
};

// IQM output for one pulse-shaped sample s (amplitude-normalised by inv_max), LO sample lo  (devices.py:210-216;
// calcMZM, core.py:1103-1108: sqrt(1+g) PM(E/2, (u+Vb)/2) + sqrt(1-g) PM(E/2, -(u+Vb)/2) with the input E/sqrt(2))
__device__ __forceinline__ double2 iqm_sample(float2 s, double inv_max, double2 lo, const IqmConst& c) {
    double sI, cI, sQ, cQ;
    sincos(c.kI * ((double)s.x * inv_max) + c.oI, &sI, &cI);
    sincos(c.kQ * ((double)s.y * inv_max) + c.oQ, &sQ, &cQ);
    const double2 mI = make_double2(c.aI * cI, c.bI * sI), mQ = make_double2(c.aQ * cQ, c.bQ * sQ);
    const double2 eI = make_double2(lo.x * mI.x - lo.y * mI.y, lo.x * mI.y + lo.y * mI.x);
    const double2 eQ = make_double2(lo.x * mQ.x - lo.y * mQ.y, lo.x * mQ.y + lo.y * mQ.x);
    return make_double2(eI.x + eQ.x * c.rx - eQ.y * c.ry, eI.y + eQ.x * c.ry + eQ.y * c.rx);
}

// power[r] = sum_n |IQM(rows[r][n])|^2   (pnorm, core.py:717)
__global__ void k_iqm_power(const float2* __restrict__ rows, const float2* __restrict__ lo, int nPol, int64_t N,
                            const unsigned* __restrict__ amax2, IqmConst c, double* __restrict__ power) {
    __shared__ double sh[32];
    const int r = blockIdx.y;
    const float2* x = rows + (int64_t)r * N;
    const float2* l = lo ? lo + (int64_t)(r / nPol) * N : nullptr;
    const double inv_max = 1.0 / sqrt((double)__uint_as_float(amax2[r]));
    double acc = 0.0;
    for (int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; n < N; n += (int64_t)gridDim.x * blockDim.x) {
        const double2 lv = l ? make_double2(l[n].x, l[n].y) : make_double2(1.0, 0.0);
        const double2 e = iqm_sample(x[n], inv_max, lv, c);
        acc += e.x * e.x + e.y * e.y;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
        atomicAdd(power + r, t);
    }
}

// out[m][n] = sum_ch sqrt(Pch / nPol) IQM(rows[ch*nPol + m][n]) / sqrt(power / N) * exp(j 2 pi f_ch n / Fs), channels
// added in ascending order like the reference's loop (tx.py:207-209)
__global__ void k_wdm_combine(const float2* __restrict__ rows, const float2* __restrict__ lo, int nCh, int nPol, int64_t N,
                              const unsigned* __restrict__ amax2, const double* __restrict__ power,
                              const double* __restrict__ ch_power, const double* __restrict__ ch_cycles, IqmConst c,
                              float2* __restrict__ out) {
    const int64_t total = (int64_t)nPol * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t n = i % N;
        const int m = (int)(i / N);
        double accx = 0.0, accy = 0.0;
        for (int ch = 0; ch < nCh; ++ch) {
            const int r = ch * nPol + m;
            const double inv_max = 1.0 / sqrt((double)__uint_as_float(amax2[r]));
            const double2 lv = lo ? make_double2(lo[(int64_t)ch * N + n].x, lo[(int64_t)ch * N + n].y) : make_double2(1.0, 0.0);
            const double2 e = iqm_sample(rows[(int64_t)r * N + n], inv_max, lv, c);
            const double g = sqrt(ch_power[ch] / (double)nPol) / sqrt(power[r] / (double)N);
            double ph = ch_cycles[ch] * (double)n;
            ph -= floor(ph);
            double s, co;
            sincospi(2.0 * ph, &s, &co);
            accx += g * (e.x * co - e.y * s);
            accy += g * (e.x * s + e.y * co);
        }
        out[i] = make_float2((float)accx, (float)accy);
    }
}

}  // namespace

extern "C" int ocb_upsample_run(const void* sym_rows, int nRows, int64_t nSym, int SpS, void* rows_out, void* stream) {
    OCB_REQUIRE(sym_rows && rows_out && nRows > 0 && nSym > 0 && SpS > 0, "upsample_run: bad argument");
    OCB_LAUNCH(k_upsample, grid_for((int64_t)nRows * nSym * SpS, 256, 2), 256, 0, (cudaStream_t)stream, (const float2*)sym_rows,
               nRows, nSym, SpS, (float2*)rows_out);
    return 0;
}

extern "C" int64_t ocb_wdm_tx_workspace_bytes(int nCh, int nPol) {
    if (nCh <= 0 || nPol <= 0) return -1;
    const int64_t rows = (int64_t)nCh * nPol;
    return 256 + rows * 4 + 256 + rows * 8 + 256 + 2 * (int64_t)nCh * 8 + 256;
}

extern "C" int ocb_wdm_tx_combine_run(const void* shaped_rows, const void* lo_rows, int nCh, int nPol, int64_t N,
                                      const ocb_wdm_tx_params* q, const double* ch_power_w, const double* ch_freq_hz,
                                      double Fs, void* out_rows, void* workspace, int64_t workspace_bytes, void* stream) {
    OCB_REQUIRE(shaped_rows && out_rows && q && ch_power_w && ch_freq_hz && workspace, "wdm_tx_combine_run: NULL argument");
    OCB_REQUIRE(nCh > 0 && nPol > 0 && N > 0 && Fs > 0, "wdm_tx_combine_run: bad geometry");
    OCB_REQUIRE(workspace_bytes >= ocb_wdm_tx_workspace_bytes(nCh, nPol), "wdm_tx_combine_run: workspace too small");
    OCB_REQUIRE(q->Vpi != 0.0, "wdm_tx_combine_run: Vpi must be non-zero");
    cudaStream_t st = (cudaStream_t)stream;
    const int rows = nCh * nPol;
    auto up = [](int64_t x) { return (x + 255) / 256 * 256; };
    char* w = (char*)workspace;
    unsigned* amax2 = (unsigned*)w; w += up((int64_t)rows * 4);
    double* power = (double*)w; w += up((int64_t)rows * 8);
    double* chp = (double*)w;
    double* chc = chp + nCh;
    OCB_CUDA(cudaMemsetAsync(amax2, 0, (size_t)rows * 4, st));
    OCB_CUDA(cudaMemsetAsync(power, 0, (size_t)rows * 8, st));
    // small per-channel tables from pageable host memory: cudaMemcpyAsync returns once such a buffer has been staged,
    // so neither the caller's arrays nor the local one need to outlive this call and no stream synchronisation is needed
    std::vector<double> cyc(nCh);
    for (int k = 0; k < nCh; ++k) {
        OCB_REQUIRE(ch_power_w[k] > 0.0, "wdm_tx_combine_run: channel powers must be positive");
        cyc[k] = ch_freq_hz[k] / Fs;
    }
    OCB_CUDA(cudaMemcpyAsync(chp, ch_power_w, (size_t)nCh * 8, cudaMemcpyHostToDevice, st));
    OCB_CUDA(cudaMemcpyAsync(chc, cyc.data(), (size_t)nCh * 8, cudaMemcpyHostToDevice, st));
    IqmConst c;
    const double pi = 3.14159265358979323846;
    auto arm = [&](double ER, double Vb, double* a, double* b, double* k, double* o) {
        const double er = pow(10.0, ER / 10.0), g = 2.0 * sqrt(er) / (er + 1.0);  // core.py:1101-1102
        const double p = sqrt(1.0 + g), m = sqrt(1.0 - g), s = 1.0 / (2.0 * sqrt(2.0));
        *a = (p + m) * s; *b = (p - m) * s;
        *k = pi * q->mzmScale / (2.0 * q->Vpi); *o = pi * Vb / (2.0 * q->Vpi);
    };
    arm(q->ERI, q->VbI, &c.aI, &c.bI, &c.kI, &c.oI);
    arm(q->ERQ, q->VbQ, &c.aQ, &c.bQ, &c.kQ, &c.oQ);
    c.rx = cos(pi * q->Vphi / q->Vpi); c.ry = sin(pi * q->Vphi / q->Vpi);
    const int gx = (int)std::min<int64_t>((N + 256 * 8 - 1) / (256 * 8), 4 * kNumSMs);
    OCB_LAUNCH(k_row_absmax2, dim3(gx, rows), 256, 0, st, (const float2*)shaped_rows, N, amax2);
    OCB_LAUNCH(k_iqm_power, dim3(gx, rows), 256, 0, st, (const float2*)shaped_rows, (const float2*)lo_rows, nPol, N, amax2, c, power);
    OCB_LAUNCH(k_wdm_combine, grid_for((int64_t)nPol * N, 256, 1), 256, 0, st, (const float2*)shaped_rows, (const float2*)lo_rows,
               nCh, nPol, N, amax2, power, chp, chc, c, (float2*)out_rows);
    return 0;
}
