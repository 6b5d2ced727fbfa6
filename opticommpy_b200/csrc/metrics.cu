// Hard decisions and Monte-Carlo error counting on the device:
//   minEuclid       optic/comm/modulation.py:271-299   idx = argmin_c |x - const[c]|, first index on ties
//   demodulateGray  optic/comm/modulation.py:369-408   bits of idx, most significant first (the bit map of the
//                   Gray-ordered constellation is the binary expansion of the index: minEuclid(const, const) = id)
//   fastBERcalc     optic/comm/metrics.py:110-195      per column: phase-ambiguity rotation mean(tx/rx), pnorm of
//                   both, SNR = P(tx)/P(rx - tx), decisions on sqrt(Es)*x, bit and symbol error counts
// Arithmetic is float64 whatever the storage type; all reductions are two-stage with a fixed grid, so results
// are run-to-run identical.  Samples are (L, nModes) interleaved, like the reference arrays.
#include <math.h>

#include <algorithm>
#include <vector>

#include "../../include/opticomm_b200.h"
#include "common.cuh"

using namespace ocb;

namespace {

constexpr int kMetricBlocks = 2 * kNumSMs;  // blocks per mode of the reduction passes
constexpr int kMaxConst = 4096;             // constellation points staged in shared memory (64 KB of double2)

template <typename T>
__device__ __forceinline__ double2 ldc(const T* __restrict__ p, int64_t i);
template <>
__device__ __forceinline__ double2 ldc<float2>(const float2* __restrict__ p, int64_t i) {
    const float2 v = p[i];
    return make_double2((double)v.x, (double)v.y);
}
template <>
__device__ __forceinline__ double2 ldc<double2>(const double2* __restrict__ p, int64_t i) { return p[i]; }

__device__ __forceinline__ void stage_const(double2* sc, const double2* __restrict__ c, int M) {
    for (int i = threadIdx.x; i < M; i += blockDim.x) sc[i] = c[i];
    __syncthreads();
}

// nearest point; strict '<' keeps the first index on exact ties (np.argmin)
__device__ __forceinline__ int decide(double2 v, const double2* sc, int M) {
    double best = INFINITY;
    int bi = 0;
    for (int c = 0; c < M; ++c) {
        const double dx = v.x - sc[c].x, dy = v.y - sc[c].y;
        const double d = dx * dx + dy * dy;
        if (d < best) { best = d; bi = c; }
    }
    return bi;
}

template <typename T>
__global__ void __launch_bounds__(256)
k_min_euclid(const T* __restrict__ x, int64_t n, const double2* __restrict__ constSymb, int M, int nbits,
             int64_t* __restrict__ idx, int64_t* __restrict__ bits) {
    extern __shared__ double2 sc_me[];
    stage_const(sc_me, constSymb, M);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int a = decide(ldc<T>(x, i), sc_me, M);
        if (idx) idx[i] = a;
        if (bits)
            for (int b = 0; b < nbits; ++b) bits[i * nbits + b] = (a >> (nbits - 1 - b)) & 1;
    }
}

// block-level sum of NV doubles per thread -> dst[0..NV) written by thread 0 (fixed order)
template <int NV>
__device__ __forceinline__ void block_sum_store(double (&v)[NV], double* dst) {
    __shared__ double sh[NV][8];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        v[k] = warp_sum(v[k]);
        if ((threadIdx.x & 31) == 0) sh[k][threadIdx.x >> 5] = v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += sh[k][w];
            dst[k] = t;
        }
    }
}

// pass 1 (grid: kMetricBlocks x nModes): partial sums of tx/rx, |rx|², |tx|² of column blockIdx.y
template <typename T>
__global__ void __launch_bounds__(256)
k_ber_sums(const T* __restrict__ rx, const T* __restrict__ tx, int64_t L, int nModes, double* __restrict__ partials) {
    const int m = blockIdx.y;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < L; k += (int64_t)gridDim.x * blockDim.x) {
        const double2 r = ldc<T>(rx, k * nModes + m), t = ldc<T>(tx, k * nModes + m);
        const double r2 = r.x * r.x + r.y * r.y;
        acc[0] += (t.x * r.x + t.y * r.y) / r2;  // t / r = t conj(r) / |r|²
        acc[1] += (t.y * r.x - t.x * r.y) / r2;
        acc[2] += r2;
        acc[3] += t.x * t.x + t.y * t.y;
    }
    block_sum_store<4>(acc, partials + ((int64_t)m * gridDim.x + blockIdx.x) * 4);
}

// per column: rot = mean(tx/rx) (or 1), norms of the rotated rx and of tx   (metrics.py:176-182)
__global__ void k_ber_scalars(const double* __restrict__ partials, int nblk, int64_t L, int nModes, int rotate,
                              double* __restrict__ scal) {
    const int m = threadIdx.x;
    if (m >= nModes) return;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int b = 0; b < nblk; ++b)
        for (int k = 0; k < 4; ++k) s[k] += partials[((int64_t)m * nblk + b) * 4 + k];
    double2 rot = rotate ? make_double2(s[0] / (double)L, s[1] / (double)L) : make_double2(1.0, 0.0);
    scal[m * 4 + 0] = rot.x;
    scal[m * 4 + 1] = rot.y;
    scal[m * 4 + 2] = sqrt((rot.x * rot.x + rot.y * rot.y) * s[2] / (double)L);
    scal[m * 4 + 3] = sqrt(s[3] / (double)L);
}

// pass 2: normalise, accumulate signal / error power, decide both sequences, count bit and symbol errors
template <typename T>
__global__ void __launch_bounds__(256)
k_ber_count(const T* __restrict__ rx, const T* __restrict__ tx, int64_t L, int nModes,
            const double2* __restrict__ constSymb, int M, double sqrtEs, const double* __restrict__ scal,
            double* __restrict__ fpart, unsigned long long* __restrict__ ipart) {
    extern __shared__ double2 sc_bc[];
    stage_const(sc_bc, constSymb, M);
    const int m = blockIdx.y;
    const double2 rot = make_double2(scal[m * 4 + 0], scal[m * 4 + 1]);
    const double nrx = scal[m * 4 + 2], ntx = scal[m * 4 + 3];
    double acc[2] = {0.0, 0.0};
    unsigned long long nbit = 0, nsym = 0;
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < L; k += (int64_t)gridDim.x * blockDim.x) {
        const double2 r = ldc<T>(rx, k * nModes + m), t = ldc<T>(tx, k * nModes + m);
        const double2 rn = make_double2((rot.x * r.x - rot.y * r.y) / nrx, (rot.x * r.y + rot.y * r.x) / nrx);
        const double2 tn = make_double2(t.x / ntx, t.y / ntx);
        acc[0] += tn.x * tn.x + tn.y * tn.y;
        const double ex = rn.x - tn.x, ey = rn.y - tn.y;
        acc[1] += ex * ex + ey * ey;
        const int a = decide(make_double2(sqrtEs * rn.x, sqrtEs * rn.y), sc_bc, M);
        const int b = decide(make_double2(sqrtEs * tn.x, sqrtEs * tn.y), sc_bc, M);
        nbit += __popc(a ^ b);
        nsym += (a != b);
    }
    block_sum_store<2>(acc, fpart + ((int64_t)m * gridDim.x + blockIdx.x) * 2);
    __shared__ unsigned long long si[2][8];
    for (int o = 16; o > 0; o >>= 1) {
        nbit += __shfl_xor_sync(0xffffffffu, nbit, o);
        nsym += __shfl_xor_sync(0xffffffffu, nsym, o);
    }
    if ((threadIdx.x & 31) == 0) { si[0][threadIdx.x >> 5] = nbit; si[1][threadIdx.x >> 5] = nsym; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tb = 0, ts = 0;
        for (int w = 0; w < 8; ++w) { tb += si[0][w]; ts += si[1][w]; }
        ipart[((int64_t)m * gridDim.x + blockIdx.x) * 2 + 0] = tb;
        ipart[((int64_t)m * gridDim.x + blockIdx.x) * 2 + 1] = ts;
    }
}

// per column totals: res[m] = {Σ|tn|², Σ|rn - tn|², bit errors, symbol errors} (counts stored as doubles < 2^53)
__global__ void k_ber_totals(const double* __restrict__ fpart, const unsigned long long* __restrict__ ipart, int nblk,
                             int nModes, double* __restrict__ res) {
    const int m = threadIdx.x;
    if (m >= nModes) return;
    double a = 0.0, e = 0.0;
    unsigned long long nb = 0, ns = 0;
    for (int b = 0; b < nblk; ++b) {
        a += fpart[((int64_t)m * nblk + b) * 2 + 0];
        e += fpart[((int64_t)m * nblk + b) * 2 + 1];
        nb += ipart[((int64_t)m * nblk + b) * 2 + 0];
        ns += ipart[((int64_t)m * nblk + b) * 2 + 1];
    }
    res[m * 4 + 0] = a;
    res[m * 4 + 1] = e;
    res[m * 4 + 2] = (double)nb;
    res[m * 4 + 3] = (double)ns;
}

int ilog2i(int M) {
    int b = 0;
    while ((1 << (b + 1)) <= M) ++b;
    return b;
}

template <typename T>
int min_euclid_impl(const void* x, int64_t n, const void* constSymb, int M, int nbits, int64_t* idx, int64_t* bits,
                    cudaStream_t st) {
    const size_t smem = (size_t)M * sizeof(double2);
    if (smem > 48 * 1024)
        OCB_CUDA(cudaFuncSetAttribute(k_min_euclid<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    OCB_LAUNCH(k_min_euclid<T>, grid_for(n, 256, 4), 256, smem, st, (const T*)x, n, (const double2*)constSymb, M, nbits, idx, bits);
    return 0;
}

template <typename T>
int ber_impl(const void* rx, const void* tx, int64_t L, int nModes, const void* constSymb, int M, int rotate,
             double sqrtEs, double* fpart1, double* scal, double* fpart2, unsigned long long* ipart, double* res,
             cudaStream_t st) {
    const int nblk = (int)std::min<int64_t>(kMetricBlocks, (L + 255) / 256);
    const dim3 grid(nblk, nModes);
    const size_t smem = (size_t)M * sizeof(double2);
    if (smem > 48 * 1024)
        OCB_CUDA(cudaFuncSetAttribute(k_ber_count<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    OCB_LAUNCH(k_ber_sums<T>, grid, 256, 0, st, (const T*)rx, (const T*)tx, L, nModes, fpart1);
    OCB_LAUNCH(k_ber_scalars, 1, 32, 0, st, fpart1, nblk, L, nModes, rotate, scal);
    OCB_LAUNCH(k_ber_count<T>, grid, 256, smem, st, (const T*)rx, (const T*)tx, L, nModes, (const double2*)constSymb, M,
               sqrtEs, scal, fpart2, ipart);
    OCB_LAUNCH(k_ber_totals, 1, 32, 0, st, fpart2, ipart, nblk, nModes, res);
    return 0;
}

}  // namespace

extern "C" int ocb_min_euclid(const void* x_dev, int x_dtype, int64_t n, const void* constSymb, int M,
                              int64_t* idx_out, int64_t* bits_out, void* stream) {
    OCB_REQUIRE(x_dev && constSymb && (idx_out || bits_out), "null pointer");
    OCB_REQUIRE(n >= 0 && M >= 1 && M <= kMaxConst, "constellation size out of range");
    OCB_REQUIRE(x_dtype == OCB_C64 || x_dtype == OCB_C128, "unsupported sample dtype");
    const int nbits = ilog2i(M);
    OCB_REQUIRE(!bits_out || (1 << nbits) == M, "bit demapping needs M to be a power of two");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    return x_dtype == OCB_C64 ? min_euclid_impl<float2>(x_dev, n, constSymb, M, nbits, idx_out, bits_out, st)
                              : min_euclid_impl<double2>(x_dev, n, constSymb, M, nbits, idx_out, bits_out, st);
}

extern "C" int64_t ocb_ber_workspace_bytes(int nModes) {
    if (nModes < 1) return 0;
    // pass-1 partials (4) + pass-2 partials (2 double + 2 integer) per block and column, scalars and totals per column
    return (int64_t)nModes * kMetricBlocks * 8 * (int64_t)sizeof(double) + (int64_t)nModes * 8 * (int64_t)sizeof(double) + 256;
}

extern "C" int ocb_ber_count(const void* rx_dev, const void* tx_dev, int dtype, int64_t L, int nModes,
                             const void* constSymb, int M, int rotate, double sqrtEs, double* ber_host,
                             double* ser_host, double* snr_host, int64_t* counts_host, void* workspace,
                             int64_t workspace_bytes, void* stream) {
    OCB_REQUIRE(rx_dev && tx_dev && constSymb && workspace, "null pointer");
    OCB_REQUIRE(L >= 1 && nModes >= 1 && nModes <= 32, "shape out of range (1 <= nModes <= 32)");
    OCB_REQUIRE(M >= 2 && M <= kMaxConst && (M & (M - 1)) == 0, "M must be a power of two <= 4096");
    OCB_REQUIRE(dtype == OCB_C64 || dtype == OCB_C128, "unsupported sample dtype");
    OCB_REQUIRE(workspace_bytes >= ocb_ber_workspace_bytes(nModes), "workspace too small");
    OCB_REQUIRE(((uintptr_t)workspace & 7) == 0, "workspace must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    double* fpart1 = (double*)workspace;
    double* fpart2 = fpart1 + (int64_t)nModes * kMetricBlocks * 4;
    unsigned long long* ipart = (unsigned long long*)(fpart2 + (int64_t)nModes * kMetricBlocks * 2);
    double* scal = (double*)(ipart + (int64_t)nModes * kMetricBlocks * 2);
    double* res = scal + (int64_t)nModes * 4;
    const int rc = dtype == OCB_C64
                       ? ber_impl<float2>(rx_dev, tx_dev, L, nModes, constSymb, M, rotate, sqrtEs, fpart1, scal, fpart2, ipart, res, st)
                       : ber_impl<double2>(rx_dev, tx_dev, L, nModes, constSymb, M, rotate, sqrtEs, fpart1, scal, fpart2, ipart, res, st);
    if (rc) return rc;
    std::vector<double> h((size_t)nModes * 4);
    OCB_CUDA(cudaMemcpyAsync(h.data(), res, h.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    OCB_CUDA(cudaStreamSynchronize(st));
    const int nbits = ilog2i(M);
    for (int m = 0; m < nModes; ++m) {
        const double pt = h[m * 4 + 0] / (double)L, pe = h[m * 4 + 1] / (double)L;
        if (snr_host) snr_host[m] = 10.0 * log10(pt / pe);                           // metrics.py:185
        if (ber_host) ber_host[m] = h[m * 4 + 2] / ((double)L * (double)nbits);      // :191
        if (ser_host) ser_host[m] = h[m * 4 + 3] / (double)L;                        // :192
        if (counts_host) { counts_host[m] = (int64_t)h[m * 4 + 2]; counts_host[nModes + m] = (int64_t)h[m * 4 + 3]; }
    }
    return 0;
}

// =============================================================================================
// Rx front-end glue: decimate (optic/dsp/core.py:435-491).
//   varVector[p] = var(sigIn[p::SpSin, k])  (numpy var of a complex column: mean |x - mean|^2)
//   sampDelay[k] = first p with varVector[p] == max      (maximum-variance sampling instant)
//   sigOut[i, k] = roll(sigIn[:, k], -sampDelay[k])[i * decFactor]
// Planar rows x[nModes][N] complex64 in, y[nModes][ceil(N / decFactor)] out; sums in float64.
// =============================================================================================
namespace {

__global__ void __launch_bounds__(256)
k_phase_moments(const float2* __restrict__ x, int64_t N, int SpSin, double* __restrict__ mom /* [nModes][SpSin][3] */) {
    const int p = blockIdx.x, mode = blockIdx.y;
    const float2* xs = x + (int64_t)mode * N;
    double sr = 0.0, si = 0.0, sq = 0.0;
    for (int64_t i = p + (int64_t)threadIdx.x * SpSin; i < N; i += (int64_t)blockDim.x * SpSin) {
        const float2 v = xs[i];
        sr += (double)v.x; si += (double)v.y;
        sq += (double)v.x * (double)v.x + (double)v.y * (double)v.y;
    }
    __shared__ double sh[3][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    sr = warp_sum(sr); si = warp_sum(si); sq = warp_sum(sq);
    if (lane == 0) { sh[0][wid] = sr; sh[1][wid] = si; sh[2][wid] = sq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0, b = 0, c = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += sh[0][w]; b += sh[1][w]; c += sh[2][w]; }
        double* o = mom + ((int64_t)mode * SpSin + p) * 3;
        o[0] = a; o[1] = b; o[2] = c;
    }
}
__global__ void k_pick_delay(const double* __restrict__ mom, int64_t N, int SpSin, int nModes, int32_t* __restrict__ delay) {
    const int mode = blockIdx.x * blockDim.x + threadIdx.x;
    if (mode >= nModes) return;
    const double cnt = (double)(N / SpSin);
    double best = -1.0;
    int bi = 0;
    for (int p = 0; p < SpSin; ++p) {
        const double* o = mom + ((int64_t)mode * SpSin + p) * 3;
        const double mr = o[0] / cnt, mi = o[1] / cnt;
        const double var = o[2] / cnt - (mr * mr + mi * mi);
        if (var > best) { best = var; bi = p; }  // first index of the maximum (core.py:477)
    }
    delay[mode] = bi;
}
__global__ void k_decimate_gather(const float2* __restrict__ x, float2* __restrict__ y, int64_t N, int64_t Nout,
                                  int decFactor, int nModes, const int32_t* __restrict__ delay) {
    const int64_t total = Nout * nModes;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t mode = i / Nout, j = i % Nout;
        int64_t src = j * decFactor + delay[mode];  // np.roll(x, -d)[j*dec] = x[(j*dec + d) mod N]
        if (src >= N) src -= N;
        y[mode * Nout + j] = x[mode * N + src];
    }
}

}  // namespace

extern "C" int64_t ocb_decimate_workspace_bytes(int nModes, int SpSin) {
    if (nModes <= 0 || SpSin <= 0) return -1;
    return (int64_t)nModes * SpSin * 3 * (int64_t)sizeof(double) + 256;
}

extern "C" int ocb_decimate_run(const void* x_rows, void* y_rows, int64_t N, int nModes, int SpSin, int decFactor,
                                void* delays_dev, void* workspace, int64_t workspace_bytes, void* stream) {
    OCB_REQUIRE(x_rows && y_rows && delays_dev && workspace, "decimate_run: NULL argument");
    OCB_REQUIRE(N > 0 && nModes > 0 && SpSin > 0 && decFactor > 0, "decimate_run: bad sizes");
    OCB_REQUIRE(N % SpSin == 0, "decimate_run: the signal length must be a multiple of SpSin (reshape(-1, SpSin), core.py:475)");
    OCB_REQUIRE(SpSin <= 65535 && nModes <= 65535, "decimate_run: SpSin / nModes too large");
    OCB_REQUIRE(workspace_bytes >= ocb_decimate_workspace_bytes(nModes, SpSin), "decimate_run: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    double* mom = (double*)workspace;
    const int64_t Nout = (N + decFactor - 1) / decFactor;
    OCB_LAUNCH(k_phase_moments, dim3(SpSin, nModes), 256, 0, st, (const float2*)x_rows, N, SpSin, mom);
    OCB_LAUNCH(k_pick_delay, (nModes + 63) / 64, 64, 0, st, mom, N, SpSin, nModes, (int32_t*)delays_dev);
    OCB_LAUNCH(k_decimate_gather, grid_for(Nout * nModes, 256, 2), 256, 0, st, (const float2*)x_rows, (float2*)y_rows, N,
               Nout, decFactor, nModes, (const int32_t*)delays_dev);
    return 0;
}
