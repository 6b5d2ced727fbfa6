// Device-side carrier phase recovery wrapper: everything optic.dsp.carrierRecovery.cpr does around
// bps (optic/dsp/carrierRecovery.py:110-169), in float64 like the reference:
//   fourthPowerFOE (:333-371)  ->  pnorm (core.py:702-717)  ->  bps (:172-223, ocb_bps_run)
//   ->  unwrap(4φ)/4 (:154)  ->  pnorm(x · e^{jφ}) (:162)
// Samples are (L, nModes) interleaved complex128, like the reference arrays.
#include <math.h>

#include <vector>

#include "../../include/opticomm_b200.h"
#include "common.cuh"
#include "plan_cache.cuh"

using namespace ocb;

namespace {

constexpr int kScanBlock = 1024;

// planar z[n][k] = x[k][n]^M   (M-th power spectrum input of the FOE, :366)
__global__ void k_foe_power(const double2* __restrict__ x, double2* __restrict__ z, int64_t L, int nModes, int M) {
    const int64_t total = L * nModes;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = i / nModes;
        const int n = (int)(i % nModes);
        const double2 v = x[i];
        double2 r = v;
        for (int p = 1; p < M; ++p) r = make_double2(r.x * v.x - r.y * v.y, r.x * v.y + r.y * v.x);
        z[(int64_t)n * L + k] = r;
    }
}

// argmax over the fftshift-ed spectrum position j (first maximum wins, :366-367); one block per mode
__global__ void __launch_bounds__(1024)
k_foe_argmax(const double2* __restrict__ Z, int64_t L, int64_t* __restrict__ pos_out) {
    __shared__ double sv[1024];
    __shared__ long long sj[1024];
    const double2* z = Z + (int64_t)blockIdx.x * L;
    const int64_t half = L / 2;  // numpy.fft.fftshift rolls by L//2: shifted[j] = orig[(j - L//2) mod L]
    double best = -1.0;
    long long bj = 0x7fffffffffffffffLL;
    for (int64_t j = threadIdx.x; j < L; j += blockDim.x) {
        int64_t k = j - half;
        if (k < 0) k += L;
        const double2 v = z[k];
        const double a = v.x * v.x + v.y * v.y;
        if (a > best) { best = a; bj = j; }
    }
    sv[threadIdx.x] = best; sj[threadIdx.x] = bj;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            const double ov = sv[threadIdx.x + s];
            const long long oj = sj[threadIdx.x + s];
            if (ov > sv[threadIdx.x] || (ov == sv[threadIdx.x] && oj < sj[threadIdx.x])) { sv[threadIdx.x] = ov; sj[threadIdx.x] = oj; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) pos_out[blockIdx.x] = sj[0];
}

// x[k][n] *= exp(-j 2π fo_n t_k), t_k = k / Fs   (:363, :369), rounding sequence of the numpy expression
__global__ void k_foe_apply(double2* __restrict__ x, int64_t L, int nModes, const double* __restrict__ fo, double Fs) {
    const int64_t total = L * nModes;
    const double two_pi = 2.0 * 3.14159265358979323846;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = i / nModes;
        const int n = (int)(i % nModes);
        const double t = (double)k / Fs;
        const double arg = -((two_pi * fo[n]) * t);
        double s, c;
        sincos(arg, &s, &c);
        const double2 v = x[i];
        x[i] = make_double2(v.x * c - v.y * s, v.x * s + v.y * c);
    }
}

// partial sums of |x|² (two-stage, deterministic): partials[blockIdx.x]
__global__ void __launch_bounds__(256)
k_power_partials(const double2* __restrict__ x, int64_t n, double* __restrict__ partials) {
    __shared__ double sh[8];
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double2 v = x[i];
        acc += v.x * v.x + v.y * v.y;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sh[w];
        partials[blockIdx.x] = t;
    }
}
// x *= 1/sqrt(mean |x|²)   (pnorm); every thread sums the (few hundred) partials itself
__global__ void k_pnorm_scale(double2* __restrict__ x, int64_t n, const double* __restrict__ partials, int nparts) {
    double tot = 0.0;
    for (int i = 0; i < nparts; ++i) tot += partials[i];
    const double g = 1.0 / sqrt(tot / (double)n);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double2 v = x[i];
        x[i] = make_double2(v.x * g, v.y * g);
    }
}

// ---- unwrap(4φ)/4 in index space: φ = b·(π/2)/B, so 4φ moves on a circle of B points ------------------
// wrapped increment of the phase index (np.unwrap's rule: |d| <= B/2 keeps d)
__device__ __forceinline__ int wrap_inc(int d, int B) {
    if (2 * d > B) return d - B;
    if (2 * d < -B) return d + B;
    return d;
}
// pass 1: per block of 1024 symbols (one mode): inclusive scan of the wrapped increments, block totals
__global__ void __launch_bounds__(kScanBlock)
k_unwrap_scan1(const int32_t* __restrict__ idx, int32_t* __restrict__ u, int32_t* __restrict__ blocksum, int64_t L,
               int nModes, int B) {
    __shared__ int sh[kScanBlock];
    const int n = blockIdx.y;
    const int64_t k = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
    int w = 0;
    if (k < L) {
        const int cur = idx[k * nModes + n];
        w = (k == 0) ? cur : wrap_inc(cur - idx[(k - 1) * nModes + n], B);
    }
    sh[threadIdx.x] = w;
    __syncthreads();
    for (int off = 1; off < kScanBlock; off <<= 1) {
        const int v = ((int)threadIdx.x >= off) ? sh[threadIdx.x - off] : 0;
        __syncthreads();
        sh[threadIdx.x] += v;
        __syncthreads();
    }
    if (k < L) u[k * nModes + n] = sh[threadIdx.x];
    if (threadIdx.x == kScanBlock - 1) blocksum[(int64_t)n * gridDim.x + blockIdx.x] = sh[kScanBlock - 1];
}
// pass 2: exclusive scan of the block totals of one mode (single block, serial over <= 8192 entries per thread chunk)
__global__ void k_unwrap_scan2(int32_t* __restrict__ blocksum, int nblocks) {
    if (threadIdx.x != 0) return;
    int32_t* b = blocksum + (int64_t)blockIdx.x * nblocks;
    int run = 0;
    for (int i = 0; i < nblocks; ++i) { const int t = b[i]; b[i] = run; run += t; }
}
// pass 3: phase = (u + offset)·((π/2)/B);  y = x·e^{jφ}
__global__ void __launch_bounds__(256)
k_unwrap_apply(const int32_t* __restrict__ u, const int32_t* __restrict__ blockoff, int nblocks, const double2* __restrict__ x,
               double2* __restrict__ y, double* __restrict__ phase, int64_t L, int nModes, int B) {
    const int64_t total = L * nModes;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = i / nModes;
        const int n = (int)(i % nModes);
        const int ui = u[i] + blockoff[(int64_t)n * nblocks + k / kScanBlock];
        const double ph = ((double)ui * (M_PI / 2.0)) / (double)B;
        phase[i] = ph;
        double s, c;
        sincos(ph, &s, &c);
        const double2 v = x[i];
        y[i] = make_double2(v.x * c - v.y * s, v.x * s + v.y * c);
    }
}

// fo[m] = fftshift(Fs * fftfreq(L))[pos[m]] / foeM  (carrierRecovery.py:358-359, 368), same double arithmetic as numpy
__global__ void k_foe_freq(const int64_t* __restrict__ pos, double* __restrict__ fo, int64_t L, int nModes, double Fs,
                           int foeM) {
    for (int m = threadIdx.x; m < nModes; m += blockDim.x) {
        const int64_t j = pos[m];
        int64_t k = j - L / 2;
        if (k < 0) k += L;
        const int64_t kk = (k <= (L - 1) / 2) ? k : k - L;
        fo[m] = (Fs * ((double)kk / (double)L)) / (double)foeM;
    }
}

}  // namespace

extern "C" int64_t ocb_cpr_workspace_bytes(int64_t L, int nModes) {
    if (L <= 0 || nModes <= 0) return -1;
    const int64_t n = L * nModes;
    const int64_t nblk = (L + kScanBlock - 1) / kScanBlock;
    return n * 16 /*X*/ + n * 16 /*Z*/ + n * 4 /*idx*/ + n * 4 /*u*/ + nblk * nModes * 4 + 4096 * 8 + nModes * 16 + 1024;
}

extern "C" int ocb_cpr_bps_run(const void* x_dev, int x_dtype, int64_t L, int nModes, const void* constSymb, int M, int B,
                               int Nhalf, int runFOE, double Fs, int foeM, void* y_out, void* phase_out,
                               double* fo_host, void* workspace, int64_t workspace_bytes, void* stream) {
    OCB_REQUIRE(x_dev && constSymb && y_out && phase_out && workspace, "cpr_bps_run: NULL argument");
    OCB_REQUIRE(L > 0 && nModes > 0 && M > 0 && B > 0 && Nhalf >= 0, "cpr_bps_run: bad sizes");
    OCB_REQUIRE(workspace_bytes >= ocb_cpr_workspace_bytes(L, nModes), "cpr_bps_run: workspace too small");
    OCB_REQUIRE(!runFOE || (Fs > 0 && foeM >= 1 && foeM <= 16), "cpr_bps_run: bad FOE parameters");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = L * nModes;
    const int nblk = (int)((L + kScanBlock - 1) / kScanBlock);
    char* c = (char*)workspace;
    double2* X = (double2*)c; c += n * 16;
    double2* Z = (double2*)c; c += n * 16;
    int32_t* idx = (int32_t*)c; c += n * 4;
    int32_t* u = (int32_t*)c; c += n * 4;
    int32_t* bsum = (int32_t*)c; c += (int64_t)nblk * nModes * 4;
    c = (char*)(((uintptr_t)c + 255) & ~(uintptr_t)255);
    double* partials = (double*)c; c += 4096 * 8;
    int64_t* pos = (int64_t*)c; c += nModes * 8;
    double* fo_dev = (double*)c;

    if (ocb_cast_complex(x_dev, x_dtype, X, OCB_C128, n, stream)) return 1;
    const int g = grid_for(n, 256, 2);
    const int gp = g > 4096 ? 4096 : g;

    if (runFOE) {
        OCB_REQUIRE(L < (1ll << 31), "cpr_bps_run: L too large for the FOE transform");
        OCB_LAUNCH(k_foe_power, g, 256, 0, st, X, Z, L, nModes, foeM);
        cufftHandle plan;  // cached per (device, L, nModes): plan_cache.cuh
        OCB_CUFFT(fft_plan_cached(CUFFT_Z2Z, (int)L, nModes, st, &plan));
        OCB_CUFFT(cufftExecZ2Z(plan, (cufftDoubleComplex*)Z, (cufftDoubleComplex*)Z, CUFFT_FORWARD));
        k_foe_argmax<<<nModes, 1024, 0, st>>>(Z, L, pos);
        launch_counter()++;
        // f = fftshift(Fs * fftfreq(L)); fo = f[indFO] / M   (:358-359, :368) — evaluated on the device, the host
        // copy of fo is fetched only when the caller asked for it
        OCB_LAUNCH(k_foe_freq, 1, 32, 0, st, pos, fo_dev, L, nModes, Fs, foeM);
        OCB_LAUNCH(k_foe_apply, g, 256, 0, st, X, L, nModes, fo_dev, Fs);
        OCB_LAUNCH(k_power_partials, gp, 256, 0, st, X, n, partials);  // pnorm (:130)
        OCB_LAUNCH(k_pnorm_scale, g, 256, 0, st, X, n, partials, gp);
        if (fo_host) {
            OCB_CUDA(cudaMemcpyAsync(fo_host, fo_dev, nModes * sizeof(double), cudaMemcpyDeviceToHost, st));
            OCB_CUDA(cudaStreamSynchronize(st));
        }
    } else if (fo_host) {
        for (int m = 0; m < nModes; ++m) fo_host[m] = 0.0;
    }

    if (ocb_bps_run(X, L, nModes, constSymb, M, B, Nhalf, idx, phase_out, stream)) return 1;  // :138
    dim3 sgrid(nblk, nModes);
    OCB_LAUNCH(k_unwrap_scan1, sgrid, kScanBlock, 0, st, idx, u, bsum, L, nModes, B);
    OCB_LAUNCH(k_unwrap_scan2, nModes, 32, 0, st, bsum, nblk);
    OCB_LAUNCH(k_unwrap_apply, g, 256, 0, st, u, bsum, nblk, X, (double2*)y_out, (double*)phase_out, L, nModes, B);  // :154, :162
    OCB_LAUNCH(k_power_partials, gp, 256, 0, st, (const double2*)y_out, n, partials);
    OCB_LAUNCH(k_pnorm_scale, g, 256, 0, st, (double2*)y_out, n, partials, gp);
    return 0;
}

// pnorm (optic/dsp/core.py:702-717): x / sqrt(mean(|x|^2)) over the WHOLE array, in place, complex128.
// workspace: >= 4096 doubles of reduction scratch.
extern "C" int ocb_pnorm_run(void* x_dev, int64_t n, void* workspace, int64_t workspace_bytes, void* stream) {
    OCB_REQUIRE(x_dev && workspace && n > 0, "pnorm_run: bad argument");
    OCB_REQUIRE(workspace_bytes >= 4096 * 8, "pnorm_run: workspace too small (4096 doubles)");
    cudaStream_t st = (cudaStream_t)stream;
    double* partials = (double*)workspace;
    const int g = grid_for(n, 256, 2);
    const int gp = g < 1024 ? g : 1024;
    OCB_LAUNCH(k_power_partials, gp, 256, 0, st, (const double2*)x_dev, n, partials);
    OCB_LAUNCH(k_pnorm_scale, g, 256, 0, st, (double2*)x_dev, n, partials, gp);
    return 0;
}
