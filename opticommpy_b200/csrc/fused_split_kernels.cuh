// Pair-split (16 samples per thread) versions of the fused Manakov kernels for the N = 2^20 geometry
// (N1 = N2 = 1024).  Same mathematics and data flow as k_time / k_freq in fused_kernels.cuh; the only
// differences are the thread decomposition (fft_split.cuh) and, as a consequence, the order inside a
// W row:  position p = slot*64 + g  holds  k1 = x(g) + 32*(2*brev16(slot) + h(g)),
//         x(g) = (g>>5)*16 + (g&15),  h(g) = (g>>4)&1.
#pragma once
#include "fft_split.cuh"
#include "fused_kernels.cuh"

namespace ocb {

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// 128-thread CTA = the x and y tasks of one time row; 64 threads (2 warps) per task.
#ifndef OCB_TIME_S_MINB
#define OCB_TIME_S_MINB 7
#endif
template <int MODE>
__global__ void __launch_bounds__(128, OCB_TIME_S_MINB)
k_time_s(const TimeArgs A) {
    using namespace fft;
    constexpr int N1 = 1024, STR = 33, GBUF = 32 * STR;
    if constexpr (MODE == TM_ITER) {
        if (A.ext.mail && *reinterpret_cast<volatile long long*>(A.ext.converged_step) == A.ext.step_id) return;
    }
    __shared__ float xbuf[2 * 2 * GBUF];
    __shared__ float pbuf[(MODE == TM_FIRST || MODE == TM_ITER) ? 2 * N1 : 1];

    const int tid = threadIdx.x, task = tid >> 6, g = tid & 63;
    const int x = ((g >> 5) << 4) | (g & 15), h = (g >> 4) & 1;
    const int row = blockIdx.x, pol = task;
    float* xr = xbuf + task * 2 * GBUF;
    float* xi = xr + GBUF;
    const int64_t base = (int64_t)pol * A.N + (int64_t)row * N1;
    auto gsync = [task] { named_bar_sync(1 + task, 64); };
    const float2* tw = A.tw;
    const float2 wV = __ldg(A.tabV + (int64_t)row * 32 + x);
    const float2* Urow = A.tabU + (int64_t)row * 32;

    if constexpr (MODE == TM_FIRST || MODE == TM_ITER) {  // HBM->L2 prefetch of the pointwise streams
        constexpr int LINES = N1 * 8 / 128;
        if (g < LINES) {
            prefetch_l2(reinterpret_cast<const char*>(A.aux0 + base) + g * 128);
            if constexpr (MODE == TM_ITER) prefetch_l2(reinterpret_cast<const char*>(A.ehd + base) + g * 128);
        }
        if constexpr (MODE == TM_ITER)
            if (pol == 0 && g < LINES / 2) prefetch_l2(reinterpret_cast<const char*>(A.pch + (int64_t)row * N1) + g * 128);
    }

    float2 v[16];
    if constexpr (MODE == TM_FWD) {
        const float2* src = A.in + base;
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = ld_stream(src + 32 * (16 * h + j) + x);
    } else {
        const float2* src = A.in + base;
#pragma unroll
        for (int s = 0; s < 16; ++s) v[s] = ld_stream(src + s * 64 + g);
        static_for<0, 16>([&](auto kk) {
            constexpr int K = decltype(kk)::value, SLOT = brev<16>(K);
            const float2 w = cmul(wV, __ldg(Urow + 2 * K + h));
            v[SLOT] = cmul_conj(v[SLOT], w);
        });
        coop1024s_inverse<1, 1, 16>(v, xr, xi, tw, x, h, 0, gsync);  // v[m] = sample n1 = 32*(16h+m) + x
        gsync();
    }

    float s_num = 0.f, s_den = 0.f, s_max = 0.f;
    if constexpr (MODE == TM_FIRST || MODE == TM_ITER) {
        float* pown = pbuf + task * N1;
        const float* poth = pbuf + (task ^ 1) * N1;
        if constexpr (MODE == TM_FIRST) {
            float2* ehd_out = A.aux1 + base;
            const float2* ech = A.aux0 + base;
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int n1 = 32 * (16 * h + m) + x;
                st_stream(ehd_out + n1, v[m]);
                pown[n1] = cabs2(ld_stream(ech + n1));
            }
        } else {
            const float2* ec = A.aux0 + base;
            float2* ec_new = A.aux1 + base;
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int n1 = 32 * (16 * h + m) + x;
                const float2 e = ld_stream(ec + n1);
                s_num += cabs2(make_float2(v[m].x - e.x, v[m].y - e.y));
                s_den += cabs2(e);
                st_stream(ec_new + n1, v[m]);
                pown[n1] = cabs2(v[m]);
            }
        }
        __syncthreads();
        float* pch = A.pch + (int64_t)row * N1;
        const float2* ehd = (MODE == TM_ITER) ? A.ehd + base : nullptr;
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int n1 = 32 * (16 * h + m) + x;
            const float P = pown[n1] + poth[n1];
            float ph;
            if constexpr (MODE == TM_FIRST) {
                if (pol == 0) st_stream(pch + n1, P);
                ph = A.cphi * P;
            } else {
                s_max = fmaxf(s_max, P);
                ph = A.cphi * (ld_stream(pch + n1) + P);
                v[m] = ld_stream(ehd + n1);
            }
            v[m] = cmul(v[m], phase_rot(ph));
        }
    }

    coop1024s_forward<1, 1, 16>(v, xr, xi, tw, x, h, 0, gsync);
    {
        float2* dst = A.out + base;
        static_for<0, 16>([&](auto kk) {
            constexpr int K = decltype(kk)::value, SLOT = brev<16>(K);
            const float2 w = cmul(wV, __ldg(Urow + 2 * K + h));
            st_stream(dst + SLOT * 64 + g, cmul(v[SLOT], w));
        });
    }
    if constexpr (MODE == TM_ITER) {
        if (pol == 1) s_max = 0.f;
        block_reduce3_finalize(s_num, s_den, s_max, A.partials, A.sums, A.ticket, &A.ext);
    }
}

// 64*C threads = one tile of C adjacent W positions, all 1024 rows of one polarisation.
// tid = (x>>1)*4C + (x&1)*2C + h*C + c  (the pair partner is lane ^ C).
template <int C>
__global__ void __launch_bounds__(64 * C, (64 * C <= 512) ? 2 : 1)
k_freq_s(float2* __restrict__ W, const float2* __restrict__ LP, const float2* __restrict__ tw, int N1,
         const long long* __restrict__ converged_step, long long step_id) {
    using namespace fft;
    if (converged_step && *reinterpret_cast<const volatile long long*>(converged_step) == step_id) return;
    constexpr int PAD = 16, STR = 32 * C + PAD;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* xr = reinterpret_cast<float*>(smem_raw);
    float* xi = xr + 32 * STR;
    const int tid = threadIdx.x, c = tid % C, h = (tid / C) & 1;
    const int x = ((tid / (4 * C)) << 1) | ((tid / (2 * C)) & 1);
    const int tiles_per_pol = N1 / C;
    const int pol = blockIdx.x / tiles_per_pol, tile = blockIdx.x % tiles_per_pol;
    float2* base = W + ((int64_t)pol * 1024) * N1 + tile * C + c;
    const float2* lp = LP + (int64_t)tile * 16 * (64 * C) + tid;
    auto bsync = [] { __syncthreads(); };

    float2 v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = ld_stream(base + (int64_t)(32 * (16 * h + j) + x) * N1);
    coop1024s_forward<C, PAD, C>(v, xr, xi, tw, x, h, c, bsync);
#pragma unroll
    for (int s = 0; s < 16; ++s) v[s] = cmul(v[s], ld_stream(lp + s * (64 * C)));
    __syncthreads();
    coop1024s_inverse<C, PAD, C>(v, xr, xi, tw, x, h, c, bsync);
#pragma unroll
    for (int j = 0; j < 16; ++j) st_stream(base + (int64_t)(32 * (16 * h + j) + x) * N1, v[j]);
}

// Operator table for the split layout (N1 = N2 = 1024).  Entry i = (tile*16 + slot2)*(64C) + tid2.
__global__ void k_tab_linop_perm_s(float2* __restrict__ LP, int C, int64_t N, double a, double b, double Fs,
                                   double h, double scale) {
    const double two_pi = 6.283185307179586476925286766559;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const int tid2 = (int)(i % (64 * C));
        const int slot2 = (int)((i / (64 * C)) % 16);
        const int64_t tile = i / ((int64_t)64 * C * 16);
        const int c = tid2 % C, h2 = (tid2 / C) & 1;
        const int x2 = ((tid2 / (4 * C)) << 1) | ((tid2 / (2 * C)) & 1);
        const int k2 = x2 + 32 * (2 * brev_rt(slot2, 4) + h2);
        const int p = (int)(tile * C + c);
        const int s1 = p / 64, g1 = p % 64;
        const int x1 = ((g1 >> 5) << 4) | (g1 & 15), h1 = (g1 >> 4) & 1;
        const int k1 = x1 + 32 * (2 * brev_rt(s1, 4) + h1);
        const int64_t k = k1 + (int64_t)1024 * k2;
        const int64_t kk = (k <= (N - 1) / 2) ? k : k - N;
        const double w = two_pi * Fs * ((double)kk / (double)N);
        const double amp = scale * exp(a * h);
        double sn, cs;
        sincos(b * (w * w) * h, &sn, &cs);
        LP[i] = make_float2((float)(amp * cs), (float)(amp * sn));
    }
}

}  // namespace ocb
