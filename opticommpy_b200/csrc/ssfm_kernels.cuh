// Elementwise / reduction kernels of the split-step propagators (cuFFT-driven engine).
// All fields are planar rows[R][N] complex64 (R = 2K: x rows then y rows).
#pragma once
#include "common.cuh"

namespace ocb {

// ------------------------------------------------------------------------------------------
// Linear-operator table:  T[k] = scale * exp( (a + j b ω_k²) · h )^{pw}
//   ω_k = 2π Fs fftfreq(N)[k]                      (channels.py:203, 360)
//   a = ∓α/2, b = ±β2/2                            (channels.py:368 ; equalization.py:1077)
//   h = hz/2                                       (channels.py:213, 406)
// The phase is evaluated in float64 (it reaches 1e2..1e3 rad at the band edge) and only the
// final (cos, sin) pair is rounded to float32.
// ------------------------------------------------------------------------------------------
__global__ void k_linop_table(float2* __restrict__ T, int64_t N, double a, double b, double Fs,
                              double h, double scale) {
    const double two_pi = 6.283185307179586476925286766559;
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < N;
         k += (int64_t)gridDim.x * blockDim.x) {
        int64_t kk = (k <= (N - 1) / 2) ? k : k - N;  // numpy.fft.fftfreq bin order
        double w = two_pi * Fs * ((double)kk / (double)N);
        double ph = b * (w * w) * h;
        double amp = scale * exp(a * h);
        double s, c;
        sincos(ph, &s, &c);
        T[k] = make_float2((float)(amp * c), (float)(amp * s));
    }
}

// F[r][k] *= T[k]   (frequency-domain multiply; 1/N of the unnormalised inverse cuFFT is folded in T)
template <int VEC>
__global__ void k_mul_table(float2* __restrict__ F, const float2* __restrict__ T, int64_t N, int R) {
    const int64_t per_row = N / VEC;
    const int64_t total = per_row * R;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t k = (i % per_row) * VEC;
        int64_t off = (i / per_row) * N + k;
        if (VEC == 2) {
            float4 f = *reinterpret_cast<const float4*>(F + off);
            float4 t = ldg4(T + k);
            float2 a = cmul(make_float2(f.x, f.y), make_float2(t.x, t.y));
            float2 b = cmul(make_float2(f.z, f.w), make_float2(t.z, t.w));
            stg4(F + off, make_float4(a.x, a.y, b.x, b.y));
        } else {
            F[off] = cmul(F[off], __ldg(T + k));
        }
    }
}

// ------------------------------------------------------------------------------------------
// Block-level reduction of (sum, sum, max) with a deterministic "last block finalises" tail.
// partials: [gridDim.x][3] doubles ; out: 3 doubles ; ticket: 1 uint (self-resetting).
// ------------------------------------------------------------------------------------------
// Host mailbox (mapped pinned memory) + device-side convergence flag used by the step loop to avoid
// a stream synchronisation per fixed-point iteration: the finalising block publishes the sums and
// the decision `lim < tol` (channels.py:429), speculatively enqueued launches of the same step read
// the flag and exit.
struct Mail {
    double sums[3];
    unsigned long long seq;  // written last, after a system-scope fence
    int converged;
    int pad;
};
constexpr int kMailSlots = 8;  // ring: the device runs at most one speculative iteration ahead of the host
struct FinalizeExt {
    Mail* mail;                 // ring of kMailSlots entries, slot = seq % kMailSlots (nullptr: plain reduction)
    long long* converged_step;  // device flag
    long long step_id;
    unsigned long long seq;
    double tol;
    long long* final_step;      // TM_ITERF only: set to step_id when the predicted-last iteration did converge
};

// The reduction is split in two so that a kernel can POST its partial sums as soon as they are final and
// only later — after the rest of its work — ask whether it was the last block: the fence + atomic round
// trip of the ticket then overlaps that work instead of extending every block's lifetime.
//   post : block-level sums -> partials[blockIdx.x]; thread 0 takes a ticket (its value is meaningful in thread 0)
//   final: the block that drew the last ticket adds up all partials in a fixed order, publishes them and,
//          with `ext`, the decision lim < tol (channels.py:517-519, 429) + the host mail
template <typename TS>  // float: one row per thread ; double: persistent kernels that sum several rows per thread
__device__ __forceinline__ unsigned block_reduce3_post(TS s0, TS s1, float m, double* __restrict__ partials,
                                                       unsigned* __restrict__ ticket) {
    __shared__ double sh0[32], sh1[32], sh2[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    // warp level in the caller's type (float: <= 32 addends), block level in double
    TS w0 = warp_sum(s0), w1 = warp_sum(s1);
    float wm = warp_max(m);
    if (lane == 0) { sh0[wid] = (double)w0; sh1[wid] = (double)w1; sh2[wid] = (double)wm; }
    __syncthreads();
    unsigned t = 0;
    if (wid == 0) {
        double a = lane < nw ? sh0[lane] : 0.0, b = lane < nw ? sh1[lane] : 0.0;
        double c = lane < nw ? sh2[lane] : 0.0;
        a = warp_sum(a); b = warp_sum(b);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c = fmax(c, __shfl_xor_sync(0xffffffffu, c, o));
        if (lane == 0) {
            partials[3 * blockIdx.x + 0] = a;
            partials[3 * blockIdx.x + 1] = b;
            partials[3 * blockIdx.x + 2] = c;
            __threadfence();
            t = atomicAdd(ticket, 1u);
        }
    }
    return t;
}
__device__ __forceinline__ void block_reduce3_final(unsigned my_ticket, double* __restrict__ partials,
                                                    double* __restrict__ out, unsigned* __restrict__ ticket,
                                                    const FinalizeExt* ext = nullptr) {
    __shared__ double sh0[32], sh1[32], sh2[32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (threadIdx.x == 0) is_last = (my_ticket == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    // fixed-order strided accumulation -> deterministic for a fixed grid
    double a = 0.0, b = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
        a += __ldcg(partials + 3 * i + 0);
        b += __ldcg(partials + 3 * i + 1);
        c = fmax(c, __ldcg(partials + 3 * i + 2));
    }
    a = warp_sum(a); b = warp_sum(b);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c = fmax(c, __shfl_xor_sync(0xffffffffu, c, o));
    if (lane == 0) { sh0[wid] = a; sh1[wid] = b; sh2[wid] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0, tc = 0.0;
        for (int w = 0; w < nw; ++w) { ta += sh0[w]; tb += sh1[w]; tc = fmax(tc, sh2[w]); }
        out[0] = ta; out[1] = tb; out[2] = tc;
        *ticket = 0u;  // ready for the next launch
        if (ext && ext->mail) {
            const int conv = (sqrt(ta) / sqrt(tb) < ext->tol) ? 1 : 0;  // channels.py:517-519, 429
            if (conv) {
                *ext->converged_step = ext->step_id;
                if (ext->final_step) *ext->final_step = ext->step_id;
            }
            volatile Mail* mb = ext->mail + (ext->seq % kMailSlots);
            mb->sums[0] = ta; mb->sums[1] = tb; mb->sums[2] = tc;
            mb->converged = conv;
            __threadfence_system();
            mb->seq = ext->seq;
        }
    }
}
template <typename TS>
__device__ __forceinline__ void block_reduce3_finalize(TS s0, TS s1, float m, double* __restrict__ partials,
                                                       double* __restrict__ out, unsigned* __restrict__ ticket,
                                                       const FinalizeExt* ext = nullptr) {
    const unsigned t = block_reduce3_post(s0, s1, m, partials, ticket);
    block_reduce3_final(t, partials, out, ticket, ext);
}

// phase rotation exp(j ph): short polynomial for the small per-step phases (|ph| < 0.5 rad,
// truncation error < 1e-9); the rare large phase takes an out-of-line libm call so that the 32
// unrolled call sites stay small (instruction-cache footprint).
__device__ __noinline__ float2 phase_rot_slow(float ph) {
    float s, c;
    sincosf(ph, &s, &c);
    return make_float2(c, s);
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ float2 phase_rot(float ph) {
    float s, c;
    if (fabsf(ph) >= 0.5f) return phase_rot_slow(ph);
    {
        const float x2 = ph * ph;
        s = ph * fmaf(x2, fmaf(x2, fmaf(x2, -1.9841270e-4f, 8.3333333e-3f), -1.6666667e-1f), 1.0f);
        c = fmaf(x2, fmaf(x2, fmaf(x2, fmaf(x2, 2.4801587e-5f, -1.3888889e-3f), 4.1666667e-2f), -0.5f), 1.0f);
    }
    return make_float2(c, s);
}

// ------------------------------------------------------------------------------------------
// Fused Manakov nonlinear pass (the kernel the HBM-roofline target is quoted on).
//   FIRST = true  : start of a step.  Ec is the step-start field, so φ = (8/9)γ·P
//                   (channels.py:388-390 with Ex_conv == Ech_x) ; Pch is written.
//                   Reads Ehd(16)+Ec(16), writes out(16)+Pch(4) bytes per 2-pol sample.
//   FIRST = false : after a second half-step.  Accumulates the convergence sums
//                   Σ|Efd-Ec|², Σ|Ec|² (channels.py:517-519) and max|Efd|² (for the next
//                   adaptive step), forms φ = (8/9)γ(Pch+|Efd_x|²+|Efd_y|²)/2 (channels.py:436,493)
//                   and writes the next rotated field out = Ehd·exp(j·dir·φ·hz) (channels.py:414-417).
//                   Reads Efd(16)+Ec(16)+Ehd(16)+Pch(4), writes out(16): 68 B per 2-pol sample.
// cphi = dir * hz * (8/9)γ  (FIRST)   or   dir * hz * (8/9)γ / 2  (!FIRST)
// ------------------------------------------------------------------------------------------
template <bool FIRST, int VEC>
__global__ void __launch_bounds__(256, 3)
k_manakov_nl(const float2* __restrict__ Ehd, const float2* __restrict__ Efd,
             const float2* __restrict__ Ec, float* __restrict__ Pch, float2* __restrict__ out,
             int64_t N, int K, float cphi, double* __restrict__ partials,
             double* __restrict__ sums, unsigned* __restrict__ ticket) {
    const int64_t per_row = N / VEC;
    const int64_t total = per_row * K;
    const int64_t ystride = (int64_t)K * N;  // y row of pair p is row K+p
    float s_num = 0.f, s_den = 0.f, s_max = 0.f;

    // Software pipeline: the loads of item i+stride are issued BEFORE the arithmetic and the stores of
    // item i, so every thread keeps one full item (7 x 16 B) in flight while it computes and HBM never
    // idles between grid-stride trips.  The grid is an exact multiple of the SM count (host side).
    struct Item {
        float2 hx[VEC], hy[VEC], cx[VEC], cy[VEC], fx[VEC], fy[VEC];
        float pc[VEC];
        int64_t off;
    };
    auto load = [&](Item& it, int64_t i) {
        it.off = (i / per_row) * N + (i % per_row) * VEC;  // pair row p, sample n
        if (VEC == 2) {
            float4 t;
            t = ldg4(Ehd + it.off);           it.hx[0] = make_float2(t.x, t.y); it.hx[1] = make_float2(t.z, t.w);
            t = ldg4(Ehd + it.off + ystride); it.hy[0] = make_float2(t.x, t.y); it.hy[1] = make_float2(t.z, t.w);
            t = ldg4(Ec + it.off);            it.cx[0] = make_float2(t.x, t.y); it.cx[1] = make_float2(t.z, t.w);
            t = ldg4(Ec + it.off + ystride);  it.cy[0] = make_float2(t.x, t.y); it.cy[1] = make_float2(t.z, t.w);
            if (!FIRST) {
                t = ldg4(Efd + it.off);           it.fx[0] = make_float2(t.x, t.y); it.fx[1] = make_float2(t.z, t.w);
                t = ldg4(Efd + it.off + ystride); it.fy[0] = make_float2(t.x, t.y); it.fy[1] = make_float2(t.z, t.w);
                float2 p2 = __ldg(reinterpret_cast<const float2*>(Pch + it.off));
                it.pc[0] = p2.x; it.pc[1] = p2.y;
            }
        } else {
            it.hx[0] = __ldg(Ehd + it.off); it.hy[0] = __ldg(Ehd + it.off + ystride);
            it.cx[0] = __ldg(Ec + it.off);  it.cy[0] = __ldg(Ec + it.off + ystride);
            if (!FIRST) {
                it.fx[0] = __ldg(Efd + it.off); it.fy[0] = __ldg(Efd + it.off + ystride);
                it.pc[0] = __ldg(Pch + it.off);
            }
        }
    };
    auto finish = [&](Item& it) {
        float2 ox[VEC], oy[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            float ph;
            if (FIRST) {
                float P = cabs2(it.cx[v]) + cabs2(it.cy[v]);
                it.pc[v] = P;
                ph = cphi * P;
            } else {
                float2 dx = make_float2(it.fx[v].x - it.cx[v].x, it.fx[v].y - it.cx[v].y);
                float2 dy = make_float2(it.fy[v].x - it.cy[v].x, it.fy[v].y - it.cy[v].y);
                s_num += cabs2(dx) + cabs2(dy);
                s_den += cabs2(it.cx[v]) + cabs2(it.cy[v]);
                float Pf = cabs2(it.fx[v]) + cabs2(it.fy[v]);
                s_max = fmaxf(s_max, Pf);
                ph = cphi * (it.pc[v] + Pf);
            }
            const float2 rot = phase_rot(ph);
            ox[v] = cmul(it.hx[v], rot);
            oy[v] = cmul(it.hy[v], rot);
        }
        if (VEC == 2) {
            stg4(out + it.off, make_float4(ox[0].x, ox[0].y, ox[1].x, ox[1].y));
            stg4(out + it.off + ystride, make_float4(oy[0].x, oy[0].y, oy[1].x, oy[1].y));
            if (FIRST) *reinterpret_cast<float2*>(Pch + it.off) = make_float2(it.pc[0], it.pc[1]);
        } else {
            out[it.off] = ox[0];
            out[it.off + ystride] = oy[0];
            if (FIRST) Pch[it.off] = it.pc[0];
        }
    };
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    Item A, B;
    if (i < total) load(A, i);
    while (i < total) {  // ping-pong: A is loaded; fetch B, finish A; then fetch A, finish B
        const int64_t j = i + stride;
        if (j < total) load(B, j);
        finish(A);
        if (j >= total) break;
        const int64_t k = j + stride;
        if (k < total) load(A, k);
        finish(B);
        i = k;
    }
    if (!FIRST) block_reduce3_finalize(s_num, s_den, s_max, partials, sums, ticket);
}

// max over samples of |Ex|²+|Ey|² (adaptive step at span start; channels.py:394) and Σ power
template <int VEC>
__global__ void __launch_bounds__(256)
k_power_stats(const float2* __restrict__ E, int64_t N, int K, double* __restrict__ partials,
              double* __restrict__ sums, unsigned* __restrict__ ticket) {
    const int64_t per_row = N / VEC;
    const int64_t total = per_row * K;
    const int64_t ystride = (int64_t)K * N;
    float s_pow = 0.f, s_max = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t off = (i / per_row) * N + (i % per_row) * VEC;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            float P = cabs2(__ldg(E + off + v)) + cabs2(__ldg(E + off + ystride + v));
            s_pow += P;
            s_max = fmaxf(s_max, P);
        }
    }
    block_reduce3_finalize(0.f, s_pow, s_max, partials, sums, ticket);
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator + Box-Muller (EDFA ASE noise, devices.py:722-726).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}
__device__ __forceinline__ float2 gauss_pair(uint32_t a, uint32_t b) {
    float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
    float u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    return make_float2(r * c, r * s);
}

// E[r][n] = E[r][n]*g + noise     (gain-only when sigma == 0 and noise == nullptr)
//   injected: noise[(r % noise_rows)][n]  — same realisation for x and y rows, every span
//   philox  : CN(0, 2σ²) with counter (r*N + n, stream_id), n = NATURAL sample index: when the field is kept in
//             the fused engine's transposed layout (position N1*n2 + n1 holds sample N2*n1 + n2; tN1 > 0) the
//             counter is mapped back, so a seed gives the same realisation whichever engine runs the span
__global__ void k_amp(float2* __restrict__ E, int R, int64_t N, float g, float sigma,
                      const float2* __restrict__ noise, int noise_rows, uint64_t seed,
                      uint64_t stream_id, int tN1, int tN2) {
    const int64_t total = (int64_t)R * N;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        float2 e = E[i];
        e.x *= g; e.y *= g;
        if (noise) {
            int64_t r = i / N, n = i % N;
            float2 w = __ldg(noise + (r % noise_rows) * N + n);
            e.x += w.x; e.y += w.y;
        } else if (sigma > 0.f) {
            int64_t ci = i;
            if (tN1 > 0) {
                const int64_t pos = i % N;
                ci = (i - pos) + (int64_t)tN2 * (pos % tN1) + pos / tN1;
            }
            uint4 c = make_uint4((uint32_t)ci, (uint32_t)(ci >> 32), (uint32_t)stream_id,
                                 (uint32_t)(stream_id >> 32));
            uint4 rnd = philox4x32_10(c, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
            float2 w = gauss_pair(rnd.x, rnd.y);
            e.x = fmaf(sigma, w.x, e.x); e.y = fmaf(sigma, w.y, e.y);
        }
        E[i] = e;
    }
}

// scalar NLSE nonlinear step: E *= exp(j c |E|²), c = γ·hz   (channels.py:225)
template <int VEC>
__global__ void k_nlse_phase(float2* __restrict__ E, int64_t total, float c) {
    for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * VEC; i < total;
         i += (int64_t)gridDim.x * blockDim.x * VEC) {
        float2 e[VEC];
        if (VEC == 2) {
            float4 t = *reinterpret_cast<const float4*>(E + i);
            e[0] = make_float2(t.x, t.y); e[1] = make_float2(t.z, t.w);
        } else e[0] = E[i];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            float s, cs;
            sincosf(c * cabs2(e[v]), &s, &cs);
            e[v] = cmul(e[v], make_float2(cs, s));
        }
        if (VEC == 2) stg4(E + i, make_float4(e[0].x, e[0].y, e[1].x, e[1].y));
        else E[i] = e[0];
    }
}

// (N, C) interleaved columns (complex64 | complex128)  ->  planar rows[C][N] complex64
template <typename TIN>
__global__ void k_pack(const TIN* __restrict__ src, float2* __restrict__ rows, int64_t N, int C,
                       int pairs) {
    const int64_t total = N * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t n = i / C;
        int c = (int)(i % C);
        int r = pairs ? ((c & 1) * (C / 2) + (c >> 1)) : c;
        TIN v = src[i];
        rows[(int64_t)r * N + n] = make_float2((float)v.x, (float)v.y);
    }
}
template <typename TOUT>
__global__ void k_unpack(const float2* __restrict__ rows, TOUT* __restrict__ dst, int64_t N, int C,
                         int pairs) {
    const int64_t total = N * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t n = i / C;
        int c = (int)(i % C);
        int r = pairs ? ((c & 1) * (C / 2) + (c >> 1)) : c;
        float2 v = rows[(int64_t)r * N + n];
        TOUT o; o.x = v.x; o.y = v.y;
        dst[i] = o;
    }
}

}  // namespace ocb
