// RLS / DD-RLS stages of the adaptive equalizer (optic/dsp/equalization.py:576-644, 712-785).
//
// Per symbol and per INPUT mode N the reference updates an nTaps x nTaps inverse correlation matrix
//     u = conj(x_N[window]) ;  A = Sd u ;  B = u^H Sd ;  C = u^H A ;  Sd <- (Sd - A B / (lambda + C)) / lambda ;  Y = Sd u
// and then every output mode m does  H[m + N nModes, :] += err_m Y  with err = symbRef - out (rls) or the
// decision error (dd-rls).  Sd — and therefore the gain vector Y_N(s) — depends on the INPUT only, so the work
// splits into two kernels:
//   k_rls_gain : one warp per (stream, input mode): the O(nTaps^2) matrix recursion, Y_N(s) for every symbol of
//                the stage -> gain buffer [stream][N][s][32].  Lane i keeps row i of Sd in registers and mirrors it
//                in shared memory, where it reads column i: A (row dot), B (column dot) and the rank-1 update need
//                no cross-lane reduction; only C does.
//   k_mimo_eq_rls : one warp per (stream, output mode): the tap recurrence with the precomputed gain vectors.
// Sd starts from the identity in every stage call, like the reference (equalization.py:447-451 — the reference
// leaves Sd undefined for 'dd-rls'; the identity is used here for both).
// nTaps <= 32: one tap / one matrix row per lane, the row also in registers.  32 < nTaps <= 64 (RPL = 2, e.g. the 35 taps
// of examples/test_WDM_transmission.ipynb): lane i owns rows / taps i and i + 32 and the matrix lives in shared memory
// only (k_rls_gain2).  Included by rxdsp.cu inside its anonymous namespace.
#pragma once

__device__ __forceinline__ float2 cdiv(float2 a, float2 b) {  // a / b
    const float d = 1.0f / fmaf(b.x, b.x, b.y * b.y);
    return make_float2((a.x * b.x + a.y * b.y) * d, (a.y * b.x - a.x * b.y) * d);
}

template <int NM>
__global__ void __launch_bounds__(32)
k_rls_gain(const float2* __restrict__ X, float2* __restrict__ G, int64_t xStride, int64_t L, int nTaps, int SpS,
           float lambda) {
    // ONE copy of Sd: row `lane` in registers (for A, the rank-1 update and Y), mirrored in shared memory (row
    // stride 33, conflict-free both ways) so that lane i can read COLUMN i for B.  (Keeping separate row and
    // column copies is not an option: in complex64 the two drift apart and the recursion diverges.)
    __shared__ float2 Ss[32 * 33];
    __shared__ float2 su[32], sB[32];
    const int lane = threadIdx.x;
    const int stream = blockIdx.x / NM, N = blockIdx.x % NM;
    const float2* x = X + (int64_t)stream * xStride;
    float2* g = G + ((int64_t)stream * NM + N) * L * 32;
    const bool live = lane < nTaps;
    float2 R[32];  // row `lane` of Sd
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        R[j] = make_float2((j == lane && live) ? 1.f : 0.f, 0.f);
        Ss[lane * 33 + j] = R[j];
    }
    const float inv_lambda = 1.0f / lambda;
    __syncwarp();
    for (int64_t s = 0; s < L; ++s) {
        float2 xi = make_float2(0.f, 0.f);
        if (live) xi = x[(s * SpS + lane) * NM + N];
        const float2 ui = make_float2(xi.x, -xi.y);  // u = conj(x)  (:627)
        su[lane] = ui;
        __syncwarp();
        float2 A = make_float2(0.f, 0.f), B = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float2 uj = su[j];
            const float2 a = cmul(R[j], uj);                    // A_i = sum_j Sd[i][j] u_j        (:632)
            A.x += a.x; A.y += a.y;
            const float2 b = cmul_conj(Ss[j * 33 + lane], uj);  // B_i = sum_j conj(u_j) Sd[j][i]  (:633)
            B.x += b.x; B.y += b.y;
        }
        float2 c = cmul_conj(A, ui);                // conj(u_i) A_i
        c.x = warp_sum(c.x); c.y = warp_sum(c.y);   // C = u^H A                       (:634)
        const float2 den = make_float2(lambda + c.x, c.y);
        sB[lane] = B;
        __syncwarp();  // every lane has read its column of Ss; B is visible
        const float2 Ad = cdiv(A, den);
        float2 Y = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float2 r = cmul(Ad, sB[j]);       // (A B)[i][j] / (lambda + C)      (:635-637)
            R[j] = make_float2((R[j].x - r.x) * inv_lambda, (R[j].y - r.y) * inv_lambda);
            Ss[lane * 33 + j] = R[j];
            const float2 y = cmul(R[j], su[j]);     // Y = Sd_new u                     (:639)
            Y.x += y.x; Y.y += y.y;
        }
        g[s * 32 + lane] = Y;
        __syncwarp();  // Ss / su / sB are rewritten in the next trip
    }
}

// 32 < nTaps <= 64: two rows per lane (r = lane + 32 k), Sd in shared memory only (row stride 65: conflict-free row and
// column walks).  Same recursion, same order of operations per matrix entry as k_rls_gain.
template <int NM>
__global__ void __launch_bounds__(32)
k_rls_gain2(const float2* __restrict__ X, float2* __restrict__ G, int64_t xStride, int64_t L, int nTaps, int SpS,
            float lambda) {
    constexpr int T = 64, STR = 65;
    extern __shared__ float2 rls_smem[];
    float2* Ss = rls_smem;            // [T][STR]
    float2* su = Ss + T * STR;        // [T]
    float2* sB = su + T;              // [T]
    const int lane = threadIdx.x;
    const int stream = blockIdx.x / NM, N = blockIdx.x % NM;
    const float2* x = X + (int64_t)stream * xStride;
    float2* g = G + ((int64_t)stream * NM + N) * L * T;
    for (int k = 0; k < 2; ++k) {
        const int r = lane + 32 * k;
        for (int j = 0; j < T; ++j) Ss[r * STR + j] = make_float2((j == r && r < nTaps) ? 1.f : 0.f, 0.f);
    }
    const float inv_lambda = 1.0f / lambda;
    __syncwarp();
    for (int64_t s = 0; s < L; ++s) {
        float2 ui[2];
        for (int k = 0; k < 2; ++k) {
            const int r = lane + 32 * k;
            float2 xi = make_float2(0.f, 0.f);
            if (r < nTaps) xi = x[(s * SpS + r) * NM + N];
            ui[k] = make_float2(xi.x, -xi.y);  // u = conj(x)
            su[r] = ui[k];
        }
        __syncwarp();
        float2 A[2], B[2];
        float2 c = make_float2(0.f, 0.f);
        for (int k = 0; k < 2; ++k) {
            const int r = lane + 32 * k;
            A[k] = make_float2(0.f, 0.f);
            B[k] = make_float2(0.f, 0.f);
            for (int j = 0; j < T; ++j) {
                const float2 uj = su[j];
                const float2 a = cmul(Ss[r * STR + j], uj);
                A[k].x += a.x; A[k].y += a.y;
                const float2 b = cmul_conj(Ss[j * STR + r], uj);
                B[k].x += b.x; B[k].y += b.y;
            }
            const float2 ck = cmul_conj(A[k], ui[k]);
            c.x += ck.x; c.y += ck.y;
        }
        c.x = warp_sum(c.x); c.y = warp_sum(c.y);
        const float2 den = make_float2(lambda + c.x, c.y);
        sB[lane] = B[0];
        sB[lane + 32] = B[1];
        __syncwarp();  // every lane has read its columns of Ss; B is visible
        for (int k = 0; k < 2; ++k) {
            const int r = lane + 32 * k;
            const float2 Ad = cdiv(A[k], den);
            float2 Yv = make_float2(0.f, 0.f);
            for (int j = 0; j < T; ++j) {
                const float2 rr = cmul(Ad, sB[j]);
                const float2 o = Ss[r * STR + j];
                const float2 nv = make_float2((o.x - rr.x) * inv_lambda, (o.y - rr.y) * inv_lambda);
                Ss[r * STR + j] = nv;
                const float2 y = cmul(nv, su[j]);
                Yv.x += y.x; Yv.y += y.y;
            }
            g[s * T + r] = Yv;
        }
        __syncwarp();
    }
}

// One warp per (stream, output mode m); lane = tap.  DD: decision-directed error (dd-rls), else reference symbols.
template <int NM, bool DD, int RPL>
__global__ void __launch_bounds__(32 * NM)
k_mimo_eq_rls(const float2* __restrict__ X, const float2* __restrict__ REF, const float2* __restrict__ G,
              float2* __restrict__ Hg, float2* __restrict__ Y, float* __restrict__ ERR, float2* __restrict__ HIT,
              int64_t xStride, int64_t refStride, int64_t yStride, int64_t errStride, int64_t errModeStride, int64_t L,
              int nTaps, int SpS, const float2* __restrict__ constSymb, int M) {
    const int lane = threadIdx.x & 31, m = threadIdx.x >> 5;
    const int stream = blockIdx.x;
    const float2* x = X + (int64_t)stream * xStride;
    const float2* ref = REF ? REF + (int64_t)stream * refStride : nullptr;
    constexpr int GS = 32 * RPL;  // gain-vector stride per symbol
    const float2* g = G + (int64_t)stream * NM * L * GS;
    float2* Hs = Hg + (int64_t)stream * NM * NM * nTaps;
    float2* y = Y + (int64_t)stream * yStride;
    float* err = ERR + (int64_t)stream * errStride + (int64_t)m * errModeStride;
    float2* hit = HIT ? HIT + (int64_t)stream * L * NM * NM * nTaps : nullptr;
    float2 H[NM][RPL];
#pragma unroll
    for (int n = 0; n < NM; ++n)
#pragma unroll
        for (int k = 0; k < RPL; ++k) {
            const int t = lane + 32 * k;
            H[n][k] = t < nTaps ? Hs[(m + n * NM) * nTaps + t] : make_float2(0.f, 0.f);
        }
    for (int64_t s = 0; s < L; ++s) {
        float2 o = make_float2(0.f, 0.f);
        float2 gy[NM][RPL];
#pragma unroll
        for (int n = 0; n < NM; ++n)
#pragma unroll
            for (int k = 0; k < RPL; ++k) {
                const int t = lane + 32 * k;
                const float2 w = t < nTaps ? x[(s * SpS + t) * NM + n] : make_float2(0.f, 0.f);
                gy[n][k] = g[((int64_t)n * L + s) * GS + t];
                const float2 pr = cmul(H[n][k], w);  // equalization.py:464-468
                o.x += pr.x; o.y += pr.y;
            }
        o.x = warp_sum(o.x); o.y = warp_sum(o.y);
        float2 target;
        if constexpr (DD) {  // nearest constellation point, first index on ties (:751-753)
            float best = 3.4e38f;
            int bi = 0x7fffffff;
            for (int c = lane; c < M; c += 32) {
                const float2 sc = __ldg(constSymb + c);
                const float dd = cabs2(make_float2(o.x - sc.x, o.y - sc.y));
                if (dd < best) { best = dd; bi = c; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, off);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            target = __ldg(constSymb + bi);
        } else {
            target = ref[s * NM + m];
        }
        const float2 e = make_float2(target.x - o.x, target.y - o.y);  // :614 / :754
        if (lane == 0) {
            y[s * NM + m] = o;  // :473
            err[s] = cabs2(e);
        }
#pragma unroll
        for (int n = 0; n < NM; ++n)
#pragma unroll
            for (int k = 0; k < RPL; ++k) {
                const float2 u = cmul(e, gy[n][k]);  // H[m + n nModes, :] += err_m Y_n   (:641)
                H[n][k].x += u.x; H[n][k].y += u.y;
            }
        if (hit) {
#pragma unroll
            for (int n = 0; n < NM; ++n)
#pragma unroll
                for (int k = 0; k < RPL; ++k)
                    if (lane + 32 * k < nTaps) hit[(s * NM * NM + m + n * NM) * nTaps + lane + 32 * k] = H[n][k];
        }
    }
#pragma unroll
    for (int n = 0; n < NM; ++n)
#pragma unroll
        for (int k = 0; k < RPL; ++k)
            if (lane + 32 * k < nTaps) Hs[(m + n * NM) * nTaps + lane + 32 * k] = H[n][k];
}

template <int NM>
int launch_rls(cudaStream_t st, const float2* X, const float2* REF, float2* G, float2* H, float2* Y, float* ERR,
               float2* HIT, int nStreams, int64_t xs, int64_t rs, int64_t ys, int64_t es, int64_t ems, int64_t L,
               int nTaps, int SpS, bool dd, float lambda, const float2* cs, int M) {
    if (nTaps <= 32) {
        OCB_LAUNCH((k_rls_gain<NM>), nStreams * NM, 32, 0, st, X, G, xs, L, nTaps, SpS, lambda);
        if (dd) OCB_LAUNCH((k_mimo_eq_rls<NM, true, 1>), nStreams, 32 * NM, 0, st, X, REF, G, H, Y, ERR, HIT, xs, rs, ys, es, ems, L, nTaps, SpS, cs, M);
        else OCB_LAUNCH((k_mimo_eq_rls<NM, false, 1>), nStreams, 32 * NM, 0, st, X, REF, G, H, Y, ERR, HIT, xs, rs, ys, es, ems, L, nTaps, SpS, cs, M);
        return 0;
    }
    const size_t smem = (size_t)(64 * 65 + 128) * sizeof(float2);  // 34 KB: below the 48 KB default limit
    OCB_LAUNCH((k_rls_gain2<NM>), nStreams * NM, 32, smem, st, X, G, xs, L, nTaps, SpS, lambda);
    if (dd) OCB_LAUNCH((k_mimo_eq_rls<NM, true, 2>), nStreams, 32 * NM, 0, st, X, REF, G, H, Y, ERR, HIT, xs, rs, ys, es, ems, L, nTaps, SpS, cs, M);
    else OCB_LAUNCH((k_mimo_eq_rls<NM, false, 2>), nStreams, 32 * NM, 0, st, X, REF, G, H, Y, ERR, HIT, xs, rs, ys, es, ems, L, nTaps, SpS, cs, M);
    return 0;
}
