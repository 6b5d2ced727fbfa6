// Host-side driver of the split-step propagators (plan lifecycle + step loops) behind the
// C-ABI in include/opticomm_b200.h.  cuFFT provides the length-N transforms; every other
// operation of a step is one of the fused kernels in ssfm_kernels.cuh.
//
// Reference control flow restated here (not code): optic/models/channels.py:380-456
// (manakovSSF), optic/dsp/equalization.py:1088-1161 (manakovDBP), channels.py:215-238 (ssfm).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/opticomm_b200.h"
#include "ssfm_kernels.cuh"
#include "fused_kernels.cuh"
#include "fused_time_bulk.cuh"
#include "fused_freq_tma.cuh"

using namespace ocb;

namespace ocb {
std::string& last_error() {
    static thread_local std::string e;
    return e;
}
int fail(const char* what, const char* file, int line) {
    char b[768];
    snprintf(b, sizeof b, "%s (%s:%d)", what, file, line);
    last_error() = b;
    return 1;
}
int64_t& launch_counter() {
    static thread_local int64_t c = 0;
    return c;
}
}  // namespace ocb

struct ocb_ssfm_plan {
    int64_t N = 0;
    int rows = 0;
    cufftHandle fft = 0;
    bool fft_ok = false;
    size_t fft_work = 0;
    // workspace partition (device)
    void* ws = nullptr;
    int64_t ws_bytes = 0;
    float2 *Ehd = nullptr, *A = nullptr, *B = nullptr, *G = nullptr, *T1 = nullptr, *T2 = nullptr;
    float* Pch = nullptr;
    double* partials = nullptr;  // [max_blocks][3]
    double* sums = nullptr;      // 3 doubles
    unsigned* ticket = nullptr;
    void* fft_area = nullptr;
    // pinned host mailbox for the convergence scalars
    double* h_sums = nullptr;
    // zero-copy mailbox (mapped pinned memory) + device flag for the sync-free fixed-point loop
    Mail* h_mail = nullptr;   // host view
    Mail* d_mail = nullptr;   // device view of the same memory
    long long* conv_flag = nullptr;  // device: id of the last step whose loop converged
    long long* final_flag = nullptr; // device: id of the last step whose predicted-last iteration (TM_ITERF) converged
    unsigned long long mail_seq = 0;
    long long step_counter = 0;
    // staging for the _host variants
    void* stage_dev = nullptr;
    int64_t stage_bytes = 0;
    int max_blocks = kNumSMs * 8;
    // fused four-step engine (power-of-two N = (32 q1) x (32 q2), single pol-pair): geometry + tables
    int engine = 0;       // OCB_ENGINE_*: 0 auto, 1 cuFFT, 2 fused
    bool fused_ok = false, fused_tables_ready = false;
    int q1 = 0, q2 = 0;
    float2 *tw1 = nullptr, *tw2 = nullptr, *tabV = nullptr, *tabU = nullptr;
    float2 *Cb = nullptr, *Nb = nullptr;  // third rotating field buffer, engine-layout noise copy
    // tuning / A-B knobs, read from the environment when the plan is created (fused_engine.inl)
    int knob_time_kernel = 0;       // OCB_TIME_KERNEL: 0 default (one-wave kernel), 1 "bulk", 2 "plain" (= default)
    bool knob_freq_tma = false;     // OCB_FREQ_TMA=1: tensor-map (TMA) tile loads / stores, k_freq_tma, instead of k_freq
    bool knob_freq_lockstep = false;  // OCB_FREQ_LOCKSTEP=1: CTA-wide instead of per-group barriers in k_freq_tma
    bool knob_freq_tma_store = true;  // OCB_FREQ_TMA_STORE=0: tensor loads, per-thread streaming stores
    int knob_freq_c = 16;           // OCB_FREQ_C=8: tile width of the classic k_freq at N2 = 1024
    CUtensorMap wmap;      // W buffer as float32 [rows][N2][2*N1] for the TMA-fed frequency pass (N2 = 1024)
    CUtensorMap* wmap_dev = nullptr;  // its copy in the workspace (device memory, 64-byte aligned)
    bool wmap_ok = false;
    // table cache keys
    double t1_h = NAN, t1_a = NAN, t1_b = NAN, t1_scale = NAN;
    // optional in-situ kernel timing (CUDA events on the launching stream); kinds:
    // 0 = fused NL iteration pass, 1 = NL first pass, 2 = linear half step (fft + multiply + ifft)
    bool prof_on = false;
    static constexpr int kProfKinds = 3, kProfCap = 2048;
    std::vector<cudaEvent_t> prof_ev[kProfKinds];
    int prof_n[kProfKinds] = {0, 0, 0};
};

namespace {
struct ProfScope {  // records an event pair around the launches issued while it is alive
    ocb_ssfm_plan* p; int kind; cudaStream_t st; bool active;
    ProfScope(ocb_ssfm_plan* p_, int kind_, cudaStream_t st_) : p(p_), kind(kind_), st(st_) {
        active = p->prof_on && p->prof_n[kind] < ocb_ssfm_plan::kProfCap;
        if (active) cudaEventRecord(p->prof_ev[kind][2 * p->prof_n[kind]], st);
    }
    ~ProfScope() {
        if (active) { cudaEventRecord(p->prof_ev[kind][2 * p->prof_n[kind] + 1], st); p->prof_n[kind]++; }
    }
};
}  // namespace

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

static bool fused_geometry(int64_t N, int* q1, int* q2) {
    if (N <= 0 || (N & (N - 1))) return false;
    int n = 0;
    while ((1ll << n) < N) ++n;
    const int n1 = (n + 1) / 2, n2 = n / 2;  // N1 >= N2
    if (n2 < 8 || n1 > 10) return false;      // 2^16 .. 2^20
    *q1 = (1 << n1) / 32;
    *q2 = (1 << n2) / 32;
    return true;
}
static int64_t fused_table_bytes(const ocb_ssfm_plan* p) {
    if (!p->fused_ok) return 0;
    // twiddle tables with their lo parts (double-single, fft_core.cuh): tw1, tw2 (hi, hi^T, lo, lo^T), V, U (hi, lo)
    return align_up(4 * 32ll * p->q1 * 8, 256) + align_up(4 * 32ll * p->q2 * 8, 256) +
           align_up(2 * 32ll * p->q2 * 32 * 8, 256) + align_up(2 * 32ll * p->q2 * p->q1 * 8, 256) +
           align_up((int64_t)p->rows * p->N * 8, 256) + align_up(p->N * 8, 256);
}
static int fused_manakov_run(ocb_ssfm_plan*, void*, const ocb_manakov_params*, const void*, const int32_t*, void*,
                             ocb_manakov_stats*, cudaStream_t);
static int fused_nlse_run(ocb_ssfm_plan*, void*, const ocb_nlse_params*, const void*, cudaStream_t);
static bool use_fused(const ocb_ssfm_plan* p) { return p->fused_ok && p->engine != OCB_ENGINE_CUFFT; }

// Tensor map of the W buffer for k_freq_tma: float32 elements, dims (fastest first) {2*N1, N2, rows}, box
// {2*C, 256, 1} = one 32-byte row segment x 256 rows.  cuTensorMapEncodeTiled is fetched through the runtime
// (cudaGetDriverEntryPoint), so the library does not link libcuda.  A driver without it leaves wmap_ok = false and the
// classic k_freq runs.
static int make_w_tensor_map(ocb_ssfm_plan* p) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return 0;
    }
    const cuuint64_t N1 = 32ull * p->q1, N2 = 32ull * p->q2;
    cuuint64_t dims[3] = {2 * N1, N2, (cuuint64_t)p->rows};
    cuuint64_t strides[2] = {N1 * 8, N1 * N2 * 8};  // bytes, dims 1 and 2
    cuuint32_t box[3] = {2 * FreqTmaCfg::C, FreqTmaCfg::BOX_ROWS, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = ((EncodeFn)fn)(&p->wmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, p->G, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char b[128]; snprintf(b, sizeof b, "cuTensorMapEncodeTiled failed (%d)", (int)r);
        return fail(b, __FILE__, __LINE__);
    }
    OCB_CUDA(cudaMemcpy(p->wmap_dev, &p->wmap, sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    p->wmap_ok = true;
    return 0;
}

extern "C" int ocb_abi_version(void) { return OCB_ABI_VERSION; }
extern "C" const char* ocb_last_error(void) { return last_error().c_str(); }
extern "C" int64_t ocb_launch_count(void) { return launch_counter(); }
extern "C" void ocb_launch_count_reset(void) { launch_counter() = 0; }

extern "C" int ocb_ssfm_plan_create(int64_t N, int rows, ocb_ssfm_plan** out) {
    OCB_REQUIRE(out != nullptr, "plan_create: out is NULL");
    OCB_REQUIRE(N >= 2 && N < (1ll << 31), "plan_create: N out of range");
    OCB_REQUIRE(rows >= 1 && rows <= 4096, "plan_create: rows out of range");
    ocb_ssfm_plan* p = new ocb_ssfm_plan();
    p->N = N;
    p->rows = rows;
    p->fused_ok = (rows == 1 || rows == 2) && fused_geometry(N, &p->q1, &p->q2);
    {
        const char* v = getenv("OCB_TIME_KERNEL");
        p->knob_time_kernel = (v && !strcmp(v, "bulk")) ? 1 : (v && !strcmp(v, "plain")) ? 2 : 0;
        p->knob_freq_tma = getenv("OCB_FREQ_TMA") && atoi(getenv("OCB_FREQ_TMA")) != 0;
        p->knob_freq_lockstep = getenv("OCB_FREQ_LOCKSTEP") && atoi(getenv("OCB_FREQ_LOCKSTEP")) != 0;
        p->knob_freq_tma_store = !(getenv("OCB_FREQ_TMA_STORE") && atoi(getenv("OCB_FREQ_TMA_STORE")) == 0);
        p->knob_freq_c = (getenv("OCB_FREQ_C") && atoi(getenv("OCB_FREQ_C")) == 8) ? 8 : 16;
    }
    if (cufftCreate(&p->fft) != CUFFT_SUCCESS) { delete p; return fail("cufftCreate failed", __FILE__, __LINE__); }
    p->fft_ok = true;
    if (cufftSetAutoAllocation(p->fft, 0) != CUFFT_SUCCESS) { ocb_ssfm_plan_destroy(p); return fail("cufftSetAutoAllocation failed", __FILE__, __LINE__); }
    int n[1] = {(int)N};
    cufftResult r = cufftMakePlanMany(p->fft, 1, n, nullptr, 1, (int)N, nullptr, 1, (int)N, CUFFT_C2C,
                                      rows, &p->fft_work);
    if (r != CUFFT_SUCCESS) {
        ocb_ssfm_plan_destroy(p);
        char b[128]; snprintf(b, sizeof b, "cufftMakePlanMany failed (%d)", (int)r);
        return fail(b, __FILE__, __LINE__);
    }
    cudaError_t e = cudaHostAlloc((void**)&p->h_sums, 8 * sizeof(double), cudaHostAllocDefault);
    if (e != cudaSuccess) { ocb_ssfm_plan_destroy(p); return fail("cudaHostAlloc failed", __FILE__, __LINE__); }
    e = cudaHostAlloc((void**)&p->h_mail, kMailSlots * sizeof(Mail), cudaHostAllocMapped);
    if (e != cudaSuccess) { ocb_ssfm_plan_destroy(p); return fail("cudaHostAlloc(mapped) failed", __FILE__, __LINE__); }
    memset(p->h_mail, 0, kMailSlots * sizeof(Mail));
    e = cudaHostGetDevicePointer((void**)&p->d_mail, p->h_mail, 0);
    if (e != cudaSuccess) { ocb_ssfm_plan_destroy(p); return fail("cudaHostGetDevicePointer failed", __FILE__, __LINE__); }
    *out = p;
    return 0;
}

extern "C" int64_t ocb_ssfm_plan_workspace_bytes(const ocb_ssfm_plan* p) {
    if (!p) return -1;
    const int64_t field = align_up((int64_t)p->rows * p->N * sizeof(float2), 256);
    int64_t b = 0;
    b += 4 * field;                                                   // Ehd, A, B, G
    b += 2 * align_up(p->N * (int64_t)sizeof(float2), 256);           // T1, T2
    b += align_up(((int64_t)(p->rows + 1) / 2) * p->N * sizeof(float), 256);  // Pch
    b += align_up((int64_t)p->max_blocks * 3 * sizeof(double), 256);  // partials
    b += 256;                                                         // sums + ticket
    b += 256;                                                         // tensor map of the W buffer
    b += align_up((int64_t)p->fft_work, 256);
    b += fused_table_bytes(p);
    return b;
}

extern "C" int ocb_ssfm_plan_bind_workspace(ocb_ssfm_plan* p, void* dev_ptr, int64_t bytes) {
    OCB_REQUIRE(p && dev_ptr, "bind_workspace: NULL argument");
    OCB_REQUIRE(bytes >= ocb_ssfm_plan_workspace_bytes(p), "bind_workspace: workspace too small");
    OCB_REQUIRE(((uintptr_t)dev_ptr & 255) == 0, "bind_workspace: pointer must be 256-byte aligned");
    const int64_t field = align_up((int64_t)p->rows * p->N * sizeof(float2), 256);
    char* c = (char*)dev_ptr;
    p->ws = dev_ptr; p->ws_bytes = bytes;
    p->Ehd = (float2*)c; c += field;
    p->A = (float2*)c; c += field;
    p->B = (float2*)c; c += field;
    p->G = (float2*)c; c += field;
    p->T1 = (float2*)c; c += align_up(p->N * (int64_t)sizeof(float2), 256);
    p->T2 = (float2*)c; c += align_up(p->N * (int64_t)sizeof(float2), 256);
    p->Pch = (float*)c; c += align_up(((int64_t)(p->rows + 1) / 2) * p->N * sizeof(float), 256);
    p->partials = (double*)c; c += align_up((int64_t)p->max_blocks * 3 * sizeof(double), 256);
    p->sums = (double*)c; p->ticket = (unsigned*)(c + 64); p->conv_flag = (long long*)(c + 128); p->final_flag = (long long*)(c + 192); c += 256;
    p->wmap_dev = (CUtensorMap*)c; c += 256;
    p->fft_area = c; c += align_up((int64_t)p->fft_work, 256);
    if (p->fft_work > 0) OCB_CUFFT(cufftSetWorkArea(p->fft, p->fft_area));
    if (p->fused_ok) {
        p->tw1 = (float2*)c; c += align_up(4 * 32ll * p->q1 * 8, 256);
        p->tw2 = (float2*)c; c += align_up(4 * 32ll * p->q2 * 8, 256);
        p->tabV = (float2*)c; c += align_up(2 * 32ll * p->q2 * 32 * 8, 256);
        p->tabU = (float2*)c; c += align_up(2 * 32ll * p->q2 * p->q1 * 8, 256);
        p->Cb = (float2*)c; c += align_up((int64_t)p->rows * p->N * 8, 256);
        p->Nb = (float2*)c; c += align_up(p->N * 8, 256);
        p->fused_tables_ready = false;
        p->wmap_ok = false;
        if (p->q2 == 32 && make_w_tensor_map(p)) return 1;
    }
    OCB_CUDA(cudaMemset(p->sums, 0, 256));
    p->t1_h = NAN;
    return 0;
}

extern "C" int ocb_ssfm_plan_set_engine(ocb_ssfm_plan* p, int engine) {
    OCB_REQUIRE(p != nullptr, "plan_set_engine: NULL plan");
    OCB_REQUIRE(engine >= OCB_ENGINE_AUTO && engine <= OCB_ENGINE_FUSED, "plan_set_engine: unknown engine");
    OCB_REQUIRE(engine != OCB_ENGINE_FUSED || p->fused_ok,
                "plan_set_engine: the fused four-step engine needs N = 2^16..2^20 and a single pol-pair");
    p->engine = engine;
    return 0;
}
extern "C" int ocb_ssfm_plan_engine(const ocb_ssfm_plan* p) {
    if (!p) return -1;
    return use_fused(p) ? OCB_ENGINE_FUSED : OCB_ENGINE_CUFFT;
}

extern "C" int ocb_ssfm_plan_profile(ocb_ssfm_plan* p, int enable) {
    OCB_REQUIRE(p != nullptr, "plan_profile: NULL plan");
    if (enable && p->prof_ev[0].empty()) {
        for (int k = 0; k < ocb_ssfm_plan::kProfKinds; ++k) {
            p->prof_ev[k].resize(2 * ocb_ssfm_plan::kProfCap);
            for (auto& e : p->prof_ev[k]) OCB_CUDA(cudaEventCreate(&e));
        }
    }
    p->prof_on = enable != 0;
    for (int k = 0; k < ocb_ssfm_plan::kProfKinds; ++k) p->prof_n[k] = 0;
    return 0;
}

// out[2k] = summed duration [ms] of the recorded launches of kind k, out[2k+1] = their count
extern "C" int ocb_ssfm_plan_profile_read(ocb_ssfm_plan* p, double* out6) {
    OCB_REQUIRE(p && out6, "plan_profile_read: NULL argument");
    OCB_CUDA(cudaDeviceSynchronize());
    for (int k = 0; k < ocb_ssfm_plan::kProfKinds; ++k) {
        double tot = 0.0;
        for (int i = 0; i < p->prof_n[k]; ++i) {
            float ms = 0.f;
            OCB_CUDA(cudaEventElapsedTime(&ms, p->prof_ev[k][2 * i], p->prof_ev[k][2 * i + 1]));
            tot += ms;
        }
        out6[2 * k] = tot;
        out6[2 * k + 1] = (double)p->prof_n[k];
        p->prof_n[k] = 0;
    }
    return 0;
}

extern "C" int ocb_ssfm_plan_destroy(ocb_ssfm_plan* p) {
    if (!p) return 0;
    // lines of the operator tables that were marked persisting in L2 during fused runs go back to normal before the
    // workspace can be freed by its owner
    if (p->fused_ok && p->ws) { cudaCtxResetPersistingL2Cache(); cudaGetLastError(); }
    for (int k = 0; k < ocb_ssfm_plan::kProfKinds; ++k)
        for (auto& e : p->prof_ev[k]) cudaEventDestroy(e);
    if (p->fft_ok) cufftDestroy(p->fft);
    if (p->h_sums) cudaFreeHost(p->h_sums);
    if (p->h_mail) cudaFreeHost(p->h_mail);
    if (p->stage_dev) cudaFree(p->stage_dev);
    delete p;
    return 0;
}

// ---- layout conversion -------------------------------------------------------------------
extern "C" int ocb_pack_fields(const void* src, int src_dtype, int64_t N, int C, int pairs, void* rows,
                               void* stream) {
    OCB_REQUIRE(src && rows && N > 0 && C > 0, "pack_fields: bad argument");
    OCB_REQUIRE(!pairs || (C % 2 == 0), "pack_fields: pairs layout needs an even column count");
    cudaStream_t st = (cudaStream_t)stream;
    int g = grid_for(N * C, 256, 1);
    if (src_dtype == OCB_C64) OCB_LAUNCH(k_pack<float2>, g, 256, 0, st, (const float2*)src, (float2*)rows, N, C, pairs);
    else if (src_dtype == OCB_C128) OCB_LAUNCH(k_pack<double2>, g, 256, 0, st, (const double2*)src, (float2*)rows, N, C, pairs);
    else return fail("pack_fields: unknown dtype", __FILE__, __LINE__);
    return 0;
}
extern "C" int ocb_unpack_fields(const void* rows, int64_t N, int C, int pairs, void* dst, int dst_dtype,
                                 void* stream) {
    OCB_REQUIRE(dst && rows && N > 0 && C > 0, "unpack_fields: bad argument");
    OCB_REQUIRE(!pairs || (C % 2 == 0), "unpack_fields: pairs layout needs an even column count");
    cudaStream_t st = (cudaStream_t)stream;
    int g = grid_for(N * C, 256, 1);
    if (dst_dtype == OCB_C64) OCB_LAUNCH(k_unpack<float2>, g, 256, 0, st, (const float2*)rows, (float2*)dst, N, C, pairs);
    else if (dst_dtype == OCB_C128) OCB_LAUNCH(k_unpack<double2>, g, 256, 0, st, (const float2*)rows, (double2*)dst, N, C, pairs);
    else return fail("unpack_fields: unknown dtype", __FILE__, __LINE__);
    return 0;
}

// dtype conversion without reordering (complex64 <-> complex128), e.g. host arrays uploaded raw
template <typename TI, typename TO>
__global__ void k_cast_complex(const TI* __restrict__ src, TO* __restrict__ dst, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        TI v = src[i];
        TO o; o.x = v.x; o.y = v.y;
        dst[i] = o;
    }
}
extern "C" int ocb_cast_complex(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream) {
    OCB_REQUIRE(src && dst && n >= 0, "cast_complex: bad argument");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int g = grid_for(n, 256, 2);
    if (src_dtype == OCB_C128 && dst_dtype == OCB_C64) OCB_LAUNCH((k_cast_complex<double2, float2>), g, 256, 0, st, (const double2*)src, (float2*)dst, n);
    else if (src_dtype == OCB_C64 && dst_dtype == OCB_C128) OCB_LAUNCH((k_cast_complex<float2, double2>), g, 256, 0, st, (const float2*)src, (double2*)dst, n);
    else if (src_dtype == OCB_C64 && dst_dtype == OCB_C64) OCB_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    else if (src_dtype == OCB_C128 && dst_dtype == OCB_C128) OCB_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * 16, cudaMemcpyDeviceToDevice, st));
    else return fail("cast_complex: unknown dtype", __FILE__, __LINE__);
    return 0;
}

// ---- small launch helpers --------------------------------------------------------------------
static int launch_table(ocb_ssfm_plan* p, float2* T, double a, double b, double Fs, double h,
                        double scale, cudaStream_t st) {
    OCB_LAUNCH(k_linop_table, grid_for(p->N, 256, 1), 256, 0, st, T, p->N, a, b, Fs, h, scale);
    return 0;
}
static int launch_mul(ocb_ssfm_plan* p, float2* F, const float2* T, cudaStream_t st) {
    if (p->N % 2 == 0) OCB_LAUNCH(k_mul_table<2>, grid_for(p->N / 2 * p->rows, 256, 1), 256, 0, st, F, T, p->N, p->rows);
    else OCB_LAUNCH(k_mul_table<1>, grid_for(p->N * p->rows, 256, 1), 256, 0, st, F, T, p->N, p->rows);
    return 0;
}
static int launch_amp(float2* E, int R, int64_t N, double g, double sigma, const float2* noise,
                      int noise_rows, uint64_t seed, uint64_t stream_id, cudaStream_t st, int tN1 = 0, int tN2 = 0) {
    OCB_LAUNCH(k_amp, grid_for((int64_t)R * N, 256, 2), 256, 0, st, E, R, N, (float)g, (float)sigma, noise,
               noise_rows, seed, stream_id, tN1, tN2);
    return 0;
}
static int nl_grid(const ocb_ssfm_plan* p, int64_t items) {
    // an exact multiple of the SM count (2 CTAs of 256 threads per SM; software-pipelined grid-stride loop)
    int64_t need = (items + 511) / 512;
    static const int per_sm = getenv("OCB_NL_CTAS") ? atoi(getenv("OCB_NL_CTAS")) : 2;  // tuning knob (2 measured best)
    int g = kNumSMs * (per_sm > 0 ? per_sm : 2);
    if (need < g) g = (int)(need < 1 ? 1 : need);
    return g > p->max_blocks ? p->max_blocks : g;
}
static int launch_nl(ocb_ssfm_plan* p, bool first, const float2* Ehd, const float2* Efd, const float2* Ec,
                     float* Pch, float2* out, int64_t N, int K, float cphi, double* partials, double* sums,
                     unsigned* ticket, cudaStream_t st) {
    const bool v2 = (N % 2 == 0);
    const int g = nl_grid(p, v2 ? N / 2 * K : N * K);
    if (first) {
        if (v2) OCB_LAUNCH((k_manakov_nl<true, 2>), g, 256, 0, st, Ehd, Efd, Ec, Pch, out, N, K, cphi, partials, sums, ticket);
        else OCB_LAUNCH((k_manakov_nl<true, 1>), g, 256, 0, st, Ehd, Efd, Ec, Pch, out, N, K, cphi, partials, sums, ticket);
    } else {
        if (v2) OCB_LAUNCH((k_manakov_nl<false, 2>), g, 256, 0, st, Ehd, Efd, Ec, Pch, out, N, K, cphi, partials, sums, ticket);
        else OCB_LAUNCH((k_manakov_nl<false, 1>), g, 256, 0, st, Ehd, Efd, Ec, Pch, out, N, K, cphi, partials, sums, ticket);
    }
    return 0;
}
static int launch_power_stats(ocb_ssfm_plan* p, const float2* E, int64_t N, int K, cudaStream_t st) {
    const bool v2 = (N % 2 == 0);
    const int g = nl_grid(p, v2 ? N / 2 * K : N * K);
    if (v2) OCB_LAUNCH(k_power_stats<2>, g, 256, 0, st, E, N, K, p->partials, p->sums, p->ticket);
    else OCB_LAUNCH(k_power_stats<1>, g, 256, 0, st, E, N, K, p->partials, p->sums, p->ticket);
    return 0;
}
// Wait until the finalising block of the launch tagged `seq` has published its mail (spin on mapped
// pinned memory; the stream is polled now and then so that a device fault cannot hang the host).
static int wait_mail(ocb_ssfm_plan* p, unsigned long long seq, cudaStream_t st) {
    volatile Mail* mb = p->h_mail + (seq % kMailSlots);
    unsigned spins = 0;
    while (mb->seq != seq) {
        if ((++spins & 0xFFFF) == 0) {
            cudaError_t e = cudaStreamQuery(st);
            if (e != cudaSuccess && e != cudaErrorNotReady) return fail(cudaGetErrorString(e), __FILE__, __LINE__);
            if (e == cudaSuccess && mb->seq != seq) return fail("wait_mail: stream drained without the expected mail", __FILE__, __LINE__);
        }
    }
    __sync_synchronize();
    p->h_sums[0] = mb->sums[0]; p->h_sums[1] = mb->sums[1]; p->h_sums[2] = mb->sums[2];
    p->h_sums[3] = (double)mb->converged;
    return 0;
}

static int fetch_sums(ocb_ssfm_plan* p, cudaStream_t st) {
    OCB_CUDA(cudaMemcpyAsync(p->h_sums, p->sums, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    OCB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int ocb_manakov_nl_pass(const void* Ehd, const void* Efd, const void* Ec, void* Pch, void* out,
                                   void* sums3_dev, int64_t N, int K, double gamma, double hz, int direction,
                                   void* stream) {
    OCB_REQUIRE(Ehd && Ec && Pch && out && N > 0 && K > 0, "nl_pass: bad argument");
    OCB_REQUIRE(direction == 1 || direction == -1, "nl_pass: direction must be +1 or -1");
    cudaStream_t st = (cudaStream_t)stream;
    ocb_ssfm_plan tmp;  // only max_blocks is used
    const bool first = (Efd == nullptr);
    // reduction scratch: allocated once per device and reused (the ticket resets itself after every launch)
    static thread_local double* scratch = nullptr;
    static thread_local int scratch_dev = -1;
    int dev = 0;
    OCB_CUDA(cudaGetDevice(&dev));
    const size_t part_bytes = (size_t)tmp.max_blocks * 3 * sizeof(double);
    if (!first && (scratch == nullptr || scratch_dev != dev)) {
        OCB_REQUIRE(sums3_dev != nullptr, "nl_pass: sums3_dev is NULL");
        OCB_CUDA(cudaMalloc((void**)&scratch, part_bytes + 64));
        OCB_CUDA(cudaMemset(scratch, 0, part_bytes + 64));
        scratch_dev = dev;
    }
    if (!first) OCB_REQUIRE(sums3_dev != nullptr, "nl_pass: sums3_dev is NULL");
    const double c = (double)direction * hz * (8.0 / 9.0) * gamma * (first ? 1.0 : 0.5);
    return launch_nl(&tmp, first, (const float2*)Ehd, (const float2*)Efd, (const float2*)Ec, (float*)Pch,
                     (float2*)out, N, K, (float)c, first ? nullptr : scratch, (double*)sums3_dev,
                     first ? nullptr : (unsigned*)((char*)scratch + part_bytes), st);
}

extern "C" int ocb_edfa_apply(void* rows_inout, int rows, int64_t N, double gain_lin, double noise_var,
                              int noise_mode, const void* noise_dev, int noise_rows, uint64_t seed,
                              uint64_t stream_id, void* stream) {
    OCB_REQUIRE(rows_inout && rows > 0 && N > 0, "edfa_apply: bad argument");
    OCB_REQUIRE(gain_lin > 0.0 && noise_var >= 0.0, "edfa_apply: gain must be > 0 and noise_var >= 0");
    if (noise_mode == OCB_NOISE_INJECTED) OCB_REQUIRE(noise_dev && noise_rows > 0, "edfa_apply: injected noise buffer missing");
    return launch_amp((float2*)rows_inout, rows, N, sqrt(gain_lin),
                      noise_mode == OCB_NOISE_PHILOX ? sqrt(noise_var / 2.0) : 0.0,
                      noise_mode == OCB_NOISE_INJECTED ? (const float2*)noise_dev : nullptr,
                      noise_rows > 0 ? noise_rows : 1, seed, stream_id, (cudaStream_t)stream);
}

// ---- Manakov SSF / DBP -------------------------------------------------------------------------
extern "C" int ocb_manakov_run(ocb_ssfm_plan* p, void* rows_inout, const ocb_manakov_params* q,
                               const void* noise_dev, const int32_t* save_spans, void* save_dev,
                               ocb_manakov_stats* stats, void* stream) {
    OCB_REQUIRE(p && rows_inout && q, "manakov_run: NULL argument");
    OCB_REQUIRE(p->ws != nullptr, "manakov_run: workspace not bound");
    OCB_REQUIRE(p->rows % 2 == 0, "manakov_run: plan rows must be even (x/y pairs)");
    OCB_REQUIRE(q->direction == 1 || q->direction == -1, "manakov_run: direction must be +1/-1");
    OCB_REQUIRE(q->maxIter >= 1, "manakov_run: maxIter must be >= 1");
    OCB_REQUIRE(q->Lspan > 0, "manakov_run: Lspan must be > 0");
    OCB_REQUIRE(q->nlprMethod || q->hz > 0, "manakov_run: hz must be > 0");
    // nlprMethod=True with gamma == 0: maxNlinPhaseRot / 0 = +inf in IEEE doubles, so the whole span is one linear
    // step, exactly what the reference does (channels.py:394-397, with a RuntimeWarning)
    if (q->amp_mode == OCB_AMP_EDFA && q->direction == 1 && q->noise_mode == OCB_NOISE_INJECTED)
        OCB_REQUIRE(noise_dev != nullptr, "manakov_run: injected noise buffer missing");
    OCB_REQUIRE(q->n_save == 0 || (save_spans && save_dev), "manakov_run: snapshot buffers missing");

    cudaStream_t st = (cudaStream_t)stream;
    if (use_fused(p) && p->rows == 2)
        return fused_manakov_run(p, rows_inout, q, noise_dev, save_spans, save_dev, stats, st);
    OCB_CUFFT(cufftSetStream(p->fft, st));
    const int64_t N = p->N;
    const int R = p->rows, K = R / 2;
    const double dir = (double)q->direction;
    // argLimOp = -(α/2) + j(β2/2)ω²  (channels.py:368) ; DBP: +(α/2) - j(β2/2)ω² (equalization.py:1077)
    const double a = -dir * q->alpha_lin / 2.0, b = dir * q->beta2 / 2.0;
    const size_t field_bytes = (size_t)R * N * sizeof(float2);

    float2* bufs[3] = {(float2*)rows_inout, p->A, p->B};
    int cur = 0;
    ocb_manakov_stats S = {0, 0, 0, 0.0, 0.0};
    double table_h = NAN;
    int next_save = 0;
    uint64_t amp_calls = 0;

    for (int span = 1; span <= q->n_spans; ++span) {
        if (q->direction < 0 && q->amp_mode != OCB_AMP_NONE) {
            // undo the span gain first (equalization.py:1090-1092)
            if (launch_amp(bufs[cur], R, N, exp(-q->alpha_lin / 2.0 * q->Lspan), 0.0, nullptr, 1, 0, 0, st)) return 1;
        }
        double maxP = 0.0;
        if (q->nlprMethod) {
            if (launch_power_stats(p, bufs[cur], N, K, st)) return 1;
            if (fetch_sums(p, st)) return 1;
            maxP = p->h_sums[2];
        }
        double z = 0.0;
        while (z < q->Lspan) {  // channels.py:387
            double hz_;
            if (q->nlprMethod) {  // channels.py:392-397
                const double phimax = (8.0 / 9.0) * q->gamma * maxP;
                const double cand = q->maxNlinPhaseRot / phimax;
                hz_ = (q->Lspan - z >= cand) ? cand : (q->Lspan - z);
            } else if (q->Lspan - z < q->hz) {  // channels.py:398-401
                hz_ = q->Lspan - z;
            } else {
                hz_ = q->hz;
            }
            if (!(table_h == hz_)) {  // linOperator = exp(argLimOp*hz_/2), rebuilt only when hz_ changes
                if (launch_table(p, p->T1, a, b, q->Fs, hz_ / 2.0, 1.0 / (double)N, st)) return 1;
                table_h = hz_;
            }
            // first half step: Ehd = ifft(fft(E)·L)    (channels.py:409-410)
            {
                ProfScope ps(p, 2, st);
                OCB_CUFFT(cufftExecC2C(p->fft, bufs[cur], p->G, CUFFT_FORWARD));
                if (launch_mul(p, p->G, p->T1, st)) return 1;
                OCB_CUFFT(cufftExecC2C(p->fft, p->G, p->Ehd, CUFFT_INVERSE));
            }

            int ec = cur, dst = (cur + 1) % 3;
            const float c_first = (float)(dir * hz_ * (8.0 / 9.0) * q->gamma);
            const float c_iter = (float)(dir * hz_ * (8.0 / 9.0) * q->gamma * 0.5);
            {
                ProfScope ps(p, 1, st);
                if (launch_nl(p, true, p->Ehd, nullptr, bufs[ec], p->Pch, p->G, N, K, c_first, nullptr, nullptr, nullptr, st)) return 1;
            }
            for (int it = 0; it < q->maxIter; ++it) {  // channels.py:413
                // second half step on the rotated field  (channels.py:420-421)
                {
                    ProfScope ps(p, 2, st);
                    OCB_CUFFT(cufftExecC2C(p->fft, p->G, p->G, CUFFT_FORWARD));
                    if (launch_mul(p, p->G, p->T1, st)) return 1;
                    OCB_CUFFT(cufftExecC2C(p->fft, p->G, bufs[dst], CUFFT_INVERSE));
                }
                // convergence sums + speculative next rotation in one pass (channels.py:424, 436, 414-417)
                {
                    ProfScope ps(p, 0, st);
                    if (launch_nl(p, false, p->Ehd, bufs[dst], bufs[ec], p->Pch, p->G, N, K, c_iter, p->partials, p->sums, p->ticket, st)) return 1;
                }
                if (fetch_sums(p, st)) return 1;
                const double lim = sqrt(p->h_sums[0]) / sqrt(p->h_sums[1]);  // channels.py:517-519
                S.iterations++;
                S.last_lim = lim;
                const int third = 3 - ec - dst;
                ec = dst;      // Ex_conv = Ech_x_fd  (channels.py:426-427)
                dst = third;
                if (lim < q->tol) break;  // channels.py:429
                if (it == q->maxIter - 1) S.nonconverged++;  // channels.py:431-434 (warning only)
            }
            cur = ec;  // Ech = Ech_fd (channels.py:438-439)
            maxP = p->h_sums[2];
            z += hz_;  // channels.py:441
            S.steps++;
            S.z_last_step = hz_;
        }
        if (q->direction > 0) {  // amplification (channels.py:443-451)
            if (q->amp_mode == OCB_AMP_EDFA) {
                const bool inj = (q->noise_mode == OCB_NOISE_INJECTED);
                if (launch_amp(bufs[cur], R, N, sqrt(q->edfa_gain_lin), inj ? 0.0 : sqrt(q->edfa_noise_var / 2.0),
                               inj ? (const float2*)noise_dev : nullptr, K, q->seed, amp_calls++, st)) return 1;
            } else if (q->amp_mode == OCB_AMP_IDEAL) {
                if (launch_amp(bufs[cur], R, N, exp(q->alpha_lin / 2.0 * q->Lspan), 0.0, nullptr, 1, 0, 0, st)) return 1;
            }
        }
        if (next_save < q->n_save && save_spans[next_save] == span) {  // channels.py:453-456
            OCB_CUDA(cudaMemcpyAsync((char*)save_dev + (size_t)next_save * field_bytes, bufs[cur], field_bytes,
                                     cudaMemcpyDeviceToDevice, st));
            next_save++;
        }
    }
    if (cur != 0) OCB_CUDA(cudaMemcpyAsync(rows_inout, bufs[cur], field_bytes, cudaMemcpyDeviceToDevice, st));
    if (stats) *stats = S;
    return 0;
}

static int ensure_stage(ocb_ssfm_plan* p, int64_t bytes) {
    if (p->stage_bytes >= bytes) return 0;
    if (p->stage_dev) cudaFree(p->stage_dev);
    p->stage_dev = nullptr; p->stage_bytes = 0;
    OCB_CUDA(cudaMalloc(&p->stage_dev, (size_t)bytes));
    p->stage_bytes = bytes;
    return 0;
}

extern "C" int ocb_manakov_run_host(ocb_ssfm_plan* p, const void* Ei_host, int in_dtype, void* Eo_host,
                                    int out_dtype, const ocb_manakov_params* q, const void* noise_host,
                                    const int32_t* save_spans, ocb_manakov_stats* stats, void* stream) {
    OCB_REQUIRE(p && Ei_host && Eo_host && q, "manakov_run_host: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t N = p->N;
    const int R = p->rows, K = R / 2;
    const int64_t in_elem = in_dtype == OCB_C128 ? 16 : 8, out_elem = out_dtype == OCB_C128 ? 16 : 8;
    const int nsnap = q->n_save > 0 ? q->n_save : 1;
    const int64_t field = align_up((int64_t)R * N * 8, 256);
    // staging: [raw in/out (max of both)] [rows] [snapshots] [noise]
    const int64_t raw = align_up(std::max<int64_t>(R * N * in_elem, (int64_t)R * nsnap * N * out_elem), 256);
    const int64_t noise_b = noise_host ? align_up((int64_t)K * N * 8, 256) : 0;
    if (ensure_stage(p, raw + field + (q->n_save > 0 ? (int64_t)nsnap * field : 0) + noise_b)) return 1;
    char* c = (char*)p->stage_dev;
    void* d_raw = c; c += raw;
    void* d_rows = c; c += field;
    void* d_save = nullptr;
    if (q->n_save > 0) { d_save = c; c += (int64_t)nsnap * field; }
    void* d_noise = nullptr;
    if (noise_host) { d_noise = c; }

    OCB_CUDA(cudaMemcpyAsync(d_raw, Ei_host, (size_t)(R * N * in_elem), cudaMemcpyHostToDevice, st));
    if (noise_host) OCB_CUDA(cudaMemcpyAsync(d_noise, noise_host, (size_t)K * N * 8, cudaMemcpyHostToDevice, st));
    if (ocb_pack_fields(d_raw, in_dtype, N, R, 1, d_rows, stream)) return 1;
    if (ocb_manakov_run(p, d_rows, q, d_noise, save_spans, d_save, stats, stream)) return 1;
    if (q->n_save > 0) {
        // output (N, 2K*n_save): snapshot s occupies columns [2K s, 2K (s+1))  (channels.py:454-455, K=1)
        // unpack each snapshot into a (N, R) block, then the host interleaves blocks column-wise.
        for (int s = 0; s < nsnap; ++s)
            if (ocb_unpack_fields((char*)d_save + (int64_t)s * field, N, R, 1,
                                  (char*)d_raw + (int64_t)s * R * N * out_elem, out_dtype, stream)) return 1;
        OCB_CUDA(cudaMemcpyAsync(Eo_host, d_raw, (size_t)((int64_t)nsnap * R * N * out_elem), cudaMemcpyDeviceToHost, st));
    } else {
        if (ocb_unpack_fields(d_rows, N, R, 1, d_raw, out_dtype, stream)) return 1;
        OCB_CUDA(cudaMemcpyAsync(Eo_host, d_raw, (size_t)(R * N * out_elem), cudaMemcpyDeviceToHost, st));
    }
    OCB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// ---- scalar NLSE ---------------------------------------------------------------------------------
extern "C" int ocb_nlse_run(ocb_ssfm_plan* p, void* row_inout, const ocb_nlse_params* q, const void* noise_dev,
                            void* stream) {
    OCB_REQUIRE(p && row_inout && q, "nlse_run: NULL argument");
    OCB_REQUIRE(p->ws != nullptr, "nlse_run: workspace not bound");
    OCB_REQUIRE(q->n_steps >= 0 && q->n_spans >= 0, "nlse_run: negative step/span count");
    if (q->amp_mode == OCB_AMP_EDFA && q->noise_mode == OCB_NOISE_INJECTED)
        OCB_REQUIRE(noise_dev != nullptr, "nlse_run: injected noise buffer missing");
    cudaStream_t st = (cudaStream_t)stream;
    if (use_fused(p) && p->rows == 1) return fused_nlse_run(p, row_inout, q, noise_dev, st);
    OCB_CUFFT(cufftSetStream(p->fft, st));
    const int64_t N = p->N;
    const int R = p->rows;
    float2* E = (float2*)row_inout;
    const double a = -q->alpha_lin / 2.0, b = q->beta2 / 2.0;
    // T1 = L/N (entering / leaving the frequency domain), T2 = L²/N (between two steps):
    // channels.py:221 and :229 are two multiplies by the same operator, merged here.
    if (launch_table(p, p->T1, a, b, q->Fs, q->hz / 2.0, 1.0 / (double)N, st)) return 1;
    if (launch_table(p, p->T2, a, b, q->Fs, q->hz, 1.0 / (double)N, st)) return 1;
    p->t1_h = NAN;
    const float cnl = (float)(q->gamma * q->hz);
    const int64_t total = (int64_t)R * N;
    for (int span = 0; span < q->n_spans; ++span) {
        if (q->n_steps > 0) {
            OCB_CUFFT(cufftExecC2C(p->fft, E, E, CUFFT_FORWARD));  // channels.py:216
            for (int s = 0; s < q->n_steps; ++s) {
                if (launch_mul(p, E, s == 0 ? p->T1 : p->T2, st)) return 1;  // :221 (+ :229 of the previous step)
                OCB_CUFFT(cufftExecC2C(p->fft, E, E, CUFFT_INVERSE));        // :224
                if (total % 2 == 0 && N % 2 == 0) OCB_LAUNCH(k_nlse_phase<2>, grid_for(total / 2, 256, 1), 256, 0, st, E, total, cnl);  // :225
                else OCB_LAUNCH(k_nlse_phase<1>, grid_for(total, 256, 1), 256, 0, st, E, total, cnl);
                OCB_CUFFT(cufftExecC2C(p->fft, E, E, CUFFT_FORWARD));        // :228
            }
            if (launch_mul(p, E, p->T1, st)) return 1;                       // :229 of the last step
            OCB_CUFFT(cufftExecC2C(p->fft, E, E, CUFFT_INVERSE));            // :232
        }
        if (q->amp_mode == OCB_AMP_EDFA) {  // :233-234
            const bool inj = (q->noise_mode == OCB_NOISE_INJECTED);
            if (launch_amp(E, R, N, sqrt(q->edfa_gain_lin), inj ? 0.0 : sqrt(q->edfa_noise_var / 2.0),
                           inj ? (const float2*)noise_dev : nullptr, R, q->seed, (uint64_t)span, st)) return 1;
        } else if (q->amp_mode == OCB_AMP_IDEAL) {  // :235-236
            if (launch_amp(E, R, N, exp(q->alpha_lin / 2.0 * q->n_steps * q->hz), 0.0, nullptr, 1, 0, 0, st)) return 1;
        }
    }
    return 0;
}

extern "C" int ocb_nlse_run_host(ocb_ssfm_plan* p, const void* Ei_host, int in_dtype, void* Eo_host, int out_dtype,
                                 const ocb_nlse_params* q, const void* noise_host, void* stream) {
    OCB_REQUIRE(p && Ei_host && Eo_host && q, "nlse_run_host: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t N = p->N;
    const int R = p->rows;
    const int64_t in_elem = in_dtype == OCB_C128 ? 16 : 8, out_elem = out_dtype == OCB_C128 ? 16 : 8;
    const int64_t raw = align_up(R * N * std::max(in_elem, out_elem), 256);
    const int64_t field = align_up((int64_t)R * N * 8, 256);
    if (ensure_stage(p, raw + field + (noise_host ? field : 0))) return 1;
    char* c = (char*)p->stage_dev;
    void* d_raw = c; c += raw;
    void* d_rows = c; c += field;
    void* d_noise = noise_host ? c : nullptr;
    OCB_CUDA(cudaMemcpyAsync(d_raw, Ei_host, (size_t)(R * N * in_elem), cudaMemcpyHostToDevice, st));
    if (noise_host) OCB_CUDA(cudaMemcpyAsync(d_noise, noise_host, (size_t)R * N * 8, cudaMemcpyHostToDevice, st));
    if (ocb_pack_fields(d_raw, in_dtype, N, R, 0, d_rows, stream)) return 1;
    if (ocb_nlse_run(p, d_rows, q, d_noise, stream)) return 1;
    if (ocb_unpack_fields(d_rows, N, R, 0, d_raw, out_dtype, stream)) return 1;
    OCB_CUDA(cudaMemcpyAsync(Eo_host, d_raw, (size_t)(R * N * out_elem), cudaMemcpyDeviceToHost, st));
    OCB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

#include "fused_engine.inl"
