// Shared helpers for the sm_100a hot-path kernels: error plumbing, launch accounting,
// vectorised complex64 access and warp/block reductions.
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace ocb {

// ---- error state (thread-local, read through ocb_last_error) ---------------------------
std::string& last_error();
int fail(const char* what, const char* file, int line);
int64_t& launch_counter();

#define OCB_CUDA(expr)                                                              \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            char _b[512];                                                           \
            snprintf(_b, sizeof _b, "%s -> %s", #expr, cudaGetErrorString(_e));     \
            return ::ocb::fail(_b, __FILE__, __LINE__);                             \
        }                                                                           \
    } while (0)

#define OCB_CUFFT(expr)                                                             \
    do {                                                                            \
        cufftResult _r = (expr);                                                    \
        if (_r != CUFFT_SUCCESS) {                                                  \
            char _b[512];                                                           \
            snprintf(_b, sizeof _b, "%s -> cufft error %d", #expr, (int)_r);        \
            return ::ocb::fail(_b, __FILE__, __LINE__);                             \
        }                                                                           \
    } while (0)

#define OCB_REQUIRE(cond, msg)                                                      \
    do {                                                                            \
        if (!(cond)) return ::ocb::fail(msg, __FILE__, __LINE__);                   \
    } while (0)

// Every kernel launch of the library goes through this so that bench.py can report how many
// of OUR kernels ran inside a timed region.
#define OCB_LAUNCH(kernel, grid, block, smem, stream, ...)                          \
    do {                                                                            \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                 \
        ::ocb::launch_counter()++;                                                  \
        OCB_CUDA(cudaGetLastError());                                               \
    } while (0)

// Launch with programmatic dependent launch allowed: the kernel may be scheduled while its predecessor
// in the stream drains; it must execute pdl_wait() before it touches anything the predecessor wrote.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                             bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#define OCB_LAUNCH_PDL(kernel, grid, block, smem, stream, pdl, ...)                  \
    do {                                                                            \
        ::ocb::launch_counter()++;                                                  \
        OCB_CUDA(::ocb::launch_ex(kernel, (grid), (block), (smem), (stream), (pdl), __VA_ARGS__)); \
    } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

inline int grid_for(int64_t work_items, int block, int per_thread, int max_waves = 8) {
    int64_t blocks = (work_items + (int64_t)block * per_thread - 1) / ((int64_t)block * per_thread);
    int64_t cap = (int64_t)kNumSMs * max_waves;
    if (blocks > cap) blocks = cap;  // grid-stride beyond this
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// ---- device helpers ------------------------------------------------------------------------
// Programmatic dependent launch: wait until the preceding kernel of the stream has completed and its
// writes are visible (no-op when the launch did not allow PDL) / let the next kernel be scheduled.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {  // a * conj(b)
    return make_float2(fmaf(a.x, b.x, a.y * b.y), fmaf(a.y, b.x, -a.x * b.y));
}
__device__ __forceinline__ float cabs2(float2 a) { return fmaf(a.x, a.x, a.y * a.y); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming 128-bit accesses (two complex64 per transaction)
__device__ __forceinline__ float4 ldg4(const float2* p) {
    float4 v;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
        : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg4(float2* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};"
                 ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Streaming accesses that do not allocate in L1: the field arrays are touched once per kernel, while
// the small twiddle tables must stay L1-resident (an L1 miss on a table load costs an L2 round trip in
// the middle of a dependent FFT chain).
__device__ __forceinline__ float2 ld_stream(const float2* p) {
    float2 v;
    asm("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_stream(const float* p) {
    float v;
    asm("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
// Same load, but ordered against the (volatile) st_stream stores: for buffers that a kernel updates in
// place (volatile asm statements keep their program order).
__device__ __forceinline__ float2 ld_stream_ordered(const float2* p) {
    float2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p) : "memory");
    return v;
}
// Same loads with a fixed position in the instruction stream (volatile, no memory clobber): used where a
// batch of loads must be issued together ahead of the code that consumes it.
__device__ __forceinline__ float2 ld_stream_pinned(const float2* p) {
    float2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_stream_pinned(const float* p) {
    float v;
    asm volatile("ld.global.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(float2* p, float2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream(float* p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// ---- mbarrier / bulk-copy primitives (PTX ISA: mbarrier, cp.async.bulk) ---------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra D_%=;\n"
        "bra W_%=;\n"
        "D_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (bytes: multiple of 16, both addresses 16-byte aligned), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// orders earlier generic-proxy accesses of shared memory before later async-proxy (bulk copy) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace ocb
