// Frequency pass of the fused four-step engine for N2 = 1024 (N = 2^20), fed by the Tensor Memory Accelerator.
//
// Same arithmetic as k_freq<32, C> (fused_kernels.cuh): for a tile of C = 4 adjacent W positions p (all 1024 rows of
// one polarisation) FFT_N2 over n2 -> x linear operator -> IFFT_N2, in place.  The data path is different:
//
//  * The W buffer is described once per plan by a CUtensorMap over float32 [pol][n2][2*N1], kept in GLOBAL memory (the
//    plan's workspace) so that its address — the key of the TMA unit's descriptor cache — is the same in every launch
//    (a __grid_constant__ kernel parameter lives at a new address per launch: measured 4 % slower in the step loop).  A tile is a strided
//    1024 x 32-byte column block: four `cp.async.bulk.tensor.3d` box loads (8 floats x 256 rows) bring it into
//    shared memory behind one mbarrier, and four tensor stores write the result back — the per-thread strided
//    ld.global / st.global of the classic kernel (32 + 32 LSU instructions per thread, the floor of its ablation) are
//    gone.  Every row segment is exactly one 32-byte sector.
//  * One 512-thread CTA = 4 groups of 128 threads, one tile each (tile = group * gridDim.x + blockIdx.x; the grid is
//    tiles / 4 CTAs, so every group has a tile).
//  * The tile's landing buffer (32 KB) becomes the exchange buffer of the cooperative transform once the tile is in
//    registers, and the staging buffer of the tensor store at the end.
//  * The tile's slice of the operator table (32 KB, stored in consumption order) streams through a two-slot ring of
//    8 KB bulk copies; the first two are issued BEFORE griddepcontrol.wait together with the twiddle table staging.
//
// Shared memory per CTA: 4 x (32 KB tile + 16 KB operator ring) + 16 KB twiddle table (hi + lo) = 208 KB.
#pragma once
#include <cuda.h>

#include "fused_time_bulk.cuh"

namespace ocb {

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, int x, int y, int z, const void* smem_src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(smem_src)) : "memory");
}
__device__ __forceinline__ void tma_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

struct FreqTmaCfg {
    static constexpr int Q2 = 32, C = 4, NT = Q2 * C, GROUPS = 4, ROWS = 1024, BOX_ROWS = 256;
    static constexpr int TILE_BYTES = ROWS * C * 8;                 // 32 KB
    static constexpr int LP_CHUNK_SLOTS = 8, LP_CHUNKS = 32 / LP_CHUNK_SLOTS;
    static constexpr int LP_CHUNK_BYTES = LP_CHUNK_SLOTS * NT * 8;  // 8 KB
    static constexpr int TW_HALF = 1024;
    static constexpr int TW_ENTRIES = (fft::kLO ? 2 : 1) * TW_HALF;
    static constexpr int OFF_TW = 0;
    static constexpr int OFF_TILE = OFF_TW + TW_ENTRIES * 8;
    static constexpr int OFF_LP = OFF_TILE + GROUPS * TILE_BYTES;
    static constexpr int OFF_BAR = OFF_LP + GROUPS * 2 * LP_CHUNK_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + GROUPS * 4 * 8;
    static constexpr int LP_ENTRIES = 32 * Q2 * C;                  // operator entries per tile
};

template <bool LOCKSTEP, bool STORE_TMA>
__global__ void __launch_bounds__(512, 1)
k_freq_tma(const CUtensorMap* __restrict__ wmap_ptr, float2* __restrict__ W, const float2* __restrict__ LP,
           const float2* __restrict__ tw,
           int N1, int NP, const long long* __restrict__ converged_step, long long step_id,
           const long long* __restrict__ need_flag, long long need_id) {
    using namespace fft;
    using Cfg = FreqTmaCfg;
    constexpr int Q2 = Cfg::Q2, C = Cfg::C, NT = Cfg::NT;
    extern __shared__ __align__(128) unsigned char smem[];
    float2* tws = reinterpret_cast<float2*>(smem + Cfg::OFF_TW);
    const float2* tws_lo = tws + Cfg::TW_HALF;
    const int tid = threadIdx.x, g = tid / NT, tg = tid % NT, q = tg / C, c = tg % C;
    unsigned char* tile_buf = smem + Cfg::OFF_TILE + g * Cfg::TILE_BYTES;
    float2* tile2 = reinterpret_cast<float2*>(tile_buf);       // landing / staging view: [row][C] float2
    float* xr = reinterpret_cast<float*>(tile_buf);            // exchange view (one plane, two rounds)
    unsigned char* lp_ring = smem + Cfg::OFF_LP + g * 2 * Cfg::LP_CHUNK_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR) + g * 4;
    uint64_t* dbar = bars;          // tile landed
    uint64_t* lbar = bars + 1;      // [2] operator ring slots
    // LOCKSTEP: the four groups meet at CTA-wide barriers; otherwise each group synchronises on its own named barrier,
    // so a group starts transforming as soon as ITS tile has landed and the groups' load / compute / store phases
    // overlap (OCB_FREQ_LOCKSTEP=1 selects the former at launch; A/B knob)
    auto gsync = [g] {
        if constexpr (LOCKSTEP) __syncthreads();
        else named_bar_sync(1 + g, NT);
    };

    const int tiles_per_pol = N1 / C;
    const int tile_id = g * gridDim.x + blockIdx.x;
    const bool has_tile = tile_id < NP * tiles_per_pol;
    const int pol = tile_id / tiles_per_pol, tile = tile_id % tiles_per_pol;
    const char* lp_src = reinterpret_cast<const char*>(LP + (int64_t)tile * Cfg::LP_ENTRIES);

    // ---- prologue (independent of the preceding kernel): barriers, operator chunks 0 and 1, twiddle table ----------
    if (tg == 0) {
        mbar_init(dbar, 1);
        mbar_init(lbar, 1);
        mbar_init(lbar + 1, 1);
        mbar_fence_init();
        if (has_tile) {
            tma_prefetch_desc(wmap_ptr);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                mbar_expect_tx(lbar + j, Cfg::LP_CHUNK_BYTES);
                bulk_g2s(lp_ring + j * Cfg::LP_CHUNK_BYTES, lp_src + j * Cfg::LP_CHUNK_BYTES, Cfg::LP_CHUNK_BYTES, lbar + j);
            }
        }
    }
    for (int i = tid; i < Cfg::TW_HALF; i += 512) {
        tws[i] = __ldg(tw + i);
        if constexpr (kLO) tws[Cfg::TW_HALF + i] = __ldg(tw + 64 * Q2 + i);
    }
    __syncthreads();

    pdl_wait();  // W was written by the preceding time pass (programmatic dependent launch)
    pdl_launch_dependents();
    const bool skip = (converged_step && *reinterpret_cast<const volatile long long*>(converged_step) == step_id) ||
                      (need_flag && *reinterpret_cast<const volatile long long*>(need_flag) != need_id);
    if (skip) {  // let the two operator chunks land before the shared memory goes away
        mbar_wait(lbar, 0);
        mbar_wait(lbar + 1, 0);
        return;
    }

    // ---- tile in: four box loads (8 floats x 256 rows each) behind one barrier ----------------------------------------
    if (tg == 0) {
        mbar_expect_tx(dbar, Cfg::TILE_BYTES);
#pragma unroll
        for (int j = 0; j < Cfg::ROWS / Cfg::BOX_ROWS; ++j)
            tma_load_3d(tile_buf + j * (Cfg::BOX_ROWS * C * 8), wmap_ptr, 2 * tile * C, j * Cfg::BOX_ROWS, pol, dbar);
    }
    mbar_wait(dbar, 0);
    float2 v[32];
#pragma unroll
    for (int a = 0; a < 32; ++a) v[a] = tile2[(Q2 * a + q) * C + c];
    gsync();  // the landing buffer becomes the exchange buffer

    coop_fft_forward<Q2, C, C, true, false>(v, xr, nullptr, tws, tws_lo, q, c, gsync);

    // ---- x linear operator: chunk j of the slice sits in ring slot j % 2 ---------------------------------------------
#pragma unroll
    for (int j = 0; j < Cfg::LP_CHUNKS; ++j) {
        const float2* lps = reinterpret_cast<const float2*>(lp_ring + (j & 1) * Cfg::LP_CHUNK_BYTES);
        mbar_wait(lbar + (j & 1), (unsigned)(j >> 1));
#pragma unroll
        for (int s = 0; s < Cfg::LP_CHUNK_SLOTS; ++s)
            v[j * Cfg::LP_CHUNK_SLOTS + s] = cmul(v[j * Cfg::LP_CHUNK_SLOTS + s], lps[s * NT + tg]);
        if (j + 2 < Cfg::LP_CHUNKS) {
            gsync();  // every thread of the group has read the slot
            if (tg == 0) {
                mbar_expect_tx(lbar + (j & 1), Cfg::LP_CHUNK_BYTES);
                bulk_g2s(lp_ring + (j & 1) * Cfg::LP_CHUNK_BYTES, lp_src + (j + 2) * Cfg::LP_CHUNK_BYTES,
                         Cfg::LP_CHUNK_BYTES, lbar + (j & 1));
            }
        }
    }
    gsync();  // (the forward transform's last exchange reads are long done; keep the groups' phases aligned per tile)

    coop_fft_inverse<Q2, C, C, true, false>(v, xr, nullptr, tws, tws_lo, q, c, gsync);

    // ---- tile out ------------------------------------------------------------------------------------------------------
    if constexpr (STORE_TMA) {
        // stage in the landing layout, four tensor stores
        gsync();  // last exchange reads done before the buffer is overwritten
#pragma unroll
        for (int a = 0; a < 32; ++a) tile2[(Q2 * a + q) * C + c] = v[a];
        fence_proxy_async();  // generic-proxy writes -> visible to the bulk-copy engine
        gsync();
        if (tg == 0) {
#pragma unroll
            for (int j = 0; j < Cfg::ROWS / Cfg::BOX_ROWS; ++j)
                tma_store_3d(wmap_ptr, 2 * tile * C, j * Cfg::BOX_ROWS, pol, tile_buf + j * (Cfg::BOX_ROWS * C * 8));
            tma_commit_group();
            tma_wait_group_read0();  // shared memory may be released once the engine has read it
        }
    } else {
        // per-thread streaming stores (one 32-byte sector per 4 lanes), issued as each thread finishes its transform:
        // the write traffic overlaps the other groups' arithmetic instead of arriving in one burst at the end
        float2* base = W + ((int64_t)pol * (32 * Q2)) * N1 + tile * C + c;
#pragma unroll
        for (int a = 0; a < 32; ++a) st_stream(base + (int64_t)(Q2 * a + q) * N1, v[a]);
    }
}

}  // namespace ocb
