// Small per-process cache of cuFFT plans for the Rx-DSP entry points (EDC / firFilter overlap-save blocks, the FOE
// spectrum of cpr): creating a plan costs a device allocation and milliseconds of host time, and these entry points
// are called once per stage of a receiver chain with the same geometry again and again.  Key = (device, stream,
// transform type, length, batch); least-recently-used eviction.  The stream is part of the key because a cuFFT plan
// owns ONE work area: host threads that drive independent units on their own streams (sharding.run_concurrent) must
// never execute the same plan at the same time.
#pragma once
#include <cufft.h>
#include <cuda_runtime.h>

#include <mutex>

namespace ocb {

struct FftPlanEntry {
    int dev = -1, type = 0, n = 0, batch = 0;
    cudaStream_t st = nullptr;
    cufftHandle h = 0;
    unsigned long long stamp = 0;
    bool used = false;
};

// returns CUFFT_SUCCESS and a plan bound to `st`
inline cufftResult fft_plan_cached(cufftType type, int n, int batch, cudaStream_t st, cufftHandle* out) {
    constexpr int kSlots = 64;
    static FftPlanEntry slots[kSlots];
    static unsigned long long clock = 0;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    int dev = 0;
    cudaGetDevice(&dev);
    int victim = 0;
    for (int i = 0; i < kSlots; ++i) {
        FftPlanEntry& e = slots[i];
        if (e.used && e.dev == dev && e.st == st && e.type == (int)type && e.n == n && e.batch == batch) {
            e.stamp = ++clock;
            *out = e.h;
            return CUFFT_SUCCESS;
        }
        if (!e.used) victim = i;
        else if (slots[victim].used && e.stamp < slots[victim].stamp) victim = i;
    }
    FftPlanEntry& v = slots[victim];
    if (v.used) { cufftDestroy(v.h); v.used = false; }
    int len[1] = {n};
    cufftResult r = cufftPlanMany(&v.h, 1, len, nullptr, 1, n, nullptr, 1, n, type, batch);
    if (r != CUFFT_SUCCESS) return r;
    v.dev = dev; v.st = st; v.type = (int)type; v.n = n; v.batch = batch; v.stamp = ++clock; v.used = true;
    *out = v.h;
    return cufftSetStream(v.h, st);
}

}  // namespace ocb
