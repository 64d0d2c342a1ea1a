// common.cuh -- shared device/host helpers for liblccrf (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/lccrf.h"

namespace lccrf {

// ------------------------------------------------------------------ errors
void set_error(const std::string &msg);
int fail(int code, const std::string &msg);

#define LCCRF_CUDA(expr)                                                                          \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return ::lccrf::fail(LCCRF_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

#define LCCRF_TRY(expr)            \
    do {                           \
        int _rc = (expr);          \
        if (_rc != LCCRF_OK) return _rc; \
    } while (0)

// ------------------------------------------------------------------ launch geometry
constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs
constexpr int kThreads = 256;

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
// persistent / grid-stride grid: a multiple of the SM count (8 CTAs of 256 threads per SM = full occupancy)
static inline int persistent_grid(long long work_items, int threads = kThreads, int ctas_per_sm = 8) {
    long long need = (work_items + threads - 1) / threads;
    long long cap = (long long)kNumSMs * ctas_per_sm;
    if (need >= cap) return (int)cap;
    // round up to a multiple of the SM count when there is at least one wave
    if (need > kNumSMs) need = (need + kNumSMs - 1) / kNumSMs * kNumSMs;
    return (int)(need < 1 ? 1 : need);
}

#ifdef __CUDACC__
// ------------------------------------------------------------------ hash of a 64-bit key word
__device__ __forceinline__ uint32_t mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return (uint32_t)x;
}

// upper_bound(arr, n, x) - 1 : index b with arr[b] <= x < arr[b+1]  (arr ascending, arr[0] <= x)
__device__ __forceinline__ int find_segment(const int *__restrict__ arr, int n, int x) {
    int lo = 0, hi = n;  // invariant: arr[lo] <= x, answer in [lo, hi)
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(arr + mid) <= x) lo = mid;
        else hi = mid;
    }
    return lo;
}
#endif

}  // namespace lccrf
