// Library-level plumbing: version, error text, launch counter, SM-count cache.
#include <atomic>
#include <cstdio>
#include <cstring>

#include "dm_common.cuh"

namespace dm {

static thread_local char g_err[512] = {0};
static std::atomic<int64_t> g_launches{0};

void set_cuda_error(cudaError_t e, const char* where) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorString(e),
             cudaGetErrorName(e));
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
    // Immutable hardware facts cached per device; not mutable state in the API sense.
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int v = cache[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
            v = 148;
        cache[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

}  // namespace dm

extern "C" {

int dm_version(void) { return 100; }

const char* dm_error_string(int code) {
    switch (code) {
        case DM_OK: return "ok";
        case DM_EINVAL: return "invalid argument";
        case DM_ECUDA: return "CUDA error";
        case DM_EUNSUPPORTED: return "unsupported configuration";
        default: return "unknown error";
    }
}

const char* dm_last_cuda_error(void) { return dm::g_err; }

int64_t dm_launch_count(void) { return dm::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
