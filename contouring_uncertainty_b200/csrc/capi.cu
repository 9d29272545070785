// Error plumbing + version queries of the C ABI (include/cub200.h).
#include <cstdio>
#include <cstring>
#include "common.cuh"

#include <atomic>

namespace cub200 {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void note_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void set_last_error(const char* what, cudaError_t e) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}
int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);          // every launch site is followed by exactly one check
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_last_error(what, e);
        return CUB200_ECUDA;
    }
    return 0;
}
}  // namespace cub200

extern "C" const char* cub200_last_error(void) { return cub200::g_err; }
extern "C" int cub200_abi_version(void) { return 19; }
extern "C" int cub200_built_for_sm(void) { return 100; }
extern "C" int64_t cub200_launch_count(void) { return cub200::g_launches.load(std::memory_order_relaxed); }
