// Shared helpers for the cub200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/cub200.h"

namespace cub200 {

void set_last_error(const char* what, cudaError_t e);
int check_launch(const char* what);

#define CUB200_CUDA_OK(expr)                                   \
    do {                                                       \
        cudaError_t _e = (expr);                               \
        if (_e != cudaSuccess) {                               \
            cub200::set_last_error(#expr, _e);                 \
            return CUB200_ECUDA;                               \
        }                                                      \
    } while (0)

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// streaming 128-bit load that does not allocate in L1 (each logit is touched exactly once)
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// order-preserving float -> int map (for redux.sync max on floats)
__device__ __forceinline__ int float_ordered(float f) {
    int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ordered_float(int i) {
    return __int_as_float(i ^ ((i >> 31) & 0x7fffffff));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

}  // namespace cub200
