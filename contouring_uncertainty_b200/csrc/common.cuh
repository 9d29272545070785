// Shared helpers for the cub200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/cub200.h"

namespace cub200 {

void set_last_error(const char* what, cudaError_t e);
int check_launch(const char* what);     // after a kernel launch: error check; counts ONE launch (cub200_launch_count)
void note_launches(int n);              // further launches of the same entry point that share one check

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

// ---------------------------------------------------------------- Philox4x32-10 + Box-Muller
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
    // (k + 1/2) 2^-23 with a 23-bit k is an exact fp32 number in [2^-24, 1 - 2^-24]; a 24-bit k + 1/2 would round up to 1
    const float u1 = ((float)(a >> 9) + 0.5f) * (1.0f / 8388608.0f);
    const float u2 = ((float)(b >> 9) + 0.5f) * (1.0f / 8388608.0f);
    const float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(2.0f * u2, &s, &c);
    n0 = r * c;
    n1 = r * s;
}

// The draws of global sample gc (Philox counter (gc, q), key = seed): 4 normals per call, z[4q .. 4q+3].
__device__ __forceinline__ void philox_normals4(uint64_t gc, int q, uint64_t seed, float& n0, float& n1, float& n2, float& n3) {
    uint32_t r[4];
    philox4x32_10((uint32_t)gc, (uint32_t)(gc >> 32), (uint32_t)q, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    box_muller(r[0], r[1], n0, n1);
    box_muller(r[2], r[3], n2, n3);
}

}  // namespace cub200
