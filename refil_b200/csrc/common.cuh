// Shared helpers for the refil_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/refil_b200.h"  // the C ABI: every extern "C" definition is checked against it

void refil_set_error(const char* fmt, ...);

#define REFIL_CHECK_ARG(cond, ...)                 \
    do {                                           \
        if (!(cond)) {                             \
            refil_set_error(__VA_ARGS__);          \
            return REFIL_ERR_ARG;                  \
        }                                          \
    } while (0)

#define REFIL_CHECK_LAUNCH(name)                                                    \
    do {                                                                            \
        cudaError_t e__ = cudaGetLastError();                                       \
        if (e__ != cudaSuccess) {                                                   \
            refil_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
            return REFIL_ERR_CUDA;                                                  \
        }                                                                           \
    } while (0)

static inline int refil_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// number of SMs on the current device (cached); B200 = 148
int refil_num_sms();
