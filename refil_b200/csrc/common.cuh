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

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------------------
// The step is a chain of short dependent kernels; with the launch attribute below kernel N+1 may be scheduled while kernel N is
// still running (on SMs N does not use, or as N's CTAs retire), run its prologue -- barrier init, TMEM allocation, the split of
// its resident weight tile, register-resident recurrent weights -- and only then block in pdl_wait() until N has completed and
// its writes are visible.  Rules kept by every kernel launched through refil_launch(..., pdl = true):
//   * nothing written by an earlier kernel of the step is read, and nothing is written to global memory, before pdl_wait();
//   * every thread executes pdl_wait() (so a kernel never completes before its predecessor: ordering stays transitive).
// MEASURED (B200, r2l): with one stream per network the early-scheduled CTAs sit on SMs that independent kernels of the other
// streams could have used -- 16-episode shard 1.11 -> 1.23 ms per step, full batch unchanged -- so the attribute is OFF by default;
// REFIL_PDL=1 in the environment turns it on (single-stream callers, where it only hides prologues).
bool refil_pdl_enabled();

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <class... KArgs, class... Args>
static inline cudaError_t refil_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                                       Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && refil_pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
