// Shared helpers for the distar_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/distar_b200.h"

namespace dsb {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return DSB_ERR_CUDA;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return DSB_OK;
}

#define DSB_REQUIRE(cond, ...)              \
    do {                                    \
        if (!(cond)) {                      \
            dsb::set_error(__VA_ARGS__);    \
            return DSB_ERR_ARG;             \
        }                                   \
    } while (0)

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

}  // namespace dsb
