// Shared helpers for the nvo_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include "nvo_b200.h"

void nvo_set_error(const char* fmt, ...);

#define NVO_CHECK(cond, ...)                    \
    do {                                        \
        if (!(cond)) {                          \
            nvo_set_error(__VA_ARGS__);         \
            return 1;                           \
        }                                       \
    } while (0)

void nvo_count_launch();

#define NVO_CUDA_LAUNCH_CHECK(name)                                                   \
    do {                                                                              \
        nvo_count_launch();                                                           \
        cudaError_t e__ = cudaGetLastError();                                         \
        if (e__ != cudaSuccess) {                                                     \
            nvo_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
            return 2;                                                                 \
        }                                                                             \
    } while (0)

static inline unsigned int nvo_blocks(int64_t work, int threads) { return (unsigned int)((work + threads - 1) / threads); }

// experiment switches: callers keep the result in a function-local static, i.e. one getenv per process
int nvo_env_int(const char* name, int fallback);

// 148 SMs on B200; persistent kernels size their grid from this (queried once)
int nvo_sm_count();

__device__ __forceinline__ float nvo_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double nvo_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// inclusive warp scan
template <typename T>
__device__ __forceinline__ T nvo_warp_scan_incl(T v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

__device__ __forceinline__ void nvo_red_add_v2(float* addr, float a, float b) {
    // vectorised 8-byte reduction (sm_90+): one L2 atomic transaction for both features of a table row
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
