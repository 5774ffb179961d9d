// Shared host/device helpers for libspecyolo (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>

#include "../../include/specyolo.h"

namespace specyolo {

// Thread-local error text returned by specyolo_last_error().
void set_error(const char* fmt, ...);

#define SY_CHECK(cond, code, ...)                                   \
    do {                                                            \
        if (!(cond)) {                                              \
            ::specyolo::set_error(__VA_ARGS__);                     \
            return (code);                                          \
        }                                                           \
    } while (0)

#define SY_CUDA(expr)                                                                   \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            ::specyolo::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                  __FILE__, __LINE__);                                  \
            return SPECYOLO_ERR_CUDA;                                                   \
        }                                                                               \
    } while (0)

// Launch check used after every kernel launch (does not synchronise).
#define SY_LAUNCH_CHECK()                                                               \
    do {                                                                                \
        cudaError_t _e = cudaGetLastError();                                            \
        if (_e != cudaSuccess) {                                                        \
            ::specyolo::set_error("kernel launch failed: %s (%s:%d)",                   \
                                  cudaGetErrorString(_e), __FILE__, __LINE__);          \
            return SPECYOLO_ERR_CUDA;                                                   \
        }                                                                               \
    } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Launch counter (bench.py reports gpu_launches from it).
void count_launch(int n = 1);

}  // namespace specyolo

#ifdef __CUDACC__
namespace specyolo {

__device__ __forceinline__ float silu_f(float x) {
    // x * sigmoid(x); __expf error is far below bf16 output resolution
    return __fdividef(x, 1.0f + __expf(-x));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + __expf(-x)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}

}  // namespace specyolo
#endif
