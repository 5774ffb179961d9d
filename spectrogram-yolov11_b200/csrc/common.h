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

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel: `cache` is the caller's
// `static size_t [kMaxDevices]` of what has been set so far (the attribute is raised, never lowered).
static constexpr int kMaxDevices = 64;
template <typename Kernel>
inline cudaError_t ensure_dynamic_smem(Kernel kernel, size_t bytes, size_t* cache) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    size_t& have = cache[dev % kMaxDevices];
    if (bytes > have) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e == cudaSuccess) have = bytes;
    }
    return e;
}

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

// Division by a launch-time constant without the ~100-cycle integer-divide sequence: the single-thread TMA / MMA
// issue loops are latency chains, so every runtime `/` or `%` in them costs more than the instruction they feed.
// q = (n * m) >> 40 with m = ceil(2^40 / d) is exact while n * d < 2^40 (checked by the launchers).
struct FastDiv {
    uint64_t m;
    uint32_t d;
};
static inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    f.m = ((1ull << 40) + d - 1) / d;
    return f;
}
static inline bool fastdiv_ok(uint64_t n_max, uint32_t d) { return n_max * (uint64_t)d < (1ull << 40); }

// Launch counter (bench.py reports gpu_launches from it).
void count_launch(int n = 1);

}  // namespace specyolo

#ifdef __CUDACC__
namespace specyolo {

__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
    return (uint32_t)(((uint64_t)n * f.m) >> 40);
}
// n -> (n / d, n % d)
__device__ __forceinline__ void fdivmod(uint32_t n, const FastDiv& f, uint32_t& q, uint32_t& r) {
    q = fdiv(n, f);
    r = n - q * f.d;
}

__device__ __forceinline__ float silu_f(float x) {
    // x * sigmoid(x); __expf error is far below bf16 output resolution
    return __fdividef(x, 1.0f + __expf(-x));
}
// SiLU with one MUFU op: x*sigmoid(x) = h + h*tanh(h), h = x/2 (tanh.approx error ~2^-11, below bf16 resolution)
__device__ __forceinline__ float silu_tanh(float x) {
    const float h = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
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

// Packed fp32 pairs (sm_100 FFMA2 / FADD2): one issue slot for two lanes of arithmetic — the conv epilogues and the
// Fusion kernels are issue-bound, not bandwidth-bound.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5};\n\t"
        "add.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
// SiLU of a pair given the accumulator pair a and the HALF-scaled bias pair hb: h = a/2 + hb, silu = h + h*tanh(h)
__device__ __forceinline__ float2 silu2_half(float2 a, float2 hb) {
    const float2 h = ffma2(a, make_float2(0.5f, 0.5f), hb);
    float2 t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(h.x));
    asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(h.y));
    return ffma2(h, t, h);
}

}  // namespace specyolo
#endif
