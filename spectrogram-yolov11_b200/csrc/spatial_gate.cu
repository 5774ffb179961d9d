// SobelSpatialAttention (ultralytics/nn/modules/conv.py:1184-1198), the gate that ends ConvHCA (conv.py:829-844) in
// the *_convHCA sibling configs:
//
//   mm   = cat(mean_c x, max_c x)                         [B, 2, H, W]
//   a    = cv1( sum_k sobel_k (*) mm )                     three depthwise 3x3 convs (groups = 2, zero padding), summed,
//                                                          then a 2 -> 1 1x1 conv: LINEAR in mm, so the caller folds
//                                                          the seven small weights into one 2 x 3 x 3 stencil w
//   y    = x * sigmoid(a)
//
// Two HBM-bound passes over the bf16 NHWC activation (the stencil needs the neighbours' statistics before any pixel
// can be gated): chan_meanmax_kernel reads x once and writes the fp32 mean / max planes; sobel_gate_kernel reads the
// planes (L2-resident: 8 bytes per pixel) and x, and writes y (y may alias x).  A group of LPP lanes owns a pixel, a
// lane a 16-byte channel vector; four pixels per group are in flight per iteration.
#include "common.h"
#include "tma_host.h"

namespace specyolo {

struct GateParams {
    specyolo_spatial_gate_t a;
    long npix;            // B * H * W
    int vpp;              // 16-byte vectors per pixel = C / 8
    float inv_c;
};

__device__ __forceinline__ uint4 gate_ldg_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float2 gate_bf2(uint32_t w) {
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

static constexpr int kGateThreads = 256;
static constexpr int kGateUnroll = 4;

// LPP lanes per pixel (power of two <= 32); a lane walks vectors lane_in_group, +LPP, ... of its pixel.
template <int LPP>
__global__ void __launch_bounds__(kGateThreads)
chan_meanmax_kernel(const __grid_constant__ GateParams p) {
    const specyolo_spatial_gate_t& a = p.a;
    constexpr int GPB = kGateThreads / LPP;                        // pixel groups per block
    const int g = threadIdx.x / LPP, l = threadIdx.x % LPP;
    const long HW = (long)a.H * a.W;
    const long step = (long)gridDim.x * GPB * kGateUnroll;
    for (long base = ((long)blockIdx.x * GPB + g) * kGateUnroll; ; base += step) {
        if (base - (long)g * kGateUnroll >= p.npix) break;           // block-uniform exit
        float s[kGateUnroll], m[kGateUnroll];
#pragma unroll
        for (int u = 0; u < kGateUnroll; ++u) { s[u] = 0.f; m[u] = -INFINITY; }
        for (int v = l; v < p.vpp; v += LPP) {
            uint4 q[kGateUnroll];
#pragma unroll
            for (int u = 0; u < kGateUnroll; ++u) {
                const long pix = base + u;
                q[u] = pix < p.npix ? gate_ldg_v4(reinterpret_cast<const __nv_bfloat16*>(a.x) + (size_t)pix * a.x_pixstride + v * 8)
                                    : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < kGateUnroll; ++u) {
                const uint32_t w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = gate_bf2(w[j]);
                    s[u] += f.x + f.y;
                    m[u] = fmaxf(m[u], fmaxf(f.x, f.y));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kGateUnroll; ++u) {
#pragma unroll
            for (int d = LPP >> 1; d > 0; d >>= 1) {
                s[u] += __shfl_xor_sync(0xffffffffu, s[u], d);
                m[u] = fmaxf(m[u], __shfl_xor_sync(0xffffffffu, m[u], d));
            }
            const long pix = base + u;
            if (l == 0 && pix < p.npix) {
                const long b = pix / HW, r = pix - b * HW;
                a.mm[(b * 2) * HW + r] = s[u] * p.inv_c;
                a.mm[(b * 2 + 1) * HW + r] = m[u];
            }
        }
    }
}

template <int LPP>
__global__ void __launch_bounds__(kGateThreads)
sobel_gate_kernel(const __grid_constant__ GateParams p) {
    const specyolo_spatial_gate_t& a = p.a;
    constexpr int GPB = kGateThreads / LPP;
    const int g = threadIdx.x / LPP, l = threadIdx.x % LPP;
    const long HW = (long)a.H * a.W;
    const long step = (long)gridDim.x * GPB * kGateUnroll;
    for (long base = ((long)blockIdx.x * GPB + g) * kGateUnroll; ; base += step) {
        if (base - (long)g * kGateUnroll >= p.npix) break;           // block-uniform exit
        float gate[kGateUnroll];
#pragma unroll
        for (int u = 0; u < kGateUnroll; ++u) {
            const long pix = min(base + u, p.npix - 1);
            const long b = pix / HW, r = pix - b * HW;
            const int y0 = (int)(r / a.W), x0 = (int)(r - (long)y0 * a.W);
            float acc = 0.f;
            for (int t = l; t < 18; t += LPP) {                       // tap t = c * 9 + ky * 3 + kx
                const int c = t / 9, k = t - c * 9, ky = k / 3, kx = k - ky * 3;
                const int yy = y0 + ky - 1, xx = x0 + kx - 1;
                if (yy >= 0 && yy < a.H && xx >= 0 && xx < a.W)
                    acc = fmaf(a.w[t], __ldg(a.mm + (b * 2 + c) * HW + (long)yy * a.W + xx), acc);
            }
#pragma unroll
            for (int d = LPP >> 1; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
            gate[u] = 1.0f / (1.0f + __expf(-acc));
        }
        for (int v = l; v < p.vpp; v += LPP) {
            uint4 q[kGateUnroll];
#pragma unroll
            for (int u = 0; u < kGateUnroll; ++u) {
                const long pix = base + u;
                q[u] = pix < p.npix ? gate_ldg_v4(reinterpret_cast<const __nv_bfloat16*>(a.x) + (size_t)pix * a.x_pixstride + v * 8)
                                    : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < kGateUnroll; ++u) {
                const long pix = base + u;
                if (pix >= p.npix) continue;
                uint32_t w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = gate_bf2(w[j]);
                    const __nv_bfloat162 o = __floats2bfloat162_rn(f.x * gate[u], f.y * gate[u]);
                    w[j] = *reinterpret_cast<const uint32_t*>(&o);
                }
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.y) + (size_t)pix * a.y_pixstride + v * 8) =
                    make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }
}

template <int LPP>
static void gate_launch_t(const GateParams& p, cudaStream_t stream) {
    constexpr int GPB = kGateThreads / LPP;
    const long groups = (p.npix + kGateUnroll - 1) / kGateUnroll;
    long blocks = (groups + GPB - 1) / GPB;
    const long cap = (long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    chan_meanmax_kernel<LPP><<<(unsigned)blocks, kGateThreads, 0, stream>>>(p);
    count_launch();
    sobel_gate_kernel<LPP><<<(unsigned)blocks, kGateThreads, 0, stream>>>(p);
    count_launch();
}

int spatial_gate_launch(const specyolo_spatial_gate_t* a, cudaStream_t stream) {
    SY_CHECK(a->C % 8 == 0 && a->C >= 8, SPECYOLO_ERR_UNSUPPORTED, "spatial gate: C must be a multiple of 8 (got %d)", a->C);
    SY_CHECK(a->x_pixstride % 8 == 0 && a->y_pixstride % 8 == 0 && a->x_pixstride >= a->C && a->y_pixstride >= a->C,
             SPECYOLO_ERR_INVALID, "spatial gate: pixel strides must be multiples of 8 elements and >= C");
    SY_CHECK(((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->y)) & 15) == 0, SPECYOLO_ERR_INVALID,
             "spatial gate: x / y must be 16-byte aligned");
    GateParams p{};
    p.a = *a;
    p.npix = (long)a->B * a->H * a->W;
    p.vpp = a->C / 8;
    p.inv_c = 1.0f / (float)a->C;
    if (p.vpp >= 32) gate_launch_t<32>(p, stream);
    else if (p.vpp >= 16) gate_launch_t<16>(p, stream);
    else gate_launch_t<8>(p, stream);
    SY_LAUNCH_CHECK();
    return SPECYOLO_OK;
}

}  // namespace specyolo
