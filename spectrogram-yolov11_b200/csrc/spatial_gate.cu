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


// ---------------------------------------------------------------------------------------------------------------------
// MSCSpatialAttention (ultralytics/nn/modules/conv.py:1200-1243), the inner block of C3x (block.py:522-529) in the
// *_OMN sibling config:
//
//   mm  = cat(mean_c x, max_c x)
//   s   = relu(conv31x31(mm)) + relu(conv3x3(mm))           two 2 -> 1 convs, zero padding, no bias     [B, 1, H, W]
//   g   = relu(fc(mean_hw(x * s)))                          x8 == x9 in the reference                    [B, C]
//   y   = x * s * g + x
//
// Four launches: chan_meanmax_kernel (above), msc_smap_kernel (the two small-channel convs on CUDA cores: register-
// blocked 4 outputs per thread, weights broadcast from shared memory), msc_pool_kernel (per-image, per-chunk partial
// sums of x * s in a fixed order: deterministic), msc_apply_kernel (every CTA folds the partials of its image and
// evaluates the C x C fc itself, then streams x once more).  x is read three times and y written once.
// ---------------------------------------------------------------------------------------------------------------------
static constexpr int kMscK = 31, kMscR = kMscK / 2;
static constexpr int kMscTW = 32, kMscTH = 16;                 // outputs per CTA
static constexpr int kMscPW = 64;                              // floats per row of the padded statistic tile (>= TW + 2R)
static constexpr int kMscRows = kMscTH + 2 * kMscR;

struct MscParams {
    specyolo_msc_gate_t a;
    float* mm;            // [B][2][H][W]
    float* smap;          // [B][H][W]
    float* partial;       // [B][nchunk][C]
    int nchunk;
    int vpp;
    float inv_hw;
};

__global__ void __launch_bounds__(128)
msc_smap_kernel(const __grid_constant__ MscParams p) {
    const specyolo_msc_gate_t& a = p.a;
    __shared__ __align__(16) float s_mm[2][kMscRows][kMscPW];
    __shared__ __align__(16) float s_w[2][kMscK][32];
    __shared__ float s_w3[18];
    const int tid = threadIdx.x, b = blockIdx.z;
    const int x0 = blockIdx.x * kMscTW, y0 = blockIdx.y * kMscTH;
    const long HW = (long)a.H * a.W;
    for (int i = tid; i < 2 * kMscRows * kMscPW; i += 128) {
        const int c = i / (kMscRows * kMscPW), r = i - c * (kMscRows * kMscPW), yy = r / kMscPW, xx = r - yy * kMscPW;
        const int gy = y0 - kMscR + yy, gx = x0 - kMscR + xx;
        (&s_mm[0][0][0])[i] = (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) ? __ldg(p.mm + ((long)b * 2 + c) * HW + (long)gy * a.W + gx) : 0.f;
    }
    for (int i = tid; i < 2 * kMscK * 32; i += 128) {
        const int kx = i & 31, r = i >> 5;                              // r = c * 31 + ky
        (&s_w[0][0][0])[i] = kx < kMscK ? __ldg(a.w_big + r * kMscK + kx) : 0.f;
    }
    if (tid < 18) s_w3[tid] = __ldg(a.w_small + tid);
    __syncthreads();

    const int tx = tid & 7, ty = tid >> 3;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
#pragma unroll 1
        for (int ky = 0; ky < kMscK; ++ky) {
            float d[36], w[32];
            const float4* row = reinterpret_cast<const float4*>(&s_mm[c][ty + ky][4 * tx]);
            const float4* wr = reinterpret_cast<const float4*>(&s_w[c][ky][0]);
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                const float4 v = row[j];
                d[4 * j] = v.x; d[4 * j + 1] = v.y; d[4 * j + 2] = v.z; d[4 * j + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 v = wr[j];
                w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
            }
#pragma unroll
            for (int kx = 0; kx < kMscK; ++kx) {
#pragma unroll
                for (int o = 0; o < 4; ++o) acc[o] = fmaf(w[kx], d[o + kx], acc[o]);
            }
        }
    }
    float acc3[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float w = s_w3[c * 9 + ky * 3 + kx];
#pragma unroll
                for (int o = 0; o < 4; ++o) acc3[o] = fmaf(w, s_mm[c][ty + kMscR + ky - 1][4 * tx + kMscR + kx - 1 + o], acc3[o]);
            }
    const int gy = y0 + ty;
    if (gy < a.H) {
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int gx = x0 + 4 * tx + o;
            if (gx < a.W) p.smap[(long)b * HW + (long)gy * a.W + gx] = fmaxf(acc[o], 0.f) + fmaxf(acc3[o], 0.f);
        }
    }
}

static constexpr int kMscMaxV = 2;     // 16-byte vectors per lane: C <= 32 * 2 * 8 = 512

// partial[b][chunk][c] = sum over the chunk's pixels of x[b, pix, c] * s[b, pix]
template <int LPP>
__global__ void __launch_bounds__(kGateThreads)
msc_pool_kernel(const __grid_constant__ MscParams p) {
    const specyolo_msc_gate_t& a = p.a;
    constexpr int GPB = kGateThreads / LPP;
    extern __shared__ float s_red[];                                   // [GPB][C]
    const int g = threadIdx.x / LPP, l = threadIdx.x % LPP;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int HW = a.H * a.W;
    const int per = (HW + p.nchunk - 1) / p.nchunk;
    const int lo = chunk * per, hi = min(HW, lo + per);
    float acc[kMscMaxV][8];
#pragma unroll
    for (int j = 0; j < kMscMaxV; ++j)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[j][e] = 0.f;
    const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(a.x) + (size_t)b * HW * a.x_pixstride;
    const float* sb = p.smap + (size_t)b * HW;
    for (int base = lo + g * kGateUnroll; base < hi; base += GPB * kGateUnroll) {
#pragma unroll
        for (int j = 0; j < kMscMaxV; ++j) {
            const int v = l + j * LPP;
            if (v >= p.vpp) break;
            uint4 q[kGateUnroll];
            float sv[kGateUnroll];
#pragma unroll
            for (int u = 0; u < kGateUnroll; ++u) {
                const int pix = base + u;
                const bool ok = pix < hi;
                q[u] = ok ? gate_ldg_v4(xb + (size_t)pix * a.x_pixstride + v * 8) : make_uint4(0, 0, 0, 0);
                sv[u] = ok ? __ldg(sb + pix) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < kGateUnroll; ++u) {
                const uint32_t w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = gate_bf2(w[e]);
                    acc[j][2 * e] = fmaf(f.x, sv[u], acc[j][2 * e]);
                    acc[j][2 * e + 1] = fmaf(f.y, sv[u], acc[j][2 * e + 1]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kMscMaxV; ++j) {
        const int v = l + j * LPP;
        if (v < p.vpp)
#pragma unroll
            for (int e = 0; e < 8; ++e) s_red[g * a.C + v * 8 + e] = acc[j][e];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < a.C; c += kGateThreads) {
        float t = 0.f;
        for (int gg = 0; gg < GPB; ++gg) t += s_red[gg * a.C + c];     // fixed order
        p.partial[((size_t)b * p.nchunk + chunk) * a.C + c] = t;
    }
}

template <int LPP>
__global__ void __launch_bounds__(kGateThreads)
msc_apply_kernel(const __grid_constant__ MscParams p) {
    const specyolo_msc_gate_t& a = p.a;
    constexpr int GPB = kGateThreads / LPP;
    extern __shared__ float s_g[];                                     // [C] pooled, then [C] gate
    float* s_pool = s_g + a.C;
    const int g = threadIdx.x / LPP, l = threadIdx.x % LPP;
    const int b = blockIdx.y;
    const int HW = a.H * a.W;
    for (int c = threadIdx.x; c < a.C; c += kGateThreads) {
        float t = 0.f;
        for (int ch = 0; ch < p.nchunk; ++ch) t += p.partial[((size_t)b * p.nchunk + ch) * a.C + c];
        s_pool[c] = t * p.inv_hw;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < a.C; c += kGateThreads) {
        float t = __ldg(a.fc_b + c);
        const float* wr = a.fc_w + (size_t)c * a.C;
        for (int j = 0; j < a.C; ++j) t = fmaf(__ldg(wr + j), s_pool[j], t);
        s_g[c] = fmaxf(t, 0.f);
    }
    __syncthreads();
    float gate[kMscMaxV][8];
#pragma unroll
    for (int j = 0; j < kMscMaxV; ++j) {
        const int v = l + j * LPP;
#pragma unroll
        for (int e = 0; e < 8; ++e) gate[j][e] = v < p.vpp ? s_g[v * 8 + e] : 0.f;
    }
    const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(a.x) + (size_t)b * HW * a.x_pixstride;
    __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(a.y) + (size_t)b * HW * a.y_pixstride;
    const float* sb = p.smap + (size_t)b * HW;
    for (int base = (blockIdx.x * GPB + g) * kGateUnroll; base < HW; base += gridDim.x * GPB * kGateUnroll) {
#pragma unroll
        for (int j = 0; j < kMscMaxV; ++j) {
            const int v = l + j * LPP;
            if (v >= p.vpp) break;
            uint4 q[kGateUnroll];
            float sv[kGateUnroll];
#pragma unroll
            for (int u = 0; u < kGateUnroll; ++u) {
                const int pix = base + u;
                const bool ok = pix < HW;
                q[u] = ok ? gate_ldg_v4(xb + (size_t)pix * a.x_pixstride + v * 8) : make_uint4(0, 0, 0, 0);
                sv[u] = ok ? __ldg(sb + pix) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < kGateUnroll; ++u) {
                const int pix = base + u;
                if (pix >= HW) continue;
                uint32_t w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f = gate_bf2(w[e]);
                    // y = x * s * g + x  (conv.py:1240-1243: x10 + x11 + x with x8 == x9)
                    const __nv_bfloat162 o = __floats2bfloat162_rn(fmaf(f.x * sv[u], gate[j][2 * e], f.x),
                                                                   fmaf(f.y * sv[u], gate[j][2 * e + 1], f.y));
                    w[e] = *reinterpret_cast<const uint32_t*>(&o);
                }
                *reinterpret_cast<uint4*>(yb + (size_t)pix * a.y_pixstride + v * 8) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }
}

static int msc_nchunk(int H, int W) {
    const int hw = H * W;
    int n = (hw + 511) / 512;
    return n < 1 ? 1 : (n > 32 ? 32 : n);
}

size_t msc_ws_bytes(int B, int H, int W, int C) {
    return ((size_t)B * 3 * H * W + (size_t)B * msc_nchunk(H, W) * C) * sizeof(float);
}

template <int LPP>
static void msc_launch_t(const MscParams& p, const GateParams& gp, cudaStream_t stream) {
    constexpr int GPB = kGateThreads / LPP;
    const specyolo_msc_gate_t& a = p.a;
    {
        const long groups = (gp.npix + kGateUnroll - 1) / kGateUnroll;
        long blocks = (groups + GPB - 1) / GPB;
        const long cap = (long)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        chan_meanmax_kernel<LPP><<<(unsigned)blocks, kGateThreads, 0, stream>>>(gp);
        count_launch();
    }
    dim3 gs((unsigned)ceil_div(a.W, kMscTW), (unsigned)ceil_div(a.H, kMscTH), (unsigned)a.B);
    msc_smap_kernel<<<gs, 128, 0, stream>>>(p);
    count_launch();
    msc_pool_kernel<LPP><<<dim3((unsigned)p.nchunk, (unsigned)a.B), kGateThreads, (size_t)GPB * a.C * sizeof(float), stream>>>(p);
    count_launch();
    const int HW = a.H * a.W;
    int per_img = ceil_div(HW, GPB * kGateUnroll);
    const int cap = max(1, sm_count() * 8 / a.B);
    if (per_img > cap) per_img = cap;
    msc_apply_kernel<LPP><<<dim3((unsigned)per_img, (unsigned)a.B), kGateThreads, (size_t)2 * a.C * sizeof(float), stream>>>(p);
    count_launch();
}

int msc_gate_launch(const specyolo_msc_gate_t* a, cudaStream_t stream) {
    SY_CHECK(a->C % 8 == 0 && a->C >= 8 && a->C <= 512, SPECYOLO_ERR_UNSUPPORTED, "msc gate: C must be a multiple of 8, <= 512 (got %d)", a->C);
    SY_CHECK(a->B <= 65535, SPECYOLO_ERR_UNSUPPORTED, "msc gate: batch %d exceeds the grid limit", a->B);
    SY_CHECK(a->k_big == kMscK, SPECYOLO_ERR_UNSUPPORTED, "msc gate: only the reference's 31 x 31 kernel is built (got %d)", a->k_big);
    SY_CHECK(a->x_pixstride % 8 == 0 && a->y_pixstride % 8 == 0 && a->x_pixstride >= a->C && a->y_pixstride >= a->C,
             SPECYOLO_ERR_INVALID, "msc gate: pixel strides must be multiples of 8 elements and >= C");
    SY_CHECK(((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->y)) & 15) == 0, SPECYOLO_ERR_INVALID,
             "msc gate: x / y must be 16-byte aligned");
    MscParams p{};
    p.a = *a;
    const size_t HW = (size_t)a->H * a->W;
    p.mm = a->ws;
    p.smap = a->ws + (size_t)a->B * 2 * HW;
    p.partial = a->ws + (size_t)a->B * 3 * HW;
    p.nchunk = msc_nchunk(a->H, a->W);
    p.vpp = a->C / 8;
    p.inv_hw = 1.0f / (float)HW;
    GateParams gp{};
    gp.a.x = a->x; gp.a.x_pixstride = a->x_pixstride; gp.a.y = a->y; gp.a.y_pixstride = a->y_pixstride;
    gp.a.B = a->B; gp.a.H = a->H; gp.a.W = a->W; gp.a.C = a->C;
    gp.a.mm = p.mm;
    gp.npix = (long)a->B * a->H * a->W;
    gp.vpp = p.vpp;
    gp.inv_c = 1.0f / (float)a->C;
    if (p.vpp >= 32) msc_launch_t<32>(p, gp, stream);
    else if (p.vpp >= 16) msc_launch_t<16>(p, gp, stream);
    else msc_launch_t<8>(p, gp, stream);
    SY_LAUNCH_CHECK();
    return SPECYOLO_OK;
}

}  // namespace specyolo
