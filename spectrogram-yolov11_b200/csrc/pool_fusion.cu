// HBM-bound glue kernels of the neck: SPPF pooling and Fusion('ESChannel').
#include "common.h"
#include "tma_host.h"
#include "ptx.cuh"

namespace specyolo {

// ------------------------------------------------------------------------------------------------
// SPPF (ultralytics/nn/modules/block.py:194-198): y1 = mp5(y0), y2 = mp5(y1), y3 = mp5(y2) with
// MaxPool2d(5, 1, 2) (implicit -inf padding).  y0 already sits in channels [0,c) of the 4c-channel
// concat buffer (written there by cv1's epilogue); this kernel fills [c,4c).  One CTA = one image x
// 16 channels: the H x W x 16 slab is staged in shared memory once and the three pools are chained
// there, so HBM sees one read and three writes of the slab.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 max_bf16x8(const uint4& a, const uint4& b) {
    uint4 r;
    const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
    __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) pr[i] = __hmax2(pa[i], pb[i]);
    return r;
}

__global__ void __launch_bounds__(256)
sppf_pool_kernel(__nv_bfloat16* __restrict__ buf, int H, int W, int c, int pixstride) {
    extern __shared__ __align__(16) uint8_t sp_smem[];
    ptx::grid_dep_launch();
    ptx::grid_dep_wait();
    const int HW = H * W;
    uint4* s0 = reinterpret_cast<uint4*>(sp_smem);  // [HW][2] (16 channels = 2 x 16 B)
    uint4* s1 = s0 + HW * 2;
    const int b = blockIdx.y;
    const int c0 = blockIdx.x * 16;
    __nv_bfloat16* img = buf + (size_t)b * HW * pixstride;
    for (int i = threadIdx.x; i < HW * 2; i += blockDim.x) {
        const int pix = i >> 1, half = i & 1;
        s0[i] = *reinterpret_cast<const uint4*>(img + (size_t)pix * pixstride + c0 + half * 8);
    }
    __syncthreads();
    uint4* src = s0;
    uint4* dst = s1;
    for (int level = 1; level <= 3; ++level) {
        for (int i = threadIdx.x; i < HW * 2; i += blockDim.x) {
            const int pix = i >> 1, half = i & 1;
            const int h = pix / W, w = pix % W;
            uint4 m = src[i];
            for (int dy = -2; dy <= 2; ++dy) {
                const int hh = h + dy;
                if (hh < 0 || hh >= H) continue;
                for (int dx = -2; dx <= 2; ++dx) {
                    const int ww = w + dx;
                    if (ww < 0 || ww >= W) continue;
                    m = max_bf16x8(m, src[(hh * W + ww) * 2 + half]);
                }
            }
            dst[i] = m;
            *reinterpret_cast<uint4*>(img + (size_t)pix * pixstride + level * c + c0 + half * 8) = m;
        }
        __syncthreads();
        uint4* t = src; src = dst; dst = t;
    }
}

int sppf_pool_launch(void* buf, int B, int H, int W, int c, int pixstride, cudaStream_t stream) {
    SY_CHECK(c % 16 == 0 && pixstride % 8 == 0 && pixstride >= 4 * c, SPECYOLO_ERR_INVALID,
             "sppf: c %% 16 == 0 and pixstride >= 4c required");
    SY_CHECK((reinterpret_cast<uintptr_t>(buf) & 15) == 0, SPECYOLO_ERR_INVALID, "sppf: unaligned buffer");
    const size_t smem = (size_t)H * W * 2 * 16 * 2;
    SY_CHECK(smem <= 200 * 1024, SPECYOLO_ERR_UNSUPPORTED, "sppf: feature map too large (%dx%d)", H, W);
    static size_t attr_smem = 0;
    if (smem > 48 * 1024 && smem > attr_smem) {
        SY_CUDA(cudaFuncSetAttribute(sppf_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    dim3 grid((unsigned)(c / 16), (unsigned)B);
    SY_CUDA(launch_pdl(sppf_pool_kernel, grid, dim3(256), smem, stream, reinterpret_cast<__nv_bfloat16*>(buf), H, W, c, pixstride));
    SY_LAUNCH_CHECK();
    count_launch();
    return SPECYOLO_OK;
}

// ------------------------------------------------------------------------------------------------
// Fusion('ESChannel')  (ultralytics/nn/modules/conv.py:2113-2127)
//   A = cat(x_1..x_k);  G = GCT(A) (conv.py:2296-2301);  out = sum_i ( G_i + sab(x_i) )
//   GCT:  e_c = alpha_c * sqrt(sum_hw A_c^2 + eps);  n_c = gamma_c / sqrt(mean_c(e^2) + eps);
//         gate_c = 1 + tanh(e_c * n_c + beta_c);  G = A * gate
//   sab (WeightedSpatialAttention, conv.py:1850-1852): x * sigmoid(conv3x3([mean_c x, max_c x]))
// nn.Upsample(x2, nearest) feeding a Fusion input is folded into the reads (upshift): its statistics are taken
// at the SOURCE resolution (sum of squares x 4; mean / max maps stored at source size and read through (h/2, w/2)
// by the full-resolution 3x3 spatial-attention conv).
//
// Two launches.  fusion_stats_kernel: every CTA streams a contiguous run of source pixels of ONE (image, input) —
// pointer increments only, a lane owns one 8-channel vector — and writes per-channel partial sums of squares plus the
// per-pixel channel mean / max maps; the last CTA of an image (atomic ticket, partials summed in a fixed order, so the
// result is deterministic) evaluates the k*c gates.  fusion_apply_kernel: a CTA owns kFusRows rows of one image,
// builds the k spatial-attention maps of those rows in shared memory and writes out = sum_i x_i * (gate_i + sab_i)
// with the gates of its channel vector in registers.  v1 of these kernels was issue-bound (runtime div/mod per
// vector, scalar LDS per FMA, a serial 400-step partial reduction): 291 us for the 80x80 level.
// ------------------------------------------------------------------------------------------------
static constexpr int kFusRows = 4;      // image rows per apply CTA
static constexpr int kFusMaxParts = 16;

struct FusionParams {
    specyolo_fusion_t a;
    int parts;            // stats CTAs per (image, input)
    int len[3];           // source pixels per stats CTA of input i (multiple of the pixels processed per step)
    float* part;          // [B][parts][k*c]
    float* mm;            // [B][k][2][H*W]   (input i uses the first HWs_i entries of each plane)
    float* gate;          // [B][k*c]
    unsigned int* ticket; // [B]
};

__device__ __forceinline__ float redux_max_f32(float v, unsigned mask) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, %2;" : "=f"(r) : "f"(v), "r"(mask));
    return r;
}

__global__ void __launch_bounds__(256)
fusion_stats_kernel(const __grid_constant__ FusionParams p) {
    ptx::grid_dep_launch();
    ptx::grid_dep_wait();
    const specyolo_fusion_t& a = p.a;
    // work item -> (image, input, part)
    int item = blockIdx.x;
    const int q = item % p.parts;
    item /= p.parts;
    const int i = item % a.k;
    const int b = item / a.k;
    const int sh = a.upshift[i];
    const int HWs = (a.H >> sh) * (a.W >> sh);
    const int HW = a.H * a.W;
    const int KC = a.k * a.c;
    const int vpp = a.c >> 3;                  // 16-byte vectors (= lanes) per pixel: 4, 8, 16 or 32
    const int pix_par = 256 / vpp;             // pixels per step
    const int v = threadIdx.x % vpp;
    const int psub = threadIdx.x / vpp;
    const int lane = threadIdx.x & 31;
    const unsigned gmask = vpp == 32 ? 0xffffffffu : (((1u << vpp) - 1u) << (lane & ~(vpp - 1)));
    __shared__ float red[256 * 8];
    __shared__ bool is_last;

    const int begin = q * p.len[i];
    const int end = min(HWs, begin + p.len[i]);
    const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(a.x[i]) + ((size_t)b * HWs) * a.pixstride[i] + v * 8;
    float* mm = p.mm + ((size_t)(b * a.k + i) * 2) * HW;
    const float inv_c = 1.0f / (float)a.c;

    float ssq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ssq[j] = 0.f;
    // 4 pixel steps per iteration, all four 16-byte loads issued before any use: one load per thread in flight left
    // the kernel latency-bound (uniform trip count: the lane reductions below are convergent)
    for (int pix0 = begin; pix0 < end; pix0 += 4 * pix_par) {
        uint4 u[4];
        bool ok[4];
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
            const int pix = pix0 + s4 * pix_par + psub;
            ok[s4] = pix < end;
            u[s4] = ok[s4] ? __ldg(reinterpret_cast<const uint4*>(src + (size_t)pix * a.pixstride[i])) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
            const int pix = pix0 + s4 * pix_par + psub;
            const uint32_t uu[4] = {u[s4].x, u[s4].y, u[s4].z, u[s4].w};
            float s = 0.f, m = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16x2(uu[j]);
                ssq[2 * j] = fmaf(f.x, f.x, ssq[2 * j]);
                ssq[2 * j + 1] = fmaf(f.y, f.y, ssq[2 * j + 1]);
                s += f.x + f.y;
                m = fmaxf(m, fmaxf(f.x, f.y));
            }
            // channel mean / max of the pixel: reduce over its vpp lanes
            for (int d = vpp >> 1; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
            m = redux_max_f32(m, gmask);
            if (ok[s4] && v == 0) {
                mm[pix] = s * inv_c;
                mm[HW + pix] = m;
            }
        }
    }
    // deterministic reduction of ssq over the pix_par pixel groups
#pragma unroll
    for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = ssq[j];
    __syncthreads();
    if ((int)threadIdx.x < a.c) {
        const int ch = threadIdx.x;
        const int vv = ch >> 3, jj = ch & 7;
        float t = 0.f;
        for (int ps = 0; ps < pix_par; ++ps) t += red[(ps * vpp + vv) * 8 + jj];
        p.part[((size_t)b * p.parts + q) * KC + i * a.c + ch] = t * (float)(1 << (2 * sh));   // upsampled: each source pixel counts 4x
    }
    // ---- last CTA of this image: gates ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(p.ticket + b, 1u);
        is_last = (t == (unsigned)(a.k * p.parts) - 1u);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    float* s_e = red;                  // KC <= 1024 floats
    float* s_red = red + 1024;
    float e2_local = 0.f;
    for (int ch = threadIdx.x; ch < KC; ch += 256) {
        float t = 0.f;
        for (int qq = 0; qq < p.parts; ++qq) t += __ldcg(p.part + ((size_t)b * p.parts + qq) * KC + ch);
        const float e = sqrtf(t + a.gct_eps) * a.alpha[ch];
        s_e[ch] = e;
        e2_local += e * e;
    }
    s_red[threadIdx.x] = e2_local;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) s_red[threadIdx.x] += s_red[threadIdx.x + d];
        __syncthreads();
    }
    const float inv = 1.0f / sqrtf(s_red[0] / (float)KC + a.gct_eps);
    for (int ch = threadIdx.x; ch < KC; ch += 256)
        p.gate[(size_t)b * KC + ch] = 1.0f + tanhf(s_e[ch] * (a.gamma[ch] * inv) + a.beta[ch]);
}

__global__ void __launch_bounds__(256)
fusion_apply_kernel(const __grid_constant__ FusionParams p) {
    ptx::grid_dep_launch();
    ptx::grid_dep_wait();
    const specyolo_fusion_t& a = p.a;
    const int b = blockIdx.y;
    const int h0 = blockIdx.x * kFusRows;
    const int rows = min(kFusRows, a.H - h0);
    const int HW = a.H * a.W;
    const int KC = a.k * a.c;
    const int vpp = a.c >> 3;
    const int pix_par = 256 / vpp;
    const int v = threadIdx.x % vpp;
    const int psub = threadIdx.x / vpp;
    extern __shared__ float s_sab[];          // [k][rows][W]

    // gates of this thread's channel vector: registers
    float g[3][8];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) g[i][j] = (i < a.k) ? __ldg(p.gate + (size_t)b * KC + i * a.c + v * 8 + j) : 0.f;

    // ---- spatial attention of the CTA's rows: sigmoid(conv3x3([mean, max])) at each input's source resolution ----
    const int npix = rows * a.W;
    for (int t = threadIdx.x; t < a.k * npix; t += 256) {
        const int i = t / npix;
        const int r = t - i * npix;
        const int hr = r / a.W, w = r - hr * a.W;
        const int sh = a.upshift[i];
        const int Ws = a.W >> sh;
        const int h = h0 + hr;
        const float* mm = p.mm + ((size_t)(b * a.k + i) * 2) * HW;
        // the 3x3 conv runs on the FULL-resolution (upsampled) maps: neighbours (h+-1, w+-1) -> source (>> sh)
        float acc = 0.f;
#pragma unroll
        for (int ci = 0; ci < 2; ++ci)
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int hh = h + ky - 1;
                if (hh < 0 || hh >= a.H) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int ww = w + kx - 1;
                    if (ww < 0 || ww >= a.W) continue;
                    acc = fmaf(__ldg(a.sab_w + ci * 9 + ky * 3 + kx), mm[(size_t)ci * HW + (hh >> sh) * Ws + (ww >> sh)], acc);
                }
            }
        s_sab[t] = 1.0f / (1.0f + __expf(-acc));
    }
    __syncthreads();

    // ---- out = sum_i x_i * (gate_i + sab_i) ----
    const __nv_bfloat16* xb[3];
    int wshift[3], rowstride[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int ii = i < a.k ? i : 0;
        const int sh = a.upshift[ii];
        const int Hs = a.H >> sh, Ws = a.W >> sh;
        xb[i] = reinterpret_cast<const __nv_bfloat16*>(a.x[ii]) + ((size_t)b * Hs * Ws) * a.pixstride[ii] + v * 8;
        wshift[i] = sh;
        rowstride[i] = Ws;
    }
    __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(a.y) + ((size_t)b * HW) * a.y_pixstride + v * 8;
    for (int hr = 0; hr < rows; ++hr) {
        const int h = h0 + hr;
        // two pixels per iteration, their 2k loads issued before any use
        for (int w0 = psub; w0 < a.W; w0 += 2 * pix_par) {
            uint4 u[2][3];
            bool ok[2];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int w = w0 + t * pix_par;
                ok[t] = w < a.W;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    if (i < a.k && ok[t]) {
                        const size_t spix = (size_t)(h >> wshift[i]) * rowstride[i] + (w >> wshift[i]);
                        u[t][i] = __ldg(reinterpret_cast<const uint4*>(xb[i] + spix * a.pixstride[i]));
                    } else {
                        u[t][i] = make_uint4(0, 0, 0, 0);
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                if (!ok[t]) continue;
                const int w = w0 + t * pix_par;
                float acc[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    if (i >= a.k) break;
                    const uint32_t uu[4] = {u[t][i].x, u[t][i].y, u[t][i].z, u[t][i].w};
                    const float sab = s_sab[(i * rows + hr) * a.W + w];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = unpack_bf16x2(uu[j]);
                        acc[2 * j] = fmaf(f.x, g[i][2 * j] + sab, acc[2 * j]);
                        acc[2 * j + 1] = fmaf(f.y, g[i][2 * j + 1] + sab, acc[2 * j + 1]);
                    }
                }
                uint4 o;
                o.x = pack_bf16x2(acc[0], acc[1]);
                o.y = pack_bf16x2(acc[2], acc[3]);
                o.z = pack_bf16x2(acc[4], acc[5]);
                o.w = pack_bf16x2(acc[6], acc[7]);
                *reinterpret_cast<uint4*>(yb + ((size_t)h * a.W + w) * a.y_pixstride) = o;
            }
        }
    }
}

// stats CTAs per (image, input): a function of the image size ONLY, so that the fp32 summation order of the per-channel
// sums of squares — and with it every gate, bit for bit — does not depend on the batch an image is part of
static int fusion_parts(int HW) {
    int parts = HW / 800;               // 80x80 -> 8, 40x40 -> 2, 20x20 -> 1
    if (parts < 1) parts = 1;
    if (parts > kFusMaxParts) parts = kFusMaxParts;
    return parts;
}

size_t fusion_ws_bytes(int k, int B, int H, int W, int c) {
    return ((size_t)B * kFusMaxParts * k * c + (size_t)B * k * 2 * H * W + (size_t)B * k * c) * sizeof(float) +
           (size_t)B * sizeof(unsigned int) + 512;
}

int fusion_launch(const specyolo_fusion_t* a, cudaStream_t stream) {
    SY_CHECK(a->k == 2 || a->k == 3, SPECYOLO_ERR_INVALID, "fusion: k must be 2 or 3");
    // c/8 lanes per pixel must be a power of two in [4,32]
    SY_CHECK(a->c % 8 == 0 && ((a->c / 8) & (a->c / 8 - 1)) == 0 && a->c / 8 >= 4 && a->c / 8 <= 32,
             SPECYOLO_ERR_UNSUPPORTED, "fusion: c must be 32, 64, 128 or 256 (c=%d)", a->c);
    for (int i = 0; i < a->k; ++i) {
        SY_CHECK(a->pixstride[i] % 8 == 0 && (reinterpret_cast<uintptr_t>(a->x[i]) & 15) == 0,
                 SPECYOLO_ERR_INVALID, "fusion: input %d not 16-byte aligned", i);
        SY_CHECK(a->upshift[i] == 0 || (a->upshift[i] == 1 && a->H % 2 == 0 && a->W % 2 == 0),
                 SPECYOLO_ERR_INVALID, "fusion: bad upshift");
    }
    SY_CHECK(a->y_pixstride % 8 == 0 && (reinterpret_cast<uintptr_t>(a->y) & 15) == 0, SPECYOLO_ERR_INVALID,
             "fusion: output not 16-byte aligned");
    SY_CHECK(a->ws != nullptr, SPECYOLO_ERR_INVALID, "fusion: workspace missing");
    SY_CHECK(a->k * a->c <= 1024, SPECYOLO_ERR_UNSUPPORTED, "fusion: k*c must be <= 1024");
    const size_t sab_smem = (size_t)a->k * kFusRows * a->W * sizeof(float);
    SY_CHECK(sab_smem <= 48 * 1024, SPECYOLO_ERR_UNSUPPORTED, "fusion: feature map too wide (%d)", a->W);
    FusionParams p{};
    p.a = *a;
    p.parts = fusion_parts(a->H * a->W);
    const int pix_par = 256 / (a->c / 8);
    for (int i = 0; i < a->k; ++i) {
        const int HWs = (a->H >> a->upshift[i]) * (a->W >> a->upshift[i]);
        p.len[i] = ceil_div(ceil_div(HWs, p.parts), 4 * pix_par) * 4 * pix_par;
    }
    float* ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(a->ws) + 255) & ~(uintptr_t)255);
    p.part = ws;
    p.mm = p.part + (size_t)a->B * kFusMaxParts * a->k * a->c;
    p.gate = p.mm + (size_t)a->B * a->k * 2 * a->H * a->W;
    p.ticket = reinterpret_cast<unsigned int*>(p.gate + (size_t)a->B * a->k * a->c);
    SY_CUDA(cudaMemsetAsync(p.ticket, 0, (size_t)a->B * sizeof(unsigned int), stream));
    SY_CUDA(launch_pdl(fusion_stats_kernel, dim3((unsigned)(a->B * a->k * p.parts)), dim3(256), 0, stream, p));
    SY_LAUNCH_CHECK();
    dim3 grid((unsigned)ceil_div(a->H, kFusRows), (unsigned)a->B);
    SY_CUDA(launch_pdl(fusion_apply_kernel, grid, dim3(256), sab_smem, stream, p));
    SY_LAUNCH_CHECK();
    count_launch(2);
    return SPECYOLO_OK;
}

}  // namespace specyolo
