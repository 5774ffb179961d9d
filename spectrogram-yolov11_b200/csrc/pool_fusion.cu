// HBM-bound glue kernels of the neck: SPPF pooling and Fusion('ESChannel').
#include "common.h"

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
    sppf_pool_kernel<<<grid, 256, smem, stream>>>(reinterpret_cast<__nv_bfloat16*>(buf), H, W, c, pixstride);
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
// nn.Upsample(x2, nearest) feeding a Fusion input is folded into the read (upshift).
// Pass 1 reads every input once: per-(b,channel) partial sums of squares (deterministic partials,
// no atomics) and per-pixel channel mean / max maps.  Pass 2 re-reads the inputs (L2 resident for
// the sizes here), rebuilds the k*c gates per CTA and writes the output.
// ------------------------------------------------------------------------------------------------
static constexpr int kFusPix = 64;  // pixels per pass-1 CTA

struct FusionParams {
    specyolo_fusion_t a;
    int nchunks;
    float* part;  // [B][nchunks][k*c]
    float* mm;    // [B][k][2][H*W]
    float* gate;  // [B][k*c]
};

__device__ __forceinline__ const __nv_bfloat16* fus_pix_ptr(const specyolo_fusion_t& a, int i, int b, int h, int w) {
    const int sh = a.upshift[i];
    const int Hs = a.H >> sh, Ws = a.W >> sh;
    return reinterpret_cast<const __nv_bfloat16*>(a.x[i]) +
           ((size_t)((size_t)b * Hs + (h >> sh)) * Ws + (w >> sh)) * a.pixstride[i];
}

__global__ void __launch_bounds__(256)
fusion_stats_kernel(const __grid_constant__ FusionParams p) {
    const specyolo_fusion_t& a = p.a;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int HW = a.H * a.W;
    const int vec_per_pix = a.c / 8;           // 16-byte vectors per pixel per input
    const int pix_par = 256 / vec_per_pix;     // pixels processed in parallel
    const int v = threadIdx.x % vec_per_pix;
    const int psub = threadIdx.x / vec_per_pix;
    __shared__ float red[256 * 8];
    const int pix_begin = chunk * kFusPix;
    const int pix_end = min(HW, pix_begin + kFusPix);

    for (int i = 0; i < a.k; ++i) {
        float ssq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ssq[j] = 0.f;
        for (int pix = pix_begin + psub; pix < pix_begin + kFusPix; pix += pix_par) {
            const bool ok = pix < pix_end;
            float s = 0.f, m = -INFINITY;
            if (ok) {
                const int h = pix / a.W, w = pix % a.W;
                const uint4 u = __ldg(reinterpret_cast<const uint4*>(fus_pix_ptr(a, i, b, h, w) + v * 8));
                const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = unpack_bf16x2(uu[j]);
                    ssq[2 * j] = fmaf(f.x, f.x, ssq[2 * j]);
                    ssq[2 * j + 1] = fmaf(f.y, f.y, ssq[2 * j + 1]);
                    s += f.x + f.y;
                    m = fmaxf(m, fmaxf(f.x, f.y));
                }
            }
            // reduce mean / max over the vec_per_pix lanes of this pixel (lanes are contiguous)
            for (int d = vec_per_pix >> 1; d > 0; d >>= 1) {
                s += __shfl_xor_sync(0xffffffffu, s, d);
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
            }
            if (ok && v == 0) {
                float* mm = p.mm + ((size_t)(b * a.k + i) * 2) * HW;
                mm[pix] = s / (float)a.c;
                mm[HW + pix] = m;
            }
        }
        // deterministic reduction of ssq over the pix_par sub-groups
#pragma unroll
        for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = ssq[j];
        __syncthreads();
        if (threadIdx.x < a.c) {
            const int ch = threadIdx.x;
            const int vv = ch / 8, jj = ch % 8;
            float t = 0.f;
            for (int ps = 0; ps < pix_par; ++ps) t += red[(ps * vec_per_pix + vv) * 8 + jj];
            p.part[((size_t)b * p.nchunks + chunk) * (a.k * a.c) + i * a.c + ch] = t;
        }
        __syncthreads();
    }
}

// One CTA per image: reduce the per-chunk partial sums (fixed order -> deterministic) and evaluate the k*c gates.
__global__ void __launch_bounds__(256)
fusion_gate_kernel(const __grid_constant__ FusionParams p) {
    const specyolo_fusion_t& a = p.a;
    const int b = blockIdx.x;
    const int KC = a.k * a.c;
    __shared__ float s_red[256];
    __shared__ float s_e[1024];
    float e2_local = 0.f;
    for (int ch = threadIdx.x; ch < KC; ch += 256) {
        float ssq = 0.f;
        for (int q = 0; q < p.nchunks; ++q) ssq += p.part[((size_t)b * p.nchunks + q) * KC + ch];
        const float e = sqrtf(ssq + a.gct_eps) * a.alpha[ch];
        s_e[ch] = e;
        e2_local += e * e;
    }
    s_red[threadIdx.x] = e2_local;
    __syncthreads();
    for (int d = 128; d > 0; d >>= 1) {
        if (threadIdx.x < d) s_red[threadIdx.x] += s_red[threadIdx.x + d];
        __syncthreads();
    }
    const float inv = 1.0f / sqrtf(s_red[0] / (float)KC + a.gct_eps);
    for (int ch = threadIdx.x; ch < KC; ch += 256)
        p.gate[(size_t)b * KC + ch] = 1.0f + tanhf(s_e[ch] * (a.gamma[ch] * inv) + a.beta[ch]);
}

__global__ void __launch_bounds__(256)
fusion_apply_kernel(const __grid_constant__ FusionParams p) {
    const specyolo_fusion_t& a = p.a;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int HW = a.H * a.W;
    const int KC = a.k * a.c;
    extern __shared__ float fs[];  // gate[KC]
    float* gate = fs;
    __shared__ float s_sab[3 * kFusPix];
    for (int ch = threadIdx.x; ch < KC; ch += 256) gate[ch] = p.gate[(size_t)b * KC + ch];

    // ---- spatial attention logits for this chunk's pixels ----
    const int pix_begin = chunk * kFusPix;
    for (int t = threadIdx.x; t < a.k * kFusPix; t += 256) {
        const int i = t / kFusPix, pl = t % kFusPix;
        const int pix = pix_begin + pl;
        float acc = 0.f;
        if (pix < HW) {
            const int h = pix / a.W, w = pix % a.W;
            const float* mm = p.mm + ((size_t)(b * a.k + i) * 2) * HW;
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
                        acc = fmaf(__ldg(a.sab_w + ci * 9 + ky * 3 + kx), mm[(size_t)ci * HW + hh * a.W + ww], acc);
                    }
                }
        }
        s_sab[t] = 1.0f / (1.0f + expf(-acc));
    }
    __syncthreads();

    // ---- out = sum_i x_i * (gate_i + sab_i) ----
    const int vec_per_pix = a.c / 8;
    for (int t = threadIdx.x; t < kFusPix * vec_per_pix; t += 256) {
        const int pl = t / vec_per_pix, v = t % vec_per_pix;
        const int pix = pix_begin + pl;
        if (pix >= HW) continue;
        const int h = pix / a.W, w = pix % a.W;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int i = 0; i < a.k; ++i) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(fus_pix_ptr(a, i, b, h, w) + v * 8));
            const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
            const float sab = s_sab[i * kFusPix + pl];
            const float* g = gate + i * a.c + v * 8;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16x2(uu[j]);
                acc[2 * j] = fmaf(f.x, g[2 * j] + sab, acc[2 * j]);
                acc[2 * j + 1] = fmaf(f.y, g[2 * j + 1] + sab, acc[2 * j + 1]);
            }
        }
        uint4 o;
        o.x = pack_bf16x2(acc[0], acc[1]);
        o.y = pack_bf16x2(acc[2], acc[3]);
        o.z = pack_bf16x2(acc[4], acc[5]);
        o.w = pack_bf16x2(acc[6], acc[7]);
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(a.y) +
                                  ((size_t)b * HW + pix) * a.y_pixstride + v * 8) = o;
    }
}

size_t fusion_ws_bytes(int k, int B, int H, int W, int c) {
    const int nchunks = ceil_div(H * W, kFusPix);
    return ((size_t)B * nchunks * k * c + (size_t)B * k * 2 * H * W + (size_t)B * k * c) * sizeof(float) + 256;
}

int fusion_launch(const specyolo_fusion_t* a, cudaStream_t stream) {
    SY_CHECK(a->k == 2 || a->k == 3, SPECYOLO_ERR_INVALID, "fusion: k must be 2 or 3");
    // c/8 lanes per pixel must be a power of two in [4,32] so that 256/(c/8) pixel groups divide kFusPix
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
    FusionParams p{};
    p.a = *a;
    p.nchunks = ceil_div(a->H * a->W, kFusPix);
    float* ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(a->ws) + 255) & ~(uintptr_t)255);
    p.part = ws;
    p.mm = ws + (size_t)a->B * p.nchunks * a->k * a->c;
    p.gate = p.mm + (size_t)a->B * a->k * 2 * a->H * a->W;
    SY_CHECK(a->k * a->c <= 1024, SPECYOLO_ERR_UNSUPPORTED, "fusion: k*c must be <= 1024");
    dim3 grid((unsigned)p.nchunks, (unsigned)a->B);
    fusion_stats_kernel<<<grid, 256, 0, stream>>>(p);
    SY_LAUNCH_CHECK();
    fusion_gate_kernel<<<a->B, 256, 0, stream>>>(p);
    SY_LAUNCH_CHECK();
    fusion_apply_kernel<<<grid, 256, (size_t)a->k * a->c * sizeof(float), stream>>>(p);
    SY_LAUNCH_CHECK();
    count_launch(3);
    return SPECYOLO_OK;
}

}  // namespace specyolo
