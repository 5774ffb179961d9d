// HBM-bound glue kernels of the neck: SPPF pooling and Fusion('ESChannel').
#include <cstdlib>
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

// Separable 5 x 5 max: a row pass (5-wide sliding window along w) into a scratch slab, then a column pass (along h) that
// also writes the level to global memory.  A thread owns a run of `seg` consecutive outputs of one (row | column, 8-channel
// vector) strip and keeps the window in registers: 1 shared load + 1 shared store per output and pass instead of the
// 25 loads of the direct form (the direct kernel was bound by the shared-memory pipe: 40 us for a 26 MB tensor).
// V = 8-channel vectors per pixel handled by a CTA (8 = 64 channels when the slab fits, else 2).
template <int V>
__global__ void __launch_bounds__(320)
sppf_pool_kernel(__nv_bfloat16* __restrict__ buf, int H, int W, int c, int pixstride, int seg_w, int seg_h) {
    extern __shared__ __align__(16) uint8_t sp_smem[];
    ptx::grid_dep_launch();
    ptx::grid_dep_wait();
    const int HW = H * W;
    uint4* s0 = reinterpret_cast<uint4*>(sp_smem);  // [HW][V]: current level
    uint4* s1 = s0 + HW * V;                        // [HW][V]: row-pass result
    const int b = blockIdx.y;
    const int c0 = blockIdx.x * (V * 8);
    __nv_bfloat16* img = buf + (size_t)b * HW * pixstride;
    for (int i = threadIdx.x; i < HW * V; i += blockDim.x) {
        const int pix = i / V, v = i - pix * V;
        s0[i] = *reinterpret_cast<const uint4*>(img + (size_t)pix * pixstride + c0 + v * 8);
    }
    __syncthreads();
    const uint4 NEG = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);   // bf16 -inf pairs: MaxPool2d's padding
    const int nseg_w = (W + seg_w - 1) / seg_w, nseg_h = (H + seg_h - 1) / seg_h;
    for (int level = 1; level <= 3; ++level) {
        // row pass: s0 -> s1
        for (int t = threadIdx.x; t < H * nseg_w * V; t += blockDim.x) {
            const int v = t % V, r = t / V;
            const int h = r % H, sg = r / H;
            const int w_lo = sg * seg_w, w_hi = min(W, w_lo + seg_w);
            const uint4* row = s0 + (size_t)h * W * V + v;
            uint4* orow = s1 + (size_t)h * W * V + v;
            uint4 a0 = w_lo - 2 >= 0 ? row[(w_lo - 2) * V] : NEG, a1 = w_lo - 1 >= 0 ? row[(w_lo - 1) * V] : NEG;
            uint4 a2 = row[w_lo * V], a3 = w_lo + 1 < W ? row[(w_lo + 1) * V] : NEG;
            for (int w = w_lo; w < w_hi; ++w) {
                const uint4 a4 = w + 2 < W ? row[(w + 2) * V] : NEG;
                orow[w * V] = max_bf16x8(max_bf16x8(max_bf16x8(a0, a1), max_bf16x8(a2, a3)), a4);
                a0 = a1; a1 = a2; a2 = a3; a3 = a4;
            }
        }
        __syncthreads();
        // column pass: s1 -> s0 (the next level's input) and global memory
        for (int t = threadIdx.x; t < W * nseg_h * V; t += blockDim.x) {
            const int v = t % V, r = t / V;
            const int w = r % W, sg = r / W;
            const int h_lo = sg * seg_h, h_hi = min(H, h_lo + seg_h);
            const uint4* col = s1 + (size_t)w * V + v;
            uint4* ocol = s0 + (size_t)w * V + v;
            const int hs = W * V;
            uint4 a0 = h_lo - 2 >= 0 ? col[(h_lo - 2) * hs] : NEG, a1 = h_lo - 1 >= 0 ? col[(h_lo - 1) * hs] : NEG;
            uint4 a2 = col[h_lo * hs], a3 = h_lo + 1 < H ? col[(h_lo + 1) * hs] : NEG;
            __nv_bfloat16* gcol = img + (size_t)w * pixstride + level * c + c0 + v * 8;
            for (int h = h_lo; h < h_hi; ++h) {
                const uint4 a4 = h + 2 < H ? col[(h + 2) * hs] : NEG;
                const uint4 m = max_bf16x8(max_bf16x8(max_bf16x8(a0, a1), max_bf16x8(a2, a3)), a4);
                ocol[h * hs] = m;
                *reinterpret_cast<uint4*>(gcol + (size_t)h * W * pixstride) = m;
                a0 = a1; a1 = a2; a2 = a3; a3 = a4;
            }
        }
        __syncthreads();
    }
}

int sppf_pool_launch(void* buf, int B, int H, int W, int c, int pixstride, cudaStream_t stream) {
    SY_CHECK(c % 16 == 0 && pixstride % 8 == 0 && pixstride >= 4 * c, SPECYOLO_ERR_INVALID,
             "sppf: c %% 16 == 0 and pixstride >= 4c required");
    SY_CHECK((reinterpret_cast<uintptr_t>(buf) & 15) == 0, SPECYOLO_ERR_INVALID, "sppf: unaligned buffer");
    // 64 channels per CTA when two such CTAs fit on an SM, else 16 channels
    const bool wide = (c % 64 == 0) && ((size_t)H * W * 8 * 16 * 2 <= 104 * 1024);
    const int V = wide ? 8 : 2;
    const size_t smem = (size_t)H * W * V * 16 * 2;
    SY_CHECK(smem <= 200 * 1024, SPECYOLO_ERR_UNSUPPORTED, "sppf: feature map too large (%dx%d)", H, W);
    // per-device attribute: set on every launch (host-side table lookup)
    SY_CUDA(cudaFuncSetAttribute(sppf_pool_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SY_CUDA(cudaFuncSetAttribute(sppf_pool_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    // runs of ~10 outputs per thread: enough strips x segments to fill the 320 threads
    const int seg_w = W <= 12 ? W : (W + 1) / 2 > 16 ? 16 : (W + 1) / 2;
    const int seg_h = H <= 12 ? H : (H + 1) / 2 > 16 ? 16 : (H + 1) / 2;
    dim3 grid((unsigned)(c / (V * 8)), (unsigned)B);
    if (wide)
        SY_CUDA(launch_pdl(sppf_pool_kernel<8>, grid, dim3(320), smem, stream, reinterpret_cast<__nv_bfloat16*>(buf), H, W, c, pixstride, seg_w, seg_h));
    else
        SY_CUDA(launch_pdl(sppf_pool_kernel<2>, grid, dim3(320), smem, stream, reinterpret_cast<__nv_bfloat16*>(buf), H, W, c, pixstride, seg_w, seg_h));
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
// pointer increments only, a lane owns four 8-channel vectors of a pixel — and writes per-channel partial sums of squares plus the
// per-pixel channel mean / max maps; the last CTA of an image (atomic ticket, partials summed in a fixed order, so the
// result is deterministic) evaluates the k*c gates.  fusion_apply_kernel: a CTA owns kFusRows rows of one image,
// builds the k spatial-attention maps of those rows in shared memory and writes out = sum_i x_i * (gate_i + sab_i)
// with the gates of its channel vector in registers.  v1 of these kernels was issue-bound (runtime div/mod per
// vector, scalar LDS per FMA, a serial 400-step partial reduction): 291 us for the 80x80 level; v2 (one vector per
// lane, scalar fp32 math) still spent ~120 warp instructions per 512-byte warp load: 226 us, 2.5 TB/s.
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
    unsigned int* ticket; // [B]   stats CTAs of the image that have finished
};

__device__ __forceinline__ float redux_max_f32(float v, unsigned mask) {
    float r;
    asm volatile("redux.sync.max.f32 %0, %1, %2;" : "=f"(r) : "f"(v), "r"(mask));
    return r;
}
// (packed fp32 pairs — ffma2 / fadd2 in common.h: both kernels were issue-bound, not bandwidth-bound: ncu 60 % issue
//  slots busy at 2.5 TB/s, ~120 warp instructions per 512-byte warp load)
// volatile: keeps the loads of one iteration together ahead of the math (the scheduler otherwise interleaves them
// with the first vectors' arithmetic to save registers, leaving two loads per lane in flight)
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float2 bf16x2_f2(uint32_t w) {
    return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}

// LPP = lanes per pixel = c / 32: a lane owns FOUR 16-byte vectors of one pixel (vector j*LPP + l, j = 0..3), so the
// channel mean / max of a pixel needs log2(LPP) <= 3 shuffle steps per four loads and the max runs packed in bf16.
template <int LPP>
__device__ __forceinline__ void fusion_stats_item(const FusionParams& p, int b, int item, float* red, bool* is_last_s) {
    const specyolo_fusion_t& a = p.a;
    constexpr int C = 32 * LPP;
    constexpr int PPI = 256 / LPP;             // pixels per CTA step
    // item -> (input, part)
    const int q = item % p.parts;
    const int i = item / p.parts;
    const int sh = a.upshift[i];
    const int HWs = (a.H >> sh) * (a.W >> sh);
    const int HW = a.H * a.W;
    const int KC = a.k * C;
    const int l = threadIdx.x % LPP;
    const int pg = threadIdx.x / LPP;
    const int lane = threadIdx.x & 31;
    const unsigned gmask = LPP == 32 ? 0xffffffffu : (((1u << LPP) - 1u) << (lane & ~(LPP - 1)));
    bool& is_last = *is_last_s;

    const int begin = q * p.len[i];
    const int end = min(HWs, begin + p.len[i]);
    const size_t pstride = (size_t)a.pixstride[i];
    const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(a.x[i]) + ((size_t)b * HWs + begin + pg) * pstride + l * 8;
    float* mm = p.mm + ((size_t)(b * a.k + i) * 2) * HW;
    const float inv_c = 1.0f / (float)C;

    float2 ssq[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int w = 0; w < 4; ++w) ssq[j][w] = make_float2(0.f, 0.f);
    // two pixel steps (8 x 16-byte loads per lane) in flight per iteration: with 4 the kernel was latency-bound
    // (ncu: 0.4 eligible warps per scheduler, 2.9 TB/s).  Uniform trip count: the lane reductions are convergent.
    for (int pix = begin + pg; pix - pg < end; pix += 2 * PPI, src += (size_t)(2 * PPI) * pstride) {
        uint4 u[2][4];
        bool ok[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            ok[t] = pix + t * PPI < end;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                u[t][j] = ok[t] ? ldg_nc_v4(src + (size_t)(t * PPI) * pstride + j * LPP * 8) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            float2 s2 = make_float2(0.f, 0.f);
            uint32_t mx = 0xff80ff80u;               // (-inf, -inf)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t uu[4] = {u[t][j].x, u[t][j].y, u[t][j].z, u[t][j].w};
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const float2 f = bf16x2_f2(uu[w]);
                    ssq[j][w] = ffma2(f, f, ssq[j][w]);
                    s2 = fadd2(s2, f);
                    mx = bf16x2_max(mx, uu[w]);
                }
            }
            float s = s2.x + s2.y;
            const float2 mf = bf16x2_f2(mx);
            float m = fmaxf(mf.x, mf.y);
#pragma unroll
            for (int d = LPP >> 1; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
            if (LPP > 1) m = redux_max_f32(m, gmask);
            if (ok[t] && l == 0) {
                mm[pix + t * PPI] = s * inv_c;
                mm[HW + pix + t * PPI] = m;
            }
        }
    }
    // deterministic reduction of ssq over the PPI pixel slots
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            red[threadIdx.x * 32 + j * 8 + 2 * w] = ssq[j][w].x;
            red[threadIdx.x * 32 + j * 8 + 2 * w + 1] = ssq[j][w].y;
        }
    __syncthreads();
    if ((int)threadIdx.x < C) {
        const int ch = threadIdx.x;
        const int vec = ch >> 3, e = ch & 7;
        const int j = vec / LPP, ll = vec % LPP;
        float t = 0.f;
        for (int g = 0; g < PPI; ++g) t += red[(g * LPP + ll) * 32 + j * 8 + e];
        p.part[((size_t)b * p.parts + q) * KC + i * C + ch] = t * (float)(1 << (2 * sh));   // upsampled: each source pixel counts 4x
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

// K inputs; a thread owns one 8-channel vector (its K*8 gates packed in registers) and walks the CTA's pixels.
template <int K>
__device__ __forceinline__ void fusion_apply_item(const FusionParams& p, int b, int hblk, float* s_sab) {
    const specyolo_fusion_t& a = p.a;
    const int h0 = hblk * kFusRows;
    const int rows = min(kFusRows, a.H - h0);
    const int HW = a.H * a.W;
    const int KC = K * a.c;
    const int vpp = a.c >> 3;
    const int pix_par = 256 / vpp;
    const int v = threadIdx.x % vpp;
    const int psub = threadIdx.x / vpp;
    // s_sab: [K][rows][W]
    float2 g[K][4];
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const float4* gp = reinterpret_cast<const float4*>(p.gate + (size_t)b * KC + i * a.c + v * 8);
        const float4 g0 = __ldcg(gp), g1 = __ldcg(gp + 1);
        g[i][0] = make_float2(g0.x, g0.y); g[i][1] = make_float2(g0.z, g0.w);
        g[i][2] = make_float2(g1.x, g1.y); g[i][3] = make_float2(g1.z, g1.w);
    }

    // ---- spatial attention of the CTA's rows: sigmoid(conv3x3([mean, max])) at each input's source resolution ----
    const int npix = rows * a.W;
    for (int t = threadIdx.x; t < K * npix; t += 256) {
        const int i = t / npix;
        const int r = t - i * npix;
        const int hr = r / a.W, w = r - hr * a.W;
        const int sh = a.upshift[i];
        const int Ws = a.W >> sh;
        const int h = h0 + hr;
        const float* mm = p.mm + ((size_t)(b * K + i) * 2) * HW;
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
                    acc = fmaf(__ldg(a.sab_w + ci * 9 + ky * 3 + kx), __ldcg(mm + (size_t)ci * HW + (hh >> sh) * Ws + (ww >> sh)), acc);
                }
            }
        s_sab[t] = 1.0f / (1.0f + __expf(-acc));
    }
    __syncthreads();

    // ---- out = sum_i x_i * (gate_i + sab_i) over the CTA's rows x W pixels (flattened, (hr, w) carried along) ----
    const __nv_bfloat16* xb[K];
    int wshift[K], rowstride[K], pstr[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const int sh = a.upshift[i];
        const int Hs = a.H >> sh, Ws = a.W >> sh;
        xb[i] = reinterpret_cast<const __nv_bfloat16*>(a.x[i]) + ((size_t)b * Hs * Ws) * a.pixstride[i] + v * 8;
        wshift[i] = sh;
        rowstride[i] = Ws;
        pstr[i] = a.pixstride[i];
    }
    __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(a.y) + ((size_t)b * HW + (size_t)h0 * a.W) * a.y_pixstride + v * 8;
    // two pixels (2K loads) in flight per iteration
    int hr[2], w[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        hr[t] = 0;
        w[t] = psub + t * pix_par;
        while (w[t] >= a.W) { w[t] -= a.W; ++hr[t]; }
    }
    for (int q = psub; q < npix; q += 2 * pix_par) {
        uint4 u[2][K];
        bool ok[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            ok[t] = q + t * pix_par < npix;
            const int h = h0 + hr[t];
#pragma unroll
            for (int i = 0; i < K; ++i) {
                const int spix = (h >> wshift[i]) * rowstride[i] + (w[t] >> wshift[i]);
                u[t][i] = ok[t] ? ldg_nc_v4(xb[i] + (size_t)spix * pstr[i]) : make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            if (!ok[t]) continue;
            const int qq = q + t * pix_par;
            float2 acc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < K; ++i) {
                const uint32_t uu[4] = {u[t][i].x, u[t][i].y, u[t][i].z, u[t][i].w};
                const float sab = s_sab[i * npix + qq];
                const float2 sab2 = make_float2(sab, sab);
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[j] = ffma2(bf16x2_f2(uu[j]), fadd2(g[i][j], sab2), acc[j]);
            }
            uint4 o;
            o.x = pack_bf16x2(acc[0].x, acc[0].y);
            o.y = pack_bf16x2(acc[1].x, acc[1].y);
            o.z = pack_bf16x2(acc[2].x, acc[2].y);
            o.w = pack_bf16x2(acc[3].x, acc[3].y);
            *reinterpret_cast<uint4*>(yb + (size_t)qq * a.y_pixstride) = o;
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            w[t] += 2 * pix_par;
            while (w[t] >= a.W) { w[t] -= a.W; ++hr[t]; }
        }
    }
}

// (A single-launch variant — statistics items of image b + D followed by the apply items of image b, handed out in
//  order by an atomic counter, apply items spinning on a per-image publish flag — was measured: the hoped-for L2 hits
//  on the second read of the inputs did not materialise; D = 2 .. 24 ran 206 .. 150 us against 160 us for the two
//  launches below at the 80x80 level, so the simple form stays.)
template <int LPP>
__global__ void __launch_bounds__(256)
fusion_stats_kernel(const __grid_constant__ FusionParams p) {
    __shared__ float red[256 * 32];
    __shared__ bool is_last;
    ptx::grid_dep_launch();
    ptx::grid_dep_wait();
    const int S = p.a.k * p.parts;
    fusion_stats_item<LPP>(p, blockIdx.x / S, blockIdx.x % S, red, &is_last);
}

template <int K>
__global__ void __launch_bounds__(256, 3)
fusion_apply_kernel(const __grid_constant__ FusionParams p) {
    extern __shared__ float s_sab[];
    ptx::grid_dep_launch();
    ptx::grid_dep_wait();
    fusion_apply_item<K>(p, blockIdx.y, blockIdx.x, s_sab);
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
    const int ppi = 256 / (a->c / 32);          // pixels per stats CTA step
    for (int i = 0; i < a->k; ++i) {
        const int HWs = (a->H >> a->upshift[i]) * (a->W >> a->upshift[i]);
        p.len[i] = ceil_div(ceil_div(HWs, p.parts), 2 * ppi) * 2 * ppi;
    }
    float* ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(a->ws) + 255) & ~(uintptr_t)255);
    p.part = ws;
    p.mm = p.part + (size_t)a->B * kFusMaxParts * a->k * a->c;
    p.gate = p.mm + (((size_t)a->B * a->k * 2 * a->H * a->W + 3) & ~(size_t)3);     // float4 reads of the gates
    p.ticket = reinterpret_cast<unsigned int*>(p.gate + (size_t)a->B * a->k * a->c);
    SY_CUDA(cudaMemsetAsync(p.ticket, 0, (size_t)a->B * sizeof(unsigned int), stream));
    const dim3 sgrid((unsigned)(a->B * a->k * p.parts));
    switch (a->c) {
        case 32: SY_CUDA(launch_pdl(fusion_stats_kernel<1>, sgrid, dim3(256), 0, stream, p)); break;
        case 64: SY_CUDA(launch_pdl(fusion_stats_kernel<2>, sgrid, dim3(256), 0, stream, p)); break;
        case 128: SY_CUDA(launch_pdl(fusion_stats_kernel<4>, sgrid, dim3(256), 0, stream, p)); break;
        default: SY_CUDA(launch_pdl(fusion_stats_kernel<8>, sgrid, dim3(256), 0, stream, p)); break;
    }
    SY_LAUNCH_CHECK();
    const dim3 grid((unsigned)ceil_div(a->H, kFusRows), (unsigned)a->B);
    if (a->k == 2) SY_CUDA(launch_pdl(fusion_apply_kernel<2>, grid, dim3(256), sab_smem, stream, p));
    else SY_CUDA(launch_pdl(fusion_apply_kernel<3>, grid, dim3(256), sab_smem, stream, p));
    SY_LAUNCH_CHECK();
    count_launch(2);
    return SPECYOLO_OK;
}

}  // namespace specyolo
