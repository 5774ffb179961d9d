// CUDA-core kernels around the implicit-GEMM conv: BN fold + weight repack, the 3-channel stem conv,
// depthwise 3x3, and NCHW<->NHWC layout conversion.  All are HBM-bound streaming kernels: 16-byte
// vector accesses along the contiguous channel axis, grids sized in multiples of the SM count.
#include "common.h"
#include "tma_host.h"
#include "ptx.cuh"
#include <cstdlib>

namespace specyolo {

// ------------------------------------------------------------------------------------------------
// BN fold + repack:  w_packed[g][n][ky][kx][ci] (bf16), rows n >= cout_g are zero.
// fuse_conv_and_bn (ultralytics/utils/torch_utils.py:238-265):
//   w' = w * gamma / sqrt(var + eps);  b' = (b_conv) * gamma / sqrt(var+eps) + beta - gamma*mean/sqrt(var+eps)
// ------------------------------------------------------------------------------------------------
// `merge` source groups are fused into one packed group with block-diagonal weights (zeros off the diagonal):
// a g=8, 16->16-per-group conv becomes a g=2, 64->64-per-group conv whose K chunks are full 128-byte swizzle
// rows and whose UMMA N is 64 instead of 16 — 4x fewer TMA / MMA / barrier operations for the same bytes.
__global__ void fold_pack_kernel(const float* __restrict__ w, const float* __restrict__ conv_bias,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                 int cout, int cin_g, int kh, int kw, int groups, int merge, int n_pad,
                                 __nv_bfloat16* __restrict__ wp, float* __restrict__ bias_out) {
    const int cout_g = cout / groups;
    const int taps = kh * kw;
    const int pgroups = groups / merge;           // packed groups
    const int pcin = cin_g * merge;               // packed input channels per group
    const int pcout = cout_g * merge;             // packed (valid) output channels per group
    const long per_row = (long)taps * pcin;
    const long total = (long)pgroups * n_pad * per_row;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % pcin);
        const int tap = (int)((i / pcin) % taps);
        const int n = (int)((i / per_row) % n_pad);
        const int pg = (int)(i / (per_row * n_pad));
        float v = 0.f;
        if (n < pcout && (n / cout_g) == (ci / cin_g)) {
            const int co = pg * pcout + n;        // == source group * cout_g + local index
            const float s = gamma ? gamma[co] / sqrtf(var[co] + eps) : 1.f;
            // OIHW: ((co*cin_g + ci_local)*kh + ky)*kw + kx ; tap = ky*kw + kx
            v = w[((long)co * cin_g + (ci % cin_g)) * taps + tap] * s;
        }
        wp[i] = __float2bfloat16_rn(v);
    }
    const int nb = pgroups * n_pad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += gridDim.x * blockDim.x) {
        const int n = i % n_pad, pg = i / n_pad;
        float b = 0.f;
        if (n < pcout) {
            const int co = pg * pcout + n;
            const float cb = conv_bias ? conv_bias[co] : 0.f;
            if (gamma) {
                const float s = gamma[co] / sqrtf(var[co] + eps);
                b = cb * s + beta[co] - mean[co] * s;
            } else {
                b = cb;
            }
        }
        bias_out[i] = b;
    }
}

int fold_pack_launch(const float* w, const float* conv_bias, const float* gamma, const float* beta,
                     const float* mean, const float* var, float eps, int cout, int cin_g, int kh, int kw,
                     int groups, int merge, int n_pad, void* wp, float* bias_out, cudaStream_t stream) {
    SY_CHECK(cout % groups == 0, SPECYOLO_ERR_INVALID, "cout %% groups != 0");
    SY_CHECK(merge >= 1 && groups % merge == 0, SPECYOLO_ERR_INVALID, "merge must divide groups");
    SY_CHECK(n_pad >= cout / groups * merge, SPECYOLO_ERR_INVALID, "n_pad too small");
    const long total = (long)(groups / merge) * n_pad * kh * kw * cin_g * merge;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    fold_pack_kernel<<<blocks, 256, 0, stream>>>(w, conv_bias, gamma, beta, mean, var, eps, cout, cin_g, kh,
                                                 kw, groups, merge, n_pad, reinterpret_cast<__nv_bfloat16*>(wp),
                                                 bias_out);
    SY_LAUNCH_CHECK();
    count_launch();
    return SPECYOLO_OK;
}

// ------------------------------------------------------------------------------------------------
// Stem: 3x3 stride-2 pad-1 conv on the 3-channel NCHW network input, SiLU, NHWC bf16 output.
// K = 27 is far too thin for the tensor pipe; the layer is bound by its 3.7x larger output write.
// One thread produces 8 output channels of one pixel (a 16-byte store); the 27 input taps are read
// through L1 (each input pixel is shared by ~2.25 outputs x Cout/8 threads).
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float load_in(const T* p, long i);
template <>
__device__ __forceinline__ float load_in<float>(const float* p, long i) { return __ldg(p + i); }
template <>
__device__ __forceinline__ float load_in<__nv_bfloat16>(const __nv_bfloat16* p, long i) {
    return __bfloat162float(p[i]);
}
template <>
__device__ __forceinline__ float load_in<uint8_t>(const uint8_t* p, long i) {
    return (float)__ldg(p + i) * (1.0f / 255.0f);
}

// Output tile per CTA: 8 rows x 32 columns = 256 pixels, one per thread, every output channel.  The 17 x 65 x 3
// input patch is staged once in shared memory (coalesced reads of the NCHW planes, zero padding, uint8 -> float
// scaling) with even / odd columns de-interleaved so that the stride-2 taps of neighbouring threads hit
// consecutive banks.  Weights live in shared memory as [27][Cout] and are read as broadcast float4; the FMAs
// are issued as packed fp32x2 (FFMA2), 16 output channels at a time, stored as two 16-byte vectors.
static constexpr int kStemTH = 8, kStemTW = 32;
static constexpr int kStemIH = 2 * kStemTH + 1, kStemIWH = kStemTW + 1;   // rows, columns per parity plane

template <typename T>
__global__ void __launch_bounds__(256)
stem_conv_kernel(const T* __restrict__ x, int B, int H, int W, const float* __restrict__ wgt,
                 const float* __restrict__ bias, int Cout, __nv_bfloat16* __restrict__ y, int y_pixstride,
                 int Ho, int Wo) {
    extern __shared__ __align__(16) float stem_smem[];
    float* sw = stem_smem;                       // [27][Cout]
    float* sb = sw + 27 * Cout;                  // [Cout]
    float* sin = sb + Cout;                      // [3][kStemIH][2][kStemIWH]
    const int n = blockIdx.z;
    const int oh0 = blockIdx.y * kStemTH, ow0 = blockIdx.x * kStemTW;
    for (int i = threadIdx.x; i < 27 * Cout; i += 256) {
        const int co = i % Cout, k = i / Cout;   // k = ci*9 + ky*3 + kx (OIHW inner order)
        sw[i] = wgt[co * 27 + k];
    }
    for (int i = threadIdx.x; i < Cout; i += 256) sb[i] = bias[i];
    const int ih0 = 2 * oh0 - 1, iw0 = 2 * ow0 - 1;
    for (int i = threadIdx.x; i < 3 * kStemIH * (2 * kStemTW + 1); i += 256) {
        const int c = i % (2 * kStemTW + 1);
        const int r = (i / (2 * kStemTW + 1)) % kStemIH;
        const int ci = i / ((2 * kStemTW + 1) * kStemIH);
        const int ih = ih0 + r, iw = iw0 + c;
        float v = 0.f;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = load_in<T>(x, (((long)n * 3 + ci) * H + ih) * W + iw);
        sin[((ci * kStemIH + r) * 2 + (c & 1)) * kStemIWH + (c >> 1)] = v;
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int oh = oh0 + ty, ow = ow0 + tx;
    float in[27];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const float* row = sin + ((ci * kStemIH + 2 * ty + ky) * 2) * kStemIWH;
            in[ci * 9 + ky * 3 + 0] = row[tx];                 // column 2tx   (even plane)
            in[ci * 9 + ky * 3 + 1] = row[kStemIWH + tx];      // column 2tx+1 (odd plane)
            in[ci * 9 + ky * 3 + 2] = row[tx + 1];             // column 2tx+2 (even plane)
        }
    if (oh >= Ho || ow >= Wo) return;
    __nv_bfloat16* yp = y + (((long)n * Ho + oh) * Wo + ow) * y_pixstride;
    for (int c0 = 0; c0 < Cout; c0 += 16) {
        float2 acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = make_float2(sb[c0 + 2 * j], sb[c0 + 2 * j + 1]);
#pragma unroll
        for (int k = 0; k < 27; ++k) {
            const float4* w4 = reinterpret_cast<const float4*>(sw + k * Cout + c0);
            const float2 v2 = make_float2(in[k], in[k]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 w = w4[q];
                acc[2 * q] = __ffma2_rn(v2, make_float2(w.x, w.y), acc[2 * q]);
                acc[2 * q + 1] = __ffma2_rn(v2, make_float2(w.z, w.w), acc[2 * q + 1]);
            }
        }
        uint4 o0, o1;
        o0.x = pack_bf16x2(silu_f(acc[0].x), silu_f(acc[0].y)); o0.y = pack_bf16x2(silu_f(acc[1].x), silu_f(acc[1].y));
        o0.z = pack_bf16x2(silu_f(acc[2].x), silu_f(acc[2].y)); o0.w = pack_bf16x2(silu_f(acc[3].x), silu_f(acc[3].y));
        o1.x = pack_bf16x2(silu_f(acc[4].x), silu_f(acc[4].y)); o1.y = pack_bf16x2(silu_f(acc[5].x), silu_f(acc[5].y));
        o1.z = pack_bf16x2(silu_f(acc[6].x), silu_f(acc[6].y)); o1.w = pack_bf16x2(silu_f(acc[7].x), silu_f(acc[7].y));
        reinterpret_cast<uint4*>(yp + c0)[0] = o0;
        reinterpret_cast<uint4*>(yp + c0)[1] = o1;
    }
}

int stem_conv_launch(const void* x, int x_dtype, int B, int H, int W, const float* w, const float* bias,
                     int Cout, void* y, int y_pixstride, cudaStream_t stream) {
    SY_CHECK(Cout % 16 == 0 && Cout <= 128, SPECYOLO_ERR_INVALID, "stem Cout must be a multiple of 16 (<=128)");
    SY_CHECK(y_pixstride % 8 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0, SPECYOLO_ERR_INVALID,
             "stem output must be 16-byte aligned");
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    dim3 grid((unsigned)ceil_div(Wo, kStemTW), (unsigned)ceil_div(Ho, kStemTH), (unsigned)B);
    const size_t smem = (size_t)(28 * Cout + 3 * kStemIH * 2 * kStemIWH) * sizeof(float);
    __nv_bfloat16* yy = reinterpret_cast<__nv_bfloat16*>(y);
    if (x_dtype == SPECYOLO_DT_F32)
        stem_conv_kernel<float><<<grid, 256, smem, stream>>>((const float*)x, B, H, W, w, bias, Cout, yy, y_pixstride, Ho, Wo);
    else if (x_dtype == SPECYOLO_DT_BF16)
        stem_conv_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>((const __nv_bfloat16*)x, B, H, W, w, bias, Cout, yy, y_pixstride, Ho, Wo);
    else if (x_dtype == SPECYOLO_DT_U8)
        stem_conv_kernel<uint8_t><<<grid, 256, smem, stream>>>((const uint8_t*)x, B, H, W, w, bias, Cout, yy, y_pixstride, Ho, Wo);
    else
        SY_CHECK(false, SPECYOLO_ERR_INVALID, "bad x_dtype %d", x_dtype);
    SY_LAUNCH_CHECK();
    count_launch();
    return SPECYOLO_OK;
}

// ------------------------------------------------------------------------------------------------
// Stem, tensor-core route: space-to-depth.  A 3x3 / stride-2 / pad-1 conv over 3 channels equals a 2x2 / stride-1
// conv (taps at block offsets -1, 0) over the 2x2-blocked image with 12 channels (dy, dx, c); padded to 16 channels
// that is a K = 64 implicit GEMM the tcgen05 kernels run at HBM speed, instead of 27 FMAs per output on CUDA cores.
// This kernel writes the blocked NHWC bf16 tensor y[b, Y, X, (dy*2+dx)*3 + c] = x[b, c, 2Y+dy, 2X+dx]; uint8 input is
// stored as its integer value (exact in bf16; the 1/255 of predictor.py:133-135 is folded into the weights).
// One thread per output pixel: 32-byte (2 x 16 B) stores, 2-element loads that coalesce along the row.
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float load_raw(const T* p, size_t i);
template <>
__device__ __forceinline__ float load_raw<float>(const float* p, size_t i) { return __ldg(p + i); }
template <>
__device__ __forceinline__ float load_raw<__nv_bfloat16>(const __nv_bfloat16* p, size_t i) { return __bfloat162float(p[i]); }
template <>
__device__ __forceinline__ float load_raw<uint8_t>(const uint8_t* p, size_t i) { return (float)__ldg(p + i); }

template <typename T>
__global__ void __launch_bounds__(256)
stem_s2d_kernel(const T* __restrict__ x, int B, int H, int W, __nv_bfloat16* __restrict__ y, int y_pixstride) {
    ptx::grid_dep_launch();
    ptx::grid_dep_wait();
    const int Wo = W >> 1, Ho = H >> 1;
    const size_t total = (size_t)B * Ho * Wo;
    const size_t plane = (size_t)H * W;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int X = (int)(idx % Wo);
        const size_t r = idx / Wo;
        const int Y = (int)(r % Ho);
        const size_t n = r / Ho;
        float v[12];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                const size_t off = (n * 3 + c) * plane + (size_t)(2 * Y + dy) * W + 2 * X;
                v[(dy * 2 + 0) * 3 + c] = load_raw<T>(x, off);
                v[(dy * 2 + 1) * 3 + c] = load_raw<T>(x, off + 1);
            }
        uint4 o0, o1;
        o0.x = pack_bf16x2(v[0], v[1]);   o0.y = pack_bf16x2(v[2], v[3]);
        o0.z = pack_bf16x2(v[4], v[5]);   o0.w = pack_bf16x2(v[6], v[7]);
        o1.x = pack_bf16x2(v[8], v[9]);   o1.y = pack_bf16x2(v[10], v[11]);
        o1.z = 0u;                        o1.w = 0u;
        uint4* yp = reinterpret_cast<uint4*>(y + idx * y_pixstride);
        yp[0] = o0;
        yp[1] = o1;
    }
}

int stem_s2d_launch(const void* x, int x_dtype, int B, int H, int W, void* y, int y_pixstride, cudaStream_t stream) {
    SY_CHECK(H % 2 == 0 && W % 2 == 0, SPECYOLO_ERR_INVALID, "stem space-to-depth needs even H and W");
    SY_CHECK(y_pixstride >= 16 && y_pixstride % 8 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0, SPECYOLO_ERR_INVALID,
             "stem space-to-depth output must be a 16-byte aligned NHWC window of >= 16 channels");
    const size_t total = (size_t)B * (H / 2) * (W / 2);
    size_t want = (total + 255) / 256;
    const size_t cap = (size_t)sm_count() * 16;
    const unsigned blocks = (unsigned)(want < cap ? want : cap);
    __nv_bfloat16* yy = reinterpret_cast<__nv_bfloat16*>(y);
    if (x_dtype == SPECYOLO_DT_F32)
        SY_CUDA(launch_pdl(stem_s2d_kernel<float>, dim3(blocks), dim3(256), 0, stream, (const float*)x, B, H, W, yy, y_pixstride));
    else if (x_dtype == SPECYOLO_DT_BF16)
        SY_CUDA(launch_pdl(stem_s2d_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, stream, (const __nv_bfloat16*)x, B, H, W, yy, y_pixstride));
    else if (x_dtype == SPECYOLO_DT_U8)
        SY_CUDA(launch_pdl(stem_s2d_kernel<uint8_t>, dim3(blocks), dim3(256), 0, stream, (const uint8_t*)x, B, H, W, yy, y_pixstride));
    else
        SY_CHECK(false, SPECYOLO_ERR_INVALID, "bad x_dtype %d", x_dtype);
    SY_LAUNCH_CHECK();
    count_launch();
    return SPECYOLO_OK;
}

// ------------------------------------------------------------------------------------------------
// Depthwise 3x3 stride-1 pad-1 (+bias, optional SiLU): DWConv in Detect.cv3 (head.py:51-52).
// One thread = 8 channels of one pixel (16-byte loads/stores); weights are packed bf16 [C][3][3][1]
// by fold_pack (groups=C, cin_g=1, n_pad=1) -> read as wp[c*9 + tap].
// ------------------------------------------------------------------------------------------------
// Each thread owns 4 channels (one 8-byte vector; a warp covers 128 contiguous channels = one 256-byte pixel row) of
// a run of kDwRun consecutive pixels of one image row and slides the 3x3 window along it: 3 new vector loads per
// output instead of 9.  The 36 folded weights + 4 biases of its channels live in REGISTERS (fp32) — with them in
// shared memory every FMA needed its own LDS and the kernel ran at 14 % of HBM bandwidth, LSU-bound.
template <int kDwRun, int kMinBlocks>
__global__ void __launch_bounds__(256, kMinBlocks)
dwconv3x3_kernel(const __nv_bfloat16* __restrict__ x, int x_pixstride, int B, int H, int W, int C,
                 const __nv_bfloat16* __restrict__ wp, const float* __restrict__ bias, int act,
                 __nv_bfloat16* __restrict__ y, int y_pixstride, FastDiv d_cg, FastDiv d_runs, FastDiv d_h,
                 uint32_t total) {
    ptx::grid_dep_launch();
    ptx::grid_dep_wait();
    float wt[9][4], bs[4];
    uint32_t cached = 0xffffffffu;
    // persistent grid: every thread walks many runs and keeps its channels (the grid stride is a multiple of C/4
    // whenever C/4 divides 256), so the weights are fetched once per thread, not once per run
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        uint32_t r, c4, rw, h, n;
        fdivmod(idx, d_cg, r, c4);
        fdivmod(r, d_runs, r, rw);
        fdivmod(r, d_h, n, h);
        const int c0 = (int)c4 * 4;
        if (c4 != cached) {
            cached = c4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                bs[j] = bias[c0 + j];
#pragma unroll
                for (int t = 0; t < 9; ++t) wt[t][j] = __bfloat162float(wp[(c0 + j) * 9 + t]);
            }
        }
        const int w0 = (int)rw * kDwRun;
        // the whole 3 x (run + 2) input window is requested up front (independent 8-byte loads in flight: enough
        // outstanding bytes per SM to cover HBM latency), kept packed, and unpacked column by column
        uint2 win[3][kDwRun + 2];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int ih = (int)h + ky - 1;
            const bool row_ok = ih >= 0 && ih < H;
            const __nv_bfloat16* rowp = x + (((size_t)n * H + (row_ok ? ih : 0)) * W) * x_pixstride + c0;
#pragma unroll
            for (int i = 0; i < kDwRun + 2; ++i) {
                const int iw = w0 + i - 1;
                win[ky][i] = (row_ok && iw >= 0 && iw < W)
                                 ? __ldg(reinterpret_cast<const uint2*>(rowp + (size_t)iw * x_pixstride))
                                 : make_uint2(0, 0);
            }
        }
        float col[3][3][4];        // ring of unpacked columns w-1, w, w+1
        auto unpack_col = [&](int slot, int i) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const float2 f0 = unpack_bf16x2(win[ky][i].x), f1 = unpack_bf16x2(win[ky][i].y);
                col[slot][ky][0] = f0.x; col[slot][ky][1] = f0.y;
                col[slot][ky][2] = f1.x; col[slot][ky][3] = f1.y;
            }
        };
        unpack_col(0, 0);
        unpack_col(1, 1);
        __nv_bfloat16* yrow = y + (((size_t)n * H + h) * W) * y_pixstride + c0;
#pragma unroll
        for (int i = 0; i < kDwRun; ++i) {
            const int w = w0 + i;
            unpack_col((i + 2) % 3, i + 2);
            if (w < W) {
                float acc[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[j] = bs[j];
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            acc[j] = fmaf(col[(i + kx) % 3][ky][j], wt[ky * 3 + kx][j], acc[j]);
                if (act == SPECYOLO_ACT_SILU) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[j] = silu_tanh(acc[j]);
                }
                uint2 o;
                o.x = pack_bf16x2(acc[0], acc[1]);
                o.y = pack_bf16x2(acc[2], acc[3]);
                *reinterpret_cast<uint2*>(yrow + (size_t)w * y_pixstride) = o;
            }
        }
    }
}

int dwconv3x3_launch(const specyolo_conv_t* a, cudaStream_t stream) {
    SY_CHECK(a->Cin == a->Cout && a->groups == a->Cin && a->kh == 3 && a->kw == 3 && a->stride == 1 &&
                 a->pad == 1 && a->dil == 1,
             SPECYOLO_ERR_UNSUPPORTED, "depthwise kernel supports 3x3 s1 p1 d1 only");
    SY_CHECK(a->Cin % 4 == 0 && a->x_pixstride % 4 == 0 && a->y_pixstride % 4 == 0, SPECYOLO_ERR_INVALID,
             "depthwise needs C, strides multiples of 4");
    SY_CHECK(((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->y)) & 7) == 0,
             SPECYOLO_ERR_INVALID, "depthwise needs 8-byte aligned tensors");
    SY_CHECK(!a->y_fp32 && !a->residual && a->n_pad == 1, SPECYOLO_ERR_UNSUPPORTED,
             "depthwise: bf16 output, no residual, n_pad==1");
    static int variant = -1;
    if (variant < 0) { const char* e = getenv("SPECYOLO_DW"); variant = e ? atoi(e) : 0; }
    const int run = variant == 1 ? 8 : 4;
    const int cg = a->Cin / 4;
    const int runs_w = (a->W + run - 1) / run;
    const long total = (long)a->B * a->H * runs_w * cg;
    SY_CHECK(total < (1L << 31) && fastdiv_ok((uint64_t)total, (uint32_t)(cg > a->H ? cg : a->H)) &&
                 fastdiv_ok((uint64_t)total, (uint32_t)runs_w),
             SPECYOLO_ERR_UNSUPPORTED, "depthwise: tensor too large");
    const long want = (total + 255) / 256;
    const long cap = (long)sm_count() * (variant == 1 ? 1 : 2);
    const int blocks = (int)(want < cap ? want : cap);
    const __nv_bfloat16* xx = reinterpret_cast<const __nv_bfloat16*>(a->x);
    const __nv_bfloat16* ww = reinterpret_cast<const __nv_bfloat16*>(a->w_packed);
    __nv_bfloat16* yy = reinterpret_cast<__nv_bfloat16*>(a->y);
    const FastDiv d_cg = make_fastdiv((uint32_t)cg), d_runs = make_fastdiv((uint32_t)runs_w), d_h = make_fastdiv((uint32_t)a->H);
    if (variant == 1)
        SY_CUDA(launch_pdl(dwconv3x3_kernel<8, 1>, dim3(blocks), dim3(256), 0, stream, xx, a->x_pixstride, a->B, a->H, a->W, a->Cin, ww, a->bias, a->act, yy,
                                                           a->y_pixstride, d_cg, d_runs, d_h, (uint32_t)total));
    else
        SY_CUDA(launch_pdl(dwconv3x3_kernel<4, 2>, dim3(blocks), dim3(256), 0, stream, xx, a->x_pixstride, a->B, a->H, a->W, a->Cin, ww, a->bias, a->act, yy,
                                                           a->y_pixstride, d_cg, d_runs, d_h, (uint32_t)total));
    SY_LAUNCH_CHECK();
    count_launch();
    return SPECYOLO_OK;
}

// ------------------------------------------------------------------------------------------------
// NCHW (f32|bf16|u8) -> NHWC bf16 through a shared-memory transpose tile: reads are coalesced along
// W, writes along C.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const T* __restrict__ x, float scale, int C, long HW, __nv_bfloat16* __restrict__ y,
                    int y_pixstride) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const long p0 = (long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j;
        const long pp = p0 + tx;
        float v = 0.f;
        if (c < C && pp < HW) v = load_in<T>(x, ((long)n * C + c) * HW + pp);
        tile[j][tx] = v;
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const long pp = p0 + j;
        const int c = c0 + tx;
        if (c < C && pp < HW) y[((long)n * HW + pp) * y_pixstride + c] = __float2bfloat16_rn(tile[tx][j] * scale);
    }
}
int nchw_to_nhwc_launch(const void* x, int x_dtype, float scale, int B, int C, int H, int W, void* y,
                        int y_pixstride, cudaStream_t stream) {
    const long HW = (long)H * W;
    dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)B);
    __nv_bfloat16* yy = reinterpret_cast<__nv_bfloat16*>(y);
    // note: load_in<uint8_t> already applies 1/255; `scale` multiplies on top (pass 1.0 for u8)
    if (x_dtype == SPECYOLO_DT_F32)
        nchw_to_nhwc_kernel<float><<<grid, 256, 0, stream>>>((const float*)x, scale, C, HW, yy, y_pixstride);
    else if (x_dtype == SPECYOLO_DT_BF16)
        nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, scale, C, HW, yy, y_pixstride);
    else if (x_dtype == SPECYOLO_DT_U8)
        nchw_to_nhwc_kernel<uint8_t><<<grid, 256, 0, stream>>>((const uint8_t*)x, scale, C, HW, yy, y_pixstride);
    else
        SY_CHECK(false, SPECYOLO_ERR_INVALID, "bad x_dtype %d", x_dtype);
    SY_LAUNCH_CHECK();
    count_launch();
    return SPECYOLO_OK;
}

__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, int x_pixstride, int C, long HW,
                    float* __restrict__ y) {
    __shared__ float tile[32][33];
    const int n = blockIdx.z;
    const long p0 = (long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int j = ty; j < 32; j += 8) {
        const long pp = p0 + j;
        const int c = c0 + tx;
        float v = 0.f;
        if (c < C && pp < HW) v = __bfloat162float(x[((long)n * HW + pp) * x_pixstride + c]);
        tile[j][tx] = v;  // [pixel][channel]
    }
    __syncthreads();
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j;
        const long pp = p0 + tx;
        if (c < C && pp < HW) y[((long)n * C + c) * HW + pp] = tile[tx][j];
    }
}

int nhwc_to_nchw_launch(const void* x, int x_pixstride, int B, int C, int H, int W, float* y,
                        cudaStream_t stream) {
    const long HW = (long)H * W;
    dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)B);
    nhwc_to_nchw_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), x_pixstride, C, HW, y);
    SY_LAUNCH_CHECK();
    count_launch();
    return SPECYOLO_OK;
}

// ------------------------------------------------------------------------------------------------
// nn.Upsample(None, 2, 'nearest') on an NHWC bf16 window, written straight into a channel window of the consumer's
// concat buffer (stock YOLO11 necks: Upsample -> Concat, cfg yolo11.yaml head; the Spectrogram cfg reads through the
// upsample inside Fusion instead).  A thread moves one 16-byte vector (8 channels) of a source pixel to its four
// destinations; consecutive threads take consecutive vectors of a pixel, so reads and writes are full lines.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
upsample2x_kernel(const __nv_bfloat16* __restrict__ x, int x_pixstride, int H, int W, int vecs, long total,
                  __nv_bfloat16* __restrict__ y, int y_pixstride) {
    ptx::grid_dep_launch();
    ptx::grid_dep_wait();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int v = (int)(i % vecs);
        const long pix = i / vecs;                      // source pixel (b, h, w)
        const int w = (int)(pix % W);
        const long bh = pix / W;                        // b * H + h
        const uint4 val = *reinterpret_cast<const uint4*>(x + pix * x_pixstride + v * 8);
        __nv_bfloat16* d = y + ((bh * 2) * (2L * W) + 2 * w) * y_pixstride + v * 8;     // (b, 2h, 2w)
        const long row = 2L * W * y_pixstride;
        *reinterpret_cast<uint4*>(d) = val;
        *reinterpret_cast<uint4*>(d + y_pixstride) = val;
        *reinterpret_cast<uint4*>(d + row) = val;
        *reinterpret_cast<uint4*>(d + row + y_pixstride) = val;
    }
}

int upsample2x_launch(const void* x, int x_pixstride, int B, int H, int W, int C, void* y, int y_pixstride,
                      cudaStream_t stream) {
    const int vecs = C / 8;
    const long total = (long)B * H * W * vecs;
    const long blocks = (total + 255) / 256;
    const unsigned grid = (unsigned)(blocks < 148L * 16 ? blocks : 148L * 16);
    SY_CUDA(launch_pdl(upsample2x_kernel, dim3(grid), dim3(256), 0, stream, reinterpret_cast<const __nv_bfloat16*>(x),
                       x_pixstride, H, W, vecs, total, reinterpret_cast<__nv_bfloat16*>(y), y_pixstride));
    SY_LAUNCH_CHECK();
    count_launch();
    return SPECYOLO_OK;
}

}  // namespace specyolo
