// BottleNect / FGM (ultralytics/nn/modules/block.py:782-861), the inner block of C3k2GC (block.py:1706-1714) in the *_GC
// sibling config (backbone layer 2, 32 channels at 160 x 160 for a 640 x 640 input).  Per image, with x [C, H, W]:
//
//   out   = gelu(in_conv(x))                                      1x1 conv + bias, exact (erf) GELU
//   x_att = fac_conv(mean_hw(out))                                [C]
//   x_fca = | ifft2( x_att * fft2(out) ) |  =  |x_att| * |out|    x_att is constant over the plane, so the transform pair is
//                                                                 the identity (the reference's cuFFT round trip differs
//                                                                 from it by fp32 rounding only); no FFT is run for it
//   x_sca = conv(mean_hw(x_fca)) * x_fca    =  q * |out|,  q = conv(|x_att| * mean_hw(|out|)) * |x_att|
//   FGM:    x1 = dwconv1(x_sca), x2 = dwconv2(x_sca)              1x1 convs + bias
//           o  = | ifft2( x1 * fft2(x2) ) |                       a REAL transform pair: x1 multiplies the spectrum
//   y     = relu( o * alpha + x_sca * beta )
//
// The reference calls cuFFT for the two FFT pairs; here |ifft2(Y)| = |fft2(conj Y)| / (H W), so one forward 2-D transform
// routine serves both directions.  Launches:
//   gc_stats_kernel   per (image, pixel chunk): sum_hw(out), sum_hw(|out|) per channel, fixed-order partials
//   gc_prep_kernel    every CTA folds the partials of its image into q (two C x C mat-vecs), then per pixel recomputes out
//                     and writes x_sca, x1, x2 as PLANAR fp32 planes (the transform works plane by plane)
//   gc_fft_kernel     persistent (32 warps), one plane at a time, the complex plane resident in shared memory (H x (W+1) x 8 bytes:
//                     206 KB at 160 x 160): row FFTs, column FFTs, multiply by x1 and conjugate, row FFTs, column FFTs,
//                     modulus, alpha / beta / relu.  A warp owns a row / column: sides that are multiples of 32 (<= 256)
//                     are transformed in registers (M-point DFT per lane + a 32-point transform across the lanes by
//                     shuffles), any other side by mixed-radix Stockham autosort stages (radix 4, 2 and the odd primes up to 19) between the
//                     strided line in the plane and a per-warp scratch line.
//   gc_pack_kernel    planar fp32 -> bf16 NHWC channel window
// All reductions run in a fixed order: results are bit-identical run to run.
#include "common.h"
#include "tma_host.h"

namespace specyolo {

static constexpr int kGcThreads = 256;
static constexpr int kFftThreads = 1024;
static constexpr int kFftWarps = kFftThreads / 32;
static constexpr int kMaxStages = 8;

struct GcParams {
    specyolo_bottlenect_t a;
    float* partial;        // [B][nchunk][2][C]
    float* xs;             // [B][C][H][W]  x_sca, overwritten by the block's result
    float* x1;             // [B][C][H][W]
    float* x2;             // [B][C][H][W]
    int nchunk;
    int HW;
    float inv_hw;
    int pitch;             // complex elements per plane row in shared memory (W + 1)
    int nst_w, nst_h;
    int rad_w[kMaxStages], rad_h[kMaxStages];
    int maxn;
};

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f)); }

// The 1x1 convs run with LANE = OUTPUT CHANNEL: a lane keeps its weight row in registers and reads the pixel's C inputs
// as shared-memory broadcasts (a thread per pixel would need the whole C x C matrix per output pixel from shared memory
// and C live outputs: it spilled).  C = 16 packs two pixels into a warp.
template <int C>
struct GcTile {
    static constexpr int PPW = 32 / C;                 // pixels per warp step
    static constexpr int VPP = C / 8;                  // 16-byte vectors per pixel
    static constexpr int P = kGcThreads / VPP;         // pixels per CTA tile (every thread stages one vector)
    static constexpr int STEPS = P / (kGcThreads / 32) / PPW;
};

template <int C>
__device__ __forceinline__ void gc_stage_tile(const __nv_bfloat16* xb, int pixstride, int tile0, int hi, float* s_x) {
    using T = GcTile<C>;
    const int pix = threadIdx.x / T::VPP, v = threadIdx.x % T::VPP;
    uint4 q = make_uint4(0, 0, 0, 0);
    if (tile0 + pix < hi) q = *reinterpret_cast<const uint4*>(xb + (size_t)(tile0 + pix) * pixstride + v * 8);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    float4* dst = reinterpret_cast<float4*>(s_x + pix * C + v * 8);
    dst[0] = make_float4(__uint_as_float(w[0] << 16), __uint_as_float(w[0] & 0xffff0000u), __uint_as_float(w[1] << 16), __uint_as_float(w[1] & 0xffff0000u));
    dst[1] = make_float4(__uint_as_float(w[2] << 16), __uint_as_float(w[2] & 0xffff0000u), __uint_as_float(w[3] << 16), __uint_as_float(w[3] & 0xffff0000u));
}

// packed pairs (FFMA2): the mat-vecs are issue-bound; two partial sums, added in a fixed order
template <int C>
__device__ __forceinline__ float gc_dot(const float (&w)[C], float bias, const float* x) {
    float2 t0 = make_float2(bias, 0.f), t1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < C; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(x + j);
        t0 = ffma2(make_float2(w[j], w[j + 1]), make_float2(v.x, v.y), t0);
        t1 = ffma2(make_float2(w[j + 2], w[j + 3]), make_float2(v.z, v.w), t1);
    }
    return (t0.x + t1.x) + (t0.y + t1.y);
}

template <int C>
__global__ void __launch_bounds__(kGcThreads)
gc_stats_kernel(const __grid_constant__ GcParams p) {
    using T = GcTile<C>;
    const specyolo_bottlenect_t& a = p.a;
    __shared__ __align__(16) float s_x[T::P * C];
    __shared__ float s_red[kGcThreads / 32][2 * C];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.y, chunk = blockIdx.x;
    const int c = lane % C, sub = lane / C;
    float w[C];
#pragma unroll
    for (int j = 0; j < C; ++j) w[j] = __ldg(a.in_w + c * C + j);
    const float bias = __ldg(a.in_b + c);
    const int per = (p.HW + p.nchunk - 1) / p.nchunk;
    const int lo = chunk * per, hi = min(p.HW, lo + per);
    const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(a.x) + (size_t)b * p.HW * a.x_pixstride;
    float s1 = 0.f, s2 = 0.f;
    for (int tile0 = lo; tile0 < hi; tile0 += T::P) {
        __syncthreads();
        gc_stage_tile<C>(xb, a.x_pixstride, tile0, hi, s_x);
        __syncthreads();
#pragma unroll 2
        for (int it = 0; it < T::STEPS; ++it) {
            const int pix = (warp * T::STEPS + it) * T::PPW + sub;
            if (tile0 + pix < hi) {
                const float o = gelu_erf(gc_dot<C>(w, bias, s_x + pix * C));
                s1 += o;
                s2 += fabsf(o);
            }
        }
    }
    if (T::PPW == 2) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
        s2 += __shfl_xor_sync(0xffffffffu, s2, 16);
    }
    if (lane < C) { s_red[warp][c] = s1; s_red[warp][C + c] = s2; }
    __syncthreads();
    if (tid < 2 * C) {
        float t = 0.f;
        for (int wv = 0; wv < kGcThreads / 32; ++wv) t += s_red[wv][tid];
        p.partial[((size_t)b * p.nchunk + chunk) * 2 * C + tid] = t;
    }
}

template <int C>
__global__ void __launch_bounds__(kGcThreads)
gc_prep_kernel(const __grid_constant__ GcParams p) {
    using T = GcTile<C>;
    constexpr int OP = T::P + 1;                       // padded pixel pitch of the output staging planes
    const specyolo_bottlenect_t& a = p.a;
    __shared__ __align__(16) float s_x[T::P * C];
    __shared__ float s_o[3 * C * OP];                 // x_sca, x1, x2 as [plane][channel][pixel]
    __shared__ __align__(16) float s_xs[kGcThreads / 32][T::PPW][C];
    __shared__ float s_v[4][C];                       // mean(out), mean(|out|) -> |x_att|, mean(x_fca) -> q
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.y;
    const int c = lane % C, sub = lane / C;
    if (tid < 2 * C) {
        float t = 0.f;
        for (int ch = 0; ch < p.nchunk; ++ch) t += p.partial[((size_t)b * p.nchunk + ch) * 2 * C + tid];
        s_v[tid / C][tid % C] = t * p.inv_hw;
    }
    __syncthreads();
    if (tid < C) {                                     // x_att = fac_conv(mean(out)); mean(x_fca) = |x_att| * mean(|out|)
        float t = __ldg(a.fac_b + tid);
        for (int j = 0; j < C; ++j) t = fmaf(__ldg(a.fac_w + tid * C + j), s_v[0][j], t);
        s_v[2][tid] = fabsf(t);
    }
    __syncthreads();
    if (tid < C) s_v[3][tid] = s_v[2][tid] * s_v[1][tid];
    __syncthreads();
    if (tid < C) {                                     // q = conv(mean(x_fca)) * |x_att|
        float t = __ldg(a.sca_b + tid);
        for (int j = 0; j < C; ++j) t = fmaf(__ldg(a.sca_w + tid * C + j), s_v[3][j], t);
        s_v[0][tid] = t * s_v[2][tid];
    }
    __syncthreads();
    const float q = s_v[0][c];
    float w0[C], w1[C], w2[C];
#pragma unroll
    for (int j = 0; j < C; ++j) {
        w0[j] = __ldg(a.in_w + c * C + j);
        w1[j] = __ldg(a.dw1_w + c * C + j);
        w2[j] = __ldg(a.dw2_w + c * C + j);
    }
    const float b0 = __ldg(a.in_b + c), b1 = __ldg(a.dw1_b + c), b2 = __ldg(a.dw2_b + c);
    const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(a.x) + (size_t)b * p.HW * a.x_pixstride;
    const size_t pb = (size_t)b * C * p.HW;
    float* const planes[3] = {p.xs, p.x1, p.x2};
    for (int tile0 = blockIdx.x * T::P; tile0 < p.HW; tile0 += gridDim.x * T::P) {
        __syncthreads();                               // previous tile's staging planes written out
        gc_stage_tile<C>(xb, a.x_pixstride, tile0, p.HW, s_x);
        __syncthreads();
#pragma unroll 1
        for (int it = 0; it < T::STEPS; ++it) {
            const int pix = (warp * T::STEPS + it) * T::PPW + sub;
            const float xs = q * fabsf(gelu_erf(gc_dot<C>(w0, b0, s_x + pix * C)));
            s_xs[warp][sub][c] = xs;
            s_o[c * OP + pix] = xs;
            __syncwarp();
            s_o[(C + c) * OP + pix] = gc_dot<C>(w1, b1, s_xs[warp][sub]);
            s_o[(2 * C + c) * OP + pix] = gc_dot<C>(w2, b2, s_xs[warp][sub]);
            __syncwarp();
        }
        __syncthreads();
        for (int i = tid; i < 3 * C * T::P; i += kGcThreads) {
            const int pix = i % T::P, kc = i / T::P;
            if (tile0 + pix < p.HW) planes[kc / C][pb + (size_t)(kc % C) * p.HW + tile0 + pix] = s_o[kc * OP + pix];
        }
    }
}

// ---- 1-D transform of one line by one warp -----------------------------------------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// Stockham autosort stage of radix R on a line of N points: reads src (element stride ss), writes dst (stride ds).
// Ns = product of the radices of the earlier stages; tw[k] = exp(-2 pi i k / N).
template <int R>
__device__ __forceinline__ void fft_stage(const float2* src, int ss, float2* dst, int ds, int N, int Ns, const float2* tw, int lane) {
    const int T = N / R;
    const int tstep = N / (Ns * R);                    // W_{Ns R}^{k t} = tw[k * t * tstep]
    const bool pow2 = (Ns & (Ns - 1)) == 0;            // true for every stage but those after an odd radix (odd radices run last)
    const int sh = __ffs(Ns) - 1;
    for (int j = lane; j < T; j += 32) {
        const int k = pow2 ? (j & (Ns - 1)) : (j % Ns);
        const int jb = pow2 ? (j >> sh) : (j / Ns);
        float2 v[R];
#pragma unroll
        for (int t = 0; t < R; ++t) {
            v[t] = src[(size_t)(j + t * T) * ss];
            if (t > 0) v[t] = cmul(v[t], tw[k * t * tstep]);
        }
        float2* o = dst + (size_t)(jb * Ns * R + k) * ds;
        if (R == 2) {
            o[0] = make_float2(v[0].x + v[1].x, v[0].y + v[1].y);
            o[(size_t)Ns * ds] = make_float2(v[0].x - v[1].x, v[0].y - v[1].y);
        } else if (R == 4) {
            const float2 s02 = make_float2(v[0].x + v[2].x, v[0].y + v[2].y), d02 = make_float2(v[0].x - v[2].x, v[0].y - v[2].y);
            const float2 s13 = make_float2(v[1].x + v[3].x, v[1].y + v[3].y), d13 = make_float2(v[1].x - v[3].x, v[1].y - v[3].y);
            o[0] = make_float2(s02.x + s13.x, s02.y + s13.y);
            o[(size_t)Ns * ds] = make_float2(d02.x + d13.y, d02.y - d13.x);           // d02 - i d13
            o[(size_t)2 * Ns * ds] = make_float2(s02.x - s13.x, s02.y - s13.y);
            o[(size_t)3 * Ns * ds] = make_float2(d02.x - d13.y, d02.y + d13.x);       // d02 + i d13
        } else {
            const int rstep = N / R;                   // W_R^m = tw[m * rstep]
#pragma unroll
            for (int u = 0; u < R; ++u) {
                float2 acc = v[0];
#pragma unroll
                for (int t = 1; t < R; ++t) {
                    const float2 w = tw[((u * t) % R) * rstep];
                    acc.x = fmaf(v[t].x, w.x, fmaf(-v[t].y, w.y, acc.x));
                    acc.y = fmaf(v[t].x, w.y, fmaf(v[t].y, w.x, acc.y));
                }
                o[(size_t)u * Ns * ds] = acc;
            }
        }
    }
}

// Radices 11 ... 19 (352 / 416 / 544 / 608-pixel inputs: plane sides 88, 104, 136, 152): kept out of line so that their
// register appetite (19 complex values + accumulators) does not shape the allocation of the common path.
__device__ __noinline__ void fft_stage_big(int R, const float2* src, int ss, float2* dst, int ds, int N, int Ns, const float2* tw, int lane) {
    if (R == 11) fft_stage<11>(src, ss, dst, ds, N, Ns, tw, lane);
    else if (R == 13) fft_stage<13>(src, ss, dst, ds, N, Ns, tw, lane);
    else if (R == 17) fft_stage<17>(src, ss, dst, ds, N, Ns, tw, lane);
    else fft_stage<19>(src, ss, dst, ds, N, Ns, tw, lane);
}

// In-place forward DFT of the line `data` (element stride `stride`) through the warp's scratch line.
__device__ void fft_line(float2* data, int stride, int N, const int* rad, int nst, const float2* tw, float2* scratch, int lane) {
    float2* src = data; int ss = stride;
    float2* dst = scratch; int ds = 1;
    int Ns = 1;
    for (int s = 0; s < nst; ++s) {
        const int R = rad[s];
        if (R == 4) fft_stage<4>(src, ss, dst, ds, N, Ns, tw, lane);
        else if (R == 2) fft_stage<2>(src, ss, dst, ds, N, Ns, tw, lane);
        else if (R == 3) fft_stage<3>(src, ss, dst, ds, N, Ns, tw, lane);
        else if (R == 5) fft_stage<5>(src, ss, dst, ds, N, Ns, tw, lane);
        else if (R == 7) fft_stage<7>(src, ss, dst, ds, N, Ns, tw, lane);
        else fft_stage_big(R, src, ss, dst, ds, N, Ns, tw, lane);
        __syncwarp();
        float2* t = src; src = dst; dst = t;
        const int ti = ss; ss = ds; ds = ti;
        Ns *= R;
    }
    if (src != data) {                                 // odd number of stages: the result sits in the scratch line
        for (int i = lane; i < N; i += 32) data[(size_t)i * stride] = scratch[i];
        __syncwarp();
    }
}

// ---- register path: lines of N = 32 M points, M <= 8 -------------------------------------------------------------------
// Lane l holds x[l + 32 j], j < M.  N = M x 32 Cooley-Tukey: an M-point DFT over j in registers, the twiddle
// W_N^(l k1), then — for each of the M residues k1 — a 32-point radix-2 DIF transform ACROSS the lanes with shuffles; lane l
// ends up with X[k1 + M * bitrev5(l)], stored to its natural place.  ~3x fewer instructions per line than the Stockham
// stages (no shared-memory round trip per stage, no index arithmetic), used whenever a plane side is a multiple of 32.
template <int M>
__device__ __forceinline__ void fft_line_reg(float2* data, int stride, const float2* tw, int lane) {
    constexpr int N = 32 * M;
    float2 v[M];
#pragma unroll
    for (int j = 0; j < M; ++j) v[j] = data[(lane + 32 * j) * stride];
    if (M > 1) {
        // M-point DFT through the (t, M - t) symmetry: X_u = A_u - i B_u, X_{M-u} = A_u + i B_u
        float2 x[M];
        x[0] = v[0];
#pragma unroll
        for (int t = 1; t < M; ++t) { x[0].x += v[t].x; x[0].y += v[t].y; }
#pragma unroll
        for (int u = 1; u <= M / 2; ++u) {
            float2 A = v[0], B = make_float2(0.f, 0.f);
#pragma unroll
            for (int t = 1; t <= (M - 1) / 2; ++t) {
                const float2 w = tw[((u * t) % M) * 32];            // (cos, -sin)(2 pi u t / M)
                A.x = fmaf(v[t].x + v[M - t].x, w.x, A.x); A.y = fmaf(v[t].y + v[M - t].y, w.x, A.y);
                B.x = fmaf(v[t].x - v[M - t].x, -w.y, B.x); B.y = fmaf(v[t].y - v[M - t].y, -w.y, B.y);
            }
            if (M % 2 == 0) {                                       // the self-paired middle term: (-1)^u v[M/2]
                const float sg = (u & 1) ? -1.f : 1.f;
                A.x = fmaf(sg, v[M / 2].x, A.x); A.y = fmaf(sg, v[M / 2].y, A.y);
            }
            x[u] = make_float2(A.x + B.y, A.y - B.x);               // A - i B
            if (u != M - u) x[M - u] = make_float2(A.x - B.y, A.y + B.x);
        }
#pragma unroll
        for (int u = 0; u < M; ++u) v[u] = u == 0 ? x[0] : cmul(x[u], tw[lane * u]);
    }
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int half = 16 >> s;
        const bool upper = (lane & half) != 0;
        const float sg = upper ? -1.f : 1.f;
        float2 w = make_float2(1.f, 0.f);
        if (s < 4 && upper) w = tw[((lane & (half - 1)) << s) * M];   // W_32^((l mod half) << s)
#pragma unroll
        for (int u = 0; u < M; ++u) {
            const float ox = __shfl_xor_sync(0xffffffffu, v[u].x, half), oy = __shfl_xor_sync(0xffffffffu, v[u].y, half);
            const float2 r = make_float2(fmaf(sg, v[u].x, ox), fmaf(sg, v[u].y, oy));   // lower: v + other, upper: other - v
            v[u] = s < 4 ? cmul(r, w) : r;
        }
    }
    const int k2 = (int)(__brev((unsigned)lane) >> 27);
    __syncwarp();
#pragma unroll
    for (int u = 0; u < M; ++u) data[(u + M * k2) * stride] = v[u];
    __syncwarp();
}

__device__ __forceinline__ void fft_line_any(float2* data, int stride, int N, const int* rad, int nst, const float2* tw, float2* scratch, int lane) {
    if ((N & 31) == 0 && N <= 256) {
        switch (N >> 5) {
            case 1: fft_line_reg<1>(data, stride, tw, lane); break;
            case 2: fft_line_reg<2>(data, stride, tw, lane); break;
            case 3: fft_line_reg<3>(data, stride, tw, lane); break;
            case 4: fft_line_reg<4>(data, stride, tw, lane); break;
            case 5: fft_line_reg<5>(data, stride, tw, lane); break;
            case 6: fft_line_reg<6>(data, stride, tw, lane); break;
            case 7: fft_line_reg<7>(data, stride, tw, lane); break;
            default: fft_line_reg<8>(data, stride, tw, lane); break;
        }
    } else {
        fft_line(data, stride, N, rad, nst, tw, scratch, lane);
    }
}

// linear pixel index -> offset in the padded shared-memory plane; y = floor((i + 0.5) / W) in fp32 is exact for the plane
// sizes that fit shared memory (the quotient sits 0.5 / W away from an integer, fp32 error is < 1e-4 of that)
__device__ __forceinline__ int plane_index(int i, int W, float inv_w, int pitch) {
    const int y = (int)(((float)i + 0.5f) * inv_w);
    return y * pitch + (i - y * W);
}

__global__ void __launch_bounds__(kFftThreads, 1)
gc_fft_kernel(const __grid_constant__ GcParams p) {
    const specyolo_bottlenect_t& a = p.a;
    extern __shared__ __align__(16) uint8_t gc_smem[];
    float2* s_plane = reinterpret_cast<float2*>(gc_smem);                    // [H][pitch]
    float2* s_scr = s_plane + (size_t)a.H * p.pitch;                         // [warps][maxn]
    float2* s_tww = s_scr + (size_t)kFftWarps * p.maxn;                      // [W]
    float2* s_twh = s_tww + a.W;                                             // [H]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < a.W; i += kFftThreads) {
        float sn, cs;
        sincospif(-2.0f * (float)i / (float)a.W, &sn, &cs);
        s_tww[i] = make_float2(cs, sn);
    }
    for (int i = tid; i < a.H; i += kFftThreads) {
        float sn, cs;
        sincospif(-2.0f * (float)i / (float)a.H, &sn, &cs);
        s_twh[i] = make_float2(cs, sn);
    }
    float2* scr = s_scr + (size_t)warp * p.maxn;
    const int planes = a.B * a.C;
    const float inv_w = 1.0f / (float)a.W;
    for (int pl = blockIdx.x; pl < planes; pl += gridDim.x) {
        const size_t off = (size_t)pl * p.HW;
        const float al = __ldg(a.alpha + pl % a.C), be = __ldg(a.beta + pl % a.C);
        __syncthreads();                               // twiddles written / previous plane consumed
        for (int i0 = tid; i0 < p.HW; i0 += 4 * kFftThreads) {      // four global loads in flight per thread
            float g[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) g[u] = i0 + u * kFftThreads < p.HW ? __ldg(p.x2 + off + i0 + u * kFftThreads) : 0.f;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * kFftThreads;
                if (i < p.HW) s_plane[plane_index(i, a.W, inv_w, p.pitch)] = make_float2(g[u], 0.f);
            }
        }
        __syncthreads();
        for (int pass = 0; pass < 2; ++pass) {
            for (int y = warp; y < a.H; y += kFftWarps) fft_line_any(s_plane + (size_t)y * p.pitch, 1, a.W, p.rad_w, p.nst_w, s_tww, scr, lane);
            __syncthreads();
            for (int x = warp; x < a.W; x += kFftWarps) fft_line_any(s_plane + x, p.pitch, a.H, p.rad_h, p.nst_h, s_twh, scr, lane);
            __syncthreads();
            if (pass == 0) {                           // Y = x1 * fft2(x2); |ifft2(Y)| = |fft2(conj Y)| / (H W)
                for (int i0 = tid; i0 < p.HW; i0 += 4 * kFftThreads) {
                    float g[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) g[u] = i0 + u * kFftThreads < p.HW ? __ldg(p.x1 + off + i0 + u * kFftThreads) : 0.f;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int i = i0 + u * kFftThreads;
                        if (i < p.HW) {
                            float2& z = s_plane[plane_index(i, a.W, inv_w, p.pitch)];
                            z = make_float2(g[u] * z.x, -g[u] * z.y);
                        }
                    }
                }
                __syncthreads();
            }
        }
        for (int i0 = tid; i0 < p.HW; i0 += 4 * kFftThreads) {
            float g[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) g[u] = i0 + u * kFftThreads < p.HW ? p.xs[off + i0 + u * kFftThreads] : 0.f;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * kFftThreads;
                if (i < p.HW) {
                    const float2 z = s_plane[plane_index(i, a.W, inv_w, p.pitch)];
                    const float o = sqrtf(z.x * z.x + z.y * z.y) * p.inv_hw;
                    p.xs[off + i] = fmaxf(fmaf(o, al, g[u] * be), 0.f);
                }
            }
        }
    }
}

// planar fp32 [B][C][HW] -> bf16 NHWC channel window
template <int C>
__global__ void __launch_bounds__(kGcThreads)
gc_pack_kernel(const __grid_constant__ GcParams p) {
    const specyolo_bottlenect_t& a = p.a;
    const long npix = (long)a.B * p.HW;
    for (long g = (long)blockIdx.x * kGcThreads + threadIdx.x; g < npix; g += (long)gridDim.x * kGcThreads) {
        const long b = g / p.HW, pix = g - b * p.HW;
        const float* src = p.xs + (size_t)b * C * p.HW + pix;
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(a.y) + (size_t)g * a.y_pixstride;
#pragma unroll
        for (int v = 0; v < C / 8; ++v) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const __nv_bfloat162 h = __floats2bfloat162_rn(src[(size_t)(v * 8 + 2 * e) * p.HW], src[(size_t)(v * 8 + 2 * e + 1) * p.HW]);
                w[e] = *reinterpret_cast<const uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(dst + v * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

static int gc_nchunk(int HW) {
    const int n = (HW + 2047) / 2048;
    return n < 1 ? 1 : (n > 32 ? 32 : n);
}

static bool gc_factor(int n, int* rad, int* nst) {
    int k = 0;
    while (n % 4 == 0 && k < kMaxStages) { rad[k++] = 4; n /= 4; }
    const int primes[8] = {2, 3, 5, 7, 11, 13, 17, 19};
    for (int q = 0; q < 8; ++q)
        while (n % primes[q] == 0 && k < kMaxStages) { rad[k++] = primes[q]; n /= primes[q]; }
    *nst = k;
    return n == 1;
}

static bool gc_reg_path(int n) { return (n & 31) == 0 && n <= 256; }

// per-warp scratch lines exist only when a side takes the Stockham path
static int gc_scratch_len(int H, int W) {
    const int a = gc_reg_path(W) ? 0 : W, b = gc_reg_path(H) ? 0 : H;
    return a > b ? a : b;
}

static size_t gc_fft_smem(int H, int W) {
    return ((size_t)H * (W + 1) + (size_t)kFftWarps * gc_scratch_len(H, W) + W + H) * sizeof(float2);
}

size_t bottlenect_ws_bytes(int B, int H, int W, int C) {
    const size_t HW = (size_t)H * W;
    return ((size_t)B * gc_nchunk((int)HW) * 2 * C + 3 * (size_t)B * C * HW) * sizeof(float);
}

static size_t g_fft_attr_smem[kMaxDevices] = {0};

template <int C>
static int bottlenect_launch_t(GcParams& p, cudaStream_t stream) {
    const specyolo_bottlenect_t& a = p.a;
    gc_stats_kernel<C><<<dim3((unsigned)p.nchunk, (unsigned)a.B), kGcThreads, 0, stream>>>(p);
    count_launch();
    int per_img = ceil_div(p.HW, GcTile<C>::P);
    const int cap = max(1, sm_count() * 4 / a.B);
    if (per_img > cap) per_img = cap;
    gc_prep_kernel<C><<<dim3((unsigned)per_img, (unsigned)a.B), kGcThreads, 0, stream>>>(p);
    count_launch();
    const size_t smem = gc_fft_smem(a.H, a.W);
    SY_CUDA(ensure_dynamic_smem(gc_fft_kernel, smem, g_fft_attr_smem));   // one cache for both channel-count instantiations
    const int planes = a.B * a.C;
    gc_fft_kernel<<<(unsigned)min(planes, sm_count()), kFftThreads, smem, stream>>>(p);
    count_launch();
    const long npix = (long)a.B * p.HW;
    gc_pack_kernel<C><<<(unsigned)min((long)sm_count() * 8, (npix + kGcThreads - 1) / kGcThreads), kGcThreads, 0, stream>>>(p);
    count_launch();
    SY_LAUNCH_CHECK();
    return SPECYOLO_OK;
}

int bottlenect_launch(const specyolo_bottlenect_t* a, cudaStream_t stream) {
    SY_CHECK(a->C == 16 || a->C == 32, SPECYOLO_ERR_UNSUPPORTED, "BottleNect: 16 or 32 channels (scales n, s of the *_GC config), got %d", a->C);
    SY_CHECK(a->B <= 65535, SPECYOLO_ERR_UNSUPPORTED, "BottleNect: batch %d exceeds the grid limit", a->B);
    SY_CHECK(a->x_pixstride % 8 == 0 && a->y_pixstride % 8 == 0 && a->x_pixstride >= a->C && a->y_pixstride >= a->C,
             SPECYOLO_ERR_INVALID, "BottleNect: pixel strides must be multiples of 8 elements and >= C");
    SY_CHECK(((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->y)) & 15) == 0, SPECYOLO_ERR_INVALID,
             "BottleNect: x / y must be 16-byte aligned");
    GcParams p{};
    p.a = *a;
    p.HW = a->H * a->W;
    p.inv_hw = 1.0f / (float)p.HW;
    p.nchunk = gc_nchunk(p.HW);
    p.pitch = a->W + 1;
    p.maxn = gc_scratch_len(a->H, a->W);
    SY_CHECK(gc_factor(a->W, p.rad_w, &p.nst_w) && gc_factor(a->H, p.rad_h, &p.nst_h), SPECYOLO_ERR_UNSUPPORTED,
             "BottleNect: plane %d x %d has a prime factor above 19", a->H, a->W);
    SY_CHECK(gc_fft_smem(a->H, a->W) <= 227 * 1024, SPECYOLO_ERR_UNSUPPORTED,
             "BottleNect: a %d x %d complex plane does not fit shared memory (up to 160 x 160, i.e. a 640 x 640 input)", a->H, a->W);
    const size_t plane_all = (size_t)a->B * a->C * p.HW;
    p.partial = a->ws;
    p.xs = a->ws + (size_t)a->B * p.nchunk * 2 * a->C;
    p.x1 = p.xs + plane_all;
    p.x2 = p.x1 + plane_all;
    return a->C == 16 ? bottlenect_launch_t<16>(p, stream) : bottlenect_launch_t<32>(p, stream);
}

}  // namespace specyolo
