// IQ burst -> Hann-windowed 1024-point STFT -> |X|^2 -> dBFS -> [0,1] -> letterbox to out_h x out_w.
//
// The reference has no such stage (README.md:7 only mentions spectrograms; its dataset YAMLs point
// at pre-rendered images), so the convention is frozen here and restated by oracle/stft_ref.py:
//   frames   t = 0..T-1, T = 1 + (L - nfft)/hop, frame t = samples [t*hop, t*hop + nfft)   (center=False)
//   window   periodic Hann  w[n] = 0.5 - 0.5 cos(2 pi n / nfft)
//   X[k]     = sum_n w[n] x[n] exp(-2 pi i n k / nfft);  image row r = (k + nfft/2) mod nfft (fftshift)
//   dB       = 10 log10( max(|X|^2, 1e-20) / (sum w)^2 )          (0 dBFS = unit-amplitude tone)
//   v        = clamp((dB - db_min) / (db_max - db_min), 0, 1)      image S[r][t], nfft rows x T cols
//   letterbox: LetterBox(new_shape, auto=False, scaleup=True, center=True) geometry
//              (ultralytics/data/augment.py:1566-1591) with half-pixel bilinear sampling (cv2
//              INTER_LINEAR geometry, float arithmetic), pad value 114/255, 3 identical channels.
//
// Only the frames / bins that the bilinear taps touch are computed: at the north-star geometry
// (L = 2^20, hop 256 -> 1024 x 4093 -> 160 x 640) that is 2 frames per output column.  HBM traffic is
// the IQ samples of those frames (each sample read once per CTA, coalesced 8-byte lanes) plus one
// write of the bf16 image: 8.39 + 2.46 MB per burst = the roofline of SURVEY 8(d).
//
// CTA = 128 threads = 4 warps, 16 output columns.  A WARP OWNS AN OUTPUT COLUMN: it transforms the column's two frames
// (bilinear taps t0, t0+1) TOGETHER, every register pair holding the same element of frame 0 and frame 1, so the whole
// transform — 32 x 32 four-step: 32-point radix-2 FFT in registers, twiddle, transpose through a padded shared-memory
// tile, second 32-point FFT — runs on packed fp32 pairs (add/sub/mul/fma.f32x2: one issue slot per two butterflies,
// twiddle constants shared by both halves).  The kernel is issue-bound (round 1: 4.6 k warp instructions per frame at
// one frame per warp, scalar), hence: powers -> shared memory, log2 only for the 2 x new_h bins the vertical taps read,
// row / column tap tables built once per CTA (the double-precision geometry used to be redone per frame and per
// pixel), window and (lane, k1) twiddles read from shared-memory tables laid out for the packed operands, padding
// written with 16-byte stores.
#include "common.h"

namespace specyolo {

static constexpr int NFFT = 1024;
static constexpr int kCols = 16;       // output columns per CTA
static constexpr int kWarps = 4;
static constexpr int kThreads = kWarps * 32;
static constexpr int kTilePitch = 33;  // float2 entries per tile row (conflict-free 8-byte rows and columns)

__host__ __device__ constexpr int bitrev5(int v) {
    return ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4);
}

__device__ __forceinline__ float2 fsub2(float2 a, float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5};\n\t"
        "sub.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5};\n\t"
        "mul.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd;}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}

// W32^j = exp(-2 pi i j / 32), j = 0..15
__device__ constexpr float kW32c[16] = {
    1.000000000e+00f, 9.807852804e-01f, 9.238795325e-01f, 8.314696123e-01f, 7.071067812e-01f, 5.555702330e-01f,
    3.826834324e-01f, 1.950903220e-01f, 0.0f, -1.950903220e-01f, -3.826834324e-01f, -5.555702330e-01f,
    -7.071067812e-01f, -8.314696123e-01f, -9.238795325e-01f, -9.807852804e-01f};
__device__ constexpr float kW32s[16] = {
    -0.0f, -1.950903220e-01f, -3.826834324e-01f, -5.555702330e-01f, -7.071067812e-01f, -8.314696123e-01f,
    -9.238795325e-01f, -9.807852804e-01f, -1.000000000e+00f, -9.807852804e-01f, -9.238795325e-01f,
    -8.314696123e-01f, -7.071067812e-01f, -5.555702330e-01f, -3.826834324e-01f, -1.950903220e-01f};

// In-register radix-2 DIF FFT of 32 complex points of TWO frames at once: re[i] = (Re x0[i], Re x1[i]), im likewise;
// X[k] ends up in element bitrev5(k).  46 of the 80 butterflies need no multiplication (W^0 = 1, W^8 = -i).
__device__ __forceinline__ void fft32x2(float2 (&re)[32], float2 (&im)[32]) {
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int half = 16 >> s;
#pragma unroll
        for (int g = 0; g < (1 << s); ++g) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const int i0 = g * 2 * half + j, i1 = i0 + half;
                const float2 ar = re[i0], ai = im[i0], br = re[i1], bi = im[i1];
                re[i0] = fadd2(ar, br);
                im[i0] = fadd2(ai, bi);
                const int tw = j << s;  // W_{2*half}^j = W32^(j << s)
                if (tw == 0) {
                    re[i1] = fsub2(ar, br);
                    im[i1] = fsub2(ai, bi);
                } else if (tw == 8) {                              // (dr + i di) * (-i) = di - i dr
                    re[i1] = fsub2(ai, bi);
                    im[i1] = fsub2(br, ar);
                } else {
                    const float2 dr = fsub2(ar, br), di = fsub2(ai, bi);
                    const float c = kW32c[tw], sn = kW32s[tw];
                    re[i1] = ffma2(di, make_float2(-sn, -sn), fmul2(dr, make_float2(c, c)));
                    im[i1] = ffma2(dr, make_float2(sn, sn), fmul2(di, make_float2(c, c)));
                }
            }
        }
    }
}

struct StftParams {
    specyolo_stft_t a;
    int T;                 // frames
    int new_w, new_h;      // content size after resize
    int left, top;         // content origin inside the output
    double sx, sy;         // source/dest scale (T/new_w, nfft/new_h)
    int col_tiles;         // CTAs (per burst) doing content columns
    int pad_tiles;         // CTAs (per burst) filling padding (first in the grid: they are the longest-running)
    int out_pitch;         // floats per column of the CTA's staging tile (== 2 mod 32: conflict-free both ways)
    float db_scale, db_off;  // v = 10 log10(pw) * db_scale + db_off
    const float2* twiddle; // [1024] exp(-2 pi i j / 1024)
};

__device__ __forceinline__ void store_px(const StftParams& p, size_t idx, float v) {
    if (p.a.out_fp32) reinterpret_cast<float*>(p.a.out)[idx] = v;
    else reinterpret_cast<__nv_bfloat16*>(p.a.out)[idx] = __float2bfloat16_rn(v);
}

// Everything of one burst's image that lies outside the content rectangle.  The rows above and below the band are two
// contiguous spans per plane: plain 16-byte store loops shared out over the burst's pad CTAs; the strips left and right
// of the band go row by row.  (The first version divided 64-bit indices per element and outlasted the transform.)
__device__ void fill_padding(const StftParams& p, int b, int pt) {
    const specyolo_stft_t& a = p.a;
    const int tid = threadIdx.x;
    const size_t plane = (size_t)a.out_h * a.out_w;
    const size_t base = (size_t)b * 3 * plane;
    const int vec = a.out_fp32 ? 4 : 8;            // elements per 16-byte store
    const int esize = a.out_fp32 ? 4 : 2;
    const bool vec_ok = a.out_w % vec == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0;
    uint4 fill;
    if (a.out_fp32) {
        fill.x = fill.y = fill.z = fill.w = __float_as_uint(a.pad_value);
    } else {
        const uint32_t h = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a.pad_value));
        fill.x = fill.y = fill.z = fill.w = h | (h << 16);
    }
    const int me = pt * kThreads + tid, nthr = p.pad_tiles * kThreads;
    for (int s = 0; s < 6; ++s) {                  // (plane, above / below)
        const int y0 = (s & 1) ? p.top + p.new_h : 0, y1 = (s & 1) ? a.out_h : p.top;
        if (y1 <= y0) continue;
        const size_t off = base + (size_t)(s >> 1) * plane + (size_t)y0 * a.out_w;
        const int n = (y1 - y0) * a.out_w;         // elements of the span
        if (vec_ok) {
            uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(a.out) + off * esize);
            const int nv = n / vec;
#pragma unroll 4
            for (int i = me; i < nv; i += nthr) dst[i] = fill;
        } else {
            for (int i = me; i < n; i += nthr) store_px(p, off + i, a.pad_value);
        }
    }
    if (p.left == 0 && p.new_w == a.out_w) return;
    const int right0 = p.left + p.new_w, strip = p.left + (a.out_w - right0);   // pad elements per band row
    for (int rr = pt; rr < 3 * p.new_h; rr += p.pad_tiles) {
        const int pl = rr / p.new_h, y = p.top + rr % p.new_h;
        const size_t row = base + (size_t)pl * plane + (size_t)y * a.out_w;
        for (int i = tid; i < strip; i += kThreads) store_px(p, row + (i < p.left ? i : right0 + (i - p.left)), a.pad_value);
    }
}

__global__ void __launch_bounds__(kThreads, 3)
stft_letterbox_kernel(const __grid_constant__ StftParams p) {
    extern __shared__ __align__(16) uint8_t st_smem[];
    const specyolo_stft_t& a = p.a;
    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t plane = (size_t)a.out_h * a.out_w;

    if ((int)blockIdx.x < p.pad_tiles) {
        fill_padding(p, b, blockIdx.x);
        return;
    }

    // shared memory: per-warp transpose tile | (register, lane) twiddles | window | row taps | column taps | staging
    float2* s_tile = reinterpret_cast<float2*>(st_smem) + warp * (32 * kTilePitch);     // [32][33] (frame 0, frame 1) of one component
    float4* s_tw2 = reinterpret_cast<float4*>(reinterpret_cast<float2*>(st_smem) + kWarps * (32 * kTilePitch));  // [32][32] {wr, wr, wi, wi}
    float* s_win = reinterpret_cast<float*>(s_tw2 + 32 * 32);                            // [1024]
    int2* s_rowtab = reinterpret_cast<int2*>(s_win + NFFT);                              // [new_h] {r0 | r1 << 16, wy}
    int2* s_coltab = s_rowtab + p.new_h;                                                 // [kCols] {t0, wx}
    float* s_out = reinterpret_cast<float*>(s_coltab + kCols);                           // [kCols][out_pitch]

    const int x0 = ((int)blockIdx.x - p.pad_tiles) * kCols;   // first content column of this CTA
    const int ncols = min(kCols, p.new_w - x0);
    for (int i = tid; i < 32 * 32; i += kThreads) {
        const int pp = i >> 5, n2 = i & 31;                   // register pp of lane n2 holds Y[k1 = bitrev5(pp)][n2]
        const float2 w = p.twiddle[(n2 * bitrev5(pp)) & (NFFT - 1)];
        s_tw2[i] = make_float4(w.x, w.x, w.y, w.y);
    }
    for (int i = tid; i < NFFT; i += kThreads) s_win[i] = 0.5f - 0.5f * p.twiddle[i].x;   // periodic Hann
    for (int ys = tid; ys < p.new_h; ys += kThreads) {
        double fy = ((double)ys + 0.5) * p.sy - 0.5;
        if (fy < 0.0) fy = 0.0;
        int r0 = (int)floor(fy);
        if (r0 > NFFT - 1) r0 = NFFT - 1;
        const int r1 = min(r0 + 1, NFFT - 1);
        s_rowtab[ys] = make_int2(r0 | (r1 << 16), __float_as_int((float)(fy - (double)r0)));
    }
    if (tid < kCols) {
        double fx = ((double)(x0 + tid) + 0.5) * p.sx - 0.5;
        if (fx < 0.0) fx = 0.0;
        int t0 = (int)floor(fx);
        if (t0 > p.T - 1) t0 = p.T - 1;
        s_coltab[tid] = make_int2(t0, __float_as_int((float)(fx - (double)t0)));
    }
    __syncthreads();

    const float2* iq = reinterpret_cast<const float2*>(a.iq) + (size_t)b * a.L;
    const float lg_scale = 3.010299956639812f * p.db_scale;   // 10 log10(pw) = 3.0103 log2(pw)

#pragma unroll 1
    for (int c = warp; c < ncols; c += kWarps) {
        const int t0 = s_coltab[c].x;
        const float wx = __int_as_float(s_coltab[c].y);
        const int t1 = min(t0 + 1, p.T - 1);
        const float2* src0 = iq + (size_t)t0 * a.hop;
        const float2* src1 = iq + (size_t)t1 * a.hop;

        // ---- stage A: lane = n2, register = n1; x[n1] = w[n] s[n], n = 32 n1 + n2 ----
        float2 re[32], im[32];
        if (c + kWarps < ncols) {                      // next column of this warp: pull its samples towards L2 while this one computes
            const int tn = s_coltab[c + kWarps].x;
            const char* nx = reinterpret_cast<const char*>(iq + (size_t)tn * a.hop);
            const int bytes = (int)min((long)(NFFT + a.hop), (long)a.L - (long)tn * a.hop) * 8;   // both frames, inside the burst
            for (int o = lane * 128; o < bytes; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(nx + o));
        }
        if (a.hop == 256 && t1 == t0 + 1) {
            // frame 1 is frame 0 advanced by 256 samples = 8 registers: 40 loads instead of 64
            float2 raw[40];
#pragma unroll
            for (int n1 = 0; n1 < 40; ++n1) raw[n1] = __ldg(src0 + 32 * n1 + lane);
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) {
                const float w = s_win[32 * n1 + lane];
                re[n1] = make_float2(raw[n1].x * w, raw[n1 + 8].x * w);
                im[n1] = make_float2(raw[n1].y * w, raw[n1 + 8].y * w);
            }
        } else {
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) {
                const int n = 32 * n1 + lane;
                const float2 s0 = __ldg(src0 + n), s1 = __ldg(src1 + n);
                const float w = s_win[n];
                re[n1] = make_float2(s0.x * w, s1.x * w);
                im[n1] = make_float2(s0.y * w, s1.y * w);
            }
        }
        fft32x2(re, im);
        __syncwarp();
        // twiddle W1024^(n2*k1), then the 32 x 32 transpose (register k1 of lane n2 -> register n2 of lane k1), real parts
        // of both frames first, then the imaginary parts, through an 8-byte-entry tile (register pairs move as they are;
        // a 16-byte tile would cost the third resident CTA)
#pragma unroll
        for (int pp = 0; pp < 32; ++pp) {
            const float4 w = s_tw2[pp * 32 + lane];
            const float2 wr = make_float2(w.x, w.y), wi = make_float2(w.z, w.w);
            const float2 vr = fsub2(fmul2(re[pp], wr), fmul2(im[pp], wi));
            im[pp] = ffma2(im[pp], wr, fmul2(re[pp], wi));
            re[pp] = vr;
        }
#pragma unroll
        for (int pp = 0; pp < 32; ++pp) s_tile[bitrev5(pp) * kTilePitch + lane] = re[pp];
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 32; ++n2) re[n2] = s_tile[lane * kTilePitch + n2];   // stage B: lane = k1, register = n2
        __syncwarp();
#pragma unroll
        for (int pp = 0; pp < 32; ++pp) s_tile[bitrev5(pp) * kTilePitch + lane] = im[pp];
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 32; ++n2) im[n2] = s_tile[lane * kTilePitch + n2];
        fft32x2(re, im);
        __syncwarp();
        // powers of both frames, row r = (k + 512) mod 1024, k = k1 + 32 k2
        float2* s_pow = s_tile;                                 // [1024] (frame 0, frame 1)
#pragma unroll
        for (int pp = 0; pp < 32; ++pp) {
            const int k = lane + 32 * bitrev5(pp);
            s_pow[(k + NFFT / 2) & (NFFT - 1)] = ffma2(re[pp], re[pp], fmul2(im[pp], im[pp]));
        }
        __syncwarp();
        // normalised dB of the bins the vertical taps touch, vertical taps of both frames, horizontal tap
        float* s_col = s_out + c * p.out_pitch;
        for (int ys = lane; ys < p.new_h; ys += 32) {
            const int2 rt = s_rowtab[ys];
            const float wy = __int_as_float(rt.y);
            const float2 p0 = s_pow[rt.x & 0xffff], p1 = s_pow[rt.x >> 16];
            const float a0 = __saturatef(fmaf(__log2f(fmaxf(p0.x, 1e-20f)), lg_scale, p.db_off));
            const float a1 = __saturatef(fmaf(__log2f(fmaxf(p1.x, 1e-20f)), lg_scale, p.db_off));
            const float b0 = __saturatef(fmaf(__log2f(fmaxf(p0.y, 1e-20f)), lg_scale, p.db_off));
            const float b1 = __saturatef(fmaf(__log2f(fmaxf(p1.y, 1e-20f)), lg_scale, p.db_off));
            const float u0 = a0 + wy * (a1 - a0), u1 = b0 + wy * (b1 - b0);
            s_col[ys] = u0 + wx * (u1 - u0);
        }
        __syncwarp();
    }
    __syncthreads();

    // ---- store (column fastest so a half-warp writes one 32/64-byte run), 3 identical channels ----
    const size_t obase = (size_t)b * 3 * plane + (size_t)p.top * a.out_w + (p.left + x0);
    if (!a.out_fp32 && ((p.left + x0) & 1) == 0 && (a.out_w & 1) == 0) {
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(a.out);
        for (int i = tid; i < p.new_h * (kCols / 2); i += kThreads) {
            const int c = (i % (kCols / 2)) * 2, ys = i / (kCols / 2);
            if (c >= ncols) continue;
            const size_t o = obase + (size_t)ys * a.out_w + c;
            const float v0 = s_out[c * p.out_pitch + ys];
            if (c + 1 < ncols) {
                const float v1 = s_out[(c + 1) * p.out_pitch + ys];
                const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
                *reinterpret_cast<__nv_bfloat162*>(out + o) = h;
                *reinterpret_cast<__nv_bfloat162*>(out + o + plane) = h;
                *reinterpret_cast<__nv_bfloat162*>(out + o + 2 * plane) = h;
            } else {
                const __nv_bfloat16 h = __float2bfloat16_rn(v0);
                out[o] = h;
                out[o + plane] = h;
                out[o + 2 * plane] = h;
            }
        }
        return;
    }
    for (int i = tid; i < p.new_h * kCols; i += kThreads) {
        const int c = i % kCols, ys = i / kCols;
        if (c >= ncols) continue;
        const float v = s_out[c * p.out_pitch + ys];
        const size_t o = obase + (size_t)ys * a.out_w + c;
        store_px(p, o, v);
        store_px(p, o + plane, v);
        store_px(p, o + 2 * plane, v);
    }
}

static float2* g_twiddle = nullptr;

static int ensure_twiddle() {
    if (g_twiddle) return SPECYOLO_OK;
    float2 h[NFFT];
    for (int j = 0; j < NFFT; ++j) {
        const double ang = -2.0 * 3.14159265358979323846 * (double)j / (double)NFFT;
        h[j] = make_float2((float)cos(ang), (float)sin(ang));
    }
    float2* d = nullptr;
    SY_CUDA(cudaMalloc(&d, sizeof(h)));
    SY_CUDA(cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice));
    g_twiddle = d;
    return SPECYOLO_OK;
}

// Python-style round half to even, as LetterBox uses int(round(x))
static int py_round(double v) { return (int)nearbyint(v); }

int stft_launch(const specyolo_stft_t* a, cudaStream_t stream) {
    SY_CHECK(a->nfft == NFFT, SPECYOLO_ERR_UNSUPPORTED, "only nfft == 1024 is supported (got %d)", a->nfft);
    SY_CHECK(a->hop >= 1 && a->L >= a->nfft, SPECYOLO_ERR_INVALID, "burst shorter than one frame");
    SY_CHECK(a->db_max > a->db_min, SPECYOLO_ERR_INVALID, "db_max must exceed db_min");
    SY_CHECK((reinterpret_cast<uintptr_t>(a->iq) & 7) == 0, SPECYOLO_ERR_INVALID, "iq must be 8-byte aligned");
    // the one-time twiddle upload allocates; call once outside any graph capture (the Python layer
    // does so at import)
    int rc = ensure_twiddle();
    if (rc != SPECYOLO_OK) return rc;

    StftParams p{};
    p.a = *a;
    p.T = 1 + (a->L - a->nfft) / a->hop;
    // LetterBox geometry, shape = (rows = nfft, cols = T)
    const double r = fmin((double)a->out_h / NFFT, (double)a->out_w / p.T);
    p.new_w = py_round(p.T * r);
    p.new_h = py_round(NFFT * r);
    if (p.new_w < 1) p.new_w = 1;
    if (p.new_h < 1) p.new_h = 1;
    if (p.new_w > a->out_w) p.new_w = a->out_w;
    if (p.new_h > a->out_h) p.new_h = a->out_h;
    const double dw = (a->out_w - p.new_w) / 2.0, dh = (a->out_h - p.new_h) / 2.0;
    p.top = py_round(dh - 0.1);
    p.left = py_round(dw - 0.1);
    p.sx = (double)p.T / p.new_w;
    p.sy = (double)NFFT / p.new_h;
    const double pref_db = 20.0 * log10(NFFT / 2.0);  // (sum w)^2 for the periodic Hann
    p.db_scale = 1.0f / (a->db_max - a->db_min);
    p.db_off = (float)((-pref_db - a->db_min) / (a->db_max - a->db_min));
    p.col_tiles = ceil_div(p.new_w, kCols);
    p.pad_tiles = 16;
    p.out_pitch = p.new_h + ((2 - p.new_h) & 31);
    p.twiddle = g_twiddle;

    const size_t smem = (size_t)kWarps * 32 * kTilePitch * 8 + (size_t)32 * 32 * 16 + (size_t)NFFT * 4 +
                        (size_t)p.new_h * 8 + (size_t)kCols * 8 + (size_t)kCols * p.out_pitch * 4;
    SY_CHECK(smem <= 220 * 1024, SPECYOLO_ERR_UNSUPPORTED, "content band too tall for shared memory");
    static size_t attr_smem[kMaxDevices] = {0};
    SY_CUDA(ensure_dynamic_smem(stft_letterbox_kernel, smem, attr_smem));
    dim3 grid((unsigned)(p.col_tiles + p.pad_tiles), (unsigned)a->B);
    stft_letterbox_kernel<<<grid, kThreads, smem, stream>>>(p);
    SY_LAUNCH_CHECK();
    count_launch();
    return SPECYOLO_OK;
}

int stft_init() { return ensure_twiddle(); }

}  // namespace specyolo
