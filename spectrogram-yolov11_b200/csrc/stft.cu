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
// CTA = 256 threads = 8 warps, 16 output columns = up to 32 frames, 4 per warp.  A warp computes one
// 1024-point FFT as 32 x 32 (four-step): 32-point FFT in registers, twiddle, transpose through a
// padded shared-memory tile, second 32-point FFT in registers.
#include "common.h"

namespace specyolo {

static constexpr int NFFT = 1024;
static constexpr int kCols = 16;      // output columns per CTA
static constexpr int kFrames = 2 * kCols;

__host__ __device__ constexpr int bitrev5(int v) {
    return ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4);
}

// In-register radix-2 DIF FFT of 32 complex points; X[k] ends up in x[bitrev5(k)].
__device__ __forceinline__ void fft32(float2 (&x)[32]) {
    // W32^j = exp(-2 pi i j / 32), j = 0..15: literal constants for the unrolled butterflies (W^0 = 1 and W^8 = -i are
    // special-cased below: 46 of the 80 butterflies of a 32-point FFT need no multiplication)
    constexpr float kW32c[16] = {
    1.000000000e+00f, 9.807852804e-01f, 9.238795325e-01f, 8.314696123e-01f, 7.071067812e-01f, 5.555702330e-01f,
    3.826834324e-01f, 1.950903220e-01f, 0.0f, -1.950903220e-01f, -3.826834324e-01f, -5.555702330e-01f,
    -7.071067812e-01f, -8.314696123e-01f, -9.238795325e-01f, -9.807852804e-01f};
    constexpr float kW32s[16] = {
    -0.0f, -1.950903220e-01f, -3.826834324e-01f, -5.555702330e-01f, -7.071067812e-01f, -8.314696123e-01f,
    -9.238795325e-01f, -9.807852804e-01f, -1.000000000e+00f, -9.807852804e-01f, -9.238795325e-01f,
    -8.314696123e-01f, -7.071067812e-01f, -5.555702330e-01f, -3.826834324e-01f, -1.950903220e-01f};

#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int half = 16 >> s;
#pragma unroll
        for (int g = 0; g < (1 << s); ++g) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const int i0 = g * 2 * half + j, i1 = i0 + half;
                const float2 a = x[i0], b = x[i1];
                x[i0] = make_float2(a.x + b.x, a.y + b.y);
                const float dr = a.x - b.x, di = a.y - b.y;
                const int tw = j << s;  // W_{2*half}^j = W32^(j * 32/(2*half)) = W32^(j << s)
                if (tw == 0) {
                    x[i1] = make_float2(dr, di);                 // W^0 = 1
                } else if (tw == 8) {
                    x[i1] = make_float2(di, -dr);                // W^8 = -i
                } else {
                    const float c = kW32c[tw], sn = kW32s[tw];
                    x[i1] = make_float2(dr * c - di * sn, dr * sn + di * c);
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
    int pad_tiles;         // CTAs (per burst) filling padding
    float db_scale, db_off;  // v = dB*db_scale + db_off
    const float2* twiddle; // [1024] exp(-2 pi i j / 1024)
};

__device__ __forceinline__ void store_px(const StftParams& p, size_t idx, float v) {
    if (p.a.out_fp32) reinterpret_cast<float*>(p.a.out)[idx] = v;
    else reinterpret_cast<__nv_bfloat16*>(p.a.out)[idx] = __float2bfloat16_rn(v);
}

__global__ void __launch_bounds__(256)
stft_letterbox_kernel(const __grid_constant__ StftParams p) {
    extern __shared__ __align__(16) uint8_t st_smem[];
    const specyolo_stft_t& a = p.a;
    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t plane = (size_t)a.out_h * a.out_w;

    if ((int)blockIdx.x >= p.col_tiles) {
        // ---------------- padding fill ----------------
        const int pt = blockIdx.x - p.col_tiles;
        const long total = (long)3 * plane;
        const long per = (total + p.pad_tiles - 1) / p.pad_tiles;
        const long lo = pt * per, hi = min(total, lo + per);
        for (long i = lo + tid; i < hi; i += 256) {
            const int x = (int)(i % a.out_w);
            const int y = (int)((i / a.out_w) % a.out_h);
            const bool inside = (x >= p.left) && (x < p.left + p.new_w) && (y >= p.top) && (y < p.top + p.new_h);
            if (!inside) store_px(p, (size_t)b * 3 * plane + i, a.pad_value);
        }
        return;
    }

    float2* s_tw = reinterpret_cast<float2*>(st_smem);                       // [1024]
    float2* s_tile = s_tw + NFFT + warp * (32 * 33);                          // per-warp [32][33]
    float* s_u = reinterpret_cast<float*>(s_tw + NFFT + 8 * (32 * 33));       // [kFrames][new_h]
    for (int i = tid; i < NFFT; i += 256) s_tw[i] = p.twiddle[i];
    __syncthreads();

    const int x0 = blockIdx.x * kCols;                 // first content column of this CTA
    const int ncols = min(kCols, p.new_w - x0);
    const float* iq = a.iq + (size_t)b * a.L * 2;

    for (int f = warp; f < 2 * ncols; f += 8) {
        // frame index of tap (f&1) of column x0 + f/2
        const int xs = x0 + (f >> 1);
        double fx = ((double)xs + 0.5) * p.sx - 0.5;
        if (fx < 0.0) fx = 0.0;
        int t0 = (int)floor(fx);
        if (t0 > p.T - 1) t0 = p.T - 1;
        const int t = (f & 1) ? min(t0 + 1, p.T - 1) : t0;

        // ---- stage A: lane = n2, register = n1; x[n1] = w[n] s[n], n = 32 n1 + n2 ----
        float2 x[32];
        const float2* src = reinterpret_cast<const float2*>(iq) + (size_t)t * a.hop;
#pragma unroll
        for (int n1 = 0; n1 < 32; ++n1) {
            const int n = 32 * n1 + lane;
            const float2 s = __ldg(src + n);
            const float w = 0.5f - 0.5f * s_tw[n].x;   // periodic Hann from cos(2 pi n / N)
            x[n1] = make_float2(s.x * w, s.y * w);
        }
        fft32(x);
        __syncwarp();
        // twiddle W1024^(n2*k1), write T[k1][n2]
#pragma unroll
        for (int pp = 0; pp < 32; ++pp) {
            const int k1 = bitrev5(pp);
            const float2 w = s_tw[(lane * k1) & (NFFT - 1)];
            const float2 v = x[pp];
            s_tile[k1 * 33 + lane] = make_float2(v.x * w.x - v.y * w.y, v.x * w.y + v.y * w.x);
        }
        __syncwarp();
        // ---- stage B: lane = k1, register = n2 ----
#pragma unroll
        for (int n2 = 0; n2 < 32; ++n2) x[n2] = s_tile[lane * 33 + n2];
        fft32(x);
        __syncwarp();
        // power -> normalised dB, row r = (k + 512) mod 1024, k = k1 + 32 k2
        float* s_row = reinterpret_cast<float*>(s_tile);
#pragma unroll
        for (int pp = 0; pp < 32; ++pp) {
            const int k2 = bitrev5(pp);
            const int k = lane + 32 * k2;
            const float pw = fmaxf(x[pp].x * x[pp].x + x[pp].y * x[pp].y, 1e-20f);
            // 10*log10(pw) = 3.0102999566 * log2(pw)
            float v = fmaf(__log2f(pw) * 3.010299956639812f, p.db_scale, p.db_off);
            v = fminf(fmaxf(v, 0.f), 1.f);
            s_row[(k + NFFT / 2) & (NFFT - 1)] = v;
        }
        __syncwarp();
        // vertical bilinear taps for every output row of the content band
        for (int ys = lane; ys < p.new_h; ys += 32) {
            double fy = ((double)ys + 0.5) * p.sy - 0.5;
            if (fy < 0.0) fy = 0.0;
            int r0 = (int)floor(fy);
            if (r0 > NFFT - 1) r0 = NFFT - 1;
            const int r1 = min(r0 + 1, NFFT - 1);
            const float wy = (float)(fy - (double)r0);
            s_u[f * p.new_h + ys] = s_row[r0] + wy * (s_row[r1] - s_row[r0]);
        }
        __syncwarp();
    }
    __syncthreads();

    // ---- horizontal taps + store (col fastest so a half-warp writes one 32/64-byte run) ----
    for (int i = tid; i < p.new_h * kCols; i += 256) {
        const int c = i % kCols, ys = i / kCols;
        if (c >= ncols) continue;
        const int xs = x0 + c;
        double fx = ((double)xs + 0.5) * p.sx - 0.5;
        if (fx < 0.0) fx = 0.0;
        int t0 = (int)floor(fx);
        if (t0 > p.T - 1) t0 = p.T - 1;
        const float wx = (float)(fx - (double)t0);
        const float u0 = s_u[(2 * c) * p.new_h + ys], u1 = s_u[(2 * c + 1) * p.new_h + ys];
        const float v = u0 + wx * (u1 - u0);
        const size_t o = (size_t)b * 3 * plane + (size_t)(p.top + ys) * a.out_w + (p.left + xs);
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
    p.pad_tiles = 8;
    p.twiddle = g_twiddle;

    const size_t smem = (size_t)NFFT * 8 + (size_t)8 * 32 * 33 * 8 + (size_t)kFrames * p.new_h * 4;
    SY_CHECK(smem <= 220 * 1024, SPECYOLO_ERR_UNSUPPORTED, "content band too tall for shared memory");
    static size_t attr_smem = 0;
    if (smem > attr_smem) {
        SY_CUDA(cudaFuncSetAttribute(stft_letterbox_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    dim3 grid((unsigned)(p.col_tiles + p.pad_tiles), (unsigned)a->B);
    stft_letterbox_kernel<<<grid, 256, smem, stream>>>(p);
    SY_LAUNCH_CHECK();
    count_launch();
    return SPECYOLO_OK;
}

int stft_init() { return ensure_twiddle(); }

}  // namespace specyolo
