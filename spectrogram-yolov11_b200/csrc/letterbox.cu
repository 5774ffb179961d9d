// Image ingest: LetterBox + layout step of the predictor for uint8 images, bit-exact with the reference.
// Replaces, per image, LetterBox.__call__ (ultralytics/data/augment.py:1535-1601: cv2.resize INTER_LINEAR +
// cv2.copyMakeBorder(114)) and the BGR->RGB / HWC->CHW shuffle of BasePredictor.preprocess
// (ultralytics/engine/predictor.py:125-136).  The resize reproduces OpenCV's 8-bit INTER_LINEAR arithmetic
// (third-party: opencv-python >= 4.6, modules/imgproc/src/resize.cpp, checked here against cv2 4.13):
//   x:  fx = (float)((dx + 0.5) * scale_x - 0.5), sx = floor(fx), fx -= sx; sx < 0 -> (0, 0); sx >= W-1 -> (W-1, 0);
//       alpha = round((1 - fx, fx) * 2048) (int16);   row value  D = S[sx] * alpha0 + S[sx+1] * alpha1   (int32)
//   y:  same fy / floor but NO clamping of the weights: the two row indices are clamped to [0, H-1] instead
//   out = ( ((beta0 * (D0 >> 4)) >> 16) + ((beta1 * (D1 >> 4)) >> 16) + 2 ) >> 2
//   exact 2x shrink in both directions is OpenCV's area fast path: (a + b + c + d + 2) >> 2.
// scale = 1.0 / ((double)dst / src), all in IEEE double / float as on the host (no fast-math), so the coefficients are
// computed per thread and match the host bit for bit.
#include "common.h"
#include "ptx.cuh"
#include "tma_host.h"

namespace specyolo {

struct LetterboxParams {
    const uint8_t* src;     // [B, H, W, 3]
    uint8_t* dst;           // chw: [B, 3, out_h, out_w]; else [B, out_h, out_w, 3]
    int B, H, W, out_h, out_w;
    int new_w, new_h, left, top;
    int swap_rb, chw, pad_value;
    double scale_x, scale_y;
};

__device__ __forceinline__ void lb_coef_x(int d, double scale, int n, int& s, int& a0, int& a1) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    s = (int)floorf(f);
    f -= (float)s;
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= n - 1) { f = 0.f; s = n - 1; }
    a0 = __float2int_rn((1.0f - f) * 2048.0f);
    a1 = __float2int_rn(f * 2048.0f);
}

__global__ void __launch_bounds__(256)
letterbox_u8_kernel(const __grid_constant__ LetterboxParams p) {
    ptx::grid_dep_launch();
    ptx::grid_dep_wait();
    const size_t plane = (size_t)p.out_h * p.out_w;
    const size_t total = (size_t)p.B * plane;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(idx % p.out_w);
        const int y = (int)((idx / p.out_w) % p.out_h);
        const int b = (int)(idx / plane);
        int v[3] = {p.pad_value, p.pad_value, p.pad_value};
        const int dx = x - p.left, dy = y - p.top;
        if (dx >= 0 && dx < p.new_w && dy >= 0 && dy < p.new_h) {
            const uint8_t* img = p.src + (size_t)b * p.H * p.W * 3;
            if (p.new_w == p.W && p.new_h == p.H) {
                const uint8_t* s = img + ((size_t)dy * p.W + dx) * 3;
                v[0] = s[0]; v[1] = s[1]; v[2] = s[2];
            } else if (p.W == 2 * p.new_w && p.H == 2 * p.new_h) {
                const uint8_t* s0 = img + ((size_t)(2 * dy) * p.W + 2 * dx) * 3;
                const uint8_t* s1 = s0 + (size_t)p.W * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = (s0[c] + s0[3 + c] + s1[c] + s1[3 + c] + 2) >> 2;
            } else {
                int sx, a0, a1;
                lb_coef_x(dx, p.scale_x, p.W, sx, a0, a1);
                float fy = (float)(((double)dy + 0.5) * p.scale_y - 0.5);
                const int sy = (int)floorf(fy);
                fy -= (float)sy;
                const int b0 = __float2int_rn((1.0f - fy) * 2048.0f), b1 = __float2int_rn(fy * 2048.0f);
                const int r0 = min(max(sy, 0), p.H - 1), r1 = min(max(sy + 1, 0), p.H - 1);
                const int x1 = min(sx + 1, p.W - 1);
                const uint8_t* q00 = img + ((size_t)r0 * p.W + sx) * 3;
                const uint8_t* q01 = img + ((size_t)r0 * p.W + x1) * 3;
                const uint8_t* q10 = img + ((size_t)r1 * p.W + sx) * 3;
                const uint8_t* q11 = img + ((size_t)r1 * p.W + x1) * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int d0 = q00[c] * a0 + q01[c] * a1;
                    const int d1 = q10[c] * a0 + q11[c] * a1;
                    v[c] = (((b0 * (d0 >> 4)) >> 16) + ((b1 * (d1 >> 4)) >> 16) + 2) >> 2;
                    v[c] = min(max(v[c], 0), 255);
                }
            }
        }
        if (p.swap_rb) { const int t = v[0]; v[0] = v[2]; v[2] = t; }
        if (p.chw) {
            uint8_t* o = p.dst + (size_t)b * 3 * plane + (size_t)y * p.out_w + x;
            o[0] = (uint8_t)v[0]; o[plane] = (uint8_t)v[1]; o[2 * plane] = (uint8_t)v[2];
        } else {
            uint8_t* o = p.dst + idx * 3;
            o[0] = (uint8_t)v[0]; o[1] = (uint8_t)v[1]; o[2] = (uint8_t)v[2];
        }
    }
}

int letterbox_u8_launch(const uint8_t* src, int B, int H, int W, uint8_t* dst, int out_h, int out_w, int new_w, int new_h,
                        int left, int top, int pad_value, int swap_rb, int chw, cudaStream_t stream) {
    SY_CHECK(new_w >= 1 && new_h >= 1 && left >= 0 && top >= 0 && left + new_w <= out_w && top + new_h <= out_h,
             SPECYOLO_ERR_INVALID, "letterbox: content %dx%d at (%d,%d) does not fit %dx%d", new_w, new_h, left, top, out_w, out_h);
    LetterboxParams p{};
    p.src = src; p.dst = dst; p.B = B; p.H = H; p.W = W; p.out_h = out_h; p.out_w = out_w;
    p.new_w = new_w; p.new_h = new_h; p.left = left; p.top = top;
    p.swap_rb = swap_rb; p.chw = chw; p.pad_value = pad_value;
    p.scale_x = 1.0 / ((double)new_w / (double)W);      // cv::resize: scale = 1 / inv_scale, inv_scale = dsize / ssize
    p.scale_y = 1.0 / ((double)new_h / (double)H);
    const size_t total = (size_t)B * out_h * out_w;
    const size_t want = (total + 255) / 256, cap = (size_t)sm_count() * 16;
    SY_CUDA(launch_pdl(letterbox_u8_kernel, dim3((unsigned)(want < cap ? want : cap)), dim3(256), 0, stream, p));
    SY_LAUNCH_CHECK();
    count_launch();
    return SPECYOLO_OK;
}

}  // namespace specyolo
