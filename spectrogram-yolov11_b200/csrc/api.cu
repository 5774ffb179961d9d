// extern "C" surface of libspecyolo (see include/specyolo.h): argument validation, dispatch to the
// kernel launchers, error text.  No logic beyond that lives here.
#include "common.h"

#include <atomic>
#include <cmath>
#include <cstring>
#include <cstdlib>

namespace specyolo {

static thread_local char g_err[1024] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// launchers implemented in the other translation units
int conv_igemm_launch(const specyolo_conv_t* a, cudaStream_t stream);
int conv_halo_try_launch(const specyolo_conv_t* a, cudaStream_t stream);
bool conv_halo_geometry_ok(int kh, int kw, int stride, int pad, int dil);
int dwconv3x3_launch(const specyolo_conv_t* a, cudaStream_t stream);
int fold_pack_launch(const float*, const float*, const float*, const float*, const float*, const float*, float,
                     int, int, int, int, int, int, int, void*, float*, cudaStream_t);
int stem_conv_launch(const void*, int, int, int, int, const float*, const float*, int, void*, int, cudaStream_t);
int stem_s2d_launch(const void*, int, int, int, int, void*, int, cudaStream_t);
int nchw_to_nhwc_launch(const void*, int, float, int, int, int, int, void*, int, cudaStream_t);
int nhwc_to_nchw_launch(const void*, int, int, int, int, int, float*, cudaStream_t);
int sppf_pool_launch(void*, int, int, int, int, int, cudaStream_t);
size_t fusion_ws_bytes(int, int, int, int, int);
int fusion_launch(const specyolo_fusion_t*, cudaStream_t);
int spatial_gate_launch(const specyolo_spatial_gate_t*, cudaStream_t);
int msc_gate_launch(const specyolo_msc_gate_t*, cudaStream_t);
size_t msc_ws_bytes(int, int, int, int);
int bottlenect_launch(const specyolo_bottlenect_t*, cudaStream_t);
size_t bottlenect_ws_bytes(int, int, int, int);
int det_loss_launch(const specyolo_det_loss_t*, cudaStream_t);
int ema_update_launch(float* const*, const float* const*, const long long*, const int*, const long long*, int, int, float, float, cudaStream_t);
size_t det_loss_ws_bytes(int B, const int* h, const int* w, int nl, int M, int topk);
int psa_attention_launch(const void*, int, int, int, int, int, int, int, float, const float*, const float*,
                         void*, int, cudaStream_t);
int detect_decode_launch(const specyolo_decode_t*, cudaStream_t);
size_t nms_ws_bytes(int, int, int, int);
int nms_launch(const specyolo_nms_t*, cudaStream_t);
int scale_boxes_launch(float*, const int*, int, int, float, float, float, float, float, cudaStream_t);
int stft_launch(const specyolo_stft_t*, cudaStream_t);
int dwpw_launch(const specyolo_dwpw_t*, cudaStream_t);
bool stem_pair_ok(int, int, int, int, int);
int jpeg_info(const void*, size_t, int*, int*, int*);
int jpeg_decode_bgr(const void*, size_t, void*, int, int, cudaStream_t);
int jpeg_decode_batch_bgr(const void* const*, const size_t*, void* const*, const int*, int, int, cudaStream_t);
int stem_pair_launch(const specyolo_stem_pair_t*, cudaStream_t);
int letterbox_u8_launch(const uint8_t*, int, int, int, uint8_t*, int, int, int, int, int, int, int, int, int, cudaStream_t);
int match_predictions_launch(const float*, const int*, int, int, const float*, const int*, int, const float*, int, uint8_t*,
                             cudaStream_t);
int stft_init();
int upsample2x_launch(const void*, int, int, int, int, int, void*, int, cudaStream_t);
bool bneck_pair_ok(int, int, int, int, int);
int bneck_pair_launch(const specyolo_bneck_t*, cudaStream_t);

}  // namespace specyolo

using namespace specyolo;

extern "C" {

const char* specyolo_last_error(void) { return g_err; }
int specyolo_version(void) { return 100; }
uint64_t specyolo_launch_count(void) { return g_launches.load(); }
void specyolo_reset_launch_count(void) { g_launches.store(0); }

int specyolo_init(void) {
    int dev = 0;
    SY_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    SY_CUDA(cudaGetDeviceProperties(&prop, dev));
    SY_CHECK(prop.major == 10, SPECYOLO_ERR_UNSUPPORTED,
             "libspecyolo is built for sm_100a only; device %d is sm_%d%d", dev, prop.major, prop.minor);
    return stft_init();
}

int specyolo_nchw_to_nhwc_bf16(const void* x, int x_dtype, float scale, int B, int C, int H, int W, void* y,
                               int y_pixstride, void* stream) {
    SY_CHECK(x && y && B > 0 && C > 0 && H > 0 && W > 0 && y_pixstride >= C, SPECYOLO_ERR_INVALID,
             "nchw_to_nhwc: bad arguments");
    return nchw_to_nhwc_launch(x, x_dtype, scale, B, C, H, W, y, y_pixstride, (cudaStream_t)stream);
}

int specyolo_nhwc_bf16_to_nchw_f32(const void* x, int x_pixstride, int B, int C, int H, int W, float* y,
                                   void* stream) {
    SY_CHECK(x && y && B > 0 && C > 0 && H > 0 && W > 0 && x_pixstride >= C, SPECYOLO_ERR_INVALID,
             "nhwc_to_nchw: bad arguments");
    return nhwc_to_nchw_launch(x, x_pixstride, B, C, H, W, y, (cudaStream_t)stream);
}

int specyolo_upsample2x(const void* x, int x_pixstride, int B, int H, int W, int C, void* y, int y_pixstride, void* stream) {
    SY_CHECK(x && y && B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && x_pixstride >= C && y_pixstride >= C &&
                 x_pixstride % 8 == 0 && y_pixstride % 8 == 0 &&
                 !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15),
             SPECYOLO_ERR_INVALID, "upsample2x: C and the pixel strides must be multiples of 8, tensors 16-byte aligned");
    return upsample2x_launch(x, x_pixstride, B, H, W, C, y, y_pixstride, (cudaStream_t)stream);
}

int specyolo_conv_merge(int cin, int cout, int groups, int k, int stride, int pad, int dil) {
    if (groups <= 1 || cin <= 0 || cout <= 0 || cin % groups || cout % groups) return 1;
    const int cin_g = cin / groups;
    if (cin_g == 1 && cout / groups == 1) {
        // depthwise k x k: as a 64-channel block-DIAGONAL implicit GEMM on the halo-tile kernel the layer runs at
        // ~HBM speed on the otherwise idle tensor pipe (63/64 of the multiplications are by zero, and still 2-3x
        // faster than the CUDA-core kernel, which is issue-bound at 25 instructions per output element)
        if (conv_halo_geometry_ok(k, k, stride, pad, dil) && !getenv("SPECYOLO_DW_CUDA_CORES"))
            for (int m = 64; m >= 16; m >>= 1)
                if (groups % m == 0) return m;
        return 1;                                       // dedicated CUDA-core kernel
    }
    // the halo kernel multiplies every group by its own weight box (UMMA N = cout_g): nothing to merge
    if (cin_g % 16 == 0 && conv_halo_geometry_ok(k, k, stride, pad, dil)) return 1;
    int merge = 1;
    while (cin_g * merge * 2 <= 64 && groups % (merge * 2) == 0) merge *= 2;
    return merge;
}

int specyolo_conv_npad(int cout, int groups) {
    if (groups <= 0 || cout <= 0 || cout % groups) return -1;
    const int cg = cout / groups;
    if (cg == 1) return 1;            // depthwise kernel: one row per group
    return (cg + 15) / 16 * 16;       // UMMA N granularity for M = 128
}

int specyolo_fold_pack_conv(const float* w_oihw, const float* conv_bias, const float* bn_gamma,
                            const float* bn_beta, const float* bn_mean, const float* bn_var, float bn_eps,
                            int cout, int cin_g, int kh, int kw, int groups, int merge, int n_pad,
                            void* w_packed, float* bias_out, void* stream) {
    SY_CHECK(w_oihw && w_packed && bias_out, SPECYOLO_ERR_INVALID, "fold_pack: null pointer");
    SY_CHECK((bn_gamma == nullptr) == (bn_beta == nullptr) && (bn_gamma == nullptr) == (bn_mean == nullptr) &&
                 (bn_gamma == nullptr) == (bn_var == nullptr),
             SPECYOLO_ERR_INVALID, "fold_pack: give all BN tensors or none");
    return fold_pack_launch(w_oihw, conv_bias, bn_gamma, bn_beta, bn_mean, bn_var, bn_eps, cout, cin_g, kh, kw,
                            groups, merge, n_pad, w_packed, bias_out, (cudaStream_t)stream);
}

int specyolo_conv2d_bias_act(const specyolo_conv_t* a, void* stream) {
    SY_CHECK(a && a->x && a->w_packed && a->bias && a->y, SPECYOLO_ERR_INVALID, "conv: null pointer");
    SY_CHECK(a->B > 0 && a->H > 0 && a->W > 0 && a->Cin > 0 && a->Cout > 0 && a->groups > 0,
             SPECYOLO_ERR_INVALID, "conv: bad sizes");
    SY_CHECK(a->kh >= 1 && a->kw >= 1 && a->stride >= 1 && a->dil >= 1 && a->pad >= 0, SPECYOLO_ERR_INVALID,
             "conv: bad kernel geometry");
    const int ho = (a->H + 2 * a->pad - a->dil * (a->kh - 1) - 1) / a->stride + 1;
    const int wo = (a->W + 2 * a->pad - a->dil * (a->kw - 1) - 1) / a->stride + 1;
    // Ho/Wo may be SMALLER than the full output: the bottom rows / right columns are then not computed (a conv with
    // less padding at the far edge, e.g. the 2x2 space-to-depth form of the stem)
    SY_CHECK(a->Ho >= 1 && a->Wo >= 1 && a->Ho <= ho && a->Wo <= wo, SPECYOLO_ERR_INVALID,
             "conv: Ho/Wo (%d,%d) exceed the geometry (%d,%d)", a->Ho, a->Wo, ho, wo);
    SY_CHECK(a->x_pixstride >= a->Cin && a->y_pixstride >= (a->y_s2d ? 4 : 1) * a->Cout, SPECYOLO_ERR_INVALID,
             "conv: pixel stride too small");
    SY_CHECK(!a->y_s2d || (a->Ho % 2 == 0 && a->Wo % 2 == 0 && !a->y_fp32 && a->groups == 1), SPECYOLO_ERR_INVALID,
             "conv: blocked (space-to-depth) output needs even Ho/Wo, bf16, groups == 1");
    SY_CHECK(a->act == SPECYOLO_ACT_NONE || a->act == SPECYOLO_ACT_SILU, SPECYOLO_ERR_INVALID, "conv: bad act");
    SY_CHECK(a->x_upshift == 0, SPECYOLO_ERR_UNSUPPORTED, "conv: x_upshift is not implemented");
    if (a->groups == a->Cin && a->Cin == a->Cout && a->groups > 1) {
        SY_CHECK(!a->y_s2d, SPECYOLO_ERR_UNSUPPORTED, "conv: blocked output is not available for depthwise convs");
        return dwconv3x3_launch(a, (cudaStream_t)stream);
    }
    const int r = conv_halo_try_launch(a, (cudaStream_t)stream);   // k x k convs with resident weights
    if (r >= 0) return r;
    SY_CHECK(!a->y_s2d, SPECYOLO_ERR_UNSUPPORTED, "conv: blocked output is implemented by the halo-tile kernel only");
    return conv_igemm_launch(a, (cudaStream_t)stream);
}

int specyolo_dwconv_pwconv(const specyolo_dwpw_t* a, void* stream) {
    SY_CHECK(a && a->x && a->dw_w && a->dw_b && a->pw_packed && a->pw_bias && (a->y || a->head_w), SPECYOLO_ERR_INVALID,
             "dwpw: null pointer");
    SY_CHECK(a->B > 0 && a->H > 0 && a->W > 0 && a->Cout > 0 && a->x_pixstride >= a->C && (a->head_w || a->y_pixstride >= a->Cout),
             SPECYOLO_ERR_INVALID, "dwpw: bad sizes");
    return dwpw_launch(a, (cudaStream_t)stream);
}

int specyolo_jpeg_info(const void* data, size_t nbytes, int* H, int* W, int* channels) {
    SY_CHECK(data && nbytes > 4 && H && W && channels, SPECYOLO_ERR_INVALID, "jpeg_info: bad arguments");
    return jpeg_info(data, nbytes, H, W, channels);
}

int specyolo_jpeg_decode_bgr(const void* data, size_t nbytes, void* out_dev, int H, int W, void* stream) {
    SY_CHECK(data && nbytes > 4 && out_dev && H > 0 && W > 0, SPECYOLO_ERR_INVALID, "jpeg_decode: bad arguments");
    return jpeg_decode_bgr(data, nbytes, out_dev, H, W, (cudaStream_t)stream);
}

int specyolo_jpeg_decode_batch_bgr(const void* const* data, const size_t* nbytes, void* const* out_dev, const int* W, int n,
                                   int backend, void* stream) {
    SY_CHECK(data && nbytes && out_dev && W && n > 0, SPECYOLO_ERR_INVALID, "jpeg_decode_batch: bad arguments");
    return jpeg_decode_batch_bgr(data, nbytes, out_dev, W, n, backend, (cudaStream_t)stream);
}

int specyolo_stem_pair_ok(int H, int W, int c0, int Cout, int n_pad) { return stem_pair_ok(H, W, c0, Cout, n_pad) ? 1 : 0; }

int specyolo_stem_pair(const specyolo_stem_pair_t* a, void* stream) {
    SY_CHECK(a && a->x && a->w0 && a->b0 && a->w1_packed && a->b1 && a->y, SPECYOLO_ERR_INVALID, "stem_pair: null pointer");
    SY_CHECK(a->B > 0 && a->H > 0 && a->W > 0 && a->Cout > 0 && a->y_pixstride >= a->Cout, SPECYOLO_ERR_INVALID,
             "stem_pair: bad sizes");
    return stem_pair_launch(a, (cudaStream_t)stream);
}

int specyolo_bottleneck_ok(int C, int Cmid, int Cout, int n_pad1, int n_pad2) {
    return bneck_pair_ok(C, Cmid, Cout, n_pad1, n_pad2) ? 1 : 0;
}

int specyolo_bottleneck(const specyolo_bneck_t* a, void* stream) {
    SY_CHECK(a && a->x && a->w1_packed && a->b1 && a->w2_packed && a->b2 && a->y, SPECYOLO_ERR_INVALID, "bottleneck: null pointer");
    SY_CHECK(a->B > 0 && a->H > 0 && a->W > 0 && a->C > 0 && a->x_pixstride >= a->C && a->y_pixstride >= a->Cout,
             SPECYOLO_ERR_INVALID, "bottleneck: bad sizes");
    return bneck_pair_launch(a, (cudaStream_t)stream);
}

int specyolo_stem_conv3x3s2(const void* x, int x_dtype, int B, int H, int W, const float* w, const float* bias,
                            int Cout, void* y, int y_pixstride, void* stream) {
    SY_CHECK(x && w && bias && y && B > 0 && H > 0 && W > 0, SPECYOLO_ERR_INVALID, "stem: bad arguments");
    return stem_conv_launch(x, x_dtype, B, H, W, w, bias, Cout, y, y_pixstride, (cudaStream_t)stream);
}

int specyolo_stem_space_to_depth(const void* x, int x_dtype, int B, int H, int W, void* y, int y_pixstride, void* stream) {
    SY_CHECK(x && y && B > 0 && H > 0 && W > 0, SPECYOLO_ERR_INVALID, "stem s2d: bad arguments");
    return stem_s2d_launch(x, x_dtype, B, H, W, y, y_pixstride, (cudaStream_t)stream);
}

int specyolo_sppf_pool(void* buf, int B, int H, int W, int c, int pixstride, void* stream) {
    SY_CHECK(buf && B > 0 && H > 0 && W > 0 && c > 0, SPECYOLO_ERR_INVALID, "sppf: bad arguments");
    return sppf_pool_launch(buf, B, H, W, c, pixstride, (cudaStream_t)stream);
}

size_t specyolo_fusion_ws_bytes(int k, int B, int H, int W, int c) { return fusion_ws_bytes(k, B, H, W, c); }
int specyolo_fusion_eschannel(const specyolo_fusion_t* a, void* stream) {
    SY_CHECK(a && a->y && a->alpha && a->gamma && a->beta && a->sab_w, SPECYOLO_ERR_INVALID, "fusion: null pointer");
    return fusion_launch(a, (cudaStream_t)stream);
}

size_t specyolo_det_loss_ws_bytes(int B, const int* h, const int* w, int nl, int M, int topk) {
    if (B < 1 || !h || !w || nl < 1 || nl > 4 || M < 0 || topk < 1) return 0;
    return det_loss_ws_bytes(B, h, w, nl, M, topk);
}

int specyolo_det_loss(const specyolo_det_loss_t* a, void* stream) {
    SY_CHECK(a && a->pred_distri && a->pred_scores && a->out && a->ws && a->gt_count, SPECYOLO_ERR_INVALID, "det loss: null pointer");
    SY_CHECK(a->M == 0 || (a->gt_boxes && a->gt_labels), SPECYOLO_ERR_INVALID, "det loss: ground truth missing");
    return det_loss_launch(a, (cudaStream_t)stream);
}

int specyolo_ema_update(float* const* ema, const float* const* model, const long long* numel, const int* chunk_tensor,
                        const long long* chunk_off, int nchunks, int chunk, float d, float one_minus_d, void* stream) {
    SY_CHECK(nchunks == 0 || (ema && model && numel && chunk_tensor && chunk_off), SPECYOLO_ERR_INVALID, "ema: null pointer");
    return ema_update_launch(ema, model, numel, chunk_tensor, chunk_off, nchunks, chunk, d, one_minus_d, (cudaStream_t)stream);
}

int specyolo_sobel_spatial_attention(const specyolo_spatial_gate_t* a, void* stream) {
    SY_CHECK(a && a->x && a->y && a->mm, SPECYOLO_ERR_INVALID, "spatial gate: null pointer");
    SY_CHECK(a->B > 0 && a->H > 0 && a->W > 0 && a->C > 0, SPECYOLO_ERR_INVALID, "spatial gate: bad sizes");
    return spatial_gate_launch(a, (cudaStream_t)stream);
}

size_t specyolo_bottlenect_ws_bytes(int B, int H, int W, int C) { return bottlenect_ws_bytes(B, H, W, C); }

int specyolo_bottlenect(const specyolo_bottlenect_t* a, void* stream) {
    SY_CHECK(a && a->x && a->y && a->ws && a->in_w && a->in_b && a->fac_w && a->fac_b && a->sca_w && a->sca_b && a->dw1_w &&
             a->dw1_b && a->dw2_w && a->dw2_b && a->alpha && a->beta, SPECYOLO_ERR_INVALID, "BottleNect: null pointer");
    SY_CHECK(a->B > 0 && a->H > 0 && a->W > 0 && a->C > 0, SPECYOLO_ERR_INVALID, "BottleNect: bad sizes");
    return bottlenect_launch(a, (cudaStream_t)stream);
}

size_t specyolo_msc_ws_bytes(int B, int H, int W, int C) { return msc_ws_bytes(B, H, W, C); }

int specyolo_msc_spatial_attention(const specyolo_msc_gate_t* a, void* stream) {
    SY_CHECK(a && a->x && a->y && a->ws && a->w_big && a->w_small && a->fc_w && a->fc_b, SPECYOLO_ERR_INVALID, "msc gate: null pointer");
    SY_CHECK(a->B > 0 && a->H > 0 && a->W > 0 && a->C > 0, SPECYOLO_ERR_INVALID, "msc gate: bad sizes");
    return msc_gate_launch(a, (cudaStream_t)stream);
}

int specyolo_psa_attention(const void* qkv, int qkv_pixstride, int B, int H, int W, int heads, int key_dim,
                           int head_dim, float scale, const float* pe_w, const float* pe_b, void* out,
                           int out_pixstride, void* stream) {
    SY_CHECK(qkv && pe_w && pe_b && out && B > 0 && H > 0 && W > 0 && heads > 0, SPECYOLO_ERR_INVALID,
             "psa attention: bad arguments");
    return psa_attention_launch(qkv, qkv_pixstride, B, H, W, heads, key_dim, head_dim, scale, pe_w, pe_b, out,
                                out_pixstride, (cudaStream_t)stream);
}

int specyolo_detect_decode(const specyolo_decode_t* a, void* stream) {
    SY_CHECK(a && a->B > 0 && a->nc > 0, SPECYOLO_ERR_INVALID, "decode: bad arguments");
    return detect_decode_launch(a, (cudaStream_t)stream);
}

size_t specyolo_nms_ws_bytes(int B, int nc, int A, int multi_label) { return nms_ws_bytes(B, nc, A, multi_label); }
int specyolo_nms(const specyolo_nms_t* a, void* stream) {
    SY_CHECK(a != nullptr, SPECYOLO_ERR_INVALID, "nms: null argument");
    return nms_launch(a, (cudaStream_t)stream);
}

int specyolo_scale_boxes(float* out, const int* out_count, int B, int max_det, float gain, float pad_w,
                         float pad_h, float img0_w, float img0_h, void* stream) {
    SY_CHECK(out && out_count && B > 0 && max_det > 0 && gain > 0.f, SPECYOLO_ERR_INVALID, "scale_boxes: bad arguments");
    return scale_boxes_launch(out, out_count, B, max_det, gain, pad_w, pad_h, img0_w, img0_h, (cudaStream_t)stream);
}

int specyolo_match_predictions(const float* pred, const int* pred_count, int B, int max_det, const float* labels,
                               const int* label_off, int max_labels_per_image, const float* iouv_host, int niou,
                               uint8_t* correct, void* stream) {
    SY_CHECK(pred && pred_count && label_off && iouv_host && correct && B > 0, SPECYOLO_ERR_INVALID,
             "match_predictions: bad arguments");
    SY_CHECK(labels != nullptr || max_labels_per_image == 0, SPECYOLO_ERR_INVALID, "match_predictions: labels missing");
    return match_predictions_launch(pred, pred_count, B, max_det, labels, label_off, max_labels_per_image, iouv_host, niou,
                                    correct, (cudaStream_t)stream);
}

int specyolo_letterbox_u8(const uint8_t* src_hwc, int B, int H, int W, uint8_t* dst, int out_h, int out_w, int new_w,
                          int new_h, int left, int top, int pad_value, int swap_rb, int chw, void* stream) {
    SY_CHECK(src_hwc && dst && B > 0 && H > 0 && W > 0 && out_h > 0 && out_w > 0, SPECYOLO_ERR_INVALID,
             "letterbox: bad arguments");
    return letterbox_u8_launch(src_hwc, B, H, W, dst, out_h, out_w, new_w, new_h, left, top, pad_value, swap_rb, chw,
                               (cudaStream_t)stream);
}

int specyolo_iq_to_letterbox(const specyolo_stft_t* a, void* stream) {
    SY_CHECK(a && a->iq && a->out && a->B > 0 && a->out_h > 0 && a->out_w > 0, SPECYOLO_ERR_INVALID,
             "iq_to_letterbox: bad arguments");
    return stft_launch(a, (cudaStream_t)stream);
}

}  // extern "C"
