/*
 * libspecyolo — C-ABI boundary of the B200-native (sm_100a) Spectrogram-YOLOv11 inference hot path.
 *
 * Plain C: raw device pointers, sizes, a CUDA stream handle passed as void*, int status codes.
 * No torch / C++ types cross this boundary.  Every entry point is asynchronous on `stream`
 * (nothing synchronises, nothing allocates device memory) so a whole forward can be captured
 * in a CUDA graph.  All activation tensors are NHWC ("channels last") bf16 unless stated;
 * a tensor argument is a pointer to its first element plus a *pixel stride* in elements, so a
 * channel window of a wider concat buffer is passed without a copy.
 *
 * Each function names the reference interface (file:line under the upstream repo
 * httpsLiem/Spectrogram-YOLOv11, a fork of Ultralytics 8.3.70) that it replaces.
 *
 * Errors: functions return SPECYOLO_OK or a SPECYOLO_ERR_* code; specyolo_last_error() returns a
 * thread-local human-readable message.  There is no CPU fallback anywhere behind this header.
 */
#ifndef SPECYOLO_H
#define SPECYOLO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPECYOLO_OK               0
#define SPECYOLO_ERR_INVALID      1   /* bad argument (shape, alignment, range)          */
#define SPECYOLO_ERR_CUDA         2   /* CUDA runtime / driver error                     */
#define SPECYOLO_ERR_UNSUPPORTED  3   /* valid request this build does not implement     */

#define SPECYOLO_ACT_NONE 0
#define SPECYOLO_ACT_SILU 1

/* ---- library ---------------------------------------------------------------------------- */
const char* specyolo_last_error(void);
int         specyolo_version(void);
/* one-time per-process setup on the current device (constant tables); refuses non-sm_100 devices.
 * Must be called once before the first specyolo_iq_to_letterbox and outside graph capture. */
int         specyolo_init(void);
/* number of kernels this library has launched in this process (bench.py: gpu_launches) */
uint64_t    specyolo_launch_count(void);
void        specyolo_reset_launch_count(void);

/* ---- layout plumbing -------------------------------------------------------------------- */
/* NCHW (fp32 | bf16 | uint8 scaled by 1/255) -> NHWC bf16.
 * Replaces the dtype/scale step of BasePredictor.preprocess (ultralytics/engine/predictor.py:125-136). */
#define SPECYOLO_DT_F32  0
#define SPECYOLO_DT_BF16 1
#define SPECYOLO_DT_U8   2
int specyolo_nchw_to_nhwc_bf16(const void* x, int x_dtype, float scale,
                               int B, int C, int H, int W,
                               void* y, int y_pixstride, void* stream);
/* NHWC bf16 window -> NCHW fp32 (debug / drop-in return values such as Detect's raw maps). */
int specyolo_nhwc_bf16_to_nchw_f32(const void* x, int x_pixstride, int B, int C, int H, int W,
                                   float* y, void* stream);

/* nn.Upsample(None, 2, 'nearest') (torch; cfg yolo11.yaml head layers 11 / 14) on an NHWC bf16 window [B,H,W,C] ->
 * [B,2H,2W,C], written into a channel window of a wider buffer (y_pixstride): the Upsample -> Concat pair of the stock
 * YOLO11 neck (ultralytics/nn/modules/conv.py:1810-1820 Concat) without the intermediate tensor. */
int specyolo_upsample2x(const void* x, int x_pixstride, int B, int H, int W, int C,
                        void* y, int y_pixstride, void* stream);

/* ---- Conv + BN fold + weight repack ------------------------------------------------------ */
/* Folds BatchNorm into the conv weights exactly like fuse_conv_and_bn
 * (ultralytics/utils/torch_utils.py:238-265) and repacks OIHW fp32 -> the K-major bf16 layout the
 * implicit-GEMM kernel consumes: w_packed[g'][n_pad][kh][kw][cin_g'], rows n >= cout_g' zero.
 * `merge` (>= 1, divides groups) fuses that many source groups into one packed group with block-diagonal
 * weights: g' = groups/merge, cin_g' = cin_g*merge, cout_g' = cout/groups*merge; the packed conv is then
 * run with groups = g'.  Use specyolo_conv_merge() for the factor the kernels prefer and
 * specyolo_conv_npad(cout, g') for n_pad.
 * bn_* may be NULL (plain nn.Conv2d, e.g. the last conv of Detect.cv2/cv3); conv_bias may be NULL.
 * All pointers are device pointers.  bias_out has g'*n_pad floats. */
int specyolo_fold_pack_conv(const float* w_oihw, const float* conv_bias,
                            const float* bn_gamma, const float* bn_beta,
                            const float* bn_mean, const float* bn_var, float bn_eps,
                            int cout, int cin_g, int kh, int kw, int groups, int merge, int n_pad,
                            void* w_packed, float* bias_out, void* stream);
/* group-merge factor preferred for a grouped conv with a k x k kernel of the given stride / padding / dilation
 * (1 for dense and depthwise convs, and for grouped convs the halo-tile kernel runs group by group) */
int specyolo_conv_merge(int cin, int cout, int groups, int k, int stride, int pad, int dil);
/* n_pad (per-group padded output channels) the kernels expect for a conv with `groups` (packed) groups */
int specyolo_conv_npad(int cout, int groups);

typedef struct {
    /* input window */
    const void* x;        /* bf16 NHWC, first element of the channel window            */
    int B, H, W, Cin;     /* Cin = channels read (all groups)                          */
    int x_pixstride;      /* elements between consecutive pixels of x                  */
    int x_upshift;        /* 0, or 1: x is read through a nearest-neighbour x2 upsample
                             (nn.Upsample folded into the consumer; H,W are the upsampled dims;
                             1x1 convs only)                                            */
    /* weights (from specyolo_fold_pack_conv) */
    const void*  w_packed;
    const float* bias;
    int Cout, n_pad;      /* logical output channels; padded rows per group            */
    int kh, kw, stride, pad, dil, groups;
    int act;              /* SPECYOLO_ACT_*                                            */
    /* output window */
    void* y;              /* bf16 (or fp32 if y_fp32) NHWC window                      */
    int Ho, Wo;           /* output size; may be smaller than the geometry gives (far-edge rows/columns skipped) */
    int y_pixstride;
    int y_fp32;
    /* optional residual added after the activation (Bottleneck shortcut,
       ultralytics/nn/modules/block.py:724-726); bf16 NHWC window with Cout channels */
    const void* residual;
    int r_pixstride;
    /* 1: y is written 2x2-blocked (space-to-depth): pixel (oh, ow), channel c goes to the NHWC tensor
       [B, Ho/2, Wo/2, 4*Cout] at channel ((oh%2)*2 + ow%2)*Cout + c; y_pixstride is that tensor's pixel stride.
       A following 3x3 / stride-2 / pad-1 conv then reads it as a 2x2 / stride-1 conv over 4*Cout channels
       (halo-tile kernel only; Ho, Wo even; bf16 output). */
    int y_s2d;
} specyolo_conv_t;

/* y = act(conv(x, W) + b) [+ residual]  — replaces Conv.forward_fuse
 * (ultralytics/nn/modules/conv.py:81-83) after BaseModel.fuse (ultralytics/nn/tasks.py:223-251).
 * Dense and grouped convs with Cin/groups % 16 == 0 run on the tcgen05 implicit-GEMM kernels (k x k convs whose
 * weights fit in shared memory on the halo-tile variant, everything else on the per-tap variant);
 * depthwise 3x3 and the 3-channel stem run on dedicated CUDA-core kernels. */
int specyolo_conv2d_bias_act(const specyolo_conv_t* a, void* stream);

/* Fused DWConv(3x3, s1, p1)+BN+act -> Conv(1x1)+BN+act: the two building blocks of Detect.cv3
 * (ultralytics/nn/modules/head.py:51-58, Sequential(DWConv(x, x, 3), Conv(x, c3, 1))) in one kernel — the depthwise
 * result stays in shared memory as the A operand of the pointwise GEMM.  C in {64, 128}; SiLU after both. */
typedef struct {
    const void* x; int B, H, W, C, x_pixstride;   /* bf16 NHWC input window                                  */
    const float* dw_w;                            /* [9][C] BN-folded depthwise weights, tap = ky*3+kx (fp32) */
    const float* dw_b;                            /* [C]                                                     */
    const void* pw_packed; const float* pw_bias;  /* from specyolo_fold_pack_conv of the 1x1 conv (groups 1)  */
    int Cout, n_pad;
    void* y; int y_pixstride;                     /* bf16 NHWC output window, Cout channels (may be NULL with a head) */
    /* optional fused head: the closing nn.Conv2d(c3, nc, 1) of a Detect.cv3 branch (head.py:56) for nc <= 4 classes,
     * evaluated in the epilogue on the fp32 activations — head_y[pix * head_pixstride + j] = head_b[j] +
     * sum_c act[c] * head_w[j * Cout + c].  With a head the Cout-channel tensor is NOT stored (y is ignored). */
    const float* head_w;                          /* fp32 [nc][Cout] (device) or NULL                         */
    const float* head_b;                          /* fp32 [nc]                                                */
    float* head_y; int head_nc, head_pixstride;   /* fp32 rows of head_pixstride floats per pixel             */
} specyolo_dwpw_t;
int specyolo_dwconv_pwconv(const specyolo_dwpw_t* a, void* stream);

/* Fused Bottleneck: y = [x +] cv2(cv1(x)) with cv1 = Conv(C, Cmid, 3) and cv2 = Conv(Cmid, Cout, 3), both 3x3 / stride 1 /
 * pad 1 + BN + SiLU (ultralytics/nn/modules/block.py:713-726, the inner block of C3k2 :1659-1671) in ONE kernel: the Cmid-
 * channel intermediate stays in shared memory (14 x 14-pixel output tiles, 1-pixel halo recomputed), the shortcut is
 * added in the second epilogue.  w*_packed / b* come from specyolo_fold_pack_conv of the two convs (groups 1).
 * Shapes taken: C = Cout in {32, 64}, Cmid in {16, 32} (specyolo_bottleneck_ok); others run as two conv launches. */
typedef struct {
    const void* x; int B, H, W, C, x_pixstride;                 /* bf16 NHWC input window (also the shortcut)            */
    const void* w1_packed; const float* b1; int Cmid, n_pad1;
    const void* w2_packed; const float* b2; int Cout, n_pad2;
    int add;                                                    /* 1: shortcut (Bottleneck.add: shortcut and c1 == c2)    */
    void* y; int y_pixstride;                                   /* bf16 NHWC output window, Cout channels                */
} specyolo_bneck_t;
int specyolo_bottleneck_ok(int C, int Cmid, int Cout, int n_pad1, int n_pad2);
int specyolo_bottleneck(const specyolo_bneck_t* a, void* stream);

/* Stem conv reading the NCHW network input directly (first layer, Cin = 3):
 * x is NCHW fp32/bf16/u8 (u8 is scaled by 1/255 like predictor.py:133-135), w is fp32
 * [Cout][3][3][3] *folded* weights (device), output NHWC bf16 with SiLU. */
int specyolo_stem_conv3x3s2(const void* x, int x_dtype, int B, int H, int W,
                            const float* w_folded_oihw, const float* bias, int Cout,
                            void* y, int y_pixstride, void* stream);

/* Tensor-core route of the same layer: 2x2 space-to-depth of the NCHW input into an NHWC bf16 tensor
 * y[b, Y, X, (dy*2+dx)*3 + c] = x[b, c, 2Y+dy, 2X+dx] (channels 12..15 zero; uint8 stored unscaled).  The stem is then
 * specyolo_conv2d_bias_act with a 2x2 / stride-1 / pad-1 kernel over 16 channels and Ho = H/2, Wo = W/2 (the caller
 * repacks the 3x3 weights: tap (ty,tx), channel (dy,dx,c) <- w[c][2ty+dy-1][2tx+dx-1], times 1/255 for uint8). */
int specyolo_stem_space_to_depth(const void* x, int x_dtype, int B, int H, int W,
                                 void* y, int y_pixstride, void* stream);

/* Fused stem: layers 0 and 1 of the trunk (cfg yolo11*.yaml backbone[0:2], both Conv(k=3, s=2, p=1)+BN+SiLU) in one
 * kernel for uint8 NCHW input (the /255 of predictor.py:133-135 folded into w0); the layer-0 activations stay in shared
 * memory.  w0: fp16 [c0][32], column c*9 + ky*3 + kx holds the folded layer-0 weight / 255, columns 27..31 zero; the
 * kernel feeds the tensor core fp16 values 1024 + pixel, so b0 must be the folded layer-0 bias MINUS
 * 1024 * sum_k float(w0[n][k]).
 * w1_packed / b1: specyolo_fold_pack_conv of layer 1 rewritten as a 2x2 / stride-1 conv over the 2x2-blocked layer-0
 * map (K = 4 taps x 4*c0 channels; blocked channel (dy*2+dx)*c0 + ci, tap (ty,tx) <- w1[ci][2ty+dy-1][2tx+dx-1]).
 * Output NHWC bf16 [B, H/4, W/4, Cout].  specyolo_stem_pair_ok() tells whether the shape is taken. */
typedef struct {
    const void* x; int B, H, W;                   /* uint8 NCHW [B,3,H,W], H % 4 == 0, W % 16 == 0            */
    const void* w0; const float* b0; int c0;      /* layer 0: 16 or 32 channels                               */
    const void* w1_packed; const float* b1;
    int Cout, n_pad;
    void* y; int y_pixstride;
} specyolo_stem_pair_t;
int specyolo_stem_pair_ok(int H, int W, int c0, int Cout, int n_pad);
int specyolo_stem_pair(const specyolo_stem_pair_t* a, void* stream);

/* ---- SPPF pooling (ultralytics/nn/modules/block.py:194-198) ------------------------------ */
/* buf is the 4*c-channel concat buffer whose first c channels hold cv1(x); writes the three
 * chained MaxPool2d(5,1,2) results into channels [c,2c), [2c,3c), [3c,4c). */
int specyolo_sppf_pool(void* buf, int B, int H, int W, int c, int pixstride, void* stream);

/* ---- Fusion('ESChannel') (ultralytics/nn/modules/conv.py:2113-2127, GCT :2296-2301,
 *      WeightedSpatialAttention :1850-1852) ------------------------------------------------ */
typedef struct {
    int k;                     /* number of inputs, 2 or 3                              */
    const void* x[3];          /* bf16 NHWC, c channels each                            */
    int pixstride[3];
    int upshift[3];            /* 1: input i is a x2 nearest upsample of a (H/2,W/2) map */
    int B, H, W, c;
    const float* alpha;        /* GCT params over k*c channels                          */
    const float* gamma;
    const float* beta;
    float gct_eps;
    const float* sab_w;        /* [2][3][3] weights of the shared 2->1 3x3 conv         */
    void* y; int y_pixstride;  /* bf16 NHWC, c channels                                 */
    float* ws;                 /* workspace, specyolo_fusion_ws_bytes() bytes           */
} specyolo_fusion_t;
size_t specyolo_fusion_ws_bytes(int k, int B, int H, int W, int c);
int    specyolo_fusion_eschannel(const specyolo_fusion_t* a, void* stream);

/* ---- SobelSpatialAttention (ultralytics/nn/modules/conv.py:1184-1198), the gate of ConvHCA (conv.py:829-844;
 *      cfg yolo11_fusion_sand3_new_convHCA.yaml backbone layers 3, 5, 7) ------------------------------------------
 * y = x * sigmoid( cv1( sum_k sobel_k (*) cat(mean_c x, max_c x) ) ).  The three depthwise 3x3 convs (SobelConv,
 * conv.py:1153-1182: groups = 2, zero padding, no bias) and the 2 -> 1 1x1 conv are linear in the two statistic
 * planes; the caller passes them folded: w[c*9 + ky*3 + kx] = cv1[0][c] * sum_k sobel_k[c][0][ky][kx]. */
typedef struct {
    const void* x; int x_pixstride;   /* bf16 NHWC [B,H,W,C], C % 8 == 0                      */
    void* y; int y_pixstride;         /* bf16 NHWC, may alias x                               */
    int B, H, W, C;
    float w[18];
    float* mm;                        /* workspace: B * 2 * H * W floats (mean / max planes)  */
} specyolo_spatial_gate_t;
int specyolo_sobel_spatial_attention(const specyolo_spatial_gate_t* a, void* stream);

/* ---- MSCSpatialAttention (ultralytics/nn/modules/conv.py:1200-1243), the inner block of C3x (block.py:522-529;
 *      cfg yolo11_fusion_sand3_new_OMN.yaml head layer 21) -----------------------------------------------------
 * mm = cat(mean_c x, max_c x);  s = relu(cv1(mm)) + relu(cv2(mm))  (2 -> 1 convs, 31x31 and 3x3, zero padding, no
 * bias);  g = relu(fc(mean_hw(x * s)));  y = x * s * g + x.  Four launches, x read three times. */
typedef struct {
    const void* x; int x_pixstride;   /* bf16 NHWC [B,H,W,C], C % 8 == 0, C <= 512            */
    void* y; int y_pixstride;         /* bf16 NHWC, may alias x                               */
    int B, H, W, C;
    const float* w_big; int k_big;    /* device [2][k_big][k_big] (cv1.0.weight), k_big == 31 */
    const float* w_small;             /* device [2][3][3]         (cv2.0.weight)              */
    const float* fc_w;                /* device [C][C] (out, in)  (fc.weight)                 */
    const float* fc_b;                /* device [C]               (fc.bias)                   */
    float* ws;                        /* workspace, specyolo_msc_ws_bytes() bytes             */
} specyolo_msc_gate_t;
size_t specyolo_msc_ws_bytes(int B, int H, int W, int C);
int    specyolo_msc_spatial_attention(const specyolo_msc_gate_t* a, void* stream);

/* ---- BottleNect + FGM (ultralytics/nn/modules/block.py:782-861), the inner block of C3k2GC (block.py:1706-1714;
 *      cfg yolo11_fusion_sand3_new_GC.yaml backbone layer 2).  The reference runs two torch.fft (cuFFT) round trips;
 *      see csrc/bottlenect.cu for the algebra.  All weights are fp32 device arrays, 1x1 convs as [C][C] (out, in). --- */
typedef struct {
    const void* x; int x_pixstride;   /* bf16 NHWC [B,H,W,C], C in {16, 32}                    */
    void* y; int y_pixstride;         /* bf16 NHWC, may alias x                                */
    int B, H, W, C;                   /* H, W: prime factors <= 19, plane must fit shared memory */
    const float* in_w;  const float* in_b;     /* in_conv.0                                    */
    const float* fac_w; const float* fac_b;    /* fac_conv                                     */
    const float* sca_w; const float* sca_b;    /* conv                                         */
    const float* dw1_w; const float* dw1_b;    /* fgm.dwconv1                                  */
    const float* dw2_w; const float* dw2_b;    /* fgm.dwconv2                                  */
    const float* alpha; const float* beta;     /* fgm.alpha, fgm.beta [C]                      */
    float* ws;                        /* workspace, specyolo_bottlenect_ws_bytes() bytes       */
} specyolo_bottlenect_t;
size_t specyolo_bottlenect_ws_bytes(int B, int H, int W, int C);
int    specyolo_bottlenect(const specyolo_bottlenect_t* a, void* stream);

/* ---- PSA attention core (ultralytics/nn/modules/block.py:1922-1933) ----------------------- */
/* qkv: bf16 NHWC [B,N,heads*(2*kd+hd)] as written by the qkv 1x1 conv; out[b,n,h*hd+d] =
 * sum_j softmax_j(q_n.k_j*scale) v_j[d] + pe(v)[n] where pe is the depthwise 3x3 (+folded BN)
 * positional conv; pe_w is fp32 [heads*hd][3][3], pe_b fp32 [heads*hd]. */
int specyolo_psa_attention(const void* qkv, int qkv_pixstride, int B, int H, int W,
                           int heads, int key_dim, int head_dim, float scale,
                           const float* pe_w, const float* pe_b,
                           void* out, int out_pixstride, void* stream);

/* ---- Detect decode (ultralytics/nn/modules/head.py:100-131, block.py:80-83 DFL,
 *      utils/tal.py:334-358 make_anchors/dist2bbox) ---------------------------------------- */
typedef struct {
    int nl;                    /* number of levels (3)                                  */
    const float* logits[4];    /* per level fp32 [B, h*w, no_stride]: 64 box bins then nc cls */
    int h[4], w[4];
    float stride[4];
    int no_stride;             /* floats per anchor row (>= 64+nc)                      */
    int B, nc, reg_max;        /* reg_max = 16                                          */
    float* y;                  /* dense fp32 [B, 4+nc, A] (xywh, sigmoid scores); may be NULL */
    /* fused score-threshold output (may be NULL): anchors with max score > conf_thres,
       in anchor order inside fixed segments of SPECYOLO_DECODE_SEG anchors:
       cand[(b*nseg+s)*SEG + i] = {x1,y1,x2,y2,conf,cls}, seg_count[b*nseg+s] = #valid */
    float conf_thres;
    float* cand;
    int*   seg_count;
} specyolo_decode_t;
#define SPECYOLO_DECODE_SEG 256
int specyolo_detect_decode(const specyolo_decode_t* a, void* stream);

/* ---- NMS (ultralytics/utils/ops.py:181-332 + torchvision.ops.nms at ops.py:312) ----------- */
typedef struct {
    int B, nc, A;
    /* input, one of: */
    const float* prediction;   /* dense [B, 4+nc, A] xywh + scores (ops.py input), or NULL */
    const float* cand;         /* segmented candidates from specyolo_detect_decode        */
    const int*   seg_count;
    float  conf_thres;
    double iou_thres;          /* compared as double against the fp32 IoU like torchvision's CPU kernel */
    int    agnostic;
    int    multi_label;        /* dense input only (ops.py:286-288)                      */
    int    max_det, max_nms;
    float  max_wh;
    const int* classes; int n_classes;   /* optional class filter (ops.py:294-295), device ptr */
    /* output */
    float* out;                /* [B, max_det, 6] x1,y1,x2,y2,conf,cls                   */
    int*   out_count;          /* [B]                                                    */
    int*   keep_idx;           /* [B, max_det] indices into the post-threshold candidate list
                                  (the `i` of ops.py:312-313); may be NULL               */
    int*   n_cand;             /* [B] number of candidates that entered NMS; may be NULL */
    void*  ws;                 /* workspace, specyolo_nms_ws_bytes() bytes               */
    /* > 0: the kept boxes are clamped to [0, clip_w] x [0, clip_h] as they are written — the clip_boxes that
       ends every scale_boxes of construct_result (ultralytics/utils/ops.py:124-127, 335-354), which the
       reference also runs when the source already has the network shape (gain 1, pad 0).  0: boxes as NMS saw them. */
    float  clip_w, clip_h;
} specyolo_nms_t;
size_t specyolo_nms_ws_bytes(int B, int nc, int A, int multi_label);
int    specyolo_nms(const specyolo_nms_t* a, void* stream);

/* scale_boxes + clip_boxes on the NMS output (ultralytics/utils/ops.py:92-127, 335-354);
 * out rows [b, i<count[b], 0:4] are rewritten in place.  pad_w/pad_h/gain as computed by the
 * caller from the reference formula. */
int specyolo_scale_boxes(float* out, const int* out_count, int B, int max_det,
                         float gain, float pad_w, float pad_h, float img0_w, float img0_h,
                         void* stream);

/* ---- validation: detection <-> ground-truth matching -------------------------------
 * Replaces box_iou (ultralytics/utils/metrics.py:52-73) + DetectionValidator.match_predictions
 * (ultralytics/engine/validator.py:224-264, use_scipy=False) per image of a batch.
 * pred [B, max_det, 6] / pred_count [B]: the NMS output (specyolo_nms); labels [n, 5] = cls, x1, y1, x2, y2 in the
 * same pixel coordinates, grouped by image with label_off [B+1] (device, prefix offsets); iouv_host: niou IoU
 * thresholds (HOST pointer, <= 16); correct [B, max_det, niou] uint8 (device), rows >= pred_count[b] are zero. */
int specyolo_match_predictions(const float* pred, const int* pred_count, int B, int max_det,
                               const float* labels, const int* label_off, int max_labels_per_image,
                               const float* iouv_host, int niou, uint8_t* correct, void* stream);

/* ---- training criterion (first slice) --------------------------------------------------
 * Replaces v8DetectionLoss.__call__ (ultralytics/utils/loss.py:222-275): TaskAlignedAssigner.forward
 * (utils/tal.py:57-118), BCEWithLogits class loss, BboxLoss (CIoU, utils/metrics.py:171-228) and DFLoss
 * (utils/loss.py:66-129) — forward and the gradient with respect to the head outputs.
 * pred_distri [B, A, 4*reg_max] / pred_scores [B, A, nc]: fp32 logits in the anchor order of make_anchors
 * (tal.py:334-346: level by level, row-major), i.e. the tensors loss.py:226-231 builds; gt_boxes [B, M, 4] xyxy in
 * input pixels and gt_labels [B, M], the first gt_count[b] rows of image b valid (loss.py:193-209 + mask_gt :243).
 * out[6] = box, cls, dfl (each with its gain), (box + cls + dfl) * B, max(target_scores.sum(), 1), positives.
 * grad_distri / grad_scores (may be NULL): d out[3] / d pred_distri, d out[3] / d pred_scores. */
typedef struct {
    int nl; int h[4], w[4]; float stride[4];
    int B, nc, reg_max;
    const float* pred_distri; const float* pred_scores;
    int M; const float* gt_boxes; const int* gt_labels; const int* gt_count;
    int topk; float alpha, beta, tal_eps;             /* TaskAlignedAssigner(topk=10, alpha=0.5, beta=6.0, eps=1e-9) */
    float gain_box, gain_cls, gain_dfl;               /* hyp.box, hyp.cls, hyp.dfl                                  */
    float* out; float* grad_distri; float* grad_scores;
    void* ws;                                         /* specyolo_det_loss_ws_bytes() bytes, 256-byte aligned        */
} specyolo_det_loss_t;
size_t specyolo_det_loss_ws_bytes(int B, const int* h, const int* w, int nl, int M, int topk);
int    specyolo_det_loss(const specyolo_det_loss_t* a, void* stream);

/* ModelEMA.update (ultralytics/utils/torch_utils.py:514-524): ema[t][i] = ema[t][i] * d + model[t][i] * one_minus_d for
 * every fp32 tensor t of the state_dict, in one launch over a chunk list, with the reference's roundings (three fp32
 * roundings per element, no FMA), so the result is bit-identical to the Python loop.  All tables are DEVICE arrays:
 * ema / model [n] device pointers, numel [n], chunk_tensor / chunk_off [nchunks] (chunk c covers elements
 * [chunk_off[c], min(chunk_off[c] + chunk, numel[chunk_tensor[c]])) of tensor chunk_tensor[c]). */
int specyolo_ema_update(float* const* ema, const float* const* model, const long long* numel,
                        const int* chunk_tensor, const long long* chunk_off, int nchunks, int chunk,
                        float d, float one_minus_d, void* stream);

/* ---- image ingest: LetterBox for uint8 images ----------------------------------------
 * Replaces LetterBox.__call__ (ultralytics/data/augment.py:1535-1601: cv2.resize INTER_LINEAR + copyMakeBorder) and the
 * BGR->RGB / HWC->CHW step of BasePredictor.preprocess (ultralytics/engine/predictor.py:125-136), bit-exact with
 * OpenCV's 8-bit bilinear arithmetic.  src: [B,H,W,3] uint8 (device).  The caller computes the geometry with the
 * reference's formulas (content new_w x new_h placed at (left, top) inside out_h x out_w, rest = pad_value);
 * dst: [B,3,out_h,out_w] if chw else [B,out_h,out_w,3]; swap_rb exchanges channels 0 and 2. */
int specyolo_letterbox_u8(const uint8_t* src_hwc, int B, int H, int W, uint8_t* dst, int out_h, int out_w,
                          int new_w, int new_h, int left, int top, int pad_value, int swap_rb, int chw,
                          void* stream);

/* JPEG file bytes (host) -> uint8 HWC BGR image in device memory, the layout of cv2.imread that specyolo_letterbox_u8
 * consumes.  Replaces the cv2.imread of LoadImagesAndVideos.__next__ (ultralytics/data/loaders.py:406) for JPEG sources;
 * nvJPEG (library, dlopen'ed at first use) does Huffman on the host and IDCT / upsampling / colour conversion on the GPU.
 * specyolo_jpeg_info fills the image size; out must hold H*W*3 bytes.  The decode is asynchronous on `stream`. */
int specyolo_jpeg_info(const void* data, size_t nbytes, int* H, int* W, int* channels);
int specyolo_jpeg_decode_bgr(const void* data, size_t nbytes, void* out_dev, int H, int W, void* stream);
/* n streams in one nvjpegDecodeBatched call; backend 2 = GPU-assisted Huffman, 3 = hardware engine; out_dev[i] holds
 * H_i * W[i] * 3 bytes.  Returns SPECYOLO_ERR_UNSUPPORTED when the back end / a stream is not supported (fall back to
 * specyolo_jpeg_decode_bgr). */
int specyolo_jpeg_decode_batch_bgr(const void* const* data, const size_t* nbytes, void* const* out_dev, const int* W, int n,
                                   int backend, void* stream);

/* ---- IQ -> spectrogram -> letterbox (no reference implementation: README.md:7 only) -------- */
typedef struct {
    const float* iq;           /* [B, L] complex64 interleaved (re,im)                   */
    int B, L;
    int nfft, hop;             /* nfft = 1024                                            */
    float db_min, db_max;      /* dBFS range mapped to [0,1]                             */
    int out_h, out_w;          /* 640, 640                                               */
    float pad_value;           /* 114/255                                                */
    void* out;                 /* [B,3,out_h,out_w] NCHW, bf16 or fp32                   */
    int out_fp32;
} specyolo_stft_t;
int specyolo_iq_to_letterbox(const specyolo_stft_t* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPECYOLO_H */
