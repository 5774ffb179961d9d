// Validation-path kernel: matching of detections to ground-truth boxes at the 10 mAP IoU thresholds.
// Replaces, per image, box_iou (ultralytics/utils/metrics.py:52-73) + DetectionValidator.match_predictions
// (ultralytics/engine/validator.py:224-264, the default use_scipy=False branch), which the reference runs as a
// torch N x M broadcast followed by a numpy loop over the thresholds on the host.
//
// Reference semantics, restated: IoU is zeroed where classes differ; for a threshold t take all (label, detection)
// pairs with IoU >= t, sort them by IoU descending, keep for every detection its first pair (= its highest-IoU label),
// re-order by detection index, keep for every label its first pair (= the LOWEST detection index, i.e. the most
// confident detection, NMS output being sorted by confidence).  Equivalently: best(d) = argmax_l IoU[l, d]; a
// detection is correct at t iff IoU[best(d), d] >= t and d is the smallest such detection with that best label.
// (Exact IoU ties between two labels of one detection fall to numpy's unstable argsort in the reference; here the
// lower label index wins.)
//
// One CTA per image; detections in registers/shared memory, labels streamed from global memory.
#include "common.h"
#include "ptx.cuh"
#include "tma_host.h"

namespace specyolo {

static constexpr int kMatchThreads = 256;
static constexpr int kMatchMaxDet = 2048;      // detections per image held in shared memory
static constexpr int kMatchMaxIou = 16;

struct MatchParams {
    const float* pred;          // [B, max_det, 6] x1,y1,x2,y2,conf,cls
    const int* pred_count;      // [B]
    int B, max_det;
    const float* labels;        // [n, 5] cls, x1, y1, x2, y2 (pixels), grouped by image
    const int* label_off;       // [B + 1]
    float iouv[kMatchMaxIou];
    int niou;
    uint8_t* correct;           // [B, max_det, niou]
};

__global__ void __launch_bounds__(kMatchThreads)
match_predictions_kernel(const __grid_constant__ MatchParams p) {
    ptx::grid_dep_launch();
    ptx::grid_dep_wait();
    extern __shared__ int match_smem[];
    int* best_l = match_smem;                                   // [max_det] best label (index within the image) or -1
    float* best_iou = reinterpret_cast<float*>(best_l + p.max_det);   // [max_det]
    int* win = reinterpret_cast<int*>(best_iou + p.max_det);    // [n_labels of this image] winning detection
    const int b = blockIdx.x;
    const int n = min(p.pred_count[b], p.max_det);
    const int l0 = p.label_off[b], nl = p.label_off[b + 1] - l0;
    const float* pr = p.pred + (size_t)b * p.max_det * 6;
    for (int d = threadIdx.x; d < n; d += kMatchThreads) {
        const float x1 = pr[d * 6 + 0], y1 = pr[d * 6 + 1], x2 = pr[d * 6 + 2], y2 = pr[d * 6 + 3];
        const float cls = pr[d * 6 + 5];
        const float area_d = (x2 - x1) * (y2 - y1);
        int bl = -1;
        float bi = 0.f;
        for (int l = 0; l < nl; ++l) {
            const float* g = p.labels + (size_t)(l0 + l) * 5;
            if (g[0] != cls) continue;                          // iou * correct_class (validator.py:240-241)
            // box_iou (metrics.py:68-73): inter / (area1 + area2 - inter + eps), box1 = label, box2 = detection
            const float w = fmaxf(fminf(g[3], x2) - fmaxf(g[1], x1), 0.f);
            const float h = fmaxf(fminf(g[4], y2) - fmaxf(g[2], y1), 0.f);
            const float inter = w * h;
            const float iou = inter / ((g[3] - g[1]) * (g[4] - g[2]) + area_d - inter + 1e-7f);
            if (iou > bi) { bi = iou; bl = l; }
        }
        best_l[d] = bl;
        best_iou[d] = bi;
    }
    __syncthreads();
    for (int t = 0; t < p.niou; ++t) {
        const float thr = p.iouv[t];
        for (int l = threadIdx.x; l < nl; l += kMatchThreads) win[l] = 0x7fffffff;
        __syncthreads();
        for (int d = threadIdx.x; d < n; d += kMatchThreads)
            if (best_l[d] >= 0 && best_iou[d] >= thr) atomicMin(&win[best_l[d]], d);
        __syncthreads();
        for (int d = threadIdx.x; d < p.max_det; d += kMatchThreads) {
            const bool ok = d < n && best_l[d] >= 0 && best_iou[d] >= thr && win[best_l[d]] == d;
            p.correct[((size_t)b * p.max_det + d) * p.niou + t] = ok ? 1 : 0;
        }
        __syncthreads();
    }
}

int match_predictions_launch(const float* pred, const int* pred_count, int B, int max_det, const float* labels,
                             const int* label_off, int max_labels_per_image, const float* iouv, int niou,
                             uint8_t* correct, cudaStream_t stream) {
    SY_CHECK(niou >= 1 && niou <= kMatchMaxIou, SPECYOLO_ERR_INVALID, "match: niou must be in [1, %d]", kMatchMaxIou);
    SY_CHECK(max_det >= 1 && max_det <= kMatchMaxDet, SPECYOLO_ERR_UNSUPPORTED, "match: max_det must be <= %d", kMatchMaxDet);
    const size_t smem = (size_t)max_det * 8 + (size_t)(max_labels_per_image > 0 ? max_labels_per_image : 1) * 4;
    SY_CHECK(smem <= 200 * 1024, SPECYOLO_ERR_UNSUPPORTED, "match: too many labels per image (%d)", max_labels_per_image);
    MatchParams p{};
    p.pred = pred; p.pred_count = pred_count; p.B = B; p.max_det = max_det;
    p.labels = labels; p.label_off = label_off; p.niou = niou; p.correct = correct;
    for (int i = 0; i < niou; ++i) p.iouv[i] = iouv[i];
    static size_t attr_smem[kMaxDevices] = {0};
    SY_CUDA(ensure_dynamic_smem(match_predictions_kernel, smem, attr_smem));
    SY_CUDA(launch_pdl(match_predictions_kernel, dim3((unsigned)B), dim3(kMatchThreads), smem, stream, p));
    SY_LAUNCH_CHECK();
    count_launch();
    return SPECYOLO_OK;
}

}  // namespace specyolo
