// Detection criterion for sm_100a (first slice): v8DetectionLoss.__call__ (ultralytics/utils/loss.py:222-275)
// = TaskAlignedAssigner (utils/tal.py:14-297, topk / alpha / beta from the caller) + BCE class loss + CIoU box loss
// (utils/metrics.py:171-228) + DFL (utils/loss.py:66-129), forward AND the gradient with respect to the head outputs, in
// six launches with no [B, n_max, A] tensor and no host synchronisation (the reference: ~150 ATen launches, a dozen
// B x n_max x A temporaries, `if fg_mask.sum()` / `max(target_scores.sum(), 1)` host round trips).
//
//   loss_decode_kernel     thread per (image, anchor): DFL softmax expectation -> box in grid units; clears the per-anchor
//                          assignment map and the box-logit gradient
//   tal_candidates_kernel  CTA per (ground-truth box, image): in-box test, CIoU (clamped at 0) and the alignment metric
//                          score^alpha * iou^beta of every anchor into shared memory, top-k by k block-wide arg-max rounds
//                          (ties: lower anchor index; zero-metric anchors are never taken — they carry target score 0 and
//                          change no loss term, see oracle/loss_ref.py)
//   tal_resolve_kernel     CTA per image: anchors claimed by several boxes go to the box with the largest CIoU over ALL
//                          boxes (tal.py:276-297), per-box maxima of metric / CIoU, normalised target score per positive
//   loss_fg_kernel         four lanes per positive (one per box side): CIoU with forward-mode derivatives, DFL
//                          cross-entropies, their gradients through the softmax expectation, the BCE terms of the target
//   loss_cls_kernel        thread per (image, anchor): sum softplus(x) and d/dx = (sigmoid(x) - target) * scale
//   loss_final_kernel      deterministic reduction of the per-block partials, gains, loss.sum() * B
// All sums are reduced in a fixed order (no floating-point atomics): the result is run-to-run identical.
#include "common.h"
#include "tma_host.h"

#include <cfloat>

namespace specyolo {

static constexpr int kLossMaxLevels = 4;
static constexpr int kLossThreads = 256;
static constexpr int kTalMaxTopk = 16;

struct LossParams {
    specyolo_det_loss_t a;
    int A;                            // anchors per image
    int lvl_off[kLossMaxLevels + 1];  // anchor offset of each level
    int no;                           // 4 * reg_max + nc
    // workspace carve-up
    float* box;          // [B][A][4]  decoded boxes, grid units
    int* amap;           // [B][A]     index of the anchor's positive entry (-1: background)
    int* cand_idx;       // [B][M][topk]
    float* cand_val;     // [B][M][topk][2]  metric, overlap
    int* ent_anchor;     // [B][M*topk]   positives of an image: anchor, box, target score
    int* ent_gt;
    float* ent_t;
    int* ent_count;      // [B]
    float* tss_part;     // [B]
    float* part_fg;      // [fg_blocks][3]  box, dfl, cls-correction partial sums
    float* part_cls;     // [cls_blocks]
    int fg_blocks, cls_blocks;
};

__device__ __forceinline__ void anchor_geom(const LossParams& p, int a, float& ax, float& ay, float& stride) {
    int l = 0;
    while (l + 1 < p.a.nl && a >= p.lvl_off[l + 1]) ++l;
    const int local = a - p.lvl_off[l];
    const int w = p.a.w[l];
    const int y = local / w, x = local - y * w;
    ax = (float)x + 0.5f;
    ay = (float)y + 0.5f;
    stride = p.a.stride[l];
}

// ---- CIoU (metrics.py:199-228, xywh=False), plain and with derivatives with respect to box 1 -----------------------
__device__ __forceinline__ float ciou_plain(float x11, float y11, float x12, float y12, float x21, float y21, float x22, float y22) {
    const float eps = 1e-7f;
    const float w1 = x12 - x11, h1 = y12 - y11 + eps, w2 = x22 - x21, h2 = y22 - y21 + eps;
    const float inter = fmaxf(fminf(x12, x22) - fmaxf(x11, x21), 0.f) * fmaxf(fminf(y12, y22) - fmaxf(y11, y21), 0.f);
    const float uni = w1 * h1 + w2 * h2 - inter + eps;
    const float iou = inter / uni;
    const float cw = fmaxf(x12, x22) - fminf(x11, x21), ch = fmaxf(y12, y22) - fminf(y11, y21);
    const float c2 = cw * cw + ch * ch + eps;
    const float dx = x21 + x22 - x11 - x12, dy = y21 + y22 - y11 - y12;
    const float rho2 = (dx * dx + dy * dy) * 0.25f;
    const float da = atanf(w2 / h2) - atanf(w1 / h1);
    const float v = 0.4052847345693511f * da * da;          // 4 / pi^2
    const float alpha = v / (v - iou + (1.0f + eps));
    return iou - (rho2 / c2 + v * alpha);
}

struct D4 {            // value + derivatives with respect to (x1, y1, x2, y2) of the predicted box
    float v, d[4];
};
__device__ __forceinline__ D4 dconst(float c) { return D4{c, {0.f, 0.f, 0.f, 0.f}}; }
__device__ __forceinline__ D4 dvar(float c, int i) { D4 r = dconst(c); r.d[i] = 1.f; return r; }
__device__ __forceinline__ D4 operator+(D4 a, D4 b) { D4 r; r.v = a.v + b.v; for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
__device__ __forceinline__ D4 operator-(D4 a, D4 b) { D4 r; r.v = a.v - b.v; for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
__device__ __forceinline__ D4 operator*(D4 a, D4 b) { D4 r; r.v = a.v * b.v; for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
__device__ __forceinline__ D4 operator/(D4 a, D4 b) {
    D4 r; const float inv = 1.0f / b.v; r.v = a.v * inv;
    for (int i = 0; i < 4; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
    return r;
}
__device__ __forceinline__ D4 dclamp0(D4 a) { return a.v > 0.f ? a : dconst(fmaxf(a.v, 0.f)); }
__device__ __forceinline__ D4 datan(D4 a) {
    D4 r; r.v = atanf(a.v); const float g = 1.0f / (1.0f + a.v * a.v);
    for (int i = 0; i < 4; ++i) r.d[i] = a.d[i] * g;
    return r;
}

// CIoU(pred, target) and its gradient with respect to the predicted box; alpha is a constant (torch.no_grad, :226).
// minimum / maximum at equal arguments: torch splits the gradient evenly; the target is a constant, so either choice
// sends half of it to a constant — reproduced by averaging the two one-sided derivatives where the values tie.
__device__ __forceinline__ D4 dmin_t(D4 a, float b) {
    if (a.v < b) return a;
    if (a.v > b) return dconst(b);
    D4 r = a; for (int i = 0; i < 4; ++i) r.d[i] *= 0.5f; return r;
}
__device__ __forceinline__ D4 dmax_t(D4 a, float b) {
    if (a.v > b) return a;
    if (a.v < b) return dconst(b);
    D4 r = a; for (int i = 0; i < 4; ++i) r.d[i] *= 0.5f; return r;
}
__device__ __forceinline__ D4 ciou_grad(float px1, float py1, float px2, float py2, float tx1, float ty1, float tx2, float ty2) {
    const float eps = 1e-7f;
    const D4 x11 = dvar(px1, 0), y11 = dvar(py1, 1), x12 = dvar(px2, 2), y12 = dvar(py2, 3);
    const D4 w1 = x12 - x11, h1 = (y12 - y11) + dconst(eps);
    const float w2 = tx2 - tx1, h2 = ty2 - ty1 + eps;
    const D4 iw = dclamp0(dmin_t(x12, tx2) - dmax_t(x11, tx1));
    const D4 ih = dclamp0(dmin_t(y12, ty2) - dmax_t(y11, ty1));
    const D4 inter = iw * ih;
    const D4 uni = w1 * h1 + dconst(w2 * h2) - inter + dconst(eps);
    const D4 iou = inter / uni;
    const D4 cw = dmax_t(x12, tx2) - dmin_t(x11, tx1), ch = dmax_t(y12, ty2) - dmin_t(y11, ty1);
    const D4 c2 = cw * cw + ch * ch + dconst(eps);
    const D4 dx = dconst(tx1 + tx2) - x11 - x12, dy = dconst(ty1 + ty2) - y11 - y12;
    const D4 rho2 = (dx * dx + dy * dy) * dconst(0.25f);
    const D4 da = dconst(atanf(w2 / h2)) - datan(w1 / h1);
    const D4 v = da * da * dconst(0.4052847345693511f);
    const float alpha = v.v / (v.v - iou.v + (1.0f + eps));
    return iou - (rho2 / c2 + v * dconst(alpha));
}

// ---------------------------------------------------------------------------------------------------------------------
// four lanes per anchor (lane = box side): a warp reads 2 KB of consecutive logits per step and clears the same span of
// the gradient (thread-per-anchor reads were 256-byte strided: 7 % of the DRAM bandwidth under ncu)
__global__ void __launch_bounds__(kLossThreads)
loss_decode_kernel(const __grid_constant__ LossParams p) {
    const specyolo_det_loss_t& a = p.a;
    const long t = (long)blockIdx.x * kLossThreads + threadIdx.x;
    const long i = t >> 2;                       // (image, anchor)
    const int side = (int)(t & 3);
    const bool live = i < (long)a.B * p.A;       // whole groups of four lanes are live or not
    float dist = 0.f;
    if (live) {
        const float* row = a.pred_distri + i * (4 * a.reg_max) + side * a.reg_max;
        float mx = -FLT_MAX;
        for (int j = 0; j < a.reg_max; ++j) mx = fmaxf(mx, row[j]);
        float den = 0.f, num = 0.f;
        for (int j = 0; j < a.reg_max; ++j) {
            const float e = expf(row[j] - mx);
            den += e;
            num += e * (float)j;
        }
        dist = num / den;
        if (a.grad_distri) {
            float* g = a.grad_distri + i * (4 * a.reg_max) + side * a.reg_max;
            for (int j = 0; j < a.reg_max; ++j) g[j] = 0.f;
        }
    }
    const int base = (threadIdx.x & 31) & ~3;
    const float d0 = __shfl_sync(0xffffffffu, dist, base), d1 = __shfl_sync(0xffffffffu, dist, base + 1);
    const float d2 = __shfl_sync(0xffffffffu, dist, base + 2), d3 = __shfl_sync(0xffffffffu, dist, base + 3);
    if (live && side == 0) {
        float ax, ay, st;
        anchor_geom(p, (int)(i % p.A), ax, ay, st);
        reinterpret_cast<float4*>(p.box)[i] = make_float4(ax - d0, ay - d1, ax + d2, ay + d3);
        p.amap[i] = -1;
    }
}

// metric of (box g, anchor anc) of image b; returns false if the anchor centre is not strictly inside the box
__device__ __forceinline__ bool tal_metric(const LossParams& p, int b, int anc, const float4 gt, int label, float& metric, float& ov) {
    const specyolo_det_loss_t& a = p.a;
    float ax, ay, st;
    anchor_geom(p, anc, ax, ay, st);
    const float px = ax * st, py = ay * st;
    const float dmin = fminf(fminf(px - gt.x, py - gt.y), fminf(gt.z - px, gt.w - py));
    metric = 0.f;
    ov = 0.f;
    if (!(dmin > 1e-9f)) return false;
    const float4 bx = reinterpret_cast<const float4*>(p.box)[(long)b * p.A + anc];
    ov = fmaxf(ciou_plain(gt.x, gt.y, gt.z, gt.w, bx.x * st, bx.y * st, bx.z * st, bx.w * st), 0.f);
    const float x = a.pred_scores[((long)b * p.A + anc) * a.nc + label];
    const float sc = 1.0f / (1.0f + expf(-x));
    metric = powf(sc, a.alpha) * powf(ov, a.beta);
    return true;
}

__global__ void __launch_bounds__(kLossThreads)
tal_candidates_kernel(const __grid_constant__ LossParams p) {
    extern __shared__ float s_metric[];               // [A]
    __shared__ float s_val[kLossThreads / 32];
    __shared__ int s_idx[kLossThreads / 32];
    __shared__ int s_win;
    const specyolo_det_loss_t& a = p.a;
    const int g = blockIdx.x, b = blockIdx.y;
    int* cidx = p.cand_idx + ((long)b * a.M + g) * a.topk;
    float* cval = p.cand_val + ((long)b * a.M + g) * a.topk * 2;
    if (g >= a.gt_count[b]) {
        for (int k = threadIdx.x; k < a.topk; k += kLossThreads) cidx[k] = -1;
        return;
    }
    const float4 gt = reinterpret_cast<const float4*>(a.gt_boxes)[(long)b * a.M + g];
    const int label = a.gt_labels[(long)b * a.M + g];
    for (int anc = threadIdx.x; anc < p.A; anc += kLossThreads) {
        float m, ov;
        tal_metric(p, b, anc, gt, label, m, ov);
        s_metric[anc] = m;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = 0; k < a.topk; ++k) {
        float best = 0.f;               // only strictly positive metrics are taken
        int bi = 0x7fffffff;
        for (int anc = threadIdx.x; anc < p.A; anc += kLossThreads) {
            const float m = s_metric[anc];
            if (m > best) { best = m; bi = anc; }                       // ascending anc: first index wins ties
        }
        for (int d = 16; d > 0; d >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, d);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, d);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            float bb = s_val[0];
            int ii = s_idx[0];
            for (int w = 1; w < kLossThreads / 32; ++w)
                if (s_val[w] > bb || (s_val[w] == bb && s_idx[w] < ii)) { bb = s_val[w]; ii = s_idx[w]; }
            if (bb > 0.f) {
                float m, ov;
                tal_metric(p, b, ii, gt, label, m, ov);
                cidx[k] = ii;
                cval[2 * k] = m;
                cval[2 * k + 1] = ov;
                s_metric[ii] = -1.f;
                s_win = 1;
            } else {
                cidx[k] = -1;
                s_win = 0;
            }
        }
        __syncthreads();
        if (!s_win) {                   // fewer than topk positive metrics: the rest of the list is empty
            for (int kk = k + 1 + threadIdx.x; kk < a.topk; kk += kLossThreads) cidx[kk] = -1;
            break;
        }
    }
}

__global__ void __launch_bounds__(kLossThreads)
tal_resolve_kernel(const __grid_constant__ LossParams p) {
    extern __shared__ int s_raw[];
    const specyolo_det_loss_t& a = p.a;
    const int b = blockIdx.x;
    const int M = a.gt_count[b];
    const int words = (p.A + 31) / 32;
    unsigned* seen = reinterpret_cast<unsigned*>(s_raw);          // [words]
    unsigned* multi = seen + words;                               // [words]
    int* pos_align = reinterpret_cast<int*>(multi + words);       // [a.M] float bits (values >= 0: integer order = float order)
    int* pos_ov = pos_align + a.M;                                // [a.M]
    __shared__ int s_count;
    __shared__ float s_red[kLossThreads];
    for (int i = threadIdx.x; i < 2 * words; i += kLossThreads) seen[i] = 0u;
    for (int i = threadIdx.x; i < 2 * a.M; i += kLossThreads) pos_align[i] = 0;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    const int ncand = M * a.topk;
    const int* cidx = p.cand_idx + (long)b * a.M * a.topk;
    const float* cval = p.cand_val + (long)b * a.M * a.topk * 2;
    for (int c = threadIdx.x; c < ncand; c += kLossThreads) {
        const int anc = cidx[c];
        if (anc < 0) continue;
        const unsigned bit = 1u << (anc & 31);
        const unsigned old = atomicOr(&seen[anc >> 5], bit);
        if (old & bit) atomicOr(&multi[anc >> 5], bit);
    }
    __syncthreads();
    // every candidate decides whether it emits the positive of its anchor: a singly claimed anchor keeps its box; a
    // multiply claimed one goes to arg-max_g CIoU(g, anchor) over all boxes (first index on ties), emitted once
    int* e_anchor = p.ent_anchor + (long)b * a.M * a.topk;
    int* e_gt = p.ent_gt + (long)b * a.M * a.topk;
    float* e_t = p.ent_t + (long)b * a.M * a.topk;        // holds the metric until the normalisation below
    for (int c = threadIdx.x; c < ncand; c += kLossThreads) {
        const int anc = cidx[c];
        if (anc < 0) continue;
        int g = c / a.topk;
        float metric = cval[2 * c], ov = cval[2 * c + 1];
        if (multi[anc >> 5] & (1u << (anc & 31))) {
            bool first = true;                                     // is this the first candidate (in list order) of the anchor?
            for (int c2 = 0; c2 < c; ++c2)
                if (cidx[c2] == anc) { first = false; break; }
            if (!first) continue;
            float best_ov = -1.f, best_m = 0.f;
            int best_g = 0;
            for (int g2 = 0; g2 < M; ++g2) {
                float m2, o2;
                tal_metric(p, b, anc, reinterpret_cast<const float4*>(a.gt_boxes)[(long)b * a.M + g2], a.gt_labels[(long)b * a.M + g2], m2, o2);
                if (o2 > best_ov) { best_ov = o2; best_m = m2; best_g = g2; }
            }
            g = best_g; metric = best_m; ov = best_ov;
        }
        const int slot = atomicAdd(&s_count, 1);
        e_anchor[slot] = anc;
        e_gt[slot] = g;
        e_t[slot] = metric;
        atomicMax(&pos_align[g], __float_as_int(metric));
        atomicMax(&pos_ov[g], __float_as_int(ov));
    }
    __syncthreads();
    const int n = s_count;
    // the slot order above depends on the atomics: normalise (tal.py:112-116), then sort the (few) entries by anchor —
    // through the candidate lists, which nobody reads any more — so that every later sum runs in a fixed order
    int* t_anchor = p.cand_idx + (long)b * a.M * a.topk;
    float* t_val = p.cand_val + (long)b * a.M * a.topk * 2;
    for (int i = threadIdx.x; i < n; i += kLossThreads) {
        const int g = e_gt[i];
        const float pa = __int_as_float(pos_align[g]), po = __int_as_float(pos_ov[g]);
        t_anchor[i] = e_anchor[i];
        t_val[2 * i] = e_t[i] * po / (pa + a.tal_eps);
        t_val[2 * i + 1] = __int_as_float(g);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kLossThreads) {
        const int anc = t_anchor[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) rank += t_anchor[j] < anc;
        p.amap[(long)b * p.A + anc] = rank;
        e_anchor[rank] = anc;
        e_t[rank] = t_val[2 * i];
        e_gt[rank] = __float_as_int(t_val[2 * i + 1]);
    }
    __syncthreads();
    float local = 0.f;
    // target_scores.sum() of this image, fixed order
    for (int i = threadIdx.x; i < n; i += kLossThreads) local += e_t[i];
    s_red[threadIdx.x] = local;
    __syncthreads();
    for (int d = kLossThreads / 2; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) s_red[threadIdx.x] += s_red[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        p.ent_count[b] = n;
        p.tss_part[b] = s_red[0];
    }
}

__device__ __forceinline__ float loss_tss(const LossParams& p) {
    float t = 0.f;
    for (int b = 0; b < p.a.B; ++b) t += p.tss_part[b];
    return fmaxf(t, 1.0f);
}

// four lanes per positive entry (lane = box side), 64 entries per block
__global__ void __launch_bounds__(kLossThreads)
loss_fg_kernel(const __grid_constant__ LossParams p) {
    const specyolo_det_loss_t& a = p.a;
    __shared__ float s_red[3][kLossThreads / 4];
    const int side = threadIdx.x & 3;
    const int slot = threadIdx.x >> 2;
    const int per_img = a.M * a.topk;
    const long e = (long)blockIdx.x * (kLossThreads / 4) + slot;      // flat entry index over [B][M*topk]
    const int b = (int)(e / per_img), i = (int)(e - (long)b * per_img);
    const bool live = b < a.B && i < p.ent_count[b];
    float l_box = 0.f, l_dfl = 0.f, l_cls = 0.f;
    const float tss = loss_tss(p);
    const float scale = (float)a.B / tss;
    if (live) {
        const int anc = p.ent_anchor[(long)b * per_img + i];
        const int g = p.ent_gt[(long)b * per_img + i];
        const float t = p.ent_t[(long)b * per_img + i];
        float ax, ay, st;
        anchor_geom(p, anc, ax, ay, st);
        const float4 gt = reinterpret_cast<const float4*>(a.gt_boxes)[(long)b * a.M + g];
        const float inv_st = 1.0f / st;
        const float tx1 = gt.x * inv_st, ty1 = gt.y * inv_st, tx2 = gt.z * inv_st, ty2 = gt.w * inv_st;
        const long row = ((long)b * p.A + anc) * (4 * a.reg_max) + side * a.reg_max;
        const float* lg = a.pred_distri + row;
        float prob[32];
        float mx = -FLT_MAX;
        for (int j = 0; j < a.reg_max; ++j) mx = fmaxf(mx, lg[j]);
        float den = 0.f, num = 0.f;
        for (int j = 0; j < a.reg_max; ++j) {
            prob[j] = expf(lg[j] - mx);
            den += prob[j];
            num += prob[j] * (float)j;
        }
        const float inv_den = 1.0f / den;
        for (int j = 0; j < a.reg_max; ++j) prob[j] *= inv_den;
        const float E = num * inv_den;
        const float4 bx = reinterpret_cast<const float4*>(p.box)[(long)b * p.A + anc];
        const D4 c = ciou_grad(bx.x, bx.y, bx.z, bx.w, tx1, ty1, tx2, ty2);
        // d box / d dist: x1 = ax - l, y1 = ay - t, x2 = ax + r, y2 = ay + b
        const float dE = (side < 2 ? -c.d[side] : c.d[side]);            // d CIoU / d dist_side
        const float g_box = -t * dE * a.gain_box * scale;                // d (sum loss * B) / d dist_side
        // DFL target of this side (tal.py:361-364, loss.py:82-92)
        const float tgt_raw = side == 0 ? ax - tx1 : (side == 1 ? ay - ty1 : (side == 2 ? tx2 - ax : ty2 - ay));
        const float tgt = fminf(fmaxf(tgt_raw, 0.f), (float)(a.reg_max - 1) - 0.01f);
        const int tl = (int)tgt;
        const float wl = (float)(tl + 1) - tgt, wr = 1.0f - wl;
        const float lse = mx + logf(den);
        const float ce = (lse - lg[tl]) * wl + (lse - lg[tl + 1]) * wr;
        const float g_dfl = t * 0.25f * a.gain_dfl * scale;
        if (a.grad_distri) {
            float* gd = a.grad_distri + row;
            for (int j = 0; j < a.reg_max; ++j) {
                float gj = g_box * prob[j] * ((float)j - E) + g_dfl * prob[j];
                if (j == tl) gj -= g_dfl * wl;
                if (j == tl + 1) gj -= g_dfl * wr;
                gd[j] = gj;
            }
        }
        l_dfl = ce * 0.25f * t;
        if (side == 0) {
            l_box = (1.0f - c.v) * t;
            const int label = a.gt_labels[(long)b * a.M + g];
            l_cls = -a.pred_scores[((long)b * p.A + anc) * a.nc + label] * t;     // BCE(x, t) = softplus(x) - x t
        }
    }
    l_dfl += __shfl_xor_sync(0xffffffffu, l_dfl, 1);
    l_dfl += __shfl_xor_sync(0xffffffffu, l_dfl, 2);
    if (side == 0) { s_red[0][slot] = l_box; s_red[1][slot] = l_dfl; s_red[2][slot] = l_cls; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float t = 0.f;
        for (int k = 0; k < kLossThreads / 4; ++k) t += s_red[threadIdx.x][k];
        p.part_fg[(long)blockIdx.x * 3 + threadIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kLossThreads)
loss_cls_kernel(const __grid_constant__ LossParams p) {
    const specyolo_det_loss_t& a = p.a;
    __shared__ float s_red[kLossThreads];
    const long i = (long)blockIdx.x * kLossThreads + threadIdx.x;
    const float tss = loss_tss(p);
    const float scale = a.gain_cls * (float)a.B / tss;
    float local = 0.f;
    if (i < (long)a.B * p.A) {
        const int b = (int)(i / p.A);
        const int ent = p.amap[i];
        int label = -1;
        float t = 0.f;
        if (ent >= 0) {
            const int per_img = a.M * a.topk;
            label = a.gt_labels[(long)b * a.M + p.ent_gt[(long)b * per_img + ent]];
            t = p.ent_t[(long)b * per_img + ent];
        }
        const float* x = a.pred_scores + i * a.nc;
        float* gx = a.grad_scores ? a.grad_scores + i * a.nc : nullptr;
        for (int c = 0; c < a.nc; ++c) {
            const float v = x[c];
            const float e = expf(-fabsf(v));
            local += fmaxf(v, 0.f) + log1pf(e);                              // softplus
            if (gx) {
                const float sg = v >= 0.f ? 1.0f / (1.0f + e) : e / (1.0f + e);
                gx[c] = (sg - (c == label ? t : 0.f)) * scale;
            }
        }
    }
    s_red[threadIdx.x] = local;
    __syncthreads();
    for (int d = kLossThreads / 2; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d) s_red[threadIdx.x] += s_red[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) p.part_cls[blockIdx.x] = s_red[0];
}

__global__ void __launch_bounds__(kLossThreads)
loss_final_kernel(const __grid_constant__ LossParams p) {
    __shared__ double s_red[4][kLossThreads];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < p.fg_blocks; i += kLossThreads)
        for (int k = 0; k < 3; ++k) acc[k] += (double)p.part_fg[(long)i * 3 + k];
    for (int i = threadIdx.x; i < p.cls_blocks; i += kLossThreads) acc[3] += (double)p.part_cls[i];
    for (int k = 0; k < 4; ++k) s_red[k][threadIdx.x] = acc[k];
    __syncthreads();
    for (int d = kLossThreads / 2; d > 0; d >>= 1) {
        if ((int)threadIdx.x < d)
            for (int k = 0; k < 4; ++k) s_red[k][threadIdx.x] += s_red[k][threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double tss = (double)loss_tss(p);
        const double box = s_red[0][0] / tss * p.a.gain_box;
        const double dfl = s_red[1][0] / tss * p.a.gain_dfl;
        const double cls = (s_red[3][0] + s_red[2][0]) / tss * p.a.gain_cls;
        p.a.out[0] = (float)box;
        p.a.out[1] = (float)cls;
        p.a.out[2] = (float)dfl;
        p.a.out[3] = (float)((box + cls + dfl) * p.a.B);
        p.a.out[4] = (float)tss;
        int n = 0;
        for (int b = 0; b < p.a.B; ++b) n += p.ent_count[b];
        p.a.out[5] = (float)n;
    }
}

// ---- host side -------------------------------------------------------------------------------------------------------
static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

struct LossLayout {
    size_t box, amap, cand_idx, cand_val, ent_anchor, ent_gt, ent_t, ent_count, tss_part, part_fg, part_cls, total;
    int fg_blocks, cls_blocks;
};
static LossLayout loss_layout(int B, long A, int M, int topk) {
    LossLayout L{};
    size_t off = 0;
    const size_t per = (size_t)B * (M > 0 ? M : 1) * topk;
    L.fg_blocks = (int)((per + kLossThreads / 4 - 1) / (kLossThreads / 4));
    L.cls_blocks = (int)(((size_t)B * A + kLossThreads - 1) / kLossThreads);
    L.box = off; off = align_up(off + (size_t)B * A * 16);
    L.amap = off; off = align_up(off + (size_t)B * A * 4);
    L.cand_idx = off; off = align_up(off + per * 4);
    L.cand_val = off; off = align_up(off + per * 8);
    L.ent_anchor = off; off = align_up(off + per * 4);
    L.ent_gt = off; off = align_up(off + per * 4);
    L.ent_t = off; off = align_up(off + per * 4);
    L.ent_count = off; off = align_up(off + (size_t)B * 4);
    L.tss_part = off; off = align_up(off + (size_t)B * 4);
    L.part_fg = off; off = align_up(off + (size_t)L.fg_blocks * 12);
    L.part_cls = off; off = align_up(off + (size_t)L.cls_blocks * 4);
    L.total = off;
    return L;
}

size_t det_loss_ws_bytes(int B, const int* h, const int* w, int nl, int M, int topk) {
    long A = 0;
    for (int l = 0; l < nl; ++l) A += (long)h[l] * w[l];
    return loss_layout(B, A, M, topk).total;
}

int det_loss_launch(const specyolo_det_loss_t* a, cudaStream_t stream) {
    SY_CHECK(a->nl >= 1 && a->nl <= kLossMaxLevels, SPECYOLO_ERR_INVALID, "det loss: 1..4 levels");
    SY_CHECK(a->reg_max >= 2 && a->reg_max <= 32, SPECYOLO_ERR_UNSUPPORTED, "det loss: reg_max must be in 2..32");
    SY_CHECK(a->topk >= 1 && a->topk <= kTalMaxTopk, SPECYOLO_ERR_UNSUPPORTED, "det loss: topk must be in 1..16");
    SY_CHECK(a->B >= 1 && a->nc >= 1 && a->M >= 0, SPECYOLO_ERR_INVALID, "det loss: bad sizes");
    SY_CHECK((reinterpret_cast<uintptr_t>(a->pred_distri) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->ws) & 255) == 0 &&
                 (!a->grad_distri || (reinterpret_cast<uintptr_t>(a->grad_distri) & 15) == 0) &&
                 (a->M == 0 || (reinterpret_cast<uintptr_t>(a->gt_boxes) & 15) == 0),
             SPECYOLO_ERR_INVALID, "det loss: pred_distri / grad_distri / gt_boxes must be 16-byte aligned, ws 256-byte aligned");
    LossParams p{};
    p.a = *a;
    long A = 0;
    for (int l = 0; l < a->nl; ++l) {
        p.lvl_off[l] = (int)A;
        A += (long)a->h[l] * a->w[l];
    }
    p.lvl_off[a->nl] = (int)A;
    SY_CHECK(A > 0 && A * 4 <= 200 * 1024, SPECYOLO_ERR_UNSUPPORTED, "det loss: %ld anchors per image do not fit the candidate kernel", A);
    p.A = (int)A;
    p.no = 4 * a->reg_max + a->nc;
    const LossLayout L = loss_layout(a->B, A, a->M, a->topk);
    uint8_t* ws = reinterpret_cast<uint8_t*>(a->ws);
    p.box = reinterpret_cast<float*>(ws + L.box);
    p.amap = reinterpret_cast<int*>(ws + L.amap);
    p.cand_idx = reinterpret_cast<int*>(ws + L.cand_idx);
    p.cand_val = reinterpret_cast<float*>(ws + L.cand_val);
    p.ent_anchor = reinterpret_cast<int*>(ws + L.ent_anchor);
    p.ent_gt = reinterpret_cast<int*>(ws + L.ent_gt);
    p.ent_t = reinterpret_cast<float*>(ws + L.ent_t);
    p.ent_count = reinterpret_cast<int*>(ws + L.ent_count);
    p.tss_part = reinterpret_cast<float*>(ws + L.tss_part);
    p.part_fg = reinterpret_cast<float*>(ws + L.part_fg);
    p.part_cls = reinterpret_cast<float*>(ws + L.part_cls);
    p.fg_blocks = a->M > 0 ? L.fg_blocks : 0;
    p.cls_blocks = L.cls_blocks;

    const unsigned nblk = (unsigned)(((long)a->B * A * 4 + kLossThreads - 1) / kLossThreads);
    loss_decode_kernel<<<nblk, kLossThreads, 0, stream>>>(p);
    count_launch();
    if (a->M > 0) {
        const size_t smem_c = (size_t)A * 4;
        static size_t attr_c[kMaxDevices] = {0};
        SY_CUDA(ensure_dynamic_smem(tal_candidates_kernel, smem_c, attr_c));
        tal_candidates_kernel<<<dim3((unsigned)a->M, (unsigned)a->B), kLossThreads, smem_c, stream>>>(p);
        count_launch();
    }
    {
        const size_t smem_r = (size_t)2 * ((A + 31) / 32) * 4 + (size_t)2 * (a->M > 0 ? a->M : 1) * 4;
        SY_CHECK(smem_r <= 48 * 1024, SPECYOLO_ERR_UNSUPPORTED, "det loss: too many ground-truth boxes per image (%d)", a->M);
        tal_resolve_kernel<<<(unsigned)a->B, kLossThreads, smem_r, stream>>>(p);
        count_launch();
    }
    if (a->M > 0) {
        loss_fg_kernel<<<(unsigned)L.fg_blocks, kLossThreads, 0, stream>>>(p);
        count_launch();
    }
    loss_cls_kernel<<<(unsigned)L.cls_blocks, kLossThreads, 0, stream>>>(p);
    count_launch();
    loss_final_kernel<<<1, kLossThreads, 0, stream>>>(p);
    count_launch();
    SY_LAUNCH_CHECK();
    return SPECYOLO_OK;
}

}  // namespace specyolo
