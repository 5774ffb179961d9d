// Fused Detect decode: DFL softmax-integral + dist2bbox + x stride + sigmoid(cls) + score threshold.
//
// Replaces Detect._inference (ultralytics/nn/modules/head.py:100-131), DFL.forward
// (ultralytics/nn/modules/block.py:80-83), make_anchors / dist2bbox (ultralytics/utils/tal.py:334-358)
// and the candidate selection of non_max_suppression (ultralytics/utils/ops.py:250, 289-291).
//
// One thread per anchor, 256 anchors (= one SPECYOLO_DECODE_SEG segment) per CTA.  Input rows are the
// fp32 head logits [B, h*w, no_stride] written by the last 1x1 convs of Detect.cv2 / cv3 (64 DFL
// bins then nc class logits).  HBM-bound: reads no*4 B per anchor, writes (4+nc)*4 B dense output
// (optional) plus 24 B per surviving candidate.  Everything is fp32.
#include "common.h"
#include "ptx.cuh"
#include "tma_host.h"

namespace specyolo {

struct DecodeParams {
    specyolo_decode_t a;
    int A;          // total anchors
    int nseg;       // segments per image
    int lvl_off[5]; // anchor offset of each level
};

__global__ void __launch_bounds__(SPECYOLO_DECODE_SEG)
detect_decode_kernel(const __grid_constant__ DecodeParams p) {
    ptx::grid_dep_launch();
    ptx::grid_dep_wait();
    const specyolo_decode_t& a = p.a;
    const int b = blockIdx.y;
    const int seg = blockIdx.x;
    const int anchor = seg * SPECYOLO_DECODE_SEG + threadIdx.x;
    const bool valid = anchor < p.A;

    // ---- many classes (nc >= 16, no dense output): the per-anchor maximum class logit is gathered COOPERATIVELY — eight
    // lanes read one anchor's class logits as consecutive float4 (128-byte runs; a thread walking its own 576-byte-strided
    // row issues nc 4-byte requests of 32 sectors each; measured at nc = 80, 1280^2, batch 128: 771 -> 458 us), four anchors per warp step,
    // every warp serving its own 32 anchors.  The value is the same maximum of the same numbers: decisions unchanged.
    __shared__ float s_lmax[SPECYOLO_DECODE_SEG];
    const bool coop = a.y == nullptr && a.nc >= 16;
    if (coop) {
        const int lane = threadIdx.x & 31, sub = lane & 7, grp = lane >> 3;
        for (int it = 0; it < 8; ++it) {
            const int t = (threadIdx.x & ~31) + it * 4 + grp;
            const int anc = seg * SPECYOLO_DECODE_SEG + t;
            float m = -INFINITY;
            if (anc < p.A) {
                int l = 0;
                while (l + 1 < a.nl && anc >= p.lvl_off[l + 1]) ++l;
                const float* cls = a.logits[l] + ((size_t)b * a.h[l] * a.w[l] + (anc - p.lvl_off[l])) * a.no_stride + 4 * a.reg_max;
                for (int j = sub * 4; j < a.nc; j += 32) {
                    if (j + 3 < a.nc) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(cls + j));
                        m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
                    } else {
                        for (int c = j; c < a.nc; ++c) m = fmaxf(m, __ldg(cls + c));
                    }
                }
            }
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            if (sub == 0) s_lmax[t] = m;
        }
        __syncwarp();
    }

    float cx = 0.f, cy = 0.f, bw = 0.f, bh = 0.f, conf = -1.f;
    int best = 0;
    if (valid) {
        int l = 0;
        while (l + 1 < a.nl && anchor >= p.lvl_off[l + 1]) ++l;
        const int local = anchor - p.lvl_off[l];
        const int gw = a.w[l];
        const float* row = a.logits[l] + ((size_t)b * a.h[l] * gw + local) * a.no_stride;
        const float* cls = row + 4 * a.reg_max;
        float* ydense = a.y ? a.y + (size_t)b * (4 + a.nc) * p.A + anchor : nullptr;

        // ---- scores first.  Without the dense output only survivors of the threshold need their box, and >= 97 % of the
        // anchors fail it: those read just the nc class logits of their row (not the 64 DFL bins) and evaluate ONE sigmoid.
        // sigmoid is monotonic as computed (expf, 1 + e, reciprocal), so  max_c sigmoid(l_c) == sigmoid(max_c l_c)  as a
        // number: the threshold decision is bit-identical to the reference order (sigmoid every class, then max,
        // head.py:100-131 + ops.py:250); survivors then redo the exact loop so that ties pick the reference's class.
        bool need_box = ydense != nullptr;
        if (!need_box) {
            float lmax;
            if (coop) {
                lmax = s_lmax[threadIdx.x];
            } else {
                lmax = __ldg(cls);
                for (int c = 1; c < a.nc; ++c) lmax = fmaxf(lmax, __ldg(cls + c));
            }
            need_box = (1.0f / (1.0f + expf(-lmax))) > a.conf_thres;
        }
        if (need_box) {
            for (int c = 0; c < a.nc; ++c) {
                const float sc = 1.0f / (1.0f + expf(-__ldg(cls + c)));
                if (ydense) ydense[(size_t)(4 + c) * p.A] = sc;
                if (sc > conf) { conf = sc; best = c; }   // strict '>' keeps the first maximum (torch.max)
            }
            const float ax = (float)(local % gw) + 0.5f;   // make_anchors: cell centre
            const float ay = (float)(local / gw) + 0.5f;
            // DFL: softmax over reg_max bins, expectation with weights 0..reg_max-1
            float dist[4];
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                float v[16];
                const float4* r4 = reinterpret_cast<const float4*>(row + s * 16);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t = __ldg(r4 + q);
                    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
                }
                float m = v[0];
#pragma unroll
                for (int k = 1; k < 16; ++k) m = fmaxf(m, v[k]);
                float sum = 0.f, wsum = 0.f;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const float e = expf(v[k] - m);
                    sum += e;
                    wsum = fmaf((float)k, e, wsum);
                }
                dist[s] = wsum / sum;
            }
            // dist2bbox (xywh=True) then * stride
            const float x1 = ax - dist[0], y1 = ay - dist[1];
            const float x2 = ax + dist[2], y2 = ay + dist[3];
            const float st = a.stride[l];
            cx = (x1 + x2) * 0.5f * st;
            cy = (y1 + y2) * 0.5f * st;
            bw = (x2 - x1) * st;
            bh = (y2 - y1) * st;
            if (ydense) {
                ydense[0] = cx;
                ydense[(size_t)p.A] = cy;
                ydense[(size_t)2 * p.A] = bw;
                ydense[(size_t)3 * p.A] = bh;
            }
        }
    }

    if (a.cand == nullptr) return;
    // ordered compaction of the survivors of this segment
    const bool pass = valid && (conf > a.conf_thres);
    __shared__ int warp_cnt[SPECYOLO_DECODE_SEG / 32];
    const unsigned ball = __ballot_sync(0xffffffffu, pass);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[warp] = __popc(ball);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < SPECYOLO_DECODE_SEG / 32; ++w) {
        if (w < warp) base += warp_cnt[w];
        total += warp_cnt[w];
    }
    if (pass) {
        const int slot = base + __popc(ball & ((1u << lane) - 1u));
        float* o = a.cand + ((size_t)(b * p.nseg + seg) * SPECYOLO_DECODE_SEG + slot) * 6;
        const float hw = bw * 0.5f, hh = bh * 0.5f;   // xywh2xyxy (ops.py:445-448)
        o[0] = cx - hw; o[1] = cy - hh; o[2] = cx + hw; o[3] = cy + hh;
        o[4] = conf; o[5] = (float)best;
    }
    if (threadIdx.x == 0) a.seg_count[b * p.nseg + seg] = total;
}

int detect_decode_launch(const specyolo_decode_t* a, cudaStream_t stream) {
    SY_CHECK(a->nl >= 1 && a->nl <= 4, SPECYOLO_ERR_INVALID, "nl must be 1..4");
    SY_CHECK(a->reg_max == 16, SPECYOLO_ERR_UNSUPPORTED, "only reg_max == 16 is supported");
    SY_CHECK(a->no_stride >= 4 * a->reg_max + a->nc && a->no_stride % 4 == 0, SPECYOLO_ERR_INVALID,
             "no_stride must be >= 64+nc and a multiple of 4");
    SY_CHECK((a->cand == nullptr) == (a->seg_count == nullptr), SPECYOLO_ERR_INVALID,
             "cand and seg_count must be given together");
    DecodeParams p{};
    p.a = *a;
    int off = 0;
    for (int l = 0; l < a->nl; ++l) {
        SY_CHECK((reinterpret_cast<uintptr_t>(a->logits[l]) & 15) == 0, SPECYOLO_ERR_INVALID,
                 "logits must be 16-byte aligned");
        p.lvl_off[l] = off;
        off += a->h[l] * a->w[l];
    }
    p.lvl_off[a->nl] = off;
    p.A = off;
    p.nseg = ceil_div(off, SPECYOLO_DECODE_SEG);
    dim3 grid((unsigned)p.nseg, (unsigned)a->B);
    SY_CUDA(launch_pdl(detect_decode_kernel, grid, dim3(SPECYOLO_DECODE_SEG), 0, stream, p));
    SY_LAUNCH_CHECK();
    count_launch();
    return SPECYOLO_OK;
}

}  // namespace specyolo
