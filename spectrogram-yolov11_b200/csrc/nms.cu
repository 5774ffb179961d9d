// Batched class-aware NMS, bit-exact with the reference's
//   non_max_suppression (ultralytics/utils/ops.py:181-332)  ->  torchvision.ops.nms (ops.py:312).
//
// One 1024-thread CTA per image, everything for that image in a single launch (the reference does a
// Python loop over images with ~6 host syncs each):
//   0. candidates in anchor order (dense input: conf mask + best class or multi-label expansion,
//      ordered block compaction; fused input: the segments written by detect_decode),
//   1. stable descending sort by score: bitonic sort of 64-bit keys (~score_bits << 32 | index) —
//      equal scores keep the lower candidate index first, like torchvision's stable sort,
//   2. greedy suppression in sorted order.  Up to 768 candidates (every predict() call in practice): all pairwise
//      "i suppresses j" bits (j behind i in score order) are computed in parallel into an n x n/32 bitmask in shared
//      memory, then ONE warp walks the candidates in order with the 'removed' set spread over its lanes (lane l = word l)
//      — torchvision's own CUDA scheme, the same greedy result as the sequential definition.  Beyond that,
//      1024 candidates per sweep: every thread tests its box
//      against the kept list (shared memory, <= max_det entries), then the 32 warps resolve their 32
//      candidates in turn with a warp-level bitmask (ballot + shuffles), publishing newly kept boxes
//      to the warps behind them.  Stops as soon as max_det boxes are kept (ops.py:313).
// IoU arithmetic reproduces torchvision's CPU kernel operation by operation in fp32 with explicit
// round-to-nearest intrinsics (no FMA contraction): boxes are offset by cls*max_wh in fp32
// (ops.py:305-311), inter/(area_i+area_j-inter) > thr with thr compared as a double.
#include "common.h"
#include "ptx.cuh"
#include "tma_host.h"

namespace specyolo {

static constexpr int kNmsThreads = 1024;
static constexpr int kSmemSort = 4096;   // keys held in shared memory up to this many (padded) entries
static constexpr int kSmemBoxes = 2048;  // sorted boxes held in shared memory up to this many
static constexpr int kMaxKeep = 2048;    // upper bound for max_det
static constexpr int kMaskN = 768;       // candidates up to which the suppression bitmask (n x n/32 words) fits in shared memory
static constexpr int kMaskWords = kMaskN / 32;

struct NmsWs {
    float4* cand_box;   // [B][cap] xyxy (not offset)
    float* cand_conf;   // [B][cap]
    int* cand_cls;      // [B][cap]
    unsigned long long* keys;  // [B][cap_pow2]
    float4* sorted_box; // [B][cap] offset boxes in sorted order (global fallback)
    int cap, cap_pow2;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

static size_t nms_ws_layout(int B, int nc, int A, int multi_label, void* base, NmsWs* ws) {
    const int cap = multi_label ? A * (nc > 1 ? nc : 1) : A;
    const int cap2 = next_pow2(cap);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_box = take((size_t)B * cap * sizeof(float4));
    const size_t o_conf = take((size_t)B * cap * sizeof(float));
    const size_t o_cls = take((size_t)B * cap * sizeof(int));
    const size_t o_keys = take((size_t)B * cap2 * sizeof(unsigned long long));
    const size_t o_sbox = take((size_t)B * cap * sizeof(float4));
    if (ws) {
        char* b = reinterpret_cast<char*>(base);
        ws->cand_box = reinterpret_cast<float4*>(b + o_box);
        ws->cand_conf = reinterpret_cast<float*>(b + o_conf);
        ws->cand_cls = reinterpret_cast<int*>(b + o_cls);
        ws->keys = reinterpret_cast<unsigned long long*>(b + o_keys);
        ws->sorted_box = reinterpret_cast<float4*>(b + o_sbox);
        ws->cap = cap;
        ws->cap_pow2 = cap2;
    }
    return off;
}

struct NmsParams {
    specyolo_nms_t a;
    NmsWs ws;
    int nseg;
    long long* dbg;      // SPECYOLO_NMS_DBG=1: [B][5] cycles at the end of phases 0, 1, sorted boxes, 2, output; [B][5] = n
};

// torchvision CPU nms_kernel_impl arithmetic, fp32, no contraction
__device__ __forceinline__ float box_area(const float4& b) {
    return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}
__device__ __forceinline__ bool iou_gt(const float4& bi, float ai, const float4& bj, float aj, double thr) {
    const float xx1 = fmaxf(bi.x, bj.x), yy1 = fmaxf(bi.y, bj.y);
    const float xx2 = fminf(bi.z, bj.z), yy2 = fminf(bi.w, bj.w);
    const float w = fmaxf(0.f, __fsub_rn(xx2, xx1));
    const float h = fmaxf(0.f, __fsub_rn(yy2, yy1));
    const float inter = __fmul_rn(w, h);
    // Disjoint boxes (the overwhelming majority of pairs): quotient 0 (or NaN for a degenerate union) is never > thr >= 0.
    if (!(inter > 0.f)) return false;
    const float uni = __fsub_rn(__fadd_rn(ai, aj), inter);
    // Certain decisions without the IEEE division (~35 instructions, and n^2/2 of them bounded the kernel: 210 k of its
    // 230 k cycles at n = 409): outside a 1e-4 relative band around thr * union the comparison cannot flip; inside the
    // band — and for degenerate unions — the exact fp32 quotient is compared with the double threshold as torchvision does.
    const float q = __fmul_rn(uni, (float)thr);
    if (uni > 0.f) {
        if (inter < q * 0.9999f) return false;
        if (inter > q * 1.0001f) return true;
    }
    const float ovr = __fdiv_rn(inter, uni);
    return (double)ovr > thr;
}

// block-wide exclusive scan of one int per thread (1024 threads); returns exclusive prefix, total in *total
__device__ __forceinline__ int block_exscan(int v, int* warp_sums, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int s = warp_sums[lane];
        int sinc = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, sinc, d);
            if (lane >= d) sinc += t;
        }
        warp_sums[lane] = sinc - s;  // exclusive
        if (lane == 31) warp_sums[32] = sinc;
    }
    __syncthreads();
    const int res = warp_sums[warp] + inc - v;
    *total = warp_sums[32];
    __syncthreads();
    return res;
}

__device__ __forceinline__ bool class_allowed(const int* classes, int n_classes, int c) {
    if (n_classes <= 0) return true;
    for (int i = 0; i < n_classes; ++i)
        if (classes[i] == c) return true;
    return false;
}

__global__ void __launch_bounds__(kNmsThreads, 1)
nms_kernel(const __grid_constant__ NmsParams p) {
    extern __shared__ __align__(16) uint8_t nms_smem[];
    ptx::grid_dep_launch();
    ptx::grid_dep_wait();
    unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(nms_smem);           // kSmemSort
    float4* s_boxes = reinterpret_cast<float4*>(nms_smem + (size_t)kSmemSort * 8);           // kSmemBoxes
    float4* s_kept = s_boxes + kSmemBoxes;                                                    // kMaxKeep
    float* s_kept_area = reinterpret_cast<float*>(s_kept + kMaxKeep);                         // kMaxKeep
    int* s_kept_rank = reinterpret_cast<int*>(s_kept_area + kMaxKeep);                        // kMaxKeep
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_kept_rank + kMaxKeep);                   // kMaskN x kMaskWords
    __shared__ int s_scan[33];
    __shared__ int s_n, s_kept_n, s_new_lo;

    const specyolo_nms_t& a = p.a;
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cap = p.ws.cap;
    const long long t_start = p.dbg ? clock64() : 0;
#define NMS_MARK(k) do { if (p.dbg && tid == 0) p.dbg[b * 6 + (k)] = clock64() - t_start; } while (0)
    float4* cbox = p.ws.cand_box + (size_t)b * cap;
    float* cconf = p.ws.cand_conf + (size_t)b * cap;
    int* ccls = p.ws.cand_cls + (size_t)b * cap;

    // ---------------- phase 0: candidates in anchor order ----------------
    int n = 0;
    if (a.prediction) {
        const float* pred = a.prediction + (size_t)b * (4 + a.nc) * a.A;
        for (int base = 0; base < a.A; base += kNmsThreads) {
            const int anc = base + tid;
            int cnt = 0;
            float best = -1.f;
            int bestc = 0;
            if (anc < a.A) {
                for (int c = 0; c < a.nc; ++c) {
                    const float sc = __ldg(pred + (size_t)(4 + c) * a.A + anc);
                    if (a.multi_label) {
                        if (sc > a.conf_thres && class_allowed(a.classes, a.n_classes, c)) ++cnt;
                    }
                    if (sc > best) { best = sc; bestc = c; }
                }
                if (!a.multi_label) {
                    cnt = (best > a.conf_thres && class_allowed(a.classes, a.n_classes, bestc)) ? 1 : 0;
                } else if (!(best > a.conf_thres)) {
                    cnt = 0;  // xc mask (ops.py:250) — implied by the per-class test, kept for clarity
                }
            }
            int total;
            int slot = n + block_exscan(cnt, s_scan, &total);
            if (cnt > 0) {
                const float x = __ldg(pred + anc), y = __ldg(pred + (size_t)a.A + anc);
                const float hw = __ldg(pred + (size_t)2 * a.A + anc) * 0.5f;
                const float hh = __ldg(pred + (size_t)3 * a.A + anc) * 0.5f;
                const float4 bx = make_float4(x - hw, y - hh, x + hw, y + hh);  // xywh2xyxy
                if (!a.multi_label) {
                    cbox[slot] = bx; cconf[slot] = best; ccls[slot] = bestc;
                } else {
                    for (int c = 0; c < a.nc; ++c) {
                        const float sc = __ldg(pred + (size_t)(4 + c) * a.A + anc);
                        if (sc > a.conf_thres && class_allowed(a.classes, a.n_classes, c)) {
                            cbox[slot] = bx; cconf[slot] = sc; ccls[slot] = c; ++slot;
                        }
                    }
                }
            }
            n += total;
        }
    } else {
        // fused input: segments of SPECYOLO_DECODE_SEG anchors written by detect_decode
        for (int sbase = 0; sbase < p.nseg; sbase += kNmsThreads) {
            const int s = sbase + tid;
            const int cnt_seg = (s < p.nseg) ? a.seg_count[b * p.nseg + s] : 0;
            // class filter is applied per record below, so scan the unfiltered counts first
            int total;
            const int off = n + block_exscan(cnt_seg, s_scan, &total);
            // stash per-segment offsets in the (not yet used) keys array — nseg <= 2*kSmemSort ints
            if (s < p.nseg) reinterpret_cast<int*>(s_keys)[s] = off;
            n += total;
        }
        __syncthreads();
        // copy records; one warp per segment
        for (int s = warp; s < p.nseg; s += kNmsThreads / 32) {
            const int off = reinterpret_cast<int*>(s_keys)[s];
            const int cnt_seg = a.seg_count[b * p.nseg + s];
            const float* src = a.cand + (size_t)(b * p.nseg + s) * SPECYOLO_DECODE_SEG * 6;
            for (int i = lane; i < cnt_seg; i += 32) {
                const float* r = src + i * 6;
                cbox[off + i] = make_float4(r[0], r[1], r[2], r[3]);
                cconf[off + i] = r[4];
                ccls[off + i] = (int)r[5];
            }
        }
        __syncthreads();
        if (a.n_classes > 0) {
            // optional class filter: stable in-place compaction (rarely used on the fused path)
            const int n_in = n;
            int n_out = 0;
            for (int base = 0; base < n_in; base += kNmsThreads) {
                const int i = base + tid;
                float4 bx = make_float4(0, 0, 0, 0); float cf = 0.f; int cl = 0; int keep = 0;
                if (i < n_in) {
                    bx = cbox[i]; cf = cconf[i]; cl = ccls[i];
                    keep = class_allowed(a.classes, a.n_classes, cl) ? 1 : 0;
                }
                int total;
                const int slot = n_out + block_exscan(keep, s_scan, &total);
                if (keep) { cbox[slot] = bx; cconf[slot] = cf; ccls[slot] = cl; }
                n_out += total;
                __syncthreads();
            }
            n = n_out;
        }
    }
    __syncthreads();
    NMS_MARK(0);
    if (p.dbg && tid == 0) p.dbg[b * 6 + 5] = n;
    if (a.n_cand && tid == 0) a.n_cand[b] = n < a.max_nms ? n : a.max_nms;

    int* out_count = a.out_count + b;
    if (n == 0) {
        if (tid == 0) *out_count = 0;
        return;
    }

    // ---------------- phase 1: stable descending sort by score ----------------
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    unsigned long long* keys = (n2 <= kSmemSort) ? s_keys : (p.ws.keys + (size_t)b * p.ws.cap_pow2);
    for (int i = tid; i < n2; i += kNmsThreads) {
        unsigned long long k = ~0ull;  // padding sorts last
        if (i < n) {
            const unsigned sb = __float_as_uint(cconf[i]);  // scores are positive: bit order == value order
            k = ((unsigned long long)(~sb) << 32) | (unsigned)i;
        }
        keys[i] = k;
    }
    __syncthreads();
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < (n2 >> 1); t += kNmsThreads) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const unsigned long long k0 = keys[lo], k1 = keys[hi];
                if ((k0 > k1) == up) { keys[lo] = k1; keys[hi] = k0; }
            }
            __syncthreads();
        }
    }
    NMS_MARK(1);
    // max_nms truncation (ops.py:301-302): keep the max_nms best; candidate indices then refer to
    // the score-sorted list, as in the reference after `x = x[argsort[:max_nms]]`.
    const bool truncated = n > a.max_nms;
    if (truncated) n = a.max_nms;

    // sorted, class-offset boxes
    float4* sboxes = (n <= kSmemBoxes) ? s_boxes : (p.ws.sorted_box + (size_t)b * cap);
    for (int r = tid; r < n; r += kNmsThreads) {
        const int idx = (int)(keys[r] & 0xffffffffu);
        float4 bx = cbox[idx];
        const float c = a.agnostic ? 0.f : __fmul_rn((float)ccls[idx], a.max_wh);
        bx.x = __fadd_rn(bx.x, c); bx.y = __fadd_rn(bx.y, c);
        bx.z = __fadd_rn(bx.z, c); bx.w = __fadd_rn(bx.w, c);
        sboxes[r] = bx;
    }
    if (tid == 0) { s_kept_n = 0; }
    __syncthreads();
    NMS_MARK(2);

    // ---------------- phase 2: greedy suppression ----------------
    const int max_det = a.max_det < kMaxKeep ? a.max_det : kMaxKeep;
    const double thr = a.iou_thres;
    if (n <= kMaskN) {
        // bitmask path: word (i, w) holds, for the 32 candidates 32w .. 32w+31 that come AFTER i, whether i suppresses them
        const int words = (n + 31) >> 5;
        float* s_area = s_kept_area;                  // areas of the sorted boxes (the kept-area array is unused on this path)
        for (int r = tid; r < n; r += kNmsThreads) s_area[r] = box_area(sboxes[r]);
        __syncthreads();
        // one warp per ROW i of the mask, lane = candidate 32w + lane: box reads are consecutive (a thread per word read
        // 32 boxes at a 512-byte lane stride: 32-way bank conflicts), box i is loaded once per row, only the words behind i
        // are visited and there is no index arithmetic in the loop (~20 instructions per word; a flat word loop with a
        // runtime division spent 108 k cycles here at n = 409)
        for (int i = warp; i < n; i += kNmsThreads / 32) {
            const float4 bi = sboxes[i];
            const float ai = s_area[i];
            for (int w = i >> 5; w < words; ++w) {
                const int c = 32 * w + lane;
                const bool sup = (c > i && c < n) && iou_gt(bi, ai, sboxes[c < n ? c : i], s_area[c < n ? c : i], thr);
                const uint32_t bits = __ballot_sync(0xffffffffu, sup);
                if (lane == 0) s_mask[i * words + w] = bits;
            }
        }
        __syncthreads();
        if (warp == 0) {
            // walk the candidates in score order; lane l holds word l of the 'removed' set; jump from kept box to kept box
            uint32_t removed = 0;
            int kept_n = 0, i = 0;
            while (i < n) {
                const int w = i >> 5;
                const uint32_t rw = __shfl_sync(0xffffffffu, removed, w);
                uint32_t alive = ~rw & (0xffffffffu << (i & 31));
                if (w == words - 1 && (n & 31)) alive &= (1u << (n & 31)) - 1u;
                if (!alive) { i = (w + 1) << 5; continue; }          // warp-uniform
                i = (w << 5) + __ffs((int)alive) - 1;
                if (lane == 0) s_kept_rank[kept_n] = i;
                if (++kept_n >= max_det) break;
                if (lane < words) removed |= s_mask[i * words + lane];    // (words behind i only: earlier ones are never read again)
                ++i;
            }
            if (lane == 0) s_kept_n = kept_n;
        }
        __syncthreads();
    } else
    for (int base = 0; base < n; base += kNmsThreads) {
        const int r = base + tid;
        bool alive = r < n;
        float4 mine = make_float4(0, 0, 0, 0);
        float my_area = 0.f;
        if (alive) { mine = sboxes[r]; my_area = box_area(mine); }
        int kept_n = s_kept_n;
        for (int k = 0; k < kept_n && alive; ++k)
            if (iou_gt(s_kept[k], s_kept_area[k], mine, my_area, thr)) alive = false;

        const int nwarps_here = min(kNmsThreads / 32, (n - base + 31) / 32);
        for (int w = 0; w < nwarps_here; ++w) {
            if (warp == w) {
                // resolve this warp's 32 candidates in score order
                for (int i = 0; i < 32; ++i) {
                    const unsigned am = __ballot_sync(0xffffffffu, alive);
                    if (!((am >> i) & 1u)) continue;
                    const float bx = __shfl_sync(0xffffffffu, mine.x, i);
                    const float by = __shfl_sync(0xffffffffu, mine.y, i);
                    const float bz = __shfl_sync(0xffffffffu, mine.z, i);
                    const float bw = __shfl_sync(0xffffffffu, mine.w, i);
                    const float ba = __shfl_sync(0xffffffffu, my_area, i);
                    if (lane > i && alive && iou_gt(make_float4(bx, by, bz, bw), ba, mine, my_area, thr))
                        alive = false;
                }
                const unsigned am = __ballot_sync(0xffffffffu, alive);
                const int cur = s_kept_n;
                const int slot = cur + __popc(am & ((1u << lane) - 1u));
                if (alive && slot < kMaxKeep) {
                    s_kept[slot] = mine; s_kept_area[slot] = my_area; s_kept_rank[slot] = r;
                }
                __syncwarp();
                if (lane == 0) {
                    s_new_lo = cur;
                    int nn = cur + __popc(am);
                    s_kept_n = nn < kMaxKeep ? nn : kMaxKeep;
                }
            }
            __syncthreads();
            const int lo = s_new_lo, hi = s_kept_n;
            if (warp > w && alive) {
                for (int k = lo; k < hi && alive; ++k)
                    if (iou_gt(s_kept[k], s_kept_area[k], mine, my_area, thr)) alive = false;
            }
            if (hi >= max_det) break;   // uniform: read from shared after the barrier
            __syncthreads();            // s_new_lo/s_kept_n are rewritten by the next warp
        }
        __syncthreads();
        if (s_kept_n >= max_det) break;
    }
    __syncthreads();

    NMS_MARK(3);
    // ---------------- output ----------------
    const int kept = s_kept_n < max_det ? s_kept_n : max_det;
    for (int k = tid; k < kept; k += kNmsThreads) {
        const int r = s_kept_rank[k];
        const int idx = (int)(keys[r] & 0xffffffffu);
        float4 bx = cbox[idx];
        if (a.clip_w > 0.f) {       // clip_boxes (ops.py:335-354)
            bx.x = fminf(fmaxf(bx.x, 0.f), a.clip_w); bx.z = fminf(fmaxf(bx.z, 0.f), a.clip_w);
            bx.y = fminf(fmaxf(bx.y, 0.f), a.clip_h); bx.w = fminf(fmaxf(bx.w, 0.f), a.clip_h);
        }
        float* o = a.out + ((size_t)b * a.max_det + k) * 6;
        o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
        o[4] = cconf[idx]; o[5] = (float)ccls[idx];
        if (a.keep_idx) a.keep_idx[(size_t)b * a.max_det + k] = truncated ? r : idx;
    }
    if (tid == 0) *out_count = kept;
    NMS_MARK(4);
}

size_t nms_ws_bytes(int B, int nc, int A, int multi_label) {
    return nms_ws_layout(B, nc, A, multi_label, nullptr, nullptr) + 256;
}

int nms_launch(const specyolo_nms_t* a, cudaStream_t stream) {
    SY_CHECK(a->B > 0 && a->nc > 0 && a->A > 0, SPECYOLO_ERR_INVALID, "bad NMS sizes");
    SY_CHECK(a->conf_thres >= 0.f && a->conf_thres <= 1.f, SPECYOLO_ERR_INVALID,
             "Invalid Confidence threshold %f, valid values are between 0.0 and 1.0", a->conf_thres);
    SY_CHECK(a->iou_thres >= 0.0 && a->iou_thres <= 1.0, SPECYOLO_ERR_INVALID,
             "Invalid IoU %f, valid values are between 0.0 and 1.0", a->iou_thres);
    SY_CHECK((a->prediction != nullptr) != (a->cand != nullptr), SPECYOLO_ERR_INVALID,
             "exactly one of prediction / cand must be given");
    SY_CHECK(a->cand == nullptr || (a->seg_count != nullptr && !a->multi_label), SPECYOLO_ERR_INVALID,
             "fused candidates need seg_count and multi_label == 0");
    SY_CHECK(a->max_det >= 1 && a->max_det <= kMaxKeep, SPECYOLO_ERR_UNSUPPORTED, "max_det must be 1..%d", kMaxKeep);
    SY_CHECK(a->max_nms >= 1, SPECYOLO_ERR_INVALID, "max_nms must be >= 1");
    SY_CHECK(a->ws != nullptr && a->out != nullptr && a->out_count != nullptr, SPECYOLO_ERR_INVALID, "null output/workspace");
    NmsParams p{};
    p.a = *a;
    // multi_label &= nc > 1 (ops.py:255)
    if (a->nc <= 1) p.a.multi_label = 0;
    char* wsb = reinterpret_cast<char*>(a->ws);
    wsb = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(wsb) + 255) & ~(uintptr_t)255);
    nms_ws_layout(a->B, a->nc, a->A, a->multi_label, wsb, &p.ws);
    p.nseg = ceil_div(a->A, SPECYOLO_DECODE_SEG);
    SY_CHECK(p.nseg <= kSmemSort * 2, SPECYOLO_ERR_UNSUPPORTED, "too many anchors");
    const size_t smem = (size_t)kSmemSort * 8 + (size_t)(kSmemBoxes + kMaxKeep) * 16 + (size_t)kMaxKeep * 8 +
                        (size_t)kMaskN * kMaskWords * 4;
    // the attribute is per device: set it on every launch (a host-side table lookup, no device work)
    SY_CUDA(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    static long long* dbg_dev = nullptr;
    const bool dbg = env_flag("SPECYOLO_NMS_DBG");
    if (dbg) {
        if (!dbg_dev) SY_CUDA(cudaMalloc(&dbg_dev, 4096 * 6 * sizeof(long long)));
        SY_CUDA(cudaMemsetAsync(dbg_dev, 0, 4096 * 6 * sizeof(long long), stream));
        p.dbg = a->B <= 4096 ? dbg_dev : nullptr;
    }
    SY_CUDA(launch_pdl(nms_kernel, dim3(a->B), dim3(kNmsThreads), smem, stream, p));
    SY_LAUNCH_CHECK();
    count_launch();
    if (dbg && p.dbg) {
        static long long h[4096 * 6];
        SY_CUDA(cudaStreamSynchronize(stream));
        SY_CUDA(cudaMemcpy(h, dbg_dev, (size_t)a->B * 6 * sizeof(long long), cudaMemcpyDeviceToHost));
        long long mx[6] = {0, 0, 0, 0, 0, 0};
        int worst = 0;
        for (int i = 0; i < a->B; ++i) {
            if (h[i * 6 + 4] > h[worst * 6 + 4]) worst = i;
            for (int k = 0; k < 6; ++k) mx[k] = h[i * 6 + k] > mx[k] ? h[i * 6 + k] : mx[k];
        }
        fprintf(stderr, "[nms dbg] B=%d max n=%lld | cycles at end of: candidates %lld sort %lld sorted boxes %lld suppression %lld output %lld "
                "(slowest image %d: n=%lld, %lld %lld %lld %lld %lld)\n", a->B, mx[5], mx[0], mx[1], mx[2], mx[3], mx[4], worst, h[worst * 6 + 5],
                h[worst * 6], h[worst * 6 + 1], h[worst * 6 + 2], h[worst * 6 + 3], h[worst * 6 + 4]);
    }
    return SPECYOLO_OK;
}

// scale_boxes + clip_boxes (ultralytics/utils/ops.py:92-127, 335-354)
__global__ void scale_boxes_kernel(float* out, const int* cnt, int B, int max_det, float gain, float pad_w,
                                   float pad_h, float w0, float h0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * max_det) return;
    const int b = i / max_det, k = i % max_det;
    if (k >= cnt[b]) return;
    float* o = out + (size_t)i * 6;
    float x1 = (o[0] - pad_w) / gain, y1 = (o[1] - pad_h) / gain;
    float x2 = (o[2] - pad_w) / gain, y2 = (o[3] - pad_h) / gain;
    o[0] = fminf(fmaxf(x1, 0.f), w0); o[1] = fminf(fmaxf(y1, 0.f), h0);
    o[2] = fminf(fmaxf(x2, 0.f), w0); o[3] = fminf(fmaxf(y2, 0.f), h0);
}

int scale_boxes_launch(float* out, const int* cnt, int B, int max_det, float gain, float pad_w, float pad_h,
                       float w0, float h0, cudaStream_t stream) {
    const int total = B * max_det;
    scale_boxes_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(out, cnt, B, max_det, gain, pad_w, pad_h, w0, h0);
    SY_LAUNCH_CHECK();
    count_launch();
    return SPECYOLO_OK;
}

}  // namespace specyolo
