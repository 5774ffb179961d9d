// Fused DWConv(3x3, s1, p1)+BN+SiLU -> Conv(1x1)+BN+SiLU for sm_100a — the two building blocks of Detect.cv3
// (ultralytics/nn/modules/head.py:51-58: Sequential(DWConv(x, x, 3), Conv(x, c3, 1))).
//
// Run separately, the depthwise layer writes and the pointwise layer re-reads a full [B,H,W,C] tensor, and the
// depthwise layer itself is either issue-bound on CUDA cores or multiplies 63/64 zeros on the tensor pipe.  Here the
// depthwise result never leaves the SM:
//   warp 0      TMA producer: per 16x8-pixel output tile one (18 x 10 pixel) halo box per 64 channels
//   warps 2-17  depthwise on CUDA cores: a thread owns a channel pair (its 9 folded weight pairs in registers, packed
//               fp32x2 math) of one or two tile rows, slides the 3x3 window along them reading the halo box from shared memory, applies bias + SiLU and
//               writes bf16 into the 128 x C tile laid out as the K-major SWIZZLE_128B A operand
//   warp 1      MMA issuer: A tile (shared memory, double-buffered) x resident 1x1 weights -> fp32 accumulator in TMEM
//               (tcgen05.mma, M = 128, N = Cout, K = C), double-buffered
//   warps 18-25 epilogue (epilogue.cuh): bias + SiLU -> bf16 -> staged TMA store
// HBM traffic: one read of the input (x1.4 halo) + one write of the output; the intermediate is 0 bytes.
#include "common.h"
#include "ptx.cuh"
#include "tma_host.h"
#include "epilogue.cuh"

namespace specyolo {

static constexpr int kDwpwThreads = 64 + 512 + 256;     // TMA + MMA warps, 16 depthwise warps, 8 epilogue warps
static constexpr int kDwpwTW = 8, kDwpwTH = 16, kDwpwHW = 10, kDwpwHH = 18;
static constexpr uint32_t kDwpwBoxBytes = kDwpwHW * kDwpwHH * 128;          // 23 040: one 64-channel halo box
static constexpr uint32_t kDwpwBoxStride = 23552;                            // rounded up to 1024
static constexpr uint32_t kDwpwAChunk = 128 * 128;                           // 16 KB: 128 pixels x 64 channels
static constexpr int kDwpwMaxDynSmem = 222 * 1024;

struct DwpwParams {
    int B, H, W, C, chunks;          // C channels = chunks * 64
    int tiles_w, tiles_h, spatial_tiles;
    FastDiv d_img, d_tw;
    int n_pad, cout;
    const float* dw_w;               // [9][C] folded depthwise weights (fp32)
    const float* dw_b;               // [C]
    const float* pw_b;               // [n_pad]
    uint32_t x_stage_bytes, w_bytes, a_buf_bytes, x_off, a_off, st_off;
    uint32_t tmem_cols;
    void* y;
    int y_pixstride;
    int store_bw, pair_stores;
    uint32_t store_row_bytes, store_swz_mask;
    // fused head (kHead): logits[pix][j] = head_b[j] + sum_c SiLU(pw(x))[c] * head_w[j][c], j < head_nc <= kDwpwMaxNc
    const float* head_w;
    const float* head_b;
    float* head_y;
    int head_nc, head_pixstride;
};
static constexpr int kDwpwMaxNc = 4;

// kHead: the pointwise result is not stored; it feeds the Detect branch's closing Conv2d(c3, nc, 1) (head.py:56) inside
// the epilogue — nc <= 4 dot products per pixel on the fp32 SiLU outputs — and only the nc class logits go to memory.
template <bool kHead>
__global__ void __launch_bounds__(kDwpwThreads, 1)
dwpw_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
            const __grid_constant__ CUtensorMap map_y, const __grid_constant__ DwpwParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t x_full[2], x_empty[2], a_full[2], a_empty[2], tmem_full_bar[2], tmem_empty_bar[2], w_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float bias_s[256];
    __shared__ __align__(16) float head_ws[kHead ? kDwpwMaxNc * 256 : 4];     // [nc][n_pad] fp32 (zero beyond cout)
    __shared__ __align__(16) float head_part[kHead ? 2 * 128 * kDwpwMaxNc : 4]; // [tile parity][row][nc]: partial sums of the second column half

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    ptx::grid_dep_launch();

    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* base = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* w_s = base;                       // resident 1x1 weights: `chunks` boxes of [n_pad rows x 128 B]
    uint8_t* x_ring = base + p.x_off;          // 2 stages x chunks halo boxes
    uint8_t* a_buf = base + p.a_off;           // 2 A tiles x chunks x 16 KB
    uint8_t* st_buf = base + p.st_off;         // epilogue staging

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&map_x);
        ptx::prefetch_tmap(&map_w);
        if (p.store_bw) ptx::prefetch_tmap(&map_y);
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&x_full[s], 1);
            ptx::mbar_init(&x_empty[s], 16);
            ptx::mbar_init(&a_full[s], 16);
            ptx::mbar_init(&a_empty[s], 1);
            ptx::mbar_init(&tmem_full_bar[s], 1);
            ptx::mbar_init(&tmem_empty_bar[s], kEpiWarps);
        }
        ptx::mbar_init(&w_bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc(&tmem_base_smem, p.tmem_cols);
    for (int i = threadIdx.x; i < p.n_pad; i += kDwpwThreads) bias_s[i] = 0.5f * p.pw_b[i];     // SiLU form (epilogue.cuh)
    if (kHead)
        for (int i = threadIdx.x; i < p.head_nc * p.n_pad; i += kDwpwThreads) {
            const int j = i / p.n_pad, c = i - j * p.n_pad;
            head_ws[j * p.n_pad + c] = c < p.cout ? p.head_w[j * p.cout + c] : 0.f;
        }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    if (warp != 0) ptx::grid_dep_wait();
    const int cta = blockIdx.x, ctas = gridDim.x;

    if (warp == 0) {
        // ===================== TMA producer =====================
        const bool leader = ptx::elect_one();
        if (leader) {
            ptx::mbar_expect_tx(&w_bar, p.w_bytes);
            for (int c = 0; c < p.chunks; ++c)
                ptx::tma_load_2d(w_s + (size_t)c * p.n_pad * 128, &map_w, &w_bar, c * 64, 0);
        }
        ptx::grid_dep_wait();
        uint32_t tl = 0;
        for (int tile = cta; tile < p.spatial_tiles; tile += ctas, ++tl) {
            uint32_t n, r, th_i, tw_i;
            fdivmod((uint32_t)tile, p.d_img, n, r);
            fdivmod(r, p.d_tw, th_i, tw_i);
            const uint32_t s = tl & 1u;
            ptx::mbar_wait(&x_empty[s], ((tl >> 1) & 1u) ^ 1u);
            if (leader) {
                ptx::mbar_expect_tx(&x_full[s], (uint32_t)p.chunks * kDwpwBoxBytes);
                for (int c = 0; c < p.chunks; ++c)
                    ptx::tma_load_4d(x_ring + s * p.x_stage_bytes + (uint32_t)c * kDwpwBoxStride, &map_x, &x_full[s], c * 64,
                                     (int)tw_i * kDwpwTW - 1, (int)th_i * kDwpwTH - 1, (int)n);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const bool leader = ptx::elect_one();
        const uint32_t idesc = ptx::umma_idesc_bf16(128, p.n_pad);
        const uint32_t hi = (uint32_t)(ptx::umma_smem_desc(0, 128) >> 32);
        const uint32_t a16 = ptx::smem_u32(a_buf) >> 4, w16 = ptx::smem_u32(w_s) >> 4;
        ptx::mbar_wait(&w_bar, 0);
        uint32_t tl = 0;
        for (int tile = cta; tile < p.spatial_tiles; tile += ctas, ++tl) {
            const uint32_t b = tl & 1u, ph = (tl >> 1) & 1u;
            ptx::mbar_wait(&tmem_empty_bar[b], ph ^ 1u);
            ptx::mbar_wait(&a_full[b], ph);
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + b * (uint32_t)p.n_pad;
            uint32_t acc = 0;
            for (int c = 0; c < p.chunks; ++c) {
                const uint32_t ac = a16 + ((b * p.a_buf_bytes + (uint32_t)c * kDwpwAChunk) >> 4);
                const uint32_t wc = w16 + (((uint32_t)c * (uint32_t)p.n_pad * 128u) >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (leader) ptx::umma_bf16_lohi(d_tmem, ac + 2u * k, hi, wc + 2u * k, hi, idesc, acc);
                    acc = 1;
                }
            }
            if (leader) {
                ptx::umma_commit(&a_empty[b]);
                ptx::umma_commit(&tmem_full_bar[b]);
            }
        }
    } else if (warp < 18) {
        // ===================== depthwise 3x3 + bias + SiLU -> A tile (warps 2..17) =====================
        // a thread owns ONE channel pair (its 9 folded weight pairs in registers) of one or two tile rows: 512 threads,
        // four warps per scheduler (with 4 channels per thread and 8 warps this role ran at ~6 k cycles per tile and
        // bounded the kernel)
        const int t = threadIdx.x - 64;                 // 0..511
        const int pairs = p.C >> 1;                     // channel pairs per pixel: 32 or 64
        const int cq = t % pairs;                       // this thread's channel pair
        const int rgrp = t / pairs;                     // row group: 512/pairs of them
        const int rows_per = (kDwpwTH * pairs) >> 9;    // tile rows per thread: 1 (C=64) or 2 (C=128)
        const int c0 = cq * 2;
        const int chunk = c0 >> 6;                      // which 64-channel box / A chunk
        const uint32_t unit = (uint32_t)(c0 & 63) >> 3; // 16-byte unit inside the 128-byte row
        const uint32_t sub = (uint32_t)(c0 & 7) * 2;    // byte offset inside the unit: 0, 4, 8 or 12
        // packed fp32 pairs (FFMA2): 9 multiply-adds per pixel, SiLU on the pair
        float2 wgt[9];
        const float2 hbs = make_float2(0.5f * p.dw_b[c0], 0.5f * p.dw_b[c0 + 1]);
#pragma unroll
        for (int k = 0; k < 9; ++k) wgt[k] = make_float2(p.dw_w[k * p.C + c0], p.dw_w[k * p.C + c0 + 1]);
        uint32_t tl = 0;
        for (int tile = cta; tile < p.spatial_tiles; tile += ctas, ++tl) {
            const uint32_t s = tl & 1u, ph = (tl >> 1) & 1u;
            ptx::mbar_wait(&x_full[s], ph);
            ptx::mbar_wait(&a_empty[s], ph ^ 1u);       // the MMAs that read this A buffer two tiles ago have retired
            const uint8_t* xs = x_ring + s * p.x_stage_bytes + (uint32_t)chunk * kDwpwBoxStride;
            uint8_t* as = a_buf + s * p.a_buf_bytes + (uint32_t)chunk * kDwpwAChunk;
            for (int rr = 0; rr < rows_per; ++rr) {
                const int th = rgrp * rows_per + rr;    // output row inside the tile
                // sliding 3x3 window over the 10 halo columns of rows th, th+1, th+2
                float2 col[3][3];
                const uint8_t* xrow = xs + (uint32_t)(th * kDwpwHW) * 128u + (uint32_t)((c0 & 63) * 2);
                auto load_col = [&](int slot, int hx) {
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        // halo boxes are loaded UNswizzled: only these warps read them, and a warp reads the 32 consecutive
                        // channel pairs of one pixel = one 128-byte row, conflict-free as it is; the address is then a
                        // compile-time offset from the row base instead of ~4 integer ops per load
                        col[slot][ky] = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(xrow + (uint32_t)((ky * kDwpwHW + hx) * 128)));
                    }
                };
                load_col(0, 0);
                load_col(1, 1);
#pragma unroll
                for (int tw = 0; tw < kDwpwTW; ++tw) {
                    load_col((tw + 2) % 3, tw + 2);
                    // two accumulation chains per pixel for instruction-level parallelism
                    float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky) {
                            if ((kx * 3 + ky) & 1) a1 = ffma2(col[(tw + kx) % 3][ky], wgt[ky * 3 + kx], a1);
                            else a0 = ffma2(col[(tw + kx) % 3][ky], wgt[ky * 3 + kx], a0);
                        }
                    const float2 sv = silu2_half(fadd2(a0, a1), hbs);      // SiLU(acc + b) = h + h tanh(h), h = acc/2 + b/2
                    const uint32_t m = (uint32_t)(th * kDwpwTW + tw);      // A row = pixel inside the tile
                    *reinterpret_cast<uint32_t*>(as + m * 128u + ((unit ^ (m & 7u)) << 4) + sub) = pack_bf16x2(sv.x, sv.y);
                }
            }
            ptx::fence_proxy_async();                   // generic-proxy writes of A -> visible to UMMA
            __syncwarp();
            if (lane == 0) {
                ptx::mbar_arrive(&a_full[s]);
                ptx::mbar_arrive(&x_empty[s]);
            }
        }
    } else {
        // ===================== epilogue (warps 18..25, see epilogue.cuh) =====================
        const int quad = warp & 3;
        const int half = (warp - 18) >> 2;
        const int m = quad * 32 + lane;
        const int tw = m & (kDwpwTW - 1), th = m >> 3;
        EpiOut eo{p.y, p.y_pixstride, nullptr, 0, p.pair_stores != 0};
        EpiStage st = epi_make_stage(st_buf, &map_y, p.n_pad, p.store_bw, p.store_row_bytes, p.store_swz_mask, warp - 16, half,
                                     lane, m);
        EpiCols ec;
        ec.ncols = p.n_pad;
        ec.n_pad = 1 << 20;
        ec.d_npad = FastDiv{1ull << 20, 1u << 20};
        ec.cout_g = p.cout;
        ec.within0 = 0;
        ec.gch0 = 0;
        uint32_t tl = 0;
        for (int tile = cta; tile < p.spatial_tiles; tile += ctas, ++tl) {
            uint32_t n, r, th_i, tw_i;
            fdivmod((uint32_t)tile, p.d_img, n, r);
            fdivmod(r, p.d_tw, th_i, tw_i);
            const uint32_t b = tl & 1u, ph = (tl >> 1) & 1u;
            const int ow = (int)tw_i * kDwpwTW + tw, oh = (int)th_i * kDwpwTH + th;
            const bool row_ok = (ow < p.W) && (oh < p.H);
            const size_t pix = ((size_t)n * p.H + oh) * p.W + ow;
            const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + b * (uint32_t)p.n_pad;
            ptx::mbar_wait(&tmem_full_bar[b], ph);
            ptx::tc_fence_after();
            if (kHead) {
                // this warp's half of the 16-column chunks -> SiLU -> nc running dot products; the two halves of a row meet in
                // shared memory (double-buffered by tile parity: one named barrier per tile orders both directions)
                const int nchunks = p.n_pad >> 4, cut = (nchunks + 1) >> 1;
                const int ch_begin = half ? cut : 0, ch_end = half ? nchunks : cut;
                // packed fp32 pairs, two independent chains per output: the kernel is issue-bound (16 depthwise warps share the
                // schedulers), a scalar 64-deep FMA chain per output made this epilogue the slowest role
                float2 acc2[kDwpwMaxNc][2];
#pragma unroll
                for (int j = 0; j < kDwpwMaxNc; ++j) acc2[j][0] = acc2[j][1] = make_float2(0.f, 0.f);
                uint32_t va[16], vb[16];
                ptx::tmem_ld16(t_addr + (uint32_t)(ch_begin << 4), va);
#pragma unroll 1
                for (int ch = ch_begin; ch < ch_end; ch += 2) {
#pragma unroll
                    for (int hlf = 0; hlf < 2; ++hlf) {
                        const int chunk = ch + hlf;
                        if (chunk >= ch_end) break;
                        uint32_t(&v)[16] = hlf ? vb : va;
                        uint32_t(&vn)[16] = hlf ? va : vb;
                        ptx::tmem_ld_wait();
                        if (chunk + 1 < ch_end) ptx::tmem_ld16(t_addr + (uint32_t)((chunk + 1) << 4), vn);
                        float2 f2[8];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 hb = *reinterpret_cast<const float4*>(bias_s + (chunk << 4) + 4 * q);
                            f2[2 * q] = silu2_half(make_float2(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1])), make_float2(hb.x, hb.y));
                            f2[2 * q + 1] = silu2_half(make_float2(__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3])), make_float2(hb.z, hb.w));
                        }
#pragma unroll
                        for (int j = 0; j < kDwpwMaxNc; ++j)
                            if (j < p.head_nc) {
                                const float4* wv = reinterpret_cast<const float4*>(head_ws + j * p.n_pad + (chunk << 4));
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const float4 w4 = wv[q];
                                    acc2[j][0] = ffma2(f2[2 * q], make_float2(w4.x, w4.y), acc2[j][0]);
                                    acc2[j][1] = ffma2(f2[2 * q + 1], make_float2(w4.z, w4.w), acc2[j][1]);
                                }
                            }
                    }
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[b]);
                float acc[kDwpwMaxNc];
#pragma unroll
                for (int j = 0; j < kDwpwMaxNc; ++j) acc[j] = (acc2[j][0].x + acc2[j][0].y) + (acc2[j][1].x + acc2[j][1].y);
                float* part = head_part + (b * 128 + m) * kDwpwMaxNc;
                if (half) *reinterpret_cast<float4*>(part) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                ptx::named_bar_sync(3, 256);
                if (!half && row_ok) {
                    float* yo = p.head_y + pix * p.head_pixstride;
#pragma unroll
                    for (int j = 0; j < kDwpwMaxNc; ++j)
                        if (j < p.head_nc) yo[j] = acc[j] + part[j] + p.head_b[j];
                }
                continue;
            }
            st.c0 = 0;
            st.c1 = (int)tw_i * kDwpwTW; st.c2 = (int)th_i * kDwpwTH; st.c3 = (int)n;
            epi_tile<true, false, false>(t_addr, ec, bias_s, eo, pix, row_ok, half, lane, st, EpiResSmem{nullptr, 1, 0, 0, 0});
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[b]);
        }
        if (st.enabled && st.issuer) ptx::bulk_wait_read0();
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, p.tmem_cols);
}

int dwpw_launch(const specyolo_dwpw_t* a, cudaStream_t stream) {
    SY_CHECK(a->C == 64 || a->C == 128, SPECYOLO_ERR_UNSUPPORTED, "dwpw: C must be 64 or 128 (got %d)", a->C);
    SY_CHECK(a->n_pad % 16 == 0 && a->n_pad >= a->Cout && a->n_pad <= 256 && a->n_pad >= 64, SPECYOLO_ERR_UNSUPPORTED,
             "dwpw: n_pad must be a multiple of 16 in [64, 256]");
    SY_CHECK(a->x_pixstride % 8 == 0 && !(reinterpret_cast<uintptr_t>(a->x) & 15) && !(reinterpret_cast<uintptr_t>(a->pw_packed) & 15),
             SPECYOLO_ERR_INVALID, "dwpw: x / weights must be 16-byte aligned");
    EncodeTiledFn encode = get_encode_fn();
    SY_CHECK(encode != nullptr, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    DwpwParams p{};
    p.B = a->B; p.H = a->H; p.W = a->W; p.C = a->C; p.chunks = a->C / 64;
    p.tiles_w = ceil_div(a->W, kDwpwTW);
    p.tiles_h = ceil_div(a->H, kDwpwTH);
    const long spatial = (long)a->B * p.tiles_w * p.tiles_h;
    SY_CHECK(spatial > 0 && fastdiv_ok((uint64_t)spatial, (uint32_t)(p.tiles_w * p.tiles_h)), SPECYOLO_ERR_INVALID, "dwpw: bad tile count");
    p.spatial_tiles = (int)spatial;
    p.d_img = make_fastdiv((uint32_t)(p.tiles_w * p.tiles_h));
    p.d_tw = make_fastdiv((uint32_t)p.tiles_w);
    p.n_pad = a->n_pad; p.cout = a->Cout;
    p.dw_w = a->dw_w; p.dw_b = a->dw_b; p.pw_b = a->pw_bias;
    p.y = a->y; p.y_pixstride = a->y_pixstride;
    const bool head = a->head_w != nullptr;
    if (head) {
        SY_CHECK(a->head_nc >= 1 && a->head_nc <= kDwpwMaxNc && a->head_b && a->head_y && a->head_pixstride >= a->head_nc,
                 SPECYOLO_ERR_UNSUPPORTED, "dwpw: fused head takes 1..%d outputs", kDwpwMaxNc);
        p.head_w = a->head_w; p.head_b = a->head_b; p.head_y = a->head_y; p.head_nc = a->head_nc; p.head_pixstride = a->head_pixstride;
    } else {
        SY_CHECK(a->y != nullptr, SPECYOLO_ERR_INVALID, "dwpw: null output");
    }
    p.w_bytes = (uint32_t)p.chunks * (uint32_t)a->n_pad * 128u;
    p.x_stage_bytes = (uint32_t)p.chunks * kDwpwBoxStride;
    p.a_buf_bytes = (uint32_t)p.chunks * kDwpwAChunk;
    p.x_off = (p.w_bytes + 1023u) & ~1023u;
    p.a_off = p.x_off + 2u * p.x_stage_bytes;
    p.st_off = p.a_off + 2u * p.a_buf_bytes;
    int store_bw = epi_stage_box_cols(a->n_pad, 2);
    if (head || (reinterpret_cast<uintptr_t>(a->y) & 15) || ((size_t)a->y_pixstride * 2) % 16) store_bw = 0;
    uint32_t stage_out = epi_stage_bytes(a->n_pad, store_bw, 2);
    if (1024 + p.st_off + stage_out > (uint32_t)kDwpwMaxDynSmem) { store_bw = 0; stage_out = 0; }
    const size_t smem_bytes = 1024 + (size_t)p.st_off + stage_out;
    SY_CHECK(smem_bytes <= (size_t)kDwpwMaxDynSmem - (head ? 10 * 1024 : 0), SPECYOLO_ERR_UNSUPPORTED,
             "dwpw: shared memory budget exceeded");
    p.store_bw = store_bw;
    p.pair_stores = 1;
    p.store_row_bytes = (uint32_t)(store_bw * 2);
    p.store_swz_mask = p.store_row_bytes == 128 ? 7u : (p.store_row_bytes == 64 ? 3u : 1u);
    uint32_t cols = 32;
    while (cols < 2u * (uint32_t)a->n_pad) cols <<= 1;
    p.tmem_cols = cols;

    CUtensorMap map_x, map_w, map_y;
    {
        const cuuint64_t pix_b = (cuuint64_t)a->x_pixstride * 2;
        cuuint64_t dims[4] = {(cuuint64_t)a->C, (cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->B};
        cuuint64_t strides[3] = {pix_b, pix_b * a->W, pix_b * a->W * a->H};
        cuuint32_t box[4] = {64, kDwpwHW, kDwpwHH, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a->x), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SY_CHECK(r == CUDA_SUCCESS, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled(dwpw X) failed (%d)", (int)r);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)a->C, (cuuint64_t)a->n_pad};
        cuuint64_t strides[1] = {(cuuint64_t)a->C * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)a->n_pad};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a->pw_packed), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SY_CHECK(r == CUDA_SUCCESS, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled(dwpw W) failed (%d)", (int)r);
    }
    map_y = map_w;
    if (store_bw) {
        const cuuint64_t pix_b = (cuuint64_t)a->y_pixstride * 2;
        cuuint64_t dims[4] = {(cuuint64_t)a->Cout, (cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->B};
        cuuint64_t strides[3] = {pix_b, pix_b * a->W, pix_b * a->W * a->H};
        cuuint32_t box[4] = {(cuuint32_t)store_bw, kDwpwTW, kDwpwTH, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&map_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a->y, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for((int)p.store_row_bytes), CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SY_CHECK(r == CUDA_SUCCESS, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled(dwpw Y) failed (%d)", (int)r);
    }
    // per-device attribute: set on every launch (host-side table lookup)
    SY_CUDA(cudaFuncSetAttribute(dwpw_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwpwMaxDynSmem));
    SY_CUDA(cudaFuncSetAttribute(dwpw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwpwMaxDynSmem - 10 * 1024));   // 8 KB more static
    const long resident = sm_count();
    const unsigned grid = (unsigned)(spatial < resident ? spatial : resident);
    if (head)
        SY_CUDA(launch_pdl(dwpw_kernel<true>, dim3(grid), dim3(kDwpwThreads), smem_bytes, stream, map_x, map_w, map_y, p));
    else
        SY_CUDA(launch_pdl(dwpw_kernel<false>, dim3(grid), dim3(kDwpwThreads), smem_bytes, stream, map_x, map_w, map_y, p));
    SY_LAUNCH_CHECK();
    count_launch();
    return SPECYOLO_OK;
}

}  // namespace specyolo
