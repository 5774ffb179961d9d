// tcgen05 / TMEM implicit-GEMM convolution for sm_100a.
//
//   y[n,oh,ow,co] = act( sum_{ky,kx,ci} x[n, oh*s+ky*d-p, ow*s+kx*d-p, ci] * W[co,ky,kx,ci] + b[co] ) (+ res)
//
// GEMM view: M = output pixels (128 per CTA, a TW x TH x TN box of the NHWC output), N = output
// channels of one group (n_tile <= 256 per CTA), K = taps * Cin_g walked in (tap, 16|32|64-channel) chunks.
//   * A operand: one TMA 4-D box {kc, TW, TH, TN} of the NHWC input per (tap, channel chunk).  The tap
//     offset is a coordinate shift, the conv stride is the tensor map's elementStride, padding is TMA
//     out-of-bounds zero fill.  The box lands in shared memory as 128 K-major rows with the
//     32/64/128-byte swizzle that the UMMA shared-memory descriptor names.
//   * B operand: TMA 2-D box {kc, n_tile} of the packed weights [groups*n_pad][taps*Cin_g].
//   * D: fp32 accumulator in TMEM (n_tile columns x 128 lanes), issued by one thread with
//     tcgen05.mma.cta_group::1.kind::f16, M=128, N=n_tile, K=16 per instruction.
//   * Warp roles (320 threads): warp0 = TMA producer, warp1 = TMEM owner + MMA issuer, warps2-9 = epilogue
//     (two warps per TMEM lane quadrant; tcgen05.ld -> +bias -> SiLU -> +residual -> bf16/fp32 NHWC store at a
//     channel offset, so concat buffers are written in place and no torch.cat copy exists).
//   * Persistent: a CTA walks tiles blockIdx.x + i*gridDim.x; the accumulator is double-buffered in TMEM so the
//     epilogue of tile i overlaps the loads / MMAs of tile i+1.
//
// Replaces Conv.forward_fuse (ultralytics/nn/modules/conv.py:81-83) = cuDNN conv + bias + SiLU.
#include "common.h"
#include "ptx.cuh"
#include "tma_host.h"
#include "epilogue.cuh"

#include <cuda.h>
#include <mutex>

namespace specyolo {

struct IgemmParams {
    int TW, TH, TN;
    int tiles_w, tiles_h, tiles_n;
    int n_tiles, groups;
    FastDiv d_ntiles, d_groups, d_tw, d_th;   // tile index -> (nt, g, tw_i, th_i, tn_i)
    FastDiv d_TW, d_TWTH;                      // accumulator row -> (tw, th, tn)
    int total_tiles;
    int Ho, Wo, B;
    int kh, kw, stride, pad, dil;
    int kc, cin_chunks, cin_g;
    int n_tile, n_pad, cout_g;
    int cps, stages;
    uint32_t a_chunk_bytes, b_chunk_bytes;
    uint32_t a_tx_bytes, b_tx_bytes;
    uint32_t tmem_cols;
    const float* bias;
    void* y;
    int y_pixstride, y_fp32;
    const __nv_bfloat16* residual;
    int r_pixstride;
    int act;
    // TMA-store epilogue (store_bw == 0: direct stores)
    int store_bw, pair_stores;
    uint32_t store_row_bytes, store_swz_mask, ring_bytes;
};

static constexpr int kThreads = kConvThreads;
static constexpr int kMaxStages = 8;
static constexpr int kMaxDynSmem = 200 * 1024;  // opt-in limit is 227 KB minus the static barriers
static constexpr int kMaxBias = 1024;           // groups * n_pad floats staged in shared memory

struct TileCoord {
    int w0, h0, n0, nt, g;
};
__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& p, int tile) {
    // N tile fastest, then group, then the spatial tile: neighbouring CTAs share the same A box in L2
    TileCoord t;
    uint32_t r, nt, g, tw_i, th_i;
    fdivmod((uint32_t)tile, p.d_ntiles, r, nt);
    fdivmod(r, p.d_groups, r, g);
    fdivmod(r, p.d_tw, r, tw_i);
    fdivmod(r, p.d_th, r, th_i);
    t.nt = (int)nt;
    t.g = (int)g;
    t.w0 = (int)tw_i * p.TW;
    t.h0 = (int)th_i * p.TH;
    t.n0 = (int)r * p.TN;
    return t;
}

// Persistent, warp-specialised: every CTA walks tiles blockIdx.x, +gridDim.x, ...; the TMA ring keeps
// streaming across tile boundaries and the accumulator is double-buffered in TMEM, so the epilogue of tile
// i overlaps the loads and MMAs of tile i+1.
template <bool kSilu, bool kRes, bool kFp32>
__global__ void __launch_bounds__(kThreads, 2)     // two CTAs per SM must fit the register file (<= 102 registers)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ CUtensorMap map_y, const __grid_constant__ IgemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float bias_s[kMaxBias];

    // warp index made warp-uniform for the compiler: the role loops below then run on the uniform datapath and
    // feed UTMALDG / UTCHMMA (which take uniform registers) without a per-instruction R2UR round trip
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    ptx::grid_dep_launch();     // PDL: the next kernel may be scheduled into SM slots as this grid drains

    // 1024-byte aligned operand ring (swizzle-128B atoms repeat every 1024 B)
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t ring_off = ((raw_addr + 1023u) & ~1023u) - raw_addr;
    uint8_t* ring = smem_raw + ring_off;
    uint8_t* a_ring = ring;
    uint8_t* b_ring = ring + (size_t)p.stages * p.cps * p.a_chunk_bytes;

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&map_a);
        ptx::prefetch_tmap(&map_b);
        if (p.store_bw) ptx::prefetch_tmap(&map_y);
        for (int s = 0; s < p.stages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(&tmem_full_bar[b], 1);
            ptx::mbar_init(&tmem_empty_bar[b], kEpiWarps);   // one arrival per epilogue warp
        }
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc(&tmem_base_smem, p.tmem_cols);
    for (int i = threadIdx.x; i < p.groups * p.n_pad; i += kThreads) bias_s[i] = (kSilu ? 0.5f : 1.0f) * p.bias[i];
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    // PDL: everything above (barriers, TMEM, bias = constants) overlapped the previous kernel's tail; activations
    // are read and output buffers written only from here on
    ptx::grid_dep_wait();

    const int taps = p.kh * p.kw;
    const int total_chunks = taps * p.cin_chunks;
    const int steps = (total_chunks + p.cps - 1) / p.cps;

    if (warp == 0) {
        // ===================== TMA producer =====================
        // All 32 lanes walk the loops converged (uniform control flow, no runtime divisions: the issue stream is
        // one long latency chain); the elected lane issues the TMA instructions.
        {
            const bool leader = ptx::elect_one();
            int stage = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
                const TileCoord tc = decode_tile(p, tile);
                const int ca = tc.g * p.cin_g;
                const int cw = tc.w0 * p.stride - p.pad, chh = tc.h0 * p.stride - p.pad;
                const int bn = tc.g * p.n_pad + tc.nt * p.n_tile;
                int ky = 0, kx = 0, cc = 0, kb = 0;       // running (tap row, tap column, channel chunk, weight K offset)
                int left = total_chunks;
                for (int step = 0; step < steps; ++step) {
                    ptx::mbar_wait(&empty_bar[stage], ph ^ 1u);
                    const int nch = min(p.cps, left);
                    left -= nch;
                    if (leader) ptx::mbar_expect_tx(&full_bar[stage], nch * (p.a_tx_bytes + p.b_tx_bytes));
                    uint8_t* a_dst = a_ring + (size_t)(stage * p.cps) * p.a_chunk_bytes;
                    uint8_t* b_dst = b_ring + (size_t)(stage * p.cps) * p.b_chunk_bytes;
                    for (int j = 0; j < nch; ++j) {
                        if (leader) {
                            ptx::tma_load_4d(a_dst, &map_a, &full_bar[stage], ca + cc * p.kc, cw + kx * p.dil,
                                             chh + ky * p.dil, tc.n0);
                            ptx::tma_load_2d(b_dst, &map_b, &full_bar[stage], kb, bn);
                        }
                        a_dst += p.a_chunk_bytes;
                        b_dst += p.b_chunk_bytes;
                        kb += p.kc;
                        if (++cc == p.cin_chunks) {
                            cc = 0;
                            if (++kx == p.kw) { kx = 0; ++ky; }
                        }
                    }
                    if (++stage == p.stages) { stage = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        {
            const bool leader = ptx::elect_one();
            const uint32_t idesc = ptx::umma_idesc_bf16(128, p.n_tile);
            const uint32_t row_bytes = p.kc * 2;
            const int kk = p.kc / 16;
            const uint32_t desc_hi = (uint32_t)(ptx::umma_smem_desc(0, row_bytes) >> 32);   // all but the start address
            const uint32_t a_ring_addr = ptx::smem_u32(a_ring), b_ring_addr = ptx::smem_u32(b_ring);
            int stage = 0;
            uint32_t ph = 0, tl = 0;
            for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tl) {
                const uint32_t buf = tl & 1u;
                const uint32_t bph = (tl >> 1) & 1u;
                ptx::mbar_wait(&tmem_empty_bar[buf], bph ^ 1u);   // epilogue has drained this accumulator
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * (uint32_t)p.n_tile;
                uint32_t accumulate = 0;
                int left = total_chunks;
                for (int step = 0; step < steps; ++step) {
                    ptx::mbar_wait(&full_bar[stage], ph);
                    ptx::tc_fence_after();
                    const int nch = min(p.cps, left);
                    left -= nch;
                    // descriptor start-address fields (16-byte units) of the first chunk of this stage
                    uint32_t a16 = (a_ring_addr + (uint32_t)(stage * p.cps) * p.a_chunk_bytes) >> 4;
                    uint32_t b16 = (b_ring_addr + (uint32_t)(stage * p.cps) * p.b_chunk_bytes) >> 4;
                    for (int j = 0; j < nch; ++j) {
                        for (int k = 0; k < kk; ++k) {
                            if (leader) ptx::umma_bf16_lohi(d_tmem, a16 + 2u * k, desc_hi, b16 + 2u * k, desc_hi, idesc, accumulate);
                            accumulate = 1;
                        }
                        a16 += p.a_chunk_bytes >> 4;
                        b16 += p.b_chunk_bytes >> 4;
                    }
                    if (leader) ptx::umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                    if (++stage == p.stages) { stage = 0; ph ^= 1u; }
                }
                if (leader) ptx::umma_commit(&tmem_full_bar[buf]);    // accumulator complete
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (warps 2..9, see epilogue.cuh) =====================
        const int quad = warp & 3;            // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;     // which half of the tile's column chunks
        const int m = quad * 32 + lane;       // accumulator row = pixel inside the tile
        const int npix = p.TW * p.TH * p.TN;
        uint32_t tn_u, th_u, tw_u, rem_u;
        fdivmod((uint32_t)m, p.d_TWTH, tn_u, rem_u);
        fdivmod(rem_u, p.d_TW, th_u, tw_u);
        const int tw = (int)tw_u, th = (int)th_u, tn = (int)tn_u;
        EpiOut eo{p.y, p.y_pixstride, p.residual, p.r_pixstride, p.pair_stores != 0};
        EpiStage st = epi_make_stage(ring + p.ring_bytes, &map_y, p.n_tile, p.store_bw, p.store_row_bytes, p.store_swz_mask,
                                     warp, half, lane, m);
        EpiCols ec;
        ec.ncols = p.n_tile;
        ec.n_pad = 1 << 20;                   // the tile is a slice of ONE group: column -> (0, within0 + column)
        ec.d_npad = FastDiv{1ull << 20, 1u << 20};
        ec.cout_g = p.cout_g;
        uint32_t tl = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++tl) {
            const TileCoord tc = decode_tile(p, tile);
            const uint32_t buf = tl & 1u;
            const uint32_t bph = (tl >> 1) & 1u;
            const int ow = tc.w0 + tw, oh = tc.h0 + th, on = tc.n0 + tn;
            const bool row_ok = (m < npix) && (ow < p.Wo) && (oh < p.Ho) && (on < p.B);
            const size_t pix = ((size_t)on * p.Ho + oh) * p.Wo + ow;
            ec.within0 = tc.nt * p.n_tile;            // channel offset inside the group
            ec.gch0 = tc.g * p.cout_g;                // group offset inside the output window
            const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * (uint32_t)p.n_tile;

            if (kRes && tile + (int)gridDim.x < p.total_tiles) {            // next tile's residual -> L2
                const TileCoord tn2 = decode_tile(p, tile + (int)gridDim.x);
                const int ow2 = tn2.w0 + tw, oh2 = tn2.h0 + th, on2 = tn2.n0 + tn;
                EpiCols e2 = ec;
                e2.within0 = tn2.nt * p.n_tile;
                e2.gch0 = tn2.g * p.cout_g;
                epi_prefetch_residual_l2<kRes>(e2, eo, ((size_t)on2 * p.Ho + oh2) * p.Wo + ow2,
                                               (m < npix) && (ow2 < p.Wo) && (oh2 < p.Ho) && (on2 < p.B), half);
            }
            ptx::mbar_wait(&tmem_full_bar[buf], bph);
            ptx::tc_fence_after();
            st.c0 = ec.gch0 + ec.within0;
            st.c1 = tc.w0; st.c2 = tc.h0; st.c3 = tc.n0;
            epi_tile<kSilu, kRes, kFp32>(t_addr, ec, bias_s + tc.g * p.n_pad + ec.within0, eo, pix, row_ok, half, lane, st, EpiResSmem{nullptr, 1, 0, 0, 0});
            // all TMEM reads of this accumulator are complete (wait::ld inside): hand it back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[buf]);
        }
        if (st.enabled && st.issuer) ptx::bulk_wait_read0();   // staging must outlive the last store's read
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// Pick the output tile (TW,TH,TN), TW*TH*TN <= 128, that needs the fewest CTAs.
static void choose_tile(int B, int Ho, int Wo, int stride, int& TW, int& TH, int& TN) {
    long best_tiles = -1;
    int bw = 1, bh = 1, bn = 1;
    for (int tw = 1; tw <= 128 && tw <= Wo; ++tw) {
        if (tw * stride > 256) break;
        for (int th = 1; th * tw <= 128 && th <= Ho; ++th) {
            if (th * stride > 256) break;
            int tn = 128 / (tw * th);
            if (tn > B) tn = B;
            if (tn < 1) continue;
            // a tile that spans several images must cover whole rows x whole image height only if
            // it spans them contiguously; the 4-D box makes any (tw,th,tn) legal.
            const long tiles = (long)ceil_div(Wo, tw) * ceil_div(Ho, th) * ceil_div(B, tn);
            const bool better = best_tiles < 0 || tiles < best_tiles ||
                                (tiles == best_tiles && tw > bw);
            if (better) { best_tiles = tiles; bw = tw; bh = th; bn = tn; }
        }
    }
    TW = bw; TH = bh; TN = bn;
}

int conv_igemm_launch(const specyolo_conv_t* a, cudaStream_t stream) {
    EncodeTiledFn encode = get_encode_fn();
    SY_CHECK(encode != nullptr, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");

    const int groups = a->groups;
    const int cin_g = a->Cin / groups, cout_g = a->Cout / groups;
    SY_CHECK(a->Cin % groups == 0 && a->Cout % groups == 0, SPECYOLO_ERR_INVALID, "channels not divisible by groups");
    SY_CHECK(cin_g % 16 == 0, SPECYOLO_ERR_UNSUPPORTED, "igemm conv needs Cin/groups %% 16 == 0 (got %d)", cin_g);
    SY_CHECK(a->n_pad % 16 == 0 && a->n_pad >= cout_g, SPECYOLO_ERR_INVALID, "bad n_pad %d for cout_g %d", a->n_pad, cout_g);
    SY_CHECK(a->x_pixstride % 8 == 0, SPECYOLO_ERR_INVALID, "x pixel stride must be a multiple of 8 elements");
    SY_CHECK((reinterpret_cast<uintptr_t>(a->x) & 15) == 0, SPECYOLO_ERR_INVALID, "x must be 16-byte aligned");
    SY_CHECK((reinterpret_cast<uintptr_t>(a->w_packed) & 15) == 0, SPECYOLO_ERR_INVALID, "w_packed must be 16-byte aligned");
    SY_CHECK(a->stride >= 1 && a->stride <= 8, SPECYOLO_ERR_INVALID, "stride out of range");

    IgemmParams p{};
    const int kc = (cin_g % 64 == 0) ? 64 : (cin_g % 32 == 0 ? 32 : 16);
    p.kc = kc;
    p.cin_g = cin_g;
    p.cin_chunks = cin_g / kc;
    p.kh = a->kh; p.kw = a->kw; p.stride = a->stride; p.pad = a->pad; p.dil = a->dil;
    p.n_pad = a->n_pad;
    p.cout_g = cout_g;
    // N tile: whole padded group if it fits one UMMA, else the largest multiple-of-16 divisor <= 256
    int n_tile = a->n_pad;
    if (n_tile > 256) {
        n_tile = 256;
        while (a->n_pad % n_tile != 0) n_tile -= 16;
    }
    p.n_tile = n_tile;

    // geometry: 1x1/s1/p0 convs run "flat" over all B*H*W pixels
    const bool flat = (a->kh == 1 && a->kw == 1 && a->stride == 1 && a->pad == 0);
    int gB, gH, gW, gHo, gWo;
    if (flat) {
        gB = 1; gH = 1; gW = a->B * a->H * a->W; gHo = 1; gWo = gW;
        p.TW = 128; p.TH = 1; p.TN = 1;
    } else {
        gB = a->B; gH = a->H; gW = a->W; gHo = a->Ho; gWo = a->Wo;
        choose_tile(gB, gHo, gWo, a->stride, p.TW, p.TH, p.TN);
    }
    p.B = gB; p.Ho = gHo; p.Wo = gWo;
    p.tiles_w = ceil_div(gWo, p.TW);
    p.tiles_h = ceil_div(gHo, p.TH);
    p.tiles_n = ceil_div(gB, p.TN);
    p.n_tiles = a->n_pad / n_tile;
    p.groups = groups;
    const long total_tiles = (long)p.tiles_w * p.tiles_h * p.tiles_n * p.n_tiles * groups;
    SY_CHECK(total_tiles > 0 && total_tiles < (1L << 30), SPECYOLO_ERR_INVALID, "bad tile count");
    p.total_tiles = (int)total_tiles;
    p.d_ntiles = make_fastdiv((uint32_t)p.n_tiles);
    p.d_groups = make_fastdiv((uint32_t)groups);
    p.d_tw = make_fastdiv((uint32_t)p.tiles_w);
    p.d_th = make_fastdiv((uint32_t)p.tiles_h);
    p.d_TW = make_fastdiv((uint32_t)p.TW);
    p.d_TWTH = make_fastdiv((uint32_t)(p.TW * p.TH));
    {
        uint32_t dmax = 128;
        for (int d : {p.n_tiles, groups, p.tiles_w, p.tiles_h}) dmax = (uint32_t)d > dmax ? (uint32_t)d : dmax;
        SY_CHECK(fastdiv_ok((uint64_t)total_tiles, dmax), SPECYOLO_ERR_UNSUPPORTED, "too many tiles for the fast tile decode");
    }
    SY_CHECK(groups * a->n_pad <= kMaxBias, SPECYOLO_ERR_UNSUPPORTED, "too many output channels (%d)", groups * a->n_pad);

    const int row_bytes = kc * 2;
    p.cps = 64 / kc;
    p.a_chunk_bytes = 128u * row_bytes;
    p.b_chunk_bytes = ((uint32_t)(n_tile * row_bytes) + 1023u) & ~1023u;
    p.a_tx_bytes = (uint32_t)(p.TW * p.TH * p.TN) * row_bytes;
    p.b_tx_bytes = (uint32_t)n_tile * row_bytes;
    const uint32_t stage_bytes = p.cps * (p.a_chunk_bytes + p.b_chunk_bytes);
    // two persistent CTAs per SM when the double-buffered accumulators of both fit TMEM (2*2*n_tile <= 512):
    // ring <= ~100 KB each; otherwise one CTA per SM with a deeper ring
    uint32_t cols = 32;
    while (cols < 2u * (uint32_t)n_tile) cols <<= 1;
    p.tmem_cols = cols;
    const int occ = (cols <= 256) ? 2 : 1;
    // TMA-store epilogue: needs a 16-byte aligned output window whose pixel stride is a multiple of 16 bytes, and
    // accumulator columns that map to consecutive channels (no per-group padding between groups)
    const int es = a->y_fp32 ? 4 : 2;
    int store_bw = epi_stage_box_cols(n_tile, es);
    if ((reinterpret_cast<uintptr_t>(a->y) & 15) || ((size_t)a->y_pixstride * es) % 16 || (groups > 1 && cout_g != a->n_pad) ||
        env_flag("SPECYOLO_NO_TMA_STORE"))
        store_bw = 0;
    // long-K layers are tensor-bound: the epilogue is hidden anyway and the staging buffer would cost a ring stage
    // (measured: 128->128 k3 s2 133 us direct vs 150 us staged)
    if ((long)a->kh * a->kw * cin_g >= 1024) store_bw = 0;
    const uint32_t stage_out_bytes = epi_stage_bytes(n_tile, store_bw, es);
    const uint32_t ring_budget = ((occ == 2) ? 108u * 1024u : 200u * 1024u) - 1024u - stage_out_bytes;
    int stages = (int)(ring_budget / stage_bytes);
    if (stages < 2) stages = 2;
    if (stages > kMaxStages) stages = kMaxStages;
    p.stages = stages;
    p.store_bw = store_bw;
    p.pair_stores = (n_tile >= 64 && !env_flag("SPECYOLO_NO_PAIR")) ? 1 : 0;   // thin tiles: the shuffles cost more than they save
    p.store_row_bytes = (uint32_t)(store_bw * es);
    p.store_swz_mask = p.store_row_bytes == 128 ? 7u : (p.store_row_bytes == 64 ? 3u : 1u);
    p.ring_bytes = (uint32_t)stages * stage_bytes;          // multiple of 1024: the staging buffers stay aligned
    const size_t smem_bytes = (size_t)stages * stage_bytes + 1024 + stage_out_bytes;
    SY_CHECK(smem_bytes <= (size_t)kMaxDynSmem, SPECYOLO_ERR_INVALID, "smem budget exceeded");

    p.bias = a->bias;
    p.y = a->y; p.y_pixstride = a->y_pixstride; p.y_fp32 = a->y_fp32;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual);
    p.r_pixstride = a->r_pixstride;
    p.act = a->act;

    // ---- tensor maps ----
    CUtensorMap map_a, map_b;
    {
        const cuuint64_t pix_b = (cuuint64_t)a->x_pixstride * 2;
        cuuint64_t dims[4] = {(cuuint64_t)a->Cin, (cuuint64_t)gW, (cuuint64_t)gH, (cuuint64_t)gB};
        cuuint64_t strides[3] = {pix_b, pix_b * gW, pix_b * gW * gH};
        cuuint32_t box[4] = {(cuuint32_t)kc, (cuuint32_t)(p.TW * a->stride),
                             (cuuint32_t)(p.TH * a->stride), (cuuint32_t)p.TN};
        cuuint32_t estr[4] = {1, (cuuint32_t)a->stride, (cuuint32_t)a->stride, 1};
        CUresult r = encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a->x), dims,
                            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(row_bytes),
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SY_CHECK(r == CUDA_SUCCESS, SPECYOLO_ERR_CUDA,
                 "cuTensorMapEncodeTiled(A) failed (%d): dims %llu,%llu,%llu,%llu box %u,%u,%u,%u",
                 (int)r, (unsigned long long)dims[0], (unsigned long long)dims[1],
                 (unsigned long long)dims[2], (unsigned long long)dims[3], box[0], box[1], box[2], box[3]);
    }
    {
        const cuuint64_t ktot = (cuuint64_t)a->kh * a->kw * cin_g;
        cuuint64_t dims[2] = {ktot, (cuuint64_t)groups * a->n_pad};
        cuuint64_t strides[1] = {ktot * 2};
        cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)n_tile};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a->w_packed),
                            dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            swizzle_for(row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SY_CHECK(r == CUDA_SUCCESS, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed (%d)", (int)r);
    }

    CUtensorMap map_y = map_b;      // placeholder when the direct-store epilogue is used
    if (store_bw) {
        const cuuint64_t pix_b = (cuuint64_t)a->y_pixstride * es;
        cuuint64_t dims[4] = {(cuuint64_t)a->Cout, (cuuint64_t)gWo, (cuuint64_t)gHo, (cuuint64_t)gB};
        cuuint64_t strides[3] = {pix_b, pix_b * gWo, pix_b * gWo * gHo};
        cuuint32_t box[4] = {(cuuint32_t)store_bw, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)p.TN};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&map_y, a->y_fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a->y,
                            dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for((int)p.store_row_bytes),
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SY_CHECK(r == CUDA_SUCCESS, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled(Y) failed (%d)", (int)r);
    }
    typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const IgemmParams);
    static const KernelFn kernels[8] = {
        conv_igemm_kernel<false, false, false>, conv_igemm_kernel<false, false, true>,
        conv_igemm_kernel<false, true, false>,  conv_igemm_kernel<false, true, true>,
        conv_igemm_kernel<true, false, false>,  conv_igemm_kernel<true, false, true>,
        conv_igemm_kernel<true, true, false>,   conv_igemm_kernel<true, true, true>};
    static std::once_flag attr_once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(attr_once, [] {
        for (KernelFn k : kernels) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
            if (e != cudaSuccess) attr_err = e;
        }
    });
    SY_CHECK(attr_err == cudaSuccess, SPECYOLO_ERR_CUDA, "cudaFuncSetAttribute failed: %s",
             cudaGetErrorString(attr_err));

    const long resident = (long)sm_count() * occ;
    const unsigned grid = (unsigned)(total_tiles < resident ? total_tiles : resident);
    const KernelFn kernel = kernels[(a->act == SPECYOLO_ACT_SILU ? 4 : 0) + (a->residual ? 2 : 0) + (a->y_fp32 ? 1 : 0)];
    SY_CUDA(launch_pdl(kernel, dim3(grid), dim3(kThreads), smem_bytes, stream, map_a, map_b, map_y, p));
    SY_LAUNCH_CHECK();
    count_launch();
    return SPECYOLO_OK;
}

}  // namespace specyolo
