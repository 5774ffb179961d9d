// Fused Bottleneck for sm_100a:  y = [x +] SiLU(W2 (*) SiLU(W1 (*) x + b1) + b2),  both convs 3x3 / s1 / p1, BN folded
// (ultralytics/nn/modules/block.py:713-726: Bottleneck.forward = x + cv2(cv1(x)); the inner blocks of every C3k2,
// block.py:1659-1671).
//
// Layer by layer the thin bottlenecks of the 160^2 / 80^2 / 40^2 levels (32->16->32, 64->32->64 channels) are bound by
// per-tile fixed costs, not by the tensor pipe or HBM: 8-column-thin GEMMs (N = 16 / 32) with one TMA / MMA / epilogue
// round trip per 128 pixels, the 16 / 32-channel intermediate written to HBM and read back through nine taps, and the
// shortcut read again by the second conv's epilogue.  Here the intermediate never leaves the SM.  Per tile of 14 x 14
// output pixels (persistent CTA, one per SM, 18 warps):
//   warp 0      TMA: ONE 18 x 18-pixel input box (the tile plus a 2-pixel halo; out-of-image pixels arrive as zeros =
//               conv 1's padding) per tile into a 2-3 deep ring; the resident weights of both convs, once
//   warp 1      MMA 1: conv 1 over the 16 x 16 "mid" region (tile + 1-pixel halo) = 2 M-tiles of 128 rows (8 pixels x
//               16 rows each); the A operand of tap (ky, kx) is the input box read through a UMMA descriptor whose
//               start is shifted by (ky * 18 + kx) rows and whose 8-row-group stride is 18 rows (as conv_halo.cu)
//               MMA 2: conv 2 over 2 M-tiles covering 16 x 16 output positions (the inner 14 x 14 are kept); A = the
//               mid tile in shared memory, taps = descriptor shifts of (ky * 16 + kx) rows, group stride 16 rows
//   warps 2-9   epilogue 1: conv-1 accumulators (double-buffered in TMEM) -> bias + SiLU -> bf16 -> mid tile in the
//               swizzled K-major operand layout; mid pixels outside the image are written as ZEROS (conv 2's padding).
//               The mid tile is double-buffered too.
//   warps 10-17 epilogue 2: conv-2 accumulators (double-buffered in TMEM) -> bias + SiLU (+ x from global memory, an
//               L2 hit: the tile was just loaded) -> bf16 -> [14 x 14][C] staging -> one TMA store per tile (clipped at
//               the image edges by the TMA unit)
// MMA 1 of tile t+1 is issued before MMA 2 of tile t, so the loads / conv 1 / epilogue 1 of the next tile overlap
// conv 2 / epilogue 2 / the store of the current one.
//
// Arithmetic per tile is that of the two convs on (256 + 256) / 196 of the pixels (halo recompute + the unused rim of
// the 16 x 16 output M-tiles); traffic is x once (plus the halo, from L2) and y once.
#include "common.h"
#include "ptx.cuh"
#include "tma_host.h"
#include "epilogue.cuh"

#include <cstdlib>
#include <mutex>

namespace specyolo {

static constexpr int kBpThreads = 64 + 256 + 256;           // TMA, MMA issuer, 8 + 8 epilogue warps
static constexpr int kBpIssuers = 1;          // arrivals per commit barrier
static constexpr int kBpOT = 14;             // valid output pixels per tile edge
static constexpr int kBpMW = 16;             // mid tile edge (= pitch, pixels)
static constexpr int kBpIW = 18;             // input box edge
static constexpr int kBpMidRows = 18 * 16 + 8;   // mid rows an M-tile of conv 2 may touch (rows 16, 17 + wrap: unused outputs)
static constexpr int kBpMaxStages = 4;
static constexpr int kBpMaxDynSmem = 226 * 1024;

struct BneckParams {
    int B, H, W;
    int tiles_w, tiles_h, spatial_tiles;
    FastDiv d_img, d_tw;
    int c, cm;                       // channels of x / y, of the intermediate
    int np1, np2;                    // accumulator columns of conv 1 / conv 2 (padded to 16)
    const float* b1;
    const float* b2;
    const __nv_bfloat16* x;
    int x_pixstride;
    int add;
    uint32_t a_row_bytes, m_row_bytes;            // c * 2, cm * 2
    uint32_t a_stage_bytes, a_tx_bytes, mid_bytes;
    uint32_t w1_box_bytes, w2_box_bytes, w_total_bytes;
    uint32_t off_w2, off_ring, off_mid, off_st;
    int stages;
    uint32_t tmem_cols;
    uint32_t tap_a16[9], tap_m16[9];
    uint32_t st_swz_mask, m_swz_mask;
    unsigned long long* dbg;         // SPECYOLO_BP_DBG=1: per-role wait / work cycle counters of CTA 0
};

__device__ __forceinline__ uint64_t bp_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t row_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
    d |= 1ull << 46;
    d |= layout << 61;
    return d;
}

// phase timers (debug): cycles CTA 0's role leaders spend in each wait / work phase, summed over the tiles
// (accumulated in registers — a global read-modify-write per event costs more than the phases it measures — and
//  written once when the role's loop ends)
struct BpTimers {
    unsigned long long t[16];
};
#define BP_ACC(slot, t_begin) do { if (p.dbg) tm.t[slot] += (unsigned long long)(clock64() - (t_begin)); } while (0)
__device__ __forceinline__ void bp_wait_timed(const BneckParams& p, BpTimers& tm, int slot, uint64_t* bar, uint32_t parity) {
    if (p.dbg) {
        const long long t = clock64();
        ptx::mbar_wait(bar, parity);
        tm.t[slot] += (unsigned long long)(clock64() - t);
    } else {
        ptx::mbar_wait(bar, parity);
    }
}
__device__ __forceinline__ void bp_flush(const BneckParams& p, const BpTimers& tm, int lane, int first, int last) {
    if (p.dbg && blockIdx.x == 0 && lane == 0)
        for (int i = first; i <= last; ++i) p.dbg[i] = tm.t[i];
}

template <bool kRes>
__global__ void __launch_bounds__(kBpThreads, 1)
bneck_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
                  const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_y,
                  const __grid_constant__ BneckParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kBpMaxStages], empty_bar[kBpMaxStages];
    __shared__ __align__(8) uint64_t acc1_full[2], acc1_empty[2], mid_full[2], mid_empty[2], acc2_full[2], acc2_empty[2], w_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float bias1_s[64];
    __shared__ __align__(16) float bias2_s[64];

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    ptx::grid_dep_launch();
    BpTimers tm;
#pragma unroll
    for (int i = 0; i < 16; ++i) tm.t[i] = 0;

    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* base = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* w1_s = base;                       // 9 boxes [np1 rows x c*2 B]
    uint8_t* w2_s = base + p.off_w2;            // 9 boxes [np2 rows x cm*2 B]
    uint8_t* ring = base + p.off_ring;          // input boxes: 324 rows x c*2 B
    uint8_t* mid_s = base + p.off_mid;          // 2 mid tiles: kBpMidRows rows x cm*2 B
    uint8_t* st_buf = base + p.off_st;          // 196 rows x c*2 B

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&map_x);
        ptx::prefetch_tmap(&map_w1);
        ptx::prefetch_tmap(&map_w2);
        ptx::prefetch_tmap(&map_y);
        // barriers completed by tcgen05.commit take one arrival per MMA issuer warp
        for (int s = 0; s < p.stages; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], kBpIssuers);
        }
        for (int b = 0; b < 2; ++b) {
            ptx::mbar_init(&mid_full[b], kEpiWarps);
            ptx::mbar_init(&mid_empty[b], kBpIssuers);
            ptx::mbar_init(&acc2_full[b], kBpIssuers);
            ptx::mbar_init(&acc2_empty[b], kEpiWarps);
            ptx::mbar_init(&acc1_full[b], kBpIssuers);
            ptx::mbar_init(&acc1_empty[b], kEpiWarps);
        }
        ptx::mbar_init(&w_bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc(&tmem_base_smem, p.tmem_cols);
    if ((int)threadIdx.x < p.np1) bias1_s[threadIdx.x] = 0.5f * p.b1[threadIdx.x];          // SiLU form (epilogue.cuh)
    if ((int)threadIdx.x >= 64 && (int)threadIdx.x < 64 + p.np2) bias2_s[threadIdx.x - 64] = 0.5f * p.b2[threadIdx.x - 64];
    // the mid rows no epilogue ever writes (rows 16, 17 and the wrap-around tail) feed only discarded accumulator rows;
    // zero them once anyway so that no NaN pattern ever enters the tensor pipe
    for (uint32_t i = threadIdx.x; i < 2u * p.mid_bytes / 16u; i += kBpThreads)
        reinterpret_cast<uint4*>(mid_s)[i] = make_uint4(0, 0, 0, 0);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const uint32_t acc1_col = 4u * (uint32_t)p.np2;          // [2 buffers][2 M-tiles] of conv 2 first, then the same for conv 1
    const int cta = blockIdx.x, ctas = gridDim.x;

    if (warp == 0) {
        // ===================== TMA producer =====================
        const bool leader = ptx::elect_one();
        if (leader) {
            ptx::mbar_expect_tx(&w_bar, p.w_total_bytes);
            for (int t = 0; t < 9; ++t) ptx::tma_load_2d(w1_s + (size_t)t * p.w1_box_bytes, &map_w1, &w_bar, t * p.c, 0);
            for (int t = 0; t < 9; ++t) ptx::tma_load_2d(w2_s + (size_t)t * p.w2_box_bytes, &map_w2, &w_bar, t * p.cm, 0);
        }
        ptx::grid_dep_wait();
        int stage = 0;
        uint32_t ph = 0;
        for (int tile = cta; tile < p.spatial_tiles; tile += ctas) {
            uint32_t n, r, th_i, tw_i;
            fdivmod((uint32_t)tile, p.d_img, n, r);
            fdivmod(r, p.d_tw, th_i, tw_i);
            bp_wait_timed(p, tm, 0, &empty_bar[stage], ph ^ 1u);
            if (leader) {
                ptx::mbar_expect_tx(&full_bar[stage], p.a_tx_bytes);
                ptx::tma_load_4d(ring + (size_t)stage * p.a_stage_bytes, &map_x, &full_bar[stage], 0,
                                 (int)tw_i * kBpOT - 2, (int)th_i * kBpOT - 2, (int)n);
            }
            if (++stage == p.stages) { stage = 0; ph ^= 1u; }
        }
        bp_flush(p, tm, lane, 0, 0);
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // Measured alternatives (tools/one_bneck.py, clock64 phase timers, ncu): a second issuer warp — one M-tile each,
        // or one conv each — does not help here although two warps can feed the tensor pipe concurrently
        // (tests/cuda/umma_dual_issue_probe.cu): the kernel is bound by the shared-memory DATA PIPE, which the tensor
        // operand reads of these thin-N MMAs (37 wavefronts per M=128, K=16 MMA whatever N is), the accumulator
        // read-out (tcgen05.ld) and the epilogue stores share (ncu: 2 015 + 1 374 wavefronts per tile at C = 32 against a
        // 3 700-cycle tile).
        ptx::grid_dep_wait();
        const bool leader = ptx::elect_one();
        const uint32_t idesc1 = ptx::umma_idesc_bf16(128, (uint32_t)p.np1);
        const uint32_t idesc2 = ptx::umma_idesc_bf16(128, (uint32_t)p.np2);
        const uint32_t a_hi = (uint32_t)(bp_desc(0, (uint32_t)kBpIW * p.a_row_bytes, p.a_row_bytes) >> 32);
        const uint32_t m_hi = (uint32_t)(bp_desc(0, (uint32_t)kBpMW * p.m_row_bytes, p.m_row_bytes) >> 32);
        const uint32_t w1_hi = (uint32_t)(bp_desc(0, 8u * p.a_row_bytes, p.a_row_bytes) >> 32);
        const uint32_t w2_hi = (uint32_t)(bp_desc(0, 8u * p.m_row_bytes, p.m_row_bytes) >> 32);
        const uint32_t ring16 = ptx::smem_u32(ring) >> 4, mid16 = ptx::smem_u32(mid_s) >> 4;
        const uint32_t w1_16 = ptx::smem_u32(w1_s) >> 4, w2_16 = ptx::smem_u32(w2_s) >> 4;
        const uint32_t w1_box16 = p.w1_box_bytes >> 4, w2_box16 = p.w2_box_bytes >> 4;
        const uint32_t mt_a16 = (8u * p.a_row_bytes) >> 4, mt_m16 = (8u * p.m_row_bytes) >> 4;   // second M-tile: 8 pixels right
        const int k1 = p.c >> 4, k2 = p.cm >> 4;
        ptx::mbar_wait(&w_bar, 0);
        ptx::tc_fence_after();
        int stage = 0;
        uint32_t sph = 0;
        auto mma1 = [&](uint32_t t) {
            const uint32_t b1 = t & 1u, ph1 = (t >> 1) & 1u;
            bp_wait_timed(p, tm, 1, &acc1_empty[b1], ph1 ^ 1u);
            bp_wait_timed(p, tm, 2, &full_bar[stage], sph);
            ptx::tc_fence_after();
            const uint32_t a16 = ring16 + (((uint32_t)stage * p.a_stage_bytes) >> 4);
            const long long t_i1 = p.dbg ? clock64() : 0;
            for (int mt = 0; mt < 2; ++mt) {
                const uint32_t d = tmem_base + acc1_col + (b1 * 2u + (uint32_t)mt) * (uint32_t)p.np1;
                uint32_t acc = 0;
                for (int k = 0; k < k1; ++k) {
                    const uint32_t ak = a16 + (uint32_t)mt * mt_a16 + 2u * (uint32_t)k;
                    uint32_t wk = w1_16 + 2u * (uint32_t)k;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        if (leader) ptx::umma_bf16_lohi(d, ak + p.tap_a16[tap], a_hi, wk, w1_hi, idesc1, acc);
                        acc = 1u;
                        wk += w1_box16;
                    }
                }
            }
            BP_ACC(13, t_i1);
            if (leader) {
                ptx::umma_commit(&empty_bar[stage]);
                ptx::umma_commit(&acc1_full[b1]);
            }
            if (++stage == p.stages) { stage = 0; sph ^= 1u; }
        };
        auto mma2 = [&](uint32_t t) {
            const uint32_t b = t & 1u, ph = (t >> 1) & 1u;
            bp_wait_timed(p, tm, 3, &acc2_empty[b], ph ^ 1u);
            bp_wait_timed(p, tm, 4, &mid_full[b], ph);
            ptx::tc_fence_after();
            const uint32_t m16 = mid16 + ((b * p.mid_bytes) >> 4);
            const long long t_i2 = p.dbg ? clock64() : 0;
            for (int ot = 0; ot < 2; ++ot) {
                const uint32_t d = tmem_base + (b * 2u + (uint32_t)ot) * (uint32_t)p.np2;
                uint32_t acc = 0;
                for (int k = 0; k < k2; ++k) {
                    const uint32_t ak = m16 + (uint32_t)ot * mt_m16 + 2u * (uint32_t)k;
                    uint32_t wk = w2_16 + 2u * (uint32_t)k;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        if (leader) ptx::umma_bf16_lohi(d, ak + p.tap_m16[tap], m_hi, wk, w2_hi, idesc2, acc);
                        acc = 1u;
                        wk += w2_box16;
                    }
                }
            }
            BP_ACC(14, t_i2);
            if (leader) {
                ptx::umma_commit(&mid_empty[b]);
                ptx::umma_commit(&acc2_full[b]);
            }
        };
        uint32_t tl = 0;
        const long long t_loop = p.dbg ? clock64() : 0;
        if (cta < p.spatial_tiles) mma1(0);
        for (int tile = cta; tile < p.spatial_tiles; tile += ctas, ++tl) {
            if (tile + ctas < p.spatial_tiles) mma1(tl + 1);
            mma2(tl);
        }
        BP_ACC(5, t_loop);
        bp_flush(p, tm, lane, 1, 5);
        bp_flush(p, tm, lane, 13, 14);
        if (p.dbg && blockIdx.x == 0 && lane == 0) p.dbg[15] = tl;
        __syncwarp();
    } else {
        ptx::grid_dep_wait();
        const int quad = warp & 3;
        const int half = ((warp - 2) >> 2) & 1;         // which of the two warps of a lane quadrant inside its group
        const int m = quad * 32 + lane;
        const int gy = m >> 3, gx = m & 7;               // position inside an M-tile: row group (0..15), pixel (0..7)
        const uint32_t t_quad = tmem_base + ((uint32_t)(quad * 32) << 16);
        if (warp < 10) {
            // ---- epilogue 1: conv-1 accumulators -> bias + SiLU -> bf16 -> mid tile (zeros outside the image) ----
            // tasks of this warp: cm = 16: M-tile `half`, its only chunk; cm = 32: chunk `half` of both M-tiles
            const int chunks = p.np1 >> 4;
            const int ntask = chunks == 1 ? 1 : 2;
            const int cc = chunks == 1 ? 0 : half;
            float2 hb[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) hb[i] = make_float2(bias1_s[cc * 16 + 2 * i], bias1_s[cc * 16 + 2 * i + 1]);
            uint32_t t_off[2], t_col[2];
            int t_mx[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int mt = chunks == 1 ? half : k;
                const int row = gy * kBpMW + 8 * mt + gx;                  // mid pixel (gy, 8 mt + gx)
                uint32_t off = (uint32_t)row * p.m_row_bytes + (uint32_t)cc * 32u;
                t_off[k] = off;
                t_col[k] = acc1_col + (uint32_t)(mt * p.np1 + cc * 16);       // (+ buffer * 2 * np1 per tile)
                t_mx[k] = 8 * mt + gx;
            }
            uint32_t tl = 0;
            for (int tile = cta; tile < p.spatial_tiles; tile += ctas, ++tl) {
                uint32_t n, r, th_i, tw_i;
                fdivmod((uint32_t)tile, p.d_img, n, r);
                fdivmod(r, p.d_tw, th_i, tw_i);
                const int iy = (int)th_i * kBpOT - 1 + gy;                 // image row of this thread's mid pixels
                const int ix0 = (int)tw_i * kBpOT - 1;
                const bool y_in = iy >= 0 && iy < p.H;
                const uint32_t b = tl & 1u;
                if (warp == 2) { bp_wait_timed(p, tm, 6, &mid_empty[b], ((tl >> 1) & 1u) ^ 1u); bp_wait_timed(p, tm, 7, &acc1_full[b], (tl >> 1) & 1u); }
                else { ptx::mbar_wait(&mid_empty[b], ((tl >> 1) & 1u) ^ 1u); ptx::mbar_wait(&acc1_full[b], (tl >> 1) & 1u); }
                const long long t_w1 = p.dbg ? clock64() : 0;
                ptx::tc_fence_after();
                uint8_t* mid = mid_s + b * p.mid_bytes;
                uint32_t va[16], vb[16];
                const uint32_t t_buf = t_quad + b * 2u * (uint32_t)p.np1;
                ptx::tmem_ld16(t_buf + t_col[0], va);
                if (ntask == 2) ptx::tmem_ld16(t_buf + t_col[1], vb);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    if (k < ntask) {
                        const uint32_t(&v)[16] = k ? vb : va;
                        uint32_t o[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float2 f2 = silu2_half(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), hb[i]);
                            o[i] = pack_bf16x2(f2.x, f2.y);
                        }
                        const int ix = ix0 + t_mx[k];
                        const bool in = y_in && ix >= 0 && ix < p.W;
                        uint32_t o0 = t_off[k], o1 = t_off[k] + 16u;
                        o0 ^= ((o0 >> 7) & p.m_swz_mask) << 4;
                        o1 ^= ((o1 >> 7) & p.m_swz_mask) << 4;
                        *reinterpret_cast<uint4*>(mid + o0) = in ? make_uint4(o[0], o[1], o[2], o[3]) : make_uint4(0, 0, 0, 0);
                        *reinterpret_cast<uint4*>(mid + o1) = in ? make_uint4(o[4], o[5], o[6], o[7]) : make_uint4(0, 0, 0, 0);
                    }
                }
                ptx::fence_proxy_async();                   // mid writes -> visible to UMMA
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    ptx::mbar_arrive(&mid_full[b]);
                    ptx::mbar_arrive(&acc1_empty[b]);
                }
                if (warp == 2) BP_ACC(8, t_w1);
            }
            if (warp == 2) bp_flush(p, tm, lane, 6, 8);
        } else {
            // ---- epilogue 2: conv-2 accumulators -> bias + SiLU (+ x) -> bf16 -> staging -> TMA store ----
            // tasks of this warp: chunks [half * cpw, (half + 1) * cpw) of both output M-tiles, handled M-tile by M-tile.
            // The shortcut rows of a tile are requested BEFORE the accumulator wait (they do not depend on it): an L2
            // round trip is ~700 cycles, and issued behind the wait it was the longest link of the per-tile chain.
            const int nch = p.np2 >> 4;                      // 2 or 4
            const int cpw = nch >> 1;                        // chunks per warp and M-tile: 1 or 2
            const bool issuer = (warp == 10) && lane == 0;
            const bool row_valid = gy < kBpOT;
            auto load_res = [&](size_t pix, bool in_img, uint4 (&r)[2][2]) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    r[q][0] = r[q][1] = make_uint4(0, 0, 0, 0);
                    if (kRes && q < cpw && in_img) {
                        const uint4* rp = reinterpret_cast<const uint4*>(p.x + pix * p.x_pixstride + (half * cpw + q) * 16);
                        r[q][0] = __ldg(rp);
                        r[q][1] = __ldg(rp + 1);
                    }
                }
            };
            auto finish = [&](const uint32_t (&v)[16], const uint4 (&r)[2], int chunk, bool valid, int tx) {
                float f[16];
                const float2* hbq = reinterpret_cast<const float2*>(bias2_s + chunk * 16);      // half-scaled bias (broadcast reads)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 f2 = silu2_half(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), hbq[i]);
                    f[2 * i] = f2.x;
                    f[2 * i + 1] = f2.y;
                }
                if (kRes) {
                    const uint32_t rr[8] = {r[0].x, r[0].y, r[0].z, r[0].w, r[1].x, r[1].y, r[1].z, r[1].w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 t2 = unpack_bf16x2(rr[i]);
                        f[2 * i] += t2.x;
                        f[2 * i + 1] += t2.y;
                    }
                }
                if (valid) {
                    const uint32_t row = (uint32_t)(gy * kBpOT + tx);
                    uint32_t o0 = row * p.a_row_bytes + (uint32_t)chunk * 32u, o1 = o0 + 16u;
                    o0 ^= ((o0 >> 7) & p.st_swz_mask) << 4;
                    o1 ^= ((o1 >> 7) & p.st_swz_mask) << 4;
                    *reinterpret_cast<uint4*>(st_buf + o0) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                                                        pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
                    *reinterpret_cast<uint4*>(st_buf + o1) = make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]),
                                                                        pack_bf16x2(f[12], f[13]), pack_bf16x2(f[14], f[15]));
                }
            };
            uint32_t tl = 0;
            for (int tile = cta; tile < p.spatial_tiles; tile += ctas, ++tl) {
                uint32_t n, r, th_i, tw_i;
                fdivmod((uint32_t)tile, p.d_img, n, r);
                fdivmod(r, p.d_tw, th_i, tw_i);
                const uint32_t b = tl & 1u;
                const int oy = (int)th_i * kBpOT + gy;
                const int ox0 = (int)tw_i * kBpOT + gx, ox1 = ox0 + 8;
                const bool valid0 = row_valid, valid1 = row_valid && gx + 8 < kBpOT;
                const bool in0 = valid0 && oy < p.H && ox0 < p.W, in1 = valid1 && oy < p.H && ox1 < p.W;
                const size_t pix0 = ((size_t)n * p.H + oy) * p.W + ox0;
                uint4 r0[2][2], r1[2][2];
                load_res(pix0, in0, r0);                     // shortcut rows of M-tile 0: in flight across the wait below
                if (warp == 10) bp_wait_timed(p, tm, 9, &acc2_full[b], (tl >> 1) & 1u); else ptx::mbar_wait(&acc2_full[b], (tl >> 1) & 1u);
                ptx::tc_fence_after();
                const long long t_w2 = p.dbg ? clock64() : 0;
                uint32_t va[16], vb[16];
                const uint32_t t_acc = t_quad + b * 2u * (uint32_t)p.np2 + (uint32_t)(half * cpw * 16);
                ptx::tmem_ld16(t_acc, va);
                if (cpw == 2) ptx::tmem_ld16(t_acc + 16u, vb);
                load_res(pix0 + 8, in1, r1);                 // M-tile 1's shortcut rows: in flight across M-tile 0's math
                // the previous tile's store must have finished reading the staging buffer
                if (issuer) ptx::bulk_wait_read0();
                ptx::named_bar_sync(2, 256);
                if (warp == 10) BP_ACC(10, t_w2);
                const long long t_w3 = p.dbg ? clock64() : 0;
                ptx::tmem_ld_wait();
                finish(va, r0[0], half * cpw, valid0, gx);
                if (cpw == 2) finish(vb, r0[1], half * cpw + 1, valid0, gx);
                ptx::tmem_ld16(t_acc + (uint32_t)p.np2, va);
                if (cpw == 2) ptx::tmem_ld16(t_acc + (uint32_t)p.np2 + 16u, vb);
                ptx::tmem_ld_wait();
                finish(va, r1[0], half * cpw, valid1, gx + 8);
                if (cpw == 2) finish(vb, r1[1], half * cpw + 1, valid1, gx + 8);
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&acc2_empty[b]);       // accumulator buffer free: MMA 2 of tile t+2 may start
                if (warp == 10) BP_ACC(11, t_w3);
                const long long t_w4 = p.dbg ? clock64() : 0;
                ptx::fence_proxy_async();
                ptx::named_bar_sync(2, 256);
                if (issuer) {
                    ptx::tma_store_4d(&map_y, st_buf, 0, (int)tw_i * kBpOT, (int)th_i * kBpOT, (int)n);
                    ptx::bulk_commit_group();
                }
                if (warp == 10) BP_ACC(12, t_w4);
            }
            if (warp == 10) bp_flush(p, tm, lane, 9, 12);
            if (issuer) ptx::bulk_wait_read0();
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, p.tmem_cols);
}

// Shapes the fused kernel takes (everything else runs as two conv launches).
bool bneck_pair_ok(int C, int Cmid, int Cout, int np1, int np2) {
    if (env_flag("SPECYOLO_NO_BNECK_PAIR")) return false;
    if (C != Cout || (C != 32 && C != 64)) return false;
    if (Cmid != 16 && Cmid != 32) return false;
    if (np1 != Cmid || np2 != Cout) return false;
    return true;
}

int bneck_pair_launch(const specyolo_bneck_t* a, cudaStream_t stream) {
    SY_CHECK(bneck_pair_ok(a->C, a->Cmid, a->Cout, a->n_pad1, a->n_pad2), SPECYOLO_ERR_UNSUPPORTED,
             "bottleneck: unsupported shape (C=%d Cmid=%d Cout=%d)", a->C, a->Cmid, a->Cout);
    SY_CHECK(!(reinterpret_cast<uintptr_t>(a->x) & 15) && !(reinterpret_cast<uintptr_t>(a->y) & 15) &&
                 a->x_pixstride % 8 == 0 && a->y_pixstride % 8 == 0,
             SPECYOLO_ERR_INVALID, "bottleneck: x / y must be 16-byte aligned with pixel strides that are multiples of 8");
    EncodeTiledFn encode = get_encode_fn();
    SY_CHECK(encode != nullptr, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    BneckParams p{};
    p.B = a->B; p.H = a->H; p.W = a->W;
    p.tiles_w = ceil_div(a->W, kBpOT);
    p.tiles_h = ceil_div(a->H, kBpOT);
    const long spatial = (long)a->B * p.tiles_w * p.tiles_h;
    SY_CHECK(spatial > 0 && spatial < (1L << 30) && fastdiv_ok((uint64_t)spatial, (uint32_t)(p.tiles_w * p.tiles_h)),
             SPECYOLO_ERR_INVALID, "bottleneck: bad tile count");
    p.spatial_tiles = (int)spatial;
    p.d_img = make_fastdiv((uint32_t)(p.tiles_w * p.tiles_h));
    p.d_tw = make_fastdiv((uint32_t)p.tiles_w);
    p.c = a->C; p.cm = a->Cmid; p.np1 = a->n_pad1; p.np2 = a->n_pad2;
    p.b1 = a->b1; p.b2 = a->b2;
    p.x = reinterpret_cast<const __nv_bfloat16*>(a->x);
    p.x_pixstride = a->x_pixstride;
    p.add = a->add;
    p.a_row_bytes = (uint32_t)a->C * 2;
    p.m_row_bytes = (uint32_t)a->Cmid * 2;
    p.a_tx_bytes = (uint32_t)(kBpIW * kBpIW) * p.a_row_bytes;
    p.a_stage_bytes = (p.a_tx_bytes + 1023u) & ~1023u;
    p.mid_bytes = ((uint32_t)kBpMidRows * p.m_row_bytes + 1023u) & ~1023u;
    p.w1_box_bytes = (uint32_t)p.np1 * p.a_row_bytes;
    p.w2_box_bytes = (uint32_t)p.np2 * p.m_row_bytes;
    p.w_total_bytes = 9u * (p.w1_box_bytes + p.w2_box_bytes);
    p.off_w2 = (9u * p.w1_box_bytes + 1023u) & ~1023u;
    p.off_ring = p.off_w2 + ((9u * p.w2_box_bytes + 1023u) & ~1023u);
    const uint32_t st_bytes = ((uint32_t)(kBpOT * kBpOT) * p.a_row_bytes + 1023u) & ~1023u;
    const long avail = (long)kBpMaxDynSmem - 1024 - (long)p.off_ring - 2L * p.mid_bytes - (long)st_bytes;
    int stages = (int)(avail / (long)p.a_stage_bytes);
    SY_CHECK(stages >= 2, SPECYOLO_ERR_UNSUPPORTED, "bottleneck: shared memory budget exceeded");
    if (stages > kBpMaxStages) stages = kBpMaxStages;
    p.stages = stages;
    p.off_mid = p.off_ring + (uint32_t)stages * p.a_stage_bytes;
    p.off_st = p.off_mid + 2u * p.mid_bytes;
    const size_t smem_bytes = 1024 + (size_t)p.off_st + st_bytes;
    uint32_t cols = 32;
    while (cols < 4u * (uint32_t)p.np2 + 4u * (uint32_t)p.np1) cols <<= 1;
    SY_CHECK(cols <= 512, SPECYOLO_ERR_UNSUPPORTED, "bottleneck: TMEM budget exceeded");
    p.tmem_cols = cols;
    for (int t = 0; t < 9; ++t) {
        p.tap_a16[t] = ((uint32_t)((t / 3) * kBpIW + (t % 3)) * p.a_row_bytes) >> 4;
        p.tap_m16[t] = ((uint32_t)((t / 3) * kBpMW + (t % 3)) * p.m_row_bytes) >> 4;
    }
    p.st_swz_mask = p.a_row_bytes == 128 ? 7u : (p.a_row_bytes == 64 ? 3u : 1u);
    p.m_swz_mask = p.m_row_bytes == 128 ? 7u : (p.m_row_bytes == 64 ? 3u : 1u);

    CUtensorMap map_x, map_w1, map_w2, map_y;
    {
        const cuuint64_t pix_b = (cuuint64_t)a->x_pixstride * 2;
        cuuint64_t dims[4] = {(cuuint64_t)a->C, (cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->B};
        cuuint64_t strides[3] = {pix_b, pix_b * a->W, pix_b * a->W * a->H};
        cuuint32_t box[4] = {(cuuint32_t)a->C, kBpIW, kBpIW, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a->x), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for((int)p.a_row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SY_CHECK(r == CUDA_SUCCESS, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled(bottleneck X) failed (%d)", (int)r);
    }
    {
        const cuuint64_t k1 = 9ull * a->C;
        cuuint64_t dims[2] = {k1, (cuuint64_t)p.np1};
        cuuint64_t strides[1] = {k1 * 2};
        cuuint32_t box[2] = {(cuuint32_t)a->C, (cuuint32_t)p.np1};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map_w1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a->w1_packed), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for((int)p.a_row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SY_CHECK(r == CUDA_SUCCESS, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled(bottleneck W1) failed (%d)", (int)r);
    }
    {
        const cuuint64_t k2 = 9ull * a->Cmid;
        cuuint64_t dims[2] = {k2, (cuuint64_t)p.np2};
        cuuint64_t strides[1] = {k2 * 2};
        cuuint32_t box[2] = {(cuuint32_t)a->Cmid, (cuuint32_t)p.np2};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map_w2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a->w2_packed), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for((int)p.m_row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SY_CHECK(r == CUDA_SUCCESS, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled(bottleneck W2) failed (%d)", (int)r);
    }
    {
        const cuuint64_t pix_b = (cuuint64_t)a->y_pixstride * 2;
        cuuint64_t dims[4] = {(cuuint64_t)a->Cout, (cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->B};
        cuuint64_t strides[3] = {pix_b, pix_b * a->W, pix_b * a->W * a->H};
        cuuint32_t box[4] = {(cuuint32_t)a->Cout, kBpOT, kBpOT, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&map_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a->y, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for((int)p.a_row_bytes), CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SY_CHECK(r == CUDA_SUCCESS, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled(bottleneck Y) failed (%d)", (int)r);
    }
    static unsigned long long* dbg_dev = nullptr;
    const bool dbg = env_flag("SPECYOLO_BP_DBG");
    if (dbg) {
        if (!dbg_dev) SY_CUDA(cudaMalloc(&dbg_dev, 16 * sizeof(unsigned long long)));
        SY_CUDA(cudaMemsetAsync(dbg_dev, 0, 16 * sizeof(unsigned long long), stream));
        p.dbg = dbg_dev;
    }
    typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const BneckParams);
    static const KernelFn kernels[2] = {bneck_pair_kernel<false>, bneck_pair_kernel<true>};
    for (KernelFn k : kernels)      // per-device attribute: set on every launch (host-side table lookup)
        SY_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kBpMaxDynSmem));
    const long resident = sm_count();
    unsigned grid = (unsigned)(spatial < resident ? spatial : resident);
    { const char* e = std::getenv("SPECYOLO_BP_GRID"); if (e && atoi(e) > 0 && (unsigned)atoi(e) < grid) grid = (unsigned)atoi(e); }
    SY_CUDA(launch_pdl(kernels[a->add ? 1 : 0], dim3(grid), dim3(kBpThreads), smem_bytes, stream, map_x, map_w1, map_w2, map_y, p));
    SY_LAUNCH_CHECK();
    count_launch();
    if (dbg) {
        unsigned long long h[16];
        SY_CUDA(cudaStreamSynchronize(stream));
        SY_CUDA(cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost));
        const double n = h[15] ? (double)h[15] : 1.0;
        fprintf(stderr, "[bneck dbg] C=%d Cmid=%d %dx%d tiles/CTA0=%llu stages=%d | per tile cycles: loop %.0f | producer wait-empty %.0f | "
                "mma wait acc1_empty %.0f full %.0f acc2_empty %.0f mid_full %.0f | epi1 wait mid_empty %.0f acc1_full %.0f work %.0f | "
                "epi2 wait acc2_full %.0f staging %.0f work %.0f store %.0f | issue mma1 %.0f mma2 %.0f\n", a->C, a->Cmid, a->H, a->W, h[15], p.stages,
                h[5] / n, h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[6] / n, h[7] / n, h[8] / n, h[9] / n, h[10] / n, h[11] / n, h[12] / n,
                h[13] / n, h[14] / n);
    }
    return SPECYOLO_OK;
}

}  // namespace specyolo
