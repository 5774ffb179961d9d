// Halo-tile implicit-GEMM convolution for sm_100a: k x k convs (k >= 2) whose weights fit in shared memory.
//
// The per-tap kernel in conv_igemm.cu issues one TMA box per (tap, channel chunk), so a 3x3 conv pulls its input
// through the L2 -> SM fabric nine times (a 7x7: 49 times) — that fabric, not HBM or the tensor pipe, bounds every
// small-channel k x k layer of the network.  Here each CTA
//   * loads its weights ONCE (all taps, all channel chunks of its groups) and keeps them resident in shared memory
//     while it walks its share of the output tiles (persistent CTA), and
//   * loads, per output tile and 64-channel chunk, ONE input box that already contains the halo:
//     (16 + (k-1)d) x (8 + (k-1)d) pixels for a 16 x 8 output tile.  The box lands as consecutive K-major rows
//     (one row per pixel, TMA swizzle).  The A operand of tap (ky,kx) is the SAME box read through a UMMA
//     shared-memory descriptor whose start address is shifted by (ky*d*halo_w + kx*d) rows and whose 8-row-group
//     stride (SBO) is halo_w rows: row group g of the 128-row operand = the 8 pixels of output row g.  The
//     swizzle is a function of the absolute shared-memory address, so a shifted, non-1024-aligned start is legal
//     (tests/cuda/umma_shift_probe.cu checks this on the device).
// Input traffic drops from taps x to ~1.4 x (3x3) / ~2.4 x (7x7) of the tensor.
//
// Strided convs whose taps all fall on one sub-lattice (stride s, dilation and padding multiples of s — DDWConv's
// k7/s2/d2/p6 and k3/s2/d2/p2, ultralytics/nn/modules/conv.py:694-710) are the same problem on the s-subsampled
// input: the TMA tensor map's elementStrides do the subsampling.
//
// Grouped convs are NOT merged to block-diagonal weights here: a 64-channel A box holds 64/cin_g groups, and each
// group's 16-channel K step multiplies its own [n_pad x cin_g] weight box into its own TMEM column range
// (UMMA N = n_pad = 16 for DDWConv) — no multiplications by zero.
//
// Warp roles as in conv_igemm.cu: warp 0 TMA producer, warp 1 TMEM owner + MMA issuer, warps 2-5 epilogue;
// the accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the loads / MMAs of tile i+1.
//
// Replaces Conv.forward_fuse (ultralytics/nn/modules/conv.py:81-83) for the k x k layers.
#include "common.h"
#include "ptx.cuh"
#include "tma_host.h"
#include "epilogue.cuh"

#include <cstdlib>

namespace specyolo {

struct HaloParams {
    int tiles_w, tiles_h;        // 8-wide x 16-high output tiles per image
    int B, Ho, Wo;
    int spatial_tiles;           // B * tiles_h * tiles_w
    int supert, nsuper;          // output tiles per accumulator stage (thin tiles: amortises the per-stage hand-offs)
    FastDiv d_img, d_tw;         // tile -> (image, tile row, tile column)
    FastDiv d_cin_g, d_npad;     // channel -> (group, channel in group); accumulator column -> (group, column in group)
    int kcb_log2;
    uint32_t tap_a16[49];        // A start-address shift of every tap, 16-byte units
    int gsplit;                  // CTAs with blockIdx.x % gsplit == s own channel split s
    int kw, taps;
    int dil, pad, sub;           // dilation / padding on the subsampled lattice; lattice step (TMA element stride)
    int halo_w;                  // pixels (= smem rows) per halo line
    int kc_box, boxes;           // channels per A box; A boxes per tile
    int cin_g, kc_b, bchunks;    // weight boxes: kc_b channels wide, bchunks per (group, tap)
    int groups_cta, n_pad, cout_g, cin_cta;
    int ncols;                   // accumulator columns per tile = groups_cta * n_pad
    uint32_t a_row_bytes, b_row_bytes;
    uint32_t a_stage_bytes, a_tx_bytes, b_box_bytes, b_total_bytes, b_region_bytes;
    int stages;
    uint32_t tmem_cols;
    const float* bias;
    void* y;
    int y_pixstride, y_fp32;
    const __nv_bfloat16* residual;
    int r_pixstride;
    int act;
    // TMA-store epilogue (store_bw == 0: direct stores)
    int store_bw, pair_stores;
    int y_s2d;                   // output written 2x2-blocked (see specyolo_conv_t::y_s2d)
    int thin;                    // <= 32-column tiles on the paired-task epilogue (see the epilogue branch)
    // residual tile ring (two slots after the staging buffer): the TMA producer loads the tile's residual box(es)
    // tiles ahead, the epilogue reads them from shared memory (res_box_cols == 0: residual read from global memory)
    int res_box_cols, res_slots;
    uint32_t res_row_bytes, res_swz_mask, res_slot_bytes, res_off;
    uint32_t store_row_bytes, store_swz_mask, ring_bytes;
};

static constexpr int kHaloThreads = kConvThreads;
static constexpr int kHaloTW = 8, kHaloTH = 16;
static constexpr int kHaloMaxStages = 8;
static constexpr int kHaloMaxBias = 1024;
static constexpr int kHaloMaxResSlots = 6;     // residual tiles in flight (the producer runs this many tiles ahead)
static constexpr int kHaloMaxDynSmem = 222 * 1024;

__device__ __forceinline__ uint64_t halo_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t row_bytes) {
    const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
    d |= 1ull << 46;
    d |= layout << 61;
    return d;
}

// Thin tiles (16 / 32 accumulator columns): one epilogue task = one 16-column chunk of one sub-tile of a super-tile
// stage.  A tile this thin carries ~300 cycles of useful epilogue work behind ~1000 cycles of dependent latency
// (tile decode -> tcgen05.ld -> bias -> MUFU -> pack -> store; clock64 phase timers), so every warp keeps TWO tasks in
// flight — both accumulator loads and both residual loads are issued before the single wait — and the bias chunk of
// the warp lives in registers for the whole kernel.
struct ThinTask {
    uint32_t taddr;
    bool ok;
    size_t yoff, roff;           // element offsets of this lane's row (first channel of the chunk)
};
__device__ __forceinline__ ThinTask thin_setup(const HaloParams& p, int tile, int tw, int th, int gch, uint32_t taddr) {
    ThinTask t;
    uint32_t n, r, th_i, tw_i;
    fdivmod((uint32_t)tile, p.d_img, n, r);
    fdivmod(r, p.d_tw, th_i, tw_i);
    const int ow = (int)tw_i * 8 + tw, oh = (int)th_i * 16 + th;
    t.ok = (ow < p.Wo) && (oh < p.Ho);
    t.taddr = taddr;
    const uint32_t pix = ((uint32_t)n * (uint32_t)p.Ho + (uint32_t)oh) * (uint32_t)p.Wo + (uint32_t)ow;
    if (p.y_s2d) {
        const uint32_t bpix = ((uint32_t)n * (uint32_t)(p.Ho >> 1) + (uint32_t)(oh >> 1)) * (uint32_t)(p.Wo >> 1) + (uint32_t)(ow >> 1);
        t.yoff = (size_t)bpix * (uint32_t)p.y_pixstride + (uint32_t)(gch + (((oh & 1) << 1) | (ow & 1)) * p.cout_g);
    } else {
        t.yoff = (size_t)pix * (uint32_t)p.y_pixstride + (uint32_t)gch;
    }
    t.roff = (size_t)pix * (uint32_t)p.r_pixstride + (uint32_t)gch;
    return t;
}
template <bool kSilu, bool kRes>
__device__ __forceinline__ void thin_finish(const uint32_t (&v)[16], const float (&bias)[16], const ThinTask& t,
                                            __nv_bfloat16* y, const uint4& r0, const uint4& r1) {
    float f[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float2 a = make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
        const float2 b = make_float2(bias[2 * i], bias[2 * i + 1]);
        const float2 r = kSilu ? silu2_half(a, b) : fadd2(a, b);
        f[2 * i] = r.x;
        f[2 * i + 1] = r.y;
    }
    if (kRes) {
        const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float2 q = unpack_bf16x2(rr[i]);
            f[2 * i] += q.x;
            f[2 * i + 1] += q.y;
        }
    }
    if (t.ok) {
        uint4 o0, o1;
        o0.x = pack_bf16x2(f[0], f[1]);   o0.y = pack_bf16x2(f[2], f[3]);
        o0.z = pack_bf16x2(f[4], f[5]);   o0.w = pack_bf16x2(f[6], f[7]);
        o1.x = pack_bf16x2(f[8], f[9]);   o1.y = pack_bf16x2(f[10], f[11]);
        o1.z = pack_bf16x2(f[12], f[13]); o1.w = pack_bf16x2(f[14], f[15]);
        uint4* yp = reinterpret_cast<uint4*>(y + t.yoff);
        yp[0] = o0;
        yp[1] = o1;
    }
}

// (min 2 CTAs/SM: caps the residual variants at 102 registers — at 134 they silently ran one CTA per SM and every
//  residual layer took 2-2.7x the time of its residual-free twin)
template <bool kSilu, bool kRes, bool kFp32>
__global__ void __launch_bounds__(kHaloThreads, 2)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_r,
                 const __grid_constant__ HaloParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kHaloMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kHaloMaxStages];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ __align__(8) uint64_t w_bar;
    __shared__ __align__(8) uint64_t res_full[kHaloMaxResSlots];
    __shared__ __align__(8) uint64_t res_empty[kHaloMaxResSlots];
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float bias_s[kHaloMaxBias];

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform for the compiler
    const int lane = threadIdx.x & 31;
    ptx::grid_dep_launch();     // PDL: the next kernel may be scheduled into SM slots as this grid drains
    const int split = blockIdx.x % p.gsplit;
    const int cta = blockIdx.x / p.gsplit;
    const int ctas = gridDim.x / p.gsplit;

    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    const uint32_t off = ((raw_addr + 1023u) & ~1023u) - raw_addr;
    uint8_t* b_s = smem_raw + off;                  // resident weights
    uint8_t* a_ring = b_s + p.b_region_bytes;       // halo boxes

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
            ptx::mbar_init(&tmem_empty_bar[b], kEpiWarps);
        }
        ptx::mbar_init(&w_bar, 1);
        for (int b = 0; b < kHaloMaxResSlots; ++b) {
            ptx::mbar_init(&res_full[b], 1);
            ptx::mbar_init(&res_empty[b], kEpiWarps);
        }
        if (p.res_box_cols) ptx::prefetch_tmap(&map_r);
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc(&tmem_base_smem, p.tmem_cols);
    {
        const float* bsrc = p.bias + (size_t)split * p.ncols;
        for (int i = threadIdx.x; i < p.ncols; i += kHaloThreads) bias_s[i] = (kSilu ? 0.5f : 1.0f) * bsrc[i];
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    // PDL: the prologue above and the weight loads below (constants) overlap the previous kernel's tail; activations
    // are read and outputs written only after griddepcontrol.wait
    if (warp != 0) ptx::grid_dep_wait();

    if (warp == 0) {
        // ===================== TMA producer =====================
        // all lanes walk the loops converged (uniform datapath); the elected lane issues
        {
            const bool leader = ptx::elect_one();
            // resident weights: every (group, tap, chunk) box of this CTA's channel split, once
            if (leader) ptx::mbar_expect_tx(&w_bar, p.b_total_bytes);
            {
                uint8_t* dst = b_s;
                for (int gl = 0; gl < p.groups_cta; ++gl) {
                    const int row = (split * p.groups_cta + gl) * p.n_pad;
                    int kcol = 0;
                    for (int i = 0; i < p.taps * p.bchunks; ++i) {          // (tap, chunk) boxes are consecutive K columns
                        if (leader) ptx::tma_load_2d(dst, &map_b, &w_bar, kcol, row);
                        dst += p.b_box_bytes;
                        kcol += p.kc_b;
                    }
                }
            }
            ptx::grid_dep_wait();
            int stage = 0;
            uint32_t ph = 0, rslot = 0, rph = 0;
            uint8_t* res_ring = a_ring + p.res_off;
            for (int st = cta; st < p.nsuper; st += ctas)
            for (int sub_t = 0; sub_t < p.supert; ++sub_t) {
                const int tile = st * p.supert + sub_t;
                if (tile >= p.spatial_tiles) break;
                uint32_t n, r, th_i, tw_i;
                fdivmod((uint32_t)tile, p.d_img, n, r);
                fdivmod(r, p.d_tw, th_i, tw_i);
                const int w0 = ((int)tw_i * kHaloTW - p.pad) * p.sub;
                const int h0 = ((int)th_i * kHaloTH - p.pad) * p.sub;
                if (kRes && p.res_box_cols) {       // this tile's residual box(es) into the next ring slot
                    ptx::mbar_wait(&res_empty[rslot], rph ^ 1u);
                    if (leader) {
                        ptx::mbar_expect_tx(&res_full[rslot], p.res_slot_bytes);
                        uint8_t* dst = res_ring + rslot * p.res_slot_bytes;
                        for (int c = 0; c < p.ncols; c += p.res_box_cols) {
                            ptx::tma_load_4d(dst, &map_r, &res_full[rslot], split * p.groups_cta * p.cout_g + c,
                                             (int)tw_i * kHaloTW, (int)th_i * kHaloTH, (int)n);
                            dst += 128u * p.res_row_bytes;
                        }
                    }
                    if (++rslot == (uint32_t)p.res_slots) { rslot = 0; rph ^= 1u; }
                }
                int ch = split * p.cin_cta;
                for (int box = 0; box < p.boxes; ++box) {
                    ptx::mbar_wait(&empty_bar[stage], ph ^ 1u);
                    if (leader) {
                        ptx::mbar_expect_tx(&full_bar[stage], p.a_tx_bytes);
                        ptx::tma_load_4d(a_ring + (size_t)stage * p.a_stage_bytes, &map_a, &full_bar[stage], ch, w0, h0,
                                         (int)n);
                    }
                    ch += p.kc_box;
                    if (++stage == p.stages) { stage = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // All lanes walk the loops converged so that descriptors live in uniform registers; the elected lane issues.
        // No runtime divisions: K step outermost (group / weight-box lookup once per step), taps innermost.
        {
            const bool leader = ptx::elect_one();
            const uint32_t idesc = ptx::umma_idesc_bf16(128, p.n_pad);
            const uint32_t a_hi = (uint32_t)(halo_desc(0, (uint32_t)p.halo_w * p.a_row_bytes, p.a_row_bytes) >> 32);
            const uint32_t b_hi = (uint32_t)(halo_desc(0, 8u * p.b_row_bytes, p.b_row_bytes) >> 32);
            const uint32_t a_ring16 = ptx::smem_u32(a_ring) >> 4, b_base16 = ptx::smem_u32(b_s) >> 4;
            const uint32_t b_box16 = p.b_box_bytes >> 4;
            const uint32_t b_tap16 = (uint32_t)p.bchunks * b_box16;       // next tap of the same (group, chunk)
            const uint32_t b_grp16 = (uint32_t)p.taps * b_tap16;          // next group
            const int ksteps = p.kc_box / 16;
            const bool single_group = p.groups_cta == 1;        // (then kc_box == kc_b)
            ptx::mbar_wait(&w_bar, 0);
            ptx::tc_fence_after();
            int stage = 0;
            uint32_t ph = 0, tl = 0;
            for (int st = cta; st < p.nsuper; st += ctas, ++tl) {
                const uint32_t buf = tl & 1u;
                const uint32_t bph = (tl >> 1) & 1u;
                ptx::mbar_wait(&tmem_empty_bar[buf], bph ^ 1u);
                ptx::tc_fence_after();
                for (int sub_t = 0; sub_t < p.supert; ++sub_t) {
                if (st * p.supert + sub_t >= p.spatial_tiles) break;
                const uint32_t d_tmem = tmem_base + (buf * (uint32_t)p.supert + (uint32_t)sub_t) * (uint32_t)p.ncols;
                for (int box = 0; box < p.boxes; ++box) {
                    ptx::mbar_wait(&full_bar[stage], ph);
                    ptx::tc_fence_after();
                    const uint32_t a16 = a_ring16 + (((uint32_t)stage * p.a_stage_bytes) >> 4);
                    if (single_group) {
                        // one group per CTA (every dense conv): box == weight chunk, no group / chunk lookup
                        uint32_t b16k = b_base16 + (uint32_t)box * b_box16;
                        uint32_t a16k = a16;
                        for (int k = 0; k < ksteps; ++k) {
                            uint32_t b16 = b16k;
                            uint32_t accumulate = (box | k) != 0 ? 1u : 0u;
                            for (int tap = 0; tap < p.taps; ++tap) {
                                if (leader) ptx::umma_bf16_lohi(d_tmem, a16k + p.tap_a16[tap], a_hi, b16, b_hi, idesc, accumulate);
                                accumulate = 1u;
                                b16 += b_tap16;
                            }
                            b16k += 2u;
                            a16k += 2u;
                        }
                    } else
                    for (int k = 0; k < ksteps; ++k) {
                        uint32_t gl, cio;
                        fdivmod((uint32_t)(box * p.kc_box + k * 16), p.d_cin_g, gl, cio);   // channel inside the split
                        const uint32_t bc = cio >> p.kcb_log2;
                        const uint32_t kb = (cio - (bc << p.kcb_log2)) >> 4;
                        uint32_t b16 = b_base16 + gl * b_grp16 + bc * b_box16 + 2u * kb;
                        const uint32_t a16k = a16 + 2u * (uint32_t)k;
                        const uint32_t d_col = d_tmem + gl * (uint32_t)p.n_pad;
                        uint32_t accumulate = cio == 0 ? 0u : 1u;         // first K step of this group
                        for (int tap = 0; tap < p.taps; ++tap) {
                            if (leader) ptx::umma_bf16_lohi(d_col, a16k + p.tap_a16[tap], a_hi, b16, b_hi, idesc, accumulate);
                            accumulate = 1u;
                            b16 += b_tap16;
                        }
                    }
                    if (leader) ptx::umma_commit(&empty_bar[stage]);
                    if (++stage == p.stages) { stage = 0; ph ^= 1u; }
                }
                }
                if (leader) ptx::umma_commit(&tmem_full_bar[buf]);
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue (warps 2..9, see epilogue.cuh) =====================
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;
        const int m = quad * 32 + lane;
        const int tw = m & (kHaloTW - 1), th = m >> 3;
        if (!kFp32 && p.thin) {
            // paired-task path (ThinTask above).  32 columns: this half owns chunk `half` of every sub-tile;
            // 16 columns: the halves take alternate sub-tiles.
            const bool wide = p.ncols == 32;
            const int chunk = wide ? half : 0;
            const int s_first = wide ? 0 : half, s_step = wide ? 1 : 2;
            const int gch = split * p.cout_g + chunk * 16;
            float bias_r[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) bias_r[i] = bias_s[chunk * 16 + i];
            __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(p.y);
            const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(chunk * 16);
            uint32_t tl = 0;
            for (int sti = cta; sti < p.nsuper; sti += ctas, ++tl) {
                const uint32_t buf = tl & 1u;
                const uint32_t bph = (tl >> 1) & 1u;
                const int tile0 = sti * p.supert;
                const int nsub = min(p.supert, p.spatial_tiles - tile0);
                if (kRes && sti + ctas < p.nsuper) {            // next stage's residual rows of this warp -> L2
                    const int tile1 = (sti + ctas) * p.supert;
                    const int nsub1 = min(p.supert, p.spatial_tiles - tile1);
                    for (int sb = s_first; sb < nsub1; sb += s_step) {
                        const ThinTask q = thin_setup(p, tile1 + sb, tw, th, gch, 0u);
                        if (q.ok) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.residual + q.roff));
                    }
                }
                const uint32_t t_stage = t_lane + buf * (uint32_t)p.supert * (uint32_t)p.ncols;
                ptx::mbar_wait(&tmem_full_bar[buf], bph);
                ptx::tc_fence_after();
                for (int sb = s_first; sb < nsub; sb += 2 * s_step) {
                    const bool two = sb + s_step < nsub;
                    uint32_t va[16], vb[16];
                    const ThinTask ta = thin_setup(p, tile0 + sb, tw, th, gch, t_stage + (uint32_t)sb * (uint32_t)p.ncols);
                    ptx::tmem_ld16(ta.taddr, va);
                    ThinTask tb = ta;
                    if (two) {
                        tb = thin_setup(p, tile0 + sb + s_step, tw, th, gch, t_stage + (uint32_t)(sb + s_step) * (uint32_t)p.ncols);
                        ptx::tmem_ld16(tb.taddr, vb);
                    }
                    uint4 ra0 = make_uint4(0, 0, 0, 0), ra1 = ra0, rb0 = ra0, rb1 = ra0;
                    if (kRes) {
                        if (ta.ok) {
                            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + ta.roff);
                            ra0 = __ldg(rp); ra1 = __ldg(rp + 1);
                        }
                        if (two && tb.ok) {
                            const uint4* rp = reinterpret_cast<const uint4*>(p.residual + tb.roff);
                            rb0 = __ldg(rp); rb1 = __ldg(rp + 1);
                        }
                    }
                    ptx::tmem_ld_wait();
                    thin_finish<kSilu, kRes>(va, bias_r, ta, yb, ra0, ra1);
                    if (two) thin_finish<kSilu, kRes>(vb, bias_r, tb, yb, rb0, rb1);
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[buf]);
            }
        } else {
        EpiOut eo{p.y, p.y_pixstride, p.residual, p.r_pixstride, p.pair_stores != 0};
        EpiStage st = epi_make_stage(a_ring + p.ring_bytes, &map_y, p.ncols, p.store_bw, p.store_row_bytes, p.store_swz_mask,
                                     warp, half, lane, m);
        EpiCols ec;
        ec.ncols = p.ncols;
        ec.n_pad = p.n_pad;
        ec.d_npad = p.d_npad;
        ec.cout_g = p.cout_g;
        ec.within0 = 0;
        ec.gch0 = split * p.groups_cta * p.cout_g;
        uint32_t tl = 0, rslot = 0, rph = 0;
        const bool res_ring_on = kRes && p.res_box_cols != 0;
        for (int sti = cta; sti < p.nsuper; sti += ctas, ++tl) {
            const uint32_t buf = tl & 1u;
            const uint32_t bph = (tl >> 1) & 1u;
            if (kRes && !res_ring_on && sti + ctas < p.nsuper) {            // next stage's residual rows -> L2
                for (int sub_t = 0; sub_t < p.supert; ++sub_t) {
                    const int t2 = (sti + ctas) * p.supert + sub_t;
                    if (t2 >= p.spatial_tiles) break;
                    uint32_t n2, r2, th2, tw2;
                    fdivmod((uint32_t)t2, p.d_img, n2, r2);
                    fdivmod(r2, p.d_tw, th2, tw2);
                    const int ow2 = (int)tw2 * kHaloTW + tw, oh2 = (int)th2 * kHaloTH + th;
                    epi_prefetch_residual_l2<kRes>(ec, eo, ((size_t)n2 * p.Ho + oh2) * p.Wo + ow2, (ow2 < p.Wo) && (oh2 < p.Ho), half);
                }
            }
            ptx::mbar_wait(&tmem_full_bar[buf], bph);       // all `supert` accumulators of this stage are complete
            ptx::tc_fence_after();
            for (int sub_t = 0; sub_t < p.supert; ++sub_t) {
                const int tile = sti * p.supert + sub_t;
                if (tile >= p.spatial_tiles) break;
                uint32_t n, r, th_i, tw_i;
                fdivmod((uint32_t)tile, p.d_img, n, r);
                fdivmod(r, p.d_tw, th_i, tw_i);
                const int ow = (int)tw_i * kHaloTW + tw, oh = (int)th_i * kHaloTH + th;
                const bool row_ok = (ow < p.Wo) && (oh < p.Ho);
                size_t pix = ((size_t)n * p.Ho + oh) * p.Wo + ow;
                if (p.y_s2d) {      // blocked output: pixel -> (block pixel, channel slab of the 2x2 position)
                    pix = ((size_t)n * (p.Ho >> 1) + (oh >> 1)) * (p.Wo >> 1) + (ow >> 1);
                    ec.gch0 = (((oh & 1) << 1) | (ow & 1)) * p.cout_g;
                }
                const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) +
                                        (buf * (uint32_t)p.supert + (uint32_t)sub_t) * (uint32_t)p.ncols;
                EpiResSmem rs{nullptr, 1, 0, 0, m};
                if (res_ring_on) {
                    rs.base = a_ring + p.res_off + rslot * p.res_slot_bytes;
                    rs.box_cols = p.res_box_cols;
                    rs.row_bytes = p.res_row_bytes;
                    rs.swz_mask = p.res_swz_mask;
                    ptx::mbar_wait(&res_full[rslot], rph);
                }
                st.c0 = ec.gch0;
                st.c1 = (int)tw_i * kHaloTW; st.c2 = (int)th_i * kHaloTH; st.c3 = (int)n;
                epi_tile<kSilu, kRes, kFp32>(t_addr, ec, bias_s, eo, pix, row_ok, half, lane, st, rs);
                if (res_ring_on) {
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&res_empty[rslot]);
                    if (++rslot == (uint32_t)p.res_slots) { rslot = 0; rph ^= 1u; }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[buf]);
        }
        if (st.enabled && st.issuer) ptx::bulk_wait_read0();   // staging must outlive the last store's read
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int halo_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = std::getenv("SPECYOLO_HALO");   // 0 disables the halo path (A/B measurements)
        mode = (e && e[0] == '0') ? 0 : 1;
    }
    return mode;
}

// Geometry test shared with fold_pack's merge decision: can this conv run on the halo kernel at all?
// (the weights-fit test needs n_pad and is done in conv_halo_plan)
bool conv_halo_geometry_ok(int kh, int kw, int stride, int pad, int dil) {
    if (!halo_mode()) return false;
    if (kh != kw || kh < 2 || kh > 7) return false;
    if (stride < 1 || stride > 2) return false;
    if (dil % stride || pad % stride) return false;
    return true;
}

// Paired-task epilogue (ThinTask): one full 16-channel chunk per task, plain 16-byte stores.
static bool thin_ok(const specyolo_conv_t* a, int ncols, int gcta, int cout_g, int store_bw) {
    if (env_flag("SPECYOLO_NO_THIN")) return false;
    if (ncols > 32 || a->y_fp32 || store_bw || gcta != 1 || cout_g != ncols) return false;
    if ((reinterpret_cast<uintptr_t>(a->y) & 15) || a->y_pixstride % 8) return false;
    if (a->residual && ((reinterpret_cast<uintptr_t>(a->residual) & 15) || a->r_pixstride % 8 || a->y_s2d ||
                        env_flag("SPECYOLO_RES_RING")))
        return false;
    if ((long)a->B * a->Ho * a->Wo >= (1L << 31)) return false;
    return true;
}

struct HaloPlan {
    HaloParams p;
    size_t smem_bytes;
    int occ;
    unsigned grid;
};

// Returns true and fills `plan` when the conv is eligible for the halo kernel.
static bool conv_halo_plan(const specyolo_conv_t* a, HaloPlan& plan) {
    if (!conv_halo_geometry_ok(a->kh, a->kw, a->stride, a->pad, a->dil)) return false;
    const int groups = a->groups;
    if (a->Cin % groups || a->Cout % groups) return false;
    const int cin_g = a->Cin / groups, cout_g = a->Cout / groups;
    if (cin_g % 16 || a->n_pad % 16 || a->n_pad < cout_g || a->n_pad > 256) return false;
    if (a->x_pixstride % 8 || (reinterpret_cast<uintptr_t>(a->x) & 15)) return false;

    HaloParams p{};
    p.sub = a->stride;
    p.dil = a->dil / p.sub;
    p.pad = a->pad / p.sub;
    p.kw = a->kw;
    p.taps = a->kh * a->kw;
    p.halo_w = kHaloTW + (a->kw - 1) * p.dil;
    const int halo_h = kHaloTH + (a->kh - 1) * p.dil;
    if (p.halo_w * p.sub > 256 || halo_h * p.sub > 256) return false;
    p.cin_g = cin_g;
    p.cout_g = cout_g;
    p.n_pad = a->n_pad;
    p.kc_b = (cin_g % 64 == 0) ? 64 : (cin_g % 32 == 0 ? 32 : 16);
    p.bchunks = cin_g / p.kc_b;
    p.b_row_bytes = (uint32_t)p.kc_b * 2;
    p.b_box_bytes = (uint32_t)p.n_pad * p.b_row_bytes;

    // channel split across CTAs: the smallest that lets the weights stay resident beside a >= 3-deep halo ring
    // (a 2-deep ring is accepted only if no split reaches 3)
    bool found = false;
    int best_stages = 0;
    for (int gs = 1; gs <= groups && groups % gs == 0 && best_stages < 3; gs *= 2) {
        const int gcta = groups / gs;
        const int cin_cta = gcta * cin_g;
        // one A box spans several whole groups when the groups are thinner than 64 channels
        const int kc_box = (gcta > 1 && cin_g < 64 && 64 % cin_g == 0 && cin_cta % 64 == 0) ? 64 : p.kc_b;
        const int ncols = gcta * p.n_pad;
        if (ncols > 256) continue;
        const uint32_t a_row = (uint32_t)kc_box * 2;
        const uint32_t a_tx = (uint32_t)(p.halo_w * halo_h) * a_row;
        const uint32_t a_stage = (a_tx + 1023u) & ~1023u;
        const uint32_t b_total = (uint32_t)(gcta * p.taps * p.bchunks) * p.b_box_bytes;
        const uint32_t b_region = (b_total + 1023u) & ~1023u;
        if (b_total >= (1u << 20)) continue;                      // mbarrier tx-count limit
        // TMA-store staging (two halves x 128 rows x box row): only when the columns map to consecutive channels
        const int es = a->y_fp32 ? 4 : 2;
        int store_bw = epi_stage_box_cols(ncols, es);
        if ((reinterpret_cast<uintptr_t>(a->y) & 15) || ((size_t)a->y_pixstride * es) % 16 ||
            (gcta > 1 && cout_g != p.n_pad) || a->y_s2d || env_flag("SPECYOLO_NO_TMA_STORE"))
            store_bw = 0;
        uint32_t stage_out = epi_stage_bytes(ncols, store_bw, es);
        if ((long)kHaloMaxDynSmem - 1024 - (long)b_region - (long)stage_out < 3L * a_stage) {   // keep a 3-deep ring
            store_bw = 0;
            stage_out = 0;
        }
        // residual ring: slots of [128 rows x ncols] bf16 loaded by the TMA producer (tiles up to 64 columns).  OFF by
        // default (SPECYOLO_RES_RING=1 enables it): once the residual variants were capped at two CTAs per SM, the
        // global-memory residual path with its one-tile-ahead L2 prefetch measured the same or better
        // (32->64 3x3 @80^2: 46 us vs 53 us with the ring) and leaves the shared memory to the operand ring.
        int res_box = 0, res_slots = 0;
        uint32_t res_bytes = 0;
        if (a->residual && ncols <= 64 && (ncols & (ncols - 1)) == 0 && (gcta == 1 || cout_g == p.n_pad) &&
            !(reinterpret_cast<uintptr_t>(a->residual) & 15) && (a->r_pixstride % 8) == 0 && env_flag("SPECYOLO_RES_RING")) {
            // as many slots (tiles of run-ahead) as fit beside a 4-deep A ring, 2..kHaloMaxResSlots
            const uint32_t slot = 128u * (uint32_t)ncols * 2u;
            const long room = (long)kHaloMaxDynSmem - 1024 - (long)b_region - (long)stage_out - 4L * a_stage;
            res_slots = (int)(room / (long)slot);
            if (res_slots > kHaloMaxResSlots) res_slots = kHaloMaxResSlots;
            if (res_slots >= 2) {
                res_box = ncols;
                res_bytes = (uint32_t)res_slots * slot;
            } else {
                res_slots = 0;
            }
        }
        stage_out += res_bytes;       // same budget line below; split again when the plan is recorded
        const long avail = (long)kHaloMaxDynSmem - 1024 - (long)b_region - (long)stage_out;
        if (avail < 2L * a_stage) continue;
        int stages = (int)(avail / a_stage);
        // two CTAs per SM when everything fits twice: more epilogue warps per SM for the thin-K layers
        // thin tiles: several output tiles per accumulator stage (one tmem_full / tmem_empty hand-off per stage) when
        // every CTA has many tiles to walk; the per-tile hand-off chain, not bandwidth, bounded those layers
        int supert = 1;
        bool thin = false;
        {
            const long spatial_t = (long)a->B * ceil_div(a->Wo, kHaloTW) * ceil_div(a->Ho, kHaloTH);
            const long per_cta = spatial_t / ((long)sm_count() * 2 / gs + 1);
            // (measured: pays only for the thinnest tiles — the K = 64 stem conv, 4 MMAs per tile: 188 -> 170 us;
            //  neutral to slightly negative once a tile carries >= 9 MMAs)
            const int mmas_per_tile = p.taps * (cin_cta / 16);
            thin = thin_ok(a, ncols, gcta, cout_g, store_bw);
            const int smax = thin ? 4 : (mmas_per_tile > 8 ? 1 : (ncols <= 64 ? 2 : 1));
            while (supert * 2 <= smax && per_cta >= 4L * supert * 2 && !env_flag("SPECYOLO_NO_SUPERTILE")) supert *= 2;
        }
        int occ = 1;
        uint32_t cols = 32;
        while (cols < 2u * (uint32_t)(ncols * supert)) cols <<= 1;
        // (228 KB per SM, ~5.5 KB of static + reserved shared memory per CTA: two CTAs fit with <= 108 KB dynamic each)
        const long half = 108L * 1024 - 1024 - (long)b_region - (long)stage_out;
        if (cols <= 256 && half >= 2L * a_stage) {
            occ = 2;
            stages = (int)(half / a_stage);
        }
        if (stages > kHaloMaxStages) stages = kHaloMaxStages;
        if (stages <= best_stages) continue;
        best_stages = stages;
        found = true;
        p.gsplit = gs;
        p.supert = supert;
        p.thin = thin ? 1 : 0;
        p.groups_cta = gcta;
        p.cin_cta = cin_cta;
        p.kc_box = kc_box;
        p.boxes = cin_cta / kc_box;
        p.ncols = ncols;
        p.a_row_bytes = a_row;
        p.a_tx_bytes = a_tx;
        p.a_stage_bytes = a_stage;
        p.b_total_bytes = b_total;
        p.b_region_bytes = b_region;
        p.stages = stages;
        p.tmem_cols = cols;
        plan.occ = occ;
        plan.smem_bytes = 1024 + (size_t)b_region + (size_t)stages * a_stage + stage_out;
        p.store_bw = store_bw;
        p.pair_stores = (ncols >= 64 && !env_flag("SPECYOLO_NO_PAIR")) ? 1 : 0;
        p.store_row_bytes = (uint32_t)(store_bw * es);
        p.store_swz_mask = p.store_row_bytes == 128 ? 7u : (p.store_row_bytes == 64 ? 3u : 1u);
        p.ring_bytes = (uint32_t)stages * a_stage;
        p.res_box_cols = res_box;
        p.res_row_bytes = (uint32_t)res_box * 2u;
        p.res_swz_mask = p.res_row_bytes == 128 ? 7u : (p.res_row_bytes == 64 ? 3u : 1u);
        p.res_slots = res_slots;
        p.res_slot_bytes = res_slots ? res_bytes / (uint32_t)res_slots : 0u;
        p.res_off = p.ring_bytes + (stage_out - res_bytes);
    }
    if (!found) return false;

    p.B = a->B; p.Ho = a->Ho; p.Wo = a->Wo;
    p.tiles_w = ceil_div(a->Wo, kHaloTW);
    p.tiles_h = ceil_div(a->Ho, kHaloTH);
    const long spatial = (long)a->B * p.tiles_w * p.tiles_h;
    if (spatial <= 0 || spatial >= (1L << 30)) return false;
    p.spatial_tiles = (int)spatial;
    p.nsuper = (int)((spatial + p.supert - 1) / p.supert);
    p.d_img = make_fastdiv((uint32_t)(p.tiles_w * p.tiles_h));
    p.d_tw = make_fastdiv((uint32_t)p.tiles_w);
    p.d_cin_g = make_fastdiv((uint32_t)cin_g);
    p.d_npad = make_fastdiv((uint32_t)p.n_pad);
    p.kcb_log2 = p.kc_b == 64 ? 6 : (p.kc_b == 32 ? 5 : 4);
    if (!fastdiv_ok((uint64_t)spatial, (uint32_t)(p.tiles_w * p.tiles_h)) || p.taps > 49) return false;
    for (int t = 0; t < p.taps; ++t)
        p.tap_a16[t] = ((uint32_t)(((t / p.kw) * p.halo_w + (t % p.kw)) * p.dil) * p.a_row_bytes) >> 4;
    p.bias = a->bias;
    p.y = a->y; p.y_pixstride = a->y_pixstride; p.y_fp32 = a->y_fp32;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual);
    p.r_pixstride = a->r_pixstride;
    p.act = a->act;
    p.y_s2d = a->y_s2d;
    if (a->y_s2d && (groups != 1 || a->residual)) return false;
    plan.p = p;
    const long resident = (long)sm_count() * plan.occ / p.gsplit;           // CTAs per split
    const long per_split = p.nsuper < resident ? p.nsuper : (resident < 1 ? 1 : resident);
    plan.grid = (unsigned)(per_split * p.gsplit);
    return true;
}

// Returns SPECYOLO_OK after launching, or -1 when the conv is not eligible (caller falls through to the per-tap kernel).
int conv_halo_try_launch(const specyolo_conv_t* a, cudaStream_t stream) {
    HaloPlan plan{};
    if (!conv_halo_plan(a, plan)) return -1;
    EncodeTiledFn encode = get_encode_fn();
    SY_CHECK(encode != nullptr, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const HaloParams& p = plan.p;
    const int halo_h = kHaloTH + (a->kh - 1) * p.dil;

    CUtensorMap map_a, map_b;
    {
        const cuuint64_t pix_b = (cuuint64_t)a->x_pixstride * 2;
        cuuint64_t dims[4] = {(cuuint64_t)a->Cin, (cuuint64_t)a->W, (cuuint64_t)a->H, (cuuint64_t)a->B};
        cuuint64_t strides[3] = {pix_b, pix_b * a->W, pix_b * a->W * a->H};
        cuuint32_t box[4] = {(cuuint32_t)p.kc_box, (cuuint32_t)(p.halo_w * p.sub), (cuuint32_t)(halo_h * p.sub), 1};
        cuuint32_t estr[4] = {1, (cuuint32_t)p.sub, (cuuint32_t)p.sub, 1};
        CUresult r = encode(&map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a->x), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for((int)p.a_row_bytes),
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SY_CHECK(r == CUDA_SUCCESS, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled(halo A) failed (%d): box %u,%u,%u", (int)r,
                 box[0], box[1], box[2]);
    }
    {
        const cuuint64_t ktot = (cuuint64_t)p.taps * p.cin_g;
        cuuint64_t dims[2] = {ktot, (cuuint64_t)a->groups * a->n_pad};
        cuuint64_t strides[1] = {ktot * 2};
        cuuint32_t box[2] = {(cuuint32_t)p.kc_b, (cuuint32_t)p.n_pad};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a->w_packed), dims, strides,
                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for((int)p.b_row_bytes),
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SY_CHECK(r == CUDA_SUCCESS, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled(halo B) failed (%d)", (int)r);
    }
    CUtensorMap map_y = map_b;      // placeholder when the direct-store epilogue is used
    if (p.store_bw) {
        const int es = a->y_fp32 ? 4 : 2;
        const cuuint64_t pix_b = (cuuint64_t)a->y_pixstride * es;
        cuuint64_t dims[4] = {(cuuint64_t)a->Cout, (cuuint64_t)a->Wo, (cuuint64_t)a->Ho, (cuuint64_t)a->B};
        cuuint64_t strides[3] = {pix_b, pix_b * a->Wo, pix_b * a->Wo * a->Ho};
        cuuint32_t box[4] = {(cuuint32_t)p.store_bw, (cuuint32_t)kHaloTW, (cuuint32_t)kHaloTH, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&map_y, a->y_fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a->y,
                            dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for((int)p.store_row_bytes),
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SY_CHECK(r == CUDA_SUCCESS, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled(halo Y) failed (%d)", (int)r);
    }
    CUtensorMap map_r = map_b;      // placeholder when the residual is read from global memory
    if (p.res_box_cols) {
        const cuuint64_t pix_b = (cuuint64_t)a->r_pixstride * 2;
        cuuint64_t dims[4] = {(cuuint64_t)a->Cout, (cuuint64_t)a->Wo, (cuuint64_t)a->Ho, (cuuint64_t)a->B};
        cuuint64_t strides[3] = {pix_b, pix_b * a->Wo, pix_b * a->Wo * a->Ho};
        cuuint32_t box[4] = {(cuuint32_t)p.res_box_cols, (cuuint32_t)kHaloTW, (cuuint32_t)kHaloTH, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&map_r, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(a->residual), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for((int)p.res_row_bytes),
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SY_CHECK(r == CUDA_SUCCESS, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled(halo residual) failed (%d)", (int)r);
    }
    typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const HaloParams);
    static const KernelFn kernels[8] = {
        conv_halo_kernel<false, false, false>, conv_halo_kernel<false, false, true>,
        conv_halo_kernel<false, true, false>,  conv_halo_kernel<false, true, true>,
        conv_halo_kernel<true, false, false>,  conv_halo_kernel<true, false, true>,
        conv_halo_kernel<true, true, false>,   conv_halo_kernel<true, true, true>};
    static std::once_flag attr_once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(attr_once, [] {
        for (KernelFn k : kernels) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloMaxDynSmem);
            if (e != cudaSuccess) attr_err = e;
        }
    });
    SY_CHECK(attr_err == cudaSuccess, SPECYOLO_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err));
    const KernelFn kernel = kernels[(a->act == SPECYOLO_ACT_SILU ? 4 : 0) + (a->residual ? 2 : 0) + (a->y_fp32 ? 1 : 0)];
    SY_CUDA(launch_pdl(kernel, dim3(plan.grid), dim3(kHaloThreads), plan.smem_bytes, stream, map_a, map_b, map_y, map_r, p));
    SY_LAUNCH_CHECK();
    count_launch();
    return SPECYOLO_OK;
}

}  // namespace specyolo
