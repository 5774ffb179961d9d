// Epilogue shared by the two implicit-GEMM conv kernels: TMEM accumulator -> +bias -> SiLU -> +residual -> bf16 / fp32
// NHWC store at a channel offset of the destination buffer (concat buffers are written in place).
//
// Eight epilogue warps: a warp may only read the 32 TMEM lanes of its quadrant (warp id % 4), so two warps share
// a quadrant and split the tile's 16-column chunks between them.  With a single warp per scheduler the epilogue
// is a pure latency chain (nothing to switch to while a tcgen05.ld / MUFU / LDS is in flight) and, measured with
// ncu, it — not the MMAs, TMA or HBM — bounded every conv of the network; two warps per scheduler, vector bias
// loads and a 3-instruction SiLU bring it under the MMA time of the k x k layers.
//
// SiLU with one MUFU: x*sigmoid(x) = h + h*tanh(h), h = x/2.  The bias is staged in shared memory pre-multiplied
// by 1/2 (SiLU) so that h = fma(acc, 0.5, hb) and y = fma(h, tanh h, h): 2 FFMA + 1 MUFU per element.
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace specyolo {

static constexpr int kEpiWarps = 8;
static constexpr int kConvThreads = 64 + 32 * kEpiWarps;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue

struct EpiOut {
    void* y;
    int y_pixstride;
    const __nv_bfloat16* residual;
    int r_pixstride;
    bool pair;            // direct bf16 stores: lane pairs exchange halves to write full 32-byte sectors
};

// One 16-column chunk of one accumulator row.  `bias16` points at the 16 staged (pre-scaled) bias values of the
// chunk (16-byte aligned shared memory); `gch` is the first output channel of the chunk inside the output window;
// `nvalid` (1..16) the number of real channels in it.
// Lanes 2i and 2i+1 hold adjacent accumulator rows.  A 16-byte store per lane at a row stride >= 64 B touches 32
// different 32-byte sectors and half-fills each one; the partial-sector writes doubled the L1 -> L2 write traffic
// (ncu: l1tex2xbar write bytes = 2 x the tensor).  `pair` mode: the two lanes swap one 16-byte half so that each
// store instruction writes one row's full 32-byte sector from the lane pair.
struct EpiRow {
    size_t pix;        // own pixel index
    size_t pix_pair;   // the partner lane's pixel index
    bool ok, ok_pair;  // row inside the output
};

// bias + activation + residual of one 16-column chunk -> f[16]
template <bool kSilu, bool kRes>
__device__ __forceinline__ void epi_math16(const uint32_t (&v)[16], const float* bias16, const EpiOut& o,
                                           const EpiRow& row, int gch, int nvalid, const uint4& r0, const uint4& r1,
                                           bool res_vec, float (&f)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 b = *reinterpret_cast<const float4*>(bias16 + 4 * q);
        const float2 a0 = make_float2(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]));
        const float2 a1 = make_float2(__uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        float2 r0, r1;
        if (kSilu) {        // packed pairs: 2 FFMA2 + 2 MUFU per pair instead of 4 FFMA + 2 MUFU
            r0 = silu2_half(a0, make_float2(b.x, b.y));
            r1 = silu2_half(a1, make_float2(b.z, b.w));
        } else {
            r0 = fadd2(a0, make_float2(b.x, b.y));
            r1 = fadd2(a1, make_float2(b.z, b.w));
        }
        f[4 * q] = r0.x; f[4 * q + 1] = r0.y; f[4 * q + 2] = r1.x; f[4 * q + 3] = r1.y;
    }
    if (kRes) {
        if (res_vec) {
            const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float2 t = unpack_bf16x2(rr[i]);
                f[2 * i] += t.x;
                f[2 * i + 1] += t.y;
            }
        } else if (row.ok) {
            const __nv_bfloat16* rp = o.residual + row.pix * o.r_pixstride + gch;
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i < nvalid) f[i] += __bfloat162float(rp[i]);
        }
    }
}

// direct global stores of one chunk (fallback when the output window cannot be described by a TMA tensor map)
template <bool kFp32>
__device__ __forceinline__ void epi_store16_direct(const float (&f)[16], const EpiOut& o, const EpiRow& row, int gch,
                                                   int nvalid, bool vec_ok, int lane) {
    const size_t pix = row.pix;
    if (kFp32) {
        float* y = reinterpret_cast<float*>(o.y) + pix * o.y_pixstride + gch;
        if (!row.ok) {
        } else if (vec_ok && nvalid == 16) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                reinterpret_cast<float4*>(y)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i < nvalid) y[i] = f[i];
        }
    } else {
        __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(o.y);
        if (vec_ok && nvalid == 16 && !o.pair) {
            if (row.ok) {
                uint4 o0, o1;
                o0.x = pack_bf16x2(f[0], f[1]);   o0.y = pack_bf16x2(f[2], f[3]);
                o0.z = pack_bf16x2(f[4], f[5]);   o0.w = pack_bf16x2(f[6], f[7]);
                o1.x = pack_bf16x2(f[8], f[9]);   o1.y = pack_bf16x2(f[10], f[11]);
                o1.z = pack_bf16x2(f[12], f[13]); o1.w = pack_bf16x2(f[14], f[15]);
                uint4* y = reinterpret_cast<uint4*>(yb + pix * o.y_pixstride + gch);
                y[0] = o0;
                y[1] = o1;
            }
        } else if (vec_ok && nvalid == 16) {          // warp-uniform condition: the exchange below is convergent
            uint4 o0, o1;
            o0.x = pack_bf16x2(f[0], f[1]);   o0.y = pack_bf16x2(f[2], f[3]);
            o0.z = pack_bf16x2(f[4], f[5]);   o0.w = pack_bf16x2(f[6], f[7]);
            o1.x = pack_bf16x2(f[8], f[9]);   o1.y = pack_bf16x2(f[10], f[11]);
            o1.z = pack_bf16x2(f[12], f[13]); o1.w = pack_bf16x2(f[14], f[15]);
            const bool odd = lane & 1;
            // even lane gives its upper half and receives the odd lane's lower half (and vice versa)
            uint4 give = odd ? o0 : o1, got;
            got.x = __shfl_xor_sync(0xffffffffu, give.x, 1);
            got.y = __shfl_xor_sync(0xffffffffu, give.y, 1);
            got.z = __shfl_xor_sync(0xffffffffu, give.z, 1);
            got.w = __shfl_xor_sync(0xffffffffu, give.w, 1);
            const size_t pix_e = odd ? row.pix_pair : row.pix, pix_o = odd ? row.pix : row.pix_pair;
            const bool ok_e = odd ? row.ok_pair : row.ok, ok_o = odd ? row.ok : row.ok_pair;
            const int half_off = odd ? 8 : 0;
            // store A: the even row's sector (even lane: own lower half, odd lane: the even lane's upper half)
            if (ok_e) *reinterpret_cast<uint4*>(yb + pix_e * o.y_pixstride + gch + half_off) = odd ? got : o0;
            // store B: the odd row's sector
            if (ok_o) *reinterpret_cast<uint4*>(yb + pix_o * o.y_pixstride + gch + half_off) = odd ? o1 : got;
        } else if (row.ok) {
            __nv_bfloat16* y = yb + pix * o.y_pixstride + gch;
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i < nvalid) y[i] = __float2bfloat16_rn(f[i]);
        }
    }
}

// TMA-store staging: the output tile goes to shared memory in the swizzled [row][bw columns] box layout and one
// thread per half issues cp.async.bulk.tensor stores (full 128-byte lines, clipping at the tensor edges done by
// the TMA unit).  Per-lane 16-byte global stores at a row stride cost one LSU packet per half-filled sector and,
// measured, bounded every HBM-bound conv at ~4.4 TB/s of combined traffic.
struct EpiStage {
    bool enabled;
    uint8_t* buf;             // this half's staging buffer: 128 rows x row_bytes, 1024-byte aligned
    const CUtensorMap* map_y;
    int bw;                   // columns per store box
    uint32_t row_bytes;       // bw * element size: 32 / 64 / 128
    uint32_t swz_mask;        // 1 / 3 / 7: 16-byte unit index ^= (offset >> 7) & mask
    bool joint;               // tile rows are <= 128 bytes: ONE box per tile filled by both halves (all 8 warps);
                              // otherwise every half stores its own 64-column boxes
    int bar_id, bar_threads;  // named barrier of the warps sharing a box
    bool issuer;              // this thread issues the stores of its box
    int m;                    // this thread's row inside the tile
    int c0, c1, c2, c3;       // TMA coordinates of the tile's first element; c0 = channel of accumulator column 0
};

template <bool kFp32>
__device__ __forceinline__ void epi_store16_stage(const float (&f)[16], const EpiStage& st, int col_in_box) {
    const uint32_t row_off = (uint32_t)st.m * st.row_bytes;
    if (kFp32) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            uint32_t off = row_off + (uint32_t)(col_in_box * 4 + u * 16);
            off ^= ((off >> 7) & st.swz_mask) << 4;
            *reinterpret_cast<float4*>(st.buf + off) = make_float4(f[4 * u], f[4 * u + 1], f[4 * u + 2], f[4 * u + 3]);
        }
    } else {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            uint32_t off = row_off + (uint32_t)(col_in_box * 2 + u * 16);
            off ^= ((off >> 7) & st.swz_mask) << 4;
            uint4 o4;
            o4.x = pack_bf16x2(f[8 * u + 0], f[8 * u + 1]); o4.y = pack_bf16x2(f[8 * u + 2], f[8 * u + 3]);
            o4.z = pack_bf16x2(f[8 * u + 4], f[8 * u + 5]); o4.w = pack_bf16x2(f[8 * u + 6], f[8 * u + 7]);
            *reinterpret_cast<uint4*>(st.buf + off) = o4;
        }
    }
}

// Column -> channel mapping of an accumulator tile made of `n_pad`-wide groups (the per-tap kernel's tile is one
// slice [ch_base, ch_base + ncols) of a single group: pass within0 = ch_base and a group width >= ncols).
struct EpiCols {
    int ncols;        // accumulator columns of the tile (multiple of 16)
    int n_pad;        // columns per group
    FastDiv d_npad;
    int cout_g;       // real channels per group
    int within0;      // first column's index inside its group
    int gch0;         // output-window channel of (group 0 of the tile, index 0)
};

// L2 prefetch of the residual rows of the NEXT tile of this warp (no registers held): when the epilogue is the slowest
// stage there is no accumulator wait to hide behind, so the next tile's residual is pulled into L2 one tile ahead
// and the register loads above then see L2 latency instead of DRAM latency.
template <bool kRes>
__device__ __forceinline__ void epi_prefetch_residual_l2(const EpiCols& ec, const EpiOut& o, size_t pix, bool row_ok,
                                                         int half) {
    if (!kRes || !row_ok) return;
    const int nchunks = ec.ncols >> 4;
    const int cut = (nchunks + 1) >> 1;
    const int ch_begin = half ? cut : 0, ch_end = half ? nchunks : cut;
    const __nv_bfloat16* rp = o.residual + pix * o.r_pixstride + ec.gch0 + ec.within0;
    for (int chunk = ch_begin; chunk < ch_end; ++chunk)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + (chunk << 4)));
}

// Residual tile staged in shared memory by the TMA producer (halo kernel, tiles <= 64 columns): [boxes][128 rows][box
// row], swizzled like every TMA box.  nullptr = residual read from global memory (paths above).
struct EpiResSmem {
    const uint8_t* base;      // this tile's slot, or nullptr
    int box_cols;             // columns per box (power of two, <= 64)
    uint32_t row_bytes;       // box_cols * 2
    uint32_t swz_mask;
    int m;                    // this thread's row
};
__device__ __forceinline__ void epi_res_smem_load(const EpiResSmem& rs, int c, uint4& r0, uint4& r1) {
    const int box = c / rs.box_cols;           // box_cols is a power of two
    const int col = c - box * rs.box_cols;
    const uint32_t box_off = (uint32_t)box * 128u * rs.row_bytes;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        uint32_t off = (uint32_t)rs.m * rs.row_bytes + (uint32_t)(col * 2 + u * 16);
        off ^= ((off >> 7) & rs.swz_mask) << 4;
        const uint4 t = *reinterpret_cast<const uint4*>(rs.base + box_off + off);
        if (u == 0) r0 = t; else r1 = t;
    }
}

// All rows/columns of one accumulator buffer handled by this warp.  t_addr = TMEM address of (this warp's lane
// quadrant, first column of the buffer).  `half` in {0,1}: which of the two warps sharing the quadrant this is.
template <bool kSilu, bool kRes, bool kFp32>
__device__ __forceinline__ void epi_tile(uint32_t t_addr, const EpiCols& ec, const float* bias_s, const EpiOut& o,
                                         size_t pix, bool row_ok, int half, int lane, const EpiStage& st,
                                         const EpiResSmem& rs) {
    EpiRow row;
    row.pix = pix;
    row.ok = row_ok;
    row.pix_pair = ((size_t)__shfl_xor_sync(0xffffffffu, (uint32_t)(pix >> 32), 1) << 32) |
                   (size_t)__shfl_xor_sync(0xffffffffu, (uint32_t)pix, 1);
    row.ok_pair = __shfl_xor_sync(0xffffffffu, (int)row_ok, 1) != 0;
    const int nchunks = ec.ncols >> 4;
    const int cut = (nchunks + 1) >> 1;
    const int ch_begin = half ? cut : 0, ch_end = half ? nchunks : cut;
    if (ch_begin >= ch_end) return;
    const bool y_vec = ((reinterpret_cast<uintptr_t>(o.y) & 15) == 0) && (o.y_pixstride % (kFp32 ? 4 : 8) == 0);
    const bool r_vec = kRes && ((reinterpret_cast<uintptr_t>(o.residual) & 15) == 0) && (o.r_pixstride % 8 == 0);
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
            const int c = chunk << 4;
            uint32_t gl, within;
            fdivmod((uint32_t)(ec.within0 + c), ec.d_npad, gl, within);
            const int nvalid = min(16, ec.cout_g - (int)within);
            const int gch = ec.gch0 + (int)gl * ec.cout_g + (int)within;
            const int align = kFp32 ? 3 : 7;
            const bool vec_ok = y_vec && ((gch & align) == 0);
            uint4 r0 = make_uint4(0, 0, 0, 0), r1 = make_uint4(0, 0, 0, 0);
            const bool res_vec = kRes && ((rs.base != nullptr) || (r_vec && row_ok && nvalid == 16 && ((gch & 7) == 0)));
            if (kRes && rs.base != nullptr) {
                epi_res_smem_load(rs, c, r0, r1);      // TMA-staged tile: rows / columns outside the tensor are zeros
            } else if (res_vec) {       // global residual (L2-prefetched one tile ahead): issue before waiting on TMEM
                const uint4* rp = reinterpret_cast<const uint4*>(o.residual + pix * o.r_pixstride + gch);
                r0 = __ldg(rp);
                r1 = __ldg(rp + 1);
            }
            ptx::tmem_ld_wait();
            if (chunk + 1 < ch_end) ptx::tmem_ld16(t_addr + (uint32_t)(c + 16), vn);   // prefetch the next chunk
            float f[16];
            if (st.enabled) {
                // box-by-box: wait until the previous store of this half has drained the staging buffer, fill it,
                // make the writes visible to the async proxy, let one thread issue the store
                // joint: the box is the whole tile row; per-half: boxes of bw columns inside this half's range
                const int col_in_box = st.joint ? c : ((c - (ch_begin << 4)) & (st.bw - 1));
                const bool first = st.joint ? (chunk == ch_begin) : (col_in_box == 0);
                const bool last = st.joint ? (chunk + 1 == ch_end) : (col_in_box + 16 == st.bw);
                if (first) {
                    if (st.issuer) ptx::bulk_wait_read0();
                    ptx::named_bar_sync(st.bar_id, st.bar_threads);
                }
                epi_math16<kSilu, kRes>(v, bias_s + c, o, row, gch, max(nvalid, 0), r0, r1, res_vec, f);
                epi_store16_stage<kFp32>(f, st, col_in_box);
                if (last) {
                    ptx::fence_proxy_async();
                    ptx::named_bar_sync(st.bar_id, st.bar_threads);
                    if (st.issuer) {
                        ptx::tma_store_4d(st.map_y, st.buf, st.joint ? st.c0 : st.c0 + c + 16 - st.bw, st.c1, st.c2, st.c3);
                        ptx::bulk_commit_group();
                    }
                }
                continue;
            }
            if (nvalid <= 0) continue;         // warp-uniform (padding columns of the last group chunk)
            epi_math16<kSilu, kRes>(v, bias_s + c, o, row, gch, nvalid, r0, r1, res_vec, f);
            epi_store16_direct<kFp32>(f, o, row, gch, nvalid, vec_ok, lane);
        }
    }
}

// staging bytes per CTA
static inline uint32_t epi_stage_bytes(int ncols, int store_bw, int elem_bytes) {
    if (!store_bw) return 0;
    return (store_bw == ncols ? 1u : 2u) * 128u * (uint32_t)(store_bw * elem_bytes);
}

// per-thread staging descriptor of an epilogue thread (warp 2..9, `half` = (warp - 2) / 4)
__device__ __forceinline__ EpiStage epi_make_stage(uint8_t* stage_base, const CUtensorMap* map_y, int ncols, int store_bw,
                                                   uint32_t row_bytes, uint32_t swz_mask, int warp, int half, int lane,
                                                   int m) {
    EpiStage st;
    st.enabled = store_bw != 0;
    st.joint = store_bw == ncols;
    st.buf = stage_base + (st.joint ? 0u : (uint32_t)half * 128u * row_bytes);
    st.map_y = map_y;
    st.bw = store_bw;
    st.row_bytes = row_bytes;
    st.swz_mask = swz_mask;
    st.bar_id = st.joint ? 1 : 1 + half;
    st.bar_threads = (st.joint && ncols > 16) ? 256 : 128;
    st.issuer = (warp == (st.joint ? 2 : 2 + 4 * half)) && lane == 0;
    st.m = m;
    st.c0 = st.c1 = st.c2 = st.c3 = 0;
    return st;
}

// Box width (columns): the whole tile row when it is at most 128 bytes (joint mode), else the width the two halves
// of an ncols-wide tile can both be cut into; 0 = TMA store not applicable.
static inline int epi_stage_box_cols(int ncols, int elem_bytes) {
    // thin tiles (<= 32 columns) are dominated by per-tile fixed costs: the two named barriers + store issue of the
    // staged path cost more than the LSU packets they save (measured: 16->32 k3 95 us direct vs 172 us staged)
    if (ncols * elem_bytes < 128) return 0;
    if (ncols * elem_bytes <= 128 && (ncols & (ncols - 1)) == 0) return ncols;    // swizzle spans: 32 / 64 / 128 bytes
    const int nchunks = ncols >> 4;
    const int h0 = ((nchunks + 1) >> 1) * 16, h1 = (nchunks >> 1) * 16;
    for (int bw = 128 / elem_bytes; bw >= 16; bw >>= 1)
        if (h0 % bw == 0 && h1 % bw == 0) return bw;
    return 0;
}

}  // namespace specyolo
