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
};

// One 16-column chunk of one accumulator row.  `bias16` points at the 16 staged (pre-scaled) bias values of the
// chunk (16-byte aligned shared memory); `gch` is the first output channel of the chunk inside the output window;
// `nvalid` (1..16) the number of real channels in it.
template <bool kSilu, bool kRes, bool kFp32>
__device__ __forceinline__ void epi_chunk16(const uint32_t (&v)[16], const float* bias16, const EpiOut& o, size_t pix,
                                            int gch, int nvalid, bool vec_ok, const uint4& r0, const uint4& r1,
                                            bool res_vec) {
    float f[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float4 b = *reinterpret_cast<const float4*>(bias16 + 4 * q);
        const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float a = __uint_as_float(v[4 * q + i]);
            if (kSilu) {
                const float h = fmaf(a, 0.5f, bb[i]);
                float t;
                asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
                f[4 * q + i] = fmaf(h, t, h);
            } else {
                f[4 * q + i] = a + bb[i];
            }
        }
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
        } else {
            const __nv_bfloat16* rp = o.residual + pix * o.r_pixstride + gch;
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i < nvalid) f[i] += __bfloat162float(rp[i]);
        }
    }
    if (kFp32) {
        float* y = reinterpret_cast<float*>(o.y) + pix * o.y_pixstride + gch;
        if (vec_ok && nvalid == 16) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                reinterpret_cast<float4*>(y)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i < nvalid) y[i] = f[i];
        }
    } else {
        __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(o.y) + pix * o.y_pixstride + gch;
        if (vec_ok && nvalid == 16) {
            uint4 o0, o1;
            o0.x = pack_bf16x2(f[0], f[1]);   o0.y = pack_bf16x2(f[2], f[3]);
            o0.z = pack_bf16x2(f[4], f[5]);   o0.w = pack_bf16x2(f[6], f[7]);
            o1.x = pack_bf16x2(f[8], f[9]);   o1.y = pack_bf16x2(f[10], f[11]);
            o1.z = pack_bf16x2(f[12], f[13]); o1.w = pack_bf16x2(f[14], f[15]);
            reinterpret_cast<uint4*>(y)[0] = o0;
            reinterpret_cast<uint4*>(y)[1] = o1;
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i < nvalid) y[i] = __float2bfloat16_rn(f[i]);
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

// All rows/columns of one accumulator buffer handled by this warp.  t_addr = TMEM address of (this warp's lane
// quadrant, first column of the buffer).  `half` in {0,1}: which of the two warps sharing the quadrant this is.
template <bool kSilu, bool kRes, bool kFp32>
__device__ __forceinline__ void epi_tile(uint32_t t_addr, const EpiCols& ec, const float* bias_s, const EpiOut& o,
                                         size_t pix, bool row_ok, int half) {
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
            const bool res_vec = kRes && r_vec && row_ok && nvalid == 16 && ((gch & 7) == 0);
            if (res_vec) {      // issue the residual loads before waiting on TMEM
                const uint4* rp = reinterpret_cast<const uint4*>(o.residual + pix * o.r_pixstride + gch);
                r0 = __ldg(rp);
                r1 = __ldg(rp + 1);
            }
            ptx::tmem_ld_wait();
            if (chunk + 1 < ch_end) ptx::tmem_ld16(t_addr + (uint32_t)(c + 16), vn);   // prefetch the next chunk
            if (!row_ok || nvalid <= 0) continue;
            epi_chunk16<kSilu, kRes, kFp32>(v, bias_s + c, o, pix, gch, nvalid, vec_ok, r0, r1, res_vec);
        }
    }
}

}  // namespace specyolo
