// Fused stem for sm_100a: uint8 NCHW image -> Conv(3, c0, 3, 2)+BN+SiLU -> Conv(c0, c1, 3, 2)+BN+SiLU -> bf16 NHWC
// (layers 0 and 1 of every YOLOv11 trunk, ultralytics/nn/tasks.py parse_model / cfg yolo11*.yaml backbone[0:2];
// the /255 of predictor.py:133-135 is folded into the layer-0 weights).
//
// Run layer by layer the stem moves 1.55 GB per 64-image batch (space-to-depth copy, the 320x320x32 layer-0 map
// written and re-read) for 0.29 GB of unavoidable traffic (image in, layer-1 map out) — 17 % of the DRAM bytes of the
// whole forward pass.  Here the layer-0 activations never leave the SM.  Per tile of 16 x 8 layer-1 pixels:
//   warps 2-9   the (69 x 37 pixel x 3 plane) uint8 image patch of the NEXT tile -> shared memory (cp.async, zero fill
//               outside the image = the conv padding), then im2col: 612 layer-0 pixels x 27 taps -> fp16 A0 [640 x 32] (K-major, SWIZZLE_64B)
//   warp 1      MMA 1 (tcgen05, 5 x M=128, N=c0, K=32) -> TMEM;  MMA 2 below
//   warps 10-17 epilogue 1: TMEM -> bias + SiLU -> bf16 -> A1 in shared memory, laid out as the 2x2-blocked
//               (space-to-depth) halo tile [17 x 9 blocked pixels][4*c0 channels] the layer-1 GEMM consumes
//               (K-major SWIZZLE_128B 64-channel chunks; blocked pixels outside the image are written as zeros =
//               layer 1's padding)
//   warp 1      MMA 2: layer 1 as a 2x2 / stride-1 conv over the blocked tile (taps = descriptor row shifts, like
//               conv_halo.cu), resident weights, M=128, N=c1, K=16*c0 -> TMEM (double-buffered)
//   warps 18-25 epilogue 2 (epilogue.cuh): bias + SiLU -> bf16 -> staged TMA store
// A1 and the layer-1 accumulator are double-buffered so that tile t+1's im2col / MMA 1 / epilogue 1 overlap tile t's
// MMA 2 / epilogue 2.
#include "common.h"
#include "ptx.cuh"
#include "tma_host.h"
#include "epilogue.cuh"

namespace specyolo {

static constexpr int kSpThreads = 64 + 256 + 256 + 256;     // loader + MMA, im2col, epilogue 1, epilogue 2
static constexpr int kSpTW = 8, kSpTH = 16, kSpHW = 9, kSpHH = 17;       // blocked-pixel tile / halo (taps at -1, 0)
static constexpr int kSpHalo = kSpHW * kSpHH;                             // 153 blocked halo pixels
static constexpr int kSpRows0 = kSpHalo * 4;                              // 612 layer-0 pixels per tile
static constexpr int kSpMTiles = 5;                                       // ceil(612 / 128)
static constexpr int kSpPatchW = 48, kSpPatchH = 69;                      // patch rows of 48 bytes (bytes 3..39 used)
static constexpr int kSpPatchChunks = (3 * kSpPatchH * 6 + 255) / 256;    // 8-byte chunks per im2col thread: 5
static constexpr uint32_t kSpPatchStride = 10240;
static constexpr uint32_t kSpA0Bytes = kSpMTiles * 128 * 64;              // 40 KB: 640 rows x 32 K (bf16)
static constexpr uint32_t kSpA1Chunk = 20480;                             // 153 rows x 128 B, rounded up to 1024
static constexpr int kSpMaxDynSmem = 225 * 1024;

__device__ __forceinline__ void cp_async8_zfill(uint32_t smem_dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_dst), "l"(src), "r"(src_bytes) : "memory");
}
// the mbarrier receives one arrival from this thread once all of its earlier cp.async copies have landed
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(ptx::smem_u32(bar)) : "memory");
}

struct StemPairParams {
    int B, H1, W1;                  // layer-1 output size (H/4, W/4)
    int tiles_w, tiles_h, spatial_tiles;
    FastDiv d_img, d_tw;
    int c0;                         // layer-0 channels: 16 or 32
    int chunks1;                    // 64-channel chunks of the blocked layer-0 tile: 4*c0 / 64
    int n_pad, cout;                // layer-1 accumulator columns / real channels
    const float* b0;
    const float* b1;
    const void* x;                  // uint8 NCHW image
    const void* w0;                 // fp16 [c0][32]: k = c*9 + ky*3 + kx, columns 27..31 zero
    uint32_t w1_bytes, a1_buf_bytes;
    uint32_t off_w0, off_patch, off_a0, off_a1, off_st;
    uint32_t tmem_cols;
    void* y;
    int y_pixstride;
    int store_bw, pair_stores;
    uint32_t store_row_bytes, store_swz_mask;
};

__global__ void __launch_bounds__(kSpThreads, 1)
stem_pair_kernel(const __grid_constant__ CUtensorMap map_w,
                 const __grid_constant__ CUtensorMap map_y, const __grid_constant__ StemPairParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t patch_full[2], patch_empty[2], a0_full, a0_empty, acc0_full, acc0_empty;
    __shared__ __align__(8) uint64_t a1_full[2], a1_empty[2], acc1_full[2], acc1_empty[2], w_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float bias_s[256];
    __shared__ __align__(16) float bias0_s[32];

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    ptx::grid_dep_launch();

    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* base = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* w1_s = base;                        // resident layer-1 weights: (tap, chunk) boxes of [n_pad rows x 128 B]
    uint8_t* w0_s = base + p.off_w0;             // layer-0 weights [c0 rows x 64 B], SWIZZLE_64B
    uint8_t* patch_s = base + p.off_patch;       // 2 image patches
    uint8_t* a0_s = base + p.off_a0;             // im2col tile
    uint8_t* a1_s = base + p.off_a1;             // 2 blocked layer-0 tiles x chunks1 x kSpA1Chunk
    uint8_t* st_buf = base + p.off_st;           // epilogue-2 staging

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&map_w);
        if (p.store_bw) ptx::prefetch_tmap(&map_y);
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&patch_full[s], 256);
            ptx::mbar_init(&patch_empty[s], 8);
            ptx::mbar_init(&a1_full[s], kEpiWarps);
            ptx::mbar_init(&a1_empty[s], 1);
            ptx::mbar_init(&acc1_full[s], 1);
            ptx::mbar_init(&acc1_empty[s], kEpiWarps);
        }
        ptx::mbar_init(&a0_full, 8);
        ptx::mbar_init(&a0_empty, 1);
        ptx::mbar_init(&acc0_full, 1);
        ptx::mbar_init(&acc0_empty, kEpiWarps);
        ptx::mbar_init(&w_bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc(&tmem_base_smem, p.tmem_cols);
    for (int i = threadIdx.x; i < p.n_pad; i += kSpThreads) bias_s[i] = 0.5f * p.b1[i];       // SiLU form (epilogue.cuh)
    if ((int)threadIdx.x < p.c0) bias0_s[threadIdx.x] = 0.5f * p.b0[threadIdx.x];
    // layer-0 weights -> shared memory in the swizzled K-major layout; zero the 28 padding rows of A0
    if ((int)threadIdx.x < p.c0 * 4) {
        const uint32_t n = threadIdx.x >> 2, u = threadIdx.x & 3;
        const uint4 v = reinterpret_cast<const uint4*>(p.w0)[threadIdx.x];
        *reinterpret_cast<uint4*>(w0_s + n * 64u + ((u ^ ((n >> 1) & 3u)) << 4)) = v;
    }
    for (uint32_t i = threadIdx.x; i < (uint32_t)(kSpMTiles * 128 - kSpRows0) * 4u; i += kSpThreads)
        *reinterpret_cast<uint4*>(a0_s + (uint32_t)kSpRows0 * 64u + i * 16u) = make_uint4(0, 0, 0, 0);
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    const uint32_t acc0_col = 2u * (uint32_t)p.n_pad;        // layer-1 accumulators first, then the 5 layer-0 tiles
    const int cta = blockIdx.x, ctas = gridDim.x;

    if (warp == 0) {
        // ===================== TMA: resident layer-1 weights =====================
        const bool leader = ptx::elect_one();
        if (leader) {
            ptx::mbar_expect_tx(&w_bar, p.w1_bytes);
            for (int i = 0; i < 4 * p.chunks1; ++i)
                ptx::tma_load_2d(w1_s + (size_t)i * p.n_pad * 128, &map_w, &w_bar, i * 64, 0);
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        ptx::grid_dep_wait();
        const bool leader = ptx::elect_one();
        // layer 0 runs fp16 x fp16 (A0 = 1024 + pixel, see the im2col warps): format fields 0
        const uint32_t idesc0 = (1u << 4) | (((uint32_t)p.c0 >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t idesc1 = ptx::umma_idesc_bf16(128, (uint32_t)p.n_pad);
        const uint32_t hi64 = (uint32_t)(ptx::umma_smem_desc(0, 64) >> 32);
        const uint32_t hi128 = (uint32_t)(ptx::umma_smem_desc(0, 128) >> 32);
        const uint32_t a1_hi = (uint32_t)(ptx::umma_desc(0, 0, (uint32_t)kSpHW * 128u, 128) >> 32);
        const uint32_t a0_16 = ptx::smem_u32(a0_s) >> 4, w0_16 = ptx::smem_u32(w0_s) >> 4;
        const uint32_t a1_16 = ptx::smem_u32(a1_s) >> 4, w1_16 = ptx::smem_u32(w1_s) >> 4;
        const uint32_t wbox16 = ((uint32_t)p.n_pad * 128u) >> 4;
        ptx::mbar_wait(&w_bar, 0);
        auto mma1 = [&](uint32_t t) {
            ptx::mbar_wait(&acc0_empty, (t & 1u) ^ 1u);
            ptx::mbar_wait(&a0_full, t & 1u);
            ptx::tc_fence_after();
#pragma unroll
            for (int mt = 0; mt < kSpMTiles; ++mt) {
                const uint32_t d = tmem_base + acc0_col + (uint32_t)(mt * p.c0);
                const uint32_t a = a0_16 + (uint32_t)(mt * 128 * 64 >> 4);
                if (leader) {
                    ptx::umma_bf16_lohi(d, a, hi64, w0_16, hi64, idesc0, 0u);
                    ptx::umma_bf16_lohi(d, a + 2u, hi64, w0_16 + 2u, hi64, idesc0, 1u);
                }
            }
            if (leader) {
                ptx::umma_commit(&a0_empty);
                ptx::umma_commit(&acc0_full);
            }
        };
        auto mma2 = [&](uint32_t t) {
            const uint32_t b = t & 1u, ph = (t >> 1) & 1u;
            ptx::mbar_wait(&acc1_empty[b], ph ^ 1u);
            ptx::mbar_wait(&a1_full[b], ph);
            ptx::tc_fence_after();
            const uint32_t d = tmem_base + b * (uint32_t)p.n_pad;
            uint32_t acc = 0;
            for (int c = 0; c < p.chunks1; ++c) {
                const uint32_t ac = a1_16 + ((b * p.a1_buf_bytes + (uint32_t)c * kSpA1Chunk) >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // the 16 channels of this K step belong to one position (dy, dx) of the 2x2 block; tap (ty, tx) of
                    // the blocked conv reads it only if 2 ty + dy - 1 and 2 tx + dx - 1 are taps of the 3x3 kernel:
                    // 18 of the 32 (tap, K step) products are non-zero, the others are skipped (the kernel is bound by
                    // shared-memory wavefronts, 48 per N = 64 MMA)
                    const int pos = (c * 64 + k * 16) / p.c0;
#pragma unroll
                    for (int tap = 0; tap < 4; ++tap) {
                        const bool live = ((tap >> 1) != 0 || (pos >> 1) != 0) && ((tap & 1) != 0 || (pos & 1) != 0);
                        const uint32_t shift = (uint32_t)(((tap >> 1) * kSpHW + (tap & 1)) * 128) >> 4;
                        const uint32_t wb = w1_16 + (uint32_t)(tap * p.chunks1 + c) * wbox16;
                        if (live) {
                            if (leader) ptx::umma_bf16_lohi(d, ac + shift + 2u * k, a1_hi, wb + 2u * k, hi128, idesc1, acc);
                            acc = 1;
                        }
                    }
                }
            }
            if (leader) {
                ptx::umma_commit(&a1_empty[b]);
                ptx::umma_commit(&acc1_full[b]);
            }
        };
        uint32_t tl = 0;
        if (cta < p.spatial_tiles) mma1(0);
        for (int tile = cta; tile < p.spatial_tiles; tile += ctas, ++tl) {
            if (tile + ctas < p.spatial_tiles) mma1(tl + 1);
            mma2(tl);
        }
    } else if (warp < 10) {
        // ===================== im2col: image patch -> A0 (warps 2..9) =====================
        ptx::grid_dep_wait();
        const int t = threadIdx.x - 64;                 // 0..255
        // Image patch of a tile: 3 planes x 69 rows of 48 bytes starting 8 pixels left of the tile's first layer-0 tap
        // (the first used byte is column 3 of a row), fetched ONE TILE AHEAD by these warps themselves as 8-byte
        // cp.async chunks (zero-filled outside the image = the conv padding; W % 16 == 0: an aligned chunk is entirely
        // inside or outside).  Every thread's copies signal patch_full through cp.async.mbarrier.arrive.noinc, so
        // nobody waits on memory latency.  (A single loader warp needed 4.3 k cycles per tile to issue the 1242
        // chunks and bounded the kernel; a TMA box over the uint8 planes starting at x = 32 tw - 5 never completed its
        // transaction — the innermost box offset is not a multiple of 16 bytes there; an aligned 64-byte-wide box is the
        // obvious next step.)
        const uint8_t* img = reinterpret_cast<const uint8_t*>(p.x);
        const int H = p.H1 * 4, W = p.W1 * 4;
        int c_off[kSpPatchChunks], c_rc[kSpPatchChunks];          // this thread's chunks: offset in the image, (row, col, plane)
#pragma unroll
        for (int j = 0; j < kSpPatchChunks; ++j) {
            const int idx = j * 256 + t;                          // plane * 414 + row * 6 + chunk
            const int c = idx / (kSpPatchH * 6), rem = idx - c * (kSpPatchH * 6);
            const int row = rem / 6, ck = rem - row * 6;
            c_off[j] = (c * H + row) * W + 8 * ck;
            c_rc[j] = idx < 3 * kSpPatchH * 6 ? (row | (ck << 8)) : -1;
        }
        auto load_patch = [&](int tile, uint32_t s) {
            uint32_t n, r, th_i, tw_i;
            fdivmod((uint32_t)tile, p.d_img, n, r);
            fdivmod(r, p.d_tw, th_i, tw_i);
            const int px0 = (int)tw_i * 32 - 8, py0 = (int)th_i * 64 - 5;
            const uint8_t* org = img + (size_t)n * 3 * (size_t)H * W;
            const long o0 = (long)py0 * W + px0;
            const uint32_t dst = ptx::smem_u32(patch_s + s * kSpPatchStride) + (uint32_t)t * 8u;
#pragma unroll
            for (int j = 0; j < kSpPatchChunks; ++j) {
                const int iy = py0 + (c_rc[j] & 0xff), ix = px0 + 8 * (c_rc[j] >> 8);
                const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
                if (c_rc[j] >= 0) cp_async8_zfill(dst + (uint32_t)(j * 2048), ok ? org + o0 + c_off[j] : org, ok ? 8u : 0u);
            }
            cp_async_mbar_arrive_noinc(&patch_full[s]);
        };
        if (cta < p.spatial_tiles) load_patch(cta, 0u);
        uint32_t tl = 0;
        for (int tile = cta; tile < p.spatial_tiles; tile += ctas, ++tl) {
            const uint32_t s = tl & 1u;
            if (tile + ctas < p.spatial_tiles) {
                // the other buffer was read by tile tl - 1: wait until all eight warps are done with it
                ptx::mbar_wait(&patch_empty[s ^ 1u], (((tl + 1) >> 1) & 1u) ^ 1u);
                load_patch(tile + ctas, s ^ 1u);
            }
            ptx::mbar_wait(&patch_full[s], (tl >> 1) & 1u);
            ptx::mbar_wait(&a0_empty, (tl & 1u) ^ 1u);          // MMA 1 of the previous tile has read A0
            const uint8_t* ps = patch_s + s * kSpPatchStride;
#pragma unroll 1
            for (int r = t; r < kSpRows0; r += 256) {
                // row r = (blocked halo pixel hp, position pos in its 2x2 block)
                const int hp = r >> 2, pos = r & 3;
                const int hy = hp / kSpHW, hx = hp - hy * kSpHW;
                const int ly = 4 * hy + 2 * (pos >> 1);           // patch row of tap ky = 0
                const int dx = pos & 1;
                // A0 holds fp16 values 1024 + pixel: bits 0x6400 | byte, i.e. one byte permute per PAIR of taps and no
                // int -> float conversion at all (the constant 1024 * sum(w) is folded into the layer-0 bias by the
                // caller; fp32 accumulation keeps the cancellation exact to ~1e-5).  v1 converted every tap through
                // fp32 -> bf16: 170 instructions per row, and the kernel as a whole was issue-bound.
                uint32_t q[9];                                    // [c*3 + ky]: bytes (kx0, kx1, kx2, 0x64)
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const uint32_t* row = reinterpret_cast<const uint32_t*>(ps + (c * kSpPatchH + ly + ky) * kSpPatchW + 4 * hx);
                        // taps kx = 0..2 are bytes 3 + 2 dx + kx of the 8 bytes at (patch row, 4 hx)
                        const uint32_t x3 = dx ? (row[1] >> 8) : __funnelshift_r(row[0], row[1], 24);
                        q[c * 3 + ky] = (x3 & 0x00ffffffu) | 0x64000000u;
                    }
                uint32_t wds[16];                                 // K slot k = c*9 + ky*3 + kx, two per word
#pragma unroll
                for (int i = 0; i < 14; ++i) {
                    const int k0 = 2 * i, k1 = 2 * i + 1;
                    const uint32_t sel = (uint32_t)(k0 % 3) | (3u << 4) | ((k1 < 27 ? 4u + (uint32_t)(k1 % 3) : 7u) << 8) | (7u << 12);
                    wds[i] = __byte_perm(q[k0 / 3], q[k1 < 27 ? k1 / 3 : 8], sel);
                }
                wds[14] = wds[15] = 0u;
                const uint32_t swz = ((uint32_t)r >> 1) & 3u;
                uint8_t* dst = a0_s + (uint32_t)r * 64u;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    *reinterpret_cast<uint4*>(dst + (((uint32_t)u ^ swz) << 4)) =
                        make_uint4(wds[4 * u], wds[4 * u + 1], wds[4 * u + 2], wds[4 * u + 3]);
            }
            ptx::fence_proxy_async();                   // generic-proxy writes of A0 -> visible to UMMA
            __syncwarp();
            if (lane == 0) {
                ptx::mbar_arrive(&a0_full);
                ptx::mbar_arrive(&patch_empty[s]);
            }
        }
    } else {
        // ===================== epilogues: 1 = warps 10..17, 2 = warps 18..25 =====================
        ptx::grid_dep_wait();
        const int quad = warp & 3;
        const int half = ((warp - 10) >> 2) & 1;        // which of the two warps of a lane quadrant inside its group
        const int m = quad * 32 + lane;
        const int tw = m & (kSpTW - 1), th = m >> 3;
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
        const int cchunks = p.c0 >> 4;                  // 16-column chunks of a layer-0 accumulator tile: 1 or 2
        const uint32_t t_quad = tmem_base + ((uint32_t)(quad * 32) << 16);

        if (warp < 18) {
            // ---- epilogue 1: layer-0 accumulators -> bias + SiLU -> bf16 -> blocked A1 tile ----
            // tasks (m-tile, 16-column chunk) of this lane quadrant alternate between its two warps; up to three
            // accumulator loads are in flight per wait (a single load -> wait -> math -> store chain per task left
            // the warps latency-bound: clock64 phase timers, 460 cycles per task)
            const int ntask = kSpMTiles * cchunks;
            // per-task constants of this thread (at most 5 tasks): A1 byte offset of its two 16-byte units, halo flags
            constexpr int kMaxT = kSpMTiles;
            uint32_t t_off[kMaxT], t_col[kMaxT];
            int t_flags[kMaxT];               // bit 0: halo row 0, bit 1: halo column 0, bit 2: row exists
            // every task of a warp has the same 16-column chunk (task parity == half): its half-scaled bias lives in
            // registers (a float4 shared-memory load per 4 columns cost 4 wavefronts each in a kernel that is bound
            // by shared-memory wavefronts)
            float2 hb[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int cc0 = cchunks == 2 ? half : 0;
                hb[i] = make_float2(bias0_s[cc0 * 16 + 2 * i], bias0_s[cc0 * 16 + 2 * i + 1]);
            }
#pragma unroll
            for (int k = 0; k < kMaxT; ++k) {
                const int task = half + 2 * k;
                const int mt = cchunks == 2 ? (task >> 1) : task;
                const int cc = cchunks == 2 ? (task & 1) : 0;
                const int row = mt * 128 + m;
                const int hp = row >> 2, pos = row & 3;
                const int hy = hp / kSpHW, hx = hp - hy * kSpHW;
                const uint32_t off = (uint32_t)(pos * p.c0 * 2 + cc * 32);          // byte offset inside the blocked pixel
                const uint32_t unit = (off & 127u) >> 4, sw = (uint32_t)hp & 7u;
                t_off[k] = (off >> 7) * kSpA1Chunk + (uint32_t)hp * 128u + ((unit ^ sw) << 4);
                t_off[k] |= ((((unit + 1u) ^ sw) << 4) ^ ((unit ^ sw) << 4)) << 24;  // xor distance to the second unit
                t_col[k] = acc0_col + (uint32_t)(mt * p.c0 + cc * 16);
                t_flags[k] = (hy == 0 ? 1 : 0) | (hx == 0 ? 2 : 0) | ((task < ntask && row < kSpRows0) ? 4 : 0) | (cc << 3);
            }
            uint32_t tl = 0;
            for (int tile = cta; tile < p.spatial_tiles; tile += ctas, ++tl) {
                uint32_t n, r, th_i, tw_i;
                fdivmod((uint32_t)tile, p.d_img, n, r);
                fdivmod(r, p.d_tw, th_i, tw_i);
                const int edge = (th_i == 0 ? 1 : 0) | (tw_i == 0 ? 2 : 0);     // halo row / column 0 lies outside the image
                const uint32_t b = tl & 1u;
                ptx::mbar_wait(&a1_empty[b], ((tl >> 1) & 1u) ^ 1u);    // MMA 2 of two tiles ago has read this A1 buffer
                ptx::mbar_wait(&acc0_full, tl & 1u);
                ptx::tc_fence_after();
                uint8_t* a1 = a1_s + b * p.a1_buf_bytes;
                auto finish = [&](const uint32_t (&v)[16], int k) {
                    uint32_t o[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 f2 = silu2_half(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), hb[i]);
                        o[i] = pack_bf16x2(f2.x, f2.y);
                    }
                    if (t_flags[k] & 4) {
                        uint8_t* d0 = a1 + (t_off[k] & 0x00ffffffu);
                        uint8_t* d1 = a1 + ((t_off[k] & 0x00ffffffu) ^ (t_off[k] >> 24));
                        if (t_flags[k] & edge) {        // blocked pixel outside the image: layer 1's zero padding
                            *reinterpret_cast<uint4*>(d0) = make_uint4(0, 0, 0, 0);
                            *reinterpret_cast<uint4*>(d1) = make_uint4(0, 0, 0, 0);
                        } else {
                            *reinterpret_cast<uint4*>(d0) = make_uint4(o[0], o[1], o[2], o[3]);
                            *reinterpret_cast<uint4*>(d1) = make_uint4(o[4], o[5], o[6], o[7]);
                        }
                    }
                };
                {
                    // tasks of this warp in pairs: two accumulator loads in flight per wait
                    uint32_t va[16], vb[16];
#pragma unroll
                    for (int k = 0; k < kMaxT; k += 2) {
                        if (half + 2 * k < ntask) {
                            const bool two = k + 1 < kMaxT && half + 2 * (k + 1) < ntask;
                            ptx::tmem_ld16(t_quad + t_col[k], va);
                            if (two) ptx::tmem_ld16(t_quad + t_col[k + 1 < kMaxT ? k + 1 : k], vb);
                            ptx::tmem_ld_wait();
                            finish(va, k);
                            if (two) finish(vb, k + 1 < kMaxT ? k + 1 : k);
                        }
                    }
                }
                ptx::fence_proxy_async();                   // A1 writes -> visible to UMMA
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    ptx::mbar_arrive(&a1_full[b]);
                    ptx::mbar_arrive(&acc0_empty);
                }
            }
        } else {
            // ---- epilogue 2: layer-1 accumulator -> bias + SiLU -> bf16 -> global ----
            uint32_t tl = 0;
            for (int tile = cta; tile < p.spatial_tiles; tile += ctas, ++tl) {
                uint32_t n, r, th_i, tw_i;
                fdivmod((uint32_t)tile, p.d_img, n, r);
                fdivmod(r, p.d_tw, th_i, tw_i);
                const uint32_t b = tl & 1u;
                const int ow = (int)tw_i * kSpTW + tw, oh = (int)th_i * kSpTH + th;
                const bool row_ok = (ow < p.W1) && (oh < p.H1);
                const size_t pix = ((size_t)n * p.H1 + oh) * p.W1 + ow;
                ptx::mbar_wait(&acc1_full[b], (tl >> 1) & 1u);
                ptx::tc_fence_after();
                st.c0 = 0;
                st.c1 = (int)tw_i * kSpTW; st.c2 = (int)th_i * kSpTH; st.c3 = (int)n;
                epi_tile<true, false, false>(t_quad + b * (uint32_t)p.n_pad, ec, bias_s, eo, pix, row_ok, half, lane, st,
                                             EpiResSmem{nullptr, 1, 0, 0, 0});
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&acc1_empty[b]);
            }
            if (st.enabled && st.issuer) ptx::bulk_wait_read0();
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, p.tmem_cols);
}

// Shapes the fused kernel takes (everything else runs layer by layer).
bool stem_pair_ok(int H, int W, int c0, int cout1, int n_pad1) {
    if (env_flag("SPECYOLO_NO_STEM_PAIR")) return false;
    if (H % 4 || W % 16 || H < 4) return false;
    if (c0 != 16 && c0 != 32) return false;
    if (n_pad1 % 16 || n_pad1 < 64 || n_pad1 < cout1 || n_pad1 > 128) return false;
    if ((size_t)n_pad1 * 16 * c0 * 2 > 64 * 1024) return false;        // resident layer-1 weights
    return true;
}

int stem_pair_launch(const specyolo_stem_pair_t* a, cudaStream_t stream) {
    SY_CHECK(stem_pair_ok(a->H, a->W, a->c0, a->Cout, a->n_pad), SPECYOLO_ERR_UNSUPPORTED,
             "stem_pair: unsupported shape (H=%d W=%d c0=%d cout=%d n_pad=%d)", a->H, a->W, a->c0, a->Cout, a->n_pad);
    SY_CHECK(!(reinterpret_cast<uintptr_t>(a->x) & 15) && !(reinterpret_cast<uintptr_t>(a->w0) & 15) &&
                 !(reinterpret_cast<uintptr_t>(a->w1_packed) & 15),
             SPECYOLO_ERR_INVALID, "stem_pair: x / weights must be 16-byte aligned");
    EncodeTiledFn encode = get_encode_fn();
    SY_CHECK(encode != nullptr, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    StemPairParams p{};
    p.B = a->B; p.H1 = a->H / 4; p.W1 = a->W / 4;
    p.tiles_w = ceil_div(p.W1, kSpTW);
    p.tiles_h = ceil_div(p.H1, kSpTH);
    const long spatial = (long)a->B * p.tiles_w * p.tiles_h;
    SY_CHECK(spatial > 0 && spatial < (1L << 30) && fastdiv_ok((uint64_t)spatial, (uint32_t)(p.tiles_w * p.tiles_h)),
             SPECYOLO_ERR_INVALID, "stem_pair: bad tile count");
    p.spatial_tiles = (int)spatial;
    p.d_img = make_fastdiv((uint32_t)(p.tiles_w * p.tiles_h));
    p.d_tw = make_fastdiv((uint32_t)p.tiles_w);
    p.c0 = a->c0;
    p.chunks1 = 4 * a->c0 / 64;
    p.n_pad = a->n_pad; p.cout = a->Cout;
    p.b0 = a->b0; p.b1 = a->b1; p.w0 = a->w0; p.x = a->x;
    p.y = a->y; p.y_pixstride = a->y_pixstride;
    p.w1_bytes = (uint32_t)(4 * p.chunks1) * (uint32_t)a->n_pad * 128u;
    p.a1_buf_bytes = (uint32_t)p.chunks1 * kSpA1Chunk;
    p.off_w0 = (p.w1_bytes + 1023u) & ~1023u;
    p.off_patch = p.off_w0 + 2048u;
    p.off_a0 = p.off_patch + 2u * kSpPatchStride;
    p.off_a1 = p.off_a0 + kSpA0Bytes;
    p.off_st = p.off_a1 + 2u * p.a1_buf_bytes;
    int store_bw = epi_stage_box_cols(a->n_pad, 2);
    if ((reinterpret_cast<uintptr_t>(a->y) & 15) || ((size_t)a->y_pixstride * 2) % 16 || env_flag("SPECYOLO_NO_TMA_STORE"))
        store_bw = 0;
    uint32_t stage_out = epi_stage_bytes(a->n_pad, store_bw, 2);
    if (1024 + p.off_st + stage_out > (uint32_t)kSpMaxDynSmem) { store_bw = 0; stage_out = 0; }
    const size_t smem_bytes = 1024 + (size_t)p.off_st + stage_out;
    SY_CHECK(smem_bytes <= (size_t)kSpMaxDynSmem, SPECYOLO_ERR_UNSUPPORTED, "stem_pair: shared memory budget exceeded");
    p.store_bw = store_bw;
    p.pair_stores = 1;
    p.store_row_bytes = (uint32_t)(store_bw * 2);
    p.store_swz_mask = p.store_row_bytes == 128 ? 7u : (p.store_row_bytes == 64 ? 3u : 1u);
    uint32_t cols = 32;
    while (cols < 2u * (uint32_t)a->n_pad + (uint32_t)(kSpMTiles * a->c0)) cols <<= 1;
    SY_CHECK(cols <= 512, SPECYOLO_ERR_UNSUPPORTED, "stem_pair: TMEM budget exceeded");
    p.tmem_cols = cols;

    CUtensorMap map_w, map_y;
    {
        const int K1 = 16 * a->c0;          // 4 taps x 4*c0 blocked channels
        cuuint64_t dims[2] = {(cuuint64_t)K1, (cuuint64_t)a->n_pad};
        cuuint64_t strides[1] = {(cuuint64_t)K1 * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)a->n_pad};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(a->w1_packed), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SY_CHECK(r == CUDA_SUCCESS, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled(stem_pair W) failed (%d)", (int)r);
    }
    map_y = map_w;
    if (store_bw) {
        const cuuint64_t pix_b = (cuuint64_t)a->y_pixstride * 2;
        cuuint64_t dims[4] = {(cuuint64_t)a->Cout, (cuuint64_t)p.W1, (cuuint64_t)p.H1, (cuuint64_t)a->B};
        cuuint64_t strides[3] = {pix_b, pix_b * p.W1, pix_b * p.W1 * p.H1};
        cuuint32_t box[4] = {(cuuint32_t)store_bw, kSpTW, kSpTH, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&map_y, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, a->y, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for((int)p.store_row_bytes), CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SY_CHECK(r == CUDA_SUCCESS, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled(stem_pair Y) failed (%d)", (int)r);
    }
    static std::once_flag attr_once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(attr_once, [] {
        attr_err = cudaFuncSetAttribute(stem_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSpMaxDynSmem);
    });
    SY_CHECK(attr_err == cudaSuccess, SPECYOLO_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(attr_err));
    const long resident = sm_count();
    const unsigned grid = (unsigned)(spatial < resident ? spatial : resident);
    SY_CUDA(launch_pdl(stem_pair_kernel, dim3(grid), dim3(kSpThreads), smem_bytes, stream, map_w, map_y, p));
    SY_LAUNCH_CHECK();
    count_launch();
    return SPECYOLO_OK;
}

}  // namespace specyolo
