// PSA attention core of C2PSA (ultralytics/nn/modules/block.py:1922-1933) on tcgen05 / TMEM:
//   q,k,v = qkv.view(B, heads, 2*kd+hd, N).split([kd,kd,hd], 2)
//   attn = softmax((q^T k) * scale, -1);  x = v @ attn^T  + pe(v)      (pe = depthwise 3x3 + folded BN)
// The reference materialises the B x heads x N x N matrix through two cuBLAS bmm and a softmax kernel; here the
// scores live in TMEM and the probabilities in shared memory only.
//
// CTA = 128 queries of one (image, head); kd = 32, hd = 64 (every YOLO11 scale).  Keys / values stream through a
// 2-stage TMA ring in blocks of 128 tokens read straight from the NHWC qkv tensor (3-D tensor map {channel, token,
// image}: tokens past N are zero-filled and masked).  Two sweeps over the keys, so any N works (400 tokens at 640^2,
// 1600 at 1280^2) without rescaling an accumulator:
//   sweep 1:  S = Q K_j^T  (UMMA 128x128x32, fp32 in TMEM)  ->  row maximum m
//   sweep 2:  S = Q K_j^T again -> P = exp2((S - m) * scale*log2e) as bf16, written to shared memory in the
//             K-major SWIZZLE_128B operand layout, row sums l accumulated in fp32;
//             O += P V_j  (UMMA 128x64x128; V_j is the [token][channel] tile as TMA wrote it = an MN-major B operand)
//   end:      x = O / l + pe(v) -> bf16 NHWC store.  pe(v) (9 taps x 64 channels per token) is read from L2.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-5 softmax / epilogue (one query row per
// thread: TMEM lane = row, so the row reductions need no shuffles).  TMEM: 128 (S) + 64 (O) columns -> two CTAs
// per SM overlap one tile's softmax with the other's MMAs.
#include "common.h"
#include "ptx.cuh"
#include "tma_host.h"

namespace specyolo {

static constexpr int KD = 32, HD = 64;
static constexpr int kAttThreads = 192;
static constexpr int kAttQ = 128;            // queries per CTA (UMMA M)
static constexpr int kAttKB = 128;           // keys per block
static constexpr uint32_t kQBytes = kAttQ * KD * 2;        // 8 KB, 64-byte rows, SWIZZLE_64B
static constexpr uint32_t kKBytes = kAttKB * KD * 2;       // 8 KB
static constexpr uint32_t kVBytes = kAttKB * HD * 2;       // 16 KB, 128-byte rows, SWIZZLE_128B
static constexpr uint32_t kPBytes = kAttQ * kAttKB * 2;    // 32 KB: two [128 rows x 64 keys] K-major SW128 chunks
static constexpr uint32_t kAttSmem = 1024 + kQBytes + 2 * (kKBytes + kVBytes) + kPBytes;   // 89 KB
static constexpr uint32_t kAttTmemCols = 256;              // S at column 0, O at column 128

struct AttParams {
    int N, H, W, heads, nblk;
    float scale_log2e;
    const __nv_bfloat16* qkv;
    int qkv_pixstride;
    const float* pe_w;
    const float* pe_b;
    __nv_bfloat16* out;
    int out_pixstride;
};

__global__ void __launch_bounds__(kAttThreads)
psa_attention_kernel(const __grid_constant__ CUtensorMap map_qk, const __grid_constant__ CUtensorMap map_v,
                     const __grid_constant__ AttParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t q_full, kv_full[2], kv_empty[2], s_full, s_empty, p_full, p_empty, o_full;
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float pe_ws[9 * HD];    // [tap][d]
    __shared__ __align__(16) float pe_bs[HD];

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    ptx::grid_dep_launch();
    const int q0 = blockIdx.x * kAttQ, head = blockIdx.y, b = blockIdx.z;
    const int ch0 = head * (2 * KD + HD);

    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* base = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* q_s = base;
    uint8_t* k_s = q_s + kQBytes;                 // 2 stages
    uint8_t* v_s = k_s + 2 * kKBytes;             // 2 stages
    uint8_t* p_s = v_s + 2 * kVBytes;

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&map_qk);
        ptx::prefetch_tmap(&map_v);
        ptx::mbar_init(&q_full, 1);
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(&kv_full[s], 1);
            ptx::mbar_init(&kv_empty[s], 1);
        }
        ptx::mbar_init(&s_full, 1);
        ptx::mbar_init(&s_empty, 4);
        ptx::mbar_init(&p_full, 4);
        ptx::mbar_init(&p_empty, 1);
        ptx::mbar_init(&o_full, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc(&tmem_base_smem, kAttTmemCols);
    for (int i = threadIdx.x; i < 9 * HD; i += kAttThreads) {
        const int tap = i / HD, d = i - tap * HD;
        pe_ws[i] = p.pe_w[(head * HD + d) * 9 + tap];
    }
    for (int i = threadIdx.x; i < HD; i += kAttThreads) pe_bs[i] = p.pe_b[head * HD + i];
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    ptx::grid_dep_wait();       // PDL: qkv is read (TMA, pe taps) and `out` written only below
    const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;
    const int nblk = p.nblk;

    if (warp == 0) {
        // ===================== TMA producer =====================
        const bool leader = ptx::elect_one();
        if (leader) {
            ptx::mbar_expect_tx(&q_full, kQBytes);
            ptx::tma_load_3d(q_s, &map_qk, &q_full, ch0, q0, b);
        }
        int stage = 0;
        uint32_t ph = 0;
        for (int sweep = 0; sweep < 2; ++sweep) {
            for (int j = 0; j < nblk; ++j) {
                ptx::mbar_wait(&kv_empty[stage], ph ^ 1u);
                if (leader) {
                    ptx::mbar_expect_tx(&kv_full[stage], sweep ? kKBytes + kVBytes : kKBytes);
                    ptx::tma_load_3d(k_s + stage * kKBytes, &map_qk, &kv_full[stage], ch0 + KD, j * kAttKB, b);
                    if (sweep) ptx::tma_load_3d(v_s + stage * kVBytes, &map_v, &kv_full[stage], ch0 + 2 * KD, j * kAttKB, b);
                }
                if (++stage == 2) { stage = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const bool leader = ptx::elect_one();
        const uint32_t idesc_s = ptx::umma_idesc_bf16(128, kAttKB);
        const uint32_t idesc_o = ptx::umma_idesc_bf16_bmn(128, HD);
        const uint64_t dq = ptx::umma_desc(ptx::smem_u32(q_s), 0, 8 * 64, 64);
        const uint64_t dp = ptx::umma_desc(ptx::smem_u32(p_s), 0, 8 * 128, 128);
        ptx::mbar_wait(&q_full, 0);
        int stage = 0;
        uint32_t ph = 0, s_ph = 0, p_ph = 0;
        for (int sweep = 0; sweep < 2; ++sweep) {
            for (int j = 0; j < nblk; ++j) {
                ptx::mbar_wait(&kv_full[stage], ph);
                ptx::mbar_wait(&s_empty, s_ph ^ 1u);          // softmax warps have read the previous S
                s_ph ^= 1u;
                ptx::tc_fence_after();
                const uint64_t dk = ptx::umma_desc(ptx::smem_u32(k_s + stage * kKBytes), 0, 8 * 64, 64);
                if (leader) {
                    ptx::umma_bf16(tmem_s, dq, dk, idesc_s, 0u);
                    ptx::umma_bf16(tmem_s, dq + 2, dk + 2, idesc_s, 1u);        // K step 2: +32 bytes
                    ptx::umma_commit(&s_full);
                }
                if (sweep) {
                    ptx::mbar_wait(&p_full, p_ph);             // P_j is in shared memory
                    p_ph ^= 1u;
                    ptx::tc_fence_after();
                    const uint64_t dv = ptx::umma_desc(ptx::smem_u32(v_s + stage * kVBytes), 0, 8 * 128, 128);
                    if (leader) {
#pragma unroll
                        for (int kk = 0; kk < kAttKB / 16; ++kk) {
                            // A: 16 keys = 32 bytes inside the 64-key chunk kk/4; B: 16 token rows = 2048 bytes
                            const uint64_t da = dp + (uint64_t)(((kk >> 2) * (kAttQ * 128) + (kk & 3) * 32) >> 4);
                            ptx::umma_bf16(tmem_o, da, dv + (uint64_t)((kk * 2048) >> 4), idesc_o, (j | kk) ? 1u : 0u);
                        }
                        ptx::umma_commit(&p_empty);
                    }
                }
                if (leader) ptx::umma_commit(&kv_empty[stage]);
                if (++stage == 2) { stage = 0; ph ^= 1u; }
            }
        }
        if (leader) ptx::umma_commit(&o_full);
    } else {
        // ===================== softmax / epilogue (warps 2..5) =====================
        const int quad = warp & 3;
        const int row = quad * 32 + lane;                       // query inside the tile = TMEM lane
        const uint32_t t_lane = (uint32_t)(quad * 32) << 16;
        const float c = p.scale_log2e;
        uint32_t s_ph = 0, pe_ph = 0;
        float m = -INFINITY, l = 0.f;
        // ---- sweep 1: row maximum ----
        for (int j = 0; j < nblk; ++j) {
            ptx::mbar_wait(&s_full, s_ph);
            s_ph ^= 1u;
            ptx::tc_fence_after();
            const int kvalid = min(kAttKB, p.N - j * kAttKB);
#pragma unroll 1
            for (int c0 = 0; c0 < kAttKB; c0 += 32) {
                uint32_t v[32];
                ptx::tmem_ld32(tmem_s + t_lane + c0, v);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (c0 + i < kvalid) m = fmaxf(m, __uint_as_float(v[i]));
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&s_empty);
        }
        const float mc = m * c;
        // ---- sweep 2: P = exp2(S*c - m*c) -> shared memory (bf16, UMMA A layout), l = row sum ----
        uint8_t* p_row = p_s + row * 128;
        const uint32_t sw = (uint32_t)(row & 7);
        for (int j = 0; j < nblk; ++j) {
            ptx::mbar_wait(&s_full, s_ph);
            s_ph ^= 1u;
            ptx::mbar_wait(&p_empty, pe_ph ^ 1u);               // previous P·V has finished reading P
            pe_ph ^= 1u;
            ptx::tc_fence_after();
            const int kvalid = min(kAttKB, p.N - j * kAttKB);
#pragma unroll 1
            for (int c0 = 0; c0 < kAttKB; c0 += 32) {
                uint32_t v[32];
                ptx::tmem_ld32(tmem_s + t_lane + c0, v);
                ptx::tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float e0 = exp2f(fmaf(__uint_as_float(v[2 * i]), c, -mc));
                    float e1 = exp2f(fmaf(__uint_as_float(v[2 * i + 1]), c, -mc));
                    if (c0 + 2 * i >= kvalid) e0 = 0.f;
                    if (c0 + 2 * i + 1 >= kvalid) e1 = 0.f;
                    // the sum uses the bf16-rounded probabilities the tensor core will multiply
                    const __nv_bfloat162 pb = __floats2bfloat162_rn(e0, e1);
                    const float2 pf = __bfloat1622float2(pb);
                    l += pf.x + pf.y;
                    pk[i] = *reinterpret_cast<const uint32_t*>(&pb);
                }
                // 32 keys = 4 x 16-byte units of this row; unit index inside the 64-key chunk is swizzled by row % 8
                uint8_t* chunk = p_row + (c0 >> 6) * (kAttQ * 128);
                const uint32_t u0 = (uint32_t)(c0 & 63) >> 3;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    *reinterpret_cast<uint4*>(chunk + (((u0 + u) ^ sw) << 4)) =
                        make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
            }
            ptx::tc_fence_before();
            ptx::fence_proxy_async();                           // generic-proxy writes of P -> visible to UMMA
            __syncwarp();
            if (lane == 0) {
                ptx::mbar_arrive(&s_empty);
                ptx::mbar_arrive(&p_full);
            }
        }
        // ---- epilogue: O / l + pe(v) ----
        ptx::mbar_wait(&o_full, 0);
        ptx::tc_fence_after();
        const int qi = q0 + row;
        const float inv_l = 1.0f / l;
        const int h = qi / p.W, w = qi - h * p.W;
        const __nv_bfloat16* vbase = p.qkv + (size_t)b * p.N * p.qkv_pixstride + ch0 + 2 * KD;
        __nv_bfloat16* op = p.out + ((size_t)b * p.N + qi) * p.out_pixstride + head * HD;
#pragma unroll 1
        for (int d0 = 0; d0 < HD; d0 += 16) {
            uint32_t v[16];
            ptx::tmem_ld16(tmem_o + t_lane + d0, v);
            ptx::tmem_ld_wait();
            if (qi >= p.N) continue;
            float res[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) res[i] = fmaf(__uint_as_float(v[i]), inv_l, pe_bs[d0 + i]);
            for (int ky = 0; ky < 3; ++ky) {
                const int hh = h + ky - 1;
                if (hh < 0 || hh >= p.H) continue;
                for (int kx = 0; kx < 3; ++kx) {
                    const int ww = w + kx - 1;
                    if (ww < 0 || ww >= p.W) continue;
                    const uint4* vp = reinterpret_cast<const uint4*>(vbase + (size_t)(hh * p.W + ww) * p.qkv_pixstride + d0);
                    const float* wt = pe_ws + (ky * 3 + kx) * HD + d0;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const uint4 u = __ldg(vp + half);
                        const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const float2 f = unpack_bf16x2(uu[jj]);
                            const int d = half * 8 + 2 * jj;
                            res[d] = fmaf(f.x, wt[d], res[d]);
                            res[d + 1] = fmaf(f.y, wt[d + 1], res[d + 1]);
                        }
                    }
                }
            }
            uint4 o0, o1;
            o0.x = pack_bf16x2(res[0], res[1]);   o0.y = pack_bf16x2(res[2], res[3]);
            o0.z = pack_bf16x2(res[4], res[5]);   o0.w = pack_bf16x2(res[6], res[7]);
            o1.x = pack_bf16x2(res[8], res[9]);   o1.y = pack_bf16x2(res[10], res[11]);
            o1.z = pack_bf16x2(res[12], res[13]); o1.w = pack_bf16x2(res[14], res[15]);
            reinterpret_cast<uint4*>(op + d0)[0] = o0;
            reinterpret_cast<uint4*>(op + d0)[1] = o1;
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, kAttTmemCols);
}

// ------------------------------------------------------------------------------------------------
// Small-N variant (N <= 512 tokens: the 20 x 20 map of a 640^2 input): everything of one (image, head) is resident.
// The two-sweep kernel above is a chain of ~16 barrier hand-offs per CTA (TMA -> S MMA -> softmax -> P -> PV MMA, twice
// per key block) and recomputes S; at N = 400 it took 108 us for ~0.04 us of tensor work.  Here
//   * all K / V blocks are loaded at once (<= 96 KB) and S = Q K^T for ALL keys sits in TMEM (<= 512 columns),
//   * EIGHT softmax warps (two per TMEM lane quadrant, 64 of the 128 columns of a block each) take the row maximum
//     straight from TMEM, meet once in shared memory, then re-read S (no second MMA sweep), write P = exp2(..) block by
//     block into a double-buffered operand tile while the PV MMAs of the previous block run,
//   * O accumulates in the columns of S block 0 (free once P_0 is written),
//   * the positional depthwise conv reads its nine V neighbours from the resident V tile in shared memory.
// ------------------------------------------------------------------------------------------------
// 2^x on the MUFU unit (ex2.approx: relative error ~2^-22, far below the bf16 rounding of P); exp2f() expands to ~20
// instructions of range handling, and at 256 exponentials per thread that made the softmax warps issue-bound
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

static constexpr int kAttSmallThreads = 64 + 256;
static constexpr int kAttMaxBlk = 4;

__global__ void __launch_bounds__(kAttSmallThreads, 1)
psa_attention_small_kernel(const __grid_constant__ CUtensorMap map_qk, const __grid_constant__ CUtensorMap map_v,
                           const __grid_constant__ AttParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t q_full, k_full[kAttMaxBlk], v_full[kAttMaxBlk], s_full[kAttMaxBlk], p_full[2], p_empty[2], o_full;
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float pe_ws[9 * HD];    // [tap][d]
    __shared__ __align__(16) float pe_bs[HD];
    __shared__ float xch[2][kAttQ];                  // row maximum / row sum of the two column halves

    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    ptx::grid_dep_launch();
    const int q0 = blockIdx.x * kAttQ, head = blockIdx.y, b = blockIdx.z;
    const int ch0 = head * (2 * KD + HD);
    const int nblk = p.nblk;

    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* base = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
    uint8_t* q_s = base;
    uint8_t* k_s = q_s + kQBytes;                          // nblk blocks
    uint8_t* v_s = k_s + (uint32_t)nblk * kKBytes;         // nblk blocks
    uint8_t* p_s = v_s + (uint32_t)nblk * kVBytes;         // 2 buffers

    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&map_qk);
        ptx::prefetch_tmap(&map_v);
        ptx::mbar_init(&q_full, 1);
        for (int j = 0; j < kAttMaxBlk; ++j) {
            ptx::mbar_init(&k_full[j], 1);
            ptx::mbar_init(&v_full[j], 1);
            ptx::mbar_init(&s_full[j], 1);
        }
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&p_full[i], 8);
            ptx::mbar_init(&p_empty[i], 1);
        }
        ptx::mbar_init(&o_full, 1);
        ptx::fence_mbar_init();
    }
    uint32_t tcols = 128;
    while (tcols < (uint32_t)nblk * 128u) tcols <<= 1;
    if (warp == 1) ptx::tmem_alloc(&tmem_base_smem, tcols);
    for (int i = threadIdx.x; i < 9 * HD; i += kAttSmallThreads) {
        const int tap = i / HD, d = i - tap * HD;
        pe_ws[i] = p.pe_w[(head * HD + d) * 9 + tap];
    }
    for (int i = threadIdx.x; i < HD; i += kAttSmallThreads) pe_bs[i] = p.pe_b[head * HD + i];
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    ptx::grid_dep_wait();

    if (warp == 0) {
        // ===================== TMA producer: Q, every K block, every V block =====================
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(&q_full, kQBytes);
            ptx::tma_load_3d(q_s, &map_qk, &q_full, ch0, q0, b);
            for (int j = 0; j < nblk; ++j) {
                ptx::mbar_expect_tx(&k_full[j], kKBytes);
                ptx::tma_load_3d(k_s + j * kKBytes, &map_qk, &k_full[j], ch0 + KD, j * kAttKB, b);
            }
            for (int j = 0; j < nblk; ++j) {
                ptx::mbar_expect_tx(&v_full[j], kVBytes);
                ptx::tma_load_3d(v_s + j * kVBytes, &map_v, &v_full[j], ch0 + 2 * KD, j * kAttKB, b);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const bool leader = ptx::elect_one();
        const uint32_t idesc_s = ptx::umma_idesc_bf16(128, kAttKB);
        const uint32_t idesc_o = ptx::umma_idesc_bf16_bmn(128, HD);
        const uint64_t dq = ptx::umma_desc(ptx::smem_u32(q_s), 0, 8 * 64, 64);
        ptx::mbar_wait(&q_full, 0);
        for (int j = 0; j < nblk; ++j) {
            ptx::mbar_wait(&k_full[j], 0);
            ptx::tc_fence_after();
            const uint64_t dk = ptx::umma_desc(ptx::smem_u32(k_s + j * kKBytes), 0, 8 * 64, 64);
            if (leader) {
                ptx::umma_bf16(tmem_base + (uint32_t)j * 128u, dq, dk, idesc_s, 0u);
                ptx::umma_bf16(tmem_base + (uint32_t)j * 128u, dq + 2, dk + 2, idesc_s, 1u);        // K step 2: +32 bytes
                ptx::umma_commit(&s_full[j]);
            }
        }
        for (int j = 0; j < nblk; ++j) {
            const uint32_t pb = (uint32_t)j & 1u;
            ptx::mbar_wait(&p_full[pb], (uint32_t)(j >> 1) & 1u);      // P_j is in shared memory (and S_0 fully consumed)
            ptx::mbar_wait(&v_full[j], 0);
            ptx::tc_fence_after();
            const uint64_t dp = ptx::umma_desc(ptx::smem_u32(p_s + pb * kPBytes), 0, 8 * 128, 128);
            const uint64_t dv = ptx::umma_desc(ptx::smem_u32(v_s + j * kVBytes), 0, 8 * 128, 128);
            if (leader) {
#pragma unroll
                for (int kk = 0; kk < kAttKB / 16; ++kk) {
                    const uint64_t da = dp + (uint64_t)(((kk >> 2) * (kAttQ * 128) + (kk & 3) * 32) >> 4);
                    ptx::umma_bf16(tmem_base, da, dv + (uint64_t)((kk * 2048) >> 4), idesc_o, (j | kk) ? 1u : 0u);
                }
                ptx::umma_commit(&p_empty[pb]);
            }
        }
        if (leader) ptx::umma_commit(&o_full);
    } else {
        // ===================== softmax / epilogue (warps 2..9) =====================
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;                       // which 64 of a block's 128 key columns / which 32 output channels
        const int row = quad * 32 + lane;                       // query inside the tile = TMEM lane
        const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16);
        const float c = p.scale_log2e;
        // ---- pass 1: row maximum over this warp's columns of every block ----
        float m = -INFINITY;
        for (int j = 0; j < nblk; ++j) {
            ptx::mbar_wait(&s_full[j], 0);
            ptx::tc_fence_after();
            const int kvalid = p.N - j * kAttKB - half * 64;    // valid columns of this half
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 32) {
                uint32_t v[32];
                ptx::tmem_ld32(t_row + (uint32_t)(j * 128 + half * 64 + c0), v);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (c0 + i < kvalid) m = fmaxf(m, __uint_as_float(v[i]));
            }
        }
        xch[half][row] = m;
        ptx::named_bar_sync(1, 256);
        m = fmaxf(xch[0][row], xch[1][row]);
        ptx::named_bar_sync(1, 256);                            // everyone has read the maxima: xch is reused for the sums
        const float mc = m * c;
        // ---- pass 2: P = exp2(S*c - m*c) (bf16, UMMA A layout), l = row sum of the rounded probabilities ----
        float l = 0.f;
        const uint32_t sw = (uint32_t)(row & 7);
        for (int j = 0; j < nblk; ++j) {
            const uint32_t pb = (uint32_t)j & 1u;
            if (j >= 2) ptx::mbar_wait(&p_empty[pb], (uint32_t)((j >> 1) + 1) & 1u);      // PV of block j-2 has read this buffer
            uint8_t* p_row = p_s + pb * kPBytes + half * (kAttQ * 128) + row * 128;
            const int kvalid = p.N - j * kAttKB - half * 64;
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 32) {
                uint32_t v[32];
                ptx::tmem_ld32(t_row + (uint32_t)(j * 128 + half * 64 + c0), v);
                ptx::tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float e0 = ex2_approx(fmaf(__uint_as_float(v[2 * i]), c, -mc));
                    float e1 = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), c, -mc));
                    if (c0 + 2 * i >= kvalid) e0 = 0.f;
                    if (c0 + 2 * i + 1 >= kvalid) e1 = 0.f;
                    const __nv_bfloat162 pbf = __floats2bfloat162_rn(e0, e1);
                    const float2 pf = __bfloat1622float2(pbf);
                    l += pf.x + pf.y;
                    pk[i] = *reinterpret_cast<const uint32_t*>(&pbf);
                }
                const uint32_t u0 = (uint32_t)c0 >> 3;
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    *reinterpret_cast<uint4*>(p_row + (((u0 + u) ^ sw) << 4)) =
                        make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
            }
            ptx::tc_fence_before();
            ptx::fence_proxy_async();                           // generic-proxy writes of P -> visible to UMMA
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&p_full[pb]);
        }
        xch[half][row] = l;
        ptx::named_bar_sync(1, 256);
        l = xch[0][row] + xch[1][row];
        // ---- epilogue: O / l + pe(v), this warp's 32 of the 64 channels ----
        for (int j = 0; j < nblk; ++j) ptx::mbar_wait(&v_full[j], 0);       // the positional conv reads V from shared memory
        ptx::mbar_wait(&o_full, 0);
        ptx::tc_fence_after();
        const int qi = q0 + row;
        const float inv_l = 1.0f / l;
        const int h = qi / p.W, w = qi - h * p.W;
        __nv_bfloat16* op = p.out + ((size_t)b * p.N + qi) * p.out_pixstride + head * HD;
#pragma unroll 1
        for (int d0 = half * 32; d0 < half * 32 + 32; d0 += 16) {
            uint32_t v[16];
            ptx::tmem_ld16(t_row + (uint32_t)d0, v);
            ptx::tmem_ld_wait();
            if (qi >= p.N) continue;
            float res[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) res[i] = fmaf(__uint_as_float(v[i]), inv_l, pe_bs[d0 + i]);
            for (int ky = 0; ky < 3; ++ky) {
                const int hh = h + ky - 1;
                if (hh < 0 || hh >= p.H) continue;
                for (int kx = 0; kx < 3; ++kx) {
                    const int ww = w + kx - 1;
                    if (ww < 0 || ww >= p.W) continue;
                    const int tok = hh * p.W + ww;
                    const uint8_t* vrow = v_s + (uint32_t)(tok >> 7) * kVBytes + (uint32_t)(tok & 127) * 128u;
                    const uint32_t vsw = (uint32_t)(tok & 7);
                    const float* wt = pe_ws + (ky * 3 + kx) * HD + d0;
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        const uint4 u = *reinterpret_cast<const uint4*>(vrow + ((((uint32_t)d0 >> 3) + (uint32_t)hf) ^ vsw) * 16u);
                        const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const float2 f = unpack_bf16x2(uu[jj]);
                            const int d = hf * 8 + 2 * jj;
                            res[d] = fmaf(f.x, wt[d], res[d]);
                            res[d + 1] = fmaf(f.y, wt[d + 1], res[d + 1]);
                        }
                    }
                }
            }
            uint4 o0, o1;
            o0.x = pack_bf16x2(res[0], res[1]);   o0.y = pack_bf16x2(res[2], res[3]);
            o0.z = pack_bf16x2(res[4], res[5]);   o0.w = pack_bf16x2(res[6], res[7]);
            o1.x = pack_bf16x2(res[8], res[9]);   o1.y = pack_bf16x2(res[10], res[11]);
            o1.z = pack_bf16x2(res[12], res[13]); o1.w = pack_bf16x2(res[14], res[15]);
            reinterpret_cast<uint4*>(op + d0)[0] = o0;
            reinterpret_cast<uint4*>(op + d0)[1] = o1;
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, tcols);
}

int psa_attention_launch(const void* qkv, int qkv_pixstride, int B, int H, int W, int heads, int key_dim,
                         int head_dim, float scale, const float* pe_w, const float* pe_b, void* out,
                         int out_pixstride, cudaStream_t stream) {
    SY_CHECK(key_dim == KD && head_dim == HD, SPECYOLO_ERR_UNSUPPORTED,
             "psa attention supports key_dim=32, head_dim=64 (got %d, %d)", key_dim, head_dim);
    SY_CHECK(qkv_pixstride % 8 == 0 && out_pixstride % 8 == 0 &&
                 ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
             SPECYOLO_ERR_INVALID, "psa attention: tensors must be 16-byte aligned");
    SY_CHECK(qkv_pixstride >= heads * (2 * KD + HD), SPECYOLO_ERR_INVALID, "psa attention: qkv pixel stride too small");
    EncodeTiledFn encode = get_encode_fn();
    SY_CHECK(encode != nullptr, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const int N = H * W;
    AttParams p{};
    p.N = N; p.H = H; p.W = W; p.heads = heads;
    p.nblk = ceil_div(N, kAttKB);
    p.scale_log2e = scale * 1.4426950408889634f;
    p.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv);
    p.qkv_pixstride = qkv_pixstride;
    p.pe_w = pe_w; p.pe_b = pe_b;
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.out_pixstride = out_pixstride;

    // 3-D views {channel, token, image} of the NHWC qkv tensor: 32-channel boxes for Q / K, 64-channel boxes for V
    CUtensorMap map_qk, map_v;
    const cuuint64_t pix_b = (cuuint64_t)qkv_pixstride * 2;
    cuuint64_t dims[3] = {(cuuint64_t)heads * (2 * KD + HD), (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t strides[2] = {pix_b, pix_b * N};
    cuuint32_t estr[3] = {1, 1, 1};
    {
        cuuint32_t box[3] = {KD, kAttKB, 1};
        CUresult r = encode(&map_qk, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SY_CHECK(r == CUDA_SUCCESS, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled(attention Q/K) failed (%d)", (int)r);
    }
    {
        cuuint32_t box[3] = {HD, kAttKB, 1};
        CUresult r = encode(&map_v, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SY_CHECK(r == CUDA_SUCCESS, SPECYOLO_ERR_CUDA, "cuTensorMapEncodeTiled(attention V) failed (%d)", (int)r);
    }
    dim3 grid((unsigned)ceil_div(N, kAttQ), (unsigned)heads, (unsigned)B);
    if (p.nblk <= kAttMaxBlk && !env_flag("SPECYOLO_ATT_TWO_SWEEP")) {
        // N <= 512: everything of one (image, head) resident (Q + all K / V blocks + two P buffers)
        const size_t smem = 1024 + (size_t)kQBytes + (size_t)p.nblk * (kKBytes + kVBytes) + 2 * (size_t)kPBytes;
        SY_CUDA(cudaFuncSetAttribute(psa_attention_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SY_CUDA(launch_pdl(psa_attention_small_kernel, grid, dim3(kAttSmallThreads), smem, stream, map_qk, map_v, p));
    } else {
        SY_CUDA(cudaFuncSetAttribute(psa_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAttSmem));
        SY_CUDA(launch_pdl(psa_attention_kernel, grid, dim3(kAttThreads), (size_t)kAttSmem, stream, map_qk, map_v, p));
    }
    SY_LAUNCH_CHECK();
    count_launch();
    return SPECYOLO_OK;
}

}  // namespace specyolo
