// Probe: can a UMMA K-major SWIZZLE_128B operand start at a row that is NOT 1024-byte aligned (a shifted window
// into a TMA-written tile), and which value must the descriptor's base_offset field carry?
// D[128 x 64] = A[128 x 64] * B^T with B = identity  ->  D row m must equal tile row (m + shift).
// Also probes an 8-row-group stride (SBO) that is not a multiple of 1024 B.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../spectrogram-yolov11_b200/csrc/ptx.cuh"
using namespace specyolo;

struct Cfg { int shift; int base_off; int sbo_rows; int rowbytes; };

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap map_a,
                                             const __grid_constant__ CUtensorMap map_b, Cfg c, float* out) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t bar, mbar;
    __shared__ uint32_t tmem_s;
    const uint32_t ra = ptx::smem_u32(raw);
    uint8_t* base = raw + (((ra + 1023u) & ~1023u) - ra);
    uint8_t* a_s = base;                 // 256 rows x rowbytes
    uint8_t* b_s = base + 256 * 128;     // 64 rows x rowbytes
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::mbar_init(&mbar, 1); ptx::fence_mbar_init(); }
    if (warp == 0) ptx::tmem_alloc(&tmem_s, 64);
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tm = tmem_s;
    const int kelems = c.rowbytes / 2;
    if (threadIdx.x == 0) {
        ptx::mbar_expect_tx(&bar, 256 * c.rowbytes + 64 * c.rowbytes);
        ptx::tma_load_2d(a_s, &map_a, &bar, 0, 0);
        ptx::tma_load_2d(b_s, &map_b, &bar, 0, 0);
        ptx::mbar_wait(&bar, 0);
        ptx::tc_fence_after();
        const uint32_t idesc = ptx::umma_idesc_bf16(128, 64);
        const uint64_t layout = c.rowbytes == 128 ? 2ull : (c.rowbytes == 64 ? 4ull : 6ull);
        for (int k = 0; k < kelems / 16; ++k) {
            const uint32_t a_addr = ptx::smem_u32(a_s) + c.shift * c.rowbytes + k * 32;
            uint64_t da = 0;
            da |= (uint64_t)((a_addr & 0x3FFFF) >> 4);
            da |= (uint64_t)((c.sbo_rows * c.rowbytes) >> 4) << 32;
            da |= 1ull << 46;
            da |= (uint64_t)(c.base_off & 7) << 49;
            da |= layout << 61;
            const uint64_t db = ptx::umma_smem_desc(ptx::smem_u32(b_s) + k * 32, c.rowbytes);
            ptx::umma_bf16(tm, da, db, idesc, k > 0);
        }
        ptx::umma_commit(&mbar);
    }
    ptx::mbar_wait(&mbar, 0);
    ptx::tc_fence_after();
    for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        ptx::tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
        ptx::tmem_ld_wait();
        for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 64 + c0 + i] = __uint_as_float(v[i]);
    }
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tm, 64);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* sym = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)sym;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int rowbytes : {128, 64, 32}) {
        const int K = rowbytes / 2;
        // A: 256 rows x K, value = row + col/1000 (bf16-exact enough: compare row ids via col 0)
        std::vector<__nv_bfloat16> ha(256 * K), hb(64 * K);
        for (int r = 0; r < 256; ++r) for (int cidx = 0; cidx < K; ++cidx) ha[r * K + cidx] = __float2bfloat16((float)(r % 251) + (cidx == 0 ? 0.f : 0.f) + (float)cidx * 0.0f + (cidx == 1 ? 1000.f : 0.f) * 0);
        // make columns distinguishable: A[r][c] = r if c==0 else (c==1 ? 256+r : 0)
        for (int r = 0; r < 256; ++r) for (int cidx = 0; cidx < K; ++cidx) ha[r * K + cidx] = __float2bfloat16(cidx == 0 ? (float)r : (cidx == K - 1 ? (float)(r % 64) : 0.f));
        for (int n = 0; n < 64; ++n) for (int cidx = 0; cidx < K; ++cidx) hb[n * K + cidx] = __float2bfloat16(n == cidx ? 1.f : 0.f);
        __nv_bfloat16 *da, *db; float* dout;
        cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
        cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
        cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
        CUtensorMapSwizzle sw = rowbytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : rowbytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
        CUtensorMap ma, mb;
        cuuint64_t dims[2] = {(cuuint64_t)K, 256}, st[1] = {(cuuint64_t)K * 2};
        cuuint32_t box[2] = {(cuuint32_t)K, 256}, es[2] = {1, 1};
        enc(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint64_t dimb[2] = {(cuuint64_t)K, 64}; cuuint32_t boxb[2] = {(cuuint32_t)K, 64};
        enc(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, dimb, st, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        // (shift, base_off, sbo_rows): sbo_rows = rows between consecutive 8-row groups
        const int cases[][3] = {{0, 0, 8}, {1, 0, 8}, {1, 1, 8}, {2, 0, 8}, {2, 2, 8}, {3, 3, 8}, {5, 5, 8}, {5, 0, 8},
                                {0, 0, 10}, {1, 1, 10}, {1, 0, 10}, {0, 0, 16}, {1, 1, 16}, {2, 2, 16}, {2, 0, 16}, {0, 0, 18}, {1,0,18}};
        for (auto& cs : cases) {
            Cfg c{cs[0], cs[1], cs[2], rowbytes};
            cudaMemset(dout, 0, 128 * 64 * 4);
            probe<<<1, 128, 64 * 1024>>>(ma, mb, c, dout);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("rowbytes %d case %d,%d,%d: CUDA error %s\n", rowbytes, cs[0], cs[1], cs[2], cudaGetErrorString(e)); return 1; }
            std::vector<float> ho(128 * 64);
            cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
            int bad = 0, firstbad = -1;
            for (int m = 0; m < 128; ++m) {
                const int src = cs[0] + (m / 8) * cs[2] + (m % 8);      // expected tile row feeding D row m
                const float e0 = (float)src, e1 = (float)(src % 64);
                if (ho[m * 64 + 0] != e0 || ho[m * 64 + K - 1] != e1) { if (!bad) firstbad = m; ++bad; }
            }
            printf("rowbytes %3d shift %d base_off %d sbo_rows %2d : %s (bad rows %d, first %d, D[first][0]=%g)\n", rowbytes, cs[0], cs[1],
                   cs[2], bad ? "MISMATCH" : "ok", bad, firstbad, firstbad >= 0 ? ho[firstbad * 64] : 0.f);
        }
        cudaFree(da); cudaFree(db); cudaFree(dout);
    }
    return 0;
}
