// Probe: tensor-pipe throughput when two issuer warps feed it DIFFERENT MMA shapes (N = n1 and n2, separate accumulators)
// versus each shape alone — does switching shape between consecutive MMAs cost anything?
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include "../../spectrogram-yolov11_b200/csrc/ptx.cuh"
using namespace specyolo;

__global__ void __launch_bounds__(128) probe(int n1, int n2, int count1, int count2, int rowbytes, long long* out) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_s;
    const uint32_t ra = ptx::smem_u32(raw);
    uint8_t* base = raw + (((ra + 1023u) & ~1023u) - ra);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;
    const int issuers = (count1 > 0) + (count2 > 0);
    if (threadIdx.x == 0) { ptx::mbar_init(&mbar, issuers); ptx::fence_mbar_init(); }
    if (warp == 0) ptx::tmem_alloc(&tmem_s, 512);
    ptx::fence_proxy_async();
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tm = tmem_s;
    const int n = warp == 1 ? n1 : n2, count = warp == 1 ? count1 : count2;
    if ((warp == 1 || warp == 2) && count > 0) {
        const bool leader = ptx::elect_one();
        const uint32_t idesc = ptx::umma_idesc_bf16(128, n);
        const uint64_t layout = rowbytes == 128 ? 2ull : (rowbytes == 64 ? 4ull : 6ull);
        const uint64_t hi_a = ((uint64_t)((18 * rowbytes) >> 4) << 32) | (1ull << 46) | (layout << 61);
        const uint64_t hi_b = ((uint64_t)((8 * rowbytes) >> 4) << 32) | (1ull << 46) | (layout << 61);
        const uint32_t a16 = ptx::smem_u32(base) >> 4;
        const uint32_t b16 = (ptx::smem_u32(base) + 48 * 1024) >> 4;
        const uint32_t d = tm + (uint32_t)(warp - 1) * 256u;
        const uint32_t bstep = (uint32_t)(n * rowbytes) >> 4;
        const long long t0 = clock64();
        for (int i = 0; i < count; i += 9) {
#pragma unroll
            for (int j = 0; j < 9; ++j)           // nine taps: A shifted by (j / 3) * 18 + j % 3 rows, B = box j
                if (leader) ptx::umma_bf16(d, hi_a | (uint64_t)(a16 + (((j / 3) * 18 + j % 3) * rowbytes >> 4)),
                                           hi_b | (uint64_t)(b16 + j * bstep), idesc, 1u);
        }
        if (leader) ptx::umma_commit(&mbar);
        ptx::mbar_wait(&mbar, 0);
        if (leader && ((warp == 1) || count1 == 0)) out[0] = clock64() - t0;
    }
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tm, 512);
}

int main() {
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    long long* d; cudaMalloc(&d, 8);
    const int cases[][5] = {  // n1, n2, count1, count2, rowbytes
        {16, 32, 1152, 0, 64}, {16, 32, 0, 1152, 64}, {16, 32, 1152, 576, 64}, {16, 16, 1152, 576, 64}, {32, 32, 1152, 576, 64},
        {32, 32, 1152, 0, 128}, {32, 32, 1152, 576, 128}};      // (9 B boxes of n x rowbytes must stay below 48 KB)
    for (auto& c : cases) {
        for (int rep = 0; rep < 2; ++rep) probe<<<1, 128, 100 * 1024>>>(c[0], c[1], c[2], c[3], c[4], d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("rowbytes %3d  warp A: %4d MMAs N=%3d   warp B: %4d MMAs N=%3d : %lld cycles = %.1f cycles/MMA %s\n", c[4], c[2], c[0], c[3],
               c[1], h, (double)h / (c[2] + c[3]), e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
