// Probe: tcgen05.mma throughput (M=128, N=n, K=16, bf16, SS) with a warp-uniform issue loop (descriptors in uniform
// registers, elected lane issues), unrolled x8.  cycles/MMA = (commit arrival - start) / count.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include "../../spectrogram-yolov11_b200/csrc/ptx.cuh"
using namespace specyolo;

__global__ void __launch_bounds__(128) probe(int sbo_rows, int rowbytes, int n, int astep16, long long* out, int bstep16 = 0, int commit_every = 0, int two_commits = 0) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ __align__(8) uint64_t dummy_bar[2];
    __shared__ uint32_t tmem_s;
    const uint32_t ra = ptx::smem_u32(raw);
    uint8_t* base = raw + (((ra + 1023u) & ~1023u) - ra);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { ptx::mbar_init(&mbar, 1); ptx::mbar_init(&dummy_bar[0], 100000); ptx::mbar_init(&dummy_bar[1], 100000); ptx::fence_mbar_init(); }
    if (warp == 0) ptx::tmem_alloc(&tmem_s, 256);
    ptx::fence_proxy_async();
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tm = tmem_s;
    if (warp == 1) {
        const bool leader = ptx::elect_one();
        const uint32_t idesc = ptx::umma_idesc_bf16(128, n);
        const uint64_t layout = rowbytes == 128 ? 2ull : (rowbytes == 64 ? 4ull : 6ull);
        const uint64_t hi_a = ((uint64_t)((sbo_rows * rowbytes) >> 4) << 32) | (1ull << 46) | (layout << 61);
        const uint64_t hi_b = ((uint64_t)((8 * rowbytes) >> 4) << 32) | (1ull << 46) | (layout << 61);
        const uint32_t a16 = ptx::smem_u32(base) >> 4;
        const uint32_t b16 = (ptx::smem_u32(base) + 48 * 1024) >> 4;
        const long long t0 = clock64();
        for (int i = 0; i < 64; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (leader) ptx::umma_bf16(tm, hi_a | (uint64_t)(a16 + j * astep16), hi_b | (uint64_t)(b16 + j * bstep16), idesc, 1u);
            if (commit_every && (i % commit_every) == commit_every - 1 && leader) {
                ptx::umma_commit(&dummy_bar[0]);
                if (two_commits) ptx::umma_commit(&dummy_bar[1]);
            }
        }
        if (leader) ptx::umma_commit(&mbar);
        ptx::mbar_wait(&mbar, 0);
        if (leader) out[0] = clock64() - t0;
    }
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tm, 256);
}

int main() {
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    long long* d; cudaMalloc(&d, 8);
    const int cases[][4] = {  // sbo_rows, rowbytes, n, astep16 (A start step between MMAs, 16-byte units)
        {8, 128, 16, 0}, {8, 128, 32, 0}, {8, 128, 64, 0}, {8, 128, 128, 0}, {8, 128, 256, 0},
        {10, 128, 16, 8}, {10, 128, 64, 8}, {10, 128, 128, 8}, {10, 128, 16, 2}, {10, 128, 64, 2},
        {8, 64, 16, 0}, {8, 64, 64, 0}, {10, 64, 64, 4}, {8, 32, 16, 0}, {8, 32, 64, 0}, {10, 32, 16, 2},
        // fused Bottleneck (bneck_pair.cu): 18-row pitch input box, 16-row pitch mid tile
        {18, 128, 32, 8}, {18, 128, 32, 2}, {18, 64, 16, 4}, {18, 64, 16, 2}, {16, 64, 64, 4}, {16, 32, 32, 2}, {16, 32, 32, 0},
        {18, 128, 32, 0}, {16, 64, 64, 0}, {18, 64, 16, 0}};
    for (auto& c : cases) {
        for (int rep = 0; rep < 2; ++rep) probe<<<1, 128, 100 * 1024>>>(c[0], c[1], c[2], c[3], d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("rowbytes %3d N %3d sbo_rows %2d astep16 %d : %.1f cycles/MMA %s\n", c[1], c[2], c[0], c[3], h / 512.0,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    // B operand changing from MMA to MMA (a different weight box per tap, as every k x k conv does)
    const int cases2[][5] = {{18, 128, 32, 8, 256}, {18, 64, 16, 4, 64}, {16, 64, 64, 4, 256}, {16, 32, 32, 2, 64}};
    for (auto& c : cases2) {
        for (int rep = 0; rep < 2; ++rep) probe<<<1, 128, 100 * 1024>>>(c[0], c[1], c[2], c[3], d, c[4]);
        cudaError_t e = cudaDeviceSynchronize();
        long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("rowbytes %3d N %3d sbo_rows %2d astep16 %d bstep16 %4d : %.1f cycles/MMA %s\n", c[1], c[2], c[0], c[3], c[4], h / 512.0,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    // tcgen05.commit every 8 / 16 / 32 MMAs (one or two barriers): does a commit drain the pipe?
    const int cases3[][3] = {{16, 1, 0}, {16, 1, 1}, {16, 2, 1}, {16, 4, 1}, {64, 1, 1}, {64, 4, 1}};   // n, commit_every (x8 MMAs), two
    for (auto& c : cases3) {
        for (int rep = 0; rep < 2; ++rep) probe<<<1, 128, 100 * 1024>>>(8, 64, c[0], 2, d, 0, c[1], c[2]);
        cudaError_t e = cudaDeviceSynchronize();
        long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("N %3d commit every %2d MMAs (%d barriers) : %.1f cycles/MMA %s\n", c[0], 8 * c[1], 1 + c[2], h / 512.0,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
