// Probe: can TWO warps feed the tensor pipe concurrently?  tcgen05.mma (M=128, N=n, K=16, bf16, SS) issued by one warp vs
// by two warps (different accumulators), each MMA followed by `pad` dependent uniform ALU ops (the descriptor arithmetic of
// a real conv issue loop).  If a UTCHMMA blocks its warp for the MMA's duration, a single issuer pays MMA + overhead per
// MMA while two issuers hide each other's overhead.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include "../../spectrogram-yolov11_b200/csrc/ptx.cuh"
using namespace specyolo;

__global__ void __launch_bounds__(128) probe(int n, int issuers, int pad, int total, long long* out, uint32_t seed) {
    extern __shared__ uint8_t raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_s;
    const uint32_t ra = ptx::smem_u32(raw);
    uint8_t* base = raw + (((ra + 1023u) & ~1023u) - ra);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(base)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { ptx::mbar_init(&mbar, issuers); ptx::fence_mbar_init(); }
    if (warp == 0) ptx::tmem_alloc(&tmem_s, 512);   // two accumulators of up to 256 columns
    ptx::fence_proxy_async();
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tm = tmem_s;
    if (warp >= 1 && warp <= issuers) {
        const bool leader = ptx::elect_one();
        const uint32_t idesc = ptx::umma_idesc_bf16(128, n);
        const uint64_t hi = ((uint64_t)((8 * 128) >> 4) << 32) | (1ull << 46) | (2ull << 61);
        uint32_t a16 = ptx::smem_u32(base) >> 4;
        const uint32_t b16 = (ptx::smem_u32(base) + 48 * 1024) >> 4;
        const uint32_t d = tm + (uint32_t)(warp - 1) * 256u;
        const int mine = total / issuers;
        uint32_t x = seed;                        // kernel parameter: warp-uniform, stays on the uniform datapath
        const long long t0 = clock64();
        for (int i = 0; i < mine; i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
#pragma unroll
                for (int q = 0; q < 16; ++q) if (q < pad) x = (x + 0x9E3779B9u) ^ (x >> 7);   // dependent uniform ALU chain (2 ops per step)
                if (leader) ptx::umma_bf16(d, hi | (uint64_t)(a16 + ((x >> 31) << 1)), hi | (uint64_t)b16, idesc, 1u);
            }
        }
        if (leader) ptx::umma_commit(&mbar);
        ptx::mbar_wait(&mbar, 0);
        if (leader && warp == 1) out[0] = clock64() - t0;
    }
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tm, 512);
}

int main() {
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    long long* d; cudaMalloc(&d, 8);
    const int total = 1024;
    for (int n : {16, 64, 128, 256})
        for (int pad : {0, 4, 8, 16})
            for (int issuers : {1, 2}) {
                for (int rep = 0; rep < 2; ++rep) probe<<<1, 128, 100 * 1024>>>(n, issuers, pad, total, d, 12345u);
                cudaError_t e = cudaDeviceSynchronize();
                long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                printf("N %3d pad %2d issuers %d : %.1f cycles/MMA %s\n", n, pad, issuers, (double)h / total,
                       e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    return 0;
}
