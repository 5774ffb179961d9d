// Host-side TMA helpers shared by the implicit-GEMM launchers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <mutex>
#include <cstdlib>

namespace specyolo {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency)
inline EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) ==
                cudaSuccess &&
            qres == cudaDriverEntryPointSuccess) {
            fn = reinterpret_cast<EncodeTiledFn>(sym);
        }
    });
    return fn;
}

inline CUtensorMapSwizzle swizzle_for(int row_bytes) {
    return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                            : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// debugging / A-B switches read from the environment (evaluated on every call: launch-time only, cheap)
inline bool env_flag(const char* name) {
    const char* e = std::getenv(name);
    return e && e[0] && e[0] != '0';
}

// Launch with programmatic dependent launch (PDL) allowed: the grid may be scheduled while the previous kernel of the
// stream drains (every CTA of it has executed griddepcontrol.launch_dependents or exited); the kernel must execute
// griddepcontrol.wait (ptx::grid_dep_wait) before it touches memory written or read by its predecessors.  Inside a
// CUDA-graph capture this becomes a programmatic edge.  ~110 kernels per step, most of them a few microseconds long:
// without it every launch pays the full drain + ramp-up of its neighbours.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = env_flag("SPECYOLO_NO_PDL") ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            n = 148;
    }
    return n;
}

}  // namespace specyolo
