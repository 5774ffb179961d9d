// ModelEMA.update (ultralytics/utils/torch_utils.py:514-524) for sm_100a: the reference walks the state_dict in Python —
// three ATen launches per floating-point tensor (`v *= d`, `(1 - d) * m`, `v += ...`), ~1 500 launches per training step
// for the 500 tensors of the spectrogram detector, all launch-latency bound.  Here ONE launch walks a chunk list over a
// device-resident pointer table ("multi-tensor apply"): chunk -> (tensor, offset), 16-byte vector accesses where the
// tensors allow it.  The arithmetic is the reference's, rounding for rounding: t1 = v * d, t2 = m * (1 - d), v = t1 + t2,
// each rounded to fp32 (no fused multiply-add), with d and (1 - d) converted to fp32 by the caller exactly as ATen converts
// the Python scalars.
#include "common.h"
#include "tma_host.h"

namespace specyolo {

__global__ void __launch_bounds__(256)
ema_update_kernel(float* const* __restrict__ ema, const float* const* __restrict__ model, const long long* __restrict__ numel,
                  const int* __restrict__ chunk_tensor, const long long* __restrict__ chunk_off, int chunk, float d, float omd) {
    const int t = chunk_tensor[blockIdx.x];
    const long long off = chunk_off[blockIdx.x];
    const long long n = min((long long)chunk, numel[t] - off);
    float* __restrict__ v = ema[t] + off;
    const float* __restrict__ m = model[t] + off;
    if ((((uintptr_t)v | (uintptr_t)m) & 15) == 0) {
        const long long n4 = n >> 2;
        for (long long i = threadIdx.x; i < n4; i += 256) {
            float4 a = reinterpret_cast<float4*>(v)[i];
            const float4 b = __ldg(reinterpret_cast<const float4*>(m) + i);
            a.x = __fadd_rn(__fmul_rn(a.x, d), __fmul_rn(b.x, omd));
            a.y = __fadd_rn(__fmul_rn(a.y, d), __fmul_rn(b.y, omd));
            a.z = __fadd_rn(__fmul_rn(a.z, d), __fmul_rn(b.z, omd));
            a.w = __fadd_rn(__fmul_rn(a.w, d), __fmul_rn(b.w, omd));
            reinterpret_cast<float4*>(v)[i] = a;
        }
        for (long long i = (n4 << 2) + threadIdx.x; i < n; i += 256) v[i] = __fadd_rn(__fmul_rn(v[i], d), __fmul_rn(m[i], omd));
    } else {
        for (long long i = threadIdx.x; i < n; i += 256) v[i] = __fadd_rn(__fmul_rn(v[i], d), __fmul_rn(m[i], omd));
    }
}

int ema_update_launch(float* const* ema, const float* const* model, const long long* numel, const int* chunk_tensor,
                      const long long* chunk_off, int nchunks, int chunk, float d, float one_minus_d, cudaStream_t stream) {
    SY_CHECK(nchunks >= 0 && chunk >= 4 && chunk % 4 == 0, SPECYOLO_ERR_INVALID, "ema: chunk must be a positive multiple of 4");
    if (nchunks == 0) return SPECYOLO_OK;
    ema_update_kernel<<<(unsigned)nchunks, 256, 0, stream>>>(ema, model, numel, chunk_tensor, chunk_off, chunk, d, one_minus_d);
    SY_LAUNCH_CHECK();
    count_launch();
    return SPECYOLO_OK;
}

}  // namespace specyolo
