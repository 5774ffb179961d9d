// PSA attention core of C2PSA (ultralytics/nn/modules/block.py:1922-1933):
//   q,k,v = qkv.view(B, heads, 2*kd+hd, N).split([kd,kd,hd], 2)
//   attn = softmax((q^T k) * scale, -1);  x = v @ attn^T  + pe(v)      (pe = depthwise 3x3 + folded BN)
// The reference materialises the B x heads x N x N matrix through two cuBLAS bmm and a softmax
// kernel; here the scores never leave the SM (online softmax, flash style).
//
// v1 maps the work onto CUDA cores: N is 400 (640^2) or 1600 (1280^2) tokens with kd = 32, hd = 64,
// i.e. 0.12 GFLOP of the 19 GFLOP per image.  CTA = 64 queries of one (image, head); 4 threads share
// a query, each walking a quarter of the keys of every 64-key tile staged in shared memory, with
// private running (max, sum, acc[hd]) merged by shuffles at the end.
#include "common.h"

namespace specyolo {

static constexpr int kAttQ = 64;    // queries per CTA
static constexpr int kAttK = 64;    // keys per smem tile
static constexpr int KD = 32, HD = 64;

__global__ void __launch_bounds__(256)
psa_attention_kernel(const __nv_bfloat16* __restrict__ qkv, int qkv_pixstride, int H, int W, int heads,
                     float scale_log2e, const float* __restrict__ pe_w, const float* __restrict__ pe_b,
                     __nv_bfloat16* __restrict__ out, int out_pixstride) {
    const int N = H * W;
    const int b = blockIdx.z, head = blockIdx.y;
    const int q0 = blockIdx.x * kAttQ;
    const int tid = threadIdx.x;
    const int ql = tid >> 2;       // local query
    const int part = tid & 3;      // key quarter
    const int qi = q0 + ql;
    const int per_head = 2 * KD + HD;
    const __nv_bfloat16* base = qkv + (size_t)b * N * qkv_pixstride + head * per_head;

    __shared__ __align__(16) __nv_bfloat16 sK[kAttK][KD + 8];   // +8 pad: rows 80 B apart
    __shared__ __align__(16) __nv_bfloat16 sV[kAttK][HD + 8];

    // query row in registers, pre-scaled by scale*log2(e) so that exp2f can be used
    float q[KD];
    if (qi < N) {
        const uint4* qp = reinterpret_cast<const uint4*>(base + (size_t)qi * qkv_pixstride);
#pragma unroll
        for (int i = 0; i < KD / 8; ++i) {
            const uint4 u = __ldg(qp + i);
            const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16x2(uu[j]);
                q[i * 8 + 2 * j] = f.x * scale_log2e;
                q[i * 8 + 2 * j + 1] = f.y * scale_log2e;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < KD; ++i) q[i] = 0.f;
    }

    float m = -INFINITY, l = 0.f;
    float acc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = 0.f;

    for (int k0 = 0; k0 < N; k0 += kAttK) {
        __syncthreads();
        // stage K (64 x 32) and V (64 x 64) tiles: 16-byte vectors
        for (int i = tid; i < kAttK * (KD / 8); i += 256) {
            const int r = i / (KD / 8), c = i % (KD / 8);
            uint4 u = make_uint4(0, 0, 0, 0);
            if (k0 + r < N) u = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(k0 + r) * qkv_pixstride + KD) + c);
            *reinterpret_cast<uint4*>(&sK[r][c * 8]) = u;
        }
        for (int i = tid; i < kAttK * (HD / 8); i += 256) {
            const int r = i / (HD / 8), c = i % (HD / 8);
            uint4 u = make_uint4(0, 0, 0, 0);
            if (k0 + r < N) u = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(k0 + r) * qkv_pixstride + 2 * KD) + c);
            *reinterpret_cast<uint4*>(&sV[r][c * 8]) = u;
        }
        __syncthreads();
        const int kend = min(kAttK, N - k0);
        for (int j = part; j < kend; j += 4) {
            float s = 0.f;
            const __nv_bfloat162* kr = reinterpret_cast<const __nv_bfloat162*>(&sK[j][0]);
#pragma unroll
            for (int i = 0; i < KD / 2; ++i) {
                const float2 f = __bfloat1622float2(kr[i]);
                s = fmaf(q[2 * i], f.x, s);
                s = fmaf(q[2 * i + 1], f.y, s);
            }
            const float mn = fmaxf(m, s);
            const float corr = exp2f(m - mn);   // m = -inf on first key -> 0
            const float pj = exp2f(s - mn);
            l = l * corr + pj;
            const __nv_bfloat162* vr = reinterpret_cast<const __nv_bfloat162*>(&sV[j][0]);
#pragma unroll
            for (int d = 0; d < HD / 2; ++d) {
                const float2 f = __bfloat1622float2(vr[d]);
                acc[2 * d] = fmaf(acc[2 * d], corr, pj * f.x);
                acc[2 * d + 1] = fmaf(acc[2 * d + 1], corr, pj * f.y);
            }
            m = mn;
        }
    }
    // merge the 4 partial softmax states of a query (lanes 4q..4q+3)
#pragma unroll
    for (int off = 1; off < 4; off <<= 1) {
        const float mo = __shfl_xor_sync(0xffffffffu, m, off);
        const float lo = __shfl_xor_sync(0xffffffffu, l, off);
        const float mn = fmaxf(m, mo);
        const float c0 = (m == -INFINITY) ? 0.f : exp2f(m - mn);
        const float c1 = (mo == -INFINITY) ? 0.f : exp2f(mo - mn);
        l = l * c0 + lo * c1;
#pragma unroll
        for (int d = 0; d < HD; ++d) {
            const float ao = __shfl_xor_sync(0xffffffffu, acc[d], off);
            acc[d] = acc[d] * c0 + ao * c1;
        }
        m = mn;
    }
    if (qi >= N) return;
    // each of the 4 threads finishes 16 of the 64 head channels: + pe(v) (depthwise 3x3) and store
    const float inv_l = 1.0f / l;
    const int h = qi / W, w = qi % W;
    const int dbase = part * 16;
    float res[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) {
        // static indexing of acc[] requires the unrolled select below
        float a = 0.f;
#pragma unroll
        for (int pp = 0; pp < 4; ++pp)
            if (pp == part) a = acc[pp * 16 + d];
        res[d] = a * inv_l + __ldg(pe_b + head * HD + dbase + d);
    }
    for (int ky = 0; ky < 3; ++ky) {
        const int hh = h + ky - 1;
        if (hh < 0 || hh >= H) continue;
        for (int kx = 0; kx < 3; ++kx) {
            const int ww = w + kx - 1;
            if (ww < 0 || ww >= W) continue;
            const uint4* vp = reinterpret_cast<const uint4*>(base + (size_t)(hh * W + ww) * qkv_pixstride + 2 * KD + dbase);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const uint4 u = __ldg(vp + half);
                const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = unpack_bf16x2(uu[j]);
                    const int d = half * 8 + 2 * j;
                    res[d] = fmaf(f.x, __ldg(pe_w + (head * HD + dbase + d) * 9 + ky * 3 + kx), res[d]);
                    res[d + 1] = fmaf(f.y, __ldg(pe_w + (head * HD + dbase + d + 1) * 9 + ky * 3 + kx), res[d + 1]);
                }
            }
        }
    }
    __nv_bfloat16* op = out + ((size_t)b * N + qi) * out_pixstride + head * HD + dbase;
    uint4 o0, o1;
    o0.x = pack_bf16x2(res[0], res[1]);   o0.y = pack_bf16x2(res[2], res[3]);
    o0.z = pack_bf16x2(res[4], res[5]);   o0.w = pack_bf16x2(res[6], res[7]);
    o1.x = pack_bf16x2(res[8], res[9]);   o1.y = pack_bf16x2(res[10], res[11]);
    o1.z = pack_bf16x2(res[12], res[13]); o1.w = pack_bf16x2(res[14], res[15]);
    reinterpret_cast<uint4*>(op)[0] = o0;
    reinterpret_cast<uint4*>(op)[1] = o1;
}

int psa_attention_launch(const void* qkv, int qkv_pixstride, int B, int H, int W, int heads, int key_dim,
                         int head_dim, float scale, const float* pe_w, const float* pe_b, void* out,
                         int out_pixstride, cudaStream_t stream) {
    SY_CHECK(key_dim == KD && head_dim == HD, SPECYOLO_ERR_UNSUPPORTED,
             "psa attention supports key_dim=32, head_dim=64 (got %d, %d)", key_dim, head_dim);
    SY_CHECK(qkv_pixstride % 8 == 0 && out_pixstride % 8 == 0 &&
                 ((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
             SPECYOLO_ERR_INVALID, "psa attention: tensors must be 16-byte aligned");
    const int N = H * W;
    dim3 grid((unsigned)ceil_div(N, kAttQ), (unsigned)heads, (unsigned)B);
    const float scale_log2e = scale * 1.4426950408889634f;
    psa_attention_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(qkv), qkv_pixstride, H, W,
                                                   heads, scale_log2e, pe_w, pe_b,
                                                   reinterpret_cast<__nv_bfloat16*>(out), out_pixstride);
    SY_LAUNCH_CHECK();
    count_launch();
    return SPECYOLO_OK;
}

}  // namespace specyolo
