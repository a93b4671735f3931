// dtlr_b200 -- decoder self-attention (nn.MultiheadAttention(256, 8) of reference deformable_transformer.py:847,
// 903-905: q = k = tgt + query_pos, v = tgt, optional boolean attn_mask with True = blocked).
//
// Exact-fp32 SIMT flash-style kernel: one thread owns one query row of one head (q, running max/sum and the 32-wide
// output accumulator live in registers), keys/values stream through shared memory in tiles; scores never touch HBM
// (the reference materialises (B*8, Q, Q) fp32 scores = 26 MB per image per layer).  Used by both dtypes in round 1;
// the bf16 tensor-core version is the next optimisation step.
#include "common.cuh"

namespace dtlr {

template <typename T> __device__ __forceinline__ float ldf_(const T* p);
template <> __device__ __forceinline__ float ldf_<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf_<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf_(T* p, float v);
template <> __device__ __forceinline__ void stf_<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf_<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

constexpr int ATT_DH = 32;
constexpr int ATT_QT = 128;   // queries per CTA (one per thread)
constexpr int ATT_KT = 64;    // keys per shared-memory tile

// q,k: rows of `qk` [B*Q, ld_qk] at column offsets h*32 (q) and k_off + h*32 (k); v rows of `v` [B*Q, ld_v] at h*32.
template <typename T>
__global__ void __launch_bounds__(ATT_QT)
mha_simt_kernel(const T* __restrict__ qk, int ld_qk, int k_off, const T* __restrict__ v, int ld_v,
                const unsigned char* __restrict__ mask, T* __restrict__ out, int ld_o, int Q, float scale) {
    __shared__ float Ks[ATT_KT][ATT_DH];
    __shared__ float Vs[ATT_KT][ATT_DH];
    const int b = blockIdx.z, h = blockIdx.y;
    const int qi = blockIdx.x * ATT_QT + threadIdx.x;
    const bool active = qi < Q;
    float q[ATT_DH], acc[ATT_DH];
    float m = -INFINITY, l = 0.f;
    const size_t rowq = (size_t)b * Q + (active ? qi : 0);
#pragma unroll
    for (int d = 0; d < ATT_DH; ++d) {
        q[d] = ldf_<T>(qk + rowq * ld_qk + h * ATT_DH + d) * scale;
        acc[d] = 0.f;
    }
    for (int k0 = 0; k0 < Q; k0 += ATT_KT) {
        __syncthreads();
        for (int i = threadIdx.x; i < ATT_KT * ATT_DH; i += ATT_QT) {
            const int kr = i / ATT_DH, d = i % ATT_DH;
            const int kk = k0 + kr;
            float kv = 0.f, vv = 0.f;
            if (kk < Q) {
                const size_t rk = (size_t)b * Q + kk;
                kv = ldf_<T>(qk + rk * ld_qk + k_off + h * ATT_DH + d);
                vv = ldf_<T>(v + rk * ld_v + h * ATT_DH + d);
            }
            Ks[kr][d] = kv;
            Vs[kr][d] = vv;
        }
        __syncthreads();
        const int kn = min(ATT_KT, Q - k0);
        for (int kr = 0; kr < kn; ++kr) {
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < ATT_DH; ++d) s = fmaf(q[d], Ks[kr][d], s);
            if (mask && active && mask[(size_t)qi * Q + k0 + kr]) s = -INFINITY;
            if (s > m) {
                const float c = __expf(m - s);     // m = -inf on the first key -> c = 0
                l *= c;
#pragma unroll
                for (int d = 0; d < ATT_DH; ++d) acc[d] *= c;
                m = s;
            }
            const float p = (s == -INFINITY) ? 0.f : __expf(s - m);
            l += p;
#pragma unroll
            for (int d = 0; d < ATT_DH; ++d) acc[d] = fmaf(p, Vs[kr][d], acc[d]);
        }
    }
    if (active) {
        const float inv = 1.f / l;
#pragma unroll
        for (int d = 0; d < ATT_DH; ++d) stf_<T>(out + rowq * ld_o + h * ATT_DH + d, acc[d] * inv);
    }
}

}  // namespace dtlr

using namespace dtlr;

extern "C" int dtlr_mha_self_attention(const void* qk, int ld_qk, int k_off, const void* v, int ld_v,
                                       const unsigned char* attn_mask, void* out, int ld_o, int B, int Q, int heads,
                                       int head_dim, int dtype, void* stream) {
    DTLR_CHECK_ARG(head_dim == ATT_DH, "mha: head_dim must be 32 (d_model 256 / 8 heads), got %d", head_dim);
    DTLR_CHECK_ARG(B >= 0 && Q >= 0 && heads > 0, "mha: bad sizes");
    if (B == 0 || Q == 0) return DTLR_OK;
    dim3 grid((Q + ATT_QT - 1) / ATT_QT, heads, B);
    const float scale = 1.0f / sqrtf((float)head_dim);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DTLR_F32)
        mha_simt_kernel<float><<<grid, ATT_QT, 0, st>>>((const float*)qk, ld_qk, k_off, (const float*)v, ld_v, attn_mask, (float*)out, ld_o, Q, scale);
    else if (dtype == DTLR_BF16)
        mha_simt_kernel<__nv_bfloat16><<<grid, ATT_QT, 0, st>>>((const __nv_bfloat16*)qk, ld_qk, k_off, (const __nv_bfloat16*)v, ld_v, attn_mask, (__nv_bfloat16*)out, ld_o, Q, scale);
    else { set_error("mha: unsupported dtype"); return DTLR_ERR_INVALID; }
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}
