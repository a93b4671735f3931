// dtlr_b200 -- decoder self-attention (nn.MultiheadAttention(256, 8) of reference deformable_transformer.py:847,
// 903-905: q = k = tgt + query_pos, v = tgt, optional boolean attn_mask with True = blocked).
//
// Exact-fp32 SIMT flash-style kernel: one thread owns one query row of one head (q, running max/sum and the 32-wide
// output accumulator live in registers), keys/values stream through shared memory in tiles; scores never touch HBM
// (the reference materialises (B*8, Q, Q) fp32 scores = 26 MB per image per layer).  Used by both dtypes in round 1;
// the bf16 tensor-core version is the next optimisation step.
#include <cuda_fp16.h>

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


// ---------------------------------------------------------------------------------------------- bf16 tensor-core path
// Flash-style attention for the unmasked inference case: one CTA = (image, head, half of the queries); K and V of
// the head (Q x 32 bf16 each) are staged once in shared memory (80-byte row pitch: conflict-free ldmatrix), each warp
// owns 16 queries at a time and sweeps the keys 64 at a time: S = Q K^T and O += P V on mma.sync m16n8k16 (bf16 in,
// fp32 accumulate), online softmax in registers with exp2f.  Scores never leave the register file.
constexpr int FA_WARPS = 16;
constexpr int FA_PITCH = 40;   // bf16 elements per smem row (32 + 8 padding)

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}

// The softmax arithmetic, not the tensor pipe, bounds this kernel (about 8 scalar instructions per query-key pair in a plain
// fp32 formulation), so the elementwise part runs on packed half2: scores -> half2, running max with HMNMX2, exponent argument
// with one HFMA2 and ex2.approx.f16x2, and the probabilities ARE the A fragments of the P.V MMA (V is staged as f16).  The row
// sums come out of the tensor cores too: V carries a ninth "ones" column block, so O[:, 32] = sum_k P[:, k].
__device__ __forceinline__ void mma_f16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t (&r)[2], const void* smem_row) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ uint32_t h2_as_u32(const __half2 v) { return *reinterpret_cast<const uint32_t*>(&v); }

__global__ void __launch_bounds__(FA_WARPS * 32)
mha_flash_bf16_kernel(const __nv_bfloat16* __restrict__ qk, int ld_qk, int k_off, const __nv_bfloat16* __restrict__ v, int ld_v,
                      __nv_bfloat16* __restrict__ out, int ld_o, int Q, int q_per_cta, float scale_log2) {
    extern __shared__ __align__(16) unsigned char fa_smem[];
    const int KP = (Q + 63) / 64 * 64;                        // keys padded to the 64-key sweep
    __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(fa_smem);
    __half* Vs = reinterpret_cast<__half*>(Ks + (size_t)KP * FA_PITCH);      // [KP][40]: 32 dims + ones column block
    const int b = blockIdx.z, h = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t row0 = (size_t)b * Q;

    // ---- stage K (bf16, cp.async) and V (converted to f16, plus the ones column) of this (image, head)
    for (int i = tid; i < KP * 4; i += FA_WARPS * 32) {
        const int r = i >> 2, c = i & 3;
        if (r < Q) cp_async16(Ks + (size_t)r * FA_PITCH + c * 8, qk + (row0 + r) * ld_qk + k_off + h * 32 + c * 8);
        else *reinterpret_cast<uint4*>(Ks + (size_t)r * FA_PITCH + c * 8) = make_uint4(0, 0, 0, 0);
    }
    cp_async_commit();
    for (int i = tid; i < KP * 5; i += FA_WARPS * 32) {
        const int r = i / 5, c = i - r * 5;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (r < Q) {
            if (c < 4) {
                const uint4 t = *reinterpret_cast<const uint4*>(v + (row0 + r) * ld_v + h * 32 + c * 8);
                const uint32_t w[4] = {t.x, t.y, t.z, t.w};
                uint32_t hh[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    hh[k] = h2_as_u32(__floats2half2_rn(__uint_as_float(w[k] << 16), __uint_as_float(w[k] & 0xffff0000u)));
                o = make_uint4(hh[0], hh[1], hh[2], hh[3]);
            } else {
                o.x = h2_as_u32(__floats2half2_rn(1.f, 0.f));
            }
        }
        *reinterpret_cast<uint4*>(Vs + (size_t)r * FA_PITCH + c * 8) = o;
    }
    cp_async_wait_all();
    __syncthreads();

    const int g = lane >> 2, t = lane & 3;
    const int q_begin = blockIdx.x * q_per_cta;
    const int q_end = min(Q, q_begin + q_per_cta);
    const int k_row = lane & 7, k_chunk = lane >> 3;
    const int v_row = (lane & 7) + ((lane >> 3) & 1) * 8, v_chunk = lane >> 4;
    const __half2 scale_h2 = __float2half2_rn(scale_log2);
    const __half2 ninf2 = __float2half2_rn(-INFINITY);

    for (int q0 = q_begin + warp * 16; q0 < q_end; q0 += FA_WARPS * 16) {
        uint32_t qa[2][4];
        {
            const int r_lo = min(q0 + g, Q - 1), r_hi = min(q0 + g + 8, Q - 1);
            const __nv_bfloat16* qlo = qk + (row0 + r_lo) * ld_qk + h * 32;
            const __nv_bfloat16* qhi = qk + (row0 + r_hi) * ld_qk + h * 32;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                qa[ks][0] = *reinterpret_cast<const uint32_t*>(qlo + ks * 16 + 2 * t);
                qa[ks][1] = *reinterpret_cast<const uint32_t*>(qhi + ks * 16 + 2 * t);
                qa[ks][2] = *reinterpret_cast<const uint32_t*>(qlo + ks * 16 + 8 + 2 * t);
                qa[ks][3] = *reinterpret_cast<const uint32_t*>(qhi + ks * 16 + 8 + 2 * t);
            }
        }
        float o[5][4];                                   // 4 n-blocks of head dims + the row-sum block
#pragma unroll
        for (int n = 0; n < 5; ++n)
#pragma unroll
            for (int k = 0; k < 4; ++k) o[n][k] = 0.f;
        // running maxima in scaled log2 units, kept at half precision so that the exponent offset is exactly the one applied
        float m_lo = -INFINITY, m_hi = -INFINITY;

        for (int kb = 0; kb < KP; kb += 64) {
            // ---- S = Q K^T for 64 keys
            float sc[8][4];
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                uint32_t kf[4];
                ldmatrix_x4(kf, Ks + (size_t)(kb + n * 8 + k_row) * FA_PITCH + k_chunk * 8);
                sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
                mma_bf16_16816(sc[n], qa[0], kf[0], kf[1]);
                mma_bf16_16816(sc[n], qa[1], kf[2], kf[3]);
            }
            // ---- to packed half: s_lo[n] = row g keys (2t, 2t+1) of n-block n, s_hi[n] = row g+8
            __half2 s_lo[8], s_hi[8];
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                s_lo[n] = __floats2half2_rn(sc[n][0], sc[n][1]);
                s_hi[n] = __floats2half2_rn(sc[n][2], sc[n][3]);
            }
            if (kb + 64 > Q) {    // key padding -> -inf
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const int key = kb + n * 8 + 2 * t;
                    if (key >= Q) { s_lo[n] = ninf2; s_hi[n] = ninf2; }
                    else if (key + 1 >= Q) { s_lo[n] = __halves2half2(__low2half(s_lo[n]), __high2half(ninf2)); s_hi[n] = __halves2half2(__low2half(s_hi[n]), __high2half(ninf2)); }
                }
            }
            // ---- block maxima (packed), quad reduce, new running max
            __half2 bm_lo = s_lo[0], bm_hi = s_hi[0];
#pragma unroll
            for (int n = 1; n < 8; ++n) { bm_lo = __hmax2(bm_lo, s_lo[n]); bm_hi = __hmax2(bm_hi, s_hi[n]); }
            float bl = fmaxf(__low2float(bm_lo), __high2float(bm_lo)), bh = fmaxf(__low2float(bm_hi), __high2float(bm_hi));
            bl = fmaxf(bl, __shfl_xor_sync(0xffffffffu, bl, 1));
            bl = fmaxf(bl, __shfl_xor_sync(0xffffffffu, bl, 2));
            bh = fmaxf(bh, __shfl_xor_sync(0xffffffffu, bh, 1));
            bh = fmaxf(bh, __shfl_xor_sync(0xffffffffu, bh, 2));
            const float mn_lo = __half2float(__float2half_rn(fmaxf(m_lo, bl * scale_log2)));
            const float mn_hi = __half2float(__float2half_rn(fmaxf(m_hi, bh * scale_log2)));
            const float c_lo = exp2f(m_lo - mn_lo), c_hi = exp2f(m_hi - mn_hi);      // 0 on the first block (m = -inf)
            m_lo = mn_lo; m_hi = mn_hi;
#pragma unroll
            for (int n = 0; n < 5; ++n) { o[n][0] *= c_lo; o[n][1] *= c_lo; o[n][2] *= c_hi; o[n][3] *= c_hi; }
            const __half2 nm_lo = __float2half2_rn(-mn_lo), nm_hi = __float2half2_rn(-mn_hi);
            // ---- P = exp2(s*scale - m) as packed half == A fragments of the P.V MMA
            uint32_t pa[4][4];
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const __half2 p_lo = h2exp2(__hfma2(s_lo[n], scale_h2, nm_lo));
                const __half2 p_hi = h2exp2(__hfma2(s_hi[n], scale_h2, nm_hi));
                pa[n >> 1][(n & 1) * 2] = h2_as_u32(p_lo);
                pa[n >> 1][(n & 1) * 2 + 1] = h2_as_u32(p_hi);
            }
            // ---- O += P V (4 head-dim blocks) and row sums (ones block)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                for (int nn = 0; nn < 2; ++nn) {
                    uint32_t vf[4];
                    ldmatrix_x4_trans(vf, Vs + (size_t)(kb + kk * 16 + v_row) * FA_PITCH + (nn * 2 + v_chunk) * 8);
                    mma_f16_16816(o[nn * 2], pa[kk], vf[0], vf[1]);
                    mma_f16_16816(o[nn * 2 + 1], pa[kk], vf[2], vf[3]);
                }
                uint32_t v1[2];
                ldmatrix_x2_trans(v1, Vs + (size_t)(kb + kk * 16 + (lane & 15)) * FA_PITCH + 32);
                mma_f16_16816(o[4], pa[kk], v1[0], v1[1]);
            }
        }
        // row sums sit in column 32 = element [0] / [2] of the t == 0 lane of each quad
        const float l_lo = __shfl_sync(0xffffffffu, o[4][0], lane & ~3);
        const float l_hi = __shfl_sync(0xffffffffu, o[4][2], lane & ~3);
        const float i_lo = 1.f / l_lo, i_hi = 1.f / l_hi;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            if (q0 + g < q_end)
                *reinterpret_cast<uint32_t*>(out + (row0 + q0 + g) * ld_o + h * 32 + n * 8 + 2 * t) = pack_bf16(o[n][0] * i_lo, o[n][1] * i_lo);
            if (q0 + g + 8 < q_end)
                *reinterpret_cast<uint32_t*>(out + (row0 + q0 + g + 8) * ld_o + h * 32 + n * 8 + 2 * t) = pack_bf16(o[n][2] * i_hi, o[n][3] * i_hi);
        }
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
    else if (dtype == DTLR_BF16 && !attn_mask && (ld_qk % 8) == 0 && (ld_v % 8) == 0 && (k_off % 8) == 0 && (ld_o % 2) == 0 &&
             (size_t)((Q + 63) / 64 * 64) * FA_PITCH * 2 * 2 <= (size_t)max_smem_optin()) {
        const int KP = (Q + 63) / 64 * 64;
        const size_t smem = (size_t)KP * FA_PITCH * 2 * 2;
        // one round of 16-query tiles per CTA: ceil(Q / 256) CTAs per (image, head), each staging K and V once
        const int per_round = FA_WARPS * 16;
        const int splits = (Q + per_round - 1) / per_round;
        int q_per_cta = ((Q + splits - 1) / splits + 15) / 16 * 16;
        DTLR_CHECK_CUDA(cudaFuncSetAttribute(mha_flash_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 fgrid((Q + q_per_cta - 1) / q_per_cta, heads, B);
        mha_flash_bf16_kernel<<<fgrid, FA_WARPS * 32, smem, st>>>((const __nv_bfloat16*)qk, ld_qk, k_off, (const __nv_bfloat16*)v, ld_v,
                                                                  (__nv_bfloat16*)out, ld_o, Q, q_per_cta, scale * 1.4426950408889634f);
    } else if (dtype == DTLR_BF16)
        mha_simt_kernel<__nv_bfloat16><<<grid, ATT_QT, 0, st>>>((const __nv_bfloat16*)qk, ld_qk, k_off, (const __nv_bfloat16*)v, ld_v, attn_mask, (__nv_bfloat16*)out, ld_o, Q, scale);
    else { set_error("mha: unsupported dtype"); return DTLR_ERR_INVALID; }
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}
