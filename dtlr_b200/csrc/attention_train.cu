// dtlr_b200 -- decoder self-attention of the fine-tune step, forward WITH the denoising attention mask and the log-sum-exp the backward
// needs, and the flash backward (reference: nn.MultiheadAttention(256, 8) under autograd, deformable_transformer.py:847, 903-905, with
// the (Q', Q') boolean attn_mask of dn_components.py:121-141; the reference materialises (B*8, Q', Q') fp32 scores, their softmax and
// both gradients).  head_dim 32, 16-bit operands, fp32 accumulation, mma.sync m16n8k16 + ldmatrix like the inference kernel
// (csrc/attention.cu) -- scores / probabilities never reach HBM in either direction.
//
//   forward   CTA = (image, head, query range): K, V of the head staged once in shared memory; a warp owns 16 queries and sweeps the
//             keys 64 at a time; writes O and lse2[b,h,i] = log2(sum_j exp(s_ij)) (base-2 units, the scale folded in).
//   backward  recomputes P = 2^(s*scale*log2e - lse2) instead of loading it.  Two passes, no atomics:
//     dQ pass   CTA = (image, head, query range), K / V staged:  dP = dO V^T,  dS = P o (dP - D),  dQ = scale * dS K
//     dKV pass  CTA = (image, head, key range), Q / dO staged (all queries): works on the TRANSPOSED tiles (rows = keys):
//               S^T = K Q^T,  dP^T = V dO^T,  dV = P^T dO,  dK = scale * dS^T Q
//     D[b,h,i] = <dO_i, O_i> comes from a small pre-pass.
//   The mask is a bit matrix (1 = blocked): mask_bits[i][j / 32] for the query-major passes, maskT_bits[j][i / 32] for the key-major one.
#include "common.cuh"

namespace dtlr {
namespace attn_train {

constexpr int PITCH = 40;      // 16-bit elements per shared-memory row (32 + 8 padding: conflict-free ldmatrix)

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." DTLR_OP16_PTX "." DTLR_OP16_PTX ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const void* smem_row) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], const void* smem_row) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ uint32_t pk(float lo, float hi) {
    op16x2_t t = op16_pack2(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
// A-operand fragments (16 rows x 32 columns = 2 k-steps) of rows r_lo / r_hi of a row-major 16-bit matrix in global memory
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[2][4], const op16_t* lo, const op16_t* hi, const int t) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        a[ks][0] = *reinterpret_cast<const uint32_t*>(lo + ks * 16 + 2 * t);
        a[ks][1] = *reinterpret_cast<const uint32_t*>(hi + ks * 16 + 2 * t);
        a[ks][2] = *reinterpret_cast<const uint32_t*>(lo + ks * 16 + 8 + 2 * t);
        a[ks][3] = *reinterpret_cast<const uint32_t*>(hi + ks * 16 + 8 + 2 * t);
    }
}
// stage `rows` rows of 32 16-bit channels (row r at src + r*ld) into shared memory with the padded pitch, zero rows up to rows_pad
__device__ __forceinline__ void stage_rows(op16_t* dst, const op16_t* src, const size_t ld, const int rows, const int rows_pad, const int tid,
                                           const int nthr) {
    for (int i = tid; i < rows_pad * 4; i += nthr) {
        const int r = i >> 2, c = i & 3;
        if (r < rows) cp_async16(dst + (size_t)r * PITCH + c * 8, src + (size_t)r * ld + c * 8);
        else *reinterpret_cast<uint4*>(dst + (size_t)r * PITCH + c * 8) = make_uint4(0, 0, 0, 0);
    }
}

// ---------------------------------------------------------------------------------------------- forward (mask, lse)
template <int MAXW>
__global__ void __launch_bounds__(MAXW * 32)
sa_fwd_kernel(const op16_t* __restrict__ qk, int ld_qk, int k_off, const op16_t* __restrict__ v, int ld_v,
              const uint32_t* __restrict__ mask_bits, op16_t* __restrict__ out, int ld_o, float* __restrict__ lse2, int Q, int q_per_cta,
              float scale_log2) {
    extern __shared__ __align__(16) unsigned char smem_[];
    const int KP = (Q + 63) / 64 * 64, MW = KP / 32;
    op16_t* Ks = reinterpret_cast<op16_t*>(smem_);
    op16_t* Vs = Ks + (size_t)KP * PITCH;
    const int b = blockIdx.z, h = blockIdx.y, H = gridDim.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x, nwarps = blockDim.x >> 5;
    const size_t row0 = (size_t)b * Q;
    pdl_launch_dependents();
    pdl_wait();
    stage_rows(Ks, qk + row0 * ld_qk + k_off + h * 32, ld_qk, Q, KP, tid, nthr);
    stage_rows(Vs, v + row0 * ld_v + h * 32, ld_v, Q, KP, tid, nthr);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    const int g = lane >> 2, t = lane & 3;
    const int q_begin = blockIdx.x * q_per_cta, q_end = min(Q, q_begin + q_per_cta);
    const int k_row = lane & 7, k_chunk = lane >> 3;
    const int v_row = (lane & 7) + ((lane >> 3) & 1) * 8, v_chunk = lane >> 4;
    for (int q0 = q_begin + warp * 16; q0 < q_end; q0 += nwarps * 16) {
        const int r_lo = min(q0 + g, Q - 1), r_hi = min(q0 + g + 8, Q - 1);
        uint32_t qa[2][4];
        load_a_frags(qa, qk + (row0 + r_lo) * ld_qk + h * 32, qk + (row0 + r_hi) * ld_qk + h * 32, t);
        float o[4][4];
#pragma unroll
        for (int n = 0; n < 4; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
        float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
        for (int kb = 0; kb < KP; kb += 64) {
            float sc[8][4];
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                uint32_t kf[4];
                ldsm4(kf, Ks + (size_t)(kb + n * 8 + k_row) * PITCH + k_chunk * 8);
                sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
                mma16816(sc[n], qa[0], kf[0], kf[1]);
                mma16816(sc[n], qa[1], kf[2], kf[3]);
            }
            if (kb + 64 > Q) {
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const int key = kb + n * 8 + 2 * t;
                    if (key >= Q) { sc[n][0] = -INFINITY; sc[n][2] = -INFINITY; }
                    if (key + 1 >= Q) { sc[n][1] = -INFINITY; sc[n][3] = -INFINITY; }
                }
            }
            if (mask_bits) {
                const uint2 wl = *reinterpret_cast<const uint2*>(mask_bits + (size_t)r_lo * MW + (kb >> 5));
                const uint2 wh = *reinterpret_cast<const uint2*>(mask_bits + (size_t)r_hi * MW + (kb >> 5));
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const uint32_t a = (n < 4 ? wl.x : wl.y) >> ((n & 3) * 8 + 2 * t), c = (n < 4 ? wh.x : wh.y) >> ((n & 3) * 8 + 2 * t);
                    if (a & 1u) sc[n][0] = -INFINITY;
                    if (a & 2u) sc[n][1] = -INFINITY;
                    if (c & 1u) sc[n][2] = -INFINITY;
                    if (c & 2u) sc[n][3] = -INFINITY;
                }
            }
            float mx_lo = m_lo, mx_hi = m_hi;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                mx_lo = fmaxf(mx_lo, fmaxf(sc[n][0], sc[n][1]));
                mx_hi = fmaxf(mx_hi, fmaxf(sc[n][2], sc[n][3]));
            }
            mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
            mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
            mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
            mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
            // a row whose keys so far are all blocked keeps m = -inf: use 0 as the reference so that no (-inf) - (-inf) appears
            const float rl = mx_lo == -INFINITY ? 0.f : mx_lo, rh = mx_hi == -INFINITY ? 0.f : mx_hi;
            const float c_lo = ex2_approx((m_lo - rl) * scale_log2), c_hi = ex2_approx((m_hi - rh) * scale_log2);
            m_lo = mx_lo; m_hi = mx_hi;
            l_lo *= c_lo; l_hi *= c_hi;
#pragma unroll
            for (int n = 0; n < 4; ++n) { o[n][0] *= c_lo; o[n][1] *= c_lo; o[n][2] *= c_hi; o[n][3] *= c_hi; }
            const float ml = rl * scale_log2, mh = rh * scale_log2;
            uint32_t pa[4][4];
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const float p0 = ex2_approx(fmaf(sc[n][0], scale_log2, -ml)), p1 = ex2_approx(fmaf(sc[n][1], scale_log2, -ml));
                const float p2 = ex2_approx(fmaf(sc[n][2], scale_log2, -mh)), p3 = ex2_approx(fmaf(sc[n][3], scale_log2, -mh));
                l_lo += p0 + p1;
                l_hi += p2 + p3;
                pa[n >> 1][(n & 1) * 2] = pk(p0, p1);
                pa[n >> 1][(n & 1) * 2 + 1] = pk(p2, p3);
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int nn = 0; nn < 2; ++nn) {
                    uint32_t vf[4];
                    ldsm4t(vf, Vs + (size_t)(kb + kk * 16 + v_row) * PITCH + (nn * 2 + v_chunk) * 8);
                    mma16816(o[nn * 2], pa[kk], vf[0], vf[1]);
                    mma16816(o[nn * 2 + 1], pa[kk], vf[2], vf[3]);
                }
        }
        float ll = l_lo, lh = l_hi;
        ll += __shfl_xor_sync(0xffffffffu, ll, 1);
        ll += __shfl_xor_sync(0xffffffffu, ll, 2);
        lh += __shfl_xor_sync(0xffffffffu, lh, 1);
        lh += __shfl_xor_sync(0xffffffffu, lh, 2);
        const float i_lo = ll > 0.f ? 1.f / ll : 0.f, i_hi = lh > 0.f ? 1.f / lh : 0.f;
        const int qr = q0 + g;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            if (qr < q_end)
                *reinterpret_cast<uint32_t*>(out + (row0 + qr) * ld_o + h * 32 + n * 8 + 2 * t) = pk(o[n][0] * i_lo, o[n][1] * i_lo);
            if (qr + 8 < q_end)
                *reinterpret_cast<uint32_t*>(out + (row0 + qr + 8) * ld_o + h * 32 + n * 8 + 2 * t) = pk(o[n][2] * i_hi, o[n][3] * i_hi);
        }
        if (t == 0) {
            float* lrow = lse2 + ((size_t)b * H + h) * Q;
            if (qr < q_end) lrow[qr] = (m_lo == -INFINITY ? 0.f : m_lo) * scale_log2 + log2f(fmaxf(ll, 1e-30f));
            if (qr + 8 < q_end) lrow[qr + 8] = (m_hi == -INFINITY ? 0.f : m_hi) * scale_log2 + log2f(fmaxf(lh, 1e-30f));
        }
    }
}

// ---------------------------------------------------------------------------------------------- D = rowsum(dO o O) per head
__global__ void __launch_bounds__(256)
sa_bwd_prep_kernel(const op16_t* __restrict__ o, int ld_o, const op16_t* __restrict__ dout, int ld_do, float* __restrict__ Dv, int B, int Q, int H) {
    pdl_launch_dependents();
    pdl_wait();
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);       // one warp per (image, query) row of H*32 channels
    if (row >= (long long)B * Q) return;
    const int lane = threadIdx.x & 31;
    for (int c0 = lane * 8; c0 < H * 32; c0 += 256) {
        const uint4 a = *reinterpret_cast<const uint4*>(o + (size_t)row * ld_o + c0);
        const uint4 d = *reinterpret_cast<const uint4*>(dout + (size_t)row * ld_do + c0);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, dw[4] = {d.x, d.y, d.z, d.w};
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) s += op16_lo_f32(aw[i]) * op16_lo_f32(dw[i]) + op16_hi_f32(aw[i]) * op16_hi_f32(dw[i]);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if ((lane & 3) == 0) {
            const int hh = c0 >> 5;
            const int bq = (int)(row / Q), i = (int)(row - (long long)bq * Q);
            Dv[((size_t)bq * H + hh) * Q + i] = s;
        }
    }
}

// ---------------------------------------------------------------------------------------------- backward, dQ pass
template <int MAXW>
__global__ void __launch_bounds__(MAXW * 32)
sa_bwd_dq_kernel(const op16_t* __restrict__ qk, int ld_qk, int k_off, const op16_t* __restrict__ v, int ld_v, const op16_t* __restrict__ dout,
                 int ld_do, const uint32_t* __restrict__ mask_bits, const float* __restrict__ lse2, const float* __restrict__ Dv,
                 op16_t* __restrict__ dqk, int ld_dqk, int Q, int q_per_cta, float scale_log2, float scale) {
    extern __shared__ __align__(16) unsigned char smem_[];
    const int KP = (Q + 63) / 64 * 64, MW = KP / 32;
    op16_t* Ks = reinterpret_cast<op16_t*>(smem_);
    op16_t* Vs = Ks + (size_t)KP * PITCH;
    const int b = blockIdx.z, h = blockIdx.y, H = gridDim.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x, nwarps = blockDim.x >> 5;
    const size_t row0 = (size_t)b * Q;
    pdl_launch_dependents();
    pdl_wait();
    stage_rows(Ks, qk + row0 * ld_qk + k_off + h * 32, ld_qk, Q, KP, tid, nthr);
    stage_rows(Vs, v + row0 * ld_v + h * 32, ld_v, Q, KP, tid, nthr);
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    const int g = lane >> 2, t = lane & 3;
    const int q_begin = blockIdx.x * q_per_cta, q_end = min(Q, q_begin + q_per_cta);
    const int k_row = lane & 7, k_chunk = lane >> 3;
    const int v_row = (lane & 7) + ((lane >> 3) & 1) * 8, v_chunk = lane >> 4;
    const float* lrow = lse2 + ((size_t)b * H + h) * Q;
    const float* drow = Dv + ((size_t)b * H + h) * Q;
    for (int q0 = q_begin + warp * 16; q0 < q_end; q0 += nwarps * 16) {
        const int r_lo = min(q0 + g, Q - 1), r_hi = min(q0 + g + 8, Q - 1);
        uint32_t qa[2][4], da[2][4];
        load_a_frags(qa, qk + (row0 + r_lo) * ld_qk + h * 32, qk + (row0 + r_hi) * ld_qk + h * 32, t);
        load_a_frags(da, dout + (row0 + r_lo) * ld_do + h * 32, dout + (row0 + r_hi) * ld_do + h * 32, t);
        const float ls_lo = lrow[r_lo], ls_hi = lrow[r_hi], D_lo = drow[r_lo], D_hi = drow[r_hi];
        float dq[4][4];
#pragma unroll
        for (int n = 0; n < 4; ++n) dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;
        for (int kb = 0; kb < KP; kb += 64) {
            uint32_t wlx = 0, wly = 0, whx = 0, why = 0;
            if (mask_bits) {
                const uint2 wl = *reinterpret_cast<const uint2*>(mask_bits + (size_t)r_lo * MW + (kb >> 5));
                const uint2 wh = *reinterpret_cast<const uint2*>(mask_bits + (size_t)r_hi * MW + (kb >> 5));
                wlx = wl.x; wly = wl.y; whx = wh.x; why = wh.y;
            }
            uint32_t pa[4][4];
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                float sc[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
                uint32_t kf[4], vf[4];
                ldsm4(kf, Ks + (size_t)(kb + n * 8 + k_row) * PITCH + k_chunk * 8);
                mma16816(sc, qa[0], kf[0], kf[1]);
                mma16816(sc, qa[1], kf[2], kf[3]);
                ldsm4(vf, Vs + (size_t)(kb + n * 8 + k_row) * PITCH + k_chunk * 8);
                mma16816(dp, da[0], vf[0], vf[1]);
                mma16816(dp, da[1], vf[2], vf[3]);
                float p0 = ex2_approx(fmaf(sc[0], scale_log2, -ls_lo)), p1 = ex2_approx(fmaf(sc[1], scale_log2, -ls_lo));
                float p2 = ex2_approx(fmaf(sc[2], scale_log2, -ls_hi)), p3 = ex2_approx(fmaf(sc[3], scale_log2, -ls_hi));
                const int key = kb + n * 8 + 2 * t;
                const uint32_t a = (n < 4 ? wlx : wly) >> ((n & 3) * 8 + 2 * t), c = (n < 4 ? whx : why) >> ((n & 3) * 8 + 2 * t);
                if (key >= Q || (a & 1u)) p0 = 0.f;
                if (key + 1 >= Q || (a & 2u)) p1 = 0.f;
                if (key >= Q || (c & 1u)) p2 = 0.f;
                if (key + 1 >= Q || (c & 2u)) p3 = 0.f;
                pa[n >> 1][(n & 1) * 2] = pk(p0 * (dp[0] - D_lo), p1 * (dp[1] - D_lo));
                pa[n >> 1][(n & 1) * 2 + 1] = pk(p2 * (dp[2] - D_hi), p3 * (dp[3] - D_hi));
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int nn = 0; nn < 2; ++nn) {
                    uint32_t kf[4];
                    ldsm4t(kf, Ks + (size_t)(kb + kk * 16 + v_row) * PITCH + (nn * 2 + v_chunk) * 8);
                    mma16816(dq[nn * 2], pa[kk], kf[0], kf[1]);
                    mma16816(dq[nn * 2 + 1], pa[kk], kf[2], kf[3]);
                }
        }
        const int qr = q0 + g;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            if (qr < q_end)
                *reinterpret_cast<uint32_t*>(dqk + (row0 + qr) * ld_dqk + h * 32 + n * 8 + 2 * t) = pk(dq[n][0] * scale, dq[n][1] * scale);
            if (qr + 8 < q_end)
                *reinterpret_cast<uint32_t*>(dqk + (row0 + qr + 8) * ld_dqk + h * 32 + n * 8 + 2 * t) = pk(dq[n][2] * scale, dq[n][3] * scale);
        }
    }
}

// ---------------------------------------------------------------------------------------------- backward, dK / dV pass
template <int MAXW>
__global__ void __launch_bounds__(MAXW * 32)
sa_bwd_dkv_kernel(const op16_t* __restrict__ qk, int ld_qk, int k_off, const op16_t* __restrict__ v, int ld_v, const op16_t* __restrict__ dout,
                  int ld_do, const uint32_t* __restrict__ maskT_bits, const float* __restrict__ lse2, const float* __restrict__ Dv,
                  op16_t* __restrict__ dqk, int ld_dqk, op16_t* __restrict__ dv, int ld_dv, int Q, int k_per_cta, float scale_log2, float scale) {
    extern __shared__ __align__(16) unsigned char smem_[];
    const int QP = (Q + 63) / 64 * 64, MW = QP / 32;
    op16_t* Qs = reinterpret_cast<op16_t*>(smem_);
    op16_t* Os = Qs + (size_t)QP * PITCH;                                  // dO
    float* lse_s = reinterpret_cast<float*>(Os + (size_t)QP * PITCH);
    float* D_s = lse_s + QP;
    const int b = blockIdx.z, h = blockIdx.y, H = gridDim.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x, nwarps = blockDim.x >> 5;
    const size_t row0 = (size_t)b * Q;
    pdl_launch_dependents();
    pdl_wait();
    stage_rows(Qs, qk + row0 * ld_qk + h * 32, ld_qk, Q, QP, tid, nthr);
    stage_rows(Os, dout + row0 * ld_do + h * 32, ld_do, Q, QP, tid, nthr);
    cp_async_commit();
    for (int i = tid; i < QP; i += nthr) {
        lse_s[i] = i < Q ? lse2[((size_t)b * H + h) * Q + i] : INFINITY;   // padded queries: P = 2^(-inf) = 0
        D_s[i] = i < Q ? Dv[((size_t)b * H + h) * Q + i] : 0.f;
    }
    cp_async_wait_all();
    __syncthreads();
    const int g = lane >> 2, t = lane & 3;
    const int k_begin = blockIdx.x * k_per_cta, k_end = min(Q, k_begin + k_per_cta);
    const int k_row = lane & 7, k_chunk = lane >> 3;
    const int v_row = (lane & 7) + ((lane >> 3) & 1) * 8, v_chunk = lane >> 4;
    for (int j0 = k_begin + warp * 16; j0 < k_end; j0 += nwarps * 16) {
        const int r_lo = min(j0 + g, Q - 1), r_hi = min(j0 + g + 8, Q - 1);
        uint32_t ka[2][4], va[2][4];
        load_a_frags(ka, qk + (row0 + r_lo) * ld_qk + k_off + h * 32, qk + (row0 + r_hi) * ld_qk + k_off + h * 32, t);
        load_a_frags(va, v + (row0 + r_lo) * ld_v + h * 32, v + (row0 + r_hi) * ld_v + h * 32, t);
        float dk[4][4], dvv[4][4];
#pragma unroll
        for (int n = 0; n < 4; ++n) { dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f; dvv[n][0] = dvv[n][1] = dvv[n][2] = dvv[n][3] = 0.f; }
        for (int qb = 0; qb < QP; qb += 32) {
            uint32_t wl = 0, wh = 0;
            if (maskT_bits) {
                wl = maskT_bits[(size_t)r_lo * MW + (qb >> 5)];
                wh = maskT_bits[(size_t)r_hi * MW + (qb >> 5)];
            }
            uint32_t pa[2][4], sa[2][4];
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                float st[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
                uint32_t qf[4], of[4];
                ldsm4(qf, Qs + (size_t)(qb + n * 8 + k_row) * PITCH + k_chunk * 8);
                mma16816(st, ka[0], qf[0], qf[1]);
                mma16816(st, ka[1], qf[2], qf[3]);
                ldsm4(of, Os + (size_t)(qb + n * 8 + k_row) * PITCH + k_chunk * 8);
                mma16816(dp, va[0], of[0], of[1]);
                mma16816(dp, va[1], of[2], of[3]);
                const int i0 = qb + n * 8 + 2 * t;
                const float2 ls = *reinterpret_cast<const float2*>(lse_s + i0), dd = *reinterpret_cast<const float2*>(D_s + i0);
                float p0 = ex2_approx(fmaf(st[0], scale_log2, -ls.x)), p1 = ex2_approx(fmaf(st[1], scale_log2, -ls.y));
                float p2 = ex2_approx(fmaf(st[2], scale_log2, -ls.x)), p3 = ex2_approx(fmaf(st[3], scale_log2, -ls.y));
                const uint32_t a = wl >> (n * 8 + 2 * t), c = wh >> (n * 8 + 2 * t);
                if (a & 1u) p0 = 0.f;
                if (a & 2u) p1 = 0.f;
                if (c & 1u) p2 = 0.f;
                if (c & 2u) p3 = 0.f;
                pa[n >> 1][(n & 1) * 2] = pk(p0, p1);
                pa[n >> 1][(n & 1) * 2 + 1] = pk(p2, p3);
                sa[n >> 1][(n & 1) * 2] = pk(p0 * (dp[0] - dd.x), p1 * (dp[1] - dd.y));
                sa[n >> 1][(n & 1) * 2 + 1] = pk(p2 * (dp[2] - dd.x), p3 * (dp[3] - dd.y));
            }
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                for (int nn = 0; nn < 2; ++nn) {
                    uint32_t f[4];
                    ldsm4t(f, Os + (size_t)(qb + kk * 16 + v_row) * PITCH + (nn * 2 + v_chunk) * 8);
                    mma16816(dvv[nn * 2], pa[kk], f[0], f[1]);
                    mma16816(dvv[nn * 2 + 1], pa[kk], f[2], f[3]);
                    ldsm4t(f, Qs + (size_t)(qb + kk * 16 + v_row) * PITCH + (nn * 2 + v_chunk) * 8);
                    mma16816(dk[nn * 2], sa[kk], f[0], f[1]);
                    mma16816(dk[nn * 2 + 1], sa[kk], f[2], f[3]);
                }
        }
        const int jr = j0 + g;
#pragma unroll
        for (int n = 0; n < 4; ++n) {
            if (jr < k_end) {
                *reinterpret_cast<uint32_t*>(dqk + (row0 + jr) * ld_dqk + k_off + h * 32 + n * 8 + 2 * t) = pk(dk[n][0] * scale, dk[n][1] * scale);
                *reinterpret_cast<uint32_t*>(dv + (row0 + jr) * ld_dv + h * 32 + n * 8 + 2 * t) = pk(dvv[n][0], dvv[n][1]);
            }
            if (jr + 8 < k_end) {
                *reinterpret_cast<uint32_t*>(dqk + (row0 + jr + 8) * ld_dqk + k_off + h * 32 + n * 8 + 2 * t) = pk(dk[n][2] * scale, dk[n][3] * scale);
                *reinterpret_cast<uint32_t*>(dv + (row0 + jr + 8) * ld_dv + h * 32 + n * 8 + 2 * t) = pk(dvv[n][2], dvv[n][3]);
            }
        }
    }
}

}  // namespace attn_train
}  // namespace dtlr

using namespace dtlr;
using namespace dtlr::attn_train;

static int check_common(const void* qk, int ld_qk, int k_off, const void* v, int ld_v, int B, int Q, int heads, int head_dim, int dtype) {
    DTLR_CHECK_ARG(head_dim == 32, "mha_train: head_dim must be 32, got %d", head_dim);
    DTLR_CHECK_ARG(dtype == DTLR_OP16, "mha_train: 16-bit operands of the library's flavour only");
    DTLR_CHECK_ARG(B >= 0 && Q >= 0 && heads > 0, "mha_train: bad sizes");
    DTLR_CHECK_ARG((ld_qk % 8) == 0 && (ld_v % 8) == 0 && (k_off % 8) == 0 && ((((uintptr_t)qk | (uintptr_t)v)) & 15) == 0,
                   "mha_train: rows must be 16-byte aligned");
    return DTLR_OK;
}

// per (image, head): T = ceil(Q / 16) warp tiles; cut into `splits` CTAs of W warps so that all CTAs fit ~2 waves of the SMs
static void partition(int Q, int heads, int B, int W, int* splits, int* per_cta) {
    const int T = (Q + 15) / 16;
    int best_s = 1;
    long long best = -1;
    for (int s = 1; s <= 8; ++s) {
        const int tpc = (T + s - 1) / s;
        const long long ctas = (long long)s * heads * B;
        const long long cost = ((ctas + sm_count() - 1) / sm_count()) * (((tpc + W - 1) / W) * (long long)W + 3);
        if (best < 0 || cost < best) { best = cost; best_s = s; }
    }
    *splits = best_s;
    *per_cta = ((T + best_s - 1) / best_s) * 16;
}

extern "C" int dtlr_mha_train_forward(const void* qk, int ld_qk, int k_off, const void* v, int ld_v, const uint32_t* mask_bits, void* out,
                                      int ld_o, float* lse2, int B, int Q, int heads, int head_dim, int dtype, void* stream) {
    int rc = check_common(qk, ld_qk, k_off, v, ld_v, B, Q, heads, head_dim, dtype);
    if (rc) return rc;
    if (B == 0 || Q == 0) return DTLR_OK;
    DTLR_CHECK_ARG(out && lse2 && (ld_o % 2) == 0, "mha_train_forward: bad output");
    const int KP = (Q + 63) / 64 * 64;
    const size_t smem = (size_t)KP * PITCH * 2 * 2;
    DTLR_CHECK_ARG(smem <= (size_t)max_smem_optin(), "mha_train_forward: Q = %d does not fit shared memory", Q);
    constexpr int W = 16;
    int splits, qpc;
    partition(Q, heads, B, W, &splits, &qpc);
    auto kern = sa_fwd_kernel<W>;
    DTLR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((Q + qpc - 1) / qpc, heads, B);
    DTLR_CHECK_CUDA(launch_pdl(kern, grid, dim3(W * 32), smem, (cudaStream_t)stream, (const op16_t*)qk, ld_qk, k_off, (const op16_t*)v, ld_v,
                               mask_bits, (op16_t*)out, ld_o, lse2, Q, qpc, 0.17677669529663687f * 1.4426950408889634f));
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_mha_train_backward(const void* qk, int ld_qk, int k_off, const void* v, int ld_v, const void* out, int ld_o,
                                       const void* dout, int ld_do, const uint32_t* mask_bits, const uint32_t* maskT_bits, const float* lse2,
                                       float* D_scratch, void* dqk, int ld_dqk, void* dv, int ld_dv, int B, int Q, int heads, int head_dim,
                                       int dtype, void* stream) {
    int rc = check_common(qk, ld_qk, k_off, v, ld_v, B, Q, heads, head_dim, dtype);
    if (rc) return rc;
    if (B == 0 || Q == 0) return DTLR_OK;
    DTLR_CHECK_ARG(out && dout && lse2 && D_scratch && dqk && dv, "mha_train_backward: null pointer");
    DTLR_CHECK_ARG((ld_o % 8) == 0 && (ld_do % 8) == 0 && (ld_dqk % 2) == 0 && (ld_dv % 2) == 0 && ((((uintptr_t)out | (uintptr_t)dout)) & 15) == 0,
                   "mha_train_backward: rows must be 16-byte aligned");
    DTLR_CHECK_ARG((mask_bits == nullptr) == (maskT_bits == nullptr), "mha_train_backward: give both mask orientations or neither");
    cudaStream_t st = (cudaStream_t)stream;
    const int KP = (Q + 63) / 64 * 64;
    const float scale = 0.17677669529663687f, sl2 = scale * 1.4426950408889634f;
    DTLR_CHECK_CUDA(launch_pdl(sa_bwd_prep_kernel, dim3((unsigned)(((long long)B * Q + 7) / 8)), dim3(256), 0, st, (const op16_t*)out, ld_o,
                               (const op16_t*)dout, ld_do, D_scratch, B, Q, heads));
    {
        constexpr int W = 16;
        const size_t smem = (size_t)KP * PITCH * 2 * 2;
        DTLR_CHECK_ARG(smem <= (size_t)max_smem_optin(), "mha_train_backward: Q = %d does not fit shared memory", Q);
        int splits, qpc;
        partition(Q, heads, B, W, &splits, &qpc);
        auto kern = sa_bwd_dq_kernel<W>;
        DTLR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((Q + qpc - 1) / qpc, heads, B);
        DTLR_CHECK_CUDA(launch_pdl(kern, grid, dim3(W * 32), smem, st, (const op16_t*)qk, ld_qk, k_off, (const op16_t*)v, ld_v,
                                   (const op16_t*)dout, ld_do, mask_bits, lse2, (const float*)D_scratch, (op16_t*)dqk, ld_dqk, Q, qpc, sl2, scale));
    }
    {
        constexpr int W = 16;
        const size_t smem = (size_t)KP * PITCH * 2 * 2 + (size_t)KP * 8;
        DTLR_CHECK_ARG(smem <= (size_t)max_smem_optin(), "mha_train_backward: Q = %d does not fit shared memory", Q);
        int splits, kpc;
        partition(Q, heads, B, W, &splits, &kpc);
        auto kern = sa_bwd_dkv_kernel<W>;
        DTLR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((Q + kpc - 1) / kpc, heads, B);
        DTLR_CHECK_CUDA(launch_pdl(kern, grid, dim3(W * 32), smem, st, (const op16_t*)qk, ld_qk, k_off, (const op16_t*)v, ld_v,
                                   (const op16_t*)dout, ld_do, maskT_bits, lse2, (const float*)D_scratch, (op16_t*)dqk, ld_dqk, (op16_t*)dv, ld_dv,
                                   Q, kpc, sl2, scale));
    }
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}
