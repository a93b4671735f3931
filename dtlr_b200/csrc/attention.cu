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
template <> __device__ __forceinline__ float ldf_<op16_t>(const op16_t* p) { return op16_to_f32(*p); }
template <typename T> __device__ __forceinline__ void stf_(T* p, float v);
template <> __device__ __forceinline__ void stf_<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf_<op16_t>(op16_t* p, float v) { *p = f32_to_op16(v); }

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
constexpr int FA_WARPS = 16;       // default warps per CTA (dtlr_attn_config overrides: 8..20)
constexpr int FA_WARPS_MAX = 20;   // 640 threads x 86 registers fit the register file (limit 102 per thread)
constexpr int FA_MT = 1;        // 16-query m-tiles per warp (MT = 2 with 8 warps measured slower: 305 vs 287 us/layer)
constexpr int FA_PITCH = 40;   // bf16 elements per smem row (32 + 8 padding)

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." DTLR_OP16_PTX "." DTLR_OP16_PTX ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
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
    op16x2_t t = op16_pack2(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}

// MT = 16-query m-tiles per warp: every K / V fragment fetched from shared memory feeds MT MMAs (MT = 2 halves the
// ldmatrix traffic per FLOP, which is what bounds the MT = 1 version).
// PIPE (experiment, opt-in with dtlr_debug_flags(16384)): the S = Q K^T MMAs of key block kb+1 are issued before the softmax of block
// kb (two ping-pong score arrays, the loop body instantiated twice), so that the tensor pipe could work under the MUFU / FMNMX /
// shuffle chain of the softmax.  Measured SLOWER on B200 (B=64, Q=900: 281.6 us vs 255.4 us un-pipelined, same box): the second
// score array takes the kernel from 86 to 128 registers (the cap at 512 threads) with 60 bytes of spills, and ptxas already
// interleaves the P.V MMAs with the exponentials of the same block -- the four warps per scheduler cover the rest.
// MAXW: the launch bound in warps -- 16 (86 registers, no spills) for CTAs of up to 16 warps, 20 (96 registers + 40 bytes of spills under
// the 102-register cap of 640 threads) only when the partition really asks for more than 16 warps
template <int MT, bool PIPE, int MAXW>
__global__ void __launch_bounds__(MAXW * 32)
mha_flash_bf16_kernel(const op16_t* __restrict__ qk, int ld_qk, int k_off, const op16_t* __restrict__ v, int ld_v,
                      op16_t* __restrict__ out, int ld_o, int Q, int q_per_cta, float scale_log2) {
    extern __shared__ __align__(16) unsigned char fa_smem[];
    const int KP = (Q + 63) / 64 * 64;                        // keys padded to the 64-key sweep
    op16_t* Ks = reinterpret_cast<op16_t*>(fa_smem);
    op16_t* Vs = Ks + (size_t)KP * FA_PITCH;
    const int b = blockIdx.z, h = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthr = blockDim.x, nwarps = blockDim.x >> 5;
    const size_t row0 = (size_t)b * Q;
    pdl_launch_dependents();
    pdl_wait();

    // ---- stage K and V of this (image, head): 4 x 16-byte chunks per row, zero rows for the key padding
    for (int i = tid; i < KP * 4; i += nthr) {
        const int r = i >> 2, c = i & 3;
        if (r < Q) {
            cp_async16(Ks + (size_t)r * FA_PITCH + c * 8, qk + (row0 + r) * ld_qk + k_off + h * 32 + c * 8);
            cp_async16(Vs + (size_t)r * FA_PITCH + c * 8, v + (row0 + r) * ld_v + h * 32 + c * 8);
        } else {
            *reinterpret_cast<uint4*>(Ks + (size_t)r * FA_PITCH + c * 8) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(Vs + (size_t)r * FA_PITCH + c * 8) = make_uint4(0, 0, 0, 0);
        }
    }
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();

    const int g = lane >> 2, t = lane & 3;
    const int q_begin = blockIdx.x * q_per_cta;
    const int q_end = min(Q, q_begin + q_per_cta);
    // ldmatrix lane -> row/column offsets.  K (non-trans): matrices = dh chunks 0..3 of 8 keys; lane supplies
    // key (lane & 7), dh chunk (lane >> 3).  V (trans): matrices (keys lo, n), (keys hi, n), (keys lo, n+1), (keys hi, n+1).
    const int k_row = lane & 7, k_chunk = lane >> 3;
    const int v_row = (lane & 7) + ((lane >> 3) & 1) * 8, v_chunk = lane >> 4;

    for (int q0 = q_begin + warp * (16 * MT); q0 < q_end; q0 += nwarps * 16 * MT) {
        // ---- Q fragments (A operand, 2 k-steps over dh = 32), straight from global memory
        uint32_t qa[MT][2][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const int r_lo = min(q0 + mt * 16 + g, Q - 1), r_hi = min(q0 + mt * 16 + g + 8, Q - 1);
            const op16_t* qlo = qk + (row0 + r_lo) * ld_qk + h * 32;
            const op16_t* qhi = qk + (row0 + r_hi) * ld_qk + h * 32;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                qa[mt][ks][0] = *reinterpret_cast<const uint32_t*>(qlo + ks * 16 + 2 * t);
                qa[mt][ks][1] = *reinterpret_cast<const uint32_t*>(qhi + ks * 16 + 2 * t);
                qa[mt][ks][2] = *reinterpret_cast<const uint32_t*>(qlo + ks * 16 + 8 + 2 * t);
                qa[mt][ks][3] = *reinterpret_cast<const uint32_t*>(qhi + ks * 16 + 8 + 2 * t);
            }
        }
        float o[MT][4][4];
        float m_lo[MT], m_hi[MT], l_lo[MT], l_hi[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            m_lo[mt] = m_hi[mt] = -INFINITY;
            l_lo[mt] = l_hi[mt] = 0.f;
#pragma unroll
            for (int n = 0; n < 4; ++n)
#pragma unroll
                for (int k = 0; k < 4; ++k) o[mt][n][k] = 0.f;
        }

        // ---- S = Q K^T for the 64 keys from kb: 8 n-blocks of 8 keys, each K fragment reused by the MT m-tiles
        // (tried in round 2: skipping the all-padding groups of 8 keys of the last block -- Q = 900: 4 real keys + 60 zeros -- with
        //  uniform predicates here and in the softmax / P.V: 86 -> 95 registers and 240 -> 301 us per layer on B200; reverted)
        auto qk_block = [&](float (&sc)[MT][8][4], const int kb) {
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                uint32_t kf[4];
                ldmatrix_x4(kf, Ks + (size_t)(kb + n * 8 + k_row) * FA_PITCH + k_chunk * 8);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    sc[mt][n][0] = sc[mt][n][1] = sc[mt][n][2] = sc[mt][n][3] = 0.f;
                    mma_bf16_16816(sc[mt][n], qa[mt][0], kf[0], kf[1]);
                    mma_bf16_16816(sc[mt][n], qa[mt][1], kf[2], kf[3]);
                }
            }
        };
        // ---- masking of the key padding, online softmax, O += P V for the 64 keys from kb (scores in sc)
        auto softmax_pv = [&](float (&sc)[MT][8][4], const int kb) {
            if (kb + 64 > Q) {    // key padding -> -inf
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const int key = kb + n * 8 + 2 * t;
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        if (key >= Q) { sc[mt][n][0] = -INFINITY; sc[mt][n][2] = -INFINITY; }
                        if (key + 1 >= Q) { sc[mt][n][1] = -INFINITY; sc[mt][n][3] = -INFINITY; }
                    }
                }
            }
            // online softmax (rows g and g+8 of each m-tile); a quad of lanes shares a row
            uint32_t pa[MT][4][4];    // P as A fragments: 4 k-steps of 16 keys
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                float mx_lo = m_lo[mt], mx_hi = m_hi[mt];
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    mx_lo = fmaxf(mx_lo, fmaxf(sc[mt][n][0], sc[mt][n][1]));
                    mx_hi = fmaxf(mx_hi, fmaxf(sc[mt][n][2], sc[mt][n][3]));
                }
                mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
                mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
                mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
                mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
                const float c_lo = ex2_approx((m_lo[mt] - mx_lo) * scale_log2), c_hi = ex2_approx((m_hi[mt] - mx_hi) * scale_log2);
                m_lo[mt] = mx_lo; m_hi[mt] = mx_hi;
                l_lo[mt] *= c_lo; l_hi[mt] *= c_hi;
#pragma unroll
                for (int n = 0; n < 4; ++n) { o[mt][n][0] *= c_lo; o[mt][n][1] *= c_lo; o[mt][n][2] *= c_hi; o[mt][n][3] *= c_hi; }
                const float ml = mx_lo * scale_log2, mh = mx_hi * scale_log2;
#pragma unroll
                for (int n = 0; n < 8; ++n) {
                    const float p0 = ex2_approx(fmaf(sc[mt][n][0], scale_log2, -ml)), p1 = ex2_approx(fmaf(sc[mt][n][1], scale_log2, -ml));
                    const float p2 = ex2_approx(fmaf(sc[mt][n][2], scale_log2, -mh)), p3 = ex2_approx(fmaf(sc[mt][n][3], scale_log2, -mh));
                    l_lo[mt] += p0 + p1;
                    l_hi[mt] += p2 + p3;
                    pa[mt][n >> 1][(n & 1) * 2] = pack_bf16(p0, p1);
                    pa[mt][n >> 1][(n & 1) * 2 + 1] = pack_bf16(p2, p3);
                }
            }
            // O += P V, each V fragment reused by the MT m-tiles
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                for (int nn = 0; nn < 2; ++nn) {
                    uint32_t vf[4];
                    ldmatrix_x4_trans(vf, Vs + (size_t)(kb + kk * 16 + v_row) * FA_PITCH + (nn * 2 + v_chunk) * 8);
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) {
                        mma_bf16_16816(o[mt][nn * 2], pa[mt][kk], vf[0], vf[1]);
                        mma_bf16_16816(o[mt][nn * 2 + 1], pa[mt][kk], vf[2], vf[3]);
                    }
                }
            }
        };
        if (PIPE) {
            float sa[MT][8][4], sb[MT][8][4];
            qk_block(sa, 0);
            for (int kb = 0;;) {
                if (kb + 64 < KP) qk_block(sb, kb + 64);
                softmax_pv(sa, kb);
                kb += 64;
                if (kb >= KP) break;
                if (kb + 64 < KP) qk_block(sa, kb + 64);
                softmax_pv(sb, kb);
                kb += 64;
                if (kb >= KP) break;
            }
        } else {
            for (int kb = 0; kb < KP; kb += 64) {
                float sc[MT][8][4];
                qk_block(sc, kb);
                softmax_pv(sc, kb);
            }
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            float ll = l_lo[mt], lh = l_hi[mt];
            ll += __shfl_xor_sync(0xffffffffu, ll, 1);
            ll += __shfl_xor_sync(0xffffffffu, ll, 2);
            lh += __shfl_xor_sync(0xffffffffu, lh, 1);
            lh += __shfl_xor_sync(0xffffffffu, lh, 2);
            const float i_lo = 1.f / ll, i_hi = 1.f / lh;
            const int qr = q0 + mt * 16 + g;
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                if (qr < q_end)
                    *reinterpret_cast<uint32_t*>(out + (row0 + qr) * ld_o + h * 32 + n * 8 + 2 * t) = pack_bf16(o[mt][n][0] * i_lo, o[mt][n][1] * i_lo);
                if (qr + 8 < q_end)
                    *reinterpret_cast<uint32_t*>(out + (row0 + qr + 8) * ld_o + h * 32 + n * 8 + 2 * t) = pack_bf16(o[mt][n][2] * i_hi, o[mt][n][3] * i_hi);
            }
        }
    }
}

}  // namespace dtlr

using namespace dtlr;

static int g_attn_warps = 0, g_attn_splits = 0;
// tuning hook: warps per CTA (1..20) and query splits per (image, head) of the flash kernel; 0, 0 = automatic
extern "C" int dtlr_attn_config(int warps, int splits) { g_attn_warps = warps; g_attn_splits = splits; return DTLR_OK; }

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
    else if (dtype == DTLR_OP16 && !attn_mask && (ld_qk % 8) == 0 && (ld_v % 8) == 0 && (k_off % 8) == 0 && (ld_o % 2) == 0 &&
             (size_t)((Q + 63) / 64 * 64) * FA_PITCH * 2 * 2 <= (size_t)max_smem_optin()) {
        const int KP = (Q + 63) / 64 * 64;
        const size_t smem = (size_t)KP * FA_PITCH * 2 * 2;
        // Work partition.  A (image, head) has T = ceil(Q / 16) query tiles of one warp each; it is cut into `splits` CTAs of W warps
        // (each CTA stages K and V once).  Two quantisations cost time: tiles per CTA vs W (Q = 900: 57 tiles; 4 x 16 warps = 64
        // slots = 89 %) and CTAs vs SMs (one CTA per SM: 144 KB of K / V).  Measured on B200 (B = 64, Q = 900, tools/bench_attn.py, us):
        // (W, splits) = (16,4) 255.7 [round 1], (16,2) 233.3, (15,2) 240.3, (19,1) 240.4, (20,1) 240.2, (19,3) 248.8, (12,5) 279.2: the
        // kernel is issue-bound, not latency-bound (a round costs in proportion to the warps per scheduler, so W is kept a multiple of
        // 4), and every extra split re-stages K and V (~3 warp-rounds).  Auto: minimise ceil(CTAs / SMs) * (ceil(tiles per CTA / W) * W
        // + 3) over W in {12, 16, 20}; dtlr_attn_config() overrides (A/B).
        const int T = (Q + 16 * FA_MT - 1) / (16 * FA_MT);
        int W = g_attn_warps, splits = g_attn_splits;
        if (W <= 0 || splits <= 0) {
            long long best = -1;
            for (int s_ = 1; s_ <= 8; ++s_)
                for (int w_ = 12; w_ <= FA_WARPS_MAX; w_ += 4) {
                    const int tpc = (T + s_ - 1) / s_;
                    const long long ctas = (long long)s_ * heads * B;
                    const long long cost = ((ctas + sm_count() - 1) / sm_count()) * (((tpc + w_ - 1) / w_) * (long long)w_ + 3);   // + 3: K / V staging
                    if (best < 0 || cost < best) { best = cost; W = w_; splits = s_; }
                }
        }
        W = W < 1 ? 1 : (W > FA_WARPS_MAX ? FA_WARPS_MAX : W);
        int q_per_cta = ((T + splits - 1) / splits) * 16 * FA_MT;
        dim3 fgrid((Q + q_per_cta - 1) / q_per_cta, heads, B);
        auto kern = (g_debug_flags & 16384) ? mha_flash_bf16_kernel<FA_MT, true, FA_WARPS_MAX>              // flag 16384: QK-pipelined variant (A/B)
                                            : (W <= 16 ? mha_flash_bf16_kernel<FA_MT, false, 16> : mha_flash_bf16_kernel<FA_MT, false, FA_WARPS_MAX>);
        DTLR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DTLR_CHECK_CUDA(launch_pdl(kern, fgrid, dim3(W * 32), smem, st, (const op16_t*)qk, ld_qk, k_off,
                                   (const op16_t*)v, ld_v, (op16_t*)out, ld_o, Q, q_per_cta, scale * 1.4426950408889634f));
    } else if (dtype == DTLR_OP16)
        mha_simt_kernel<op16_t><<<grid, ATT_QT, 0, st>>>((const op16_t*)qk, ld_qk, k_off, (const op16_t*)v, ld_v, attn_mask, (op16_t*)out, ld_o, Q, scale);
    else { set_error("mha: unsupported dtype"); return DTLR_ERR_INVALID; }
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}
