// dtlr_b200 -- decoder self-attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), bf16, no mask.
//
// nn.MultiheadAttention(256, 8) core of reference deformable_transformer.py:847, 903-905 at Q = 900 queries, head dim 32.
// One CTA = (image, head, up to 4 tiles of 128 queries).  K (Q x 32, 64-byte-swizzled rows) and V^T (32 x keys, 128-byte
// swizzle; produced by a small transpose pre-pass) of the head are TMA-loaded once and stay in shared memory.
// Per 128-query tile the scores never leave the SM:
//   pass 1   S_c = Q K_c^T for the eight 128-key chunks (tcgen05.mma 128x128x16, fp32 in TMEM, double buffered); the softmax
//            warps read each chunk with tcgen05.ld and keep the running row maximum;
//   pass 2   S_c is recomputed (QK^T is 2 MMAs per chunk -- cheaper than rescaling an accumulator), P_c = exp2((S_c - max) *
//            scale) is written as bf16 straight into the 128B-swizzled K-major layout an A operand needs, and
//            O += P_c V_c^T runs as tcgen05.mma 128x32x16 into a third TMEM region;
//   epilogue O / rowsum -> bf16 -> global.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer + TMEM owner, warps 2-17 softmax (four warps per TMEM lane quarter,
// each owning 32 of the 128 key columns of a chunk).  All hand-offs are mbarriers; tcgen05.commit signals MMA completion.
#include "tc_common.cuh"

namespace dtlr {

constexpr int AT_QT = 128;        // queries per tile
constexpr int AT_KC = 128;        // keys per chunk
constexpr int AT_KPAD = 1024;     // padded key count (8 chunks)
constexpr int AT_TILES = 4;       // query tiles per CTA
constexpr int AT_SMEM_K = AT_KPAD * 64;            // 64 KB
constexpr int AT_SMEM_V = 16 * 4096;               // 16 k-blocks of 32 rows x 128 B
constexpr int AT_SMEM_Q = AT_QT * 64;              // 8 KB
constexpr int AT_SMEM_P = 2 * 16384;               // one P chunk: 2 k-blocks of 128 rows x 128 B
constexpr int AT_XCH_FLOATS = 2 * 2 * 128 * 4;     // {max, sum} x tile parity x 128 rows x 4 column quarters
constexpr int AT_SMEM_TOTAL = AT_SMEM_K + AT_SMEM_V + AT_SMEM_Q + 2 * AT_SMEM_P + AT_XCH_FLOATS * 4 + 1024 + 512;

// V^T pre-pass: vt[(b*H + h)*32 + d][key] = v[b*Q + key][h*32 + d], zero for key >= Q
__global__ void vt_transpose_kernel(const __nv_bfloat16* __restrict__ v, int ld_v, __nv_bfloat16* __restrict__ vt, int Q, int H) {
    __shared__ __nv_bfloat16 tile[32][34];
    const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8 threads
    for (int r = ty; r < 32; r += 8) {
        const int key = k0 + r;
        tile[r][tx] = key < Q ? v[((size_t)b * Q + key) * ld_v + h * 32 + tx] : __float2bfloat16(0.f);
    }
    __syncthreads();
    for (int d = ty; d < 32; d += 8)
        vt[((size_t)(b * H + h) * 32 + d) * AT_KPAD + k0 + tx] = tile[tx][d];
}

__global__ void __launch_bounds__(576, 1)
mha_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ out, int ld_o, int Q, int H, int k_off,
                   float scale_log2) {
    extern __shared__ unsigned char at_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)at_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* sK = smem;
    unsigned char* sV = sK + AT_SMEM_K;
    unsigned char* sQ = sV + AT_SMEM_V;
    unsigned char* sP = sQ + AT_SMEM_Q;                     // [2][2 k-blocks][128 rows][128 B]
    float* xch = reinterpret_cast<float*>(sP + 2 * AT_SMEM_P);   // row max [2][128][2] then row sum [2][128][2]: exchange between column halves
    uint64_t* bars = reinterpret_cast<uint64_t*>(xch + AT_XCH_FLOATS);
    uint64_t* kv_full = bars;            // 1
    uint64_t* q_full = bars + 1;         // 1
    uint64_t* q_empty = bars + 2;        // 1
    uint64_t* s_full = bars + 3;         // 2
    uint64_t* s_empty = bars + 5;        // 2
    uint64_t* p_full = bars + 7;         // 2
    uint64_t* p_empty = bars + 9;        // 2
    uint64_t* o_full = bars + 11;        // 1
    uint64_t* o_empty = bars + 12;       // 1
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 13);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, h = blockIdx.y;
    const int tile0 = blockIdx.x * AT_TILES;
    const int n_qtiles = (Q + AT_QT - 1) / AT_QT;
    const int my_tiles = min(AT_TILES, n_qtiles - tile0);
    const int row_base = b * Q;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
        mbar_init(kv_full, 1); mbar_init(q_full, 1); mbar_init(q_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 16);
            mbar_init(&p_full[i], 16); mbar_init(&p_empty[i], 1);
        }
        mbar_init(o_full, 1); mbar_init(o_empty, 16);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_ptr);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tmem_S = tmem_base;          // 2 x 128 columns
    const uint32_t tmem_O = tmem_base + 256;    // 32 columns

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            mbar_expect_tx(kv_full, AT_SMEM_K + AT_SMEM_V);
            for (int j = 0; j < 4; ++j)          // K: 4 boxes of 256 keys x 32 dims (rows past this image are masked later)
                tma_load_2d(sK + j * 256 * 64, &tmK, kv_full, k_off + h * 32, row_base + j * 256);
            for (int j = 0; j < 16; ++j)         // V^T: 16 k-blocks of 64 keys x 32 dims
                tma_load_2d(sV + j * 4096, &tmV, kv_full, j * 64, (b * H + h) * 32);
            for (int t = 0; t < my_tiles; ++t) {
                mbar_wait(q_empty, (t & 1) ^ 1);
                mbar_expect_tx(q_full, AT_SMEM_Q);
                tma_load_2d(sQ, &tmQ, q_full, h * 32, row_base + (tile0 + t) * AT_QT);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t IDESC_S = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(AT_KC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t IDESC_O = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        mbar_wait(kv_full, 0);
        uint32_t si = 0;     // S-buffer use counter (buffer si & 1, phase (si >> 1) & 1)
        uint32_t pi = 0;     // P-buffer use counter
        const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
        auto issue_S = [&](int c) {
            const uint32_t sb = si & 1;
            mbar_wait(&s_empty[sb], ((si >> 1) & 1) ^ 1);
            tcgen05_fence_after();
            if (elect_one()) {
                const uint64_t dq = make_sw64_kmajor_desc(aQ);
                const uint64_t dk = make_sw64_kmajor_desc(aK + c * AT_KC * 64);
                umma_bf16(tmem_S + sb * AT_KC, dq, dk, IDESC_S, 0);
                umma_bf16(tmem_S + sb * AT_KC, dq + 2, dk + 2, IDESC_S, 1);
                umma_commit(&s_full[sb]);
            }
            __syncwarp();
            ++si;
        };
        auto issue_PV = [&](int c, bool first) {
            const uint32_t pb = pi & 1;
            mbar_wait(&p_full[pb], (pi >> 1) & 1);
            tcgen05_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    const uint64_t dp = make_sw128_kmajor_desc(aP + pb * AT_SMEM_P + kb * 16384);
                    const uint64_t dv = make_sw128_kmajor_desc(aV + (c * 2 + kb) * 4096);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_O, dp + (uint64_t)(2 * k), dv + (uint64_t)(2 * k), IDESC_O, (first && kb == 0 && k == 0) ? 0u : 1u);
                }
                umma_commit(&p_empty[pb]);
            }
            __syncwarp();
            ++pi;
        };
        for (int t = 0; t < my_tiles; ++t) {
            mbar_wait(q_full, t & 1);
            tcgen05_fence_after();
            for (int c = 0; c < 8; ++c) issue_S(c);                     // pass 1: row maxima
            mbar_wait(o_empty, (t & 1) ^ 1);                            // previous tile's O has been read out
            tcgen05_fence_after();
            for (int c = 0; c < 8; ++c) {                               // pass 2: probabilities and P V
                issue_S(c);
                if (c == 7 && elect_one()) umma_commit(q_empty);        // Q tile no longer needed once S_7 retires
                __syncwarp();
                if (c >= 1) issue_PV(c - 1, c == 1);
            }
            issue_PV(7, false);
            if (elect_one()) umma_commit(o_full);
            __syncwarp();
        }
    } else {
        // ===== softmax warps =====
        const int qd = warp & 3;
        const int part = (warp - 2) >> 2;                               // which 32 of the 128 key columns of a chunk (0..3)
        const int row = qd * 32 + lane;                                 // row of the 128-query tile == TMEM lane
        const uint32_t lane_sel = (uint32_t)(qd * 32) << 16;
        uint32_t si = 0, pi = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const int q = (tile0 + t) * AT_QT + row;
            // ---- pass 1: running maximum over this warp's 64 of every 128 key columns
            float mx = -INFINITY;
            for (int c = 0; c < 8; ++c, ++si) {
                const uint32_t sb = si & 1;
                mbar_wait(&s_full[sb], (si >> 1) & 1);
                tcgen05_fence_after();
                const bool full_chunk = (c + 1) * AT_KC <= Q;       // only the boundary chunk pays for per-key masking
                {
                    uint32_t acc[32];
                    tmem_ld32(tmem_S + sb * AT_KC + lane_sel + part * 32, acc);
                    if (full_chunk) {
#pragma unroll
                        for (int j = 0; j < 32; j += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(acc[j]), __uint_as_float(acc[j + 1])));
                    } else {
                        const int key0 = c * AT_KC + part * 32;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (key0 + j < Q) mx = fmaxf(mx, __uint_as_float(acc[j]));
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_empty[sb]);
            }
            float* xm = xch + (t & 1) * 512;
            xm[row * 4 + part] = mx;
            asm volatile("bar.sync %0, 128;" ::"r"(1 + qd) : "memory");            // the 4 warps of this lane quarter
            mx = fmaxf(fmaxf(xm[row * 4], xm[row * 4 + 1]), fmaxf(xm[row * 4 + 2], xm[row * 4 + 3]));
            const float mxs = mx * scale_log2;
            // ---- pass 2: P = exp2(S*scale - max*scale) -> bf16 -> swizzled A-operand tile; row sums
            float sum = 0.f;
            for (int c = 0; c < 8; ++c, ++si, ++pi) {
                const uint32_t sb = si & 1, pb = pi & 1;
                mbar_wait(&s_full[sb], (si >> 1) & 1);
                mbar_wait(&p_empty[pb], ((pi >> 1) & 1) ^ 1);
                tcgen05_fence_after();
                unsigned char* prow = sP + pb * AT_SMEM_P + (part >> 1) * 16384 + row * 128;
                {
                    uint32_t acc[32];
                    tmem_ld32(tmem_S + sb * AT_KC + lane_sel + part * 32, acc);
                    const int key0 = c * AT_KC + part * 32;
                    float p[32];
                    if ((c + 1) * AT_KC <= Q) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            p[j] = ex2_approx(fmaf(__uint_as_float(acc[j]), scale_log2, -mxs));
                            sum += p[j];
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            p[j] = (key0 + j < Q) ? ex2_approx(fmaf(__uint_as_float(acc[j]), scale_log2, -mxs)) : 0.f;
                            sum += p[j];
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        uint4 o;
                        __nv_bfloat162 a0 = __floats2bfloat162_rn(p[j], p[j + 1]), a1 = __floats2bfloat162_rn(p[j + 2], p[j + 3]);
                        __nv_bfloat162 a2 = __floats2bfloat162_rn(p[j + 4], p[j + 5]), a3 = __floats2bfloat162_rn(p[j + 6], p[j + 7]);
                        o.x = *reinterpret_cast<uint32_t*>(&a0); o.y = *reinterpret_cast<uint32_t*>(&a1);
                        o.z = *reinterpret_cast<uint32_t*>(&a2); o.w = *reinterpret_cast<uint32_t*>(&a3);
                        const int chunk = (part & 1) * 4 + j / 8;                // 16-byte chunk inside the 128-byte row
                        *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) * 16)) = o;
                    }
                }
                tcgen05_fence_before();
                fence_proxy_async();                                          // generic-proxy smem writes -> visible to the MMA
                __syncwarp();
                if (lane == 0) { mbar_arrive(&s_empty[sb]); mbar_arrive(&p_full[pb]); }
            }
            float* xs = xch + 1024 + (t & 1) * 512;
            xs[row * 4 + part] = sum;
            asm volatile("bar.sync %0, 128;" ::"r"(1 + qd) : "memory");
            const float inv = 1.f / ((xs[row * 4] + xs[row * 4 + 1]) + (xs[row * 4 + 2] + xs[row * 4 + 3]));
            // ---- epilogue: O (128 x 32 fp32 in TMEM) / row sum -> bf16; this warp writes 8 of the 32 head channels
            mbar_wait(o_full, t & 1);
            tcgen05_fence_after();
            uint32_t oacc[8];
            tmem_ld8(tmem_O + lane_sel + part * 8, oacc);
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_empty);
            if (q < Q) {
                __nv_bfloat16* op = out + (size_t)(row_base + q) * ld_o + h * 32 + part * 8;
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(oacc[2 * j]) * inv, __uint_as_float(oacc[2 * j + 1]) * inv);
                    w[j] = *reinterpret_cast<uint32_t*>(&t0);
                }
                *reinterpret_cast<uint4*>(op) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace dtlr

using namespace dtlr;

// returns DTLR_ERR_UNSUPPORTED when the shape does not fit this kernel (the caller then uses the mma.sync flash kernel)
extern "C" int dtlr_mha_tcgen05(const void* qk, int ld_qk, int k_off, const void* v, int ld_v, void* vt_scratch, void* out, int ld_o,
                                int B, int Q, int heads, int head_dim, void* stream) {
    if (head_dim != 32 || Q > AT_KPAD || Q < 1 || (ld_qk % 8) || (ld_v % 8) || (k_off % 8) || (ld_o % 8) || !vt_scratch ||
        (((uintptr_t)qk | (uintptr_t)v | (uintptr_t)out | (uintptr_t)vt_scratch) & 15)) {
        set_error("mha_tcgen05: unsupported shape (head_dim=%d Q=%d)", head_dim, Q);
        return DTLR_ERR_UNSUPPORTED;
    }
    if (B == 0) return DTLR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 tg(AT_KPAD / 32, heads, B);
    vt_transpose_kernel<<<tg, 256, 0, st>>>((const __nv_bfloat16*)v, ld_v, (__nv_bfloat16*)vt_scratch, Q, heads);
    DTLR_CHECK_LAUNCH();
    CUtensorMap tmQ, tmK, tmV;
    int rc;
    const long long rows = (long long)B * Q;
    if ((rc = make_tmap_2d_bf16(&tmQ, qk, rows, ld_qk, ld_qk, AT_QT, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if ((rc = make_tmap_2d_bf16(&tmK, qk, rows, ld_qk, ld_qk, 256, 32, CU_TENSOR_MAP_SWIZZLE_64B))) return rc;
    if ((rc = make_tmap_2d_bf16(&tmV, vt_scratch, (long long)B * heads * 32, AT_KPAD, AT_KPAD, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    static bool configured = false;
    if (!configured) {
        DTLR_CHECK_CUDA(cudaFuncSetAttribute(mha_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_TOTAL));
        configured = true;
    }
    const int n_qtiles = (Q + AT_QT - 1) / AT_QT;
    dim3 grid((n_qtiles + AT_TILES - 1) / AT_TILES, heads, B);
    const float scale_log2 = 1.4426950408889634f / sqrtf((float)head_dim);
    mha_tcgen05_kernel<<<grid, 576, AT_SMEM_TOTAL, st>>>(tmQ, tmK, tmV, (__nv_bfloat16*)out, ld_o, Q, heads, k_off, scale_log2);
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}
