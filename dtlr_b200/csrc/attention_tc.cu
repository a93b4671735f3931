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
#include <type_traits>

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

// V^T pre-pass: vt[(b*H + h)*32 + d][key] = v[b*Q + key][h*32 + d], zero for key >= Q.  One CTA = 64 keys of one (image, head):
// 16-byte loads (8 dims of a key), a padded shared-memory tile, 16-byte stores (8 keys of a dim; 128 contiguous bytes per dim row).
// (The first version moved single 16-bit elements: 33.7 us per call in the step's launch list for 64 MB of traffic.)
__global__ void __launch_bounds__(256) vt_transpose_kernel(const op16_t* __restrict__ v, int ld_v, op16_t* __restrict__ vt, int Q, int H) {
    __shared__ op16_t tile[64][40];                               // 80-byte pitch: 16-byte aligned rows, conflict-light column reads
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.z, h = blockIdx.y, k0 = blockIdx.x * 64;
    const int t = threadIdx.x;
    {
        const int key = k0 + (t >> 2), ch = t & 3;
        uint4 x = make_uint4(0, 0, 0, 0);
        if (key < Q) x = __ldg(reinterpret_cast<const uint4*>(v + ((size_t)b * Q + key) * ld_v + h * 32 + ch * 8));
        *reinterpret_cast<uint4*>(&tile[t >> 2][ch * 8]) = x;
    }
    __syncthreads();
    {
        const int d = t >> 3, kc = t & 7;
        const unsigned short* tp = reinterpret_cast<const unsigned short*>(&tile[kc * 8][d]);
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = (uint32_t)tp[(2 * i) * 40] | ((uint32_t)tp[(2 * i + 1) * 40] << 16);
        *reinterpret_cast<uint4*>(vt + ((size_t)(b * H + h) * 32 + d) * AT_KPAD + k0 + kc * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

__global__ void __launch_bounds__(576, 1)
mha_tcgen05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, op16_t* __restrict__ out, int ld_o, int Q, int H, int k_off,
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
        constexpr uint32_t IDESC_S = (1u << 4) | OP16_IDESC_AB | ((uint32_t)(AT_KC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t IDESC_O = (1u << 4) | OP16_IDESC_AB | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
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
                        op16x2_t a0 = op16_pack2(p[j], p[j + 1]), a1 = op16_pack2(p[j + 2], p[j + 3]);
                        op16x2_t a2 = op16_pack2(p[j + 4], p[j + 5]), a3 = op16_pack2(p[j + 6], p[j + 7]);
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
                op16_t* op = out + (size_t)(row_base + q) * ld_o + h * 32 + part * 8;
                uint32_t w[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    op16x2_t t0 = op16_pack2(__uint_as_float(oacc[2 * j]) * inv, __uint_as_float(oacc[2 * j + 1]) * inv);
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


// ---------------------------------------------------------------------------------------------- single-pass version
// The two-pass kernel above reads every score from TMEM twice and recomputes S.  Here every score is read ONCE:
//  * FOUR 128-query tiles are in flight per CTA (FlashAttention-4 style ping-pong, taken further because the head dimension is
//    only 32): 16 softmax warps, four per SM sub-partition;
//  * a softmax thread owns one query row and takes a whole 64-key chunk into registers (one tcgen05.ld.32x32b.x64), so the row
//    maximum needs no cross-warp exchange and the S buffer is handed back to the MMA warp right after the load;
//  * online softmax with the running maximum: when a chunk raises it, the 128 x 32 fp32 output accumulator in TMEM is rescaled by
//    the row's own thread (tcgen05.ld -> multiply -> tcgen05.st) -- cheap because the head dimension is 32 -- after the previous
//    P.V has completed and before the next one is issued;
//  * P is rounded to bf16 on the integer pipe (add 0x8000, high halves by one PRMT per pair; F2FP shares the SFU pipe with ex2)
//    and goes straight into the 128B-swizzled K-major A-operand tile (one 16 KB buffer per tile).
// TMEM: per tile 64 columns of S + 32 of O (4 x 96 = 384).  Warps: 0 TMA, 1 MMA issuer, 2-17 softmax (slot = (warp-2)/4).
// Measured (B = 64, Q = 900, CUDA-graph timing incl. the 20 us V^T pre-pass): 269 us vs 311 us two-pass vs 254 us for the
// mma.sync flash kernel, which therefore stays the default.  Probes (exponentials and / or TMEM loads replaced by constants)
// move the time by < 15 %: neither the SFU nor the TMEM read path bounds it, the per-chunk chain of mbarrier hand-offs between
// the softmax warps and the single MMA-issuing warp does (variants tried: 2 tiles with software-pipelined TMEM loads 419 us,
// 3 tiles with double-buffered S 294 us).  DESIGN.md 3.3.
constexpr int A2_KC = 64;                            // keys per chunk
constexpr int A2_SLOTS = 4;                          // query tiles in flight
constexpr int A2_SMEM_P = 16384;                     // P chunk of one tile: 128 rows x 128 B
constexpr int A2_TM = 128;                           // TMEM columns per slot: S chunk (64 fp32) | P chunk (32 packed 16-bit pairs) | O (32 fp32)
constexpr int A2_SMEM_TOTAL = AT_SMEM_K + AT_SMEM_V + A2_SLOTS * AT_SMEM_Q + 1024 + 512;     // (P lives in tensor memory)
static_assert(A2_SMEM_TOTAL <= 232448, "attention kernel shared memory exceeds 227 KB");

__global__ void __launch_bounds__(576, 1)
mha_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
               op16_t* __restrict__ out, int ld_o, int Q, int H, int k_off, float scale_log2, unsigned int* dbgbuf_, const int dbg) {
    extern __shared__ unsigned char at_raw[];
    unsigned int* const dbgbuf = (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) ? dbgbuf_ : nullptr;
#define A2_DBG(role, unit, slot) do { if (dbgbuf && (unit) < 16) dbgbuf[(((role) * 16 + (unit)) * 16 + (slot))] = (unsigned int)clock(); } while (0)
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)at_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* sK = smem;
    unsigned char* sV = sK + AT_SMEM_K;
    unsigned char* sQ = sV + AT_SMEM_V;                      // [slots][128 rows x 64 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sQ + A2_SLOTS * AT_SMEM_Q);
    uint64_t* kv_full = bars;                   // 1
    uint64_t* q_full = bars + 1;                // [slots]
    uint64_t* s_full = q_full + A2_SLOTS;       // commit: S chunk in TMEM
    uint64_t* s_empty = s_full + A2_SLOTS;      // 4 arrivals: S chunk is in registers
    uint64_t* p_full = s_empty + A2_SLOTS;      // 4 arrivals: P chunk in shared memory (and O rescaled)
    uint64_t* p_empty = p_full + A2_SLOTS;      // commit: P.V of the chunk complete
    uint64_t* o_full = p_empty + A2_SLOTS;      // commit: all P.V of the tile complete
    uint64_t* k_full = o_full + A2_SLOTS;       // [4] one per 256-key box of K: S = Q K^T starts when the FIRST box has landed,
    uint64_t* v_full = k_full + 4;              // [4] P.V when the first four V^T k-blocks have (not after all 128 KB of the head)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(v_full + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.z, h = blockIdx.y;
    const int tile0 = blockIdx.x * A2_SLOTS;
    const int n_qtiles = (Q + AT_QT - 1) / AT_QT;
    const int my_tiles = min(A2_SLOTS, n_qtiles - tile0);    // one tile per slot, one round per CTA
    const int row_base = b * Q;
    const int n_chunks = (Q + A2_KC - 1) / A2_KC;            // chunks that contain at least one real key

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
        mbar_init(kv_full, 1);
        for (int i = 0; i < 4; ++i) { mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); }
        for (int i = 0; i < A2_SLOTS; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4);
            mbar_init(&p_full[i], 4); mbar_init(&p_empty[i], 1);
            mbar_init(&o_full[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_ptr);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_launch_dependents();
    pdl_wait();

    // The P chunk goes back into TENSOR memory (tcgen05.st over its own 32 columns, packed 16-bit pairs) and is the A operand of
    // O += P V (tcgen05.mma [d], [a_tmem], b_desc) -- no shared-memory round trip, no proxy fence -- and the two MMA streams have
    // their own issuers: warp 1 issues S = Q K^T, warp 0 (idle after its TMA loads) issues O += P V.  The timeline of the first
    // version (tools/attn_timeline.py) showed ONE issuer warp pacing the kernel: 8 wait / issue / commit steps per 64-key round took
    // ~3,600 clk against 2,048 clk of exponentials.
    if (warp == 0) {
        // ===== TMA producer, then the P.V issuer =====
        if (elect_one()) {
            for (int sl = 0; sl < my_tiles; ++sl) {
                mbar_expect_tx(&q_full[sl], AT_SMEM_Q);
                tma_load_2d(sQ + sl * AT_SMEM_Q, &tmQ, &q_full[sl], h * 32, row_base + (tile0 + sl) * AT_QT);
            }
            for (int j = 0; j < 4; ++j) {                    // 256 keys at a time, in the order the chunks are consumed
                mbar_expect_tx(&k_full[j], AT_SMEM_K / 4);
                tma_load_2d(sK + j * 256 * 64, &tmK, &k_full[j], k_off + h * 32, row_base + j * 256);
                mbar_expect_tx(&v_full[j], AT_SMEM_V / 4);
                for (int i = 0; i < 4; ++i)
                    tma_load_2d(sV + (4 * j + i) * 4096, &tmV, &v_full[j], (4 * j + i) * 64, (b * H + h) * 32);
            }
        }
        __syncwarp();
        constexpr uint32_t IDESC_O = (1u << 4) | OP16_IDESC_AB | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        int v_ready = 0;
        const uint32_t aV = smem_u32(sV);
        // the four query tiles are independent pipelines: the issuer polls their barriers and serves whichever is ready, so a tile
        // that is late does not hold the others back (in fixed order the tiles stayed in step and took turns at the MUFU unit:
        // 2,800 clk per 64-key round against 2,048 clk of exponentials)
        int nc[A2_SLOTS] = {0, 0, 0, 0};
        for (int left = my_tiles * n_chunks; left > 0;) {
            const int before = left;
#pragma unroll
            for (int sl = 0; sl < A2_SLOTS; ++sl) {          // O_sl += P_sl(c) V_c
                const int c = nc[sl];
                if (sl >= my_tiles || c >= n_chunks || !mbar_test(&p_full[sl], c & 1)) continue;
                ++nc[sl];
                --left;
                while (v_ready <= (c >> 2)) mbar_wait(&v_full[v_ready++], 0);
                tcgen05_fence_after();
                A2_DBG(1, c + 1, sl * 2);
                if (elect_one()) {
                    const uint64_t dv = make_sw128_kmajor_desc(aV + c * 4096);
                    const uint32_t tP = tmem_base + sl * A2_TM + A2_KC, tO = tP + 32;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ts(tO, tP + k * 8, dv + (uint64_t)(2 * k), IDESC_O, (c == 0 && k == 0) ? 0u : 1u);
                    umma_commit(&p_empty[sl]);
                    if (c == n_chunks - 1) umma_commit(&o_full[sl]);
                }
                __syncwarp();
                A2_DBG(1, c + 1, sl * 2 + 1);
            }
            if (left == before) __nanosleep(40);             // nothing was ready: leave the issue slots to the softmax warps
        }
    } else if (warp == 1) {
        // ===== S = Q K^T issuer =====
        constexpr uint32_t IDESC_S = (1u << 4) | OP16_IDESC_AB | ((uint32_t)(A2_KC >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int sl = 0; sl < my_tiles; ++sl) mbar_wait(&q_full[sl], 0);
        tcgen05_fence_after();
        const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK);
        int nc[A2_SLOTS] = {0, 0, 0, 0};
        int k_ready = 0;
        for (int left = my_tiles * n_chunks; left > 0;) {
            const int before = left;
#pragma unroll
            for (int sl = 0; sl < A2_SLOTS; ++sl) {          // S_sl(c) = Q_sl K_c^T
                const int c = nc[sl];
                if (sl >= my_tiles || c >= n_chunks || !mbar_test(&s_empty[sl], (c & 1) ^ 1)) continue;
                ++nc[sl];
                --left;
                while (k_ready <= (c >> 2)) mbar_wait(&k_full[k_ready++], 0);
                tcgen05_fence_after();
                A2_DBG(0, c, sl * 2);
                if (elect_one()) {
                    const uint64_t dq = make_sw64_kmajor_desc(aQ + sl * AT_SMEM_Q);
                    const uint64_t dk = make_sw64_kmajor_desc(aK + c * A2_KC * 64);
                    umma_bf16(tmem_base + sl * A2_TM, dq, dk, IDESC_S, 0);
                    umma_bf16(tmem_base + sl * A2_TM, dq + 2, dk + 2, IDESC_S, 1);
                    umma_commit(&s_full[sl]);
                }
                __syncwarp();
                A2_DBG(0, c, sl * 2 + 1);
            }
            if (left == before) __nanosleep(40);
        }
    } else {
        // ===== softmax warps: slot sl, TMEM lane quarter qd, thread = query row =====
        const int sl = (warp - 2) >> 2;
        const int qd = warp & 3;
        const int row = qd * 32 + lane;
        const uint32_t tS = tmem_base + sl * A2_TM + ((uint32_t)(qd * 32) << 16), tP = tS + A2_KC, tO = tP + 32;
        if (sl < my_tiles) {
            const int q = (tile0 + sl) * AT_QT + row;
            float m = -INFINITY, l = 0.f;
            const int drole = (warp == 2) ? 2 : ((warp == 14) ? 3 : -1);
#define A2_DBGW(unit, slot) do { if (drole >= 0 && lane == 0) A2_DBG(drole, unit, slot); } while (0)
            // Q = 900 is 7 tiles of 128 queries + 4: three of the last tile's four warps own no query at all.  They keep the barrier
            // protocol going and skip every TMEM access and every exponential (the kernel is bound by the MUFU unit: 3 of a pair of
            // CTAs' 32 softmax warps = 9 % of the work); their P / O rows hold garbage that only their own (never stored) rows see
            const bool live = (tile0 + sl) * AT_QT + qd * 32 < Q;
            // one chunk of the online softmax over the first NK key columns of the chunk (NK = 64, or 16 for a short last chunk:
            // Q = 900 leaves 4 keys in chunk 14); the remaining P columns are zero
            auto chunk_body = [&](auto nk_tag, const int c) {
                constexpr int NK = decltype(nk_tag)::value;
                uint32_t a[64];
                tmem_ld64(tS, a);
                A2_DBGW(c, 2);
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_empty[sl]);                // the S buffer may be overwritten by the next chunk
                const int key0 = c * A2_KC;
                if (key0 + NK > Q) {                                      // boundary chunk: keys beyond Q do not exist
#pragma unroll
                    for (int j = 0; j < NK; ++j)
                        if (key0 + j >= Q) a[j] = 0xff800000u;
                }
                float c0 = -INFINITY, c1 = -INFINITY, c2 = -INFINITY, c3 = -INFINITY;
#pragma unroll
                for (int j = 0; j < NK; j += 4) {
                    c0 = fmaxf(c0, __uint_as_float(a[j])); c1 = fmaxf(c1, __uint_as_float(a[j + 1]));
                    c2 = fmaxf(c2, __uint_as_float(a[j + 2])); c3 = fmaxf(c3, __uint_as_float(a[j + 3]));
                }
                const float m_new = fmaxf(fmaxf(m, fmaxf(c0, c1)), fmaxf(c2, c3));
                const float alpha = ex2_approx((m - m_new) * scale_log2);     // 0 on the first chunk (m = -inf), 1 when unchanged
                const float mxs = m_new * scale_log2;
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                for (int j = 0; j < NK; j += 4) {                             // P overwrites S in place: word j/2 = keys j, j+1
                    const float p0 = ex2_approx(fmaf(__uint_as_float(a[j]), scale_log2, -mxs)), p1 = ex2_approx(fmaf(__uint_as_float(a[j + 1]), scale_log2, -mxs));
                    const float p2 = ex2_approx(fmaf(__uint_as_float(a[j + 2]), scale_log2, -mxs)), p3 = ex2_approx(fmaf(__uint_as_float(a[j + 3]), scale_log2, -mxs));
                    s0 += p0; s1 += p1; s2 += p2; s3 += p3;
#ifdef DTLR_BUILD_F16
                    { const op16x2_t t0 = op16_pack2(p0, p1), t1 = op16_pack2(p2, p3);
                      a[j / 2] = *reinterpret_cast<const uint32_t*>(&t0); a[j / 2 + 1] = *reinterpret_cast<const uint32_t*>(&t1); }
#else
                    // round to nearest bf16 on the integer pipe (p >= 0, finite): + 0x8000, keep the high halves
                    a[j / 2] = __byte_perm(__float_as_uint(p0) + 0x8000u, __float_as_uint(p1) + 0x8000u, 0x7632);
                    a[j / 2 + 1] = __byte_perm(__float_as_uint(p2) + 0x8000u, __float_as_uint(p3) + 0x8000u, 0x7632);
#endif
                }
#pragma unroll
                for (int j = NK / 2; j < 32; ++j) a[j] = 0u;
                l = l * alpha + ((s0 + s1) + (s2 + s3));
                m = m_new;
                A2_DBGW(c, 3);
                // the previous P.V has completed: P buffer free, O may be rescaled
                mbar_wait(&p_empty[sl], (c & 1) ^ 1);
                tcgen05_fence_after();
                A2_DBGW(c, 4);
                if (c > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
                    uint32_t o[32];
                    tmem_ld32(tO, o);
#pragma unroll
                    for (int j = 0; j < 32; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * alpha);
                    tmem_st32(tO, o);
                }
                A2_DBGW(c, 5);
                tmem_st32(tP, *reinterpret_cast<uint32_t(*)[32]>(a));      // P chunk: packed pairs, column j = keys 2j, 2j + 1
            };
            for (int c = 0; c < n_chunks; ++c) {
                A2_DBGW(c, 0);
                mbar_wait(&s_full[sl], c & 1);
                tcgen05_fence_after();
                A2_DBGW(c, 1);
                if (live) {
                    if (Q - c * A2_KC <= 16) chunk_body(std::integral_constant<int, 16>{}, c);
                    else chunk_body(std::integral_constant<int, 64>{}, c);
                } else {
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_empty[sl]);
                    mbar_wait(&p_empty[sl], (c & 1) ^ 1);
                    tcgen05_fence_after();
                }
                tcgen05_fence_before();
                __syncwarp();
                A2_DBGW(c, 6);
                if (lane == 0) mbar_arrive(&p_full[sl]);
                A2_DBGW(c, 7);
            }
            // ---- epilogue: O / l -> bf16 -> global (64 bytes per row)
            mbar_wait(&o_full[sl], 0);
            tcgen05_fence_after();
            uint32_t o[32];
            tmem_ld32(tO, o);
            if (q < Q) {
                const float inv = 1.f / l;
                uint4* op = reinterpret_cast<uint4*>(out + (size_t)(row_base + q) * ld_o + h * 32);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t w[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        op16x2_t t0 = op16_pack2(__uint_as_float(o[j * 8 + 2 * i]) * inv, __uint_as_float(o[j * 8 + 2 * i + 1]) * inv);
                        w[i] = *reinterpret_cast<uint32_t*>(&t0);
                    }
                    op[j] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace dtlr

using namespace dtlr;

// timeline probe of mha_tc2_kernel (tools/attn_timeline.py): a device buffer of 4 * 16 * 16 u32, or NULL (off)
static unsigned int* g_attn_dbgbuf = nullptr;
extern "C" int dtlr_attn_debug_buffer(void* buf) { g_attn_dbgbuf = reinterpret_cast<unsigned int*>(buf); return DTLR_OK; }

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
    dim3 tg(AT_KPAD / 64, heads, B);
    DTLR_CHECK_CUDA(launch_pdl(vt_transpose_kernel, tg, dim3(256), 0, st, (const op16_t*)v, ld_v, (op16_t*)vt_scratch, Q, heads));
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
    if (!(g_debug_flags & 256)) {                 // single-pass kernel (default of this path); flag 256: the two-pass kernel
        static bool configured2 = false;
        if (!configured2) {
            DTLR_CHECK_CUDA(cudaFuncSetAttribute(mha_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A2_SMEM_TOTAL));
            configured2 = true;
        }
        grid.x = (n_qtiles + A2_SLOTS - 1) / A2_SLOTS;
        DTLR_CHECK_CUDA(launch_pdl(mha_tc2_kernel, grid, dim3(576), A2_SMEM_TOTAL, st, tmQ, tmK, tmV, (op16_t*)out, ld_o, Q, heads, k_off, scale_log2, g_attn_dbgbuf, g_debug_flags));
        return DTLR_OK;
    }
    mha_tcgen05_kernel<<<grid, 576, AT_SMEM_TOTAL, st>>>(tmQ, tmK, tmV, (op16_t*)out, ld_o, Q, heads, k_off, scale_log2);
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}
