// dtlr_b200 -- the position-wise feed-forward block of a transformer layer as ONE tcgen05 kernel for sm_100a:
//
//     Y = LayerNorm( X + W2 . relu(W1 . X + b1) + b2 ) * gamma + beta          X, Y: [M, 256] bf16, hidden = HID
//
// which is `forward_ffn` + `norm` of reference models/dino/deformable_transformer.py:804-808,816-817 (encoder layer) and
// :876-880 (decoder layer); d_model 256, dim_feedforward 2048, ReLU, dropout 0, post-norm.
//
// Why fused: as two GEMMs the hidden activation (M x 2048 bf16 = 239 MB at B = 64) is written to HBM and read back; measured
// (tools/gemm_probe.py) linear1 is bound by that write (64 us, HBM floor 42 us) and linear2 by re-reading it (78-90 us).
// Here the hidden activation never leaves the SM -- and never touches shared memory either:
//
//   per 128-row tile of X (shared memory, 64 KB, double-buffered: A operand of linear1, the residual, finally the output staging):
//     for each chunk j of 128 hidden units (HID / 128 chunks, software-pipelined by one chunk):
//       G1(j):  Hacc[j&1] (TMEM, 128 x 128 fp32)  = X . W1[j]^T                      16 x tcgen05.mma 128x128x16 (SS)
//       E1(j):  8 epilogue warps: tcgen05.ld -> +b1 -> ReLU -> bf16 -> tcgen05.st back IN PLACE over the first 64 columns of
//               the accumulator it came from (thread = row; the two warps of a lane quarter meet at a 64-thread barrier)
//       G2(j):  Yacc (TMEM, 128 x 256 fp32)      += H[j&1] . W2[:, j]^T               16 x tcgen05.mma 128x128x16 with the A
//               operand read from TENSOR MEMORY (tcgen05.mma [d], [a_tmem], b_desc): no shared-memory round trip for H
//     final:   Yacc + b2 + X -> bf16 (packed in registers, ONE TMEM read; the accumulator is released to the next tile at once)
//              -> LayerNorm (row statistics: thread = row, the two column halves meet in shared memory) -> written over the X
//              tile in place -> TMA store.  The next tile's X is already resident in the other buffer and its G1 runs meanwhile.
//   W1 / W2 chunks (2 MB per tile in total, L2-resident) stream through a 5-stage TMA ring of 16 KB stages.
//   TMEM: Yacc 256 columns + 2 x 128 columns of Hacc / H = 512.
//
// Measured history (M = 58368, CUDA-graph timing): hidden chunk through shared memory 146 us (shared-memory-bandwidth bound:
// 6.5 MB of operand / TMA / epilogue traffic per tile) -> through TMEM 135 us -> see DESIGN.md 3.2b for the current number;
// un-fused linear1 + linear2/LN: 147 us.
//
// Warp roles as in gemm.cu: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-9 = epilogue.
#include "tc_common.cuh"

namespace dtlr {

constexpr int FF_D = 256;          // d_model (one full LayerNorm row per tile)
constexpr int FF_BM = 128;
constexpr int FF_HC = 128;         // hidden units per chunk
constexpr int FF_STAGE = 16384;    // ring stage: 128 rows x 64 k (bf16, 128B swizzle)
constexpr int FF_NS = 5;           // ring depth
constexpr int FF_MAX_HID = 2048;

struct FfnArgs {
    const float* b1;      // [HID]
    const float* b2;      // [256]
    const float* gamma;   // [256]
    const float* beta;    // [256]
    float eps;
    int M, HID;
    int dbg;              // probes (dtlr_debug_flags): 256 no E1 work, 512 no G1 MMAs, 1024 no G2 MMAs, 2048 no final epilogue work
    // HEAD variant (3-layer MLP head with a 4-wide last layer, reference models/dino/utils.py:110-122 as used for bbox_embed):
    const float* w3;      // [4][256] fp32 (nn.Linear weight of the last layer)
    const float* b3;      // [4]
    const float* ref;     // [M][4] reference boxes in (0,1) or null: out = sigmoid(delta + inverse_sigmoid(ref)) (box refinement)
    float* out4;          // [M][4]
    // PART variant (wave-quantisation tail, see dtlr_ffn_ln): CTA u works on row tile tile_base + u / nsplit and on the hidden slice
    // (u % nsplit) of width HID (the LOCAL width; b1 / W1 / W2 are offset by the slice) and stores its raw fp32 partial of
    // W2 . relu(W1 X + b1) to partial[slice][tile - tile_base][128][256]
    int tile_base, nsplit;
    float* partial;
    // stream-K kernel (ffn_ln_sk_kernel): one ready flag per CTA for the partial it leaves to its left neighbour
    int* flags;
    // timeline probe (dtlr_ffn_debug_buffer): CTA 0 records clock() at fixed points of its three roles, [role][unit 64][slot 16]
    unsigned int* dbgbuf;
};

struct FfnSmem {
    static constexpr int XS = 4 * FF_STAGE;                 // one X tile: 4 k-blocks of 128 x 64
    static constexpr int RING = FF_NS * FF_STAGE;
    static constexpr int B1 = FF_MAX_HID * 4;
    static constexpr int VEC = 3 * FF_D * 4;                // b2, gamma, beta
    static constexpr int STAT = FF_BM * 2 * 2 * 4;          // per row, per column half: sum, sum of squares
    static constexpr int BAR = 512;
    static constexpr int TOTAL = 1024 + 2 * XS + RING + B1 + VEC + STAT + BAR;
};
static_assert(FfnSmem::TOTAL <= 232448, "FFN kernel shared memory exceeds 227 KB");

__device__ __forceinline__ uint32_t ff_pack_bf16x2(float lo, float hi) {
    op16x2_t t = op16_pack2(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}

// CL: launched as clusters of two CTAs that stream the SAME weight chunks: each ring stage is fetched from L2 once by one CTA
// of the pair (alternating) and multicast into both, halving the L2 -> SM traffic (912 MB per call otherwise); a stage is refilled
// once BOTH tensor cores have released it (multicast tcgen05.commit on both CTAs' w_empty barriers)
// HEAD: the same two chained contractions with HID = 256 as the first two layers of a 3-layer MLP head -- Y = relu(W2 relu(W1 X + b1)
// + b2) stays in registers, the 4-wide last layer is 4 dot products per row on the FMA pipe, followed (optionally) by the box
// refinement sigmoid(delta + inverse_sigmoid(ref)) of reference deformable_transformer.py:734-738 / dino.py:343-345: one kernel reads
// X once and writes 16 bytes per row, instead of three GEMMs (two 256-wide intermediates through HBM) + an elementwise kernel.
template <bool CL, int VAR>
__global__ void __launch_bounds__(320, 1)
ffn_ln_tcgen05_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                      const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmO, const FfnArgs a) {
    extern __shared__ unsigned char smem_raw[];
    constexpr int NS = FF_NS;
    constexpr bool HEAD = VAR == 1, PART = VAR == 2;
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* xs = smem;                                // [2][4 k-blocks][128 rows x 128 B]
    unsigned char* ring = xs + 2 * FfnSmem::XS;
    float* b1_s = reinterpret_cast<float*>(ring + FfnSmem::RING);
    float* b2_s = b1_s + FF_MAX_HID;
    float* gamma_s = b2_s + FF_D;
    float* beta_s = gamma_s + FF_D;
    float* stat_s = beta_s + FF_D;                           // [128 rows][2 halves][2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(stat_s + FF_BM * 4);
    uint64_t* x_full = bars;             // [2]
    uint64_t* x_free = bars + 2;         // [2]  8 arrivals (epilogue warps): the X tile (output staging) may be overwritten
    uint64_t* w_full = bars + 4;         // [NS]
    uint64_t* w_empty = w_full + NS;     // [NS]
    uint64_t* hacc_full = w_empty + NS;  // [2] G1 complete: fp32 hidden chunk in TMEM
    uint64_t* h_full = hacc_full + 2;    // [2] 8 arrivals: bf16 hidden chunk written back to TMEM
    uint64_t* y_full = h_full + 2;       // [1] all G2 of the tile complete
    uint64_t* y_free = y_full + 1;       // [1] 8 arrivals: output accumulator read out
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(y_free + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_m = (a.M + FF_BM - 1) / FF_BM;
    const int NJ = a.HID / FF_HC;
    // row tiles of this CTA: persistent stride over all tiles, or (PART) exactly one tile and one hidden slice of NJ chunks from jo
    const int mt_first = PART ? a.tile_base + (int)blockIdx.x / a.nsplit : (int)blockIdx.x;
    const int mt_step = PART ? (1 << 28) : (int)gridDim.x;
    const int jo = PART ? ((int)blockIdx.x % a.nsplit) * NJ : 0;
    // CL: both CTAs of a pair (even grid, rank = blockIdx.x & 1) must consume the same number of weight chunks: the pair works on
    // tiles (2p, 2p+1), (2p, 2p+1) + grid, ... while the EVEN tile exists; an odd tile == num_m is a ghost (X rows zero-filled by
    // TMA, stores clipped by TMA)
    auto more = [&](int mt) { return CL ? ((mt & ~1) < num_m) : (mt < num_m); };
    const uint32_t cta_rank = CL ? cluster_cta_rank() : 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW1);
        tma_prefetch_desc(&tmW2);
        tma_prefetch_desc(&tmO);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&x_full[b], 1);
            mbar_init(&x_free[b], 8);
            mbar_init(&hacc_full[b], 1);
            mbar_init(&h_full[b], 8);
        }
        for (int s = 0; s < NS; ++s) {
            mbar_init(&w_full[s], 1);
            mbar_init(&w_empty[s], CL ? 2 : 1);
        }
        mbar_init(y_full, 1);
        mbar_init(y_free, 8);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_ptr);
    for (int i = threadIdx.x; i < a.HID; i += 320) b1_s[i] = __ldg(a.b1 + jo * FF_HC + i);
    float* w3_s = b1_s + FF_D;                       // HEAD: [256 columns][4 outputs] behind the 256 b1 entries (HID == 256)
    for (int i = threadIdx.x; i < FF_D; i += 320) {
        b2_s[i] = __ldg(a.b2 + i);
        if (!HEAD) {
            gamma_s[i] = __ldg(a.gamma + i);
            beta_s[i] = __ldg(a.beta + i);
        } else {
#pragma unroll
            for (int n = 0; n < 4; ++n) w3_s[i * 4 + n] = __ldg(a.w3 + n * FF_D + i);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (CL) cluster_sync_all();                      // the peer's barriers are initialised before anything is multicast at them
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t tm_y = tmem_base;                 // columns [0, 256)
    const uint32_t tm_h = tmem_base + 256;           // 2 x 128 columns

    if (warp == 0) {
        // ===== TMA producer: weight chunks in exactly the order the MMA warp consumes them; the X tile of the NEXT row tile is
        //       requested half way through the current one (its buffer was released by the tile before)
        if (elect_one()) {
            uint32_t it = 0, t = 0;
            auto load_x = [&](int mt, uint32_t tt) {
                const uint32_t xb = tt & 1;
                mbar_wait(&x_free[xb], ((tt >> 1) & 1) ^ 1);
                mbar_expect_tx(&x_full[xb], FfnSmem::XS);
                for (int kb = 0; kb < 4; ++kb) tma_load_2d(xs + xb * FfnSmem::XS + kb * FF_STAGE, &tmX, &x_full[xb], kb * 64, mt * FF_BM);
            };
            if (more(mt_first)) load_x(mt_first, 0);
            for (int mt = mt_first; more(mt); mt += mt_step, ++t) {
                for (int j = 0; j <= NJ; ++j) {
                    if (j == NJ / 2 && more(mt + mt_step)) load_x(mt + mt_step, t + 1);
                    if (j < NJ) {
                        for (int kb = 0; kb < 4; ++kb, ++it) {              // W1 rows j*128.., k columns kb*64..
                            const int s = it % NS;
                            mbar_wait(&w_empty[s], ((it / NS) & 1) ^ 1);
                            mbar_expect_tx(&w_full[s], FF_STAGE);
                            if (!CL) tma_load_2d(ring + s * FF_STAGE, &tmW1, &w_full[s], kb * 64, (jo + j) * FF_HC);
                            else if ((it & 1) == cta_rank) tma_load_2d_multicast(ring + s * FF_STAGE, &tmW1, &w_full[s], kb * 64, (jo + j) * FF_HC, 3);
                        }
                    }
                    if (j >= 1) {
                        for (int q = 0; q < 4; ++q, ++it) {                 // W2 output rows (q&1)*128.., hidden columns of chunk j-1
                            const int s = it % NS;
                            mbar_wait(&w_empty[s], ((it / NS) & 1) ^ 1);
                            mbar_expect_tx(&w_full[s], FF_STAGE);
                            if (!CL) tma_load_2d(ring + s * FF_STAGE, &tmW2, &w_full[s], (jo + j - 1) * FF_HC + (q >> 1) * 64, (q & 1) * 128);
                            else if ((it & 1) == cta_rank)
                                tma_load_2d_multicast(ring + s * FF_STAGE, &tmW2, &w_full[s], (jo + j - 1) * FF_HC + (q >> 1) * 64, (q & 1) * 128, 3);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: G1(j) one chunk ahead of G2(j-1).  Hacc[b] is rewritten by G1(j+2) only after G2(j), its last reader, in
        //       issue order; E1(j) finished with it before h_full(j), which G2(j) waited for
        constexpr uint32_t IDESC = (1u << 4) | OP16_IDESC_AB | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(FF_BM >> 4) << 24);
        uint32_t it = 0, g = 0, t = 0;
        for (int mt = mt_first; more(mt); mt += mt_step, ++t, g += NJ) {
            const uint32_t xb = t & 1;
            mbar_wait(&x_full[xb], (t >> 1) & 1);
            tcgen05_fence_after();
            for (int j = 0; j <= NJ; ++j) {
                if (j < NJ) {
                    const uint32_t b = (g + j) & 1;
                    for (int kb = 0; kb < 4; ++kb, ++it) {
                        const int s = it % NS;
                        mbar_wait(&w_full[s], (it / NS) & 1);
                        tcgen05_fence_after();
                        if (elect_one()) {
                            const uint64_t da = make_sw128_kmajor_desc(smem_u32(xs + xb * FfnSmem::XS + kb * FF_STAGE));
                            const uint64_t db = make_sw128_kmajor_desc(smem_u32(ring + s * FF_STAGE));
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (a.dbg & 512) break;
                                umma_bf16(tm_h + b * FF_HC, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDESC, (kb | k) != 0);
                            }
                            if (CL) umma_commit_multicast(&w_empty[s], 3); else umma_commit(&w_empty[s]);
                            if (kb == 3) umma_commit(&hacc_full[b]);
                        }
                        __syncwarp();
                    }
                }
                if (j >= 1) {
                    const uint32_t gj = g + j - 1, b = gj & 1, u = gj >> 1;
                    mbar_wait(&h_full[b], u & 1);                            // bf16 hidden chunk j-1 is in TMEM
                    if (j == 1) mbar_wait(y_free, (t & 1) ^ 1);              // previous tile's output accumulator read out
                    tcgen05_fence_after();
                    for (int q = 0; q < 4; ++q, ++it) {
                        const int s = it % NS;
                        const int kb2 = q >> 1, half = q & 1;
                        mbar_wait(&w_full[s], (it / NS) & 1);
                        tcgen05_fence_after();
                        if (elect_one()) {
                            const uint64_t db = make_sw128_kmajor_desc(smem_u32(ring + s * FF_STAGE));
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (a.dbg & 1024) break;
                                // A = bf16 hidden chunk in TMEM: the 64 hidden units of k-block kb2 = 32 columns, 8 per K = 16 step
                                umma_bf16_ts(tm_y + half * 128, tm_h + b * FF_HC + kb2 * 32 + k * 8, db + (uint64_t)(2 * k), IDESC,
                                             (j > 1) || kb2 > 0 || k > 0);
                            }
                            if (CL) umma_commit_multicast(&w_empty[s], 3); else umma_commit(&w_empty[s]);
                            if (q == 3 && j == NJ) umma_commit(y_full);
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else {
        // ===== epilogue warps: TMEM lane quarter qd (32 rows), column half hsel
        const int qd = warp & 3;
        const int hsel = (warp - 2) >> 2;
        const int row = qd * 32 + lane;
        const uint32_t swz = (uint32_t)(lane & 7);
        const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
        uint32_t g = 0, t = 0;
        for (int mt = mt_first; more(mt); mt += mt_step, ++t, g += NJ) {
            // ---- E1: hidden chunk j: +b1, ReLU, bf16, back into TMEM in place
            for (int j = 0; j < NJ; ++j) {
                const uint32_t gj = g + j, b = gj & 1, u = gj >> 1;
                mbar_wait(&hacc_full[b], u & 1);
                tcgen05_fence_after();
                if (a.dbg & 256) {
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&h_full[b]);
                    continue;
                }
                uint32_t acc[64];
                tmem_ld64(tm_h + b * FF_HC + lane_addr + (uint32_t)(hsel * 64), acc);
                const float* bp = b1_s + j * FF_HC + hsel * 64;
                uint32_t pk[32];
#pragma unroll
                for (int c = 0; c < 64; c += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bp + c);
                    const float v0 = fmaxf(__uint_as_float(acc[c]) + b4.x, 0.f), v1 = fmaxf(__uint_as_float(acc[c + 1]) + b4.y, 0.f);
                    const float v2 = fmaxf(__uint_as_float(acc[c + 2]) + b4.z, 0.f), v3 = fmaxf(__uint_as_float(acc[c + 3]) + b4.w, 0.f);
                    pk[c / 2] = ff_pack_bf16x2(v0, v1);
                    pk[c / 2 + 1] = ff_pack_bf16x2(v2, v3);
                }
                // this warp's 64 hidden units -> 32 packed columns at hsel*32 of the accumulator; the partner warp of the lane quarter
                // (the other 64 fp32 columns, which the columns written here overlap) must have finished ITS load first
                asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");
                tmem_st32(tm_h + b * FF_HC + lane_addr + (uint32_t)(hsel * 32), pk);
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&h_full[b]);
            }
            // ---- final: Yacc + b2 + X -> LayerNorm -> bf16 over the X tile -> TMA store
            const uint32_t xb = t & 1;
            unsigned char* xt = xs + xb * FfnSmem::XS;
            if (PART) {
                // ---- final (PART): the raw fp32 partial of this hidden slice -> scratch (summed + normalised by ffn_tail_ln_kernel)
                mbar_wait(y_full, t & 1);
                tcgen05_fence_after();
                if (lane == 0) mbar_arrive(&x_free[xb]);
                float* dst = a.partial + (((size_t)((int)blockIdx.x % a.nsplit) * (size_t)(gridDim.x / a.nsplit) + (size_t)(mt - a.tile_base)) * FF_BM + row) * FF_D;
#pragma unroll
                for (int cb = 0; cb < 2; ++cb) {
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        uint32_t acc[32];
                        tmem_ld32(tm_y + lane_addr + (uint32_t)(hsel * 128 + cb * 64 + hf * 32), acc);
                        if (cb == 1 && hf == 1) {
                            tcgen05_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(y_free);
                        }
                        float4* d4 = reinterpret_cast<float4*>(dst + hsel * 128 + cb * 64 + hf * 32);
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            d4[i] = make_float4(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1]), __uint_as_float(acc[4 * i + 2]), __uint_as_float(acc[4 * i + 3]));
                    }
                }
                continue;
            }
            if (HEAD) {
                // ---- final (HEAD): relu(Yacc + b2) . W3^T + b3 [-> box refinement] -> 16 bytes per row
                mbar_wait(y_full, t & 1);                    // every MMA of the tile is complete: X is no longer read either
                tcgen05_fence_after();
                if (lane == 0) mbar_arrive(&x_free[xb]);
                float o4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int cb = 0; cb < 2; ++cb) {
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        uint32_t acc[32];
                        tmem_ld32(tm_y + lane_addr + (uint32_t)(hsel * 128 + cb * 64 + hf * 32), acc);
                        if (cb == 1 && hf == 1) {
                            tcgen05_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(y_free);
                        }
                        const int c0 = hsel * 128 + cb * 64 + hf * 32;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float v = fmaxf(__uint_as_float(acc[i]) + b2_s[c0 + i], 0.f);
                            const float4 w = *reinterpret_cast<const float4*>(w3_s + (c0 + i) * 4);
                            o4[0] = fmaf(v, w.x, o4[0]); o4[1] = fmaf(v, w.y, o4[1]);
                            o4[2] = fmaf(v, w.z, o4[2]); o4[3] = fmaf(v, w.w, o4[3]);
                        }
                    }
                }
                if (hsel == 1) *reinterpret_cast<float4*>(stat_s + row * 4) = make_float4(o4[0], o4[1], o4[2], o4[3]);
                asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");   // the two column halves of a row meet
                if (hsel == 0) {
                    const float4 p = *reinterpret_cast<const float4*>(stat_s + row * 4);
                    const long long rg = (long long)mt * FF_BM + row;
                    if (rg < a.M) {
                        float z[4] = {o4[0] + p.x + __ldg(a.b3), o4[1] + p.y + __ldg(a.b3 + 1), o4[2] + p.z + __ldg(a.b3 + 2), o4[3] + p.w + __ldg(a.b3 + 3)};
                        if (a.ref) {
                            const float4 r4 = __ldg(reinterpret_cast<const float4*>(a.ref) + rg);
                            const float rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                            for (int n = 0; n < 4; ++n) {
                                const float x = fminf(fmaxf(rr[n], 0.f), 1.f);
                                const float x1 = fmaxf(x, 1e-3f), x2 = fmaxf(1.f - x, 1e-3f);
                                z[n] = 1.f / (1.f + expf(-(z[n] + logf(x1 / x2))));
                            }
                        }
                        reinterpret_cast<float4*>(a.out4)[rg] = make_float4(z[0], z[1], z[2], z[3]);
                    }
                }
                asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");   // partials consumed before the next tile overwrites them
                continue;
            }
            mbar_wait(&x_full[xb], (t >> 1) & 1);
            mbar_wait(y_full, t & 1);
            tcgen05_fence_after();
            if (a.dbg & 2048) {
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(y_free); mbar_arrive(&x_free[xb]); }
                continue;
            }
            float sum = 0.f, sq = 0.f;
            uint32_t xp[64];                                 // the row's 128 pre-norm values of this warp, packed bf16x2
#pragma unroll
            for (int cb = 0; cb < 2; ++cb) {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t acc[32];
                    tmem_ld32(tm_y + lane_addr + (uint32_t)(hsel * 128 + cb * 64 + hf * 32), acc);
                    if (cb == 1 && hf == 1) {                // the output accumulator is in registers: the next tile's G2 may start
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(y_free);
                    }
                    const unsigned char* xrow = xt + (hsel * 2 + cb) * FF_STAGE + row * 128;
                    const float* bp = b2_s + hsel * 128 + cb * 64 + hf * 32;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const int k = hf * 4 + kk;
                        const uint4 r4 = *reinterpret_cast<const uint4*>(xrow + ((k ^ swz) * 16));
                        const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float x0 = __uint_as_float(acc[kk * 8 + 2 * i]) + bp[kk * 8 + 2 * i] + op16_lo_f32(rw[i]);
                            const float x1 = __uint_as_float(acc[kk * 8 + 2 * i + 1]) + bp[kk * 8 + 2 * i + 1] + op16_hi_f32(rw[i]);
                            const uint32_t pk = ff_pack_bf16x2(x0, x1);
                            xp[cb * 32 + k * 4 + i] = pk;
                            const float y0 = op16_lo_f32(pk), y1 = op16_hi_f32(pk);
                            sum += y0 + y1;
                            sq = fmaf(y0, y0, sq);
                            sq = fmaf(y1, y1, sq);
                        }
                    }
                }
            }
            stat_s[(row * 2 + hsel) * 2] = sum;
            stat_s[(row * 2 + hsel) * 2 + 1] = sq;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");       // the two warps of this lane quarter
            const float mean = (stat_s[row * 4] + stat_s[row * 4 + 2]) * (1.f / 256.f);
            const float var = fmaxf((stat_s[row * 4 + 1] + stat_s[row * 4 + 3]) * (1.f / 256.f) - mean * mean, 0.f);
            const float rstd = rsqrtf(var + a.eps);
            asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");       // statistics consumed before the next tile overwrites them
#pragma unroll
            for (int cb = 0; cb < 2; ++cb) {
                unsigned char* xrow = xt + (hsel * 2 + cb) * FF_STAGE + row * 128;
                const int c0 = hsel * 128 + cb * 64;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    uint32_t o[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t pk = xp[cb * 32 + k * 4 + i];
                        const int c = c0 + k * 8 + 2 * i;
                        o[i] = ff_pack_bf16x2((op16_lo_f32(pk) - mean) * rstd * gamma_s[c] + beta_s[c],
                                              (op16_hi_f32(pk) - mean) * rstd * gamma_s[c + 1] + beta_s[c + 1]);
                    }
                    *reinterpret_cast<uint4*>(xrow + ((k ^ swz) * 16)) = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                for (int cb = 0; cb < 2; ++cb)
                    tma_store_2d(&tmO, xt + (hsel * 2 + cb) * FF_STAGE + (qd * 32) * 128, (hsel * 2 + cb) * 64, mt * FF_BM + qd * 32);
                tma_store_commit();
                tma_store_wait_read<0>();                                    // this X buffer may now be refilled
                mbar_arrive(&x_free[xb]);
            }
            __syncwarp();
        }
        if (lane == 0) tma_store_wait<0>();
    }
    tcgen05_fence_before();
    __syncthreads();
    if (CL) cluster_sync_all();                      // no CTA leaves while its peer may still multicast data / arrivals into it
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------------------------------------
// Stream-K variant of the block above (the default for more than one round of tiles).  Wave quantisation: num_m
// 128-row tiles on G = 148 CTAs take ceil(num_m / G) rounds -- 456 tiles at B = 64 are 3.08 waves, i.e. a fourth round for 12 tiles.
// Here the work is cut in UNITS = (row tile, hidden chunk of 128) instead: CTA c owns the contiguous unit range
// [c U / G, (c + 1) U / G) of the U = num_m * NJ units, so every CTA does the same amount of tensor-core work (49 or 50 units at
// B = 64) and a row tile whose chunks straddle a range boundary is shared by exactly two neighbours:
//   * CTA c + 1 starts its range in the middle of tile T: it runs the LAST chunks of T first ("tail part"), stores the raw fp32
//     partial of W2 . relu(W1 X + b1) over those chunks to workspace slot c + 1 and raises flag c + 1 (release);
//   * CTA c ends its range with the FIRST chunks of T ("head part", the last thing it does -- by then the neighbour's partial has
//     been in L2 for ~100 us): acquire flag, add the partial to its accumulator, + b2 + X -> LayerNorm -> store; flag reset to 0.
// No CTA waits on a CTA that waits (the tail part is the first thing a CTA does and depends on nothing), all G <= SM count CTAs are
// co-resident (1 CTA / SM by shared memory), so the spin cannot deadlock; the summation order is fixed (deterministic).
// The unit stream is software-pipelined ACROSS tile boundaries: G1 of unit v + 1 is issued before G2 of unit v whatever tiles they
// belong to, so the tensor pipe no longer idles between the last G2 of a tile and the first G1 of the next one.
// barrier wait of the stream-K kernel: cluster-scope acquire when arrivals / transaction bytes come from the peer CTA (PAIR)
constexpr bool g_dbg_cluster_scope = false;       // true: cluster-scope acquire on every wait (CCTL.IVALL each; first version)
template <bool PAIR>
__device__ __forceinline__ void sk_wait(uint64_t* bar, uint32_t parity) {
    if (PAIR && (g_dbg_cluster_scope)) mbar_wait_cluster(bar, parity); else mbar_wait(bar, parity);
}

template <bool PAIR>
__global__ void __launch_bounds__(320, 1)
ffn_ln_sk_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmO, const FfnArgs a) {
    extern __shared__ unsigned char smem_raw[];
    // PAIR: a 16 KB stage is this CTA's HALF of the B operand of EIGHT K = 16 steps (G1: 64 of the chunk's 128 hidden rows x 128 k;
    // G2: 128 of the 256 output rows x 64 hidden columns, N = 256 MMAs), so one barrier round trip / tcgen05.commit covers 512 clk of
    // tensor work instead of 256 (measured: the per-stage handshake, not the tensor pipe, paced the single-CTA kernel)
    constexpr int NS = FF_NS;
    constexpr int WST = FF_STAGE;
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* xs = smem;                                // [2][4 k-blocks][128 rows x 128 B]
    unsigned char* ring = xs + 2 * FfnSmem::XS;
    float* b1_s = reinterpret_cast<float*>(ring + FfnSmem::RING);
    float* b2_s = b1_s + FF_MAX_HID;
    float* gamma_s = b2_s + FF_D;
    float* beta_s = gamma_s + FF_D;
    float* stat_s = beta_s + FF_D;                           // [128 rows][2 halves][2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(stat_s + FF_BM * 4);
    uint64_t* x_full = bars;             // [2]
    uint64_t* x_free = bars + 2;         // [2]  8 arrivals (epilogue warps)
    uint64_t* w_full = bars + 4;         // [NS]
    uint64_t* w_empty = w_full + NS;     // [NS]
    uint64_t* hacc_full = w_empty + NS;  // [2]
    uint64_t* h_full = hacc_full + 2;    // [2] 8 arrivals
    uint64_t* y_full = h_full + 2;       // [1]
    uint64_t* y_free = y_full + 1;       // [1] 8 arrivals
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(y_free + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_m = (a.M + FF_BM - 1) / FF_BM;
    const int NJ = a.HID / FF_HC;
    // PAIR: the work item is a PAIR tile (256 rows: row tile 2 pt + rank for the CTA of cluster rank `rank`), split over the pairs
    constexpr int NC = PAIR ? 2 : 1;
    const uint32_t rank = PAIR ? cluster_cta_rank() : 0;
    const int num_pt = (num_m + NC - 1) / NC;
    const long long U = (long long)num_pt * NJ;
    const int cta = (int)blockIdx.x, G = (int)gridDim.x / NC, grp = cta / NC;
    const int u0 = (int)(U * grp / G), u1 = (int)(U * (grp + 1) / G);
    const int nu = u1 - u0;                                  // >= NJ (host: num_pt >= G)
    const int mt0 = u0 / NJ, j0 = u0 - mt0 * NJ;             // mt counts pair tiles when PAIR
    const bool leader = rank == 0;
    unsigned int* const dbgbuf = (cta == 0) ? a.dbgbuf : nullptr;
#define FF_DBG(role, unit, slot) do { if (dbgbuf && (unit) < 64) dbgbuf[(((role) * 64 + (unit)) * 16 + (slot))] = (unsigned int)clock(); } while (0)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW1);
        tma_prefetch_desc(&tmW2);
        tma_prefetch_desc(&tmO);
        for (int b = 0; b < 2; ++b) {
            mbar_init(&x_full[b], 1);
            mbar_init(&x_free[b], 8);
            mbar_init(&hacc_full[b], 1);
            mbar_init(&h_full[b], 8 * NC);                   // PAIR: the peer's epilogue warps arrive on the leader's barrier
        }
        for (int s = 0; s < NS; ++s) {
            mbar_init(&w_full[s], 1);
            mbar_init(&w_empty[s], 1);
        }
        mbar_init(y_full, 1);
        mbar_init(y_free, 8 * NC);
        fence_barrier_init();
    }
    if (warp == 1) { if (PAIR) tmem_alloc_pair<512>(tmem_ptr); else tmem_alloc<512>(tmem_ptr); }
    for (int i = threadIdx.x; i < a.HID; i += 320) b1_s[i] = __ldg(a.b1 + i);
    for (int i = threadIdx.x; i < FF_D; i += 320) {
        b2_s[i] = __ldg(a.b2 + i);
        gamma_s[i] = __ldg(a.gamma + i);
        beta_s[i] = __ldg(a.beta + i);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();                    // both CTAs' barriers are initialised before anything arrives at them remotely
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // shared::cluster addresses of the LEADER's barriers (PAIR: TMA completions and epilogue arrivals of both CTAs go there)
    const uint32_t ld_x_full = mapa_rank(smem_u32(x_full), 0), ld_w_full = mapa_rank(smem_u32(w_full), 0);
    const uint32_t ld_h_full = mapa_rank(smem_u32(h_full), 0), ld_y_free = mapa_rank(smem_u32(y_free), 0);
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t tm_y = tmem_base;                 // columns [0, 256)
    const uint32_t tm_h = tmem_base + 256;           // 2 x 128 columns

    if (warp == 0) {
        // ===== TMA producer: W1 chunk of unit v, then W2 chunk of unit v - 1 -- exactly the order the MMA warp consumes them
        if (elect_one()) {
            uint32_t it = 0;
            int t = 0;                                           // item (= row tile visited) index of unit v
            auto load_x = [&](int mt, uint32_t tt) {
                const uint32_t xb = tt & 1;
                sk_wait<PAIR>(&x_free[xb], ((tt >> 1) & 1) ^ 1);
                if (leader) mbar_expect_tx(&x_full[xb], NC * FfnSmem::XS);
                for (int kb = 0; kb < 4; ++kb) {
                    if (PAIR) tma_load_2d_pair(xs + xb * FfnSmem::XS + kb * FF_STAGE, &tmX, ld_x_full + xb * 8, kb * 64, (mt * NC + (int)rank) * FF_BM);
                    else tma_load_2d(xs + xb * FfnSmem::XS + kb * FF_STAGE, &tmX, &x_full[xb], kb * 64, mt * FF_BM);
                }
            };
            load_x(mt0, 0);
            int mt = mt0, j = j0, pj = 0, item_v0 = 0, item_n = (NJ - j0 < nu) ? NJ - j0 : nu;
            for (int v = 0; v <= nu; ++v) {
                FF_DBG(0, v, 0);
                if (v < nu) {
                    if (v > 0 && j == 0) {
                        ++t;
                        item_v0 = v;
                        item_n = (NJ < nu - v) ? NJ : nu - v;
                    }
                    for (int kb = 0; kb < (PAIR ? 2 : 4); ++kb, ++it) {  // W1 rows j*128.., k columns kb*64.. (PAIR: 64 rows x 2 k-blocks)
                        const int s = it % NS;
                        sk_wait<PAIR>(&w_empty[s], ((it / NS) & 1) ^ 1);
                        FF_DBG(0, v, 1 + kb);
                        if (a.dbg & 128) { if (leader) mbar_arrive(&w_full[s]); continue; }      // probe: no weight stream
                        if (leader) mbar_expect_tx(&w_full[s], NC * WST);
                        if (PAIR) {
                            tma_load_2d_pair(ring + s * WST, &tmW1, ld_w_full + s * 8, (2 * kb) * 64, j * FF_HC + (int)rank * 64);
                            tma_load_2d_pair(ring + s * WST + WST / 2, &tmW1, ld_w_full + s * 8, (2 * kb + 1) * 64, j * FF_HC + (int)rank * 64);
                        } else tma_load_2d(ring + s * WST, &tmW1, &w_full[s], kb * 64, j * FF_HC);
                    }
                }
                if (v >= 1) {
                    for (int q = 0; q < (PAIR ? 2 : 4); ++q, ++it) {    // W2 output rows (q&1)*128.., hidden columns of chunk pj
                        const int s = it % NS;                          // (PAIR: output rows rank*128.., hidden columns q*64..)
                        sk_wait<PAIR>(&w_empty[s], ((it / NS) & 1) ^ 1);
                        FF_DBG(0, v, 5 + q);
                        if (a.dbg & 128) { if (leader) mbar_arrive(&w_full[s]); continue; }
                        if (leader) mbar_expect_tx(&w_full[s], NC * WST);
                        if (PAIR) tma_load_2d_pair(ring + s * WST, &tmW2, ld_w_full + s * 8, pj * FF_HC + q * 64, (int)rank * 128);
                        else tma_load_2d(ring + s * WST, &tmW2, &w_full[s], pj * FF_HC + (q >> 1) * 64, (q & 1) * 128);
                    }
                }
                if (v < nu) {
                    // the X tile of the NEXT item, half way through this one (its buffer was released by the item before; everything
                    // that item's last G2 needs has been requested above, so this wait cannot starve it)
                    // (the first item: at once -- both buffers start free)
                    const int at = (t == 0) ? 0 : ((item_n - 1 < NJ / 2) ? item_n - 1 : NJ / 2);
                    FF_DBG(0, v, 9);
                    if (v - item_v0 == at && item_v0 + item_n < nu) load_x(mt + 1, (uint32_t)(t + 1));
                    FF_DBG(0, v, 10);
                    pj = j;
                    if (++j == NJ) { j = 0; ++mt; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: G1(v) one unit ahead of G2(v - 1), across tile boundaries (PAIR: the leader CTA issues for both, M = 256)
        constexpr uint32_t IDESC = (1u << 4) | OP16_IDESC_AB | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)((NC * FF_BM) >> 4) << 24);
        if (leader) {
        uint32_t it = 0;
        int t = -1, j = j0;
        uint32_t xb = 0;
        bool first_w = false, last_w = false;
        int tw = 0;
        for (int v = 0; v <= nu; ++v) {
            bool first_v = false, last_v = false;
            FF_DBG(1, v, 0);
            if (v < nu) {
                first_v = (v == 0) || (j == 0);
                last_v = (v == nu - 1) || (j == NJ - 1);
                if (first_v) {
                    ++t;
                    xb = (uint32_t)t & 1;
                    sk_wait<PAIR>(&x_full[xb], ((uint32_t)t >> 1) & 1);
                    tcgen05_fence_after();
                }
                FF_DBG(1, v, 1);
                const uint32_t b = (uint32_t)v & 1;
                if (PAIR) {
                    constexpr uint32_t IDESC2 = (1u << 4) | OP16_IDESC_AB | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
                    for (int st = 0; st < 2; ++st, ++it) {
                        const int s = it % NS;
                        sk_wait<PAIR>(&w_full[s], (it / NS) & 1);
                        tcgen05_fence_after();
                        FF_DBG(1, v, 2 + 2 * st);
                        if (elect_one()) {
#pragma unroll
                            for (int kb = 0; kb < 2; ++kb) {
                                const uint64_t da = make_sw128_kmajor_desc(smem_u32(xs + xb * FfnSmem::XS + (2 * st + kb) * FF_STAGE));
                                const uint64_t db = make_sw128_kmajor_desc(smem_u32(ring + s * WST + kb * (WST / 2)));
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    if (a.dbg & 512) break;                                      // probe: no G1 MMAs
                                    umma_bf16_pair(tm_h + b * FF_HC, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDESC2, (st | kb | k) != 0);
                                }
                            }
                            umma_commit_pair(&w_empty[s], 3);
                            if (st == 1) umma_commit_pair(&hacc_full[b], 3);
                        }
                        __syncwarp();
                        FF_DBG(1, v, 3 + 2 * st);
                    }
                } else
                for (int kb = 0; kb < 4; ++kb, ++it) {
                    const int s = it % NS;
                    sk_wait<PAIR>(&w_full[s], (it / NS) & 1);
                    tcgen05_fence_after();
                    if (elect_one()) {
                        const uint64_t da = make_sw128_kmajor_desc(smem_u32(xs + xb * FfnSmem::XS + kb * FF_STAGE));
                        const uint64_t db = make_sw128_kmajor_desc(smem_u32(ring + s * WST));
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (a.dbg & 512) break;                                              // probe: no G1 MMAs
                            umma_bf16(tm_h + b * FF_HC, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDESC, (kb | k) != 0);
                        }
                        umma_commit(&w_empty[s]);
                        if (kb == 3) umma_commit(&hacc_full[b]);
                    }
                    __syncwarp();
                }
                if (++j == NJ) j = 0;
            }
            if (v >= 1) {
                const uint32_t w = (uint32_t)(v - 1), b = w & 1;
                sk_wait<PAIR>(&h_full[b], (w >> 1) & 1);             // 16-bit hidden chunk of unit w is in TMEM (of both CTAs)
                FF_DBG(1, v, 6);
                if (first_w) sk_wait<PAIR>(y_free, ((uint32_t)tw & 1) ^ 1);  // previous item's output accumulator read out
                tcgen05_fence_after();
                FF_DBG(1, v, 7);
                if (PAIR) {
                    constexpr uint32_t IDESC3 = (1u << 4) | OP16_IDESC_AB | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
                    for (int kb2 = 0; kb2 < 2; ++kb2, ++it) {
                        const int s = it % NS;
                        sk_wait<PAIR>(&w_full[s], (it / NS) & 1);
                        tcgen05_fence_after();
                        FF_DBG(1, v, 8 + 2 * kb2);
                        if (elect_one()) {
                            const uint64_t db = make_sw128_kmajor_desc(smem_u32(ring + s * WST));
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (a.dbg & 1024) break;                                         // probe: no G2 MMAs
                                umma_bf16_ts_pair(tm_y, tm_h + b * FF_HC + kb2 * 32 + k * 8, db + (uint64_t)(2 * k), IDESC3,
                                                  (!first_w) || kb2 > 0 || k > 0);
                            }
                            umma_commit_pair(&w_empty[s], 3);
                            if (kb2 == 1 && last_w) umma_commit_pair(y_full, 3);
                        }
                        __syncwarp();
                        FF_DBG(1, v, 9 + 2 * kb2);
                    }
                } else
                for (int q = 0; q < 4; ++q, ++it) {
                    const int s = it % NS;
                    const int kb2 = q >> 1, half = q & 1;
                    sk_wait<PAIR>(&w_full[s], (it / NS) & 1);
                    tcgen05_fence_after();
                    if (elect_one()) {
                        const uint64_t db = make_sw128_kmajor_desc(smem_u32(ring + s * WST));
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (a.dbg & 1024) break;                                             // probe: no G2 MMAs
                            umma_bf16_ts(tm_y + half * 128, tm_h + b * FF_HC + kb2 * 32 + k * 8, db + (uint64_t)(2 * k), IDESC,
                                         (!first_w) || kb2 > 0 || k > 0);
                        }
                        umma_commit(&w_empty[s]);
                        if (q == 3 && last_w) umma_commit(y_full);
                    }
                    __syncwarp();
                }
            }
            first_w = first_v; last_w = last_v; tw = t;
        }
        }
    } else {
        // ===== epilogue warps: TMEM lane quarter qd (32 rows), column half hsel
        const int qd = warp & 3;
        const int hsel = (warp - 2) >> 2;
        const int row = qd * 32 + lane;
        const uint32_t swz = (uint32_t)(lane & 7);
        const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
        // workspace slot s: [64 column groups of 4][128 rows] float4 -- a warp's 32 rows are consecutive 16-byte words
        int mt = mt0, js = j0;
        uint32_t t = 0;
        // The LayerNorm of a finished item is split: pass 1 (TMEM + b2 + X -> 16-bit pre-norm rows written IN PLACE over the X tile,
        // row statistics) runs at once and releases the output accumulator; pass 2 (normalise in place, TMA store, X buffer release)
        // is cut into 4 column chunks that run in the idle time after the next item's first E1 steps -- measured (tools/
        // ffn_timeline.py): E1 takes ~1,200 of the ~2,300 clk of a unit, while the un-split LayerNorm stalled the MMA warp for
        // ~12,000 clk per row tile
        int pend = 0;                                        // pass-2 chunks still to do (4 .. 1), 0: none
        bool p_store = false;                                // stores issued, their shared-memory reads not yet awaited
        unsigned char* p_xt = xs;
        int p_row0 = 0;
        uint32_t p_xb = 0;
        float p_mean = 0.f, p_rstd = 0.f;
        auto pass2 = [&]() {
            if (pend > 0) {
                const int c = 4 - pend, cb = c >> 1, k0 = (c & 1) * 4;
                unsigned char* xrow = p_xt + (hsel * 2 + cb) * FF_STAGE + row * 128;
                const float* gp = gamma_s + hsel * 128 + cb * 64 + k0 * 8;
                const float* bp = beta_s + hsel * 128 + cb * 64 + k0 * 8;
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    uint4* px = reinterpret_cast<uint4*>(xrow + (((k0 + kk) ^ swz) * 16));
                    const uint4 r4 = *px;
                    const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
                    const float4 g0 = *reinterpret_cast<const float4*>(gp + kk * 8), g1 = *reinterpret_cast<const float4*>(gp + kk * 8 + 4);
                    const float4 e0 = *reinterpret_cast<const float4*>(bp + kk * 8), e1 = *reinterpret_cast<const float4*>(bp + kk * 8 + 4);
                    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                    const float ee[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
                    uint32_t o[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        o[i] = ff_pack_bf16x2((op16_lo_f32(rw[i]) - p_mean) * p_rstd * gg[2 * i] + ee[2 * i],
                                              (op16_hi_f32(rw[i]) - p_mean) * p_rstd * gg[2 * i + 1] + ee[2 * i + 1]);
                    *px = make_uint4(o[0], o[1], o[2], o[3]);
                }
                if (--pend == 0) {
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        for (int cb2 = 0; cb2 < 2; ++cb2)
                            tma_store_2d(&tmO, p_xt + (hsel * 2 + cb2) * FF_STAGE + (qd * 32) * 128, (hsel * 2 + cb2) * 64, p_row0 + qd * 32);
                        tma_store_commit();
                    }
                    p_store = true;
                }
            } else if (p_store) {
                if (lane == 0) {
                    tma_store_wait_read<0>();                                // the X buffer may now be refilled
                    mbar_arrive(&x_free[p_xb]);
                }
                __syncwarp();
                p_store = false;
            }
        };
        for (int v = 0; v < nu; ++t, ++mt, js = 0) {
            const int n = (NJ - js < nu - v) ? NJ - js : nu - v;
            const bool tail = (js > 0) && !(a.dbg & 16384), head = (js == 0) && (n < NJ) && !(a.dbg & 16384);   // probe 16384: no partial exchange
            // ---- E1: +b1, ReLU, 16-bit, back into TMEM in place
            for (int i = 0; i < n; ++i, ++v) {
                const uint32_t b = (uint32_t)v & 1, u = (uint32_t)v >> 1;
                if (threadIdx.x == 64) FF_DBG(2, v, 0);
                sk_wait<PAIR>(&hacc_full[b], u & 1);
                tcgen05_fence_after();
                if (threadIdx.x == 64) FF_DBG(2, v, 1);
                if (a.dbg & 256) {                                                               // probe: no E1 work
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) { if (PAIR) mbar_arrive_cluster(ld_h_full + b * 8); else mbar_arrive(&h_full[b]); }
                    continue;
                }
                uint32_t acc[64];
                tmem_ld64(tm_h + b * FF_HC + lane_addr + (uint32_t)(hsel * 64), acc);
                if (threadIdx.x == 64) FF_DBG(2, v, 2);
                const float* bp = b1_s + (js + i) * FF_HC + hsel * 64;
                uint32_t pk[32];
#pragma unroll
                for (int c = 0; c < 64; c += 4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(bp + c);
                    const float v0 = fmaxf(__uint_as_float(acc[c]) + b4.x, 0.f), v1 = fmaxf(__uint_as_float(acc[c + 1]) + b4.y, 0.f);
                    const float v2 = fmaxf(__uint_as_float(acc[c + 2]) + b4.z, 0.f), v3 = fmaxf(__uint_as_float(acc[c + 3]) + b4.w, 0.f);
                    pk[c / 2] = ff_pack_bf16x2(v0, v1);
                    pk[c / 2 + 1] = ff_pack_bf16x2(v2, v3);
                }
                asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");
                if (threadIdx.x == 64) FF_DBG(2, v, 3);
                tmem_st32(tm_h + b * FF_HC + lane_addr + (uint32_t)(hsel * 32), pk);
                tcgen05_fence_before();
                __syncwarp();
                if (threadIdx.x == 64) FF_DBG(2, v, 4);
                if (lane == 0) { if (PAIR) mbar_arrive_cluster(ld_h_full + b * 8); else mbar_arrive(&h_full[b]); }
                if (threadIdx.x == 64) FF_DBG(2, v, 5);
                pass2();
            }
            while (pend > 0 || p_store) pass2();             // (short item: whatever is left of the previous item's LayerNorm)
            const uint32_t xb = t & 1;
            unsigned char* xt = xs + xb * FfnSmem::XS;
            if (tail) {
                // ---- final (tail part): raw fp32 partial -> workspace slot `cta`, then the ready flag for the left neighbour
                sk_wait<PAIR>(y_full, t & 1);
                tcgen05_fence_after();
                if (lane == 0) mbar_arrive(&x_free[xb]);
                float4* dst = reinterpret_cast<float4*>(a.partial) + (size_t)cta * (FF_BM * FF_D / 4) + row;
#pragma unroll
                for (int cb = 0; cb < 2; ++cb) {
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        uint32_t acc[32];
                        tmem_ld32(tm_y + lane_addr + (uint32_t)(hsel * 128 + cb * 64 + hf * 32), acc);
                        if (cb == 1 && hf == 1) {
                            tcgen05_fence_before();
                            __syncwarp();
                            if (lane == 0) { if (PAIR) mbar_arrive_cluster(ld_y_free); else mbar_arrive(y_free); }
                        }
                        const int cg0 = (hsel * 128 + cb * 64 + hf * 32) / 4;
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            __stcg(dst + (size_t)(cg0 + i) * FF_BM,
                                   make_float4(__uint_as_float(acc[4 * i]), __uint_as_float(acc[4 * i + 1]), __uint_as_float(acc[4 * i + 2]), __uint_as_float(acc[4 * i + 3])));
                    }
                }
                __threadfence();
                asm volatile("bar.sync 5, 256;" ::: "memory");
                if (threadIdx.x == 64) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.flags + cta), "r"(1u) : "memory");
                continue;
            }
            // (PAIR: this CTA's X tile completed on the leader's barrier; y_full -- every MMA of the pair done -- implies it landed)
            if (threadIdx.x == 64) FF_DBG(2, 56 + t, 0);
            if (!PAIR) sk_wait<PAIR>(&x_full[xb], (t >> 1) & 1);
            sk_wait<PAIR>(y_full, t & 1);
            tcgen05_fence_after();
            if (threadIdx.x == 64) FF_DBG(2, 56 + t, 1);
            const float4* psrc = reinterpret_cast<const float4*>(a.partial) + (size_t)(cta + NC) * (FF_BM * FF_D / 4) + row;
            if (head) {
                // ---- head part: the right neighbour's partial over the remaining chunks of this tile (written long ago)
                if (lane == 0) {
                    uint32_t f;
                    do {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(a.flags + cta + NC) : "memory");
                    } while (f == 0u);
                }
                __syncwarp();
            }
            if (a.dbg & 2048) {                              // probe: no final epilogue work
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) { if (PAIR) mbar_arrive_cluster(ld_y_free); else mbar_arrive(y_free); mbar_arrive(&x_free[xb]); }
                continue;
            }
            float sum = 0.f, sq = 0.f;
#pragma unroll
            for (int cb = 0; cb < 2; ++cb) {
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t acc[32];
                    tmem_ld32(tm_y + lane_addr + (uint32_t)(hsel * 128 + cb * 64 + hf * 32), acc);
                    if (cb == 1 && hf == 1) {                // the output accumulator is in registers: the next item's G2 may start
                        tcgen05_fence_before();
                        __syncwarp();
                        if (lane == 0) { if (PAIR) mbar_arrive_cluster(ld_y_free); else mbar_arrive(y_free); }
                    }
                    if (head) {
                        const int cg0 = (hsel * 128 + cb * 64 + hf * 32) / 4;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 p = __ldcg(psrc + (size_t)(cg0 + i) * FF_BM);
                            acc[4 * i] = __float_as_uint(__uint_as_float(acc[4 * i]) + p.x);
                            acc[4 * i + 1] = __float_as_uint(__uint_as_float(acc[4 * i + 1]) + p.y);
                            acc[4 * i + 2] = __float_as_uint(__uint_as_float(acc[4 * i + 2]) + p.z);
                            acc[4 * i + 3] = __float_as_uint(__uint_as_float(acc[4 * i + 3]) + p.w);
                        }
                    }
                    unsigned char* xrow = xt + (hsel * 2 + cb) * FF_STAGE + row * 128;
                    const float* bp = b2_s + hsel * 128 + cb * 64 + hf * 32;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const int k = hf * 4 + kk;
                        uint4* px = reinterpret_cast<uint4*>(xrow + ((k ^ swz) * 16));
                        const uint4 r4 = *px;
                        const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
                        const float4 q0 = *reinterpret_cast<const float4*>(bp + kk * 8), q1 = *reinterpret_cast<const float4*>(bp + kk * 8 + 4);
                        const float bb[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                        uint32_t o[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float x0 = __uint_as_float(acc[kk * 8 + 2 * i]) + bb[2 * i] + op16_lo_f32(rw[i]);
                            const float x1 = __uint_as_float(acc[kk * 8 + 2 * i + 1]) + bb[2 * i + 1] + op16_hi_f32(rw[i]);
                            const uint32_t pk = ff_pack_bf16x2(x0, x1);
                            o[i] = pk;
                            const float y0 = op16_lo_f32(pk), y1 = op16_hi_f32(pk);
                            sum += y0 + y1;
                            sq = fmaf(y0, y0, sq);
                            sq = fmaf(y1, y1, sq);
                        }
                        *px = make_uint4(o[0], o[1], o[2], o[3]);      // the pre-norm row replaces the residual it was made from
                    }
                }
            }
            if (head) {
                // every warp has consumed the partial: hand the flag back (0) for the next launch on this workspace
                asm volatile("bar.sync 5, 256;" ::: "memory");
                if (threadIdx.x == 64) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(a.flags + cta + NC), "r"(0u) : "memory");
            }
            if (threadIdx.x == 64) FF_DBG(2, 56 + t, 2);
            stat_s[(row * 2 + hsel) * 2] = sum;
            stat_s[(row * 2 + hsel) * 2 + 1] = sq;
            asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");       // the two warps of this lane quarter
            p_mean = (stat_s[row * 4] + stat_s[row * 4 + 2]) * (1.f / 256.f);
            p_rstd = rsqrtf(fmaxf((stat_s[row * 4 + 1] + stat_s[row * 4 + 3]) * (1.f / 256.f) - p_mean * p_mean, 0.f) + a.eps);
            asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");       // statistics consumed before the next item overwrites them
            pend = 4;
            p_xt = xt;
            p_xb = xb;
            p_row0 = (mt * NC + (int)rank) * FF_BM;
            if (threadIdx.x == 64) FF_DBG(2, 56 + t, 3);
        }
        while (pend > 0 || p_store) pass2();                 // the last item's LayerNorm has nothing left to hide under
        if (lane == 0) tma_store_wait<0>();
    }
    tcgen05_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();                    // no CTA leaves (or frees tensor memory) while its peer may still use it
    if (warp == 1) {
        tcgen05_fence_after();
        if (PAIR) tmem_dealloc_pair<512>(tmem_base); else tmem_dealloc<512>(tmem_base);
    }
}

// The rows of the wave-quantisation tail: sum of the hidden-slice partials (fixed order: deterministic) + b2 + X -> rounded to the
// 16-bit type like the main kernel's packed pre-norm row -> LayerNorm -> Y.  One warp per row, 8 columns per lane.
__global__ void __launch_bounds__(256)
ffn_tail_ln_kernel(const float* __restrict__ partial, const int nsplit, const size_t slice_stride, const op16_t* __restrict__ x, const int ldx,
                   const float* __restrict__ b2, const float* __restrict__ gamma, const float* __restrict__ beta, const float eps,
                   op16_t* __restrict__ y, const int ldy, const int rows) {
    pdl_launch_dependents();
    pdl_wait();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int c = (threadIdx.x & 31) * 8;
    float v[8];
    {
        const uint4 xr = *reinterpret_cast<const uint4*>(x + (size_t)row * ldx + c);
        const uint32_t xw[4] = {xr.x, xr.y, xr.z, xr.w};
        const float4 ba = *reinterpret_cast<const float4*>(b2 + c), bb = *reinterpret_cast<const float4*>(b2 + c + 4);
        const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[2 * i] = bv[2 * i] + op16_lo_f32(xw[i]); v[2 * i + 1] = bv[2 * i + 1] + op16_hi_f32(xw[i]); }
    }
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int sidx = 0; sidx < nsplit; ++sidx) {
        const float4* p = reinterpret_cast<const float4*>(partial + (size_t)sidx * slice_stride + (size_t)row * FF_D + c);
        const float4 a = p[0], b = p[1];
        acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w; acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
    }
    float sum = 0.f, sq = 0.f;
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        pk[i] = ff_pack_bf16x2(acc[2 * i] + v[2 * i], acc[2 * i + 1] + v[2 * i + 1]);
        const float y0 = op16_lo_f32(pk[i]), y1 = op16_hi_f32(pk[i]);
        sum += y0 + y1;
        sq = fmaf(y0, y0, sq);
        sq = fmaf(y1, y1, sq);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); sq += __shfl_xor_sync(0xffffffffu, sq, o); }
    const float mean = sum * (1.f / 256.f);
    const float rstd = rsqrtf(fmaxf(sq * (1.f / 256.f) - mean * mean, 0.f) + eps);
    uint32_t o4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        o4[i] = ff_pack_bf16x2((op16_lo_f32(pk[i]) - mean) * rstd * gamma[c + 2 * i] + beta[c + 2 * i],
                               (op16_hi_f32(pk[i]) - mean) * rstd * gamma[c + 2 * i + 1] + beta[c + 2 * i + 1]);
    *reinterpret_cast<uint4*>(y + (size_t)row * ldy + c) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
}

static int ffn_tmap(CUtensorMap* map, const void* base, long long rows, int cols, long long ld, int box_rows) {
    return make_tmap_2d_bf16(map, base, rows, cols, ld, box_rows, 64, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace dtlr

using namespace dtlr;

// Wave quantisation: the block runs one 128-row tile per SM at a time; num_m tiles on `sm` SMs take ceil(num_m / sm)
// rounds (456 tiles at B = 64: 3.08 waves -> 4 rounds).  When the last round is mostly empty the split entry point runs the full
// rounds with the main kernel and the `rem` tail tiles with the PART variant -- each tail tile's hidden dimension cut into `nsplit`
// slices on otherwise idle SMs, partial sums to a workspace -- followed by ffn_tail_ln_kernel.
static void ffn_split_plan(int M, int hidden, int* main_rows, int* rem_tiles, int* nsplit) {
    const int num_m = (M + FF_BM - 1) / FF_BM, sm = sm_count();
    const int full = num_m / sm, rem = num_m % sm;
    *main_rows = M; *rem_tiles = 0; *nsplit = 1;
    if ((g_debug_flags & 262144) || full < 1 || rem == 0 || rem * 2 > sm) return;      // flag 262144: never split (A/B)
    int ns = 1;
    const int nj = hidden / FF_HC;
    while (ns * 2 <= nj && (nj % (ns * 2)) == 0 && rem * ns * 2 <= sm) ns *= 2;
    if (ns < 2) return;
    *main_rows = full * sm * FF_BM; *rem_tiles = rem; *nsplit = ns;
}

// Which plan dtlr_ffn_ln_ws runs for a shape: 0 = the plain persistent kernel (fewer tiles than one pair tile per CTA pair), 1 = full
// rounds + PART tail + ffn_tail_ln_kernel (round-2 first version, kept behind dtlr_debug_flags(536870912) for A/B), 2 = stream-K
// on single CTAs (ffn_ln_sk_kernel<false>; A/B flag 1073741824), 3 = stream-K on CTA pairs (ffn_ln_sk_kernel<true>, the default).
constexpr long long FF_SK_FLAG_BYTES = 1024;
static int ffn_plan(int M, int hidden) {
    if (M <= 0 || hidden <= 0 || (hidden % FF_HC) != 0 || hidden > FF_MAX_HID) return 0;
    if (g_debug_flags & (262144 | 4096)) return 0;                  // flag 262144: never split (A/B); 4096: the CTA-pair variant
    const int num_m = (M + FF_BM - 1) / FF_BM, sm = sm_count();
    if (g_debug_flags & 536870912) {
        int main_rows, rem, ns;
        ffn_split_plan(M, hidden, &main_rows, &rem, &ns);
        return rem ? 1 : 0;
    }
    if (sm * 4 > FF_SK_FLAG_BYTES || hidden < 2 * FF_HC) return 0;          // (one hidden chunk per tile: nothing to stream)
    // stream-K on CTA pairs (cta_group::2, 256-row pair tiles): the default from one pair tile per pair upwards (measured at
    // M = 58368: 92 us against 108-118 us for the single-CTA stream-K kernel); flag 1073741824: single-CTA kernels only (A/B)
    if (!(g_debug_flags & 1073741824) && (sm % 2) == 0 && (num_m + 1) / 2 >= sm / 2) return 3;
    if (num_m <= sm || (num_m % sm) == 0) return 0;
    return 2;
}

extern "C" int dtlr_ffn_plan(int M, int hidden) { return ffn_plan(M, hidden); }

// timeline probe of the stream-K kernels (tools/ffn_timeline.py): a device buffer of 3 * 64 * 16 u32, or NULL (off)
static unsigned int* g_ffn_dbgbuf = nullptr;
extern "C" int dtlr_ffn_debug_buffer(void* buf) { g_ffn_dbgbuf = reinterpret_cast<unsigned int*>(buf); return DTLR_OK; }

extern "C" long long dtlr_ffn_workspace_bytes(int M, int hidden) {
    const int plan = ffn_plan(M, hidden);
    if (plan >= 2) return FF_SK_FLAG_BYTES + (long long)sm_count() * FF_BM * FF_D * 4;
    if (plan == 1) {
        int main_rows, rem, ns;
        ffn_split_plan(M, hidden, &main_rows, &rem, &ns);
        return (long long)ns * rem * FF_BM * FF_D * 4;
    }
    return 0;
}

extern "C" int dtlr_ffn_ln(const void* X, int ldx, const void* W1, int ldw1, const float* b1, const void* W2, int ldw2,
                           const float* b2, const float* gamma, const float* beta, float eps, void* Y, int ldy, int M, int hidden,
                           void* stream);

extern "C" int dtlr_ffn_ln_ws(const void* X, int ldx, const void* W1, int ldw1, const float* b1, const void* W2, int ldw2,
                              const float* b2, const float* gamma, const float* beta, float eps, void* Y, int ldy, int M, int hidden,
                              void* workspace, long long workspace_bytes, void* stream) {
    const int plan = ffn_plan(M, hidden);
    if (!plan || !workspace || workspace_bytes < dtlr_ffn_workspace_bytes(M, hidden))
        return dtlr_ffn_ln(X, ldx, W1, ldw1, b1, W2, ldw2, b2, gamma, beta, eps, Y, ldy, M, hidden, stream);
    int rc;
    if (plan >= 2) {
        const bool pair = plan == 3;
        DTLR_CHECK_ARG(X && W1 && b1 && W2 && b2 && gamma && beta && Y, "ffn_ln: null pointer");
        DTLR_CHECK_ARG(ldx >= FF_D && ldw1 >= FF_D && ldw2 >= hidden && ldy >= FF_D, "ffn_ln: leading dimension too small");
        DTLR_CHECK_ARG((ldx % 8) == 0 && (ldw1 % 8) == 0 && (ldw2 % 8) == 0 && (ldy % 8) == 0 &&
                       ((((uintptr_t)X | (uintptr_t)W1 | (uintptr_t)W2 | (uintptr_t)Y | (uintptr_t)workspace)) & 15) == 0,
                       "ffn_ln: operands need 16-byte aligned rows");
        CUtensorMap tx, tw1, tw2, to;
        if ((rc = ffn_tmap(&tx, X, M, FF_D, ldx, FF_BM))) return rc;
        if ((rc = ffn_tmap(&tw1, W1, hidden, FF_D, ldw1, pair ? 64 : 128))) return rc;
        if ((rc = ffn_tmap(&tw2, W2, FF_D, hidden, ldw2, 128))) return rc;
        if ((rc = ffn_tmap(&to, Y, M, FF_D, ldy, 32))) return rc;
        static bool configured_sk = false;
        if (!configured_sk) {
            DTLR_CHECK_CUDA(cudaFuncSetAttribute(ffn_ln_sk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FfnSmem::TOTAL));
            DTLR_CHECK_CUDA(cudaFuncSetAttribute(ffn_ln_sk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FfnSmem::TOTAL));
            configured_sk = true;
        }
        FfnArgs a{b1, b2, gamma, beta, eps, M, hidden, g_debug_flags, nullptr, nullptr, nullptr, nullptr, 0, 1,
                  reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(workspace) + FF_SK_FLAG_BYTES), reinterpret_cast<int*>(workspace), g_ffn_dbgbuf};
        if (pair)
            DTLR_CHECK_CUDA(launch_pdl_cluster(ffn_ln_sk_kernel<true>, dim3(sm_count()), dim3(320), FfnSmem::TOTAL, (cudaStream_t)stream, 2u, tx, tw1, tw2, to, a));
        else
            DTLR_CHECK_CUDA(launch_pdl(ffn_ln_sk_kernel<false>, dim3(sm_count()), dim3(320), FfnSmem::TOTAL, (cudaStream_t)stream, tx, tw1, tw2, to, a));
        return DTLR_OK;
    }
    int main_rows = M, rem = 0, ns = 1;
    ffn_split_plan(M, hidden, &main_rows, &rem, &ns);
    rc = dtlr_ffn_ln(X, ldx, W1, ldw1, b1, W2, ldw2, b2, gamma, beta, eps, Y, ldy, main_rows, hidden, stream);
    if (rc) return rc;
    const op16_t* xt = reinterpret_cast<const op16_t*>(X) + (size_t)main_rows * ldx;
    op16_t* yt = reinterpret_cast<op16_t*>(Y) + (size_t)main_rows * ldy;
    const int mt = M - main_rows;
    CUtensorMap tx, tw1, tw2;
    if ((rc = ffn_tmap(&tx, xt, mt, FF_D, ldx, FF_BM))) return rc;
    if ((rc = ffn_tmap(&tw1, W1, hidden, FF_D, ldw1, 128))) return rc;
    if ((rc = ffn_tmap(&tw2, W2, FF_D, hidden, ldw2, 128))) return rc;
    static bool configured = false;
    if (!configured) {
        DTLR_CHECK_CUDA(cudaFuncSetAttribute(ffn_ln_tcgen05_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FfnSmem::TOTAL));
        configured = true;
    }
    const FfnArgs a{b1, b2, gamma, beta, eps, mt, hidden / ns, g_debug_flags, nullptr, nullptr, nullptr, nullptr, 0, ns, (float*)workspace};
    DTLR_CHECK_CUDA(launch_pdl(ffn_ln_tcgen05_kernel<false, 2>, dim3(rem * ns), dim3(320), FfnSmem::TOTAL, (cudaStream_t)stream, tx, tw1, tw2, tx, a));
    DTLR_CHECK_CUDA(launch_pdl(ffn_tail_ln_kernel, dim3((mt + 7) / 8), dim3(256), 0, (cudaStream_t)stream, (const float*)workspace, ns,
                               (size_t)rem * FF_BM * FF_D, xt, ldx, b2, gamma, beta, eps, yt, ldy, mt));
    return DTLR_OK;
}

extern "C" int dtlr_ffn_ln(const void* X, int ldx, const void* W1, int ldw1, const float* b1, const void* W2, int ldw2,
                           const float* b2, const float* gamma, const float* beta, float eps, void* Y, int ldy, int M, int hidden,
                           void* stream) {
    DTLR_CHECK_ARG(M >= 0 && hidden > 0, "ffn_ln: bad sizes");
    if (M == 0) return DTLR_OK;
    DTLR_CHECK_ARG(X && W1 && b1 && W2 && b2 && gamma && beta && Y, "ffn_ln: null pointer");
    DTLR_CHECK_ARG((hidden % FF_HC) == 0 && hidden <= FF_MAX_HID, "ffn_ln: hidden width %d must be a multiple of 128 and <= %d", hidden, FF_MAX_HID);
    DTLR_CHECK_ARG(ldx >= FF_D && ldw1 >= FF_D && ldw2 >= hidden && ldy >= FF_D, "ffn_ln: leading dimension too small");
    DTLR_CHECK_ARG((ldx % 8) == 0 && (ldw1 % 8) == 0 && (ldw2 % 8) == 0 && (ldy % 8) == 0 &&
                   ((((uintptr_t)X | (uintptr_t)W1 | (uintptr_t)W2 | (uintptr_t)Y)) & 15) == 0, "ffn_ln: operands need 16-byte aligned rows");
    CUtensorMap tx, tw1, tw2, to;
    int rc;
    if ((rc = ffn_tmap(&tx, X, M, FF_D, ldx, FF_BM))) return rc;
    if ((rc = ffn_tmap(&tw1, W1, hidden, FF_D, ldw1, 128))) return rc;
    if ((rc = ffn_tmap(&tw2, W2, FF_D, hidden, ldw2, 128))) return rc;
    if ((rc = ffn_tmap(&to, Y, M, FF_D, ldy, 32))) return rc;
    static bool configured = false;
    if (!configured) {
        DTLR_CHECK_CUDA(cudaFuncSetAttribute(ffn_ln_tcgen05_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FfnSmem::TOTAL));
        DTLR_CHECK_CUDA(cudaFuncSetAttribute(ffn_ln_tcgen05_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FfnSmem::TOTAL));
        configured = true;
    }
    const int num_m = (M + FF_BM - 1) / FF_BM;
    const FfnArgs a{b1, b2, gamma, beta, eps, M, hidden, g_debug_flags, nullptr, nullptr, nullptr, nullptr, 0, 1, nullptr};
    // CTA pairs with multicast weights: measured identical (129.0 vs 130.1 us): halving the L2 reads does not help because the
    // limit of the weight stream is the ~51 B/clk at which one SM's shared memory is filled (loads alone: 63 us with or without
    // multicast, with 5 or 9 ring stages).  Kept (dtlr_debug_flags(4096)) as the validated base of the cta_group::2 version, which
    // halves the bytes each SM must take in.
    if ((g_debug_flags & 4096) && num_m >= 2) {
        int grid = num_m < sm_count() ? num_m : sm_count();
        grid &= ~1;
        DTLR_CHECK_CUDA(launch_pdl_cluster(ffn_ln_tcgen05_kernel<true, 0>, dim3(grid), dim3(320), FfnSmem::TOTAL, (cudaStream_t)stream, 2u, tx, tw1, tw2, to, a));
        return DTLR_OK;
    }
    const int grid = num_m < sm_count() ? num_m : sm_count();
    DTLR_CHECK_CUDA(launch_pdl(ffn_ln_tcgen05_kernel<false, 0>, dim3(grid), dim3(320), FfnSmem::TOTAL, (cudaStream_t)stream, tx, tw1, tw2, to, a));
    return DTLR_OK;
}

// 3-layer MLP head with a 4-wide last layer (reference models/dino/utils.py:110-122: bbox_embed / enc_out_bbox_embed, 256 -> 256 ->
// 256 -> 4) fused with the box refinement that follows it (deformable_transformer.py:734-738, dino.py:343-345) -- the HEAD variant of
// the FFN kernel above.  X [M,256] 16-bit, W1 / W2 [256,256] 16-bit, b1 / b2 fp32, W3 [4,256] fp32, b3 [4] fp32, ref [M,4] fp32 or
// NULL (NULL: out4 = the raw deltas), out4 [M,4] fp32.
extern "C" int dtlr_mlp_head(const void* X, int ldx, const void* W1, int ldw1, const float* b1, const void* W2, int ldw2,
                             const float* b2, const float* W3, const float* b3, const float* ref, float* out4, int M, void* stream) {
    DTLR_CHECK_ARG(M >= 0, "mlp_head: bad sizes");
    if (M == 0) return DTLR_OK;
    DTLR_CHECK_ARG(X && W1 && b1 && W2 && b2 && W3 && b3 && out4, "mlp_head: null pointer");
    DTLR_CHECK_ARG(ldx >= FF_D && ldw1 >= FF_D && ldw2 >= FF_D, "mlp_head: leading dimension too small");
    DTLR_CHECK_ARG((ldx % 8) == 0 && (ldw1 % 8) == 0 && (ldw2 % 8) == 0 && ((((uintptr_t)X | (uintptr_t)W1 | (uintptr_t)W2)) & 15) == 0 &&
                   ((((uintptr_t)out4 | (uintptr_t)ref)) & 15) == 0, "mlp_head: operands need 16-byte aligned rows");
    CUtensorMap tx, tw1, tw2;
    int rc;
    if ((rc = ffn_tmap(&tx, X, M, FF_D, ldx, FF_BM))) return rc;
    if ((rc = ffn_tmap(&tw1, W1, FF_D, FF_D, ldw1, 128))) return rc;
    if ((rc = ffn_tmap(&tw2, W2, FF_D, FF_D, ldw2, 128))) return rc;
    static bool configured = false;
    if (!configured) {
        DTLR_CHECK_CUDA(cudaFuncSetAttribute(ffn_ln_tcgen05_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FfnSmem::TOTAL));
        configured = true;
    }
    const int num_m = (M + FF_BM - 1) / FF_BM;
    const FfnArgs a{b1, b2, nullptr, nullptr, 0.f, M, FF_D, g_debug_flags, W3, b3, ref, out4, 0, 1, nullptr};
    const int grid = num_m < sm_count() ? num_m : sm_count();
    DTLR_CHECK_CUDA(launch_pdl(ffn_ln_tcgen05_kernel<false, 1>, dim3(grid), dim3(320), FfnSmem::TOTAL, (cudaStream_t)stream, tx, tw1, tw2, tx, a));
    return DTLR_OK;
}
