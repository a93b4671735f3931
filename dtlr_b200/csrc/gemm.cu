// dtlr_b200 -- dense contractions of the DINO hot path for sm_100a.
//
//   C[M,N] = act( A[M,K] . W[N,K]^T + bias[N] ) (+ residual[M,N])
//
// which is every nn.Linear of the reference transformer (models/dino/deformable_transformer.py:787-790,852-855,
// ops/modules/ms_deform_attn.py:55-58, utils.py:110-122), nn.MultiheadAttention's in/out projections, and -- with the
// activations in NHWC -- every 1x1 convolution of the ResNet-50 trunk and of input_proj.
//
//  * bf16 operands (throughput mode): tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM), operands staged by TMA
//    into 128B-swizzled shared memory, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one elected
//    thread) + TMEM allocator, warps 2-5 = epilogue (tcgen05.ld -> bias/ReLU/residual -> global).  One 128 x BN
//    output tile per CTA, STAGES-deep mbarrier ring; two CTAs are co-resident per SM so one tile's epilogue overlaps
//    the other's main loop.
//  * fp32 operands (parity mode): a plain SIMT kernel with exact fp32 FMA accumulation -- tensor cores have no fp32
//    input type and TF32 (10-bit mantissa) cannot hold the 1e-3 end-to-end tolerance / top-k ranking of the reference.
#include "tc_common.cuh"

namespace dtlr {

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    op16x2_t t = op16_pack2(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
template <typename OutT> __device__ __forceinline__ void unpack_chunk(const uint4& d, float* f);
template <> __device__ __forceinline__ void unpack_chunk<float>(const uint4& d, float* f) {
    f[0] = __uint_as_float(d.x); f[1] = __uint_as_float(d.y); f[2] = __uint_as_float(d.z); f[3] = __uint_as_float(d.w);
}
template <> __device__ __forceinline__ void unpack_chunk<op16_t>(const uint4& d, float* f) {
    const uint32_t w[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = op16_lo_f32(w[i]); f[2 * i + 1] = op16_hi_f32(w[i]); }
}
template <typename OutT> __device__ __forceinline__ uint4 pack_chunk(const float* f);
template <> __device__ __forceinline__ uint4 pack_chunk<float>(const float* f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
}
template <> __device__ __forceinline__ uint4 pack_chunk<op16_t>(const float* f) {
    return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// ---------------------------------------------------------------------------------------------- tcgen05 GEMM
// Implicit-GEMM geometry of a stride-1 k x k convolution on NHWC activations: the A operand of k-block kb is the
// (kh,kw) tap / 64-channel slice of the input, fetched by TMA straight from the [B,H,W,C] tensor (4-D map, zero fill
// outside the image = the conv padding); a 128-pixel M tile is nseg runs of seg_w consecutive output pixels of one row.
struct ConvGeo {
    int enabled;
    int Ho, Wo, KW, pad, cblocks;     // cblocks = C / 64
    int seg_w, nseg;                  // seg_w * nseg == 128, Wo % seg_w == 0
    int stride;                       // 1, or 2: the A tensor map traverses W and H with this element stride (TMA loads every
                                      // stride-th pixel of a box that spans seg_w * stride input pixels)
};

// optional LayerNorm fused into the epilogue of a full-row (N == 256 == BN) bf16 GEMM:
//   y = LN(gemm + bias + residual) * gamma + beta ;  y2 = y + add2 (optional second output)
struct LnArgs {
    const float* gamma;      // null = no LayerNorm
    const float* beta;
    const void* add2;        // [M, ld2] bf16 or null
    void* out2;              // [M, ld2] bf16
    int ld2;
    float eps;
};

struct GemmEpi {
    const float* bias;       // [N] fp32 or null
    const void* residual;    // [M, ldr] same dtype as C, or null
    void* C;                 // [M, ldc]
    int ldr, ldc;
    int M, N, K;
    int relu;                // 0 none, 1 ReLU before the residual add, 2 ReLU after it (ResNet bottleneck)
    LnArgs ln;
    ConvGeo conv;            // conv.enabled: implicit-GEMM convolution (A fetched through the 4-D map)
    int dbg;                 // tuning only (dtlr_debug_flags): 1 skip global stores, 2 skip MMA issue, 4 skip step-1 staging
    unsigned int* dbgbuf;    // timeline probe (dtlr_gemm_debug_buffer): CTA 0 of the weight-stationary kernel records clock() stamps
    int split3 = 0;          // tile kernel, even STAGES, plain (non-conv) A: the operands are split-precision matrices A = [hi | hi | lo],
                             // W = [hi | lo | hi] over K = 3 Kl columns (Kl % 64 == 0).  Instead of walking 3 Kl columns (6 tile loads per
                             // logical k-block) the producer loads A_hi, W_hi and A_lo, W_lo once (two pipeline stages) and the issuer
                             // runs hi.hi, hi.lo, lo.hi from them: the same three products with 2/3 of the L2 -> shared-memory traffic
    int nfast = 0;           // tile kernel: column tiles vary fastest over the persistent CTAs (A tiles shared through L2) instead of row tiles
    int split_out = 0;       // fp32-output tile kernel only: write the result as the 16-bit split operand [hi | hi | lo] ([M, 3N], ldc in
                             // 16-bit elements) that dtlr_split_cast would make of it -- the next split product reads it directly
};

static unsigned int* g_gemm_dbgbuf = nullptr;
extern "C" int dtlr_gemm_debug_buffer(void* buf) { g_gemm_dbgbuf = reinterpret_cast<unsigned int*>(buf); return DTLR_OK; }

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;   // 64 bf16 = 128 B = one swizzle atom row


template <int BN, int STAGES, typename OutT>
struct GemmSmem {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_BYTES = BN * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int ROW_BYTES = BN * (int)sizeof(OutT);          // staging row pitch (16-byte chunks XOR-swizzled)
    static constexpr int STAGING_BYTES = GEMM_BM * ROW_BYTES;
    static constexpr int BIAS_BYTES = 4 * BN * 4;                     // one fp32 bias row per epilogue warp
    static constexpr int LN_BYTES = (BN == 256 && sizeof(OutT) == 2) ? (2 * GEMM_BM * 4 * 4 + 2 * BN * 4) : 0;   // row stats x2 + gamma/beta
    static constexpr int TOTAL = STAGES * STAGE_BYTES + STAGING_BYTES + BIAS_BYTES + LN_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void ln_load_blk(uint4 (&dst)[8], const op16_t* src, const int ld, const int cb, const int m0,
                                            const int qd, const int lane, const int M) {
#pragma unroll
    for (int itx = 0; itx < 8; ++itx) {
        const int r = itx * 4 + (lane >> 3), ch = lane & 7;
        const int grow = m0 + qd * 32 + r;
        dst[itx] = make_uint4(0, 0, 0, 0);
        if (src && grow < M) dst[itx] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)grow * ld + cb * 64 + ch * 8));
    }
}
// coalesced registers -> swizzled rows of the warp's staging block (after every earlier TMA store has finished reading it)
__device__ __forceinline__ void ln_stage_blk(const uint4 (&src)[8], unsigned char* buf0, const int lane) {
    if (lane == 0) tma_store_wait_read<0>();
    __syncwarp();
#pragma unroll
    for (int itx = 0; itx < 8; ++itx) {
        const int r = itx * 4 + (lane >> 3), ch = lane & 7;
        *reinterpret_cast<uint4*>(buf0 + r * 128 + ((ch ^ (r & 7)) * 16)) = src[itx];
    }
    __syncwarp();
}
__device__ __forceinline__ uint32_t ln_norm_pair(const uint32_t pk, const float mean, const float rstd, const float* g, const float* b) {
    const float y0 = (op16_lo_f32(pk) - mean) * rstd * g[0] + b[0];
    const float y1 = (op16_hi_f32(pk) - mean) * rstd * g[1] + b[1];
    return pack_bf16x2(y0, y1);
}

// Linear -> (+residual) -> LayerNorm(256) [-> second output y + add2] epilogue of one 128 x 256 accumulator tile, shared by the
// weight-stationary kernel (K <= 256) and the tile kernel (any K).  8 epilogue warps; warp (qd, hsel) holds the column blocks
// {hsel, hsel + 2} x 64 of its 32 rows (thread = row), the partner warp of the lane quarter the other half; the bf16-rounded
// pre-norm row stays packed in registers between the statistics pass and the normalisation pass.  Residual / add2 blocks are
// fetched coalesced (8 x 16 bytes per lane) and turned to thread-per-row order through the warp's 4 KB 128B-swizzled staging block,
// which then carries the outputs to the TMA store engine (tmC: y, tmC2: y + add2).
__device__ __forceinline__ void ln_epilogue_tile(const GemmEpi& e, const CUtensorMap* tmC, const CUtensorMap* tmC2, const int m0,
                                                 const uint32_t tmem_acc, uint64_t* full_bar, const uint32_t full_phase, uint64_t* empty_bar,
                                                 unsigned char* buf0, const float* bias_s, const float* gamma_s, const float* beta_s,
                                                 float* stat_s, const int qd, const int hsel, const int lane) {
    const op16_t* resp = reinterpret_cast<const op16_t*>(e.residual);
    const op16_t* add2 = reinterpret_cast<const op16_t*>(e.ln.add2);
    const int row = qd * 32 + lane;
    const uint32_t swz = (uint32_t)(lane & 7);
    unsigned char* srow = buf0 + lane * 128;
    uint4 rr[8];                                                      // one coalesced block in flight (residual, then add2)
    ln_load_blk(rr, resp, e.ldr, hsel, m0, qd, lane, e.M);
    mbar_wait(full_bar, full_phase);
    tcgen05_fence_after();
    uint32_t xp[64];                                                  // the row's 128 pre-norm values of this warp, packed bf16x2
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int blk = 0; blk < 2; ++blk) {
        const int cb = hsel + 2 * blk;
        if (resp) {
            ln_stage_blk(rr, buf0, lane);
            if (blk == 0) ln_load_blk(rr, resp, e.ldr, hsel + 2, m0, qd, lane, e.M);        // next block's residual in flight behind this block's math
        }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {                              // 32 accumulator columns at a time (register budget)
            uint32_t acc[32];
            tmem_ld32(tmem_acc + ((uint32_t)(qd * 32) << 16) + (uint32_t)(cb * 64 + hf * 32), acc);
            if (blk == 1 && hf == 1) {                                // last TMEM read of this accumulator by this warp
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(empty_bar);
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int k = hf * 4 + kk;
                uint4 r4 = make_uint4(0, 0, 0, 0);
                if (resp) r4 = *reinterpret_cast<const uint4*>(srow + ((k ^ swz) * 16));
                const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
                const float4 ba = *reinterpret_cast<const float4*>(bias_s + cb * 64 + k * 8);
                const float4 bb = *reinterpret_cast<const float4*>(bias_s + cb * 64 + k * 8 + 4);
                const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float x0 = __uint_as_float(acc[kk * 8 + 2 * i]) + bv[2 * i] + op16_lo_f32(rw[i]);
                    const float x1 = __uint_as_float(acc[kk * 8 + 2 * i + 1]) + bv[2 * i + 1] + op16_hi_f32(rw[i]);
                    const uint32_t pk = pack_bf16x2(x0, x1);
                    xp[blk * 32 + k * 4 + i] = pk;
                    const float y0 = op16_lo_f32(pk), y1 = op16_hi_f32(pk);
                    sum += y0 + y1;
                    sq = fmaf(y0, y0, sq);
                    sq = fmaf(y1, y1, sq);
                }
            }
        }
        if (resp) __syncwarp();                                       // the block is rewritten next
    }
    if (add2) ln_load_blk(rr, add2, e.ln.ld2, hsel, m0, qd, lane, e.M);                     // in flight across the statistics exchange
    stat_s[(row * 2 + hsel) * 2] = sum;
    stat_s[(row * 2 + hsel) * 2 + 1] = sq;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");       // the two warps of this lane quarter
    const float mean = (stat_s[row * 4] + stat_s[row * 4 + 2]) * (1.f / 256.f);
    const float var = fmaxf((stat_s[row * 4 + 1] + stat_s[row * 4 + 3]) * (1.f / 256.f) - mean * mean, 0.f);
    const float rstd = rsqrtf(var + e.ln.eps);
    asm volatile("bar.sync %0, 64;" ::"r"(1 + qd) : "memory");       // statistics consumed before the next tile overwrites them
#pragma unroll
    for (int blk = 0; blk < 2; ++blk) {
        const int cb = hsel + 2 * blk;
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = ln_norm_pair(xp[blk * 32 + k * 4 + i], mean, rstd, gamma_s + cb * 64 + k * 8 + 2 * i, beta_s + cb * 64 + k * 8 + 2 * i);
            *reinterpret_cast<uint4*>(srow + ((k ^ swz) * 16)) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tma_store_2d(tmC, buf0, cb * 64, m0 + qd * 32);
            tma_store_commit();
        }
        if (add2) {                                                   // second output: the ROUNDED y plus add2, as un-fused
            ln_stage_blk(rr, buf0, lane);                                            // (waits for the y store to finish reading the block)
            if (blk == 0) ln_load_blk(rr, add2, e.ln.ld2, hsel + 2, m0, qd, lane, e.M);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint4* p = reinterpret_cast<uint4*>(srow + ((k ^ swz) * 16));
                const uint4 t4 = *p;
                const uint32_t tw[4] = {t4.x, t4.y, t4.z, t4.w};
                uint32_t z[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t y = ln_norm_pair(xp[blk * 32 + k * 4 + i], mean, rstd, gamma_s + cb * 64 + k * 8 + 2 * i, beta_s + cb * 64 + k * 8 + 2 * i);
                    z[i] = pack_bf16x2(op16_lo_f32(y) + op16_lo_f32(tw[i]),
                                       op16_hi_f32(y) + op16_hi_f32(tw[i]));
                }
                *p = make_uint4(z[0], z[1], z[2], z[3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(tmC2, buf0, cb * 64, m0 + qd * 32);
                tma_store_commit();
            }
        }
    }
}

// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM owner), warps 2-9 = epilogue
// (two warps per TMEM lane quarter, each draining half of the tile's columns).
// Two TMEM accumulators: the MMA warp fills one while the epilogue drains the other.  Epilogue: TMEM -> registers
// (bias, ReLU) -> XOR-swizzled shared staging -> fully coalesced 16-byte global stores (+ coalesced residual reads).
template <int BN, int STAGES, typename OutT>
__global__ void __launch_bounds__(320, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmC2, const GemmEpi e) {
    using S = GemmSmem<BN, STAGES, OutT>;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment: required by the 128B swizzle pattern shared between TMA and the UMMA descriptors
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* staging = smem + STAGES * S::STAGE_BYTES;
    float* bias_s = reinterpret_cast<float*>(staging + S::STAGING_BYTES);
    float* ln_stat_s = reinterpret_cast<float*>(staging + S::STAGING_BYTES + S::BIAS_BYTES);     // [2][128][4]
    float* ln_gb_s = ln_stat_s + 2 * GEMM_BM * 4;                                                // gamma[BN], beta[BN]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + S::STAGING_BYTES + S::BIAS_BYTES + S::LN_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_m = (e.M + GEMM_BM - 1) / GEMM_BM, num_n = (e.N + BN - 1) / BN;
    const int num_tiles = num_m * num_n;
    const int Kl = e.split3 ? e.K / 3 : 0;                                   // logical K of a split-precision product
    // tile order.  Default: row tiles fastest (the CTAs of a wave share one weight tile).  nfast: column tiles fastest -- the CTAs of a wave
    // share their A tiles through L2.  The split-precision products stream a large A (up to 717 MB) beside an fp32 / split output of the same
    // size; with row tiles fastest every column tile re-read A from DRAM (ncu, profiles/r2_split_gemm_ncu.txt: 1.0 GB read for 0.48 GB of A
    // at N = 256, 0.69 GB for 0.06 GB at N = 2048 -- the output stream evicts it)
    const bool nfast = e.nfast != 0;
    const int num_kb = e.split3 ? 2 * (Kl / GEMM_BK) : (e.K + GEMM_BK - 1) / GEMM_BK;   // pipeline steps per tile

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], 8);          // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<2 * BN>(tmem_ptr);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_launch_dependents();
    pdl_wait();                                        // everything above overlapped the previous kernel's tail

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (nfast ? tile / num_n : tile % num_m) * GEMM_BM, n0 = (nfast ? tile % num_n : tile / num_m) * BN;
                // implicit-GEMM conv: the (image, row, column) of every pixel run of the tile is fixed for the whole K loop and the tap
                // (kh, kw, channel block) advances by counters -- round 1 recomputed both with ~8 integer divisions per k-step on this
                // single producer thread, which made the producer the bottleneck of the deep convs (layer4: 0.85 us per k-step against
                // 0.11 us of MMA; ncu: L2 / tensor pipe / l1tex all below 17 %)
                const ConvGeo& cg = e.conv;
                int seg_x[4], seg_y[4], seg_b[4], nvalid = 0, ck = 0, ckw = 0, ckh = 0;        // (registers: every loop over them is unrolled)
                const bool hoisted = cg.enabled && cg.nseg <= 4;
                if (hoisted) {
                    const int hw = cg.Ho * cg.Wo;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int p = m0 + j * cg.seg_w;
                        seg_b[j] = seg_y[j] = seg_x[j] = 0;
                        if (j < cg.nseg && p < e.M) {
                            const int bimg = p / hw, rem = p - bimg * hw;
                            const int ho = rem / cg.Wo;
                            seg_b[j] = bimg; seg_y[j] = ho * cg.stride - cg.pad; seg_x[j] = (rem - ho * cg.Wo) * cg.stride - cg.pad;
                            nvalid = j + 1;
                        }
                    }
                }
                const uint32_t conv_tx = (uint32_t)(nvalid * cg.seg_w * GEMM_BK * 2 + S::B_BYTES);
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);                   // slot free (first lap passes immediately)
                    unsigned char* sa = smem + s * S::STAGE_BYTES;
                    if (hoisted) {
                        mbar_expect_tx(&full_bar[s], conv_tx);
                        const int c0 = ck * GEMM_BK;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (j < nvalid)
                                tma_load_4d(sa + (size_t)j * cg.seg_w * (GEMM_BK * 2), &tmA, &full_bar[s], c0, seg_x[j] + ckw, seg_y[j] + ckh, seg_b[j]);
                        if (++ck == cg.cblocks) { ck = 0; if (++ckw == cg.KW) { ckw = 0; ++ckh; } }
                    } else if (cg.enabled) {                            // narrow maps (more than 4 pixel runs per tile): per-step arithmetic
                        const int kpos = kb / cg.cblocks, c0 = (kb - kpos * cg.cblocks) * GEMM_BK;
                        const int kh = kpos / cg.KW, kw = kpos - kh * cg.KW;
                        int nv = 0;
                        for (int j = 0; j < cg.nseg; ++j) nv += (m0 + j * cg.seg_w < e.M) ? 1 : 0;
                        mbar_expect_tx(&full_bar[s], (uint32_t)(nv * cg.seg_w * GEMM_BK * 2 + S::B_BYTES));
                        for (int j = 0; j < nv; ++j) {
                            const int p = m0 + j * cg.seg_w;
                            const int hw = cg.Ho * cg.Wo;
                            const int bimg = p / hw, rem = p - bimg * hw;
                            const int ho = rem / cg.Wo, wo = rem - ho * cg.Wo;
                            tma_load_4d(sa + (size_t)j * cg.seg_w * (GEMM_BK * 2), &tmA, &full_bar[s], c0, wo * cg.stride + kw - cg.pad,
                                        ho * cg.stride + kh - cg.pad, bimg);
                        }
                    } else if (e.split3) {                              // step 2j: (A_hi, W_hi) of logical block j; step 2j+1: (A_lo, W_lo)
                        const int j = kb >> 1, part = kb & 1;
                        mbar_expect_tx(&full_bar[s], S::STAGE_BYTES);
                        tma_load_2d(sa, &tmA, &full_bar[s], (part ? 2 * Kl : 0) + j * GEMM_BK, m0);
                        tma_load_2d(sa + S::A_BYTES, &tmB, &full_bar[s], (part ? Kl : 0) + j * GEMM_BK, n0);
                        continue;
                    } else {
                        mbar_expect_tx(&full_bar[s], S::STAGE_BYTES);
                        tma_load_2d(sa, &tmA, &full_bar[s], kb * GEMM_BK, m0);
                    }
                    tma_load_2d(sa + S::A_BYTES, &tmB, &full_bar[s], kb * GEMM_BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
        // A,B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
        constexpr uint32_t IDESC = (1u << 4) | OP16_IDESC_AB | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(GEMM_BM >> 4) << 24);
        uint32_t it = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
            const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
            mbar_wait(&tmem_empty_bar[as], aph ^ 1);                    // epilogue has drained this accumulator
            tcgen05_fence_after();
            const uint32_t tmem_d = tmem_base + as * BN;
            if (e.split3) {
                // split-precision product: stages (s0, s1) = (A_hi | W_hi, A_lo | W_lo) of one logical k-block; hi.hi, hi.lo, lo.hi
                for (int kb = 0; kb < num_kb; kb += 2, it += 2) {
                    const int s0 = it % STAGES, s1 = (it + 1) % STAGES;
                    mbar_wait(&full_bar[s0], (it / STAGES) & 1);
                    mbar_wait(&full_bar[s1], ((it + 1) / STAGES) & 1);
                    tcgen05_fence_after();
                    if (elect_one()) {
                        const uint32_t sa0 = smem_u32(smem + s0 * S::STAGE_BYTES), sa1 = smem_u32(smem + s1 * S::STAGE_BYTES);
                        const uint64_t da0 = make_sw128_kmajor_desc(sa0), db0 = make_sw128_kmajor_desc(sa0 + S::A_BYTES);
                        const uint64_t da1 = make_sw128_kmajor_desc(sa1), db1 = make_sw128_kmajor_desc(sa1 + S::A_BYTES);
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k) umma_bf16(tmem_d, da0 + (uint64_t)(2 * k), db0 + (uint64_t)(2 * k), IDESC, (kb | k) != 0);
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k) umma_bf16(tmem_d, da0 + (uint64_t)(2 * k), db1 + (uint64_t)(2 * k), IDESC, 1);
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k) umma_bf16(tmem_d, da1 + (uint64_t)(2 * k), db0 + (uint64_t)(2 * k), IDESC, 1);
                        umma_commit(&empty_bar[s0]);
                        umma_commit(&empty_bar[s1]);
                        if (kb == num_kb - 2) umma_commit(&tmem_full_bar[as]);
                    }
                    __syncwarp();
                }
                continue;
            }
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);                            // TMA bytes have landed
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint32_t sa = smem_u32(smem + s * S::STAGE_BYTES);
                    const uint64_t da = make_sw128_kmajor_desc(sa);
                    const uint64_t db = make_sw128_kmajor_desc(sa + S::A_BYTES);
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        if (e.dbg & 2) break;
                        // advance 16 bf16 = 32 bytes inside the swizzle atom: +2 in the (addr>>4) field
                        umma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDESC, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[s]);                         // frees the smem slot when these MMAs retire
                    if (kb == num_kb - 1) umma_commit(&tmem_full_bar[as]);   // accumulator complete -> epilogue
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue: warp w owns TMEM lane quarter (w % 4) = 32 tile rows, and column half (w-2)/4 of them =====
        const int qd = warp & 3;
        const int half = (warp - 2) >> 2;
        constexpr int HW_COLS = BN / 2;
        constexpr int CHUNKS = S::ROW_BYTES / 16;                       // 16-byte chunks per staging row
        constexpr int EPC = 16 / (int)sizeof(OutT);                     // elements per chunk
        unsigned char* my_rows = staging + (size_t)(qd * 32) * S::ROW_BYTES;
        const bool vec_ok = ((e.ldc * (int)sizeof(OutT)) % 16 == 0) && (((uintptr_t)e.C & 15) == 0) &&
                            (!e.residual || (((e.ldr * (int)sizeof(OutT)) % 16 == 0) && (((uintptr_t)e.residual & 15) == 0)));
        bool do_ln = false;
        if constexpr (S::LN_BYTES > 0) do_ln = e.ln.gamma != nullptr;
        if constexpr (S::LN_BYTES > 0) {
            if (do_ln) {
                // full-row (N == BN == 256) LayerNorm epilogue: thread = row, TMA stores (ln_epilogue_tile)
                for (int j = threadIdx.x - 64; j < BN; j += 256) {
                    ln_gb_s[j] = e.ln.gamma[j];
                    ln_gb_s[BN + j] = e.ln.beta[j];
                    bias_s[j] = e.bias ? __ldg(e.bias + j) : 0.f;
                }
                asm volatile("bar.sync 5, 256;" ::: "memory");          // the 8 epilogue warps only
                unsigned char* buf0 = staging + (warp - 2) * 4096;
                uint32_t tc = 0;
                for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tc) {
                    const uint32_t as = tc & 1, aph = (tc >> 1) & 1;
                    ln_epilogue_tile(e, &tmC, &tmC2, (tile % num_m) * GEMM_BM, tmem_base + as * BN, &tmem_full_bar[as], aph,
                                     &tmem_empty_bar[as], buf0, bias_s, ln_gb_s, ln_gb_s + BN, ln_stat_s, qd, half, lane);
                }
                if (lane == 0) tma_store_wait<0>();
            }
        }
        if (!do_ln) {
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
            const int m0 = (nfast ? tile / num_n : tile % num_m) * GEMM_BM, n0 = (nfast ? tile % num_n : tile / num_m) * BN;
            const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
            // bias row of this tile -> warp-private shared memory while the MMAs of the tile are still running
            float* my_bias = bias_s + (warp - 2) * HW_COLS - half * HW_COLS;   // indexed by tile column
            if (e.bias) {
                for (int j = half * HW_COLS + lane; j < (half + 1) * HW_COLS; j += 32) my_bias[j] = (n0 + j < e.N) ? __ldg(e.bias + n0 + j) : 0.f;
                __syncwarp();
            }
            mbar_wait(&tmem_full_bar[as], aph);
            tcgen05_fence_after();
            // ---- step 1: TMEM -> registers -> bias / ReLU -> swizzled staging (thread = row qd*32 + lane)
#pragma unroll 1
            for (int c = half * HW_COLS; c < (half + 1) * HW_COLS; c += 32) {
                if (e.dbg & 4) break;
                uint32_t acc[32];
                tmem_ld32(tmem_base + as * BN + ((uint32_t)(qd * 32) << 16) + (uint32_t)c, acc);
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
                if (e.bias) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(my_bias + c + j);
                        v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                    }
                }
                if (e.relu == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                unsigned char* srow = my_rows + (size_t)lane * S::ROW_BYTES;
                if constexpr (sizeof(OutT) == 2) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        uint4 o;
                        o.x = pack_bf16x2(v[j], v[j + 1]); o.y = pack_bf16x2(v[j + 2], v[j + 3]);
                        o.z = pack_bf16x2(v[j + 4], v[j + 5]); o.w = pack_bf16x2(v[j + 6], v[j + 7]);
                        const int chunk = (c + j) / 8;
                        *reinterpret_cast<uint4*>(srow + ((chunk ^ (lane & 7)) * 16)) = o;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const int chunk = (c + j) / 4;
                        *reinterpret_cast<float4*>(srow + ((chunk ^ (lane & 7)) * 16)) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                }
            }
            // all TMEM reads of this accumulator are done -> hand it back to the MMA warp
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
            // ---- step 2: staging -> global, coalesced (each warp stores its own 32 rows), + residual, + late ReLU
            const int ncols = (e.dbg & 1) ? 0 : min(BN, e.N - n0);
            if (vec_ok && (ncols % EPC) == 0) {
                const int nchunks = ncols / EPC;
                constexpr int CW = CHUNKS / 2;                           // chunks per row owned by this warp
                constexpr int NIT = CW;                                  // 32 rows * CW chunks / 32 lanes
                // all residual loads of this warp's 32 x (BN/2) block are issued before anything consumes them
                constexpr int GRP = NIT < 8 ? NIT : 8;                   // residual prefetch depth (register budget)
#pragma unroll 1
                for (int g0 = 0; g0 < NIT; g0 += GRP) {
                uint4 rres[GRP];
                if (e.residual) {
#pragma unroll
                    for (int it = 0; it < GRP; ++it) {
                        const int idx = (g0 + it) * 32 + lane;
                        const int r = idx / CW, ch = half * CW + idx % CW;
                        const int grow = m0 + qd * 32 + r;
                        rres[it] = make_uint4(0, 0, 0, 0);
                        if (ch < nchunks && grow < e.M)
                            rres[it] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const OutT*>(e.residual) + (size_t)grow * e.ldr + n0 + ch * EPC));
                    }
                }
#pragma unroll
                for (int it = 0; it < GRP; ++it) {
                    const int idx = (g0 + it) * 32 + lane;
                    const int r = idx / CW, ch = half * CW + idx % CW;
                    const int grow = m0 + qd * 32 + r;
                    if (ch >= nchunks || grow >= e.M) continue;
                    uint4 d = *reinterpret_cast<const uint4*>(my_rows + (size_t)r * S::ROW_BYTES + ((ch ^ (r & 7)) * 16));
                    OutT* cp = reinterpret_cast<OutT*>(e.C) + (size_t)grow * e.ldc + n0 + ch * EPC;
                    if (e.residual || e.relu == 2) {
                        float f[EPC];
                        unpack_chunk<OutT>(d, f);
                        if (e.residual) {
                            float g[EPC];
                            unpack_chunk<OutT>(rres[it], g);
#pragma unroll
                            for (int k = 0; k < EPC; ++k) f[k] = e.relu == 3 ? (g[k] > 0.f ? f[k] : 0.f) : f[k] + g[k];
                        }
                        if (e.relu == 2) {
#pragma unroll
                            for (int k = 0; k < EPC; ++k) f[k] = fmaxf(f[k], 0.f);
                        }
                        d = pack_chunk<OutT>(f);
                    }
                    if constexpr (sizeof(OutT) == 4) {
                        if (e.split_out) {       // 4 fp32 results -> hi at column n and n + N, lo = rn16(x - hi) at n + 2N (8-byte stores)
                            float f[4];
                            unpack_chunk<float>(d, f);
                            float l[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) l[k] = f[k] - op16_to_f32(f32_to_op16(f[k]));
                            const uint2 hi2 = make_uint2(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]));
                            const uint2 lo2 = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
                            op16_t* sp = reinterpret_cast<op16_t*>(e.C) + (size_t)grow * e.ldc + n0 + ch * 4;
                            *reinterpret_cast<uint2*>(sp) = hi2;
                            *reinterpret_cast<uint2*>(sp + e.N) = hi2;
                            *reinterpret_cast<uint2*>(sp + 2 * (size_t)e.N) = lo2;
                            continue;
                        }
                    }
                    *reinterpret_cast<uint4*>(cp) = d;
                }
                }
            } else {
                for (int idx = lane; idx < 32 * HW_COLS; idx += 32) {
                    const int r = idx / HW_COLS, col = half * HW_COLS + idx % HW_COLS;
                    const int grow = m0 + qd * 32 + r;
                    if (col >= ncols || grow >= e.M) continue;
                    const int ch = col / EPC;
                    const OutT x = *reinterpret_cast<const OutT*>(my_rows + (size_t)r * S::ROW_BYTES + ((ch ^ (r & 7)) * 16) + (col % EPC) * sizeof(OutT));
                    float f = (float)x;
                    if (e.residual) {
                        const float r = (float)reinterpret_cast<const OutT*>(e.residual)[(size_t)grow * e.ldr + n0 + col];
                        f = e.relu == 3 ? (r > 0.f ? f : 0.f) : f + r;
                    }
                    if (e.relu == 2) f = fmaxf(f, 0.f);
                    reinterpret_cast<OutT*>(e.C)[(size_t)grow * e.ldc + n0 + col] = (OutT)f;
                }
            }
            __syncwarp();          // staging rows are rewritten by the next tile
        }
        }   // !do_ln
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc<2 * BN>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------- weight-stationary tcgen05 GEMM
// K <= 256 (every Linear of the transformer except linear2 / ref_point_head.0, and the 1x1 convolutions of layer1-3).
// Measured on the tile kernel above (tools/gemm_probe.py, CUDA-graph timing): with K = 256 the loads need 10 us and the MMAs
// 3 us of a 18 us launch -- the EPILOGUE (TMEM -> registers -> staging -> LDS -> STG with per-lane address arithmetic) is what
// bounds it, and the weight tile is re-fetched with every output tile.  Here:
//  * a persistent CTA owns one BN-wide slice of W for its whole life (loaded once, up to 128 KB resident in shared memory)
//    and streams only A tiles through the TMA ring (16 KB stages); with BN = 256 a 256-wide layer reads A exactly once.
//    CTA c works on slice c % ns and row tiles c / ns, c / ns + cps, ... so the ns CTAs that share an A tile fetch it at the
//    same time (one HBM read, the rest L2 hits).
//  * the epilogue never touches global memory through the LSU: each of the 8 epilogue warps converts a 32-row x 128-byte block
//    (thread = row: tcgen05.ld, bias, ReLU, pack), writes it 128B-swizzled into its own 4 KB buffer (conflict-free 16-byte
//    stores) and one lane hands it to the TMA store engine (cp.async.bulk.tensor, bulk-group completion; rows / columns beyond
//    M / N are clipped by the hardware).  A residual operand arrives the same way in the other direction: TMA-loaded one
//    block ahead into the buffer the result will overwrite (two buffers per warp), read back swizzled by the row's thread.
template <int BN, typename OutT, bool RES, bool LN = false>
struct WsSmem {
    static constexpr int KB_MAX = 4;
    static constexpr int W_KB_BYTES = BN * GEMM_BK * 2;                       // one 64-wide k-block of the slice
    static constexpr int W_BYTES = KB_MAX * W_KB_BYTES;
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;                     // 16 KB
    static constexpr int BLK_BYTES = 32 * 128;                                // one epilogue block: 32 rows x 128 bytes
    static constexpr int NBUF = RES ? 2 : 1;
    static constexpr int STG_WARP = NBUF * BLK_BYTES;
    static constexpr int STG_BYTES = 8 * STG_WARP;
    static constexpr int BIAS_BYTES = BN * 4 + (LN ? 2 * BN * 4 /*gamma, beta*/ + GEMM_BM * 4 * 4 /*row statistics*/ : 0);
    static constexpr int BAR_BYTES = 512;
    static constexpr int FIXED = W_BYTES + STG_BYTES + BIAS_BYTES + BAR_BYTES + 1024 /*alignment slack*/;
    static constexpr int STAGES_FIT = (232448 - FIXED) / A_BYTES;
    static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
    static constexpr int TOTAL = FIXED + STAGES * A_BYTES;
    static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
    static_assert(STAGES >= 3, "weight slice leaves too little room for the A ring");
};

template <typename OutT, int CB> struct TmemBlock;
template <> struct TmemBlock<op16_t, 64> {
    static __device__ __forceinline__ void load(uint32_t taddr, uint32_t (&r)[64]) { tmem_ld64(taddr, r); }
};
template <> struct TmemBlock<float, 32> {
    static __device__ __forceinline__ void load(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld32(taddr, r); }
};

template <int BN, typename OutT, bool RES, bool LN = false>
__global__ void __launch_bounds__(320, 1)
gemm_ws_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const GemmEpi e,
                       const int ns, const int cps) {
    using S = WsSmem<BN, OutT, RES, LN>;
    static_assert(!LN || (BN == 256 && sizeof(OutT) == 2 && !RES), "LayerNorm epilogue: one 256-wide bf16 slice, residual through the LSU");
    constexpr int STAGES = S::STAGES;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* wreg = smem;                                          // [KB_MAX][BN rows x 128 B], 128B swizzle
    unsigned char* aring = smem + S::W_BYTES;                            // [STAGES][128 rows x 128 B]
    unsigned char* staging = aring + STAGES * S::A_BYTES;                // [8 warps][NBUF][32 rows x 128 B], 128B swizzle
    float* bias_s = reinterpret_cast<float*>(staging + S::STG_BYTES);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + S::STG_BYTES + S::BIAS_BYTES);
    uint64_t* empty_bar = full_bar + 8;
    uint64_t* w_bar = empty_bar + 8;                   // [KB_MAX]
    uint64_t* tmem_full_bar = w_bar + S::KB_MAX;       // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
    uint64_t* res_bar = tmem_empty_bar + 2;            // [8 warps][2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_bar + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slice = blockIdx.x % ns, r0 = blockIdx.x / ns;
    if (r0 >= cps) return;                                               // CTAs beyond ns * cps have no work (whole CTA leaves)
    const int n0 = slice * BN;
    const int num_m = (e.M + GEMM_BM - 1) / GEMM_BM;
    const int num_kb = (e.K + GEMM_BK - 1) / GEMM_BK;
    unsigned int* const dbgbuf = (blockIdx.x == 0) ? e.dbgbuf : nullptr;
#define WS_DBG(role, unit, slot) do { if (dbgbuf && (unit) < 64) dbgbuf[(((role) * 64 + (unit)) * 16 + (slot))] = (unsigned int)clock(); } while (0)
    WS_DBG(3, 0, warp < 15 ? warp : 15);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmC);
        if (RES) tma_prefetch_desc(&tmR);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int k = 0; k < S::KB_MAX; ++k) mbar_init(&w_bar[k], 1);
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], 8);
        }
        for (int k = 0; k < 16; ++k) mbar_init(&res_bar[k], 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<S::TMEM_COLS>(tmem_ptr);
    for (int j = threadIdx.x; j < BN; j += 320) bias_s[j] = (e.bias && n0 + j < e.N) ? __ldg(e.bias + n0 + j) : 0.f;
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_launch_dependents();
    if (warp != 0) pdl_wait();                         // (the producer first fetches the constant weight slice)
    if (warp == 1) WS_DBG(3, 1, 0);

    if (warp == 0) {
        // ===== TMA producer: the weight slice once, then A tiles
        if (elect_one()) {
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_expect_tx(&w_bar[kb], S::W_KB_BYTES);
                tma_load_2d(wreg + kb * S::W_KB_BYTES, &tmB, &w_bar[kb], kb * GEMM_BK, n0);
            }
            pdl_wait();                                // activations of the previous kernel from here on
            uint32_t it = 0;
            int tcd = 0;
            for (int mt = r0; mt < num_m; mt += cps, ++tcd) {
                WS_DBG(0, tcd, 0);
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    WS_DBG(0, tcd, 1 + kb);
                    mbar_expect_tx(&full_bar[s], S::A_BYTES);
                    tma_load_2d(aring + s * S::A_BYTES, &tmA, &full_bar[s], kb * GEMM_BK, mt * GEMM_BM);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer
        constexpr uint32_t IDESC = (1u << 4) | OP16_IDESC_AB | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(GEMM_BM >> 4) << 24);
        uint32_t it = 0, tcount = 0;
        for (int mt = r0; mt < num_m; mt += cps, ++tcount) {
            const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
            WS_DBG(1, tcount, 0);
            mbar_wait(&tmem_empty_bar[as], aph ^ 1);
            tcgen05_fence_after();
            WS_DBG(1, tcount, 1);
            const uint32_t tmem_d = tmem_base + as * BN;
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                if (tcount == 0) mbar_wait(&w_bar[kb], 0);              // weight k-block resident (first tile only)
                mbar_wait(&full_bar[s], ph);
                tcgen05_fence_after();
                WS_DBG(1, tcount, 2 + 2 * kb);
                if (elect_one()) {
                    const uint64_t da = make_sw128_kmajor_desc(smem_u32(aring + s * S::A_BYTES));
                    const uint64_t db = make_sw128_kmajor_desc(smem_u32(wreg + kb * S::W_KB_BYTES));
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        if (e.dbg & 2) break;
                        umma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDESC, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[s]);
                    if (kb == num_kb - 1) umma_commit(&tmem_full_bar[as]);
                }
                __syncwarp();
                WS_DBG(1, tcount, 3 + 2 * kb);
            }
        }
    } else {
        // ===== epilogue: warp w owns TMEM lane quarter (w % 4) = 32 tile rows and every second 128-byte column block
        const int qd = warp & 3;
        const int hsel = (warp - 2) >> 2;
        constexpr int EPC = 16 / (int)sizeof(OutT);                      // elements per 16-byte chunk
        constexpr int CB = 8 * EPC;                                      // columns per block (64 bf16 / 32 fp32)
        constexpr int NCB = BN / CB;
        unsigned char* buf0 = staging + (warp - 2) * S::STG_WARP;
        uint64_t* rbar = res_bar + (warp - 2) * 2;
        const uint32_t swz = (uint32_t)(lane & 7);
        uint32_t tcount = 0, bcount = 0;
        if constexpr (LN) {
            float* gamma_s = bias_s + BN;
            float* beta_s = gamma_s + BN;
            float* stat_s = beta_s + BN;                                 // [128 rows][2 halves][2]
            for (int j = threadIdx.x - 64; j < BN; j += 256) { gamma_s[j] = e.ln.gamma[j]; beta_s[j] = e.ln.beta[j]; }
            asm volatile("bar.sync 5, 256;" ::: "memory");               // the 8 epilogue warps only
            for (int mt = r0; mt < num_m; mt += cps, ++tcount) {
                const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
                ln_epilogue_tile(e, &tmC, &tmR, mt * GEMM_BM, tmem_base + as * BN, &tmem_full_bar[as], aph, &tmem_empty_bar[as], buf0,
                                 bias_s, gamma_s, beta_s, stat_s, qd, hsel, lane);
            }
            if (lane == 0) tma_store_wait<0>();
        } else {
        if (RES && hsel < NCB && r0 < num_m && lane == 0) {              // residual of this warp's first block
            mbar_expect_tx(&rbar[0], S::BLK_BYTES);
            tma_load_2d(buf0, &tmR, &rbar[0], n0 + hsel * CB, r0 * GEMM_BM + qd * 32);
        }
        for (int mt = r0; mt < num_m; mt += cps, ++tcount) {
            const int m0 = mt * GEMM_BM;
            const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
            bool waited = false;
            int blkd = 0;
            if (threadIdx.x == 64) WS_DBG(2, tcount, 0);
            if (!(e.dbg & 4)) {
                // TMEM reads in 32-column pieces, one piece ahead: while piece i is converted / staged, piece i + 1 is on its way
                constexpr int PPB = CB / 32;                             // pieces per block (16-bit output: 2, fp32: 1)
                constexpr int NBLK = (NCB + 1) / 2;                      // blocks of a tile this warp may own
                const uint32_t tacc = tmem_base + as * BN + ((uint32_t)(qd * 32) << 16);
                uint32_t ra[32], rb[32];
#pragma unroll
                for (int bi = 0; bi < NBLK; ++bi) {
                    const int cb = hsel + 2 * bi;
                    if (cb >= NCB) break;
                    unsigned char* buf = buf0 + (RES ? (bcount & 1) * S::BLK_BYTES : 0);
                    if (RES && lane == 0) {
                        // the other buffer was last read by the store of block bcount-1 (the newest bulk group): once that read is
                        // done, fetch the residual of block bcount+1 into it
                        tma_store_wait_read<0>();
                        int ncb = cb + 2, nmt = mt;
                        if (ncb >= NCB) { ncb = hsel; nmt = mt + cps; }
                        if (nmt < num_m) {
                            uint64_t* nb = &rbar[(bcount + 1) & 1];
                            mbar_expect_tx(nb, S::BLK_BYTES);
                            tma_load_2d(buf0 + ((bcount + 1) & 1) * S::BLK_BYTES, &tmR, nb, n0 + ncb * CB, nmt * GEMM_BM + qd * 32);
                        }
                    }
                    if (!waited) {
                        mbar_wait(&tmem_full_bar[as], aph);
                        tcgen05_fence_after();
                        waited = true;
                        tmem_ld32_issue(tacc + (uint32_t)(cb * CB), ra);             // (bi == 0: the first piece of the tile)
                    }
                    if (threadIdx.x == 64) WS_DBG(2, tcount, 1 + 5 * blkd);
                    float v[CB];
#pragma unroll
                    for (int pp = 0; pp < PPB; ++pp) {
                        tmem_wait_ld();                                              // piece (cb, pp) has arrived
                        const bool more = (pp + 1 < PPB) || (cb + 2 < NCB);
                        const uint32_t nxt = tacc + (uint32_t)((pp + 1 < PPB) ? cb * CB + (pp + 1) * 32 : (cb + 2) * CB);
                        auto step = [&](uint32_t (&cur)[32], uint32_t (&other)[32]) {
                            if (more) {
                                tmem_ld32_issue(nxt, other);
                            } else {                                                 // last TMEM read of this accumulator by this warp
                                tcgen05_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
                            }
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 b4 = *reinterpret_cast<const float4*>(bias_s + cb * CB + pp * 32 + j);
                                v[pp * 32 + j] = __uint_as_float(cur[j]) + b4.x; v[pp * 32 + j + 1] = __uint_as_float(cur[j + 1]) + b4.y;
                                v[pp * 32 + j + 2] = __uint_as_float(cur[j + 2]) + b4.z; v[pp * 32 + j + 3] = __uint_as_float(cur[j + 3]) + b4.w;
                            }
                        };
                        if (((bi * PPB + pp) & 1) == 0) step(ra, rb); else step(rb, ra);
                    }
                    if (threadIdx.x == 64) WS_DBG(2, tcount, 2 + 5 * blkd);
                    if (e.relu == 1) {
#pragma unroll
                        for (int j = 0; j < CB; ++j) v[j] = fmaxf(v[j], 0.f);
                    }
                    unsigned char* srow = buf + lane * 128;
                    if (RES) {
                        mbar_wait(&rbar[bcount & 1], (bcount >> 1) & 1);  // residual block has landed (TMA, swizzled)
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            float g[EPC];
                            unpack_chunk<OutT>(*reinterpret_cast<const uint4*>(srow + ((k ^ swz) * 16)), g);
#pragma unroll
                            for (int i = 0; i < EPC; ++i) v[k * EPC + i] = e.relu == 3 ? (g[i] > 0.f ? v[k * EPC + i] : 0.f) : v[k * EPC + i] + g[i];
                        }
                    } else {
                        if (threadIdx.x == 64) WS_DBG(2, tcount, 3 + 5 * blkd);
                        if (lane == 0) tma_store_wait_read<0>();         // the previous store has finished reading this buffer
                        __syncwarp();
                    }
                    if (threadIdx.x == 64) WS_DBG(2, tcount, 4 + 5 * blkd);
                    if (e.relu == 2) {
#pragma unroll
                        for (int j = 0; j < CB; ++j) v[j] = fmaxf(v[j], 0.f);
                    }
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        *reinterpret_cast<uint4*>(srow + ((k ^ swz) * 16)) = pack_chunk<OutT>(v + k * EPC);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0 && !(e.dbg & 1)) {
                        tma_store_2d(&tmC, buf, n0 + cb * CB, m0 + qd * 32);
                        tma_store_commit();
                    }
                    if (threadIdx.x == 64) WS_DBG(2, tcount, 5 + 5 * blkd);
                    ++blkd;
                    ++bcount;
                }
            }
            if (!waited) {                                               // no block of this tile belongs to this warp
                mbar_wait(&tmem_full_bar[as], aph);
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
            }
        }
        if (lane == 0) tma_store_wait<0>();                              // the buffers must outlive the last store
        if (threadIdx.x == 64) WS_DBG(3, 2, 0);
        }   // !LN
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc<S::TMEM_COLS>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------- fp32 SIMT GEMM (parity mode)
// 64x64 tile, 16-deep K slab, 256 threads, 4x4 register micro-tile; exact fp32 FMA accumulation in K order.
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw, const GemmEpi e) {
    __shared__ float As[16][64 + 4];
    __shared__ float Ws[16][64 + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lr = tid >> 2, lk = (tid & 3) * 4;     // loader: row 0..63, k offset 0,4,8,12
    for (int k0 = 0; k0 < e.K; k0 += 16) {
        {
            const int gm = m0 + lr, gn = n0 + lr;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gk = k0 + lk + j;
                As[lk + j][lr] = (gm < e.M && gk < e.K) ? A[(size_t)gm * lda + gk] : 0.f;
                Ws[lk + j][lr] = (gn < e.N && gk < e.K) ? W[(size_t)gn * ldw + gk] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Ws[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* C = reinterpret_cast<float*>(e.C);
    const float* R = reinterpret_cast<const float*>(e.residual);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= e.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= e.N) continue;
            float v = acc[i][j];
            if (e.bias) v += e.bias[gn];
            if (e.relu == 1) v = fmaxf(v, 0.f);
            if (R) {
                const float r = R[(size_t)gm * e.ldr + gn];
                v = e.relu == 3 ? (r > 0.f ? v : 0.f) : v + r;
            }
            if (e.relu == 2) v = fmaxf(v, 0.f);
            C[(size_t)gm * e.ldc + gn] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------- host side
// 2-D bf16 tensor map: rows x cols (cols contiguous), row pitch ld elements, box = box_rows x 64 cols, 128B swizzle
static int make_tmap_bf16(CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return DTLR_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)GEMM_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, DTLR_TMAP_OP16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for %dx%d ld=%d", (int)r, rows, cols, ld);
        return DTLR_ERR_CUDA;
    }
    return DTLR_OK;
}

// 4-D bf16 tensor map over NHWC activations: dims (C, W, H, B), box (64 channels, seg_w pixels, 1, 1), 128B swizzle
static int make_tmap_nhwc(CUtensorMap* map, const void* base, int B, int H, int W, int C, int seg_w, int stride = 1) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return DTLR_ERR_CUDA;
    }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    // traversal stride s in W and H: a box spanning seg_w * s input pixels delivers ceil(seg_w * s / s) = seg_w of them
    cuuint32_t box[4] = {(cuuint32_t)GEMM_BK, (cuuint32_t)(seg_w * stride), 1, 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = enc(map, DTLR_TMAP_OP16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (NHWC 4-D) failed (%d) for B=%d H=%d W=%d C=%d", (int)r, B, H, W, C);
        return DTLR_ERR_CUDA;
    }
    return DTLR_OK;
}

// 2-D tensor map of an output / residual matrix for the TMA epilogue: box = 32 rows x 128 bytes, 128B swizzle
template <typename OutT>
static int make_tmap_out(CUtensorMap* map, const void* base, int rows, int cols, int ld) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return DTLR_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(OutT)};
    cuuint32_t box[2] = {(cuuint32_t)(128 / sizeof(OutT)), 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, sizeof(OutT) == 2 ? DTLR_TMAP_OP16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for the %dx%d output (ld=%d)", (int)r, rows, cols, ld);
        return DTLR_ERR_CUDA;
    }
    return DTLR_OK;
}

template <int BN, typename OutT, bool RES, bool LN = false>
static int launch_ws(const void* A, int lda, const void* W, int ldw, const GemmEpi& e, cudaStream_t st) {
    using S = WsSmem<BN, OutT, RES, LN>;
    CUtensorMap ta, tb, tc, tr;
    int rc;
    if ((rc = make_tmap_bf16(&ta, A, e.M, e.K, lda, GEMM_BM))) return rc;
    if ((rc = make_tmap_bf16(&tb, W, e.N, e.K, ldw, BN))) return rc;
    if ((rc = make_tmap_out<OutT>(&tc, e.C, e.M, e.N, e.ldc))) return rc;
    tr = tc;
    if (RES && (rc = make_tmap_out<OutT>(&tr, e.residual, e.M, e.N, e.ldr))) return rc;
    if (LN && e.ln.out2 && (rc = make_tmap_out<OutT>(&tr, e.ln.out2, e.M, e.N, e.ln.ld2))) return rc;   // LayerNorm variant: tmR carries y + add2
    auto k = gemm_ws_tcgen05_kernel<BN, OutT, RES, LN>;
    static bool configured = false;
    if (!configured) {
        DTLR_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
        configured = true;
    }
    const int ns = (e.N + BN - 1) / BN, num_m = (e.M + GEMM_BM - 1) / GEMM_BM;
    int cps = sm_count() / ns;
    if (cps > num_m) cps = num_m;
    GemmEpi e2 = e;
    e2.dbgbuf = g_gemm_dbgbuf;
    DTLR_CHECK_CUDA(launch_pdl(k, dim3(ns * cps), dim3(320), S::TOTAL, st, ta, tb, tc, tr, e2, ns, cps));
    return DTLR_OK;
}

// weight-stationary path: K <= 256, N a multiple of the slice width, 16-byte aligned output / residual rows, and enough row
// tiles that every CTA amortises its weight slice over >= 2 of them
template <typename OutT>
static bool ws_try(const void* A, int lda, const void* W, int ldw, const GemmEpi& e, cudaStream_t st, int* rc) {
    if ((g_debug_flags & 32) || e.split_out) return false;
    constexpr int EPC = 16 / (int)sizeof(OutT);
    // only the row PITCH has to be 16-byte aligned (TMA store clips the columns beyond N, the weight rows beyond N are zero-filled)
    if (e.K > 256 || (e.ldc % EPC) != 0 || ((uintptr_t)e.C & 15) != 0) return false;
    // the TMA store clips at 16-byte granularity: with N % EPC != 0 the last chunk of a row is written whole, so the columns up to
    // the next 16-byte boundary receive zeros -- only allowed when they are padding by construction (pitch == N rounded up)
    if ((e.N % EPC) != 0 && (e.ldc != (e.N + EPC - 1) / EPC * EPC || e.residual)) return false;
    if (e.residual && ((e.ldr % EPC) != 0 || ((uintptr_t)e.residual & 15) != 0)) return false;
    const long long num_m = (e.M + GEMM_BM - 1) / GEMM_BM;
    int bn = 0;
    if (e.residual) {                    // two epilogue buffers per warp: the weight slice is at most 128 wide
        if ((e.N % 128) == 0) bn = 128;
        else if ((e.N % 64) == 0) bn = 64;
    } else {
        if ((e.N % 256) == 0 && !(g_debug_flags & 16777216)) bn = 256;       // flag 16777216: 128-wide slices (8 A stages instead of 4), A/B
        else if ((e.N % 192) == 0 && !(g_debug_flags & 33554432)) bn = 192;  // flag 33554432: N = 384 as 3 x 128 instead of 2 x 192
        else if ((e.N % 128) == 0) bn = 128;
        else if ((e.N % 64) == 0) bn = 64;
        else if (e.N <= 256 && e.N > 16) bn = (e.N + 63) / 64 * 64;          // one ragged slice (e.g. the 166-class heads)
    }
    if (!bn) return false;
    const int ns_ = (e.N + bn - 1) / bn;
    if (ns_ > sm_count() || num_m * ns_ < 2ll * sm_count()) return false;
    if (e.residual) {
        *rc = bn == 128 ? launch_ws<128, OutT, true>(A, lda, W, ldw, e, st) : launch_ws<64, OutT, true>(A, lda, W, ldw, e, st);
        return true;
    }
    switch (bn) {
        case 256: *rc = launch_ws<256, OutT, false>(A, lda, W, ldw, e, st); break;
        case 192: *rc = launch_ws<192, OutT, false>(A, lda, W, ldw, e, st); break;
        case 128: *rc = launch_ws<128, OutT, false>(A, lda, W, ldw, e, st); break;
        default: *rc = launch_ws<64, OutT, false>(A, lda, W, ldw, e, st); break;
    }
    return true;
}

template <int BN, int STAGES, typename OutT>
static int launch_tc(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpi& e, cudaStream_t st, const CUtensorMap* tc = nullptr,
                     const CUtensorMap* tc2 = nullptr) {
    using S = GemmSmem<BN, STAGES, OutT>;
    auto k = gemm_bf16_tcgen05_kernel<BN, STAGES, OutT>;
    static bool configured = false;     // per template instantiation
    if (!configured) {
        DTLR_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
        configured = true;
    }
    const int tiles = ((e.M + GEMM_BM - 1) / GEMM_BM) * ((e.N + BN - 1) / BN);
    const int grid = tiles < sm_count() ? tiles : sm_count();
    DTLR_CHECK_CUDA(launch_pdl(k, dim3(grid), dim3(320), S::TOTAL, st, ta, tb, tc ? *tc : ta, tc2 ? *tc2 : ta, e));   // output maps: LayerNorm epilogue only
    return DTLR_OK;
}

}  // namespace dtlr

using namespace dtlr;


extern "C" int dtlr_gemm(const void* A, int lda, const void* W, int ldw, const float* bias, const void* residual,
                         int ldr, void* C, int ldc, int M, int N, int K, int in_dtype, int out_dtype, int relu,
                         void* stream) {
    DTLR_CHECK_ARG(M >= 0 && N > 0 && K > 0, "gemm: bad sizes M=%d N=%d K=%d", M, N, K);
    if (M == 0) return DTLR_OK;
    DTLR_CHECK_ARG(A && W && C, "gemm: null pointer");
    // in_dtype DTLR_SPLIT16: A = [hi | hi | lo], W = [hi | lo | hi] (K = 3 Kl columns each) -- the same product as the plain walk over K,
    // with the duplicate hi tiles loaded once where the tile kernel can (e.split3)
    bool split_in = false;
    if (in_dtype == DTLR_SPLIT16) {
        split_in = (K % 3) == 0 && ((K / 3) % GEMM_BK) == 0 && !(g_debug_flags & 67108864);      // flag 67108864: plain 3K walk, A/B
        in_dtype = DTLR_OP16;
    }
    const bool split_out = out_dtype == DTLR_SPLIT16;      // fp32 result written as the 16-bit [hi | hi | lo] operand of the next split product
    DTLR_CHECK_ARG(lda >= K && ldw >= K && ldc >= (split_out ? 3 * N : N) && (!residual || ldr >= N), "gemm: leading dimension too small");
    cudaStream_t st = (cudaStream_t)stream;
    GemmEpi e{bias, residual, C, ldr, ldc, M, N, K, relu, LnArgs{nullptr, nullptr, nullptr, nullptr, 0, 0.f}, ConvGeo{0, 0, 0, 0, 0, 0, 0, 0, 1}, g_debug_flags};
    if (split_out) {
        DTLR_CHECK_ARG(in_dtype == DTLR_OP16 && (N % 4) == 0 && (ldc % 8) == 0 && (((uintptr_t)C) & 15) == 0 &&
                       (!residual || ((ldr % 4) == 0 && (((uintptr_t)residual) & 15) == 0)),
                       "gemm: split output needs 16-bit operands, N %% 4 == 0 and 16-byte aligned rows");
        e.split_out = 1;            // e.ldc stays the pitch in 16-bit elements (the split stores use it; ldc % 8 also satisfies the fp32
        out_dtype = DTLR_F32;       // epilogue's 16-byte rule, and N % 4 == 0 keeps every tile on its vector path)
    }
    if (in_dtype == DTLR_F32) {
        DTLR_CHECK_ARG(out_dtype == DTLR_F32, "gemm: fp32 operands produce fp32 output");
        dim3 grid((M + 63) / 64, (N + 63) / 64);
        sgemm_kernel<<<grid, 256, 0, st>>>((const float*)A, lda, (const float*)W, ldw, e);
        DTLR_CHECK_LAUNCH();
        return DTLR_OK;
    }
    DTLR_CHECK_ARG(in_dtype == DTLR_OP16, "gemm: operands must be f32 or bf16");
    DTLR_CHECK_ARG(out_dtype == DTLR_OP16 || out_dtype == DTLR_F32, "gemm: output must be bf16 or f32");
    DTLR_CHECK_ARG((lda % 8) == 0 && (ldw % 8) == 0 && (((uintptr_t)A | (uintptr_t)W) & 15) == 0,
                   "gemm: bf16 operands need 16-byte aligned rows (lda=%d ldw=%d)", lda, ldw);
    CUtensorMap ta, tb;
    int rc = DTLR_OK;
    if (out_dtype == DTLR_OP16 ? ws_try<op16_t>(A, lda, W, ldw, e, st, &rc) : ws_try<float>(A, lda, W, ldw, e, st, &rc)) return rc;
    const long long row_tiles_ = (M + GEMM_BM - 1) / GEMM_BM;
    e.split3 = split_in ? 1 : 0;           // the 128- and 64-wide tile kernels below (even stage counts); 256-wide 16-bit tiles walk 3 Kl
    e.nfast = ((split_in || split_out) && !(g_debug_flags & 134217728)) ? 1 : 0;       // flag 134217728: row tiles fastest as before, A/B
    if (!split_in && out_dtype == DTLR_OP16 && (N % 256) == 0 && (K >= 512 || N >= 1024) && !(g_debug_flags & 8) &&
        (K < 1024 || ((row_tiles_ * (N / 128) + sm_count() - 1) / sm_count()) * 128 * 3 > ((row_tiles_ * (N / 256) + sm_count() - 1) / sm_count()) * 256 * 2)) {
        // 128 x 256 tiles: the A tile is shared by twice as many output columns (less L2 traffic per FLOP) and full-width
        // N = 256 layers become one tile per row block
        if ((rc = make_tmap_bf16(&ta, A, M, K, lda, GEMM_BM))) return rc;
        if ((rc = make_tmap_bf16(&tb, W, N, K, ldw, 256))) return rc;
        return launch_tc<256, 3, op16_t>(ta, tb, e, st);
    }
    // few tiles and a long K loop (e.g. the 3x3 stride-2 input_proj of the last level: M = 1024, N = 256, K = 18432 -> 16 tiles of
    // 128 x 128 on 148 SMs): 64-wide tiles double the CTAs that share the K loop
    const long long tiles128 = (long long)((M + GEMM_BM - 1) / GEMM_BM) * ((N + 127) / 128);
    if (N > 64 && !(tiles128 * 2 <= sm_count() && K >= 1024 && (N % 64) == 0)) {
        if ((rc = make_tmap_bf16(&ta, A, M, K, lda, GEMM_BM))) return rc;
        if ((rc = make_tmap_bf16(&tb, W, N, K, ldw, 128))) return rc;
        return out_dtype == DTLR_OP16 ? launch_tc<128, 4, op16_t>(ta, tb, e, st) : launch_tc<128, 4, float>(ta, tb, e, st);
    }
    if ((rc = make_tmap_bf16(&ta, A, M, K, lda, GEMM_BM))) return rc;
    if ((rc = make_tmap_bf16(&tb, W, N, K, ldw, 64))) return rc;
    return out_dtype == DTLR_OP16 ? launch_tc<64, 6, op16_t>(ta, tb, e, st) : launch_tc<64, 6, float>(ta, tb, e, st);
}

// y = LayerNorm_256(A.W^T + bias (+ residual)) * gamma + beta, optional y2 = y + add2 -- the Linear -> (+residual) -> LayerNorm
// tail of every attention / FFN block (reference deformable_transformer.py:813-814,806-807,906-907,956-957,878-879,326) in ONE
// tcgen05 kernel: the 128 x 256 tile holds whole rows, so the normalisation happens in the epilogue.
extern "C" int dtlr_gemm_ln(const void* A, int lda, const void* W, int ldw, const float* bias, const void* residual, int ldr,
                            const float* gamma, const float* beta, float eps, void* Y, int ldy, const void* add2, void* Y2, int ld2,
                            int M, int K, void* stream) {
    const int N = 256;
    DTLR_CHECK_ARG(M >= 0 && K > 0, "gemm_ln: bad sizes");
    if (M == 0) return DTLR_OK;
    DTLR_CHECK_ARG(A && W && Y && gamma && beta, "gemm_ln: null pointer");
    DTLR_CHECK_ARG(lda >= K && ldw >= K && ldy >= N && (!residual || ldr >= N) && (!Y2 || (add2 && ld2 >= N)), "gemm_ln: bad leading dimension");
    DTLR_CHECK_ARG((lda % 8) == 0 && (ldw % 8) == 0 && (ldy % 8) == 0 && (!residual || (ldr % 8) == 0) && (!Y2 || (ld2 % 8) == 0) &&
                   ((((uintptr_t)A | (uintptr_t)W | (uintptr_t)Y | (uintptr_t)residual | (uintptr_t)add2 | (uintptr_t)Y2)) & 15) == 0,
                   "gemm_ln: operands need 16-byte aligned rows");
    GemmEpi e{bias, residual, Y, ldr, ldy, M, N, K, 0, LnArgs{gamma, beta, add2, Y2, ld2, eps}, ConvGeo{0, 0, 0, 0, 0, 0, 0, 0, 1}, g_debug_flags};
    // K <= 256: weight-stationary kernel; otherwise the tile kernel -- both end in ln_epilogue_tile (TMA stores)
    if (K <= 256 && (long long)((M + GEMM_BM - 1) / GEMM_BM) >= 2ll * sm_count() && !(g_debug_flags & 32))
        return launch_ws<256, op16_t, false, true>(A, lda, W, ldw, e, (cudaStream_t)stream);
    CUtensorMap ta, tb, tc, tc2;
    int rc;
    if ((rc = make_tmap_bf16(&ta, A, M, K, lda, GEMM_BM))) return rc;
    if ((rc = make_tmap_bf16(&tb, W, N, K, ldw, 256))) return rc;
    if ((rc = make_tmap_out<op16_t>(&tc, Y, M, N, ldy))) return rc;
    tc2 = tc;
    if (Y2 && (rc = make_tmap_out<op16_t>(&tc2, Y2, M, N, ld2))) return rc;
    return launch_tc<256, 3, op16_t>(ta, tb, e, (cudaStream_t)stream, &tc, &tc2);
}

// Convolution (stride 1, "same" padding) on NHWC bf16 activations as an implicit GEMM on the tcgen05 kernel above: no im2col
// matrix ever exists; the k x k taps are TMA loads with shifted coordinates and hardware zero fill.
extern "C" int dtlr_conv2d_nhwc_strided(const void* x, const void* w, const float* bias, const void* residual, void* out, int B, int Hin,
                                        int Win, int C, int Cout, int KH, int KW, int pad, int stride, int relu, int out_dtype, void* stream);
extern "C" int dtlr_conv2d_nhwc(const void* x, const void* w, const float* bias, const void* residual, void* out, int B, int H,
                                int W, int C, int Cout, int KH, int KW, int pad, int relu, int out_dtype, void* stream) {
    return dtlr_conv2d_nhwc_strided(x, w, bias, residual, out, B, H, W, C, Cout, KH, KW, pad, 1, relu, out_dtype, stream);
}

// The same implicit GEMM for stride 1 or 2 (ResNet's stride-2 3x3 convs and strided 1x1 downsample convs, torchvision Bottleneck
// v1.5 as built by reference models/dino/backbone.py:118-120): the A tensor map traverses W and H with element stride 2, so the
// producer's per-tap TMA boxes deliver exactly the input pixels of 128 consecutive output pixels -- no im2col matrix (round 1 wrote
// and re-read a 9x larger A for these six convs).  H, W are the INPUT sizes; output (H + 2 pad - KH) / stride + 1.
extern "C" int dtlr_conv2d_nhwc_strided(const void* x, const void* w, const float* bias, const void* residual, void* out, int B, int Hin,
                                        int Win, int C, int Cout, int KH, int KW, int pad, int stride, int relu, int out_dtype, void* stream) {
    DTLR_CHECK_ARG(x && w && out, "conv2d_nhwc: null pointer");
    DTLR_CHECK_ARG(stride == 1 || stride == 2, "conv2d_nhwc: stride must be 1 or 2");
    DTLR_CHECK_ARG(KH == 2 * pad + 1 && KW == 2 * pad + 1, "conv2d_nhwc: only 'same'-padded kernels (k = 2*pad+1)");
    DTLR_CHECK_ARG(C % 64 == 0, "conv2d_nhwc: C must be a multiple of 64 (got %d)", C);
    const int H = (Hin + 2 * pad - KH) / stride + 1, W = (Win + 2 * pad - KW) / stride + 1;       // output size
    int seg_w = W >= 128 ? 128 : W;
    DTLR_CHECK_ARG(seg_w >= 8 && (128 % seg_w) == 0 && (W % seg_w) == 0,
                   "conv2d_nhwc: output width %d cannot be tiled into 128-pixel row segments (use im2col + gemm)", W);
    DTLR_CHECK_ARG((((uintptr_t)x | (uintptr_t)w) & 15) == 0, "conv2d_nhwc: operands must be 16-byte aligned");
    DTLR_CHECK_ARG(out_dtype == DTLR_OP16 || out_dtype == DTLR_F32 || out_dtype == DTLR_SPLIT16, "conv2d_nhwc: output must be 16-bit, f32 or split");
    const bool split_out = out_dtype == DTLR_SPLIT16;             // fp32 result written as [hi | hi | lo], 3 x Cout 16-bit columns per pixel
    const bool f32out = out_dtype == DTLR_F32 || split_out;       // split-precision mode (3C input channels = [hi | hi | lo]); residual fp32 too
    DTLR_CHECK_ARG(!split_out || ((Cout % 8) == 0 && (((uintptr_t)out) & 15) == 0), "conv2d_nhwc: split output needs Cout %% 8 == 0");
    const int M = B * H * W, K = KH * KW * C;
    if (M == 0) return DTLR_OK;
    GemmEpi e{bias, residual, out, Cout, split_out ? 3 * Cout : Cout, M, Cout, K, relu, LnArgs{nullptr, nullptr, nullptr, nullptr, 0, 0.f}, ConvGeo{1, H, W, KW, pad, C / 64, seg_w, 128 / seg_w, stride}, g_debug_flags};
    e.split_out = split_out ? 1 : 0;
    CUtensorMap ta, tb;
    int rc;
    if ((rc = make_tmap_nhwc(&ta, x, B, Hin, Win, C, seg_w, stride))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    // tile width by SM fill: 128 x 256 tiles halve the A traffic per FLOP, but the deep stages have few rows (layer4: M = 4096 = 32
    // row tiles x 2 column tiles of 256 = 64 CTAs on 148 SMs).  cost(BN) = rounds of CTAs x BN; a narrower tile is taken only when it
    // cuts that cost by >= 1.5x (layer4: 256 -> 128 halves it; layer3, 96 CTAs, stays at 256)
    const long long row_tiles = (M + GEMM_BM - 1) / GEMM_BM;
    auto tile_cost = [&](int bn) { const long long ctas = row_tiles * ((Cout + bn - 1) / bn); return ((ctas + sm_count() - 1) / sm_count()) * bn; };
    const bool ok256 = (Cout % 256) == 0 && !f32out, ok128 = Cout > 64;     // (fp32 tiles: 128 / 64 columns, like dtlr_gemm)
    int bn_pick = ok256 ? 256 : (ok128 ? 128 : 64);
    if (!(g_debug_flags & 1048576)) {       // flag 1048576: round-1 rule (widest tile that divides Cout), A/B
        if (bn_pick == 256 && (Cout % 128) == 0 && tile_cost(128) * 3 <= tile_cost(256) * 2) bn_pick = 128;
        if (bn_pick == 128 && (Cout % 64) == 0 && tile_cost(64) * 3 <= tile_cost(128) * 2) bn_pick = 64;
    }
    if (bn_pick == 256) {
        if ((rc = make_tmap_bf16(&tb, w, Cout, K, K, 256))) return rc;
        return launch_tc<256, 3, op16_t>(ta, tb, e, st);
    }
    if (bn_pick == 128) {
        if ((rc = make_tmap_bf16(&tb, w, Cout, K, K, 128))) return rc;
        return f32out ? launch_tc<128, 4, float>(ta, tb, e, st) : launch_tc<128, 4, op16_t>(ta, tb, e, st);
    }
    if ((rc = make_tmap_bf16(&tb, w, Cout, K, K, 64))) return rc;
    return f32out ? launch_tc<64, 6, float>(ta, tb, e, st) : launch_tc<64, 6, op16_t>(ta, tb, e, st);
}
