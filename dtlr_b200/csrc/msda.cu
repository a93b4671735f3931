// dtlr_b200 -- multi-scale deformable attention core for sm_100a.
//
// Replaces the reference op models/dino/ops/src/cuda/ms_deform_im2col_cuda.cuh:237-299 (forward, one thread per
// output scalar, every lane re-reading loc/weight scalars) and :301-403 (backward) behind the C ABI of
// include/dtlr_b200.h.  Design:
//
//  * D == 32 fast path (the only head width DTLR uses: d_model 256 / 8 heads).  One CTA = one (image, head,
//    query range).  The head's value slab -- every level, S pixels x 32 channels -- is staged ONCE in shared
//    memory with 16-byte cp.async (fp32 slab 117 KB, bf16 58 KB at S=912), with one zero pixel of padding on
//    each side, so that all 64 bilinear taps of every query are shared-memory reads.
//  * one warp = one query at a time.  Lane (pt = lane&15, row = lane>>4) computes the tap parameters of ONE
//    (sampling point, y-row) pair: pixel index of the left corner and the two x-corner weights (attention
//    weight, y weight and validity folded in).  The 32 parameter triples are then handed round by shuffles.
//  * gather: a group of 2*LPP lanes (LPP = lanes per pixel = 32*sizeof(T)/16) reads the two x-adjacent corners
//    of one (point,row) as ONE contiguous 2*32*sizeof(T)-byte run with 16-byte loads -> bank-conflict free by
//    construction; fp32 accumulation; butterfly "transpose" reduction leaves one output channel per lane.
//  * slabs that do not fit in shared memory (large S) use the same code with the taps read through L1/L2.
//  * generic path (any D, fp32/fp64): one warp per (b,q,m), lanes over channels.  Used by the fp64 KATs.
#include "tc_common.cuh"

namespace dtlr {

constexpr int MAX_LEVELS = 8;

struct Levels {
    int n;
    int H[MAX_LEVELS], W[MAX_LEVELS], start[MAX_LEVELS];
};

// ------------------------------------------------------------------------------------------------ fast path
// Blackwell packed fp32 FMA (FFMA2): d.xy = a.xy * b.xy + d.xy on 64-bit register pairs.
__device__ __forceinline__ void ffma2(unsigned long long& acc, const unsigned long long a, const unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long pack2(const float lo, const float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ unsigned long long pack2u(const uint32_t lo, const uint32_t hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(const unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

template <typename T>
struct Vec16;  // 16 bytes of T -> NP packed fp32 pairs
template <>
struct Vec16<float> {
    static constexpr int NP = 2;
    __device__ static __forceinline__ void load(const void* p, unsigned long long (&v)[2]) {
        const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(p);
        v[0] = t.x; v[1] = t.y;
    }
};
template <>
struct Vec16<op16_t> {
    static constexpr int NP = 4;
    __device__ static __forceinline__ void load(const void* p, unsigned long long (&v)[4]) {
        const uint4 t = *reinterpret_cast<const uint4*>(p);
        v[0] = pack2u(__float_as_uint(op16_lo_f32(t.x)), __float_as_uint(op16_hi_f32(t.x)));
        v[1] = pack2u(__float_as_uint(op16_lo_f32(t.y)), __float_as_uint(op16_hi_f32(t.y)));
        v[2] = pack2u(__float_as_uint(op16_lo_f32(t.z)), __float_as_uint(op16_hi_f32(t.z)));
        v[3] = pack2u(__float_as_uint(op16_lo_f32(t.w)), __float_as_uint(op16_hi_f32(t.w)));
    }
};

__device__ __forceinline__ void store_out(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_out(op16_t* p, float v) { *p = f32_to_op16(v); }

// FUSED: the kernel also does the MSDeformAttn prologue (reference ops/modules/ms_deform_attn.py:98-108): `loc` points at
// the fp32 projection rows [B*Lq, fz.ld] (M*L*P*2 offsets, then M*L*P attention logits); the softmax over the L*P
// logits of the head is a 16-lane shuffle reduction and the sampling location is ref*valid_ratio + offset-term, so the
// (B,Lq,M,L,P,2) location and (B,Lq,M,L,P) weight tensors never exist in HBM.  Requires SINGLE (L*P <= 16).
struct FusedArgs {
    const float* ref;            // [B*Lq, RD] reference points (RD = 2: encoder centres, 4: decoder boxes)
    const float* valid_ratios;   // [B, L, 2] (w, h)
    int ld;                      // row pitch of the projection matrix (elements)
    int RD;
    int proj_bf16;               // projection rows are bf16 (throughput mode) instead of fp32
};

template <typename T, bool STAGE, bool SINGLE, bool FUSED, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, NWARPS == 16 ? 2 : 1)   // <= 64 registers: 32 resident warps per SM
msda_fwd_d32_kernel(const T* __restrict__ value, const float* __restrict__ loc, const float* __restrict__ attn,
                    T* __restrict__ out, const __grid_constant__ Levels lv, const int S, const int M, const int Lq,
                    const int P, const int q_per_cta, const FusedArgs fz, const int vld /* value elements per pixel row */) {
    constexpr int ROWB = 32 * (int)sizeof(T);   // bytes of one pixel (32 channels)
    constexpr int LPP = ROWB / 16;              // lanes per pixel (8 fp32 / 4 bf16)
    constexpr int GL = 2 * LPP;                 // lanes per (point,row) group: two x-adjacent pixels
    constexpr int G = 32 / GL;                  // groups per warp (2 fp32 / 4 bf16)
    constexpr int ITER = 16 / G;                // gather iterations per y-row of a 16-point chunk
    constexpr int NP = Vec16<T>::NP;            // packed channel pairs per lane
    constexpr int NV = 2 * NP;                  // channels per lane

    extern __shared__ __align__(16) unsigned char smem[];
    const int b = blockIdx.z, m = blockIdx.y;
    const int q0 = blockIdx.x * q_per_cta;
    const int q1 = min(Lq, q0 + q_per_cta);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int LP = lv.n * P;

    const unsigned char* gbase = reinterpret_cast<const unsigned char*>(value) + ((size_t)b * S * vld + (size_t)m * 32) * sizeof(T);
    const size_t gstride = (size_t)vld * sizeof(T);    // bytes between consecutive pixels of one head (vld = M*32 when contiguous)

    if (STAGE) {
        // slab[p+1] = pixel p; slab[0] and slab[S+1] are zero padding (taps with zero weight may land there)
        for (int i = tid; i < S * LPP; i += NWARPS * 32) {
            const int row = i / LPP, ch = i - row * LPP;
            cp_async16(smem + (size_t)(row + 1) * ROWB + ch * 16, gbase + (size_t)row * gstride + ch * 16);
        }
        cp_async_commit();
        if (tid < 2 * LPP) {
            const int row = (tid < LPP) ? 0 : (S + 1);
            *reinterpret_cast<uint4*>(smem + (size_t)row * ROWB + (tid % LPP) * 16) = make_uint4(0, 0, 0, 0);
        }
        cp_async_wait_all();
        __syncthreads();
    }

    const int sub = lane % GL;        // position inside the 2-pixel run
    const int g = lane / GL;          // group id
    const int side = sub / LPP;       // 0 = left corner (x0), 1 = right corner (x0+1)
    // parameter role of this lane: sampling point (lane & 15) of the current 16-point chunk, x-side (lane >> 4).
    // It produces, for BOTH y-rows, the byte offset of the row's left pixel and the weight of its own side.
    const int pt16 = lane & 15, pside = lane >> 4;
    // the lane this lane fetches its taps from in iteration i is (i*G + g) + 16*side
    const int src0 = g + 16 * side;
    const uint32_t lane_off = STAGE ? (uint32_t)(ROWB + sub * 16) : (uint32_t)((sub - side * LPP) * 16);

    // SINGLE: the raw per-query inputs of this lane (FUSED: offset pair, logit, reference box; else location pair and
    // weight) are fetched one query ahead, so their global-memory latency hides behind the previous query's gather.
    float2 pf_a = make_float2(0.f, 0.f);
    float pf_b = 0.f;
    float4 pf_ref = make_float4(0.f, 0.f, 0.f, 0.f);
    auto prefetch = [&](int q) {
        const int pt = min(pt16, LP - 1);
        if (FUSED) {
            const size_t row = (size_t)b * Lq + q;
            if (fz.proj_bf16) {
                const op16_t* pr = reinterpret_cast<const op16_t*>(loc) + row * fz.ld;
                const uint32_t o2 = *reinterpret_cast<const uint32_t*>(pr + ((size_t)m * LP + pt) * 2);
                pf_a = make_float2(op16_lo_f32(o2), op16_hi_f32(o2));
                pf_b = op16_to_f32(pr[(size_t)M * LP * 2 + (size_t)m * LP + pt]);
            } else {
                const float* pr = loc + row * fz.ld;
                pf_a = *reinterpret_cast<const float2*>(pr + ((size_t)m * LP + pt) * 2);
                pf_b = pr[(size_t)M * LP * 2 + (size_t)m * LP + pt];
            }
            if (fz.RD == 4) pf_ref = *reinterpret_cast<const float4*>(fz.ref + row * 4);
            else { const float2 r2 = *reinterpret_cast<const float2*>(fz.ref + row * 2); pf_ref = make_float4(r2.x, r2.y, 0.f, 0.f); }
        } else {
            const size_t pb = ((size_t)((size_t)b * Lq + q) * M + m) * LP;
            pf_a = *reinterpret_cast<const float2*>(loc + (pb + pt) * 2);
            pf_b = attn[pb + pt];
        }
    };
    if (SINGLE && q0 + warp < q1) prefetch(q0 + warp);

    for (int q = q0 + warp; q < q1; q += NWARPS) {
        const size_t pbase = ((size_t)((size_t)b * Lq + q) * M + m) * LP;
        const float2 cur_a = pf_a;
        const float cur_b = pf_b;
        const float4 cur_ref = pf_ref;
        if (SINGLE && q + NWARPS < q1) prefetch(q + NWARPS);
        unsigned long long acc[NP];
#pragma unroll
        for (int k = 0; k < NP; ++k) acc[k] = 0ull;

        // SINGLE: L*P <= 16 (DTLR: 4 levels x 4 points) -> one chunk, level constants hoisted out of the query loop
        for (int c0 = 0; c0 < (SINGLE ? 1 : LP); c0 += 16) {
            // ---- tap parameters of point c0+pt16 (branch-free: out-of-range points get weight 0 and a safe offset)
            const int pt = min(c0 + pt16, LP - 1);
            const bool pt_ok = (c0 + pt16) < LP;
            const int l = pt / P;
            const int H = lv.H[l], W = lv.W[l];
            float2 xy;
            float aw;
            if (FUSED) {
                const float2 off = cur_a;                       // FUSED implies SINGLE: prefetched
                const float lg = pt_ok ? cur_b : -INFINITY;
                float mx = lg;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                const float ex = expf(lg - mx);
                float den = ex;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
                aw = ex * (1.f / den);
                const float rf[4] = {cur_ref.x, cur_ref.y, cur_ref.z, cur_ref.w};
                const float vx = fz.valid_ratios[((size_t)b * lv.n + l) * 2], vy = fz.valid_ratios[((size_t)b * lv.n + l) * 2 + 1];
                const float rx = rf[0] * vx, ry = rf[1] * vy;
                if (fz.RD == 2) {
                    xy.x = rx + off.x / (float)W;
                    xy.y = ry + off.y / (float)H;
                } else {
                    xy.x = rx + off.x / (float)P * (rf[2] * vx) * 0.5f;
                    xy.y = ry + off.y / (float)P * (rf[3] * vy) * 0.5f;
                }
            } else if (SINGLE) {
                xy = cur_a;
                aw = cur_b;
            } else {
                xy = *reinterpret_cast<const float2*>(loc + (pbase + pt) * 2);
                aw = attn[pbase + pt];
            }
            const float y = fmaf(xy.y, (float)H, -0.5f), x = fmaf(xy.x, (float)W, -0.5f);
            const bool inside = pt_ok && y > -1.f && x > -1.f && y < (float)H && x < (float)W;
            const float yf = floorf(y), xf = floorf(x);
            const int y0 = (int)yf, x0 = (int)xf;
            const float fy = y - yf, fx = x - xf;
            // weight of this lane's x-side, zero when that corner column is outside the map
            const bool col_ok = pside ? (x0 + 1 <= W - 1) : (x0 >= 0);
            const float wx = (inside && col_ok) ? (pside ? fx : 1.f - fx) * aw : 0.f;
            const bool r0_ok = inside && y0 >= 0, r1_ok = inside && y0 + 1 <= H - 1;
            const float w_r0 = r0_ok ? wx * (1.f - fy) : 0.f;
            const float w_r1 = r1_ok ? wx * fy : 0.f;
            const int pix0 = lv.start[l] + y0 * W + x0;
            int o_r0, o_r1;   // pixel index (STAGE: scaled to bytes) of the left corner of each row, 0 if the row is unused
            if (STAGE) {
                o_r0 = r0_ok ? pix0 * ROWB : 0;
                o_r1 = r1_ok ? (pix0 + W) * ROWB : 0;
            } else {
                o_r0 = r0_ok ? pix0 : 0;
                o_r1 = r1_ok ? pix0 + W : 0;
            }
            // ---- gather: iteration (r,i), group g consumes point i*G+g, y-row r
#pragma unroll
            for (int r = 0; r < 2; ++r) {
#pragma unroll
                for (int i = 0; i < ITER; ++i) {
                    const int src = src0 + i * G;
                    const int so = __shfl_sync(0xffffffffu, r ? o_r1 : o_r0, src);
                    const float sw = __shfl_sync(0xffffffffu, r ? w_r1 : w_r0, src);
                    const unsigned long long w2 = pack2(sw, sw);
                    unsigned long long v[NP];
                    if (STAGE) {
                        Vec16<T>::load(smem + (uint32_t)so + lane_off, v);
                    } else {
                        const int p = so + side;
                        if (p >= 0 && p < S && sw != 0.f) {
                            Vec16<T>::load(gbase + (size_t)p * gstride + lane_off, v);
                        } else {
#pragma unroll
                            for (int k = 0; k < NP; ++k) v[k] = 0ull;
                        }
                    }
#pragma unroll
                    for (int k = 0; k < NP; ++k) ffma2(acc[k], v[k], w2);
                }
            }
        }

        // ---- reduce over the lanes that hold the same channels (same sub % LPP): butterfly with halving,
        //      ends with exactly one channel per lane.
        float a[NV];
#pragma unroll
        for (int k = 0; k < NP; ++k) unpack2(acc[k], a[2 * k], a[2 * k + 1]);
        int ch = (sub % LPP) * NV;
        if (NV == 8) {
            {   // xor 16: keep 4
                const bool hi = (lane & 16) != 0;
                float keep[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float send = hi ? a[k] : a[k + 4];
                    const float mine = hi ? a[k + 4] : a[k];
                    keep[k] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
                }
                ch += hi ? 4 : 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) a[k] = keep[k];
            }
            {   // xor 8: keep 2
                const bool hi = (lane & 8) != 0;
                float keep[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float send = hi ? a[k] : a[k + 2];
                    const float mine = hi ? a[k + 2] : a[k];
                    keep[k] = mine + __shfl_xor_sync(0xffffffffu, send, 8);
                }
                ch += hi ? 2 : 0;
                a[0] = keep[0]; a[1] = keep[1];
            }
            {   // xor 4: keep 1
                const bool hi = (lane & 4) != 0;
                const float send = hi ? a[0] : a[1];
                const float mine = hi ? a[1] : a[0];
                a[0] = mine + __shfl_xor_sync(0xffffffffu, send, 4);
                ch += hi ? 1 : 0;
            }
        } else {  // NV == 4: lanes sharing (lane & 7): xor 16, xor 8
            {
                const bool hi = (lane & 16) != 0;
                float keep[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float send = hi ? a[k] : a[k + 2];
                    const float mine = hi ? a[k + 2] : a[k];
                    keep[k] = mine + __shfl_xor_sync(0xffffffffu, send, 16);
                }
                ch += hi ? 2 : 0;
                a[0] = keep[0]; a[1] = keep[1];
            }
            {
                const bool hi = (lane & 8) != 0;
                const float send = hi ? a[0] : a[1];
                const float mine = hi ? a[1] : a[0];
                a[0] = mine + __shfl_xor_sync(0xffffffffu, send, 8);
                ch += hi ? 1 : 0;
            }
        }
        store_out(out + ((size_t)((size_t)b * Lq + q) * M + m) * 32 + ch, a[0]);
    }
}

// ------------------------------------------------------------------------------------------------ tensor-core gather path
// bf16 values, D = 32, 4 levels x 4 points (every DTLR config).  The SIMT kernel above is bound by the shared-memory /
// shuffle crossbar (65 wavefronts per (query, head): 32 for the gather itself, 33 for handing tap parameters round and
// reducing) and by instruction issue (331 instructions per (query, head), a third of them bf16 -> fp32 unpacking).  Here:
//
//  * phase 1, lane = query (32 queries of one warp at a time): the whole prologue -- softmax over the 16 logits, sampling
//    locations, bilinear corner weights, validity -- is straight-line per-lane code with NO shuffles; each lane leaves a
//    192-byte tap table in shared memory: 32 u16 slab offsets (point x y-row) and 32 packed bf16 weight pairs.
//  * phase 2, warp = one query at a time: the 16 points are consumed two at a time by ONE ldmatrix.x4.trans + ONE
//    mma.sync.m16n8k16 (bf16 x bf16 -> fp32): the four 8x8 matrices are the four 128-byte runs (2 x-adjacent pixels x 32
//    channels) of {point a, point b} x {row y0, row y1}, each read conflict-free straight from the staged slab as the A
//    operand (M = point slot x channel-in-chunk, K = y-row x x-corner x 16-byte chunk); the B operand holds the corner
//    weights on the chunk diagonal (N = point slot x output chunk), so the tensor core does the unpack, the 64-tap
//    weighted sum and most of the reduction.  Per (query, head): 8 LDSM + 8 HMMA + 3 table loads + 2 shuffles.
constexpr int MMA_NW = 8;             // warps per CTA (two CTAs per SM: 2 x (slab + 8 tap tables))
constexpr int MMA_TSTRIDE = 208;      // bytes per query in a tap table: 64 offsets + 128 weights + 16 pad (the pad makes the
                                      // 16-byte stores of 8 consecutive lanes hit 8 different bank groups)
constexpr int MMA_TAB_BYTES = 32 * MMA_TSTRIDE;
constexpr int MMA_RAW_PITCH = 112;    // staged raw projection row of a query (96 B used): lanes 0-7 hit 8 different 16-byte bank groups

__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const uint32_t saddr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void mma_bf16_m16n8k16(float (&c)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." DTLR_OP16_PTX "." DTLR_OP16_PTX ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t prmt(const uint32_t a, const uint32_t b, const uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
__device__ __forceinline__ uint32_t pack_bf16x2_rn(const float lo, const float hi) {
    op16x2_t t = op16_pack2(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}

// MODE 0: sampling locations / attention weights given (fp32 tensors of the operator boundary)
// MODE 1: fused prologue, fp32 projection rows;  MODE 2: fused prologue, bf16 projection rows
// QU: queries per phase-2 iteration (loop unroll): QU independent ldmatrix -> mma chains in flight per warp; all variants stay
// within the 128-register budget of 2 CTAs per SM
template <int MODE, int QU>
__global__ void __launch_bounds__(MMA_NW * 32, 2)
msda_fwd_mma_kernel(const op16_t* __restrict__ value, const void* __restrict__ loc_or_proj, const float* __restrict__ attn,
                    op16_t* __restrict__ out, const __grid_constant__ Levels lv, const int S, const int M, const int Lq,
                    const int q_per_cta, const FusedArgs fz, const int vld, const __grid_constant__ CUtensorMap tmV, const int tma_rows) {
    constexpr int P = 4, LP = 16;
    extern __shared__ __align__(128) unsigned char smem_base[];
    // the slab starts 64 bytes into the allocation: pixel p lives at smem + (p + 1) * 64, so the first DATA row sits on a 128-byte
    // boundary, which is what a TMA box destination needs (tma_rows > 0: the slab is staged by cp.async.bulk.tensor)
    unsigned char* const smem = smem_base + 64;
    __shared__ __align__(8) uint64_t slab_bar;
    const int b = blockIdx.z, m = blockIdx.y;
    const int q0 = blockIdx.x * q_per_cta;
    const int q1 = min(Lq, q0 + q_per_cta);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_launch_dependents();
    pdl_wait();

    unsigned char* tab = smem + (size_t)(S + 2) * 64 + (size_t)warp * MMA_TAB_BYTES;
    const int per_warp = (q1 - q0 + MMA_NW - 1) / MMA_NW;
    const int wq0 = q0 + warp * per_warp;
    const int wq1 = min(q1, wq0 + per_warp);
    // MODE 2: the raw projection rows of a 32-query batch (64 B of offsets + 32 B of logits per query, 768 B apart in HBM) are
    // fetched by cp.async with consecutive lanes on consecutive 16-byte chunks (6-11 lines per instruction instead of 32) into the
    // warp's table region (112-byte pitch: conflict-free 16-byte reads by lane = query); the first batch flies behind the slab copy
    auto stage_raw = [&](int qb) {
        const op16_t* pbase = reinterpret_cast<const op16_t*>(loc_or_proj);
#pragma unroll
        for (int u = 0; u < 6; ++u) {
            const int c = lane + 32 * u;
            const int qi = c / 6, part = c - qi * 6;
            const int q = min(qb + qi, wq1 - 1);
            const op16_t* src = pbase + ((size_t)b * Lq + q) * fz.ld + (part < 4 ? m * (LP * 2) + part * 8 : M * (LP * 2) + m * LP + (part - 4) * 8);
            cp_async16(tab + qi * MMA_RAW_PITCH + part * 16, src);
        }
        cp_async_commit();
    };
    if (MODE == 2 && wq0 < wq1) stage_raw(wq0);

    // ---- stage the head's value slab: slab pixel p+1 = pixel p, pixels 0 and S+1 are zero padding
    if (tma_rows > 0) {
        // TMA: S / tma_rows boxes of (32 channels = 64 bytes) x tma_rows pixels from the (B*S, vld) value matrix, issued by one thread,
        // completion on an mbarrier -- no per-thread address arithmetic, no register staging (the cp.async path issues S*4 copies)
        if (tid == 0) {
            mbar_init(&slab_bar, 1);
            fence_barrier_init();
            mbar_expect_tx(&slab_bar, (uint32_t)S * 64u);
            for (int r0 = 0; r0 < S; r0 += tma_rows) tma_load_2d(smem + (size_t)(r0 + 1) * 64, &tmV, &slab_bar, m * 32, b * S + r0);
        }
        cp_async_commit();          // keeps the group numbering of the MODE 2 raw-row prefetch (an empty group completes at once)
        if (tid < 8) {
            const int row = (tid < 4) ? 0 : (S + 1);
            *reinterpret_cast<uint4*>(smem + (size_t)row * 64 + (tid & 3) * 16) = make_uint4(0, 0, 0, 0);
        }
    } else {
        const unsigned char* gbase = reinterpret_cast<const unsigned char*>(value) + ((size_t)b * S * vld + (size_t)m * 32) * 2;
        const size_t gstride = (size_t)vld * 2;
        for (int i = tid; i < S * 4; i += MMA_NW * 32) {
            const int row = i >> 2, ch = i & 3;
            cp_async16(smem + (size_t)(row + 1) * 64 + ch * 16, gbase + (size_t)row * gstride + ch * 16);
        }
        cp_async_commit();
        if (tid < 8) {
            const int row = (tid < 4) ? 0 : (S + 1);
            *reinterpret_cast<uint4*>(smem + (size_t)row * 64 + (tid & 3) * 16) = make_uint4(0, 0, 0, 0);
        }
    }
    // phase-2 lane roles
    const int g = lane >> 2, j = lane & 3;
    const int mi = lane >> 3;                                        // ldmatrix: row (lane & 7) of matrix mi = 2*yrow + slot
    const uint32_t slab_lane = (uint32_t)__cvta_generic_to_shared(smem) + (lane & 7) * 16;
    const uint32_t tab_o = mi * 16;                                   // this lane's 8 offsets (one per ldmatrix)
    const uint32_t tab_w = 64 + ((g >> 2) * 2 + (j >> 1)) * 32;       // 8 weight pairs of (slot g>>2, x-corner j>>1)
    // B fragment: element e of b0/b1 is B[k = 2j+e (+8)][n = g]; k -> (x-corner j>>1, chunk 2(j&1)+e), n -> (slot g>>2, chunk g&3)
    const bool contrib = (j & 1) == ((g >> 1) & 1);
    const uint32_t sel0 = contrib ? ((g & 1) ? 0x1044u : 0x4410u) : 0x4444u;   // y-row 0 weight (low half of the pair)
    const uint32_t sel1 = contrib ? ((g & 1) ? 0x3244u : 0x4432u) : 0x4444u;   // y-row 1 weight (high half)

    // every warp passes the block barrier exactly once (after its first prologue, which overlaps the slab copy); a warp whose
    // query range is empty only waits for its copies
    int qb = wq0;
    bool first = true;
    while (true) {
        const bool have = qb < wq1;
        // =============================== phase 1: lane = query qb + lane
        if (have) {
            const int q = min(qb + lane, wq1 - 1);
            const size_t row = (size_t)b * Lq + q;
            float ox[LP], oy[LP], aw[LP];
            if (MODE == 0) {
                const float4* pl = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(loc_or_proj) + (row * M + m) * (LP * 2));
                const float4* pa = reinterpret_cast<const float4*>(attn + (row * M + m) * LP);
#pragma unroll
                for (int k = 0; k < 8; ++k) { const float4 t = __ldg(pl + k); ox[2 * k] = t.x; oy[2 * k] = t.y; ox[2 * k + 1] = t.z; oy[2 * k + 1] = t.w; }
#pragma unroll
                for (int k = 0; k < 4; ++k) { const float4 t = __ldg(pa + k); aw[4 * k] = t.x; aw[4 * k + 1] = t.y; aw[4 * k + 2] = t.z; aw[4 * k + 3] = t.w; }
            } else if (MODE == 1) {
                const float* pr = reinterpret_cast<const float*>(loc_or_proj) + row * fz.ld;
                const float4* pl = reinterpret_cast<const float4*>(pr + m * (LP * 2));
                const float4* pa = reinterpret_cast<const float4*>(pr + M * (LP * 2) + m * LP);
#pragma unroll
                for (int k = 0; k < 8; ++k) { const float4 t = __ldg(pl + k); ox[2 * k] = t.x; oy[2 * k] = t.y; ox[2 * k + 1] = t.z; oy[2 * k + 1] = t.w; }
#pragma unroll
                for (int k = 0; k < 4; ++k) { const float4 t = __ldg(pa + k); aw[4 * k] = t.x; aw[4 * k + 1] = t.y; aw[4 * k + 2] = t.z; aw[4 * k + 3] = t.w; }
            } else {
                if (!first) stage_raw(qb);                               // (the first batch was requested at kernel start)
                if (first) cp_async_wait_group<1>(); else cp_async_wait_all();   // group order: raw rows, then the slab
                __syncwarp();
                const uint4* pl = reinterpret_cast<const uint4*>(tab + lane * MMA_RAW_PITCH);
                const uint4* pa = pl + 4;
                uint4 raw[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) raw[k] = pl[k];
                (void)pa;
                __syncwarp();                                            // every lane holds its row: the region becomes the tap table
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint4 t = raw[k];
                    const uint32_t w4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) { ox[4 * k + i] = op16_lo_f32(w4[i]); oy[4 * k + i] = op16_hi_f32(w4[i]); }
                }
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const uint4 t = raw[4 + k];
                    const uint32_t w4[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) { aw[8 * k + 2 * i] = op16_lo_f32(w4[i]); aw[8 * k + 2 * i + 1] = op16_hi_f32(w4[i]); }
                }
            }
            float4 rf = make_float4(0.f, 0.f, 0.f, 0.f);
            if (MODE != 0) {
                // softmax over the head's 16 logits (reference ops/modules/ms_deform_attn.py:100)
                float mx = aw[0];
#pragma unroll
                for (int p = 1; p < LP; ++p) mx = fmaxf(mx, aw[p]);
                float den = 0.f;
#pragma unroll
                for (int p = 0; p < LP; ++p) { aw[p] = __expf(aw[p] - mx); den += aw[p]; }
                const float inv = 1.f / den;
#pragma unroll
                for (int p = 0; p < LP; ++p) aw[p] *= inv;
                if (fz.RD == 4) rf = __ldg(reinterpret_cast<const float4*>(fz.ref + row * 4));
                else { const float2 r2 = __ldg(reinterpret_cast<const float2*>(fz.ref + row * 2)); rf = make_float4(r2.x, r2.y, 0.f, 0.f); }
            }
            // pixel-space coordinate of a point of level l: x = ox * ax[l] + cx[l]  (= loc_x * W - 0.5 with the location
            // arithmetic of reference ops/modules/ms_deform_attn.py:102-108 folded into one FMA; bf16 mode, not bit-exact)
            float ax[4], ay[4], cx[4], cy[4];
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                const float Hf = (float)lv.H[l], Wf = (float)lv.W[l];
                if (MODE == 0) {
                    ax[l] = Wf; ay[l] = Hf; cx[l] = -0.5f; cy[l] = -0.5f;
                } else {
                    const float vx = fz.valid_ratios[((size_t)b * 4 + l) * 2], vy = fz.valid_ratios[((size_t)b * 4 + l) * 2 + 1];
                    cx[l] = fmaf(rf.x * vx, Wf, -0.5f);
                    cy[l] = fmaf(rf.y * vy, Hf, -0.5f);
                    ax[l] = fz.RD == 2 ? 1.f : 0.125f * (rf.z * vx) * Wf;      // off / W * W  |  off / P * w * 0.5 * W
                    ay[l] = fz.RD == 2 ? 1.f : 0.125f * (rf.w * vy) * Hf;
                }
            }
            uint32_t opk[LP], wp0[LP], wp1[LP];     // per point: offsets (row0 | row1 << 16), weight pairs of x-corner 0 / 1
#pragma unroll
            for (int p = 0; p < LP; ++p) {
                const int l = p / P;
                const int H = lv.H[l], W = lv.W[l];
                const float Hf = (float)H, Wf = (float)W;
                const float y = fmaf(oy[p], ay[l], cy[l]), x = fmaf(ox[p], ax[l], cx[l]);
                const bool inside = y > -1.f && x > -1.f && y < Hf && x < Wf;
                const float yf = floorf(y), xf = floorf(x);
                const int y0 = (int)yf, x0 = (int)xf;
                const float fy = y - yf, fx = x - xf;
                const float wx0 = (inside && x0 >= 0) ? (1.f - fx) * aw[p] : 0.f;          // left corner column inside the map
                const float wx1 = (inside && x0 + 1 <= W - 1) ? fx * aw[p] : 0.f;         // right corner column inside the map
                const bool r0_ok = inside && y0 >= 0, r1_ok = inside && y0 + 1 <= H - 1;
                const float gy0 = r0_ok ? 1.f - fy : 0.f, gy1 = r1_ok ? fy : 0.f;
                wp0[p] = pack_bf16x2_rn(wx0 * gy0, wx0 * gy1);
                wp1[p] = pack_bf16x2_rn(wx1 * gy0, wx1 * gy1);
                const int pix0 = lv.start[l] + y0 * W + x0;                  // left pixel of row y0 (>= -1 when the row is used)
                const uint32_t o0 = r0_ok ? (uint32_t)(pix0 + 1) * 64u : 0u;
                const uint32_t o1 = r1_ok ? (uint32_t)(pix0 + W + 1) * 64u : 0u;
                opk[p] = o0 | (o1 << 16);
            }
            unsigned char* t = tab + lane * MMA_TSTRIDE;
            // offsets of ldmatrix role mi = 2*yrow + slot: half-word `it` = point 2*it + slot, row yrow
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int s = r & 1;
                const uint32_t sel = (r >> 1) ? 0x7632u : 0x5410u;
                uint4 v;
                v.x = prmt(opk[0 + s], opk[2 + s], sel);
                v.y = prmt(opk[4 + s], opk[6 + s], sel);
                v.z = prmt(opk[8 + s], opk[10 + s], sel);
                v.w = prmt(opk[12 + s], opk[14 + s], sel);
                *reinterpret_cast<uint4*>(t + r * 16) = v;
            }
            // weight pairs of role 2*slot + x-corner: word `it` = point 2*it + slot
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                *reinterpret_cast<uint4*>(t + 64 + (2 * s) * 32) = make_uint4(wp0[0 + s], wp0[2 + s], wp0[4 + s], wp0[6 + s]);
                *reinterpret_cast<uint4*>(t + 64 + (2 * s) * 32 + 16) = make_uint4(wp0[8 + s], wp0[10 + s], wp0[12 + s], wp0[14 + s]);
                *reinterpret_cast<uint4*>(t + 64 + (2 * s + 1) * 32) = make_uint4(wp1[0 + s], wp1[2 + s], wp1[4 + s], wp1[6 + s]);
                *reinterpret_cast<uint4*>(t + 64 + (2 * s + 1) * 32 + 16) = make_uint4(wp1[8 + s], wp1[10 + s], wp1[12 + s], wp1[14 + s]);
            }
        }
        if (first) {
            cp_async_wait_all();
            __syncthreads();                                     // (also: every thread now sees the initialised slab barrier)
            if (tma_rows > 0) mbar_wait(&slab_bar, 0);           // the TMA boxes of the slab have landed
            first = false;
        } else {
            __syncwarp();
        }
        if (!have) break;

        // =============================== phase 2: one query at a time, 8 x (ldmatrix.x4.trans + mma)
        const int nq = min(32, wq1 - qb);
#pragma unroll (QU)
        for (int jq = 0; jq < nq; ++jq) {
            const unsigned char* t = tab + jq * MMA_TSTRIDE;
            const uint4 o4 = *reinterpret_cast<const uint4*>(t + tab_o);
            const uint4 wa = *reinterpret_cast<const uint4*>(t + tab_w);
            const uint4 wb = *reinterpret_cast<const uint4*>(t + tab_w + 16);
            const uint32_t ov[4] = {o4.x, o4.y, o4.z, o4.w};
            const uint32_t wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
            float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const uint32_t off = (it & 1) ? (ov[it >> 1] >> 16) : (ov[it >> 1] & 0xffffu);
                uint32_t a[4];
                ldsm_x4_trans(a, slab_lane + off);
                const uint32_t b0 = prmt(wv[it], 0u, sel0), b1 = prmt(wv[it], 0u, sel1);
                if (it & 1) mma_bf16_m16n8k16(c1, a, b0, b1);
                else mma_bf16_m16n8k16(c0, a, b0, b1);
            }
            // D rows 0-7 = slot 0 (useful columns n < 4: lanes j < 2), rows 8-15 = slot 1 (columns n >= 4: lanes j >= 2)
            float r0 = c0[0] + c1[0], r1 = c0[1] + c1[1];
            const float r2 = c0[2] + c1[2], r3 = c0[3] + c1[3];
            r0 += __shfl_down_sync(0xffffffffu, r2, 2);
            r1 += __shfl_down_sync(0xffffffffu, r3, 2);
            if (j < 2) {
                op16_t* o = out + ((size_t)((size_t)b * Lq + qb + jq) * M + m) * 32 + g;
                o[(2 * j) * 8] = f32_to_op16(r0);
                o[(2 * j + 1) * 8] = f32_to_op16(r1);
            }
        }
        __syncwarp();                // the table is rewritten by the next batch
        qb += 32;
        if (qb >= wq1) break;
    }
}

// ------------------------------------------------------------------------------------------------ generic path
template <typename T>
__device__ __forceinline__ T ld_as(const T* p) { return *p; }

template <typename T>
__global__ void msda_fwd_generic_kernel(const T* __restrict__ value, const T* __restrict__ loc,
                                        const T* __restrict__ attn, T* __restrict__ out, const __grid_constant__ Levels lv, const int B,
                                        const int S, const int M, const int D, const int Lq, const int P) {
    const int lane = threadIdx.x & 31;
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long total = (long long)B * Lq * M;
    if (wid >= total) return;
    const int m = (int)(wid % M);
    const long long bq = wid / M;
    const int b = (int)(bq / Lq);
    const int LP = lv.n * P;
    const T* vb = value + (size_t)b * S * M * D + (size_t)m * D;
    for (int c = lane; c < D; c += 32) {
        T acc = 0;
        for (int pt = 0; pt < LP; ++pt) {
            const int l = pt / P;
            const int H = lv.H[l], W = lv.W[l];
            const T lx = loc[((size_t)wid * LP + pt) * 2], ly = loc[((size_t)wid * LP + pt) * 2 + 1];
            const T aw = attn[(size_t)wid * LP + pt];
            const T y = ly * H - (T)0.5, x = lx * W - (T)0.5;
            if (!(y > -1 && x > -1 && y < H && x < W)) continue;
            const int y0 = (int)floor((double)y), x0 = (int)floor((double)x);
            const T fy = y - y0, fx = x - x0;
            T val = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int yy = y0 + (k >> 1), xx = x0 + (k & 1);
                if (yy < 0 || yy > H - 1 || xx < 0 || xx > W - 1) continue;
                const T cw = ((k >> 1) ? fy : 1 - fy) * ((k & 1) ? fx : 1 - fx);
                val += cw * vb[(size_t)(lv.start[l] + yy * W + xx) * M * D + c];
            }
            acc += aw * val;
        }
        out[(size_t)wid * D + c] = acc;
    }
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// backward: one warp per (b,q,m); lanes over channels; grad_value scattered with atomics (as the reference does,
// ms_deform_im2col_cuda.cuh:125-152), grad_loc / grad_attn reduced with shuffles instead of shared memory + a
// serial thread-0 sum (reference :377-393).
template <typename T>
__global__ void msda_bwd_kernel(const T* __restrict__ value, const T* __restrict__ loc, const T* __restrict__ attn,
                                const T* __restrict__ gout, T* __restrict__ gvalue, T* __restrict__ gloc,
                                T* __restrict__ gattn, const __grid_constant__ Levels lv, const int B, const int S, const int M,
                                const int D, const int Lq, const int P) {
    const int lane = threadIdx.x & 31;
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long total = (long long)B * Lq * M;
    if (wid >= total) return;
    const int m = (int)(wid % M);
    const int b = (int)((wid / M) / Lq);
    const int LP = lv.n * P;
    const size_t voff = (size_t)b * S * M * D + (size_t)m * D;
    const T* go = gout + (size_t)wid * D;
    for (int pt = 0; pt < LP; ++pt) {
        const int l = pt / P;
        const int H = lv.H[l], W = lv.W[l];
        const size_t ip = (size_t)wid * LP + pt;
        const T y = loc[ip * 2 + 1] * H - (T)0.5, x = loc[ip * 2] * W - (T)0.5;
        const T aw = attn[ip];
        T acc_w = 0, acc_x = 0, acc_y = 0;
        if (y > -1 && x > -1 && y < H && x < W) {
            const int y0 = (int)floor((double)y), x0 = (int)floor((double)x);
            const T fy = y - y0, fx = x - x0, gy = 1 - fy, gx = 1 - fx;
            for (int c = lane; c < D; c += 32) {
                const T g = go[c], ga = g * aw;
                T val = 0, ddy = 0, ddx = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int yy = y0 + (k >> 1), xx = x0 + (k & 1);
                    if (yy < 0 || yy > H - 1 || xx < 0 || xx > W - 1) continue;
                    const T cw = ((k >> 1) ? fy : gy) * ((k & 1) ? fx : gx);
                    const T dy = ((k >> 1) ? (T)1 : (T)-1) * ((k & 1) ? fx : gx);
                    const T dx = ((k & 1) ? (T)1 : (T)-1) * ((k >> 1) ? fy : gy);
                    const size_t idx = voff + (size_t)(lv.start[l] + yy * W + xx) * M * D + c;
                    const T v = value[idx];
                    val += cw * v; ddy += dy * v; ddx += dx * v;
                    atomicAdd(gvalue + idx, cw * ga);
                }
                acc_w += g * val;
                acc_x += (T)W * ddx * ga;
                acc_y += (T)H * ddy * ga;
            }
        }
        acc_w = warp_sum(acc_w); acc_x = warp_sum(acc_x); acc_y = warp_sum(acc_y);
        if (lane == 0) { gattn[ip] = acc_w; gloc[ip * 2] = acc_x; gloc[ip * 2 + 1] = acc_y; }
    }
}

// backward fast path: D = 32, fp32, L*P = 16 (every shipped DTLR config).  One warp per (b,q,m), no shared memory:
//  * lane pt (< 16) computes the tap parameters of sampling point pt ONCE (the generic kernel above recomputes them in all
//    32 lanes for every point): biased pixel index of the (y0,x0) corner, fractional weights, the four corner-validity bits;
//    4 shuffles per point hand them to the warp.
//  * lane = (corner k = lane>>3, channel group j = lane&7): ONE 16-byte load fetches the four corners of a point (4 x 128 B,
//    coalesced), d = <grad_out[4j..4j+3], v> is the only per-channel arithmetic, and ONE `red.global.add.v4.f32`
//    (REDG.E.ADD.F32x4) scatters the point's grad_value contribution: 16 vector reductions per (q,m) instead of 64 scalar
//    warp-wide atomics (the reference scatters scalar atomicAdd, ms_deform_im2col_cuda.cuh:125-152).
//  * the 48 partial sums (16 points x {grad_attn, grad_x, grad_y}) stay in registers and are reduced over the 32 lanes by a
//    transposed butterfly (24+12+6+3+3 = 48 shuffles instead of 48 x 5); it leaves point p's three sums in lanes 2p, 2p+1,
//    which store grad_attn / grad_loc coalesced (the reference: shared memory + a serial thread-0 sum, :377-393).
template <int NPT>
__device__ __forceinline__ void transposed_reduce(float (&r)[3 * NPT], const int lane) {
    static_assert(NPT == 16, "layout below assumes 48 partials over 32 lanes");
    // after the step with offset o, the lanes with bit o clear own the first half of the live values, the others the second
#pragma unroll
    for (int i = 0; i < 24; ++i) {
        const bool hi = lane & 16;
        const float send = hi ? r[i] : r[i + 24], keep = hi ? r[i + 24] : r[i];
        r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const bool hi = lane & 8;
        const float send = hi ? r[i] : r[i + 12], keep = hi ? r[i + 12] : r[i];
        r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const bool hi = lane & 4;
        const float send = hi ? r[i] : r[i + 6], keep = hi ? r[i + 6] : r[i];
        r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const bool hi = lane & 2;
        const float send = hi ? r[i] : r[i + 3], keep = hi ? r[i + 3] : r[i];
        r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) r[i] += __shfl_xor_sync(0xffffffffu, r[i], 1);
}

__device__ __forceinline__ void red_add_f32x4(float* p, const float a, const float b, const float c, const float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// same reduction without the "memory" clobber: grad_value never aliases the buffers this kernel loads from, so the compiler may
// keep later loads in flight across it (used by the GROUPED variant only)
__device__ __forceinline__ void red_add_f32x4_nc(float* p, const float a, const float b, const float c, const float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d));
}

// GROUPED (default for P = 4 since round 2; dtlr_debug_flags(65536) = the variant above; measured 433 us vs 594 us at B=32, Lq=900 -- written from the ncu reading of the default variant,
// profiles/r1_msda_bwd_ncu.txt: 16 warps per SM, one value load in flight per warp because the clobbered `red` asm pins the
// next point's load behind it): the four value loads of a group of 4 points are issued before any of their arithmetic and the
// reductions carry no memory clobber, so a warp keeps 4 L2 round trips in flight instead of 1.  Arithmetic per point is identical.
template <int PTS, int GROUP>      // PTS: sampling points per level (L * PTS == 16); GROUP: value loads in flight (0: one)
__device__ __forceinline__ void msda_bwd_d32_body(const float* __restrict__ value, const float* __restrict__ loc,
                                                  const float* __restrict__ attn, const float* __restrict__ gout,
                                                  float* __restrict__ gvalue, float* __restrict__ gloc, float* __restrict__ gattn,
                                                  const Levels& lv, const long long total, const int S, const int M, const int Lq) {
    const int lane = threadIdx.x & 31;
    const int k = lane >> 3, j = lane & 7;
    const bool ky = k >> 1, kx = k & 1;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    // the tap parameters of lane pt's level are loop-invariant
    const int pl = (lane & 15) / PTS;
    const int pH = lv.H[pl], pW = lv.W[pl], pstart = lv.start[pl];
    for (long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wid < total; wid += nwarps) {
        const int m = (int)(wid % M);
        const int b = (int)((wid / M) / Lq);
        const size_t voff = ((size_t)b * S * M + m) * 32 + 4 * j;
        const float4 g = __ldg(reinterpret_cast<const float4*>(gout + (size_t)wid * 32) + j);
        // ---- lane pt: parameters of point pt (lanes 16-31 mirror 0-15; only their shuffled-from copies in 0-15 are used)
        const size_t ip = (size_t)wid * 16 + (lane & 15);
        const float2 lxy = __ldg(reinterpret_cast<const float2*>(loc) + ip);
        const float paw = __ldg(attn + ip);
        const float y = lxy.y * pH - 0.5f, x = lxy.x * pW - 0.5f;
        const bool inside = y > -1.f && x > -1.f && y < pH && x < pW;
        const float fy0 = floorf(y), fx0 = floorf(x);
        const int y0 = (int)fy0, x0 = (int)fx0;
        const float pfy = y - fy0, pfx = x - fx0;
        unsigned ppix = 0;
        if (inside) {
            const unsigned okY0 = y0 >= 0, okY1 = y0 + 1 <= pH - 1, okX0 = x0 >= 0, okX1 = x0 + 1 <= pW - 1;
            const unsigned bits = (okY0 & okX0) | ((okY0 & okX1) << 1) | ((okY1 & okX0) << 2) | ((okY1 & okX1) << 3);
            // pixel index of corner (y0,x0), biased by one row + one pixel so that it is never negative; bits 24-27: validity
            ppix = (unsigned)(pstart + (y0 + 1) * pW + (x0 + 1)) | (bits << 24);
        }
        float r[48];
        if (GROUP > 0) {
            constexpr int G = GROUP > 0 ? GROUP : 1;
#pragma unroll
            for (int g0 = 0; g0 < 16; g0 += G) {
                float4 v[G];
                float wy[G], wx[G], aw[G];
                int off[G];                              // element offset of this lane's 4 channels of its corner; -1: corner not on the map
#pragma unroll
                for (int i = 0; i < G; ++i) {
                    const int pt = g0 + i;
                    const unsigned pix = __shfl_sync(0xffffffffu, ppix, pt);
                    const float fy = __shfl_sync(0xffffffffu, pfy, pt);
                    const float fx = __shfl_sync(0xffffffffu, pfx, pt);
                    aw[i] = __shfl_sync(0xffffffffu, paw, pt);
                    const int W = lv.W[pt / PTS];
                    wy[i] = ky ? fy : 1.f - fy;
                    wx[i] = kx ? fx : 1.f - fx;
                    const bool ok = (pix >> (24 + k)) & 1u;
                    const int pixel = (int)(pix & 0xffffffu) - W - 1 + (ky ? W : 0) + (kx ? 1 : 0);
                    off[i] = ok ? pixel * M * 32 : -1;
                    v[i] = ok ? __ldg(reinterpret_cast<const float4*>(value + voff + (size_t)off[i])) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int i = 0; i < G; ++i) {
                    const int pt = g0 + i;
                    const int W = lv.W[pt / PTS], H = lv.H[pt / PTS];
                    const float d = g.x * v[i].x + g.y * v[i].y + g.z * v[i].z + g.w * v[i].w;      // 0 for an off-map corner
                    if (off[i] >= 0) {
                        const float c = wy[i] * wx[i] * aw[i];
                        red_add_f32x4_nc(gvalue + voff + (size_t)off[i], c * g.x, c * g.y, c * g.z, c * g.w);
                    }
                    r[3 * pt] = wy[i] * wx[i] * d;
                    r[3 * pt + 1] = (kx ? wy[i] : -wy[i]) * d * aw[i] * (float)W;
                    r[3 * pt + 2] = (ky ? wx[i] : -wx[i]) * d * aw[i] * (float)H;
                }
            }
        } else
#pragma unroll
        for (int pt = 0; pt < 16; ++pt) {
            const unsigned pix = __shfl_sync(0xffffffffu, ppix, pt);
            const float fy = __shfl_sync(0xffffffffu, pfy, pt);
            const float fx = __shfl_sync(0xffffffffu, pfx, pt);
            const float aw = __shfl_sync(0xffffffffu, paw, pt);
            const int W = lv.W[pt / PTS], H = lv.H[pt / PTS];
            const float wy = ky ? fy : 1.f - fy, wx = kx ? fx : 1.f - fx;
            float d = 0.f;
            if ((pix >> (24 + k)) & 1u) {
                const int pixel = (int)(pix & 0xffffffu) - W - 1 + (ky ? W : 0) + (kx ? 1 : 0);
                const size_t idx = voff + (size_t)pixel * M * 32;
                const float4 v = __ldg(reinterpret_cast<const float4*>(value + idx));
                d = g.x * v.x + g.y * v.y + g.z * v.z + g.w * v.w;
                const float c = wy * wx * aw;
                red_add_f32x4(gvalue + idx, c * g.x, c * g.y, c * g.z, c * g.w);
            }
            r[3 * pt] = wy * wx * d;                               // -> grad_attn
            r[3 * pt + 1] = (kx ? wy : -wy) * d * aw * (float)W;   // -> grad_loc.x   (d bilinear / dx = +-wy)
            r[3 * pt + 2] = (ky ? wx : -wx) * d * aw * (float)H;   // -> grad_loc.y
        }
        transposed_reduce<16>(r, lane);
        if (!(lane & 1)) {
            const size_t op = (size_t)wid * 16 + (lane >> 1);
            gattn[op] = r[0];
            *reinterpret_cast<float2*>(gloc + 2 * op) = make_float2(r[1], r[2]);
        }
    }
}

template <int PTS>
__global__ void __launch_bounds__(256)
msda_bwd_d32_kernel(const float* __restrict__ value, const float* __restrict__ loc, const float* __restrict__ attn,
                    const float* __restrict__ gout, float* __restrict__ gvalue, float* __restrict__ gloc,
                    float* __restrict__ gattn, const __grid_constant__ Levels lv, const long long total, const int S,
                    const int M, const int Lq) {
    msda_bwd_d32_body<PTS, 0>(value, loc, attn, gout, gvalue, gloc, gattn, lv, total, S, M, Lq);
}

template <int PTS, int GROUP>
__global__ void __launch_bounds__(GROUP > 4 ? 128 : 256, GROUP > 4 ? 3 : 2)
msda_bwd_d32_grouped_kernel(const float* __restrict__ value, const float* __restrict__ loc, const float* __restrict__ attn,
                            const float* __restrict__ gout, float* __restrict__ gvalue, float* __restrict__ gloc,
                            float* __restrict__ gattn, const __grid_constant__ Levels lv, const long long total, const int S,
                            const int M, const int Lq) {
    msda_bwd_d32_body<PTS, GROUP>(value, loc, attn, gout, gvalue, gloc, gattn, lv, total, S, M, Lq);
}

// ------------------------------------------------------------------------------------------------ host side
static int fill_levels(Levels& lv, const int64_t* shapes, const int64_t* lsi, int L, int S) {
    DTLR_CHECK_ARG(L >= 1 && L <= MAX_LEVELS, "msda: n_levels %d not in [1,%d]", L, MAX_LEVELS);
    long long tot = 0;
    lv.n = L;
    for (int l = 0; l < L; ++l) {
        lv.H[l] = (int)shapes[2 * l];
        lv.W[l] = (int)shapes[2 * l + 1];
        lv.start[l] = (int)lsi[l];
        DTLR_CHECK_ARG(lv.H[l] > 0 && lv.W[l] > 0, "msda: level %d has empty shape", l);
        DTLR_CHECK_ARG(lv.start[l] == tot, "msda: level_start_index[%d]=%d, expected %lld", l, lv.start[l], tot);
        tot += (long long)lv.H[l] * lv.W[l];
    }
    DTLR_CHECK_ARG(tot == S, "msda: sum(H*W)=%lld != S=%d", tot, S);
    return DTLR_OK;
}

template <typename T, int NW>
static int launch_fwd_d32_nw(const void* value, const void* loc, const void* attn, void* out, const Levels& lv, int B,
                             int S, int M, int Lq, int P, bool stage, size_t slab, int occ, const FusedArgs* fzp,
                             cudaStream_t st, int vld) {
    const bool fused = fzp != nullptr;
    const FusedArgs fz = fused ? *fzp : FusedArgs{nullptr, nullptr, 0, 0, 0};
    const long long slots = (long long)sm_count() * occ;
    // split the query range so that the grid is several waves deep but every CTA keeps >= 64 queries
    int qsplit = (int)((4 * slots + (long long)B * M - 1) / ((long long)B * M));
    const int max_split = (Lq + 63) / 64;
    if (qsplit > max_split) qsplit = max_split;
    if (qsplit < 1) qsplit = 1;
    const int q_per_cta = (Lq + qsplit - 1) / qsplit;
    qsplit = (Lq + q_per_cta - 1) / q_per_cta;
    dim3 grid(qsplit, M, B), block(NW * 32);
    const bool single = lv.n * P <= 16;
    if (fused && !single) {
        set_error("msda fused prologue needs n_levels*n_points <= 16");
        return DTLR_ERR_UNSUPPORTED;
    }
    if (stage) {
        auto k = fused ? msda_fwd_d32_kernel<T, true, true, true, NW>
                       : (single ? msda_fwd_d32_kernel<T, true, true, false, NW> : msda_fwd_d32_kernel<T, true, false, false, NW>);
        DTLR_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)slab));
        k<<<grid, block, slab, st>>>((const T*)value, (const float*)loc, (const float*)attn, (T*)out, lv, S, M, Lq, P,
                                     q_per_cta, fz, vld);
    } else {
        auto k = fused ? msda_fwd_d32_kernel<T, false, true, true, NW>
                       : (single ? msda_fwd_d32_kernel<T, false, true, false, NW> : msda_fwd_d32_kernel<T, false, false, false, NW>);
        k<<<grid, block, 0, st>>>((const T*)value, (const float*)loc, (const float*)attn, (T*)out, lv, S, M, Lq, P,
                                  q_per_cta, fz, vld);
    }
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

// tensor-core gather kernel: bf16 values, 4 levels x 4 points, slab + tap tables resident in shared memory
static bool mma_path_ok(const Levels& lv, int S, int P, const void* loc, const void* attn, const FusedArgs* fzp, int vld) {
    if (g_debug_flags & 16) return false;
    if (lv.n != 4 || P != 4 || S > 1023 || (vld % 8) != 0) return false;
    if (64 + (size_t)(S + 2) * 64 + (size_t)MMA_NW * MMA_TAB_BYTES > (size_t)max_smem_optin()) return false;
    if (fzp) {
        if ((fzp->ld % (fzp->proj_bf16 ? 8 : 4)) != 0) return false;
        if (((uintptr_t)fzp->ref & (fzp->RD == 4 ? 15 : 7)) != 0) return false;
    } else if ((((uintptr_t)loc | (uintptr_t)attn) & 15) != 0) {
        return false;
    }
    return true;
}

static int launch_fwd_mma(const void* value, const void* loc, const void* attn, void* out, const Levels& lv, int B, int S,
                          int M, int Lq, const FusedArgs* fzp, cudaStream_t st, int vld) {
    const size_t smem = 64 + (size_t)(S + 2) * 64 + (size_t)MMA_NW * MMA_TAB_BYTES;
    // slab staging by TMA (north star: "TMA / shared-memory staging of per-level value tiles"): boxes of S / k pixels (k minimal with
    // S % k == 0, S / k <= 256 and even, so that every box lands on a 128-byte boundary); otherwise, or with dtlr_debug_flags(2097152),
    // the cp.async path
    CUtensorMap tmv;
    memset(&tmv, 0, sizeof(tmv));
    int tma_rows = 0;
    if (!(g_debug_flags & 2097152) && (((uintptr_t)value) & 15) == 0) {
        for (int k = 1; k <= 16; ++k)
            if (S % k == 0 && S / k <= 256 && ((S / k) % 2) == 0) { tma_rows = S / k; break; }
        if (tma_rows && make_tmap_2d_bf16(&tmv, value, (long long)B * S, M * 32, vld, tma_rows, 32, CU_TENSOR_MAP_SWIZZLE_NONE) != DTLR_OK) tma_rows = 0;
    }
    // one 32-query batch per warp where possible: ceil(Lq / 256) CTAs per (image, head); small problems are split further
    // (down to 64 queries per CTA) until the grid covers the machine twice
    int qsplit = (Lq + MMA_NW * 32 - 1) / (MMA_NW * 32);
    while ((long long)B * M * qsplit < 2ll * sm_count() && Lq / (qsplit + 1) >= 64) ++qsplit;
    const int q_per_cta = (Lq + qsplit - 1) / qsplit;
    qsplit = (Lq + q_per_cta - 1) / q_per_cta;
    const FusedArgs fz = fzp ? *fzp : FusedArgs{nullptr, nullptr, 0, 0, 0};
    const int mode = !fzp ? 0 : (fzp->proj_bf16 ? 2 : 1);
    // four queries per phase-2 iteration (measured on B200, fused call at B = 64, CUDA-graph timed: unroll 1 / 2 / 4 = 103.2 / 101.5 /
    // 98.2 us); dtlr_debug_flags 4194304 / 8388608: one / two queries (A/B)
    const int qu = (g_debug_flags & 4194304) ? 1 : ((g_debug_flags & 8388608) ? 2 : 4);
    auto pick = [&](auto k1, auto k2, auto k4) { return qu == 1 ? k1 : (qu == 4 ? k4 : k2); };
    auto k = mode == 0 ? pick(msda_fwd_mma_kernel<0, 1>, msda_fwd_mma_kernel<0, 2>, msda_fwd_mma_kernel<0, 4>)
                       : (mode == 1 ? pick(msda_fwd_mma_kernel<1, 1>, msda_fwd_mma_kernel<1, 2>, msda_fwd_mma_kernel<1, 4>)
                                    : pick(msda_fwd_mma_kernel<2, 1>, msda_fwd_mma_kernel<2, 2>, msda_fwd_mma_kernel<2, 4>));
    DTLR_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(qsplit, M, B), block(MMA_NW * 32);
    DTLR_CHECK_CUDA(launch_pdl(k, grid, block, smem, st, (const op16_t*)value, loc, (const float*)attn, (op16_t*)out, lv, S, M, Lq, q_per_cta, fz, vld, tmv, tma_rows));
    return DTLR_OK;
}

template <typename T>
static int launch_fwd_d32(const void* value, const void* loc, const void* attn, void* out, const Levels& lv, int B,
                          int S, int M, int Lq, int P, cudaStream_t st, const FusedArgs* fzp = nullptr, int vld = 0) {
    if (vld == 0) vld = M * 32;
    if (sizeof(T) == 2 && mma_path_ok(lv, S, P, loc, attn, fzp, vld))
        return launch_fwd_mma(value, loc, attn, out, lv, B, S, M, Lq, fzp, st, vld);
    const size_t slab = (size_t)(S + 2) * 32 * sizeof(T);
    const bool stage = slab <= (size_t)max_smem_optin();
    // resident CTAs per SM by shared memory (228 KB per SM, 1 KB reserved per CTA)
    int occ_smem = stage ? (int)min((size_t)8, (size_t)(228 * 1024) / (slab + 1024)) : 8;
    if (occ_smem < 1) occ_smem = 1;
    // 64 registers/thread -> at most 32 warps per SM: one 32-warp CTA when only one slab fits, else 16-warp CTAs
    if (occ_smem == 1)
        return launch_fwd_d32_nw<T, 32>(value, loc, attn, out, lv, B, S, M, Lq, P, stage, slab, 1, fzp, st, vld);
    return launch_fwd_d32_nw<T, 16>(value, loc, attn, out, lv, B, S, M, Lq, P, stage, slab, min(occ_smem, 2), fzp, st, vld);
}

}  // namespace dtlr

using namespace dtlr;

extern "C" int dtlr_msda_forward(const void* value, const int64_t* shapes, const int64_t* lsi, const void* loc,
                                 const void* attn, void* out, int B, int S, int M, int D, int L, int Lq, int P,
                                 int dtype, void* stream) {
    DTLR_CHECK_ARG(B >= 0 && Lq >= 0 && S > 0 && M > 0 && D > 0 && P > 0, "msda_forward: bad sizes");
    DTLR_CHECK_ARG(shapes && lsi, "msda_forward: null shapes");
    Levels lv;
    int rc = fill_levels(lv, shapes, lsi, L, S);
    if (rc) return rc;
    if (B == 0 || Lq == 0) return DTLR_OK;
    DTLR_CHECK_ARG(value && loc && attn && out, "msda_forward: null pointer");
    DTLR_CHECK_ARG((long long)B <= 65535 && (long long)M <= 65535, "msda_forward: B or M exceeds 65535");
    cudaStream_t st = (cudaStream_t)stream;
    const bool aligned = (((uintptr_t)value | (uintptr_t)loc) & 15) == 0;
    if (D == 32 && aligned && (dtype == DTLR_F32 || dtype == DTLR_OP16)) {
        return dtype == DTLR_F32 ? launch_fwd_d32<float>(value, loc, attn, out, lv, B, S, M, Lq, P, st)
                                 : launch_fwd_d32<op16_t>(value, loc, attn, out, lv, B, S, M, Lq, P, st);
    }
    const long long warps = (long long)B * Lq * M;
    const int threads = 256;
    const long long blocks = (warps * 32 + threads - 1) / threads;
    DTLR_CHECK_ARG(blocks < (1ll << 31), "msda_forward: problem too large for the generic kernel");
    if (dtype == DTLR_F32) {
        msda_fwd_generic_kernel<float><<<(unsigned)blocks, threads, 0, st>>>(
            (const float*)value, (const float*)loc, (const float*)attn, (float*)out, lv, B, S, M, D, Lq, P);
    } else if (dtype == DTLR_F64) {
        msda_fwd_generic_kernel<double><<<(unsigned)blocks, threads, 0, st>>>(
            (const double*)value, (const double*)loc, (const double*)attn, (double*)out, lv, B, S, M, D, Lq, P);
    } else {
        set_error("msda_forward: dtype %d with D=%d has no kernel (bf16 requires D=32, 16-byte aligned)", dtype, D);
        return DTLR_ERR_UNSUPPORTED;
    }
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_msda_forward_fused(const void* value, int value_ld, const int64_t* shapes, const int64_t* lsi, const void* proj,
                                       int ld_proj, int proj_dtype, const float* ref, int ref_dim, const float* valid_ratios,
                                       void* out, int B, int S, int M, int D, int L, int Lq, int P, int dtype, void* stream) {
    DTLR_CHECK_ARG(proj_dtype == DTLR_F32 || proj_dtype == DTLR_OP16, "msda_forward_fused: projection must be f32 or bf16");
    DTLR_CHECK_ARG(B >= 0 && Lq >= 0 && S > 0 && M > 0 && P > 0, "msda_forward_fused: bad sizes");
    DTLR_CHECK_ARG(D == 32 && (dtype == DTLR_F32 || dtype == DTLR_OP16), "msda_forward_fused: needs D=32, f32 or bf16 values");
    DTLR_CHECK_ARG(ref_dim == 2 || ref_dim == 4, "msda_forward_fused: reference points must have 2 or 4 coordinates");
    DTLR_CHECK_ARG(shapes && lsi, "msda_forward_fused: null shapes");
    DTLR_CHECK_ARG(ld_proj >= M * L * P * 3 && (ld_proj % 2) == 0, "msda_forward_fused: projection row pitch %d too small/odd", ld_proj);
    Levels lv;
    int rc = fill_levels(lv, shapes, lsi, L, S);
    if (rc) return rc;
    if (B == 0 || Lq == 0) return DTLR_OK;
    DTLR_CHECK_ARG(value && proj && ref && valid_ratios && out, "msda_forward_fused: null pointer");
    DTLR_CHECK_ARG((((uintptr_t)value | (uintptr_t)proj) & 15) == 0, "msda_forward_fused: value/proj must be 16-byte aligned");
    DTLR_CHECK_ARG((long long)B <= 65535 && (long long)M <= 65535, "msda_forward_fused: B or M exceeds 65535");
    DTLR_CHECK_ARG(value_ld >= M * 32 && (value_ld % 8) == 0, "msda_forward_fused: value row pitch %d too small / unaligned", value_ld);
    const FusedArgs fz{ref, valid_ratios, ld_proj, ref_dim, proj_dtype == DTLR_OP16 ? 1 : 0};
    cudaStream_t st = (cudaStream_t)stream;
    return dtype == DTLR_F32 ? launch_fwd_d32<float>(value, (const float*)proj, nullptr, out, lv, B, S, M, Lq, P, st, &fz, value_ld)
                             : launch_fwd_d32<op16_t>(value, (const float*)proj, nullptr, out, lv, B, S, M, Lq, P, st, &fz, value_ld);
}

extern "C" int dtlr_msda_backward(const void* value, const int64_t* shapes, const int64_t* lsi, const void* loc,
                                  const void* attn, const void* grad_out, void* grad_value, void* grad_loc,
                                  void* grad_attn, int B, int S, int M, int D, int L, int Lq, int P, int dtype,
                                  void* stream) {
    DTLR_CHECK_ARG(B >= 0 && Lq >= 0 && S > 0 && M > 0 && D > 0 && P > 0, "msda_backward: bad sizes");
    DTLR_CHECK_ARG(shapes && lsi, "msda_backward: null shapes");
    Levels lv;
    int rc = fill_levels(lv, shapes, lsi, L, S);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t esz = dtype == DTLR_F64 ? 8 : 4;
    DTLR_CHECK_ARG(dtype == DTLR_F32 || dtype == DTLR_F64, "msda_backward: dtype must be f32 or f64");
    if (B == 0) return DTLR_OK;
    DTLR_CHECK_ARG(grad_value, "msda_backward: null grad_value");
    DTLR_CHECK_CUDA(cudaMemsetAsync(grad_value, 0, (size_t)B * S * M * D * esz, st));
    if (Lq == 0) return DTLR_OK;
    DTLR_CHECK_ARG(value && loc && attn && grad_out && grad_loc && grad_attn, "msda_backward: null pointer");
    const long long warps = (long long)B * Lq * M;
    const int threads = 256;
    const long long blocks = (warps * 32 + threads - 1) / threads;
    DTLR_CHECK_ARG(blocks < (1ll << 31), "msda_backward: problem too large");
    int maxW = 0;
    for (int l = 0; l < L; ++l) maxW = max(maxW, lv.W[l]);
    const bool fast = dtype == DTLR_F32 && D == 32 && L * P == 16 && (P == 2 || P == 4 || P == 8 || P == 16) &&
                      (long long)S + maxW + 2 < (1ll << 24) && !(g_debug_flags & 8192) &&   // flag 8192: generic kernel (A/B)
                      ((((uintptr_t)value | (uintptr_t)grad_value | (uintptr_t)grad_out) & 15) == 0) &&
                      ((((uintptr_t)loc | (uintptr_t)grad_loc) & 7) == 0);
    if (fast) {
        // grid-stride over the (b,q,m) items: enough CTAs to fill the machine a few times over (32 x 8 warps per SM at most)
        auto launch = [&](auto kern, int threads_per_cta = 256) {
            const long long ctas = (warps * 32 + threads_per_cta - 1) / threads_per_cta;
            const unsigned g = (unsigned)min(ctas, (long long)sm_count() * 32 * (256 / threads_per_cta));
            kern<<<g, threads_per_cta, 0, st>>>((const float*)value, (const float*)loc, (const float*)attn, (const float*)grad_out,
                                                (float*)grad_value, (float*)grad_loc, (float*)grad_attn, lv, warps, S, M, Lq);
        };
        // grouped-load variant: the default since round 2 (the complete GPU suite passes with it; B = 64, Lq = 912: 855 us vs 1197 us,
        // B = 32, Lq = 900: 433 vs 594 us).  dtlr_debug_flags(65536) selects the one-load-in-flight variant (A/B); int offsets need
        // S*M*32 < 2^31
        if (!(g_debug_flags & 65536) && P == 4 && (long long)S * M * 32 < (1ll << 31)) {
            // flag 131072: 8 loads in flight, 128-thread CTAs (148 registers -> 3 CTAs per SM) -- not yet measured
            if (g_debug_flags & 131072) launch(msda_bwd_d32_grouped_kernel<4, 8>, 128);
            else launch(msda_bwd_d32_grouped_kernel<4, 4>);
        } else switch (P) {
            case 2: launch(msda_bwd_d32_kernel<2>); break;
            case 4: launch(msda_bwd_d32_kernel<4>); break;
            case 8: launch(msda_bwd_d32_kernel<8>); break;
            default: launch(msda_bwd_d32_kernel<16>); break;
        }
    } else if (dtype == DTLR_F32) {
        msda_bwd_kernel<float><<<(unsigned)blocks, threads, 0, st>>>(
            (const float*)value, (const float*)loc, (const float*)attn, (const float*)grad_out, (float*)grad_value,
            (float*)grad_loc, (float*)grad_attn, lv, B, S, M, D, Lq, P);
    } else {
        msda_bwd_kernel<double><<<(unsigned)blocks, threads, 0, st>>>(
            (const double*)value, (const double*)loc, (const double*)attn, (const double*)grad_out,
            (double*)grad_value, (double*)grad_loc, (double*)grad_attn, lv, B, S, M, D, Lq, P);
    }
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}
