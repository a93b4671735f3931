// dtlr_b200 -- kernels of the BACKWARD half of the fine-tune step (reference engine.py:172-274 train_one_epoch_CTC:
// loss.backward() + clip_grad_norm_ + AdamW.step over the model of models/dino/dino.py / deformable_transformer.py).
// The reference gets all of this from torch autograd (cuBLAS / ATen kernels); here every piece the transformer needs is a kernel:
//   wgrad          dW[n,k] += sum_r dY[r,n] X[r,k]   tcgen05, BOTH operands MN-major straight from the row-major activations
//                  (no transposed copies), split over the reduction rows, fp32 vector reductions into the gradient arena
//   colsum         bias gradients (and per-level sums for level_embed)
//   layernorm_bwd  nn.LayerNorm(256) backward with the statistics recomputed from the saved pre-norm rows
//   relu_bwd, msda_bwd_glue (softmax + sampling-location chain of MSDeformAttn), add_cast
//   pack_weights   one launch re-creates the 16-bit W and W^T operand copies of every Linear after the optimizer step
//   sumsq + adamw  clip_grad_norm_(max_norm) and torch.optim.AdamW fused over the flat parameter / gradient arenas
// dgrad needs no kernel of its own: dX = dY . W is dtlr_gemm against the packed W^T.
#include "tc_common.cuh"

#define DTLR_LAUNCH(kernel, grid, block, smem, st, ...)                                                                   \
    do {                                                                                                                  \
        cudaError_t _le = dtlr::launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__);             \
        if (_le != cudaSuccess) {                                                                                         \
            dtlr::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_le), __FILE__, __LINE__);            \
            return DTLR_ERR_CUDA;                                                                                         \
        }                                                                                                                 \
    } while (0)

#define DISPATCH_T(dtype, ...)                                          \
    if ((dtype) == DTLR_F32) { using T = float; __VA_ARGS__ }           \
    else if ((dtype) == DTLR_OP16) { using T = op16_t; __VA_ARGS__ }    \
    else { set_error("unsupported dtype %d", (int)(dtype)); return DTLR_ERR_INVALID; }

namespace dtlr {
namespace train {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <typename T> __device__ __forceinline__ void ld8(const T* p, float* f);
template <> __device__ __forceinline__ void ld8<float>(const float* p, float* f) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <> __device__ __forceinline__ void ld8<op16_t>(const op16_t* p, float* f) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = op16_lo_f32(w[i]); f[2 * i + 1] = op16_hi_f32(w[i]); }
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    op16x2_t v = op16_pack2(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
template <typename T> __device__ __forceinline__ void st8(T* p, const float* f);
template <> __device__ __forceinline__ void st8<float>(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
template <> __device__ __forceinline__ void st8<op16_t>(op16_t* p, const float* f) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
}
template <typename T> __device__ __forceinline__ float ld1(const T* p);
template <> __device__ __forceinline__ float ld1<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld1<op16_t>(const op16_t* p) { return op16_to_f32(*p); }
template <typename T> __device__ __forceinline__ void st1(T* p, float v);
template <> __device__ __forceinline__ void st1<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st1<op16_t>(op16_t* p, float v) { *p = f32_to_op16(v); }

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// ---------------------------------------------------------------------------------------------- wgrad on tcgen05
// dW[n, k] += sum_r dY[r, n] * X[r, k]   (nn.Linear weight gradient; reference: autograd of F.linear -> cuBLAS "TN" GEMM)
// The reduction runs over the ROWS of two row-major activations, i.e. both MMA operands are MN-major: A = dY^T (M-dim n contiguous
// in memory), B = X^T (N-dim k contiguous).  TMA delivers 64-row x 64-column boxes (128-byte rows, SWIZZLE_128B) -- exactly the
// canonical MN-major SW128 atom stack of the UMMA shared-memory descriptor ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: rows
// (k) 128 B apart, 8-row groups SBO = 1024 B apart, the next 64 columns LBO = one box (8 KB) further.  One MMA (K = 16) consumes
// two 8-row groups: the descriptor start address advances by 2048 B per step.  No transposed copy of any activation is made.
// CTA = one 128 x 128 tile of dW for one slice of the rows; fp32 accumulator in TMEM (128 columns); epilogue = one
// red.global.add.v4.f32 per 4 accumulator columns into the (zero-initialised, shared-by-all-slices) fp32 gradient.
constexpr int WG_BM = 128, WG_BN = 128, WG_RB = 64, WG_STAGES = 4;
constexpr int WG_ATOM = WG_RB * 128;                                   // one TMA box: 64 rows x 64 16-bit columns = 8 KB
constexpr int WG_STAGE_BYTES = (WG_BM / 64 + WG_BN / 64) * WG_ATOM;    // 32 KB
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + 1024 + 128;

__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}

__global__ void __launch_bounds__(192, 1)
wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ dW, int ldw,
                     int N, int K, int rows, int rb_per_split, int swap_lbo_sbo) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + WG_STAGES;
    uint64_t* acc_bar = empty_bar + WG_STAGES;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.x * WG_BM, k0 = blockIdx.y * WG_BN;
    const int total_rb = (rows + WG_RB - 1) / WG_RB;
    const int rb0 = blockIdx.z * rb_per_split;
    const int nrb = min(rb_per_split, total_rb - rb0);
    if (nrb <= 0) return;                                              // (uniform over the CTA)
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < WG_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(acc_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<WG_BN>(tmem_ptr);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_launch_dependents();
    pdl_wait();
    if (warp == 0) {
        if (elect_one()) {
            for (int i = 0; i < nrb; ++i) {
                const int s = i % WG_STAGES;
                const uint32_t ph = (i / WG_STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                unsigned char* sa = smem + s * WG_STAGE_BYTES;
                const int r = (rb0 + i) * WG_RB;
                mbar_expect_tx(&full_bar[s], WG_STAGE_BYTES);
                tma_load_2d(sa, &tmA, &full_bar[s], n0, r);
                tma_load_2d(sa + WG_ATOM, &tmA, &full_bar[s], n0 + 64, r);
                tma_load_2d(sa + 2 * WG_ATOM, &tmB, &full_bar[s], k0, r);
                tma_load_2d(sa + 3 * WG_ATOM, &tmB, &full_bar[s], k0 + 64, r);
            }
        }
    } else if (warp == 1) {
        // instruction descriptor: D = f32, A/B format = the library's 16-bit type, A and B MN-major (bits 15, 16), N>>3, M>>4
        constexpr uint32_t IDESC = (1u << 4) | OP16_IDESC_AB | (1u << 15) | (1u << 16) | ((uint32_t)(WG_BN >> 3) << 17) |
                                   ((uint32_t)(WG_BM >> 4) << 24);
        const uint32_t lbo = swap_lbo_sbo ? 1024u : (uint32_t)WG_ATOM, sbo = swap_lbo_sbo ? (uint32_t)WG_ATOM : 1024u;
        for (int i = 0; i < nrb; ++i) {
            const int s = i % WG_STAGES;
            const uint32_t ph = (i / WG_STAGES) & 1;
            mbar_wait(&full_bar[s], ph);
            tcgen05_fence_after();
            if (elect_one()) {
                const uint32_t sa = smem_u32(smem + s * WG_STAGE_BYTES);
                const uint64_t da = make_sw128_mnmajor_desc(sa, lbo, sbo);
                const uint64_t db = make_sw128_mnmajor_desc(sa + 2 * WG_ATOM, lbo, sbo);
#pragma unroll
                for (int k = 0; k < WG_RB / 16; ++k)                    // 16 reduction rows = two 8-row groups = 2048 B per step
                    umma_bf16(tmem_base, da + (uint64_t)(128 * k), db + (uint64_t)(128 * k), IDESC, (i | k) != 0);
                umma_commit(&empty_bar[s]);
                if (i == nrb - 1) umma_commit(acc_bar);
            }
            __syncwarp();
        }
    } else {
        const int qd = warp & 3;                                       // TMEM lane quarter this warp may read
        mbar_wait(acc_bar, 0);
        tcgen05_fence_after();
        const int n = n0 + qd * 32 + lane;
        float* drow = dW + (size_t)n * ldw + k0;
        const bool vec = ((ldw & 3) == 0) && ((((uintptr_t)dW) & 15) == 0);
#pragma unroll 1
        for (int c = 0; c < WG_BN; c += 32) {
            uint32_t acc[32];
            tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)c, acc);
            if (n < N) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const int col = k0 + c + j;
                    if (vec && col + 3 < K) {
                        red_add_v4(drow + c + j, __uint_as_float(acc[j]), __uint_as_float(acc[j + 1]), __uint_as_float(acc[j + 2]),
                                   __uint_as_float(acc[j + 3]));
                    } else {
#pragma unroll
                        for (int t = 0; t < 4; ++t)
                            if (col + t < K) atomicAdd(drow + c + j + t, __uint_as_float(acc[j + t]));
                    }
                }
            }
        }
        tcgen05_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc<WG_BN>(tmem_base);
    }
}

// fp32 parity mode: the same contraction with exact fp32 FMAs (SIMT), 64 x 64 tile of dW per CTA, rows split over blockIdx.z
__global__ void __launch_bounds__(256)
wgrad_f32_kernel(const float* __restrict__ dY, int ldy, const float* __restrict__ X, int ldx, float* __restrict__ dW, int ldw, int N, int K,
                 int rows, int rows_per_split) {
    __shared__ float sa[32][65], sb[32][65];
    pdl_launch_dependents();
    pdl_wait();
    const int n0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
    const int r_begin = blockIdx.z * rows_per_split, r_end = min(rows, r_begin + rows_per_split);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;           // thread -> 4 x 4 outputs: n = ty*4.., k = tx*4..
    float acc[4][4] = {};
    for (int r0 = r_begin; r0 < r_end; r0 += 32) {
        for (int i = threadIdx.x; i < 32 * 64; i += 256) {
            const int rr = i >> 6, cc = i & 63;
            const int r = r0 + rr;
            sa[rr][cc] = (r < r_end && n0 + cc < N) ? dY[(size_t)r * ldy + n0 + cc] : 0.f;
            sb[rr][cc] = (r < r_end && k0 + cc < K) ? X[(size_t)r * ldx + k0 + cc] : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int rr = 0; rr < 32; ++rr) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = sa[rr][ty * 4 + i]; b[i] = sb[rr][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + ty * 4 + i, k = k0 + tx * 4 + j;
            if (n < N && k < K) atomicAdd(dW + (size_t)n * ldw + k, acc[i][j]);
        }
}

// ---------------------------------------------------------------------------------------------- column sums
// out[c] += sum over the rows of x[:, c]; rows = nseg segments of seg_rows consecutive rows, segment s starting at row s*seg_stride
// (bias gradients: one segment; per-level sums over the (B, S, C) token tensor for level_embed: nseg = B, seg_stride = S).
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ x, long long ld, int N, int nseg, int seg_rows, long long seg_stride, float* __restrict__ out,
              int rows_per_block) {
    __shared__ float sm[256 * 8];
    pdl_launch_dependents();
    pdl_wait();
    const int cg_all = (N + 7) / 8;
    const int cg0 = blockIdx.y * 256;
    const int cg = min(256, cg_all - cg0);                            // column groups (8 columns each) of this block
    const int rp = 256 / cg;                                          // rows in flight per block
    const int c = threadIdx.x % cg, ry = threadIdx.x / cg;
    const int total = nseg * seg_rows;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(total, r0 + rows_per_block);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int col = (cg0 + c) * 8;
    if (ry < rp) {
        if (col + 7 < N) {
            // four rows per iteration: their loads are independent and issued together (the first version took one row at a time
            // behind a 64-bit division and was latency bound)
            int g = r0 + ry;
            for (; g + 3 * rp < r1; g += 4 * rp) {
                float f[4][8];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int gg = g + u * rp;
                    const int seg = nseg == 1 ? 0 : gg / seg_rows;
                    ld8<T>(x + ((long long)seg * seg_stride + (gg - seg * seg_rows)) * ld + col, f[u]);
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] += (f[0][k] + f[1][k]) + (f[2][k] + f[3][k]);
            }
            for (; g < r1; g += rp) {
                float f[8];
                const int seg = nseg == 1 ? 0 : g / seg_rows;
                ld8<T>(x + ((long long)seg * seg_stride + (g - seg * seg_rows)) * ld + col, f);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] += f[k];
            }
        } else {
            for (int g = r0 + ry; g < r1; g += rp) {
                const int seg = nseg == 1 ? 0 : g / seg_rows;
                const T* p = x + ((long long)seg * seg_stride + (g - seg * seg_rows)) * ld + col;
                for (int k = 0; k < 8; ++k)
                    if (col + k < N) acc[k] += ld1<T>(p + k);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) sm[threadIdx.x * 8 + k] = acc[k];
    __syncthreads();
    if (ry == 0) {
        float s[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            s[k] = 0.f;
            for (int j = 0; j < rp; ++j) s[k] += sm[(j * cg + c) * 8 + k];
        }
        // few CTAs and vector reductions: every CTA adds into the same N addresses, and same-line atomics serialise in L2 (the first
        // version: 456 CTAs x 256 scalar atomics on 8 lines = 39 us for a 15 MB input; ncu-free diagnosis from the launch list)
        if (col + 7 < N && ((((uintptr_t)(out + col)) & 15) == 0)) {
            red_add_v4(out + col, s[0], s[1], s[2], s[3]);
            red_add_v4(out + col + 4, s[4], s[5], s[6], s[7]);
        } else {
            for (int k = 0; k < 8; ++k)
                if (col + k < N) atomicAdd(out + col + k, s[k]);
        }
    }
}

// ---------------------------------------------------------------------------------------------- LayerNorm(256) backward
// y = LN(z) * gamma + beta  (deformable_transformer.py:813-814, 806-807, 906-907, 956-957, 878-879, 758).  Saved: the pre-norm rows z
// (storage dtype T); mean / rstd are recomputed (one warp per row, 8 channels per lane).  dy (+ dy2) fp32 ->
//   dz = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma;  written as fp32 (the residual gradient stream) and / or as T
//   (the operand of the dgrad / wgrad contractions that follow); dgamma += sum dy * xhat, dbeta += sum dy (per-CTA partial sums,
//   one atomicAdd per channel per CTA).
template <typename T>
__global__ void __launch_bounds__(256)
layernorm256_bwd_kernel(const T* __restrict__ z, const float* __restrict__ dy, const float* __restrict__ dy2, const float* __restrict__ gamma,
                        float* __restrict__ dz32, T* __restrict__ dzT, float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows,
                        float eps) {
    __shared__ float red[8][256];
    pdl_launch_dependents();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float g[8];
    ld8<float>(gamma + lane * 8, g);
    float ag[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, ab[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (long long row = (long long)blockIdx.x * 8 + warp; row < rows; row += (long long)gridDim.x * 8) {
        const size_t off = (size_t)row * 256 + lane * 8;
        float x[8], d[8];
        ld8<T>(z + off, x);
        ld8<float>(dy + off, d);
        if (dy2) {
            float d2[8];
            ld8<float>(dy2 + off, d2);
#pragma unroll
            for (int k = 0; k < 8; ++k) d[k] += d2[k];
        }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += x[k];
        const float mean = warp_sum(s) * (1.f / 256.f);
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { x[k] -= mean; ss += x[k] * x[k]; }
        const float rstd = rsqrtf(warp_sum(ss) * (1.f / 256.f) + eps);
        float s1 = 0.f, s2 = 0.f, gk[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x[k] *= rstd;                                             // xhat
            gk[k] = d[k] * g[k];
            s1 += gk[k];
            s2 = fmaf(gk[k], x[k], s2);
            ab[k] += d[k];
            ag[k] = fmaf(d[k], x[k], ag[k]);
        }
        s1 = warp_sum(s1) * (1.f / 256.f);
        s2 = warp_sum(s2) * (1.f / 256.f);
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = rstd * (gk[k] - s1 - x[k] * s2);
        if (dz32) st8<float>(dz32 + off, o);
        if (dzT) st8<T>(dzT + off, o);
    }
    // per-CTA sums, then one 16-byte reduction per 4 channels (same-line atomics serialise in L2: few CTAs, vector operations)
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        float* dst = pass == 0 ? dgamma : dbeta;
        if (!dst) continue;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) red[warp][lane * 8 + k] = pass == 0 ? ag[k] : ab[k];
        __syncthreads();
        if (threadIdx.x < 64) {
            float s4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                s4[j] = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) s4[j] += red[w][threadIdx.x * 4 + j];
            }
            if ((((uintptr_t)dst) & 15) == 0) red_add_v4(dst + threadIdx.x * 4, s4[0], s4[1], s4[2], s4[3]);
            else
                for (int j = 0; j < 4; ++j) atomicAdd(dst + threadIdx.x * 4 + j, s4[j]);
        }
    }
}

// dh = (h > 0) ? dh : 0, in place (ReLU of linear1 / MLP hidden layers: deformable_transformer.py:805, 877; models/dino/utils.py:120)
template <typename T>
__global__ void relu_bwd_kernel(T* __restrict__ dh, const T* __restrict__ h, long long n8) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float a[8], b[8];
        ld8<T>(dh + i * 8, a);
        ld8<T>(h + i * 8, b);
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = b[k] > 0.f ? a[k] : 0.f;
        st8<T>(dh + i * 8, a);
    }
}

// out = a (+ b) (+ c), fp32 inputs, output fp32 or the 16-bit type (gradient stream -> contraction operand)
template <typename T>
__global__ void add_cast_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c, T* __restrict__ out,
                                long long n8) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float x[8], y[8];
        ld8<float>(a + i * 8, x);
        if (b) {
            ld8<float>(b + i * 8, y);
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] += y[k];
        }
        if (c) {
            ld8<float>(c + i * 8, y);
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] += y[k];
        }
        st8<T>(out + i * 8, x);
    }
}

// ---------------------------------------------------------------------------------------------- MSDeformAttn prologue backward
// Chain rule through models/dino/ops/modules/ms_deform_attn.py:98-108 (softmax over the L*P logits of a head; sampling locations =
// reference point + offset / (W_l, H_l)  [2-d reference points]  or  + offset / P * (w, h) * 0.5  [4-d boxes]; reference points and
// valid ratios carry no gradient -- they are detached, deformable_transformer.py:737).  grad_loc / grad_attn are the outputs of
// dtlr_msda_backward; the result is the gradient of the fused projection row [offsets (M*L*P*2) | logits (M*L*P)].
struct GlueLevels { int n; int H[8], W[8]; };
template <typename T>
__global__ void msda_bwd_glue_kernel(const float* __restrict__ gloc, const float* __restrict__ gattn, const float* __restrict__ attn,
                                     const float* __restrict__ ref, int RD, const float* __restrict__ valid_ratios,
                                     const __grid_constant__ GlueLevels lv, T* __restrict__ dproj, int ld, int B, int Lq, int M, int P) {
    pdl_launch_dependents();
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * Lq * M;
    if (i >= total) return;
    const int m = (int)(i % M);
    const long long row = i / M;
    const int b = (int)(row / Lq);
    const int L = lv.n, LP = L * P;
    const float* at = attn + (size_t)i * LP;
    const float* ga = gattn + (size_t)i * LP;
    const float* gl = gloc + (size_t)i * LP * 2;
    float dot = 0.f;
    for (int k = 0; k < LP; ++k) dot = fmaf(at[k], ga[k], dot);
    T* po = dproj + (size_t)row * ld + (size_t)m * LP * 2;
    T* pl = dproj + (size_t)row * ld + (size_t)M * LP * 2 + (size_t)m * LP;
    const float* rf = ref + (size_t)row * RD;
    for (int l = 0; l < L; ++l) {
        float sx, sy;
        if (RD == 2) {
            sx = 1.f / (float)lv.W[l];
            sy = 1.f / (float)lv.H[l];
        } else {
            const float vx = valid_ratios[((size_t)b * L + l) * 2], vy = valid_ratios[((size_t)b * L + l) * 2 + 1];
            sx = (rf[2] * vx) * 0.5f / (float)P;
            sy = (rf[3] * vy) * 0.5f / (float)P;
        }
        for (int p = 0; p < P; ++p) {
            const int k = l * P + p;
            st1<T>(po + 2 * k, gl[2 * k] * sx);
            st1<T>(po + 2 * k + 1, gl[2 * k + 1] * sy);
            st1<T>(pl + k, at[k] * (ga[k] - dot));
        }
    }
}

// L * P == 16 (every shipped config: 4 levels x 4 points): the same arithmetic with the 16 points of a (query, head) in registers,
// 16-byte loads and stores (the generic kernel above walks them with scalar accesses 128 bytes apart across the warp: 113 us per call
// at 32 lines against ~30 us of memory time)
template <typename T, int NL, int NP>
__global__ void __launch_bounds__(128)
msda_bwd_glue16_kernel(const float* __restrict__ gloc, const float* __restrict__ gattn, const float* __restrict__ attn,
                       const float* __restrict__ ref, int RD, const float* __restrict__ valid_ratios, const __grid_constant__ GlueLevels lv,
                       T* __restrict__ dproj, int ld, int B, int Lq, int M) {
    static_assert(NL * NP == 16, "16 sampling points per head");
    pdl_launch_dependents();
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * Lq * M) return;
    const int m = (int)(i % M);
    const long long row = i / M;
    const int b = (int)(row / Lq);
    float at[16], ga[16], gl[32];
#pragma unroll
    for (int k = 0; k < 16; k += 8) { ld8<float>(attn + (size_t)i * 16 + k, at + k); ld8<float>(gattn + (size_t)i * 16 + k, ga + k); }
#pragma unroll
    for (int k = 0; k < 32; k += 8) ld8<float>(gloc + (size_t)i * 32 + k, gl + k);
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) dot = fmaf(at[k], ga[k], dot);
    const float* rf = ref + (size_t)row * RD;
    float rw = 0.f, rh = 0.f;
    if (RD == 4) { rw = rf[2]; rh = rf[3]; }
    float o[32], lg[16];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        float sx, sy;
        if (RD == 2) {
            sx = 1.f / (float)lv.W[l];
            sy = 1.f / (float)lv.H[l];
        } else {
            sx = (rw * valid_ratios[((size_t)b * NL + l) * 2]) * 0.5f / (float)NP;
            sy = (rh * valid_ratios[((size_t)b * NL + l) * 2 + 1]) * 0.5f / (float)NP;
        }
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int k = l * NP + p;
            o[2 * k] = gl[2 * k] * sx;
            o[2 * k + 1] = gl[2 * k + 1] * sy;
            lg[k] = at[k] * (ga[k] - dot);
        }
    }
    T* po = dproj + (size_t)row * ld + (size_t)m * 32;
    T* pl = dproj + (size_t)row * ld + (size_t)M * 32 + (size_t)m * 16;
#pragma unroll
    for (int k = 0; k < 32; k += 8) st8<T>(po + k, o + k);
#pragma unroll
    for (int k = 0; k < 16; k += 8) st8<T>(pl + k, lg + k);
}

// ---------------------------------------------------------------------------------------------- operand copies of the weights
// After every optimizer step each Linear needs its weight as a 16-bit [N, K] operand (forward, K-major) and as a 16-bit [K, N]
// operand (dgrad: dX = dY . W is dtlr_gemm against W^T).  ONE launch walks a device table of matrices; 32 x 32 tiles through shared
// memory so that both copies are written with coalesced rows.  Entry (8 x int64): src fp32 ptr, rows, cols, ld_src, dst ptr (or 0),
// ld_dst, dstT ptr (or 0), ld_dstT; tile_start[e] = first tile of entry e (int32 prefix, tile_start[n_entries] = total).
template <typename T>
__global__ void __launch_bounds__(256)
pack_weights_kernel(const long long* __restrict__ table, const int* __restrict__ tile_start, int n_entries) {
    __shared__ float tile[32][33];
    __shared__ int s_e;
    pdl_launch_dependents();
    pdl_wait();
    if (threadIdx.x == 0) {
        int lo = 0, hi = n_entries - 1;                               // last entry with tile_start <= blockIdx.x
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (tile_start[mid] <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
        }
        s_e = lo;
    }
    __syncthreads();
    const int e = s_e;
    const long long* t = table + (size_t)e * 8;
    const float* src = reinterpret_cast<const float*>(t[0]);
    const int rows = (int)t[1], cols = (int)t[2];
    const long long ld_src = t[3], ld_dst = t[5], ld_dstT = t[7];
    T* dst = reinterpret_cast<T*>(t[4]);
    T* dstT = reinterpret_cast<T*>(t[6]);
    const int tl = (int)blockIdx.x - tile_start[e];
    const int tiles_c = (cols + 31) / 32;
    const int r0 = (tl / tiles_c) * 32, c0 = (tl % tiles_c) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = r0 + ty + j * 8, c = c0 + tx;
        const float v = (r < rows && c < cols) ? src[(size_t)r * ld_src + c] : 0.f;
        tile[ty + j * 8][tx] = v;
        if (dst && r < rows && c < cols) st1<T>(dst + (size_t)r * ld_dst + c, v);
    }
    __syncthreads();
    if (dstT) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + ty + j * 8, r = r0 + tx;                // transposed: row index of dstT = c
            if (c < cols && r < rows) st1<T>(dstT + (size_t)c * ld_dstT + r, tile[tx][ty + j * 8]);
        }
    }
}

// ---------------------------------------------------------------------------------------------- ResNet / input_proj backward pieces
// (reference: torch autograd over cuDNN for models/dino/backbone.py:109-128 layer2-4 and models/dino/dino.py:118-135 input_proj)
// v = y > 0 ? dy : 0 -- the ReLU that closes a Bottleneck (y = relu(conv3 + identity)): the masked gradient is needed twice, as fp32 (it
// continues along the identity path) and as the 16-bit operand of conv3's dgrad / wgrad.  dy32 is updated in place.
template <typename T>
__global__ void relu_bwd_dual_kernel(float* __restrict__ dy32, const T* __restrict__ y, T* __restrict__ out16, long long n8) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float d[8], yy[8];
        ld8<float>(dy32 + i * 8, d);
        ld8<T>(y + i * 8, yy);
#pragma unroll
        for (int k = 0; k < 8; ++k) d[k] = yy[k] > 0.f ? d[k] : 0.f;
        st8<float>(dy32 + i * 8, d);
        if (out16) st8<T>(out16 + i * 8, d);
    }
}

// nn.GroupNorm(32, 256) backward (models/dino/dino.py:121-124) for one feature level.  x fp32 [B, HW, 256] = the saved conv output;
// dy fp32: row (b, hw) at dy + (b * dy_stride_b + hw) * 256 (a level slice of the (B, S, 256) token-gradient tensor).
// CTA = (32-channel block = 4 groups, image); thread = (row lane 0..63, group 0..3) walks the HW rows three times (statistics, the two
// projections, the output) -- the slice (<= 655 KB per image) stays in L2.  dgamma / dbeta: one atomic per channel per image.
template <typename T>
__global__ void __launch_bounds__(256)
groupnorm8_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long dy_stride_b, const float* __restrict__ gamma,
                      T* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, int HW, int C, float eps) {
    __shared__ float red[64][4][2];
    __shared__ float stat[4][2];
    __shared__ float chan[64][32];
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y, c0 = blockIdx.x * 32;
    const int gq = threadIdx.x & 3, rl = threadIdx.x >> 2;
    const int ch = c0 + gq * 8;
    const float* xb = x + (size_t)b * HW * C + ch;
    const float* db = dy + (size_t)b * dy_stride_b * C + ch;
    float g[8];
    ld8<float>(gamma + ch, g);
    auto group_reduce = [&](float a, float c, float& ra, float& rc) {
        red[rl][gq][0] = a;
        red[rl][gq][1] = c;
        __syncthreads();
        if (threadIdx.x < 8) {
            float t = 0.f;
            for (int r = 0; r < 64; ++r) t += red[r][threadIdx.x >> 1][threadIdx.x & 1];
            stat[threadIdx.x >> 1][threadIdx.x & 1] = t;
        }
        __syncthreads();
        ra = stat[gq][0];
        rc = stat[gq][1];
        __syncthreads();
    };
    float s = 0.f, ss = 0.f;
    for (int r = rl; r < HW; r += 64) {
        float v[8];
        ld8<float>(xb + (size_t)r * C, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) { s += v[k]; ss = fmaf(v[k], v[k], ss); }
    }
    float S1, S2;
    group_reduce(s, ss, S1, S2);
    const float inv_n = 1.f / (float)(HW * 8);
    const float mean = S1 * inv_n;
    const float rstd = rsqrtf(fmaxf(S2 * inv_n - mean * mean, 0.f) + eps);
    float a1 = 0.f, a2 = 0.f, dg[8], dbt[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) dg[k] = dbt[k] = 0.f;
    for (int r = rl; r < HW; r += 64) {
        float v[8], d[8];
        ld8<float>(xb + (size_t)r * C, v);
        ld8<float>(db + (size_t)r * C, d);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float xh = (v[k] - mean) * rstd, gy = d[k] * g[k];
            a1 += gy;
            a2 = fmaf(gy, xh, a2);
            dg[k] = fmaf(d[k], xh, dg[k]);
            dbt[k] += d[k];
        }
    }
    float A1, A2;
    group_reduce(a1, a2, A1, A2);
    A1 *= inv_n;
    A2 *= inv_n;
    for (int r = rl; r < HW; r += 64) {
        float v[8], d[8], o[8];
        ld8<float>(xb + (size_t)r * C, v);
        ld8<float>(db + (size_t)r * C, d);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float xh = (v[k] - mean) * rstd;
            o[k] = rstd * (d[k] * g[k] - A1 - xh * A2);
        }
        st8<T>(dx + ((size_t)b * HW + r) * C + ch, o);
    }
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        float* dst = pass == 0 ? dgamma : dbeta;
        if (!dst) continue;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) chan[rl][gq * 8 + k] = pass == 0 ? dg[k] : dbt[k];
        __syncthreads();
        if (threadIdx.x < 32) {
            float t = 0.f;
            for (int r = 0; r < 64; ++r) t += chan[r][threadIdx.x];
            atomicAdd(dst + c0 + threadIdx.x, t);
        }
    }
}

// Transposed convolution as a gather: dx[b, yi, xi, c] (+)= sum over the taps (kh, kw) whose output pixel yo = (yi + pad - kh) / stride,
// xo = (xi + pad - kw) / stride exists, of dcol[(b, yo, xo), (kh * KW + kw) * C + c].  dcol fp32 [B*Ho*Wo, ldc] = dY . W (dtlr_gemm against
// the transposed weight copy); used for the stride-2 3x3 convolutions, the strided 1x1 downsamples (KH = KW = 1: a row scatter) and
// input_proj[3] (reference: cuDNN's strided dgrad kernels).
__global__ void __launch_bounds__(256)
col2im_kernel(const float* __restrict__ dcol, int ldc, float* __restrict__ dx, int B, int H, int W, int C, int KH, int KW, int stride, int pad,
              int Ho, int Wo, int accumulate) {
    pdl_launch_dependents();
    pdl_wait();
    const long long total = (long long)B * H * W * (C / 4);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % (C / 4));
        long long p = i / (C / 4);
        const int xi = (int)(p % W);
        p /= W;
        const int yi = (int)(p % H), b = (int)(p / H);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int kh = 0; kh < KH; ++kh) {
            const int ty = yi + pad - kh;
            if (ty < 0 || (ty % stride) != 0) continue;
            const int yo = ty / stride;
            if (yo >= Ho) continue;
            for (int kw = 0; kw < KW; ++kw) {
                const int tx = xi + pad - kw;
                if (tx < 0 || (tx % stride) != 0) continue;
                const int xo = tx / stride;
                if (xo >= Wo) continue;
                const float4 v = *reinterpret_cast<const float4*>(dcol + ((size_t)(b * Ho + yo) * Wo + xo) * ldc + (size_t)(kh * KW + kw) * C + c4 * 4);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        float4* o = reinterpret_cast<float4*>(dx + (((size_t)b * H + yi) * W + xi) * C + c4 * 4);
        if (accumulate) { const float4 q = *o; acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w; }
        *o = acc;
    }
}

// Operand copies of the convolution weights with FrozenBatchNorm folded in (backbone.py:62-72: w' = w * scale[cout]), one launch over a
// device table.  Entry (10 x int64): src fp32 [Cout, Cin, taps] (torch layout), Cout, Cin, taps, scale fp32 [Cout] | 0,
// fwd dst [Cout, taps*Cin] (K order tap, cin -- the im2col / implicit-GEMM order), bwd dst | 0, bwd kind, bwd pitch, first element.
//   bwd kind 1: plain transpose [taps*Cin, pitch >= Cout]  (dgrad of 1x1 convs; dcol = dY . W' for the strided convs)
//   bwd kind 2: flipped taps    [Cin, taps*Cout]: dst[ci][(taps-1-t)*Cout + co]  (stride-1 3x3 dgrad as a convolution over dY)
template <typename T>
__global__ void __launch_bounds__(256)
pack_conv_kernel(const long long* __restrict__ table, int n_entries, long long total) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = n_entries - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (table[(size_t)mid * 10 + 9] <= e) lo = mid; else hi = mid - 1;
        }
        const long long* t = table + (size_t)lo * 10;
        const float* src = reinterpret_cast<const float*>(t[0]);
        const int Cin = (int)t[2], taps = (int)t[3];
        const int Cout = (int)t[1];
        const float* scale = reinterpret_cast<const float*>(t[4]);
        T* fwd = reinterpret_cast<T*>(t[5]);
        T* bwd = reinterpret_cast<T*>(t[6]);
        const long long el = e - t[9];
        const int tap = (int)(el % taps);
        const int ci = (int)((el / taps) % Cin);
        const int co = (int)(el / ((long long)taps * Cin));
        const float v = src[el] * (scale ? scale[co] : 1.f);
        st1<T>(fwd + (size_t)co * taps * Cin + (size_t)tap * Cin + ci, v);
        if (bwd) {
            if (t[7] == 1) st1<T>(bwd + ((size_t)tap * Cin + ci) * t[8] + co, v);
            else st1<T>(bwd + (size_t)ci * t[8] + (size_t)(taps - 1 - tap) * Cout + co, v);
        }
    }
}

// The inverse for the weight GRADIENTS: dtlr_wgrad produced dW' in the forward operand layout [Cout, taps*Cin] (a zeroed scratch arena);
// the parameter gradient is d w[co][ci][t] += scale[co] * dW'[co][t*Cin + ci].  Entry (6 x int64): scratch ptr, grad ptr, Cout, Cin, taps,
// scale | 0; elem_start prefix in entry[...] is passed separately.
__global__ void __launch_bounds__(256)
unpack_conv_grads_kernel(const long long* __restrict__ table, const long long* __restrict__ elem_start, int n_entries, long long total) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        int lo = 0, hi = n_entries - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (elem_start[mid] <= e) lo = mid; else hi = mid - 1;
        }
        const long long* t = table + (size_t)lo * 6;
        const float* tmp = reinterpret_cast<const float*>(t[0]);
        float* grad = reinterpret_cast<float*>(t[1]);
        const int Cin = (int)t[3], taps = (int)t[4];
        const float* scale = reinterpret_cast<const float*>(t[5]);
        const long long el = e - elem_start[lo];
        const int tap = (int)(el % taps);
        const int ci = (int)((el / taps) % Cin);
        const int co = (int)(el / ((long long)taps * Cin));
        grad[el] += tmp[(size_t)co * taps * Cin + (size_t)tap * Cin + ci] * (scale ? scale[co] : 1.f);
    }
}

// ---------------------------------------------------------------------------------------------- clip_grad_norm_ + AdamW
// state (device, fp32[4]): [0] = sum of squares of all gradients of the step, [1] = step count.
__global__ void optim_begin_kernel(float* state) {
    pdl_launch_dependents();
    pdl_wait();
    state[0] = 0.f;
    state[1] += 1.f;
}
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, long long n, float* __restrict__ state) {
    __shared__ float ws[8];
    pdl_launch_dependents();
    pdl_wait();
    float s = 0.f;
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(g)[i];
        s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float v = g[(n4 << 2) + threadIdx.x]; s = fmaf(v, v, s); }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += ws[w];
        atomicAdd(state, t);
    }
}
// torch.nn.utils.clip_grad_norm_(max_norm) (engine.py:238-239: coefficient = min(1, max_norm / (total_norm + 1e-6))) followed by
// torch.optim.AdamW.step (decoupled weight decay, bias-corrected moments; finetuning.py:227-231), one pass over the arena slice.
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n, float lr, float b1,
             float b2, float eps, float wd, float max_norm, const float* __restrict__ state) {
    pdl_launch_dependents();
    pdl_wait();
    const float total_norm = sqrtf(state[0]);
    const float clip = max_norm > 0.f ? fminf(1.f, max_norm / (total_norm + 1e-6f)) : 1.f;
    const float step = state[1];
    const float bc1 = 1.f - powf(b1, step), bc2 = 1.f - powf(b2, step);
    const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * clip;
        float pi = p[i] * (1.f - lr * wd);
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
        pi -= step_size * (mi / denom);
        p[i] = pi;
    }
}

}  // namespace train
}  // namespace dtlr

using namespace dtlr;
using namespace dtlr::train;

static inline unsigned grid_cap(long long work_items, int block, int per_sm) {
    long long g = (work_items + block - 1) / block;
    const long long cap = (long long)sm_count() * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

extern "C" int dtlr_wgrad(const void* dY, int ldy, const void* X, int ldx, float* dW, int ldw, int rows, int N, int K, int dtype,
                          void* stream) {
    DTLR_CHECK_ARG(rows >= 0 && N > 0 && K > 0, "wgrad: bad sizes rows=%d N=%d K=%d", rows, N, K);
    if (rows == 0) return DTLR_OK;
    DTLR_CHECK_ARG(dY && X && dW, "wgrad: null pointer");
    DTLR_CHECK_ARG(ldy >= N && ldx >= K && ldw >= K, "wgrad: leading dimension too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DTLR_F32) {
        const int tiles = ((N + 63) / 64) * ((K + 63) / 64);
        int splits = (2 * sm_count() + tiles - 1) / tiles;
        const int max_splits = (rows + 127) / 128;
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
        int rps = (rows + splits - 1) / splits;
        rps = (rps + 31) / 32 * 32;
        splits = (rows + rps - 1) / rps;
        dim3 grid((N + 63) / 64, (K + 63) / 64, splits);
        DTLR_LAUNCH(wgrad_f32_kernel, grid, 256, 0, st, (const float*)dY, ldy, (const float*)X, ldx, dW, ldw, N, K, rows, rps);
        DTLR_CHECK_LAUNCH();
        return DTLR_OK;
    }
    DTLR_CHECK_ARG(dtype == DTLR_OP16, "wgrad: operands must be f32 or the library's 16-bit type");
    DTLR_CHECK_ARG((ldy % 8) == 0 && (ldx % 8) == 0 && ((((uintptr_t)dY | (uintptr_t)X)) & 15) == 0,
                   "wgrad: 16-bit operands need 16-byte aligned rows (ldy=%d ldx=%d)", ldy, ldx);
    CUtensorMap ta, tb;
    int rc;
    if ((rc = make_tmap_2d_bf16(&ta, dY, rows, N, ldy, WG_RB, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    if ((rc = make_tmap_2d_bf16(&tb, X, rows, K, ldx, WG_RB, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
    const int tiles = ((N + WG_BM - 1) / WG_BM) * ((K + WG_BN - 1) / WG_BN);
    const int total_rb = (rows + WG_RB - 1) / WG_RB;
    int splits = (2 * sm_count() + tiles - 1) / tiles;                 // ~2 waves of CTAs, at least 4 pipeline stages of work each
    const int max_splits = (total_rb + 3) / 4;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    const int rbps = (total_rb + splits - 1) / splits;
    splits = (total_rb + rbps - 1) / rbps;
    static bool configured = false;
    if (!configured) {
        DTLR_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
        configured = true;
    }
    dim3 grid((N + WG_BM - 1) / WG_BM, (K + WG_BN - 1) / WG_BN, splits);
    DTLR_LAUNCH(wgrad_tcgen05_kernel, grid, 192, WG_SMEM, st, ta, tb, dW, ldw, N, K, rows, rbps, (g_debug_flags & 134217728) ? 1 : 0);
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_colsum(const void* x, long long ld, int N, long long nseg, long long seg_rows, long long seg_stride, float* out,
                           int dtype, void* stream) {
    DTLR_CHECK_ARG(N > 0 && nseg >= 0 && seg_rows >= 0 && ld >= N, "colsum: bad sizes");
    const long long total = nseg * seg_rows;
    if (total == 0) return DTLR_OK;
    DTLR_CHECK_ARG(total < (1ll << 31) && nseg < (1ll << 31), "colsum: more than 2^31 rows");
    DTLR_CHECK_ARG((dtype == DTLR_F32 ? (ld % 4) == 0 : (ld % 8) == 0) && (((uintptr_t)x) & 15) == 0, "colsum: rows must be 16-byte aligned");
    const int cg_all = (N + 7) / 8;
    const int ny = (cg_all + 255) / 256;
    const int rp = 256 / (cg_all < 256 ? cg_all : 256);
    long long blocks = (long long)sm_count() * 2 / ny;                 // one wave; every CTA ends with N / 4 vector reductions
    if (blocks < 1) blocks = 1;
    long long rpb = (total + blocks - 1) / blocks;
    if (rpb < 16LL * rp) rpb = 16LL * rp;
    blocks = (total + rpb - 1) / rpb;
    dim3 grid((unsigned)blocks, ny);
    DISPATCH_T(dtype, DTLR_LAUNCH((colsum_kernel<T>), grid, 256, 0, (cudaStream_t)stream, (const T*)x, ld, N, (int)nseg, (int)seg_rows, seg_stride, out, (int)rpb);)
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_layernorm_bwd(const void* z, const float* dy, const float* dy2, const float* gamma, float* dz32, void* dz16,
                                  float* dgamma, float* dbeta, long long rows, int C, float eps, int dtype, void* stream) {
    DTLR_CHECK_ARG(C == 256, "layernorm_bwd: only C=256 (d_model of every DTLR config) is implemented, got %d", C);
    if (rows == 0) return DTLR_OK;
    const unsigned grid = grid_cap(rows, 8, 4);
    DISPATCH_T(dtype, DTLR_LAUNCH((layernorm256_bwd_kernel<T>), grid, 256, 0, (cudaStream_t)stream, (const T*)z, dy, dy2, gamma, dz32, (T*)dz16, dgamma, dbeta, rows, eps);)
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_relu_bwd(void* dh, const void* h, long long n, int dtype, void* stream) {
    DTLR_CHECK_ARG((n % 8) == 0, "relu_bwd: element count must be a multiple of 8");
    if (n == 0) return DTLR_OK;
    DISPATCH_T(dtype, DTLR_LAUNCH((relu_bwd_kernel<T>), grid_cap(n / 8, 256, 16), 256, 0, (cudaStream_t)stream, (T*)dh, (const T*)h, n / 8);)
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_add_cast(const float* a, const float* b, const float* c, void* out, long long n, int out_dtype, void* stream) {
    DTLR_CHECK_ARG((n % 8) == 0, "add_cast: element count must be a multiple of 8");
    if (n == 0) return DTLR_OK;
    DISPATCH_T(out_dtype, DTLR_LAUNCH((add_cast_kernel<T>), grid_cap(n / 8, 256, 16), 256, 0, (cudaStream_t)stream, a, b, c, (T*)out, n / 8);)
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_msda_bwd_glue(const float* grad_loc, const float* grad_attn, const float* attn, const float* ref, int ref_dim,
                                  const float* valid_ratios, const int64_t* shapes, int L, void* dproj, int ld, int B, int Lq, int M, int P,
                                  int out_dtype, void* stream) {
    DTLR_CHECK_ARG(ref_dim == 2 || ref_dim == 4, "msda_bwd_glue: reference points must have 2 or 4 coordinates");
    DTLR_CHECK_ARG(L >= 1 && L <= 8 && ld >= M * L * P * 3, "msda_bwd_glue: bad level count / row pitch");
    GlueLevels lv;
    lv.n = L;
    for (int l = 0; l < L; ++l) { lv.H[l] = (int)shapes[2 * l]; lv.W[l] = (int)shapes[2 * l + 1]; }
    const long long total = (long long)B * Lq * M;
    if (total == 0) return DTLR_OK;
    const bool rows16 = out_dtype == DTLR_F32 ? ((ld % 4) == 0) : ((ld % 8) == 0);
    if (L == 4 && P == 4 && rows16 && ((((uintptr_t)dproj | (uintptr_t)grad_loc | (uintptr_t)grad_attn | (uintptr_t)attn)) & 15) == 0 &&
        !(g_debug_flags & 268435456)) {                  // flag 268435456: the generic kernel (A/B)
        DISPATCH_T(out_dtype, DTLR_LAUNCH((msda_bwd_glue16_kernel<T, 4, 4>), (unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream, grad_loc, grad_attn, attn, ref, ref_dim, valid_ratios, lv, (T*)dproj, ld, B, Lq, M);)
        DTLR_CHECK_LAUNCH();
        return DTLR_OK;
    }
    DISPATCH_T(out_dtype, DTLR_LAUNCH((msda_bwd_glue_kernel<T>), (unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream, grad_loc, grad_attn, attn, ref, ref_dim, valid_ratios, lv, (T*)dproj, ld, B, Lq, M, P);)
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_pack_weights(const long long* table, const int* tile_start, int n_entries, int total_tiles, int dtype, void* stream) {
    if (n_entries <= 0 || total_tiles <= 0) return DTLR_OK;
    DISPATCH_T(dtype, DTLR_LAUNCH((pack_weights_kernel<T>), (unsigned)total_tiles, 256, 0, (cudaStream_t)stream, table, tile_start, n_entries);)
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_optim_begin(float* state, void* stream) {
    DTLR_LAUNCH(optim_begin_kernel, 1, 1, 0, (cudaStream_t)stream, state);
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}
extern "C" int dtlr_grad_sumsq(const float* g, long long n, float* state, void* stream) {
    if (n == 0) return DTLR_OK;
    DTLR_CHECK_ARG((((uintptr_t)g) & 15) == 0, "grad_sumsq: the gradient arena must be 16-byte aligned");
    DTLR_LAUNCH(sumsq_kernel, grid_cap(n / 4 + 1, 256, 8), 256, 0, (cudaStream_t)stream, g, n, state);
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}
extern "C" int dtlr_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                          float weight_decay, float max_norm, const float* state, void* stream) {
    if (n == 0) return DTLR_OK;
    DTLR_LAUNCH(adamw_kernel, grid_cap(n, 256, 16), 256, 0, (cudaStream_t)stream, p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, max_norm, state);
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_relu_bwd_dual(float* dy32, const void* y, void* out16, long long n, int dtype, void* stream) {
    DTLR_CHECK_ARG((n % 8) == 0, "relu_bwd_dual: element count must be a multiple of 8");
    if (n == 0) return DTLR_OK;
    DISPATCH_T(dtype, DTLR_LAUNCH((relu_bwd_dual_kernel<T>), grid_cap(n / 8, 256, 16), 256, 0, (cudaStream_t)stream, dy32, (const T*)y, (T*)out16, n / 8);)
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_groupnorm_bwd(const float* x, const float* dy, long long dy_stride_b, const float* gamma, void* dx, float* dgamma,
                                  float* dbeta, int B, int HW, int C, int G, float eps, int out_dtype, void* stream) {
    DTLR_CHECK_ARG(C % G == 0 && C / G == 8 && (C % 32) == 0, "groupnorm_bwd: 8 channels per group (GroupNorm(32, 256)) only, got C=%d G=%d", C, G);
    if (B == 0 || HW == 0) return DTLR_OK;
    dim3 grid(C / 32, B);
    DISPATCH_T(out_dtype, DTLR_LAUNCH((groupnorm8_bwd_kernel<T>), grid, 256, 0, (cudaStream_t)stream, x, dy, dy_stride_b, gamma, (T*)dx, dgamma, dbeta, HW, C, eps);)
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_col2im(const float* dcol, int ldc, float* dx, int B, int H, int W, int C, int KH, int KW, int stride, int pad, int Ho,
                           int Wo, int accumulate, void* stream) {
    DTLR_CHECK_ARG((C % 4) == 0 && (ldc % 4) == 0 && ldc >= KH * KW * C && stride >= 1, "col2im: bad sizes");
    const long long total = (long long)B * H * W * (C / 4);
    if (total == 0) return DTLR_OK;
    DTLR_LAUNCH(col2im_kernel, grid_cap(total, 256, 16), 256, 0, (cudaStream_t)stream, dcol, ldc, dx, B, H, W, C, KH, KW, stride, pad, Ho, Wo, accumulate);
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_pack_conv(const long long* table, int n_entries, long long total_elems, int dtype, void* stream) {
    if (n_entries <= 0 || total_elems <= 0) return DTLR_OK;
    DISPATCH_T(dtype, DTLR_LAUNCH((pack_conv_kernel<T>), grid_cap(total_elems, 256, 16), 256, 0, (cudaStream_t)stream, table, n_entries, total_elems);)
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_unpack_conv_grads(const long long* table, const long long* elem_start, int n_entries, long long total_elems, void* stream) {
    if (n_entries <= 0 || total_elems <= 0) return DTLR_OK;
    DTLR_LAUNCH(unpack_conv_grads_kernel, grid_cap(total_elems, 256, 16), 256, 0, (cudaStream_t)stream, table, elem_start, n_entries, total_elems);
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}
