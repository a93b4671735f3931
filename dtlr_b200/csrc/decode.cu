// dtlr_b200 -- the "CTC view" decode tail, fused (reference models/dino/dino.py:472-502 + engine.py:512-530):
//   sort the queries of each line by box centre x, sigmoid the class logits, synthesise the blank probability
//   (eps = 0.003 in the training/eval loop, 0.03/C in evaluation.py:141), argmax over [blank, classes].
// The reference materialises three (B,Q,C+1) tensors (53 MB/image at C=7356); here one warp reduces a query row to a
// label in registers, and one CTA per line sorts the (cx, query) pairs in shared memory (bitonic, index tie-break) and
// emits the labels in reading order.  Optionally also writes new_pred_logits (B,Q,C+1) for callers that need it
// (n-gram rescoring).
#include "common.cuh"

namespace dtlr {

__device__ __forceinline__ float warp_sum_d(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per (b,q): label = 0 (blank) or 1 + argmax class.  HBM bound (B*Q*C*4 bytes read once; 848 MB at C = 7356, B = 32):
// VEC = 4 reads the row with 16-byte loads, two per lane in flight per iteration (the scalar version reached 26 % of the copy
// peak at C = 7356: too few bytes in flight per warp); rows are 16-byte aligned when ld % 4 == 0.  Exact expf: the blank / class
// decision is compared bit-for-bit with the reference's argmax.
template <int VEC>
__global__ void __launch_bounds__(256)
ctc_row_label_kernel(const float* __restrict__ logits, int ld, int C, float eps, float pscale,
                     int* __restrict__ label, float* __restrict__ row_sum, long long rows) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float* x = logits + (size_t)row * ld;
    float s = 0.f, best = -1.f;
    int arg = 0x7fffffff;
    auto take = [&](const float xv, const int c) {
        const float p = pscale * (1.f / (1.f + expf(-xv)));
        s += p;
        if (p > best) { best = p; arg = c; }          // strict: the first maximum of this lane's ascending indices wins
    };
    int c0 = 0;
    if (VEC == 4) {
        const int n4 = C >> 2;
        const float4* x4 = reinterpret_cast<const float4*>(x);
        int i = lane;
        for (; i + 32 < n4; i += 64) {                // two independent 16-byte loads per lane per iteration
            const float4 a = __ldg(x4 + i), b = __ldg(x4 + i + 32);
            take(a.x, 4 * i); take(a.y, 4 * i + 1); take(a.z, 4 * i + 2); take(a.w, 4 * i + 3);
            take(b.x, 4 * i + 128); take(b.y, 4 * i + 129); take(b.z, 4 * i + 130); take(b.w, 4 * i + 131);
        }
        if (i < n4) {
            const float4 a = __ldg(x4 + i);
            take(a.x, 4 * i); take(a.y, 4 * i + 1); take(a.z, 4 * i + 2); take(a.w, 4 * i + 3);
        }
        c0 = n4 << 2;                                  // scalar tail: C % 4 classes
    }
    for (int c = c0 + lane; c < C; c += 32) take(x[c], c);
    s = warp_sum_d(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }     // first maximum wins, like argmax
    }
    if (lane == 0) {
        const bool low = s < 1.f - eps;
        const float blank = low ? 1.f - s : eps;
        const float top = low ? best : (1.f - eps) * best / s;
        label[row] = (blank >= top) ? 0 : arg + 1;
        if (row_sum) row_sum[row] = s;
    }
}

// one CTA per line: bitonic sort of (cx, q) ascending (ties by q), then frames[b,pos] = label[b, perm[pos]]
__global__ void ctc_sort_emit_kernel(const float* __restrict__ boxes, const int* __restrict__ label, int* __restrict__ frames,
                                     int* __restrict__ perm_out, int Q, int n_pow2) {
    extern __shared__ unsigned char dsm[];
    float* key = reinterpret_cast<float*>(dsm);
    int* val = reinterpret_cast<int*>(key + n_pow2);
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        key[i] = i < Q ? boxes[((size_t)b * Q + i) * 4] : INFINITY;
        val[i] = i < Q ? i : 0x7fffffff;
    }
    __syncthreads();
    for (int k = 2; k <= n_pow2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const float ka = key[i], kb = key[p];
                    const int va = val[i], vb = val[p];
                    const bool a_gt_b = (ka > kb) || (ka == kb && va > vb);
                    const bool asc = (i & k) == 0;
                    if (a_gt_b == asc) { key[i] = kb; key[p] = ka; val[i] = vb; val[p] = va; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < Q; i += blockDim.x) {
        const int q = val[i];
        frames[(size_t)b * Q + i] = label[(size_t)b * Q + q];
        if (perm_out) perm_out[(size_t)b * Q + i] = q;
    }
}

// optional: new_pred_logits[b, pos, :] from logits[b, perm[pos], :]   (one warp per output row)
__global__ void ctc_new_pred_kernel(const float* __restrict__ logits, int ld, int C, float eps, float pscale,
                                    const int* __restrict__ perm, const float* __restrict__ row_sum,
                                    float* __restrict__ new_pred, int Q, long long rows) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const long long b = row / Q;
    const long long src = b * Q + perm[row];
    const float s = row_sum[src];
    const bool low = s < 1.f - eps;
    const float scale = low ? 1.f : (1.f - eps) / s;
    const float* x = logits + (size_t)src * ld;
    float* o = new_pred + (size_t)row * (C + 1);
    if (lane == 0) o[0] = low ? 1.f - s : eps;
    for (int c = lane; c < C; c += 32) {
        const float p = pscale * (1.f / (1.f + expf(-x[c])));
        o[c + 1] = low ? p : (1.f - eps) * p / s;
    }
    (void)scale;
}

}  // namespace dtlr

using namespace dtlr;

extern "C" int dtlr_ctc_decode_scaled(const float* logits, int ld, const float* boxes, int* frames, int* perm, float* new_pred,
                                      int* scratch_label, float* scratch_sum, int B, int Q, int C, float eps, float prob_scale,
                                      void* stream) {
    DTLR_CHECK_ARG(B >= 0 && Q >= 0 && C > 0 && ld >= C, "ctc_decode: bad sizes");
    if (B == 0 || Q == 0) return DTLR_OK;
    DTLR_CHECK_ARG(logits && boxes && frames && scratch_label, "ctc_decode: null pointer");
    DTLR_CHECK_ARG(!new_pred || (perm && scratch_sum), "ctc_decode: new_pred needs perm and scratch_sum buffers");
    int n = 1;
    while (n < Q) n <<= 1;
    DTLR_CHECK_ARG((size_t)n * 8 <= (size_t)max_smem_optin(), "ctc_decode: %d queries per line exceed the shared-memory sort", Q);
    cudaStream_t st = (cudaStream_t)stream;
    const long long rows = (long long)B * Q;
    // 16-byte-load kernel whenever the rows are 16-byte aligned (dtlr_debug_flags(32768): scalar-load kernel, A/B).  Measured at
    // C = 7356, B = 32 (848 MB of logits, L2 flushed): 497 -> 208 us = 4.07 TB/s = 63 % of the measured copy peak
    if (!(g_debug_flags & 32768) && (ld % 4) == 0 && (((uintptr_t)logits) & 15) == 0 && C >= 4)
        ctc_row_label_kernel<4><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(logits, ld, C, eps, prob_scale, scratch_label, scratch_sum, rows);
    else
        ctc_row_label_kernel<1><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(logits, ld, C, eps, prob_scale, scratch_label, scratch_sum, rows);
    DTLR_CHECK_LAUNCH();
    const size_t smem = (size_t)n * 8;
    if (smem > 48 * 1024)
        DTLR_CHECK_CUDA(cudaFuncSetAttribute(ctc_sort_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctc_sort_emit_kernel<<<B, n < 1024 ? n : 1024, smem, st>>>(boxes, scratch_label, frames, perm, Q, n);
    DTLR_CHECK_LAUNCH();
    if (new_pred) {
        ctc_new_pred_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(logits, ld, C, eps, prob_scale, perm, scratch_sum, new_pred, Q, rows);
        DTLR_CHECK_LAUNCH();
    }
    return DTLR_OK;
}

extern "C" int dtlr_ctc_decode(const float* logits, int ld, const float* boxes, int* frames, int* perm, float* new_pred,
                               int* scratch_label, float* scratch_sum, int B, int Q, int C, float eps, void* stream) {
    return dtlr_ctc_decode_scaled(logits, ld, boxes, frames, perm, new_pred, scratch_label, scratch_sum, B, Q, C, eps, 1.f, stream);
}
