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

// =====================================================================================================================
// Fused CTC loss of the fine-tuning step (reference models/dino/dino.py:457-551): sort by cx -> sigmoid -> blank synthesis ->
// a hard-blank frame interleaved after every query -> nn.CTCLoss(blank=0, zero_infinity=True, reduction='mean') on the log.
// The reference materialises new_pred_logits (B,Q,C+1), the interleaved (B,2Q,C+1) tensor and its log (53 MB per image at
// C = 7356) and lets autograd walk back through all of them.  Here the (B,2Q,C+1) tensor never exists:
//   1. ctc_row_label_kernel + ctc_sort_emit_kernel (above): row sums and the cx permutation;
//   2. ctc_lp_kernel:      the only probabilities a CTC lattice reads -- blank and the line's L target labels -- per frame, as logs
//                          (B,Q,L+1 instead of B,2Q,C+1);
//   3. ctc_lattice_kernel: alpha and beta recursions over the 2Q frames (the odd, hard-blank frames are constants), one CTA per
//                          line, one thread per lattice position, one barrier per frame; emits -log p(target) and
//                          d(-log p)/d(new_pred[frame, position]);
//   4. ctc_grad_kernel:    chain rule through the blank synthesis and the sigmoid, written densely to grad_logits (B,Q,C) in the
//                          ORIGINAL query order (the sort is an index permutation; boxes get no gradient, as in the reference).
// torch's ctc_loss returns d/dlog_probs under the assumption that log_probs are log-softmax outputs (it adds exp(lp)); that extra
// term cancels exactly through the normalised rows of new_pred (their sum is constant in the logits), so the gradient written here
// equals what autograd produces for the reference's chain.
namespace dtlr {

#define CTC_NEG_INF (-INFINITY)

__device__ __forceinline__ float lse3(float a, float b, float c) {
    const float m = fmaxf(a, fmaxf(b, c));
    if (m == CTC_NEG_INF) return CTC_NEG_INF;
    return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

// lp[(b*Q + i)*(Lmax+1) + 0] = log blank prob of the i-th frame in reading order, [.. + 1 + k] = log prob of target label k
__global__ void ctc_lp_kernel(const float* __restrict__ logits, int ld, const int* __restrict__ perm, const float* __restrict__ row_sum,
                              const int* __restrict__ targets, const int* __restrict__ tlen, int Lmax, float eps,
                              float* __restrict__ lp, int Q) {
    const int b = blockIdx.y, i = blockIdx.x;
    const int q = perm[(size_t)b * Q + i];
    const float s = row_sum[(size_t)b * Q + q];
    const bool low = s < 1.f - eps;
    const int L = min(tlen[b], Lmax);
    float* o = lp + ((size_t)b * Q + i) * (Lmax + 1);
    const float* x = logits + ((size_t)b * Q + q) * ld;
    for (int k = threadIdx.x; k <= L; k += blockDim.x) {
        float y;
        if (k == 0) {
            y = low ? 1.f - s : eps;
        } else {
            const float p = 1.f / (1.f + expf(-x[targets[(size_t)b * Lmax + k - 1]]));
            y = low ? p : (1.f - eps) * p / s;
        }
        o[k] = logf(y);
    }
}

// One CTA per line; thread s = lattice position (even: blank, odd: target label (s-1)/2), S = 2L+1 <= blockDim.x.
// alpha_even [B, Q, Sp] scratch (alpha at the query frames), gext [B, Q, Sp] out: d(-log p)/d new_pred[frame i, class of position s]
// (0 when the path mass is 0), nll [B] out.  The frame loop is sequential (2Q steps, one __syncthreads each); lp columns are
// prefetched 8 query frames ahead so no step waits for HBM.
__global__ void __launch_bounds__(1024)
ctc_lattice_kernel(const float* __restrict__ lp, const int* __restrict__ targets, const int* __restrict__ tlen, int Lmax, int Q,
                   int Sp, float* __restrict__ alpha_even, float* __restrict__ gext, float* __restrict__ nll) {
    extern __shared__ float sm[];          // 2 buffers of (blockDim.x + 4): two -inf guard cells on each side
    const int b = blockIdx.x, s = threadIdx.x, nthr = blockDim.x;
    const int L = min(tlen[b], Lmax), S = 2 * L + 1;
    const bool valid = s < S, is_lab = (s & 1) != 0;
    const int k = (s - 1) >> 1;
    const int* tg = targets + (size_t)b * Lmax;
    const int lab = (valid && is_lab) ? tg[k] : -1;
    const bool skip_dn = valid && is_lab && s >= 3 && tg[k - 1] != lab;            // alpha: may come from s-2
    const bool skip_up = valid && is_lab && s + 2 < S && tg[k + 1] != lab;         // beta: may go to s+2
    const int col = is_lab ? 1 + k : 0;
    const float lp_odd = is_lab ? logf(1e-5f) : 0.f;                               // the hard-blank frame: (1, 1e-5, 1e-5, ...)
    const int pitch = nthr + 4;
    float* buf0 = sm + 2;
    float* buf1 = sm + pitch + 2;
    for (int i = s; i < 2 * pitch; i += nthr) sm[i] = CTC_NEG_INF;
    __syncthreads();
    const float* lpb = lp + (size_t)b * Q * (Lmax + 1) + col;
    float* ae = alpha_even + (size_t)b * Q * Sp + s;
    float* ge = gext + (size_t)b * Q * Sp + s;
    const int T = 2 * Q;
    constexpr int PF = 8;
    float cur[PF], nxt[PF];

    // ------------------------------------------------------------------ alpha
#pragma unroll
    for (int j = 0; j < PF; ++j) nxt[j] = (valid && j < Q) ? lpb[(size_t)j * (Lmax + 1)] : CTC_NEG_INF;
    float* prev = buf0;
    float* next = buf1;
    for (int i0 = 0; i0 < Q; i0 += PF) {
#pragma unroll
        for (int j = 0; j < PF; ++j) cur[j] = nxt[j];
#pragma unroll
        for (int j = 0; j < PF; ++j) {
            const int i = i0 + PF + j;
            nxt[j] = (valid && i < Q) ? lpb[(size_t)i * (Lmax + 1)] : CTC_NEG_INF;
        }
#pragma unroll
        for (int j = 0; j < PF; ++j) {
            const int i = i0 + j;
            if (i >= Q) break;
            // query frame t = 2i
            float a;
            if (i == 0) a = (s <= 1 && valid) ? cur[j] : CTC_NEG_INF;
            else a = valid ? cur[j] + lse3(prev[s], prev[s - 1], skip_dn ? prev[s - 2] : CTC_NEG_INF) : CTC_NEG_INF;
            next[s] = a;
            if (valid) ae[(size_t)i * Sp] = a;
            __syncthreads();
            { float* t_ = prev; prev = next; next = t_; }
            // hard-blank frame t = 2i + 1
            a = valid ? lp_odd + lse3(prev[s], prev[s - 1], skip_dn ? prev[s - 2] : CTC_NEG_INF) : CTC_NEG_INF;
            next[s] = a;
            __syncthreads();
            { float* t_ = prev; prev = next; next = t_; }
        }
    }
    // log p(target) = lse(alpha_{T-1}(S-1), alpha_{T-1}(S-2))
    const float logp = lse3(prev[S - 1], S >= 2 ? prev[S - 2] : CTC_NEG_INF, CTC_NEG_INF);
    __syncthreads();
    if (s == 0) nll[b] = -logp;
    const bool feasible = logp > CTC_NEG_INF;

    // ------------------------------------------------------------------ beta (frames T-1 .. 0), gradient at the query frames
    for (int i = s; i < 2 * pitch; i += nthr) sm[i] = CTC_NEG_INF;
    __syncthreads();
    prev = buf0;           // holds beta_{t+1}
    next = buf1;
#pragma unroll
    for (int j = 0; j < PF; ++j) {
        const int i = Q - 1 - j;
        nxt[j] = (valid && i >= 0) ? lpb[(size_t)i * (Lmax + 1)] : CTC_NEG_INF;
    }
    for (int i0 = Q - 1; i0 >= 0; i0 -= PF) {
#pragma unroll
        for (int j = 0; j < PF; ++j) cur[j] = nxt[j];
#pragma unroll
        for (int j = 0; j < PF; ++j) {
            const int i = i0 - PF - j;
            nxt[j] = (valid && i >= 0) ? lpb[(size_t)i * (Lmax + 1)] : CTC_NEG_INF;
        }
#pragma unroll
        for (int j = 0; j < PF; ++j) {
            const int i = i0 - j;
            if (i < 0) break;
            // hard-blank frame t = 2i + 1
            float be;
            if (i == Q - 1) be = (valid && s >= S - 2) ? lp_odd : CTC_NEG_INF;
            else be = valid ? lp_odd + lse3(prev[s], prev[s + 1], skip_up ? prev[s + 2] : CTC_NEG_INF) : CTC_NEG_INF;
            next[s] = be;
            __syncthreads();
            { float* t_ = prev; prev = next; next = t_; }
            // query frame t = 2i
            const float lpq = cur[j];
            be = valid ? lpq + lse3(prev[s], prev[s + 1], skip_up ? prev[s + 2] : CTC_NEG_INF) : CTC_NEG_INF;
            next[s] = be;
            if (valid) {
                const float al = ae[(size_t)i * Sp];
                float g = 0.f;
                if (feasible && al > CTC_NEG_INF && be > CTC_NEG_INF && lpq > CTC_NEG_INF) g = -expf(al + be - 2.f * lpq - logp);
                ge[(size_t)i * Sp] = g;
            }
            __syncthreads();
            { float* t_ = prev; prev = next; next = t_; }
        }
    }
    (void)T;
}

// one CTA per (frame i, line b): grad_logits[b, perm[i], :] = w_b * d(-log p)/d logit through the blank synthesis and the sigmoid
__global__ void __launch_bounds__(256)
ctc_grad_kernel(const float* __restrict__ logits, int ld, const int* __restrict__ perm, const float* __restrict__ row_sum,
                const int* __restrict__ targets, const int* __restrict__ tlen, int Lmax, int Sp, float eps,
                const float* __restrict__ gext, const float* __restrict__ nll, int zero_infinity, float* __restrict__ grad,
                int Q, int C, int B) {
    extern __shared__ float gsm[];         // G[0..L]: summed d(-log p)/d new_pred per class slot (0 = blank), then pk[1..L]
    __shared__ float red[8];
    const int b = blockIdx.y, i = blockIdx.x, tid = threadIdx.x;
    const int q = perm[(size_t)b * Q + i];
    const float s = row_sum[(size_t)b * Q + q];
    const bool low = s < 1.f - eps;
    const int L = min(tlen[b], Lmax), S = 2 * L + 1;
    const float* ge = gext + ((size_t)b * Q + i) * Sp;
    const float* x = logits + ((size_t)b * Q + q) * ld;
    const int* tg = targets + (size_t)b * Lmax;
    float* G = gsm;
    float* pk = gsm + (Lmax + 1);
    float part = 0.f;
    for (int p = tid; p < S; p += blockDim.x) {
        const float g = ge[p];
        if (p & 1) {
            const int k = (p - 1) >> 1;
            G[1 + k] = g;
            pk[1 + k] = 1.f / (1.f + expf(-x[tg[k]]));
        } else part += g;
    }
    part = warp_sum_d(part);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    float g0 = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) g0 += red[w];
    const float nl = nll[b];
    const bool dead = !(nl < INFINITY) && zero_infinity;            // infeasible alignment: loss and gradient are zeroed
    const float wb = dead ? 0.f : 1.f / ((float)max(L, 1) * (float)B);
    // low:  new_0 = 1 - sum p, new_c = p_c            -> dL/dp_c = G_c - G_0
    // high: new_0 = eps,      new_c = (1-eps) p_c / s -> dL/dp_c = (1-eps)/s * (G_c - sum_k G_k p_k / s)
    float dense, f;
    if (low) { dense = -g0; f = 1.f; }
    else {
        float kk = 0.f;
        for (int k = 1; k <= L; ++k) kk += G[k] * pk[k];            // L <= ~100 terms; every thread computes the same sum
        f = (1.f - eps) / s;
        dense = -kk / s;
    }
    float* go = grad + ((size_t)b * Q + q) * C;
    for (int c = tid; c < C; c += blockDim.x) {
        const float p = 1.f / (1.f + expf(-x[c]));
        go[c] = wb * f * dense * p * (1.f - p);
    }
    __syncthreads();
    for (int k = 1 + tid; k <= L; k += blockDim.x) {
        const float p = pk[k];
        atomicAdd(go + tg[k - 1], wb * f * G[k] * p * (1.f - p));   // repeated labels of a line add up
    }
}

}  // namespace dtlr

extern "C" int dtlr_ctc_loss(const float* logits, int ld, const float* boxes, const int* targets, const int* target_len, int Lmax,
                             float eps, int zero_infinity, float* nll, float* grad_logits, int* perm, float* row_sum,
                             int* scratch_label, int* scratch_frames, float* lp, float* alpha, float* gext, int B, int Q, int C,
                             void* stream) {
    DTLR_CHECK_ARG(B >= 0 && Q > 0 && C > 0 && ld >= C && Lmax >= 0, "ctc_loss: bad sizes");
    if (B == 0) return DTLR_OK;
    DTLR_CHECK_ARG(logits && boxes && target_len && nll && perm && row_sum && scratch_label && scratch_frames && lp && alpha && gext,
                   "ctc_loss: null pointer");
    DTLR_CHECK_ARG(Lmax == 0 || targets, "ctc_loss: null targets");
    const int S = 2 * Lmax + 1;
    DTLR_CHECK_ARG(S <= 1024, "ctc_loss: target length %d exceeds the one-thread-per-lattice-position kernel (max 511)", Lmax);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = dtlr_ctc_decode_scaled(logits, ld, boxes, scratch_frames, perm, nullptr, scratch_label, row_sum, B, Q, C, eps, 1.f, stream);
    if (rc != DTLR_OK) return rc;
    const int Sp = S;
    ctc_lp_kernel<<<dim3(Q, B), 128, 0, st>>>(logits, ld, perm, row_sum, targets, target_len, Lmax, eps, lp, Q);
    DTLR_CHECK_LAUNCH();
    const int nthr = (S + 31) / 32 * 32;
    ctc_lattice_kernel<<<B, nthr, 2 * (nthr + 4) * sizeof(float), st>>>(lp, targets, target_len, Lmax, Q, Sp, alpha, gext, nll);
    DTLR_CHECK_LAUNCH();
    if (grad_logits) {
        ctc_grad_kernel<<<dim3(Q, B), 256, 2 * (Lmax + 1) * sizeof(float), st>>>(logits, ld, perm, row_sum, targets, target_len, Lmax, Sp,
                                                                                  eps, gext, nll, zero_infinity, grad_logits, Q, C, B);
        DTLR_CHECK_LAUNCH();
    }
    return DTLR_OK;
}
