// dtlr_b200 -- linear sum assignment on the GPU for the Hungarian matcher of the DINO detection loss
// (reference models/dino/matcher.py:94: `scipy.optimize.linear_sum_assignment(c[i])` on a cost matrix copied to the CPU,
// seven times per training step -- SURVEY.md §8f.1).  One CTA per image solves the rectangular problem
// (T targets x Q queries, T <= Q) with the shortest-augmenting-path method scipy uses (Crouse's variant of Jonker-Volgenant):
// the Dijkstra relaxation over the Q columns and the arg-min are block-parallel, duals and the partial assignment live in
// shared memory, arithmetic in fp64 like scipy.  Ties: lowest reduced cost, then an unassigned column, then the lowest index.
#include "common.cuh"

namespace dtlr {

constexpr int LSAP_THREADS = 512;

struct ArgMin {
    double v;
    int j;        // column; 0x7fffffff = none
    int freej;    // 1 if the column is unassigned (preferred among equal costs)
};
__device__ __forceinline__ bool better(const ArgMin& a, const ArgMin& b) {   // is a better than b
    if (a.v != b.v) return a.v < b.v;
    if (a.freej != b.freej) return a.freej > b.freej;
    return a.j < b.j;
}

// ---------------------------------------------------------------------------------------------------------------------
// matching cost of reference matcher.py:57-88 for the (image, target, query) triples of the block diagonal only (the
// reference builds the full (B*Q) x sum(T) matrix and throws the cross-image blocks away), written target-major so that
// the solver's row scan is coalesced:  cost[p][t][q],  p = problem (layer * B + image).
//   class: alpha (1-p)^2 (-log(p+1e-8)) - (1-alpha) p^2 (-log(1-p+1e-8)),  p = sigmoid(logit[q][label_t])
//   bbox : L1(cxcywh)          giou: -GIoU(xyxy)  (union + 1e-6 / area + 1e-6 like util/box_ops.py:24-64)
__global__ void __launch_bounds__(256)
match_cost_kernel(const float* __restrict__ logits, const float* __restrict__ boxes, const int64_t* __restrict__ tgt_labels,
                  const float* __restrict__ tgt_boxes, const int* __restrict__ t_off, const int* __restrict__ t_cnt,
                  int B, int Q, int C, int Tmax, float w_class, float w_bbox, float w_giou, float alpha,
                  float* __restrict__ cost) {
    const int p = blockIdx.y, b = p % B;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int T = t_cnt[b], off = t_off[b];
    if (q >= Q) return;
    const float* lg = logits + ((size_t)p * Q + q) * C;
    const float4 ob = *reinterpret_cast<const float4*>(boxes + ((size_t)p * Q + q) * 4);
    const float ox0 = ob.x - 0.5f * ob.z, oy0 = ob.y - 0.5f * ob.w, ox1 = ob.x + 0.5f * ob.z, oy1 = ob.y + 0.5f * ob.w;
    const float oarea = (ox1 - ox0) * (oy1 - oy0);
    for (int t = 0; t < T; ++t) {
        const int lab = (int)tgt_labels[off + t];
        const float4 tb = *reinterpret_cast<const float4*>(tgt_boxes + (size_t)(off + t) * 4);
        const float pr = 1.f / (1.f + expf(-lg[lab]));
        const float neg = (1.f - alpha) * (pr * pr) * (-logf(1.f - pr + 1e-8f));
        const float pos = alpha * ((1.f - pr) * (1.f - pr)) * (-logf(pr + 1e-8f));
        const float cc = pos - neg;
        const float l1 = fabsf(ob.x - tb.x) + fabsf(ob.y - tb.y) + fabsf(ob.z - tb.z) + fabsf(ob.w - tb.w);
        const float tx0 = tb.x - 0.5f * tb.z, ty0 = tb.y - 0.5f * tb.w, tx1 = tb.x + 0.5f * tb.z, ty1 = tb.y + 0.5f * tb.w;
        const float tarea = (tx1 - tx0) * (ty1 - ty0);
        const float iw = fmaxf(fminf(ox1, tx1) - fmaxf(ox0, tx0), 0.f), ih = fmaxf(fminf(oy1, ty1) - fmaxf(oy0, ty0), 0.f);
        const float inter = iw * ih, uni = oarea + tarea - inter;
        const float iou = inter / (uni + 1e-6f);
        const float cw = fmaxf(fmaxf(ox1, tx1) - fminf(ox0, tx0), 0.f), ch = fmaxf(fmaxf(oy1, ty1) - fminf(oy0, ty0), 0.f);
        const float carea = cw * ch;
        const float giou = iou - (carea - uni) / (carea + 1e-6f);
        cost[((size_t)p * Tmax + t) * Q + q] = w_bbox * l1 + w_class * cc + w_giou * (-giou);
    }
}

// cost[p][t][q] (fp32, target-major) -> q_of_t[p][t];  image of problem p = p % B
__global__ void __launch_bounds__(LSAP_THREADS)
lsap_kernel(const float* __restrict__ cost, int B, int Q, const int* __restrict__ t_cnt, int Tmax, int* __restrict__ q_of_t) {
    extern __shared__ unsigned char ls_raw[];
    const int b = blockIdx.x;
    const int T = t_cnt[b % B];
    double* v = reinterpret_cast<double*>(ls_raw);             // [Q] column duals
    double* sp = v + Q;                                         // [Q] shortest path costs
    double* u = sp + Q;                                         // [Tmax] row duals
    int* path = reinterpret_cast<int*>(u + Tmax);               // [Q]
    int* row4col = path + Q;                                    // [Q]
    int* col4row = row4col + Q;                                 // [Tmax]
    unsigned char* SC = reinterpret_cast<unsigned char*>(col4row + Tmax);   // [Q]
    unsigned char* SR = SC + Q;                                 // [Tmax]
    __shared__ ArgMin red[LSAP_THREADS / 32];
    __shared__ int s_i, s_sink;
    __shared__ double s_minval;

    const float* cb = cost + (size_t)b * Tmax * Q;
    for (int j = threadIdx.x; j < Q; j += blockDim.x) { v[j] = 0.0; row4col[j] = -1; }
    for (int i = threadIdx.x; i < T; i += blockDim.x) { u[i] = 0.0; col4row[i] = -1; }
    __syncthreads();

    for (int cur = 0; cur < T; ++cur) {
        for (int j = threadIdx.x; j < Q; j += blockDim.x) { sp[j] = INFINITY; SC[j] = 0; path[j] = -1; }
        for (int i = threadIdx.x; i < T; i += blockDim.x) SR[i] = 0;
        if (threadIdx.x == 0) { s_i = cur; s_sink = -1; s_minval = 0.0; }
        __syncthreads();
        while (true) {
            const int i = s_i;
            const double minval = s_minval, ui = u[i];
            const float* crow = cb + (size_t)i * Q;
            ArgMin mine{INFINITY, 0x7fffffff, 0};
            for (int j = threadIdx.x; j < Q; j += blockDim.x) {
                if (SC[j]) continue;
                const double r = minval + (double)crow[j] - ui - v[j];
                double s = sp[j];
                if (r < s) { s = r; sp[j] = r; path[j] = i; }
                const ArgMin cand{s, j, row4col[j] == -1 ? 1 : 0};
                if (better(cand, mine)) mine = cand;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ArgMin other;
                other.v = __shfl_xor_sync(0xffffffffu, mine.v, o);
                other.j = __shfl_xor_sync(0xffffffffu, mine.j, o);
                other.freej = __shfl_xor_sync(0xffffffffu, mine.freej, o);
                if (better(other, mine)) mine = other;
            }
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mine;
            __syncthreads();
            if (threadIdx.x == 0) {
                ArgMin bm = red[0];
                for (int w = 1; w < LSAP_THREADS / 32; ++w)
                    if (better(red[w], bm)) bm = red[w];
                if (bm.j == 0x7fffffff || !(bm.v < INFINITY)) {
                    s_sink = -2;                                 // infeasible (all-inf row): leave unassigned
                } else {
                    SR[i] = 1;
                    SC[bm.j] = 1;
                    s_minval = bm.v;
                    if (row4col[bm.j] == -1) s_sink = bm.j; else s_i = row4col[bm.j];
                }
            }
            __syncthreads();
            if (s_sink != -1) break;
        }
        const int sink = s_sink;
        if (sink >= 0) {
            const double minval = s_minval;
            // dual updates (scipy: u[cur] += minVal; u[i in SR, i != cur] += minVal - sp[col4row[i]]; v[j in SC] -= minVal - sp[j])
            for (int i = threadIdx.x; i < T; i += blockDim.x) {
                if (i == cur) u[i] += minval;
                else if (SR[i]) u[i] += minval - sp[col4row[i]];
            }
            for (int j = threadIdx.x; j < Q; j += blockDim.x)
                if (SC[j]) v[j] -= minval - sp[j];
            __syncthreads();
            if (threadIdx.x == 0) {                              // augment along the alternating path
                int j = sink;
                while (true) {
                    const int i = path[j];
                    row4col[j] = i;
                    const int prev = col4row[i];
                    col4row[i] = j;
                    j = prev;
                    if (i == cur) break;
                }
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < Tmax; i += blockDim.x) q_of_t[(size_t)b * Tmax + i] = i < T ? col4row[i] : -1;
}

}  // namespace dtlr

using namespace dtlr;

extern "C" int dtlr_lsap(const float* cost, int P, int B, int Q, const int* t_cnt, int Tmax, int* q_of_t, void* stream) {
    DTLR_CHECK_ARG(P >= 0 && B > 0 && Q > 0 && Tmax >= 0, "lsap: bad sizes");
    if (P == 0 || Tmax == 0) return DTLR_OK;
    DTLR_CHECK_ARG(cost && t_cnt && q_of_t, "lsap: null pointer");
    DTLR_CHECK_ARG(Tmax <= Q, "lsap: more targets (%d) than queries (%d) in an image", Tmax, Q);
    size_t smem = (size_t)Q * 16 + (size_t)Tmax * 8 + (size_t)Q * 8 + (size_t)Tmax * 4 + (size_t)Q + Tmax + 64;
    smem = (smem + 15) & ~(size_t)15;
    DTLR_CHECK_ARG(smem <= (size_t)max_smem_optin(), "lsap: problem too large for shared memory (Q=%d, T=%d)", Q, Tmax);
    if (smem > 48 * 1024) DTLR_CHECK_CUDA(cudaFuncSetAttribute(lsap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lsap_kernel<<<P, LSAP_THREADS, smem, (cudaStream_t)stream>>>(cost, B, Q, t_cnt, Tmax, q_of_t);
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_match_cost(const float* logits, const float* boxes, const int64_t* tgt_labels, const float* tgt_boxes,
                               const int* t_off, const int* t_cnt, int P, int B, int Q, int C, int Tmax, float w_class,
                               float w_bbox, float w_giou, float alpha, float* cost, void* stream) {
    DTLR_CHECK_ARG(P >= 0 && B > 0 && Q > 0 && C > 0 && Tmax >= 0 && P % B == 0, "match_cost: bad sizes");
    if (P == 0 || Tmax == 0) return DTLR_OK;
    DTLR_CHECK_ARG(logits && boxes && tgt_labels && tgt_boxes && t_off && t_cnt && cost, "match_cost: null pointer");
    dim3 grid((Q + 255) / 256, P);
    match_cost_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(logits, boxes, tgt_labels, tgt_boxes, t_off, t_cnt, B, Q, C, Tmax,
                                                             w_class, w_bbox, w_giou, alpha, cost);
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}
