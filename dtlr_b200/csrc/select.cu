// dtlr_b200 -- selection kernels of the hot path (index work: results are compared bit-for-bit / index-for-index with the reference):
//   * two-stage query selection (reference models/dino/deformable_transformer.py:345-353): top-K of the S encoder-token scores per
//     line + the three gathers that follow it (reference anchors, sigmoid of the proposals, the selected memory rows);
//   * PostProcess (reference models/dino/dino.py:1008-1046): top-`num_select` of the Q*C sigmoid scores per line, box conversion /
//     scaling, class-agnostic NMS (torchvision.ops.nms semantics), and the reading of reference evaluation.py:94-115 (keep scores
//     above the threshold, order by box centre x).
// One CTA per line: exact radix select (12 + 12 + 8 bits) over the monotone integer image of the probabilities, candidates taken
// in index order (ties at the threshold resolve to the lowest flat index), bitonic sorts in shared memory.
#include "common.cuh"

namespace dtlr {

__device__ __forceinline__ uint32_t f2key(float f) {           // order-preserving float -> uint
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// all threads of the CTA; n a power of two; ascending by (key, val); caller has synchronised the writes of key/val
__device__ __forceinline__ void bitonic_sort_asc(uint32_t* key, int* val, const int n) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const uint32_t ka = key[i], kb = key[p];
                    const int va = val[i], vb = val[p];
                    const bool a_gt_b = (ka > kb) || (ka == kb && va > vb);
                    const bool asc = (i & k) == 0;
                    if (a_gt_b == asc) { key[i] = kb; key[p] = ka; val[i] = vb; val[p] = va; }
                }
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------- two-stage select
// one CTA per line: all S scores sorted descending (ties: lowest token index first), the first K indices written as int64
__global__ void topk_sort_kernel(const float* __restrict__ scores, int S, int K, int n_pow2, long long* __restrict__ idx_out) {
    extern __shared__ __align__(16) unsigned char dsm[];
    uint32_t* key = reinterpret_cast<uint32_t*>(dsm);
    int* val = reinterpret_cast<int*>(key + n_pow2);
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        key[i] = i < S ? ~f2key(scores[(size_t)b * S + i]) : 0xFFFFFFFFu;
        val[i] = i < S ? i : 0x7fffffff;
    }
    __syncthreads();
    bitonic_sort_asc(key, val, n_pow2);
    for (int i = threadIdx.x; i < K; i += blockDim.x) idx_out[(size_t)b * K + i] = val[i];
}

// one warp per selected (line, rank): ref = sigmoid((delta + prop)[idx]) (the decoder's first reference points = the interm boxes),
// init_box = sigmoid(prop[idx]), tgt = mem[idx] (row of d elements)
template <typename T>
__global__ void __launch_bounds__(256)
select_gather_kernel(const long long* __restrict__ idx, const float* __restrict__ coord, const float* __restrict__ prop,
                     const T* __restrict__ mem, float* __restrict__ refpoint, float* __restrict__ initbox, T* __restrict__ tgt,
                     int S, int K, int d, long long rows) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const long long b = row / K;
    const long long src = b * S + idx[row];
    if (lane < 4) refpoint[row * 4 + lane] = 1.f / (1.f + expf(-(coord[src * 4 + lane] + prop[src * 4 + lane])));
    else if (lane < 8) initbox[row * 4 + lane - 4] = 1.f / (1.f + expf(-prop[src * 4 + lane - 4]));
    const T* s = mem + (size_t)src * d;
    T* o = tgt + (size_t)row * d;
    if (((size_t)d * sizeof(T)) % 16 == 0 && ((((uintptr_t)mem) | ((uintptr_t)tgt)) & 15) == 0) {
        const int n16 = (int)((size_t)d * sizeof(T) / 16);
        const uint4* s4 = reinterpret_cast<const uint4*>(s);
        uint4* o4 = reinterpret_cast<uint4*>(o);
        for (int i = lane; i < n16; i += 32) o4[i] = s4[i];
    } else {
        for (int i = lane; i < d; i += 32) o[i] = s[i];
    }
}

// ---------------------------------------------------------------------------------------------------------------- PostProcess
// warp 0: the bin (from the top) in which the running count reaches `need`; writes (bin, need - count above the bin)
__device__ __forceinline__ void find_bin(const unsigned* hist, const int nb, const unsigned need, unsigned* out_bin, unsigned* out_need) {
    const int lane = threadIdx.x & 31;
    const int chunk = nb >> 5;
    const int hi = nb - 1 - lane * chunk;
    unsigned csum = 0;
    for (int i = 0; i < chunk; ++i) csum += hist[hi - i];
    unsigned incl = csum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const unsigned excl = incl - csum;
    if (excl < need && need <= incl) {
        unsigned acc = excl;
        for (int i = 0; i < chunk; ++i) {
            const unsigned h = hist[hi - i];
            if (acc + h >= need) { *out_bin = (unsigned)(hi - i); *out_need = need - acc; break; }
            acc += h;
        }
    }
}

// box_mode 0: cxcywh -> xyxy; 1: as stored (not_to_xyxy); 2: `test` (x0, y0, w, h).  sizes [B,2] = (img_h, img_w).
__global__ void __launch_bounds__(1024)
pp_topk_kernel(const float* __restrict__ logits, int ld, int Q, int C, const float* __restrict__ boxes, const float* __restrict__ sizes,
               int K, int n_pow2, int box_mode, float* __restrict__ scores, int* __restrict__ labels, float* __restrict__ boxes_out) {
    __shared__ unsigned hist[4096];
    __shared__ unsigned sel_bin, sel_need;
    __shared__ unsigned wt_g[32], wt_e[32];
    extern __shared__ __align__(16) unsigned char dsm[];
    uint32_t* ckey = reinterpret_cast<uint32_t*>(dsm);
    int* cidx = reinterpret_cast<int*>(ckey + n_pow2);
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = Q * C;
    const float* lg = logits + (size_t)b * Q * ld;
    auto keyof = [&](const int e) -> uint32_t {
        const int q = e / C, c = e - q * C;
        const float p = 1.f / (1.f + expf(-lg[(size_t)q * ld + c]));      // torch.sigmoid in fp32; p >= 0: its bits are monotone
        return __float_as_uint(p);
    };
    // ---- exact K-th largest key by three histogram passes
    uint32_t prefix = 0;
    unsigned need = (unsigned)K;
    const int shifts[3] = {20, 8, 0}, widths[3] = {12, 12, 8}, above[3] = {32, 20, 8};
    for (int pass = 0; pass < 3; ++pass) {
        const int nb = 1 << widths[pass];
        for (int i = tid; i < nb; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (int e = tid; e < N; e += blockDim.x) {
            const uint32_t k = keyof(e);
            if (pass == 0 || (k >> above[pass]) == prefix) atomicAdd(&hist[(k >> shifts[pass]) & (nb - 1)], 1u);
        }
        __syncthreads();
        if (warp == 0) find_bin(hist, nb, need, &sel_bin, &sel_need);
        __syncthreads();
        prefix = (prefix << widths[pass]) | sel_bin;
        need = sel_need;
        __syncthreads();
    }
    const uint32_t T = prefix;                  // the K-th largest key; `need` of the entries equal to it are taken, lowest index first
    const int n_gt = K - (int)need;
    for (int i = tid; i < n_pow2; i += blockDim.x) { ckey[i] = 0xFFFFFFFFu; cidx[i] = 0x7fffffff; }
    __syncthreads();
    // ---- candidates in index order
    int cnt_g = 0, cnt_e = 0;
    for (int base = 0; base < N; base += blockDim.x) {
        const int e = base + tid;
        const uint32_t k = e < N ? keyof(e) : 0u;
        const bool gt = e < N && k > T, eq = e < N && k == T;
        if (!__syncthreads_or(gt || eq)) continue;
        const unsigned bg = __ballot_sync(0xffffffffu, gt), be = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) { wt_g[warp] = __popc(bg); wt_e[warp] = __popc(be); }
        __syncthreads();
        int pg = __popc(bg & ((1u << lane) - 1u)), pe = __popc(be & ((1u << lane) - 1u));
        int tg = 0, te = 0;
        const int nw = blockDim.x >> 5;
        for (int w = 0; w < nw; ++w) {
            const int g = wt_g[w], q = wt_e[w];
            if (w < warp) { pg += g; pe += q; }
            tg += g; te += q;
        }
        if (gt) { ckey[cnt_g + pg] = ~k; cidx[cnt_g + pg] = e; }
        if (eq && cnt_e + pe < (int)need) { ckey[n_gt + cnt_e + pe] = ~k; cidx[n_gt + cnt_e + pe] = e; }
        cnt_g += tg; cnt_e += te;
        __syncthreads();
    }
    bitonic_sort_asc(ckey, cidx, n_pow2);       // ~key ascending = score descending, ties by flat index
    const float ih = sizes[b * 2], iw = sizes[b * 2 + 1];
    for (int i = tid; i < K; i += blockDim.x) {
        const int e = cidx[i];
        const int q = e / C, c = e - q * C;
        scores[(size_t)b * K + i] = __uint_as_float(~ckey[i]);
        labels[(size_t)b * K + i] = c;
        const float4 bx = *reinterpret_cast<const float4*>(boxes + ((size_t)b * Q + q) * 4);
        float x0 = bx.x, y0 = bx.y, x1 = bx.z, y1 = bx.w;
        if (box_mode != 1) {
            x0 = bx.x - 0.5f * bx.z; y0 = bx.y - 0.5f * bx.w; x1 = bx.x + 0.5f * bx.z; y1 = bx.y + 0.5f * bx.w;
            if (box_mode == 2) { x1 = x1 - x0; y1 = y1 - y0; }
        }
        *reinterpret_cast<float4*>(boxes_out + ((size_t)b * K + i) * 4) = make_float4(x0 * iw, y0 * ih, x1 * iw, y1 * ih);
    }
}

// class-agnostic NMS over the K score-sorted boxes of a line (torchvision.ops.nms: suppress j > i when IoU(i, j) > thr), then the
// reading of evaluation.py:97-115: kept detections with score > score_thr, ordered by (x0 + x1) / 2.
__global__ void __launch_bounds__(1024)
pp_nms_read_kernel(const float* __restrict__ boxes_k, const float* __restrict__ scores, const int* __restrict__ labels, int K, int n_pow2,
                   float iou_thr, float score_thr, unsigned char* __restrict__ keep, int* __restrict__ read_labels,
                   int* __restrict__ read_count) {
    extern __shared__ __align__(16) unsigned char dsm[];
    const int W = (K + 31) >> 5;
    float4* box = reinterpret_cast<float4*>(dsm);
    uint32_t* skey = reinterpret_cast<uint32_t*>(box + K);
    int* sval = reinterpret_cast<int*>(skey + n_pow2);
    unsigned char* kf = reinterpret_cast<unsigned char*>(sval + n_pow2);
    uint32_t* mask = reinterpret_cast<uint32_t*>(kf + ((K + 15) & ~15));
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < K; i += blockDim.x) {
        box[i] = *reinterpret_cast<const float4*>(boxes_k + ((size_t)b * K + i) * 4);
        kf[i] = 1;
    }
    __syncthreads();
    if (iou_thr > 0.f) {
        for (int t = tid; t < K * W; t += blockDim.x) {
            const int i = t / W, w = t - i * W;
            if (w * 32 + 31 <= i) { mask[t] = 0; continue; }        // only j > i can be suppressed by i
            const float4 a = box[i];
            const float sa = (a.z - a.x) * (a.w - a.y);
            uint32_t bits = 0;
            for (int jj = 0; jj < 32; ++jj) {
                const int j = w * 32 + jj;
                if (j > i && j < K) {
                    const float4 c = box[j];
                    const float iw = fmaxf(fminf(a.z, c.z) - fmaxf(a.x, c.x), 0.f), ih = fmaxf(fminf(a.w, c.w) - fmaxf(a.y, c.y), 0.f);
                    const float inter = iw * ih, sb = (c.z - c.x) * (c.w - c.y);
                    if (inter / (sa + sb - inter) > iou_thr) bits |= 1u << jj;
                }
            }
            mask[t] = bits;
        }
        __syncthreads();
        if (tid < 32) {                           // the greedy scan is sequential in i: lane w carries word w of the removed set
            uint32_t removed = 0;
            for (int i = 0; i < K; ++i) {
                const uint32_t word = __shfl_sync(0xffffffffu, removed, i >> 5);
                const bool dead = (word >> (i & 31)) & 1u;
                if (!dead && lane < W) removed |= mask[i * W + lane];
                if (lane == 0) kf[i] = dead ? 0 : 1;
            }
        }
        __syncthreads();
    }
    int valid_i = 0;
    for (int i = tid; i < n_pow2; i += blockDim.x) {
        const bool v = i < K && kf[i] && scores[(size_t)b * K + i] > score_thr;
        skey[i] = v ? f2key((box[i].x + box[i].z) / 2.f) : 0xFFFFFFFFu;
        sval[i] = v ? i : 0x7fffffff;
        valid_i += v ? 1 : 0;
        if (i < K) keep[(size_t)b * K + i] = kf[i];
    }
    // block-wide sum of valid_i
    __shared__ int vsum[32];
    int v = valid_i;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) vsum[tid >> 5] = v;
    __syncthreads();
    int total = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) total += vsum[w];
    bitonic_sort_asc(skey, sval, n_pow2);
    for (int i = tid; i < K; i += blockDim.x) read_labels[(size_t)b * K + i] = i < total ? labels[(size_t)b * K + sval[i]] : -1;
    if (tid == 0) read_count[b] = total;
}

}  // namespace dtlr

using namespace dtlr;

static int pow2_at_least(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

extern "C" int dtlr_topk_select(const float* scores, int B, int S, int K, long long* idx, void* stream) {
    DTLR_CHECK_ARG(B >= 0 && S > 0 && K > 0, "topk_select: bad sizes");
    DTLR_CHECK_ARG(K <= S, "topk_select: selected index k out of range (k = %d > %d tokens)", K, S);     // torch.topk's failure mode (quirk Q2)
    if (B == 0) return DTLR_OK;
    DTLR_CHECK_ARG(scores && idx, "topk_select: null pointer");
    const int n = pow2_at_least(S);
    const size_t smem = (size_t)n * 8;
    DTLR_CHECK_ARG(smem <= (size_t)max_smem_optin(), "topk_select: %d tokens per line exceed the shared-memory sort", S);
    if (smem > 48 * 1024) DTLR_CHECK_CUDA(cudaFuncSetAttribute(topk_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    topk_sort_kernel<<<B, n < 1024 ? n : 1024, smem, (cudaStream_t)stream>>>(scores, S, K, n, idx);
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_select_gather(const long long* idx, const float* coord, const float* prop, const void* mem, float* refpoint,
                                  float* initbox, void* tgt, int B, int S, int K, int d, int dtype, void* stream) {
    DTLR_CHECK_ARG(B >= 0 && S > 0 && K > 0 && d > 0, "select_gather: bad sizes");
    if (B == 0) return DTLR_OK;
    DTLR_CHECK_ARG(idx && coord && prop && mem && refpoint && initbox && tgt, "select_gather: null pointer");
    const long long rows = (long long)B * K;
    const unsigned grid = (unsigned)((rows + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DTLR_F32)
        select_gather_kernel<float><<<grid, 256, 0, st>>>(idx, coord, prop, (const float*)mem, refpoint, initbox, (float*)tgt, S, K, d, rows);
    else if (dtype == DTLR_OP16)
        select_gather_kernel<op16_t><<<grid, 256, 0, st>>>(idx, coord, prop, (const op16_t*)mem, refpoint, initbox,
                                                                  (op16_t*)tgt, S, K, d, rows);
    else
        DTLR_CHECK_ARG(false, "select_gather: unsupported dtype %d", dtype);
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_postprocess(const float* logits, int ld, const float* boxes, const float* sizes, int B, int Q, int C, int K,
                                int box_mode, float nms_iou, float score_thr, float* scores, int* labels, float* boxes_out,
                                unsigned char* keep, int* read_labels, int* read_count, void* stream) {
    DTLR_CHECK_ARG(B >= 0 && Q > 0 && C > 0 && ld >= C && K > 0, "postprocess: bad sizes");
    DTLR_CHECK_ARG((long long)Q * C < (1ll << 31), "postprocess: Q*C too large");
    DTLR_CHECK_ARG((long long)K <= (long long)Q * C, "postprocess: selected index k out of range (num_select %d > %lld scores)", K, (long long)Q * C);
    DTLR_CHECK_ARG(box_mode >= 0 && box_mode <= 2, "postprocess: bad box mode");
    if (B == 0) return DTLR_OK;
    DTLR_CHECK_ARG(logits && boxes && sizes && scores && labels && boxes_out, "postprocess: null pointer");
    DTLR_CHECK_ARG((((uintptr_t)boxes | (uintptr_t)boxes_out) & 15) == 0, "postprocess: boxes must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int n = pow2_at_least(K);
    const size_t smem1 = (size_t)n * 8;
    DTLR_CHECK_ARG(smem1 + 17 * 1024 <= (size_t)max_smem_optin(), "postprocess: num_select %d too large", K);
    if (smem1 > 30 * 1024) DTLR_CHECK_CUDA(cudaFuncSetAttribute(pp_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    pp_topk_kernel<<<B, 1024, smem1, st>>>(logits, ld, Q, C, boxes, sizes, K, n, box_mode, scores, labels, boxes_out);
    DTLR_CHECK_LAUNCH();
    if (keep) {
        DTLR_CHECK_ARG(read_labels && read_count, "postprocess: keep needs read_labels and read_count");
        DTLR_CHECK_ARG(K <= 1024, "postprocess: NMS covers num_select <= 1024 (got %d)", K);
        const int W = (K + 31) / 32;
        const size_t smem2 = (size_t)K * 16 + (size_t)n * 8 + ((K + 15) & ~15) + (size_t)K * W * 4;
        DTLR_CHECK_ARG(smem2 <= (size_t)max_smem_optin(), "postprocess: NMS shared memory");
        if (smem2 > 48 * 1024) DTLR_CHECK_CUDA(cudaFuncSetAttribute(pp_nms_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        pp_nms_read_kernel<<<B, 1024, smem2, st>>>(boxes_out, scores, labels, K, n, nms_iou, score_thr, keep, read_labels, read_count);
        DTLR_CHECK_LAUNCH();
    }
    return DTLR_OK;
}
