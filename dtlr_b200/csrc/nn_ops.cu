// dtlr_b200 -- the non-GEMM kernels of the DINO forward (all HBM-bound elementwise / small-reduction work):
// im2col (NHWC), max-pool, GroupNorm, sine position embedding, residual LayerNorm, deformable-attention prologue
// (softmax + sampling locations), proposal generation, sine query embedding, iterative box refinement, row max.
// Each kernel states the reference lines it restates.  T is the activation type (float = parity, bf16 = throughput);
// statistics and transcendental math are always fp32.
#include "common.cuh"


// every kernel of this file starts with pdl_launch_dependents(); pdl_wait(); and is launched with the programmatic-dependent-launch
// attribute: its launch latency overlaps the tail of its predecessor in the stream (common.cuh)
#define DTLR_LAUNCH(kernel, grid, block, smem, st, ...)                                                                   \
    do {                                                                                                                  \
        cudaError_t _le = dtlr::launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__);             \
        if (_le != cudaSuccess) {                                                                                         \
            dtlr::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_le), __FILE__, __LINE__);            \
            return DTLR_ERR_CUDA;                                                                                         \
        }                                                                                                                 \
    } while (0)

namespace dtlr {

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<op16_t>(const op16_t* p) { return op16_to_f32(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<op16_t>(op16_t* p, float v) { *p = f32_to_op16(v); }

// 8 consecutive channels <-> 8 floats
template <typename T> __device__ __forceinline__ void ld8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void ld8<float>(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void ld8<op16_t>(const op16_t* p, float (&v)[8]) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = op16_lo_f32(w[i]); v[2 * i + 1] = op16_hi_f32(w[i]); }
}
template <typename T> __device__ __forceinline__ void st8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void st8<float>(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void st8<op16_t>(op16_t* p, const float (&v)[8]) {
    uint4 o;
    op16x2_t a = op16_pack2(v[0], v[1]), b = op16_pack2(v[2], v[3]);
    op16x2_t c = op16_pack2(v[4], v[5]), d = op16_pack2(v[6], v[7]);
    o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
    o.z = *reinterpret_cast<uint32_t*>(&c); o.w = *reinterpret_cast<uint32_t*>(&d);
    *reinterpret_cast<uint4*>(p) = o;
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---------------------------------------------------------------------------------------------- im2col
// NHWC (or NCHW fp32 for the network input) -> patch matrix [B*Ho*Wo, ldo], K ordered (kh, kw, cin), zero padding.
// Makes every k>1 / strided convolution of the ResNet-50 trunk (torchvision resnet50 as wrapped by reference
// models/dino/backbone.py:109-128) and of input_proj[3] (dino.py:126-135) a dtlr_gemm call.
template <typename TI, typename TO, bool NCHW_IN>
__global__ void im2col_kernel(const TI* __restrict__ x, TO* __restrict__ out, int B, int H, int W, int C, int KH, int KW,
                              int stride, int pad, int Ho, int Wo, int ldo) {
    pdl_launch_dependents();
    pdl_wait();
    const int K = KH * KW * C;
    const long long total = (long long)B * Ho * Wo * ldo;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % ldo);
        const long long r = i / ldo;
        float v = 0.f;
        if (k < K) {
            const int c = k % C, kw = (k / C) % KW, kh = k / (C * KW);
            const int wo = (int)(r % Wo), ho = (int)((r / Wo) % Ho), b = (int)(r / ((long long)Wo * Ho));
            const int hi = ho * stride - pad + kh, wi = wo * stride - pad + kw;
            if (hi >= 0 && hi < H && wi >= 0 && wi < W) {
                const size_t idx = NCHW_IN ? (((size_t)b * C + c) * H + hi) * W + wi : (((size_t)b * H + hi) * W + wi) * C + c;
                v = ldf<TI>(x + idx);
            }
        }
        stf<TO>(out + i, v);
    }
}

// 8-channel vectorised variant for NHWC inputs with C % 8 == 0 (every conv except the stem)
template <typename T>
__global__ void im2col_vec8_kernel(const T* __restrict__ x, T* __restrict__ out, int B, int H, int W, int C, int KH, int KW,
                                   int stride, int pad, int Ho, int Wo) {
    pdl_launch_dependents();
    pdl_wait();
    const int C8 = C / 8;
    const int K8 = KH * KW * C8;
    const long long total = (long long)B * Ho * Wo * K8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % K8);
        const long long r = i / K8;
        const int c8 = k % C8, kw = (k / C8) % KW, kh = k / (C8 * KW);
        const int wo = (int)(r % Wo), ho = (int)((r / Wo) % Ho), b = (int)(r / ((long long)Wo * Ho));
        const int hi = ho * stride - pad + kh, wi = wo * stride - pad + kw;
        float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (hi >= 0 && hi < H && wi >= 0 && wi < W) ld8<T>(x + ((((size_t)b * H + hi) * W + wi) * C + c8 * 8), v);
        st8<T>(out + (size_t)i * 8, v);
    }
}

// ---------------------------------------------------------------------------------------------- max-pool 3x3 s2 p1 (NHWC)
template <typename T>
__global__ void maxpool3x3s2_kernel(const T* __restrict__ x, T* __restrict__ out, int B, int H, int W, int C, int Ho, int Wo) {
    pdl_launch_dependents();
    pdl_wait();
    const int C8 = C / 8;
    const long long total = (long long)B * Ho * Wo * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8);
        const long long r = i / C8;
        const int wo = (int)(r % Wo), ho = (int)((r / Wo) % Ho), b = (int)(r / ((long long)Wo * Ho));
        float m[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
        for (int dh = 0; dh < 3; ++dh)
            for (int dw = 0; dw < 3; ++dw) {
                const int hi = ho * 2 - 1 + dh, wi = wo * 2 - 1 + dw;
                if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
                float v[8];
                ld8<T>(x + ((((size_t)b * H + hi) * W + wi) * C + c8 * 8), v);
#pragma unroll
                for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], v[k]);
            }
        st8<T>(out + (size_t)i * 8, m);
    }
}

// ---------------------------------------------------------------------------------------------- GroupNorm(32, C) on NHWC
// reference models/dino/dino.py:121-124 (nn.GroupNorm(32, hidden_dim), eps 1e-5) applied to one feature level; the
// result is written straight into the level's slice of the flattened token buffer (deformable_transformer.py:278-288).
// grid (groups, B); x [B, HW, C] (fp32 from the projection GEMM), out rows b*out_stride_b + hw.
template <typename TO>
__global__ void groupnorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 TO* __restrict__ out, int HW, int C, int G, long long out_stride_b, float eps) {
    pdl_launch_dependents();
    pdl_wait();
    const int g = blockIdx.x, b = blockIdx.y;
    const int cpg = C / G;
    const float* xb = x + (size_t)b * HW * C + g * cpg;
    const int n = HW * cpg;
    float s = 0.f, ss = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = xb[(size_t)(i / cpg) * C + (i % cpg)];
        s += v;
    }
    __shared__ float red[32];
    __shared__ float stat[2];
    s = warp_sum_f(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        t = warp_sum_f(t);
        if (threadIdx.x == 0) stat[0] = t / n;
    }
    __syncthreads();
    const float mean = stat[0];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float d = xb[(size_t)(i / cpg) * C + (i % cpg)] - mean;
        ss += d * d;
    }
    ss = warp_sum_f(ss);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        t = warp_sum_f(t);
        if (threadIdx.x == 0) stat[1] = rsqrtf(t / n + eps);
    }
    __syncthreads();
    const float rstd = stat[1];
    TO* ob = out + (size_t)b * out_stride_b * C + g * cpg;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int c = i % cpg;
        const size_t off = (size_t)(i / cpg) * C + c;
        stf<TO>(ob + off, (xb[off] - mean) * rstd * gamma[g * cpg + c] + beta[g * cpg + c]);
    }
}

// Coalesced variant for 8 channels per group (GroupNorm(32, 256)): one CTA per (image, 32-channel quad = 4 groups), every warp reads
// full 128-byte row segments with 16-byte loads (lane & 7 = float4 within the segment, so a thread always sees ONE group), rows
// strided over the 32 row slots of the CTA.  With HW <= 32 * RPT the CTA's slice stays in registers: one read of x instead of three.
// Two-pass statistics (mean, then centred sum of squares) like the kernel above -- same arithmetic, same eps placement.
template <typename TO, int RPT>
__global__ void __launch_bounds__(256)
groupnorm8_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                  TO* __restrict__ out, int HW, int C, long long out_stride_b, float eps) {
    pdl_launch_dependents();
    pdl_wait();
    const int cq = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f4 = lane & 7;                              // float4 index inside the 32-channel segment
    const int grp = f4 >> 1;                              // group (of this CTA's 4) the thread works for
    const int slot = warp * 4 + (lane >> 3);              // row slot 0..31
    const float* xb = x + (size_t)b * HW * C + cq * 32 + f4 * 4;
    __shared__ float red[8][4];
    __shared__ float stat[2][4];
    const bool cached = HW <= 32 * RPT;
    float4 v[RPT];
    float s = 0.f;
    if (cached) {
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const int r = slot + 32 * i;
            v[i] = r < HW ? *reinterpret_cast<const float4*>(xb + (size_t)r * C) : make_float4(0.f, 0.f, 0.f, 0.f);
            s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    } else {
        for (int r = slot; r < HW; r += 32) {
            const float4 t = *reinterpret_cast<const float4*>(xb + (size_t)r * C);
            s += (t.x + t.y) + (t.z + t.w);
        }
    }
    const float inv_n = 1.f / (float)(HW * 8);
    auto group_reduce = [&](float val, float* dst) {       // sum over the threads of one group: lanes ^1, ^8, ^16, then the 8 warps
        val += __shfl_xor_sync(0xffffffffu, val, 1);
        val += __shfl_xor_sync(0xffffffffu, val, 8);
        val += __shfl_xor_sync(0xffffffffu, val, 16);
        if ((lane & 25) == 0) red[warp][grp] = val;         // lanes 0, 2, 4, 6: one per group
        __syncthreads();
        if (tid < 4) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += red[w][tid];
            dst[tid] = t;
        }
        __syncthreads();
    };
    group_reduce(s, stat[0]);
    const float mean = stat[0][grp] * inv_n;
    float ss = 0.f;
    if (cached) {
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            if (slot + 32 * i < HW) {
                const float a = v[i].x - mean, bq = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
                ss += (a * a + bq * bq) + (c * c + d * d);
            }
        }
    } else {
        for (int r = slot; r < HW; r += 32) {
            const float4 t = *reinterpret_cast<const float4*>(xb + (size_t)r * C);
            const float a = t.x - mean, bq = t.y - mean, c = t.z - mean, d = t.w - mean;
            ss += (a * a + bq * bq) + (c * c + d * d);
        }
    }
    group_reduce(ss, stat[1]);
    const float rstd = rsqrtf(stat[1][grp] * inv_n + eps);
    const int c0 = cq * 32 + f4 * 4;
    const float4 gm = *reinterpret_cast<const float4*>(gamma + c0), bt = *reinterpret_cast<const float4*>(beta + c0);
    TO* ob = out + (size_t)b * out_stride_b * C + c0;
    auto emit = [&](const int r, const float4 t) {
        TO* o = ob + (size_t)r * C;
        stf<TO>(o, (t.x - mean) * rstd * gm.x + bt.x);
        stf<TO>(o + 1, (t.y - mean) * rstd * gm.y + bt.y);
        stf<TO>(o + 2, (t.z - mean) * rstd * gm.z + bt.z);
        stf<TO>(o + 3, (t.w - mean) * rstd * gm.w + bt.w);
    };
    if (cached) {
#pragma unroll
        for (int i = 0; i < RPT; ++i)
            if (slot + 32 * i < HW) emit(slot + 32 * i, v[i]);
    } else {
        for (int r = slot; r < HW; r += 32) emit(r, *reinterpret_cast<const float4*>(xb + (size_t)r * C));
    }
}

// ---------------------------------------------------------------------------------------------- sine position embedding
// reference models/dino/position_encoding.py:79-108 (+ level_embed, deformable_transformer.py:281-282).
// mask [B,H,W] uint8 (1 = padding); out rows b*out_stride_b + (y*W+x), C = 2*npf channels: [pos_y | pos_x].
template <typename TO>
__global__ void pos_sine_kernel(const unsigned char* __restrict__ mask, const float* __restrict__ level_embed,
                                TO* __restrict__ out, int B, int H, int W, int npf, float temp_h, float temp_w,
                                long long out_stride_b) {
    pdl_launch_dependents();
    pdl_wait();
    // one 128-thread block per token.  Warp 0 counts the unmasked pixels of the token's column (cumsum over H), warp 1 those
    // of its row (cumsum over W); then every thread produces one (sin, cos) pair per axis: dim_t[2k] == dim_t[2k+1].
    const int C = 2 * npf;
    const long long tok = blockIdx.x;
    const int x = (int)(tok % W), y = (int)((tok / W) % H), b = (int)(tok / ((long long)W * H));
    const unsigned char* mb = mask + (size_t)b * H * W;
    __shared__ float emb[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < 2) {
        float cum = 0.f, tot = 0.f;
        if (warp == 0) {
            for (int i = lane; i < H; i += 32) { const float nm = mb[i * W + x] ? 0.f : 1.f; tot += nm; if (i <= y) cum += nm; }
        } else {
            for (int j = lane; j < W; j += 32) { const float nm = mb[y * W + j] ? 0.f : 1.f; tot += nm; if (j <= x) cum += nm; }
        }
        cum = warp_sum_f(cum);
        tot = warp_sum_f(tot);
        if (lane == 0) emb[warp] = cum / (tot + 1e-6f) * 6.283185307179586f;
    }
    __syncthreads();
    TO* o = out + ((size_t)b * out_stride_b + (size_t)y * W + x) * C;
    const int half_pairs = npf / 2;                              // (sin, cos) pairs per axis
    for (int k = threadIdx.x; k < 2 * half_pairs; k += blockDim.x) {
        const bool is_x = k >= half_pairs;
        const int i = is_x ? k - half_pairs : k;
        const float dim_t = powf(is_x ? temp_w : temp_h, 2.f * (float)i / (float)npf);
        const float a = (is_x ? emb[1] : emb[0]) / dim_t;
        float sn, cs;
        sincosf(a, &sn, &cs);
        const int c = (is_x ? npf : 0) + 2 * i;
        stf<TO>(o + c, sn + (level_embed ? level_embed[c] : 0.f));
        stf<TO>(o + c + 1, cs + (level_embed ? level_embed[c + 1] : 0.f));
    }
}

// ---------------------------------------------------------------------------------------------- (residual +) LayerNorm, C = 256
// reference nn.LayerNorm(256) eps 1e-5 after every attention / FFN block (deformable_transformer.py:813-814,806-807,
// 906-907,956-957,878-879), enc_output_norm (:326) and decoder.norm (:758).  One warp per row, 8 channels per lane.
// y = LN(x (+ res)); optional second output y2 = y + add2 (the "+pos" query of the next block).
// R consecutive rows per warp: all of their loads (x, res, add2) are issued before the first reduction, so a warp keeps R (x 2-3)
// 512-byte rows in flight -- with one row per warp the kernel ran at ~55 % of the copy bandwidth (latency bound).
template <typename T, int R>
__global__ void __launch_bounds__(256)
add_layernorm256_kernel(const T* __restrict__ x, const T* __restrict__ res, const float* __restrict__ gamma,
                        const float* __restrict__ beta, T* __restrict__ y, const T* __restrict__ add2,
                        T* __restrict__ y2, int rows, float eps) {
    pdl_launch_dependents();
    pdl_wait();
    const int row0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R;
    if (row0 >= rows) return;
    const int lane = threadIdx.x & 31;
    float v[R][8], a[R][8];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const size_t off = (size_t)min(row0 + i, rows - 1) * 256 + lane * 8;
        ld8<T>(x + off, v[i]);
        if (res) {
            float r[8];
            ld8<T>(res + off, r);
#pragma unroll
            for (int k = 0; k < 8; ++k) v[i][k] += r[k];
        }
        if (y2) ld8<T>(add2 + off, a[i]);
    }
    float g[8], bt[8];
    ld8<float>(gamma + lane * 8, g);
    ld8<float>(beta + lane * 8, bt);
#pragma unroll
    for (int i = 0; i < R; ++i) {
        if (row0 + i >= rows) break;
        const size_t off = (size_t)(row0 + i) * 256 + lane * 8;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += v[i][k];
        const float mean = warp_sum_f(s) * (1.f / 256.f);
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float d = v[i][k] - mean; ss += d * d; }
        const float rstd = rsqrtf(warp_sum_f(ss) * (1.f / 256.f) + eps);
        float o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = (v[i][k] - mean) * rstd * g[k] + bt[k];
        st8<T>(y + off, o);
        if (y2) {
#pragma unroll
            for (int k = 0; k < 8; ++k) a[i][k] += o[k];
            st8<T>(y2 + off, a[i]);
        }
    }
}

// out = a + b (8-wide), optionally zeroing rows where rowmask != 0 (value.masked_fill, ms_deform_attn.py:95-96)
template <typename T>
__global__ void add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ out, long long n8) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        float x[8], y[8];
        ld8<T>(a + i * 8, x);
        ld8<T>(b + i * 8, y);
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] += y[k];
        st8<T>(out + i * 8, x);
    }
}
template <typename T>
__global__ void zero_masked_rows_kernel(T* __restrict__ x, const unsigned char* __restrict__ rowmask, long long rows, int C8) {
    pdl_launch_dependents();
    pdl_wait();
    const long long total = rows * C8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        if (rowmask[i / C8]) {
            const float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            st8<T>(x + i * 8, z);
        }
    }
}

// ---------------------------------------------------------------------------------------------- MSDA prologue
// reference ops/modules/ms_deform_attn.py:98-108: softmax over the L*P attention logits of each head and
//   2-coord refs (encoder): loc = ref_l + off / (W_l, H_l)
//   4-coord refs (decoder): loc = ref_l[:2] + off / P * ref_l[2:] * 0.5
// with ref_l = ref * valid_ratio_l (deformable_transformer.py:491, 686-687).
// proj [rows, ld] fp32: columns [0, M*L*P*2) offsets (m,l,p,xy), then M*L*P logits.  One thread per (row, head).
struct PrepLevels { int n; int H[8], W[8]; };
__global__ void msda_prep_kernel(const float* __restrict__ proj, int ld, const float* __restrict__ ref, int RD,
                                 const float* __restrict__ valid_ratios, float* __restrict__ loc, float* __restrict__ attn,
                                 const __grid_constant__ PrepLevels lv, int B, int Lq, int M, int P) {
    pdl_launch_dependents();
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * Lq * M;
    if (i >= total) return;
    const int m = (int)(i % M);
    const long long row = i / M;
    const int b = (int)(row / Lq);
    const int L = lv.n, LP = L * P;
    const float* pr = proj + (size_t)row * ld;
    const float* off = pr + (size_t)m * LP * 2;
    const float* lg = pr + (size_t)M * LP * 2 + (size_t)m * LP;
    float mx = -INFINITY;
    for (int k = 0; k < LP; ++k) mx = fmaxf(mx, lg[k]);
    float den = 0.f;
    for (int k = 0; k < LP; ++k) den += expf(lg[k] - mx);
    const float inv = 1.f / den;
    const float* rf = ref + (size_t)row * RD;
    float* lo = loc + (size_t)i * LP * 2;
    float* at = attn + (size_t)i * LP;
    for (int l = 0; l < L; ++l) {
        const float vx = valid_ratios[((size_t)b * L + l) * 2], vy = valid_ratios[((size_t)b * L + l) * 2 + 1];
        const float rx = rf[0] * vx, ry = rf[1] * vy;
        for (int p = 0; p < P; ++p) {
            const int k = l * P + p;
            const float ox = off[2 * k], oy = off[2 * k + 1];
            float x, y;
            if (RD == 2) {
                x = rx + ox / (float)lv.W[l];
                y = ry + oy / (float)lv.H[l];
            } else {
                x = rx + ox / (float)P * (rf[2] * vx) * 0.5f;
                y = ry + oy / (float)P * (rf[3] * vy) * 0.5f;
            }
            lo[2 * k] = x;
            lo[2 * k + 1] = y;
            at[k] = expf(lg[k] - mx) * inv;
        }
    }
}

// 4 levels x 4 points (every shipped config), fp32 projection rows with 16-byte aligned pitch: the 16 points of a (query, head) in
// registers, 16-byte loads / stores; arithmetic identical to the generic kernel above (used by the fine-tune step, which needs
// sampling_locations / attention_weights in HBM for the backward; the inference path fuses all of this into the gather kernel)
__global__ void __launch_bounds__(128)
msda_prep16_kernel(const float* __restrict__ proj, int ld, const float* __restrict__ ref, int RD, const float* __restrict__ valid_ratios,
                   float* __restrict__ loc, float* __restrict__ attn, const __grid_constant__ PrepLevels lv, int B, int Lq, int M) {
    pdl_launch_dependents();
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * Lq * M) return;
    const int m = (int)(i % M);
    const long long row = i / M;
    const int b = (int)(row / Lq);
    const float* pr = proj + (size_t)row * ld;
    float off[32], lg[16];
#pragma unroll
    for (int k = 0; k < 32; k += 4) *reinterpret_cast<float4*>(off + k) = *reinterpret_cast<const float4*>(pr + (size_t)m * 32 + k);
#pragma unroll
    for (int k = 0; k < 16; k += 4) *reinterpret_cast<float4*>(lg + k) = *reinterpret_cast<const float4*>(pr + (size_t)M * 32 + (size_t)m * 16 + k);
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 16; ++k) mx = fmaxf(mx, lg[k]);
    float den = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) den += expf(lg[k] - mx);
    const float inv = 1.f / den;
    const float* rf = ref + (size_t)row * RD;
    const float r0 = rf[0], r1 = rf[1];
    float rw = 0.f, rh = 0.f;
    if (RD == 4) { rw = rf[2]; rh = rf[3]; }
    float lo[32], at[16];
#pragma unroll
    for (int l = 0; l < 4; ++l) {
        const float vx = valid_ratios[((size_t)b * 4 + l) * 2], vy = valid_ratios[((size_t)b * 4 + l) * 2 + 1];
        const float rx = r0 * vx, ry = r1 * vy;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int k = l * 4 + p;
            if (RD == 2) {
                lo[2 * k] = rx + off[2 * k] / (float)lv.W[l];
                lo[2 * k + 1] = ry + off[2 * k + 1] / (float)lv.H[l];
            } else {
                lo[2 * k] = rx + off[2 * k] / 4.f * (rw * vx) * 0.5f;
                lo[2 * k + 1] = ry + off[2 * k + 1] / 4.f * (rh * vy) * 0.5f;
            }
            at[k] = expf(lg[k] - mx) * inv;
        }
    }
#pragma unroll
    for (int k = 0; k < 32; k += 4) *reinterpret_cast<float4*>(loc + (size_t)i * 32 + k) = *reinterpret_cast<const float4*>(lo + k);
#pragma unroll
    for (int k = 0; k < 16; k += 4) *reinterpret_cast<float4*>(attn + (size_t)i * 16 + k) = *reinterpret_cast<const float4*>(at + k);
}

// encoder reference points before the per-level valid-ratio product (deformable_transformer.py:479-490):
// ref[b, tok] = ((x+0.5)/(vr_w*W_l), (y+0.5)/(vr_h*H_l)) for the token's own level l.
__global__ void enc_ref_kernel(const float* __restrict__ valid_ratios, float* __restrict__ ref, const __grid_constant__ PrepLevels lv,
                               int B, int S) {
    pdl_launch_dependents();
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * S) return;
    const int b = (int)(i / S);
    int t = (int)(i % S), l = 0;
    while (l < lv.n - 1 && t >= lv.H[l] * lv.W[l]) { t -= lv.H[l] * lv.W[l]; ++l; }
    const int y = t / lv.W[l], x = t % lv.W[l];
    const float vx = valid_ratios[((size_t)b * lv.n + l) * 2], vy = valid_ratios[((size_t)b * lv.n + l) * 2 + 1];
    ref[i * 2] = ((float)x + 0.5f) / (vx * (float)lv.W[l]);
    ref[i * 2 + 1] = ((float)y + 0.5f) / (vy * (float)lv.H[l]);
}

// ---------------------------------------------------------------------------------------------- two-stage proposals
// reference models/dino/utils.py:15-64: anchor (cx,cy,w,h) per token, validity, logit; zeroes invalid/padded memory rows.
// valid_hw [B, L, 2] = (valid_H, valid_W) counts from the level masks.
// One warp per token (8 tokens per CTA), 16-byte loads / stores when the row allows: HBM bound (one read + one write of the memory).
template <typename T>
__global__ void __launch_bounds__(256)
proposals_kernel(const T* __restrict__ memory, const unsigned char* __restrict__ pad, const int* __restrict__ valid_hw,
                 T* __restrict__ out_memory, float* __restrict__ proposals, const __grid_constant__ PrepLevels lv,
                 int B, int S, int C, float default_hw) {
    pdl_launch_dependents();
    pdl_wait();
    const long long tok = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tok >= (long long)B * S) return;
    const int lane = threadIdx.x & 31;
    const int b = (int)(tok / S);
    int t = (int)(tok % S), l = 0;
    while (l < lv.n - 1 && t >= lv.H[l] * lv.W[l]) { t -= lv.H[l] * lv.W[l]; ++l; }
    const int y = t / lv.W[l], x = t % lv.W[l];
    const float vh = (float)valid_hw[((size_t)b * lv.n + l) * 2], vw = (float)valid_hw[((size_t)b * lv.n + l) * 2 + 1];
    float p[4];
    p[0] = ((float)x + 0.5f) / vw;
    p[1] = ((float)y + 0.5f) / vh;
    p[2] = p[3] = default_hw * exp2f((float)l);
    bool valid = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) valid = valid && (p[k] > 0.01f) && (p[k] < 0.99f);
    const bool keep = valid && !pad[tok];
    if (lane < 4) proposals[tok * 4 + lane] = keep ? logf(p[lane] / (1.f - p[lane])) : INFINITY;
    const T* src = memory + (size_t)tok * C;
    T* dst = out_memory + (size_t)tok * C;
    if (((size_t)C * sizeof(T)) % 16 == 0 && ((((uintptr_t)memory) | ((uintptr_t)out_memory)) & 15) == 0) {
        const int n16 = (int)((size_t)C * sizeof(T) / 16);
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        for (int i = lane; i < n16; i += 32) d4[i] = keep ? s4[i] : make_uint4(0, 0, 0, 0);
    } else {
        for (int c = lane; c < C; c += 32) dst[c] = keep ? src[c] : (T)0.f;
    }
}

// row-wise max over the first N columns (two-stage class score, deformable_transformer.py:345)
__global__ void rowmax_kernel(const float* __restrict__ x, int ld, int N, float* __restrict__ out, long long rows) {
    pdl_launch_dependents();
    pdl_wait();
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    float m = -INFINITY;
    for (int c = threadIdx.x & 31; c < N; c += 32) m = fmaxf(m, x[(size_t)row * ld + c]);
    m = warp_max_f(m);
    if ((threadIdx.x & 31) == 0) out[row] = m;
}

// ---------------------------------------------------------------------------------------------- decoder helpers
// reference models/dino/utils.py:141-167 on ref*valid_ratio[level 0] (deformable_transformer.py:686-691):
// (x,y,w,h) -> 512-dim sine embedding in order (y, x, w, h), T = 10000, scale 2*pi.  One block per query.
template <typename TO>
__global__ void sine_embed_kernel(const float* __restrict__ ref, const float* __restrict__ valid_ratios, TO* __restrict__ out,
                                  int B, int Q, int L) {
    pdl_launch_dependents();
    pdl_wait();
    // 64 threads per query row handle one (sin, cos) pair of each of the 4 components: dim_t[2k] == dim_t[2k+1]
    __shared__ float dim_t[64];
    if (threadIdx.x < 64) dim_t[threadIdx.x] = powf(10000.f, 2.f * (float)threadIdx.x / 128.f);
    __syncthreads();
    const int rows_per_block = blockDim.x / 64;
    const long long row = (long long)blockIdx.x * rows_per_block + threadIdx.x / 64;
    if (row >= (long long)B * Q) return;
    const int k = threadIdx.x & 63;
    const int b = (int)(row / Q);
    const float vx = valid_ratios[(size_t)b * L * 2], vy = valid_ratios[(size_t)b * L * 2 + 1];
    const float4 r = *reinterpret_cast<const float4*>(ref + row * 4);
    const float comp[4] = {r.y * vy, r.x * vx, r.z * vx, r.w * vy};   // y, x, w, h
    const float two_pi = 6.283185307179586f;
#pragma unroll
    for (int part = 0; part < 4; ++part) {
        // same operation order as the reference: (v * 2pi) / dim_t
        const float a = comp[part] * two_pi / dim_t[k];
        float sn, cs;
        sincosf(a, &sn, &cs);
        TO* o = out + row * 512 + part * 128 + 2 * k;
        stf<TO>(o, sn);
        stf<TO>(o + 1, cs);
    }
}

// bf16 throughput mode of the same embedding: arguments lie in [0, 2*pi*valid_ratio] (reference points are sigmoids), so the SFU
// sin/cos are accurate to ~1e-6 absolute -- far below the bf16 output rounding; 1/dim_t from one ex2 per thread, 32 rows per
// block, (sin, cos) pairs stored as one packed 4-byte word (256-byte coalesced rows per component).
__global__ void __launch_bounds__(256)
sine_embed_bf16_kernel(const float* __restrict__ ref, const float* __restrict__ valid_ratios, op16_t* __restrict__ out,
                       const long long rows, const int Q, const int L) {
    pdl_launch_dependents();
    pdl_wait();
    const int k = threadIdx.x & 63;
    // 2*pi / 10000^(2k/128)
    const float w = 6.283185307179586f * exp2f(-(float)k * (13.287712379549449f / 64.f));
    const long long row0 = (long long)blockIdx.x * 32 + (threadIdx.x >> 6);
#pragma unroll 2
    for (int it = 0; it < 8; ++it) {
        const long long row = row0 + it * 4;
        if (row >= rows) break;
        const int b = (int)(row / Q);
        const float vx = __ldg(valid_ratios + (size_t)b * L * 2), vy = __ldg(valid_ratios + (size_t)b * L * 2 + 1);
        const float4 r = __ldg(reinterpret_cast<const float4*>(ref + row * 4));
        const float comp[4] = {r.y * vy, r.x * vx, r.z * vx, r.w * vy};   // y, x, w, h
        uint32_t* o = reinterpret_cast<uint32_t*>(out + row * 512) + k;
#pragma unroll
        for (int part = 0; part < 4; ++part) {
            float sn, cs;
            __sincosf(comp[part] * w, &sn, &cs);
            op16x2_t t = op16_pack2(sn, cs);
            o[part * 64] = *reinterpret_cast<uint32_t*>(&t);
        }
    }
}

// new_ref = sigmoid(delta + inverse_sigmoid(ref)), eps 1e-3 (deformable_transformer.py:734-738, dino.py:343-345,
// util/misc.py:575-579).  ref_is_logit: the reference is already in logit space (two-stage init: refpoint.sigmoid()).
__global__ void box_refine_kernel(const float* __restrict__ delta, int ldd, const float* __restrict__ ref, float* __restrict__ out,
                                  long long n4) {
    pdl_launch_dependents();
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const long long row = i / 4;
    const int c = (int)(i % 4);
    float x = fminf(fmaxf(ref[i], 0.f), 1.f);
    const float x1 = fmaxf(x, 1e-3f), x2 = fmaxf(1.f - x, 1e-3f);
    const float z = delta[row * ldd + c] + logf(x1 / x2);
    out[i] = 1.f / (1.f + expf(-z));
}

__global__ void sigmoid_kernel(const float* __restrict__ x, float* __restrict__ out, long long n) {
    pdl_launch_dependents();
    pdl_wait();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = 1.f / (1.f + expf(-x[i]));
}

template <typename TI, typename TO>
__global__ void cast_kernel(const TI* __restrict__ x, TO* __restrict__ out, long long n) {
    pdl_launch_dependents();
    pdl_wait();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        stf<TO>(out + i, ldf<TI>(x + i));
}

// ---------------------------------------------------------------------------------------------- split-precision operand
// fp32 activations [M, K] (row pitch ldx) -> 16-bit [M, 3K]: [hi | hi | lo] with hi = rn16(x), lo = rn16(x - hi).  Against a weight
// matrix packed as [hi | lo | hi] (engine.py:_split_w) the ordinary 16-bit tcgen05 GEMM over K' = 3K accumulates
// hi.hi + hi.lo + lo.hi in fp32 -- the 3-term split product (2 x 11 significand bits with fp16 operands; the dropped lo.lo term is
// 2^-22 relative): the tensor-core form of the fp32 parity mode.  One thread = 8 consecutive elements of a row:
// two 16-byte loads, three 16-byte stores.
__global__ void __launch_bounds__(256)
split_cast_kernel(const float* __restrict__ x, long long ldx, op16_t* __restrict__ out, long long M, int K) {
    pdl_launch_dependents();
    pdl_wait();
    const int k8 = K >> 3;
    const long long n = M * k8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / k8;
        const int k = (int)(i - row * k8) << 3;
        const float4 a = *reinterpret_cast<const float4*>(x + row * ldx + k);
        const float4 b = *reinterpret_cast<const float4*>(x + row * ldx + k + 4);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        __align__(16) op16_t hi[8];
        __align__(16) op16_t lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            hi[j] = f32_to_op16(v[j]);
            lo[j] = f32_to_op16(v[j] - op16_to_f32(hi[j]));
        }
        op16_t* o = out + row * (3ll * K) + k;
        *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(hi);
        *reinterpret_cast<uint4*>(o + K) = *reinterpret_cast<const uint4*>(hi);
        *reinterpret_cast<uint4*>(o + 2ll * K) = *reinterpret_cast<const uint4*>(lo);
    }
}

// ---------------------------------------------------------------------------------------------- ResNet stem, direct
// conv1 7x7 / stride 2 / pad 3, 3 -> 64 channels, + folded FrozenBatchNorm + ReLU (torchvision resnet50 stem as wrapped by
// reference models/dino/backbone.py:109-128), straight from the fp32 NCHW network input to NHWC activations.  K = 147 is
// too thin for the tensor-core path to pay for a 200 MB im2col round trip, so this one convolution runs on FFMA:
// CTA = 2 x 64 output pixels x 64 channels, input patch (parity-split columns -> conflict-free stride-2 taps) and the
// 147 x 64 weight matrix in shared memory, thread = two pixels (rows r, r+2) x 32 output channels, so every weight
// vector fetched from shared memory feeds two FMAs (the one-pixel version was shared-memory bound at 37 % FMA use).
constexpr int STEM_TH = 4, STEM_TW = 64, STEM_PR = STEM_TH * 2 + 5, STEM_PC = STEM_TW * 2 + 5;   // patch 13 x 133
constexpr int STEM_PCH = (STEM_PC + 1) / 2;                                                      // columns per parity plane (67)
template <typename TO>
__global__ void __launch_bounds__(256)
stem_conv7x7_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, TO* __restrict__ out,
                    int H, int W, int Ho, int Wo) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ float stem_smem[];
    float* ws = stem_smem;                                   // [147][64]
    float* patch = stem_smem + 147 * 64;                     // [3][STEM_PR][2 parities][STEM_PCH + 1]
    constexpr int PPL = STEM_PCH + 1;
    const int b = blockIdx.z, oh0 = blockIdx.y * STEM_TH, ow0 = blockIdx.x * STEM_TW;
    for (int i = threadIdx.x; i < 147 * 64; i += 256) ws[i] = w[i];
    const int ih0 = oh0 * 2 - 3, iw0 = ow0 * 2 - 3;
    for (int i = threadIdx.x; i < 3 * STEM_PR * STEM_PC; i += 256) {
        const int c = i / (STEM_PR * STEM_PC), rem = i % (STEM_PR * STEM_PC);
        const int r = rem / STEM_PC, col = rem % STEM_PC;
        const int ih = ih0 + r, iw = iw0 + col;
        float v = 0.f;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = x[(((size_t)b * 3 + c) * H + ih) * W + iw];
        patch[((c * STEM_PR + r) * 2 + (col & 1)) * PPL + (col >> 1)] = v;
    }
    __syncthreads();
    const int pix = threadIdx.x & 127, half = threadIdx.x >> 7;
    const int pr = pix / STEM_TW, pc = pix % STEM_TW;          // pr in {0,1}; this thread also owns row pr + 2
    float acc0[32], acc1[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) { acc0[k] = 0.f; acc1[k] = 0.f; }
    for (int kh = 0; kh < 7; ++kh) {
#pragma unroll
        for (int kw = 0; kw < 7; ++kw) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int col = pc + (kw >> 1);
                const float v0 = patch[((c * STEM_PR + pr * 2 + kh) * 2 + (kw & 1)) * PPL + col];
                const float v1 = patch[((c * STEM_PR + (pr + 2) * 2 + kh) * 2 + (kw & 1)) * PPL + col];
                const float4* wp = reinterpret_cast<const float4*>(ws + ((kh * 7 + kw) * 3 + c) * 64 + half * 32);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4 w4 = wp[k];
                    acc0[4 * k] = fmaf(v0, w4.x, acc0[4 * k]);         acc1[4 * k] = fmaf(v1, w4.x, acc1[4 * k]);
                    acc0[4 * k + 1] = fmaf(v0, w4.y, acc0[4 * k + 1]); acc1[4 * k + 1] = fmaf(v1, w4.y, acc1[4 * k + 1]);
                    acc0[4 * k + 2] = fmaf(v0, w4.z, acc0[4 * k + 2]); acc1[4 * k + 2] = fmaf(v1, w4.z, acc1[4 * k + 2]);
                    acc0[4 * k + 3] = fmaf(v0, w4.w, acc0[4 * k + 3]); acc1[4 * k + 3] = fmaf(v1, w4.w, acc1[4 * k + 3]);
                }
            }
        }
    }
    const int ow = ow0 + pc;
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        const int oh = oh0 + pr + rr * 2;
        if (oh < Ho && ow < Wo) {
            TO* o = out + (((size_t)b * Ho + oh) * Wo + ow) * 64 + half * 32;
#pragma unroll
            for (int k = 0; k < 32; k += 8) {
                float v8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v8[j] = fmaxf((rr ? acc1[k + j] : acc0[k + j]) + bias[half * 32 + k + j], 0.f);
                st8<TO>(o + k, v8);
            }
        }
    }
}

static inline int grid_for(long long n, int threads) {
    long long g = (n + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 32;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace dtlr

using namespace dtlr;

#define DISPATCH_T(dtype, ...)                                          \
    if ((dtype) == DTLR_F32) { using T = float; __VA_ARGS__ }           \
    else if ((dtype) == DTLR_OP16) { using T = op16_t; __VA_ARGS__ } \
    else { set_error("unsupported dtype %d", (int)(dtype)); return DTLR_ERR_INVALID; }

// Stem patches for the tensor-core path: fp32 NCHW network input (C = 3) -> bf16 rows [B*Ho*Wo, ldo] in (kh, kw, c) order for a
// 7x7 / stride 2 / pad 3 convolution.  One CTA = 128 consecutive output pixels of one output row: the 7 x 3 input row segments
// it touches (261 columns) are staged once in shared memory as bf16 (coalesced reads), then every thread emits 16-byte chunks
// of 8 consecutive k, consecutive threads = consecutive chunks of a row (fully coalesced 304-byte rows).  HBM-bound by its
// output (B*Ho*Wo*ldo*2 bytes); the generic per-element kernel below runs 20x slower on this shape.
constexpr int SIC_SEG = 128;                     // output pixels per CTA
constexpr int SIC_COLS = 2 * SIC_SEG + 5;        // input columns touched
constexpr int SIC_ROW = SIC_COLS * 3 + 1;        // bf16 elements per staged input row (column-major pixels, channel fastest)
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const float* __restrict__ x, op16_t* __restrict__ out, const int H, const int W, const int Ho,
                   const int Wo, const int ldo) {
    pdl_launch_dependents();
    pdl_wait();
    // tile[kh][col * 3 + c]: for a fixed kh the 21 values (kw, c) of output pixel p are the 21 CONSECUTIVE elements from 6 * p
    __shared__ op16_t tile[7 * SIC_ROW];
    const int ox0 = blockIdx.x * SIC_SEG, oy = blockIdx.y, b = blockIdx.z;
    const int ix0 = 2 * ox0 - 3;
    constexpr int NLD = (21 * SIC_COLS + 255) / 256;                   // all loads of a thread in flight together
    float v[NLD];
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
        const int i = threadIdx.x + u * 256;
        const int rowi = i / SIC_COLS, col = i - rowi * SIC_COLS;      // rowi = kh * 3 + c (coalesced along col)
        const int kh = rowi / 3, c = rowi - kh * 3;
        const int iy = 2 * oy + kh - 3, ix = ix0 + col;
        v[u] = 0.f;
        if (i < 21 * SIC_COLS && iy >= 0 && iy < H && ix >= 0 && ix < W) v[u] = __ldg(x + ((size_t)(b * 3 + c) * H + iy) * W + ix);
    }
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
        const int i = threadIdx.x + u * 256;
        const int rowi = i / SIC_COLS, col = i - rowi * SIC_COLS;
        const int kh = rowi / 3, c = rowi - kh * 3;
        if (i < 21 * SIC_COLS) tile[kh * SIC_ROW + col * 3 + c] = f32_to_op16(v[u]);
    }
    __syncthreads();
    const int nchunk = ldo / 8;                                        // 16-byte chunks per output row
    const int npx = min(SIC_SEG, Wo - ox0);
    const size_t row0 = ((size_t)b * Ho + oy) * Wo + ox0;
    const unsigned short* t16 = reinterpret_cast<const unsigned short*>(tile);
    for (int idx = threadIdx.x; idx < npx * nchunk; idx += 256) {
        const int p = idx / nchunk, j = idx - p * nchunk;
        int k = 8 * j;
        int kh = k / 21, r = k - kh * 21;                              // walk (kh, r) incrementally over the chunk's 8 elements
        int a = kh * SIC_ROW + 6 * p + r;
        uint32_t w4[4];
#pragma unroll
        for (int e2 = 0; e2 < 4; ++e2) {
            uint32_t h[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                h[q] = (k < 147) ? (uint32_t)t16[a] : 0u;
                ++k; ++r; ++a;
                if (r == 21) { r = 0; a += SIC_ROW - 21; }
            }
            w4[e2] = h[0] | (h[1] << 16);
        }
        *reinterpret_cast<uint4*>(out + (row0 + p) * ldo + 8 * j) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
    }
}

extern "C" int dtlr_im2col(const void* x, void* out, int B, int H, int W, int C, int KH, int KW, int stride, int pad,
                           int Ho, int Wo, int ldo, int in_dtype, int out_dtype, int nchw_input, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    DTLR_CHECK_ARG(ldo >= KH * KW * C, "im2col: ldo too small");
    const long long total = (long long)B * Ho * Wo * ldo;
    if (total == 0) return DTLR_OK;
    if (!nchw_input && in_dtype == out_dtype && C % 8 == 0 && ldo == KH * KW * C) {
        const long long t8 = total / 8;
        DISPATCH_T(in_dtype, DTLR_LAUNCH((im2col_vec8_kernel<T>), grid_for(t8, 256), 256, 0, st, (const T*)x, (T*)out, B, H, W, C, KH, KW, stride, pad, Ho, Wo);)
    } else if (nchw_input && in_dtype == DTLR_F32 && out_dtype == DTLR_OP16 && C == 3 && KH == 7 && KW == 7 && stride == 2 && pad == 3 &&
               (ldo % 8) == 0 && (((uintptr_t)out) & 15) == 0 && B <= 65535 && Ho <= 65535) {
        dim3 grid((Wo + SIC_SEG - 1) / SIC_SEG, Ho, B);
        DTLR_LAUNCH((stem_im2col_kernel), grid, 256, 0, st, (const float*)x, (op16_t*)out, H, W, Ho, Wo, ldo);
    } else if (nchw_input && in_dtype == DTLR_F32) {
        DISPATCH_T(out_dtype, DTLR_LAUNCH((im2col_kernel<float, T, true>), grid_for(total, 256), 256, 0, st, (const float*)x, (T*)out, B, H, W, C, KH, KW, stride, pad, Ho, Wo, ldo);)
    } else if (!nchw_input && in_dtype == out_dtype) {
        DISPATCH_T(in_dtype, DTLR_LAUNCH((im2col_kernel<T, T, false>), grid_for(total, 256), 256, 0, st, (const T*)x, (T*)out, B, H, W, C, KH, KW, stride, pad, Ho, Wo, ldo);)
    } else {
        set_error("im2col: unsupported dtype/layout combination");
        return DTLR_ERR_UNSUPPORTED;
    }
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_stem_conv(const float* x, const float* w, const float* bias, void* out, int B, int H, int W, int Ho, int Wo,
                              int out_dtype, void* stream) {
    DTLR_CHECK_ARG(Ho == (H + 6 - 7) / 2 + 1 && Wo == (W + 6 - 7) / 2 + 1, "stem_conv: output size does not match 7x7/s2/p3");
    if (B == 0) return DTLR_OK;
    DTLR_CHECK_ARG(B <= 65535, "stem_conv: batch too large");
    const size_t smem = (size_t)(147 * 64 + 3 * STEM_PR * 2 * (STEM_PCH + 1)) * sizeof(float);
    dim3 grid((Wo + STEM_TW - 1) / STEM_TW, (Ho + STEM_TH - 1) / STEM_TH, B);
    cudaStream_t st = (cudaStream_t)stream;
    if (out_dtype == DTLR_F32) {
        DTLR_CHECK_CUDA(cudaFuncSetAttribute(stem_conv7x7_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DTLR_LAUNCH((stem_conv7x7_kernel<float>), grid, 256, smem, st, x, w, bias, (float*)out, H, W, Ho, Wo);
    } else if (out_dtype == DTLR_OP16) {
        DTLR_CHECK_CUDA(cudaFuncSetAttribute(stem_conv7x7_kernel<op16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        DTLR_LAUNCH((stem_conv7x7_kernel<op16_t>), grid, 256, smem, st, x, w, bias, (op16_t*)out, H, W, Ho, Wo);
    } else { set_error("stem_conv: unsupported dtype"); return DTLR_ERR_INVALID; }
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_maxpool3x3s2(const void* x, void* out, int B, int H, int W, int C, int Ho, int Wo, int dtype, void* stream) {
    DTLR_CHECK_ARG(C % 8 == 0, "maxpool: C must be a multiple of 8");
    const long long total = (long long)B * Ho * Wo * (C / 8);
    if (total == 0) return DTLR_OK;
    DISPATCH_T(dtype, DTLR_LAUNCH((maxpool3x3s2_kernel<T>), grid_for(total, 256), 256, 0, (cudaStream_t)stream, (const T*)x, (T*)out, B, H, W, C, Ho, Wo);)
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_groupnorm(const float* x, const float* gamma, const float* beta, void* out, int B, int HW, int C, int G,
                              long long out_stride_b, float eps, int out_dtype, void* stream) {
    DTLR_CHECK_ARG(C % G == 0, "groupnorm: C %% G != 0");
    if (B == 0 || HW == 0) return DTLR_OK;
    if (C / G == 8 && (C % 32) == 0 && ((((uintptr_t)x | (uintptr_t)gamma | (uintptr_t)beta)) & 15) == 0 && !(g_debug_flags & 524288)) {
        dim3 grid8(C / 32, B);          // flag 524288: the one-CTA-per-group kernel (A/B)
        DISPATCH_T(out_dtype, DTLR_LAUNCH((groupnorm8_kernel<T, 20>), grid8, 256, 0, (cudaStream_t)stream, x, gamma, beta, (T*)out, HW, C, out_stride_b, eps);)
        DTLR_CHECK_LAUNCH();
        return DTLR_OK;
    }
    dim3 grid(G, B);
    DISPATCH_T(out_dtype, DTLR_LAUNCH((groupnorm_kernel<T>), grid, 256, 0, (cudaStream_t)stream, x, gamma, beta, (T*)out, HW, C, G, out_stride_b, eps);)
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_pos_sine(const unsigned char* mask, const float* level_embed, void* out, int B, int H, int W, int npf,
                             float temp_h, float temp_w, long long out_stride_b, int out_dtype, void* stream) {
    const long long total = (long long)B * H * W;
    if (total == 0) return DTLR_OK;
    DISPATCH_T(out_dtype, DTLR_LAUNCH((pos_sine_kernel<T>), (unsigned)total, 128, 0, (cudaStream_t)stream, mask, level_embed, (T*)out, B, H, W, npf, temp_h, temp_w, out_stride_b);)
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_add_layernorm(const void* x, const void* res, const float* gamma, const float* beta, void* y, const void* add2,
                                  void* y2, long long rows, int C, float eps, int dtype, void* stream) {
    DTLR_CHECK_ARG(C == 256, "add_layernorm: only C=256 (d_model of every DTLR config) is implemented, got %d", C);
    if (rows == 0) return DTLR_OK;
    const int wpb = 8;
    if (rows >= 4096 && !(g_debug_flags & 67108864)) {       // two rows per warp (flag 67108864: one, A/B)
        const unsigned grid2 = (unsigned)((rows + 2 * wpb - 1) / (2 * wpb));
        DISPATCH_T(dtype, DTLR_LAUNCH((add_layernorm256_kernel<T, 2>), grid2, wpb * 32, 0, (cudaStream_t)stream, (const T*)x, (const T*)res, gamma, beta, (T*)y, (const T*)add2, (T*)y2, (int)rows, eps);)
        DTLR_CHECK_LAUNCH();
        return DTLR_OK;
    }
    const unsigned grid = (unsigned)((rows + wpb - 1) / wpb);
    DISPATCH_T(dtype, DTLR_LAUNCH((add_layernorm256_kernel<T, 1>), grid, wpb * 32, 0, (cudaStream_t)stream, (const T*)x, (const T*)res, gamma, beta, (T*)y, (const T*)add2, (T*)y2, (int)rows, eps);)
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_add(const void* a, const void* b, void* out, long long n, int dtype, void* stream) {
    DTLR_CHECK_ARG(n % 8 == 0, "add: n must be a multiple of 8");
    if (n == 0) return DTLR_OK;
    DISPATCH_T(dtype, DTLR_LAUNCH((add_kernel<T>), grid_for(n / 8, 256), 256, 0, (cudaStream_t)stream, (const T*)a, (const T*)b, (T*)out, n / 8);)
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_zero_masked_rows(void* x, const unsigned char* rowmask, long long rows, int C, int dtype, void* stream) {
    DTLR_CHECK_ARG(C % 8 == 0, "zero_masked_rows: C %% 8 != 0");
    if (rows == 0) return DTLR_OK;
    DISPATCH_T(dtype, DTLR_LAUNCH((zero_masked_rows_kernel<T>), grid_for(rows * (C / 8), 256), 256, 0, (cudaStream_t)stream, (T*)x, rowmask, rows, C / 8);)
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

static int fill_prep_levels(PrepLevels& lv, const int64_t* shapes, int L) {
    DTLR_CHECK_ARG(L >= 1 && L <= 8, "n_levels %d not in [1,8]", L);
    lv.n = L;
    for (int l = 0; l < L; ++l) { lv.H[l] = (int)shapes[2 * l]; lv.W[l] = (int)shapes[2 * l + 1]; }
    return DTLR_OK;
}

extern "C" int dtlr_msda_prep(const float* proj, int ld, const float* ref, int ref_dim, const float* valid_ratios,
                              const int64_t* shapes, int L, float* loc, float* attn, int B, int Lq, int M, int P, void* stream) {
    DTLR_CHECK_ARG(ref_dim == 2 || ref_dim == 4, "msda_prep: reference points must have 2 or 4 coordinates");
    PrepLevels lv;
    int rc = fill_prep_levels(lv, shapes, L);
    if (rc) return rc;
    const long long total = (long long)B * Lq * M;
    if (total == 0) return DTLR_OK;
    if (L == 4 && P == 4 && (ld % 4) == 0 && ((((uintptr_t)proj | (uintptr_t)loc | (uintptr_t)attn)) & 15) == 0 && !(g_debug_flags & 268435456)) {
        DTLR_LAUNCH((msda_prep16_kernel), (unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream, proj, ld, ref, ref_dim, valid_ratios, loc, attn, lv, B, Lq, M);
        DTLR_CHECK_LAUNCH();
        return DTLR_OK;
    }
    DTLR_LAUNCH((msda_prep_kernel), (unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream, proj, ld, ref, ref_dim, valid_ratios, loc, attn, lv, B, Lq, M, P);
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_enc_ref_points(const float* valid_ratios, const int64_t* shapes, int L, float* ref, int B, int S, void* stream) {
    PrepLevels lv;
    int rc = fill_prep_levels(lv, shapes, L);
    if (rc) return rc;
    const long long total = (long long)B * S;
    if (total == 0) return DTLR_OK;
    DTLR_LAUNCH((enc_ref_kernel), (unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream, valid_ratios, ref, lv, B, S);
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_encoder_proposals(const void* memory, const unsigned char* pad, const int* valid_hw, const int64_t* shapes, int L,
                                      void* out_memory, float* proposals, int B, int S, int C, float default_hw, int dtype, void* stream) {
    PrepLevels lv;
    int rc = fill_prep_levels(lv, shapes, L);
    if (rc) return rc;
    const long long total = (long long)B * S;
    if (total == 0) return DTLR_OK;
    DISPATCH_T(dtype, DTLR_LAUNCH((proposals_kernel<T>), (unsigned)((total + 7) / 8), 256, 0, (cudaStream_t)stream, (const T*)memory, pad, valid_hw, (T*)out_memory, proposals, lv, B, S, C, default_hw);)
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_rowmax(const float* x, int ld, int N, float* out, long long rows, void* stream) {
    if (rows == 0) return DTLR_OK;
    DTLR_LAUNCH((rowmax_kernel), (unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream, x, ld, N, out, rows);
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_sine_embed(const float* ref, const float* valid_ratios, void* out, int B, int Q, int L, int out_dtype, void* stream) {
    const long long rows = (long long)B * Q;
    if (rows == 0) return DTLR_OK;
    if (out_dtype == DTLR_OP16 && !(g_debug_flags & 128)) {
        DTLR_LAUNCH((sine_embed_bf16_kernel), (unsigned)((rows + 31) / 32), 256, 0, (cudaStream_t)stream, ref, valid_ratios, (op16_t*)out, rows, Q, L);
        return DTLR_OK;
    }
    DISPATCH_T(out_dtype, DTLR_LAUNCH((sine_embed_kernel<T>), (unsigned)((rows + 3) / 4), 256, 0, (cudaStream_t)stream, ref, valid_ratios, (T*)out, B, Q, L);)
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_box_refine(const float* delta, int ldd, const float* ref, float* out, long long rows, void* stream) {
    const long long n4 = rows * 4;
    if (n4 == 0) return DTLR_OK;
    DTLR_LAUNCH((box_refine_kernel), (unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream, delta, ldd, ref, out, n4);
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_sigmoid(const float* x, float* out, long long n, void* stream) {
    if (n == 0) return DTLR_OK;
    DTLR_LAUNCH((sigmoid_kernel), (unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream, x, out, n);
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_split_cast(const float* x, long long ldx, void* out, long long M, int K, void* stream) {
    DTLR_CHECK_ARG(M >= 0 && K > 0 && (K % 8) == 0 && ldx >= K && (ldx % 4) == 0, "split_cast: K must be a multiple of 8 and the row pitch of 4 (K=%d)", K);
    if (M == 0) return DTLR_OK;
    DTLR_CHECK_ARG(x && out && ((((uintptr_t)x) | ((uintptr_t)out)) & 15) == 0, "split_cast: operands must be 16-byte aligned");
    const long long n = M * (K / 8);
    DTLR_LAUNCH((split_cast_kernel), grid_for(n, 256), 256, 0, (cudaStream_t)stream, x, ldx, (op16_t*)out, M, K);
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

extern "C" int dtlr_cast(const void* x, void* out, long long n, int in_dtype, int out_dtype, void* stream) {
    if (n == 0) return DTLR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (in_dtype == DTLR_F32 && out_dtype == DTLR_OP16)
        DTLR_LAUNCH((cast_kernel<float, op16_t>), grid_for(n, 256), 256, 0, st, (const float*)x, (op16_t*)out, n);
    else if (in_dtype == DTLR_OP16 && out_dtype == DTLR_F32)
        DTLR_LAUNCH((cast_kernel<op16_t, float>), grid_for(n, 256), 256, 0, st, (const op16_t*)x, (float*)out, n);
    else if (in_dtype == out_dtype && in_dtype == DTLR_F32)
        DTLR_LAUNCH((cast_kernel<float, float>), grid_for(n, 256), 256, 0, st, (const float*)x, (float*)out, n);
    else if (in_dtype == out_dtype && in_dtype == DTLR_OP16)
        DTLR_LAUNCH((cast_kernel<op16_t, op16_t>), grid_for(n, 256), 256, 0, st, (const op16_t*)x, (op16_t*)out, n);
    else { set_error("cast: unsupported dtypes"); return DTLR_ERR_INVALID; }
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}
