// dtlr_b200 -- GPU input stage: ToTensor + Normalize + batch padding + padding mask in ONE kernel.
//
// Replaces, for already-resized 8-bit line images, the per-image host chain of the reference:
//   torchvision F.to_tensor (datasets/transforms.py:247-249: u8 HWC -> f32 CHW / 255),
//   F.normalize           (datasets/transforms.py:552-558: (x - mean) / std per channel),
//   nested_tensor_from_tensor_list (util/misc.py:375-397: zero-pad to the batch max H x W, mask = True on padding)
// followed by the 12-bytes-per-pixel fp32 host->device copy.  Here the batch crosses PCIe as packed u8 (1 byte per pixel
// for grayscale lines, which datasets/IAM.py:86-88 replicates to RGB on the host) and one HBM-bound kernel writes the
// (B,3,H,W) fp32 tensor and the (B,H,W) mask.  Arithmetic is IEEE fp32 division / subtraction / division in the order torch
// performs them, so the result is bit-identical to the reference chain.
// Algorithmic bytes per launch: B*Hmax*Wmax*(3*4 + 1) written + sum(h*w*channels) read.
#include "common.cuh"

namespace dtlr {

struct Norm3 { float mean[3], stdv[3]; };

// one thread = 4 consecutive output pixels of one row (Wmax % 4 == 0: three 16-byte plane stores + one 4-byte mask store per
// thread, fully coalesced); VEC = 1 is the scalar path for other widths.
template <int CH, int VEC>
__global__ void __launch_bounds__(256)
preprocess_u8_kernel(const uint8_t* __restrict__ packed, const long long* __restrict__ offsets, const int* __restrict__ hw,
                     float* __restrict__ out, uint8_t* __restrict__ mask, const int Hmax, const int Wmax, const Norm3 nm) {
    const int b = blockIdx.z, y = blockIdx.y;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (x0 >= Wmax) return;
    const int h = hw[2 * b], w = hw[2 * b + 1];
    const size_t plane = (size_t)Hmax * Wmax;
    const size_t o = (size_t)b * 3 * plane + (size_t)y * Wmax + x0;
    const uint8_t* src = packed + offsets[b] + ((size_t)y * w + x0) * CH;
    float v[3][VEC];
    uint8_t mk[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const bool in = y < h && x0 + i < w;
        mk[i] = in ? 0 : 1;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float r = 0.f;
            if (in) {
                const float u = (float)src[i * CH + (CH == 3 ? c : 0)];
                r = __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.f), nm.mean[c]), nm.stdv[c]);
            }
            v[c][i] = r;
        }
    }
    if (VEC == 4) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
            *reinterpret_cast<float4*>(out + o + c * plane) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
        *reinterpret_cast<uchar4*>(mask + (size_t)b * plane + (size_t)y * Wmax + x0) = make_uchar4(mk[0], mk[1], mk[2], mk[3]);
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) out[o + c * plane] = v[c][0];
        mask[(size_t)b * plane + (size_t)y * Wmax + x0] = mk[0];
    }
}

}  // namespace dtlr

using namespace dtlr;

extern "C" int dtlr_preprocess_u8(const uint8_t* packed, const long long* offsets, const int* hw, int channels, float* out,
                                  uint8_t* mask, int B, int Hmax, int Wmax, const float* mean3_host, const float* std3_host,
                                  void* stream) {
    DTLR_CHECK_ARG(B >= 0 && Hmax > 0 && Wmax > 0, "preprocess_u8: bad sizes");
    DTLR_CHECK_ARG(channels == 1 || channels == 3, "preprocess_u8: channels must be 1 (grayscale) or 3 (RGB, HWC)");
    DTLR_CHECK_ARG(mean3_host && std3_host, "preprocess_u8: null mean/std");
    if (B == 0) return DTLR_OK;
    DTLR_CHECK_ARG(packed && offsets && hw && out && mask, "preprocess_u8: null pointer");
    DTLR_CHECK_ARG(B <= 65535 && Hmax <= 65535, "preprocess_u8: B or Hmax exceeds 65535");
    Norm3 nm;
    for (int c = 0; c < 3; ++c) {
        nm.mean[c] = mean3_host[c];
        nm.stdv[c] = std3_host[c];
        DTLR_CHECK_ARG(nm.stdv[c] != 0.f, "preprocess_u8: std[%d] is zero", c);
    }
    cudaStream_t st = (cudaStream_t)stream;
    // 16-byte plane stores need Wmax % 4 == 0 (then every row start of the cudaMalloc-aligned planes is 16-byte aligned too)
    const bool vec = (Wmax % 4) == 0 && (((uintptr_t)out & 15) == 0) && (((uintptr_t)mask & 3) == 0);
    const int per_thread = vec ? 4 : 1;
    const int threads = 128;
    dim3 grid((Wmax + threads * per_thread - 1) / (threads * per_thread), Hmax, B), block(threads);
    if (channels == 1) {
        if (vec) preprocess_u8_kernel<1, 4><<<grid, block, 0, st>>>(packed, offsets, hw, out, mask, Hmax, Wmax, nm);
        else preprocess_u8_kernel<1, 1><<<grid, block, 0, st>>>(packed, offsets, hw, out, mask, Hmax, Wmax, nm);
    } else {
        if (vec) preprocess_u8_kernel<3, 4><<<grid, block, 0, st>>>(packed, offsets, hw, out, mask, Hmax, Wmax, nm);
        else preprocess_u8_kernel<3, 1><<<grid, block, 0, st>>>(packed, offsets, hw, out, mask, Hmax, Wmax, nm);
    }
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// 8-bit bilinear resize with PIL's arithmetic: the reference resizes every line on the host with
// torchvision F.resize -> PIL Image.resize(BILINEAR) (datasets/transforms.py:107-108, Pillow src/libImaging/Resample.c): a
// triangle filter whose support grows with the down-scale factor, coefficients in 22-bit fixed point, a horizontal pass into an
// 8-bit intermediate and a vertical pass.  The coefficient tables are built on the host in double precision exactly as
// precompute_coeffs / normalize_coeffs_8bpc do (dtlr_b200/input.py: resample_tables); the two kernels below are the integer
// multiply-accumulate passes, so the result is bit-identical to PIL.  One launch pair handles a whole ragged batch.
namespace dtlr {

constexpr int RS_META = 12;   // int64 per image: in_off, tmp_off, out_off, h, w, oh, ow, xtab_off, ytab_off, ksx, ksy, reserved
constexpr int RS_PRECISION_BITS = 32 - 8 - 2;

// table of one axis at tab + off: xmin[n_out], count[n_out], coef[n_out * ksize]
template <int CH, bool VERTICAL>
__global__ void __launch_bounds__(128)
resize_pass_kernel(const uint8_t* __restrict__ src_base, uint8_t* __restrict__ dst_base, const long long* __restrict__ meta,
                   const int* __restrict__ tab) {
    const long long* m = meta + (size_t)blockIdx.z * RS_META;
    const int h = (int)m[3], w = (int)m[4], oh = (int)m[5], ow = (int)m[6];
    const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y;
    // horizontal: (h, w) -> (h, ow);  vertical: (h, ow) -> (oh, ow)
    const int rows = VERTICAL ? oh : h;
    if (ox >= ow || oy >= rows) return;
    const uint8_t* src = src_base + (VERTICAL ? m[1] : m[0]);
    uint8_t* dst = dst_base + (VERTICAL ? m[2] : m[1]);
    const int n_out = VERTICAL ? oh : ow, ks = (int)(VERTICAL ? m[10] : m[9]);
    const int* t = tab + (VERTICAL ? m[8] : m[7]);
    const int o = VERTICAL ? oy : ox;
    const int lo = t[o], cnt = t[n_out + o];
    const int* k = t + 2 * n_out + (size_t)o * ks;
    int acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = 1 << (RS_PRECISION_BITS - 1);
    for (int i = 0; i < cnt; ++i) {
        const int kv = k[i];
        const uint8_t* p = VERTICAL ? src + ((size_t)(lo + i) * ow + ox) * CH : src + ((size_t)oy * w + lo + i) * CH;
#pragma unroll
        for (int c = 0; c < CH; ++c) acc[c] += (int)p[c] * kv;
    }
    uint8_t* q = dst + ((size_t)oy * ow + ox) * CH;
#pragma unroll
    for (int c = 0; c < CH; ++c) q[c] = (uint8_t)min(255, max(0, acc[c] >> RS_PRECISION_BITS));
}

}  // namespace dtlr

extern "C" int dtlr_resize_u8_bilinear(const uint8_t* in, const long long* meta, const int* tables, uint8_t* tmp, uint8_t* out,
                                       int B, int channels, int max_h, int max_oh, int max_ow, void* stream) {
    DTLR_CHECK_ARG(B >= 0 && max_h > 0 && max_oh > 0 && max_ow > 0, "resize_u8: bad sizes");
    DTLR_CHECK_ARG(channels == 1 || channels == 3, "resize_u8: channels must be 1 or 3");
    if (B == 0) return DTLR_OK;
    DTLR_CHECK_ARG(in && meta && tables && tmp && out, "resize_u8: null pointer");
    DTLR_CHECK_ARG(B <= 65535 && max_h <= 65535 && max_oh <= 65535, "resize_u8: B or a height exceeds 65535");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 block(128), gh((max_ow + 127) / 128, max_h, B), gv((max_ow + 127) / 128, max_oh, B);
    if (channels == 1) {
        resize_pass_kernel<1, false><<<gh, block, 0, st>>>(in, tmp, meta, tables);
        resize_pass_kernel<1, true><<<gv, block, 0, st>>>(tmp, out, meta, tables);
    } else {
        resize_pass_kernel<3, false><<<gh, block, 0, st>>>(in, tmp, meta, tables);
        resize_pass_kernel<3, true><<<gv, block, 0, st>>>(tmp, out, meta, tables);
    }
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}
