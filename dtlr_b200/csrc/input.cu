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
