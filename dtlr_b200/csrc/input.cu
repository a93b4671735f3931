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

template <int CH>
__global__ void __launch_bounds__(256)
preprocess_u8_kernel(const uint8_t* __restrict__ packed, const long long* __restrict__ offsets, const int* __restrict__ hw,
                     float* __restrict__ out, uint8_t* __restrict__ mask, const int Hmax, const int Wmax, const Norm3 nm) {
    const int b = blockIdx.z, y = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= Wmax) return;
    const int h = hw[2 * b], w = hw[2 * b + 1];
    const bool in = y < h && x < w;
    const size_t plane = (size_t)Hmax * Wmax;
    const size_t o = (size_t)b * 3 * plane + (size_t)y * Wmax + x;
    float v[3] = {0.f, 0.f, 0.f};
    if (in) {
        const uint8_t* src = packed + offsets[b] + ((size_t)y * w + x) * CH;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float u = (float)src[CH == 3 ? c : 0];
            v[c] = __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.f), nm.mean[c]), nm.stdv[c]);
        }
    }
    out[o] = v[0];
    out[o + plane] = v[1];
    out[o + 2 * plane] = v[2];
    mask[(size_t)b * plane + (size_t)y * Wmax + x] = in ? 0 : 1;
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
    dim3 grid((Wmax + 255) / 256, Hmax, B), block(256);
    cudaStream_t st = (cudaStream_t)stream;
    if (channels == 1)
        preprocess_u8_kernel<1><<<grid, block, 0, st>>>(packed, offsets, hw, out, mask, Hmax, Wmax, nm);
    else
        preprocess_u8_kernel<3><<<grid, block, 0, st>>>(packed, offsets, hw, out, mask, Hmax, Wmax, nm);
    DTLR_CHECK_LAUNCH();
    return DTLR_OK;
}
