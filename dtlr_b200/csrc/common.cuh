// dtlr_b200 -- shared host/device helpers for the sm_100a kernels behind include/dtlr_b200.h
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/dtlr_b200.h"

// ---- the 16-bit operand / activation type of the throughput mode.  The library is built twice from the same sources:
//   libdtlr_b200.so      op16 = bf16 (8 significand bits)   -- torch.bfloat16 models
//   libdtlr_b200_f16.so  op16 = fp16 (11 significand bits)  -- torch.float16 models: same tcgen05 / HMMA rate and the same bytes, 8x
//                        finer operand rounding (the error budget of the benchmarked mode)
// Everything that depends on the flavour is below: the types, the conversions (incl. the bit tricks bf16 allows), the PTX type
// token of mma.sync, the tcgen05 instruction-descriptor format bits, the TMA element type and the C-ABI dtype code.
#ifdef DTLR_BUILD_F16
#include <cuda_fp16.h>
namespace dtlr {
typedef __half op16_t;
typedef __half2 op16x2_t;
#define DTLR_OP16_PTX "f16"
#define DTLR_TMAP_OP16 CU_TENSOR_MAP_DATA_TYPE_FLOAT16
constexpr uint32_t OP16_IDESC_AB = 0u;                                  // A format [7,10) = F16, B format [10,13) = F16
constexpr int DTLR_OP16 = DTLR_F16;
__device__ __forceinline__ float op16_lo_f32(uint32_t u) { return __half2float(__ushort_as_half((unsigned short)(u & 0xffffu))); }
__device__ __forceinline__ float op16_hi_f32(uint32_t u) { return __half2float(__ushort_as_half((unsigned short)(u >> 16))); }
__device__ __forceinline__ op16x2_t op16_pack2(float lo, float hi) { return __floats2half2_rn(lo, hi); }
__device__ __forceinline__ float op16_to_f32(op16_t v) { return __half2float(v); }
__device__ __forceinline__ op16_t f32_to_op16(float v) { return __float2half_rn(v); }
}  // namespace dtlr
#else
namespace dtlr {
typedef __nv_bfloat16 op16_t;
typedef __nv_bfloat162 op16x2_t;
#define DTLR_OP16_PTX "bf16"
#define DTLR_TMAP_OP16 CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
constexpr uint32_t OP16_IDESC_AB = (1u << 7) | (1u << 10);              // A format [7,10) = BF16, B format [10,13) = BF16
constexpr int DTLR_OP16 = DTLR_BF16;
__device__ __forceinline__ float op16_lo_f32(uint32_t u) { return __uint_as_float(u << 16); }           // bf16 = the top half of fp32
__device__ __forceinline__ float op16_hi_f32(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ op16x2_t op16_pack2(float lo, float hi) { return __floats2bfloat162_rn(lo, hi); }
__device__ __forceinline__ float op16_to_f32(op16_t v) { return __bfloat162float(v); }
__device__ __forceinline__ op16_t f32_to_op16(float v) { return __float2bfloat16_rn(v); }
}  // namespace dtlr
#endif

namespace dtlr {

void set_error(const char* fmt, ...);
// tuning / A-B switches set through dtlr_debug_flags() (capi.cu): 1,2,4,8 = GEMM probes (gemm.cu), 16 = MSDA on the SIMT
// kernel instead of the tensor-core gather kernel (msda.cu)
extern int g_debug_flags;

inline int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

inline int max_smem_optin() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (n <= 0) n = 227 * 1024;
    }
    return n;
}

#define DTLR_CHECK_ARG(cond, ...)                    \
    do {                                             \
        if (!(cond)) {                               \
            dtlr::set_error(__VA_ARGS__);            \
            return DTLR_ERR_INVALID;                 \
        }                                            \
    } while (0)

#define DTLR_CHECK_CUDA(expr)                                                                  \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            dtlr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                            __LINE__);                                                         \
            return DTLR_ERR_CUDA;                                                              \
        }                                                                                      \
    } while (0)

#define DTLR_CHECK_LAUNCH() DTLR_CHECK_CUDA(cudaGetLastError())

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Programmatic dependent launch (PDL): a kernel launched with launch_pdl() may start while its predecessor in the stream is still
// draining; it runs its private prologue (barrier init, TMEM allocation, descriptor prefetch, constant weights) and must execute
// pdl_wait() before its first access to memory the predecessor may touch.  pdl_launch_dependents() lets the NEXT kernel start
// early (no effect when that one is launched normally).  dtlr_debug_flags(64) turns the launch attribute off (A/B).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// 2^x on the SFU without exp2f's denormal-range pre/post scaling (2 FMUL + 1 FSETP per call): in the softmax the result
// either is <= 1 with an argument <= 0 or is flushed to zero, so the plain approximation is exactly what is wanted.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// same with a thread-block cluster of `cluster_x` CTAs along x
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cluster_x, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (g_debug_flags & 64) ? 0 : 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = cluster_x;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (g_debug_flags & 64) ? 0 : 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace dtlr
