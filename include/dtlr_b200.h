/* dtlr_b200 C ABI -- the drop-in boundary of the B200-native DTLR hot path.
 *
 * Plain C, plain pointers and sizes; no torch / ATen types.  Every pointer marked "dev" is a CUDA device
 * pointer on the current device, every "host" pointer is ordinary host memory.  All entry points are
 * asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream), re-entrant, and
 * return 0 on success or a dtlr_status code; dtlr_last_error() gives the message for the calling thread.
 * Nothing here ever falls back to the CPU.
 *
 * The reference interface each entry point replaces is cited as file:line relative to raphael-baena/DTLR.
 */
#ifndef DTLR_B200_H
#define DTLR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    DTLR_OK = 0,
    DTLR_ERR_INVALID = 1,     /* bad argument (shape, dtype, alignment, null pointer)            */
    DTLR_ERR_CUDA = 2,        /* a CUDA runtime / driver call failed                             */
    DTLR_ERR_UNSUPPORTED = 3  /* valid request that this build has no kernel for                 */
} dtlr_status;

typedef enum { DTLR_F32 = 0, DTLR_BF16 = 1, DTLR_F64 = 2 } dtlr_dtype;

/* library identification: (major<<16 | minor<<8 | patch), and the SM architecture the kernels were built for */
int dtlr_version(void);
int dtlr_built_for_sm(void);
const char* dtlr_last_error(void);

/* ---------------------------------------------------------------------------------------------------------
 * Multi-scale deformable attention core.
 *
 * Replaces  MultiScaleDeformableAttention.ms_deform_attn_forward
 *           models/dino/ops/src/vision.cpp:14, src/ms_deform_attn.h:21-40,
 *           src/cuda/ms_deform_attn_cuda.cu:20-80, src/cuda/ms_deform_im2col_cuda.cuh:237-299
 *
 *   out[b,q,m,:] = sum_{l,p} attn[b,q,m,l,p] * bilinear(value_l[b,:,m,:], loc[b,q,m,l,p])
 *   pixel coords x = loc_x*W_l - 0.5, y = loc_y*H_l - 0.5; zero padding outside the map.
 *
 * value   dev  (B,S,M,D) contiguous, dtype `dtype`      (S = sum_l H_l*W_l)
 * shapes  host (L,2) int64  = (H_l, W_l)                [the reference passes a device tensor and reads it
 * lsi     host (L)   int64  = level start index          inside the kernel; here the Python shim copies it once]
 * loc     dev  (B,Lq,M,L,P,2) contiguous: fp32 when dtype is F32/BF16, fp64 when F64   (x,y) normalised
 * attn    dev  (B,Lq,M,L,P)   contiguous, same type as loc
 * out     dev  (B,Lq,M*D)     contiguous, dtype `dtype`, fully overwritten
 *
 * No im2col_step / batch divisibility requirement (reference ms_deform_attn_cuda.cu:50-52 needs
 * B % min(B,64) == 0; any B works here).  B, Lq may be 0 (no-op).
 */
int dtlr_msda_forward(const void* value, const int64_t* shapes, const int64_t* lsi, const void* loc,
                      const void* attn, void* out, int B, int S, int M, int D, int L, int Lq, int P,
                      int dtype, void* stream);

/* Replaces  MultiScaleDeformableAttention.ms_deform_attn_backward
 *           models/dino/ops/src/vision.cpp:15, src/cuda/ms_deform_attn_cuda.cu:83-153,
 *           src/cuda/ms_deform_im2col_cuda.cuh:87-159, 301-403
 * grad_out (B,Lq,M*D); grad_value (B,S,M,D), grad_loc (B,Lq,M,L,P,2), grad_attn (B,Lq,M,L,P) are fully
 * overwritten (grad_value is zero-filled inside).  dtype F32 or F64.
 */
int dtlr_msda_backward(const void* value, const int64_t* shapes, const int64_t* lsi, const void* loc,
                       const void* attn, const void* grad_out, void* grad_value, void* grad_loc,
                       void* grad_attn, int B, int S, int M, int D, int L, int Lq, int P, int dtype,
                       void* stream);


/* ---------------------------------------------------------------------------------------------------------
 * Dense contraction with fused epilogue:  C[M,N] = act(A[M,K] . W[N,K]^T + bias[N]) (+ residual[M,N])
 *
 * Replaces every torch.nn.Linear / F.linear call on the path (cuBLAS in the reference):
 *   models/dino/ops/modules/ms_deform_attn.py:55-58,94,98,99,125; models/dino/deformable_transformer.py:787-790,
 *   804-808,852-855,876-880,326,341; models/dino/utils.py:110-122 (MLP); nn.MultiheadAttention in/out projections
 *   (deformable_transformer.py:847); and, on NHWC activations, the 1x1 convolutions of backbone.py / dino.py:118-125.
 *
 * A dev [M,lda], W dev [N,ldw] (both K contiguous), bias dev fp32 [N] or NULL, residual dev [M,ldr] (dtype of C) or
 * NULL, C dev [M,ldc].  in_dtype DTLR_BF16: tcgen05 tensor-core path (fp32 accumulate), out_dtype BF16 or F32,
 * rows 16-byte aligned (lda,ldw multiples of 8).  in_dtype DTLR_F32: exact-fp32 SIMT path (parity mode), out F32.
 */
int dtlr_gemm(const void* A, int lda, const void* W, int ldw, const float* bias, const void* residual, int ldr,
              void* C, int ldc, int M, int N, int K, int in_dtype, int out_dtype, int relu, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DTLR_B200_H */
