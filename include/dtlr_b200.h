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

/* DTLR_BF16 is served by libdtlr_b200.so, DTLR_F16 by libdtlr_b200_f16.so (the same sources built with fp16 as the 16-bit type);
 * each library rejects the other 16-bit code. */
typedef enum { DTLR_F32 = 0, DTLR_BF16 = 1, DTLR_F64 = 2, DTLR_F16 = 3,
               /* OUTPUT code of dtlr_gemm / dtlr_conv2d_nhwc[_strided] only: the fp32 result is written as the 16-bit split operand
                * [hi | hi | lo] ([M, 3N], ldc in 16-bit elements, N % 4 == 0) that dtlr_split_cast would make of it */
               DTLR_SPLIT16 = 4 } dtlr_dtype;

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

/* The same core with the MSDeformAttn prologue fused in (models/dino/ops/modules/ms_deform_attn.py:98-108 + :116-123):
 * proj fp32 [B*Lq, ld_proj] = sampling_offsets (M*L*P*2) then attention logits (M*L*P) of each query; the softmax over
 * the L*P logits and loc = ref*valid_ratio + offset/(W,H) (ref_dim 2) or + offset/P*wh*0.5 (ref_dim 4) happen in the
 * kernel, so sampling_locations / attention_weights never exist in HBM.  Needs D = 32 and L*P <= 16.
 * ref fp32 [B*Lq, ref_dim], valid_ratios fp32 [B, L, 2] = (w, h) (deformable_transformer.py:239-246, 491, 686-687). */
int dtlr_msda_forward_fused(const void* value, int value_ld /* elements per pixel row, >= M*32 */, const int64_t* shapes, const int64_t* lsi, const void* proj, int ld_proj,
                            int proj_dtype /* DTLR_F32 or DTLR_BF16 */, const float* ref, int ref_dim, const float* valid_ratios, void* out, int B, int S, int M,
                            int D, int L, int Lq, int P, int dtype, void* stream);

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
 * When N is no multiple of 16 bytes of output elements and ldc is exactly N rounded up to that multiple (a padded row pitch),
 * the pad columns of C are scratch: the TMA-store epilogue may write zeros there.
 * out_dtype DTLR_SPLIT16 (16-bit operands only): C is 16-bit [M, 3N] = [hi | hi | lo] of the fp32 result (residual, if any, fp32) --
 * the A operand of the next split-precision product (dtlr_split_cast) without the fp32 round trip.
 */
int dtlr_gemm(const void* A, int lda, const void* W, int ldw, const float* bias, const void* residual, int ldr,
              void* C, int ldc, int M, int N, int K, int in_dtype, int out_dtype, int relu, void* stream);
/* Linear -> (+residual) -> LayerNorm(256) [-> second output y + add2] in one tcgen05 kernel (bf16 operands, N = 256 = d_model):
 * Y = LN(A.W^T + bias + residual) * gamma + beta;  Y2 = Y + add2 when Y2 != NULL.  Replaces the nn.Linear + dropout(0) + residual +
 * nn.LayerNorm tail of every block: models/dino/deformable_transformer.py:813-814, 806-807, 906-907, 956-957, 878-879, 326. */
int dtlr_gemm_ln(const void* A, int lda, const void* W, int ldw, const float* bias, const void* residual, int ldr,
                 const float* gamma, const float* beta, float eps, void* Y, int ldy, const void* add2, void* Y2, int ld2,
                 int M, int K, void* stream);
/* The whole position-wise FFN block of a transformer layer in ONE tcgen05 kernel (bf16, d_model 256, hidden <= 2048 and a
 * multiple of 128):  Y = LN(X + W2 . relu(W1 . X + b1) + b2) * gamma + beta.  The hidden activation lives only in TMEM / shared
 * memory.  Replaces forward_ffn + norm of models/dino/deformable_transformer.py:804-808,816-817 (encoder) and :876-880 (decoder).
 * X [M,ldx], W1 [hidden,ldw1] (256 used), W2 [256,ldw2] (hidden used), Y [M,ldy]; b1 [hidden], b2/gamma/beta [256] fp32. */
int dtlr_ffn_ln(const void* X, int ldx, const void* W1, int ldw1, const float* b1, const void* W2, int ldw2, const float* b2,
                const float* gamma, const float* beta, float eps, void* Y, int ldy, int M, int hidden, void* stream);
/* The same block without the wave-quantisation loss: with more than one round of 128-row tiles per SM the call runs
 * the stream-K kernel -- every CTA owns an equal range of (row tile, 128-wide hidden chunk) units; a tile that straddles a range
 * boundary is summed from two neighbouring CTAs' fp32 partials, exchanged through `workspace` behind a ready flag.
 * dtlr_ffn_workspace_bytes(M, hidden) = the workspace the plan for this shape needs (0: none; then, or with a NULL / too small
 * workspace, the call is dtlr_ffn_ln).  The first 1024 bytes of the workspace are the flags: they must be ZERO before the first
 * call and are handed back zero by every completed call; calls that share a workspace must be ordered on one stream.
 * Deterministic (fixed summation order).  dtlr_ffn_plan: 0 plain, 1 full rounds + split tail (A/B), 2 stream-K. */
long long dtlr_ffn_workspace_bytes(int M, int hidden);
int dtlr_ffn_plan(int M, int hidden);
int dtlr_ffn_ln_ws(const void* X, int ldx, const void* W1, int ldw1, const float* b1, const void* W2, int ldw2,
                   const float* b2, const float* gamma, const float* beta, float eps, void* Y, int ldy, int M, int hidden,
                   void* workspace, long long workspace_bytes, void* stream);
/* tuning aid only: sets kernel debug flags (0 = normal operation), returns the previous value */
int dtlr_debug_flags(int flags);
/* relu: 0 none, 1 ReLU before the residual add (FFN linear1), 2 ReLU after it (ResNet bottleneck output), 3 ReLU BACKWARD: the
 * `residual` operand is the saved forward activation h and the result is masked, C = h > 0 ? A.W^T : 0 (dgrad of a Linear + ReLU pair) */

/* ---------------------------------------------------------------------------------------------------------
 * Convolution front end (NHWC activations).  A k x k / strided convolution of the ResNet-50 trunk (torchvision
 * resnet50 as wrapped by models/dino/backbone.py:109-128) or of input_proj[3] (models/dino/dino.py:126-135) is
 * dtlr_im2col + dtlr_gemm with the FrozenBatchNorm (backbone.py:62-72) folded into the weights and bias.
 * x: NHWC [B,H,W,C] of in_dtype, or (nchw_input=1) the fp32 NCHW network input; out [B*Ho*Wo, ldo] of out_dtype,
 * K ordered (kh, kw, cin), columns >= KH*KW*C zero-filled.
 */
int dtlr_im2col(const void* x, void* out, int B, int H, int W, int C, int KH, int KW, int stride, int pad, int Ho,
                int Wo, int ldo, int in_dtype, int out_dtype, int nchw_input, void* stream);
/* Stride-1 "same" k x k convolution on NHWC bf16 activations as an implicit GEMM on the tcgen05 kernel: the taps are TMA
 * loads from x [B,H,W,C] with shifted coordinates and hardware zero fill (no im2col matrix).  w bf16 [Cout, KH*KW*C]
 * (K ordered kh, kw, cin; FrozenBatchNorm folded), bias fp32 [Cout] or NULL, residual/out bf16 [B*H*W, Cout].
 * Needs C % 64 == 0 and W a power of two <= 128 or a multiple of 128; other shapes use dtlr_im2col + dtlr_gemm.
 * (the 3x3 conv2 of every torchvision Bottleneck, models/dino/backbone.py:109-128) */
int dtlr_conv2d_nhwc(const void* x, const void* w, const float* bias, const void* residual, void* out, int B, int H, int W,
                     int C, int Cout, int KH, int KW, int pad, int relu, int out_dtype, void* stream);
/* The same for stride 1 or 2 (the stride-2 3x3 conv2 and the strided 1x1 downsample of the first Bottleneck of layer2-4): the A tensor
 * map traverses W and H with element stride 2, so no im2col matrix is built.  Hin, Win: INPUT size; output (Hin + 2 pad - KH) / stride
 * + 1 by (Win + 2 pad - KW) / stride + 1, its width subject to the tiling rule above; residual / out [B*Hout*Wout, Cout].
 * out_dtype: the 16-bit type, or DTLR_F32 (fp32 result and residual), or DTLR_SPLIT16 (out [B*Hout*Wout, 3*Cout] = [hi | hi | lo] of the
 * fp32 result, Cout % 8 == 0) -- the split-precision mode runs its 3x3 convs here on pixels of 3C channels with per-tap [hi | lo | hi]
 * weights. */
int dtlr_conv2d_nhwc_strided(const void* x, const void* w, const float* bias, const void* residual, void* out, int B, int Hin,
                             int Win, int C, int Cout, int KH, int KW, int pad, int stride, int relu, int out_dtype, void* stream);
/* ResNet stem: conv1 7x7/s2/p3 (3->64) + folded FrozenBatchNorm + ReLU, direct (no im2col): x fp32 NCHW [B,3,H,W],
 * w fp32 [7][7][3][64] (BN scale folded), bias fp32 [64] -> out NHWC [B*Ho*Wo, 64] of out_dtype.
 * (torchvision resnet50.conv1/bn1/relu as wrapped by models/dino/backbone.py:109-128, FrozenBatchNorm2d :62-72) */
int dtlr_stem_conv(const float* x, const float* w, const float* bias, void* out, int B, int H, int W, int Ho, int Wo,
                   int out_dtype, void* stream);
/* 3x3 / stride 2 / pad 1 max-pool of the ResNet stem, NHWC, C % 8 == 0 */
int dtlr_maxpool3x3s2(const void* x, void* out, int B, int H, int W, int C, int Ho, int Wo, int dtype, void* stream);
/* nn.GroupNorm(G, C), eps (models/dino/dino.py:121-124) over one feature level: x fp32 [B,HW,C] -> out rows
 * (b*out_stride_b + hw) of a [.., C] buffer of out_dtype, i.e. directly into the flattened multi-level token tensor
 * (models/dino/deformable_transformer.py:278-288). */
int dtlr_groupnorm(const float* x, const float* gamma, const float* beta, void* out, int B, int HW, int C, int G,
                   long long out_stride_b, float eps, int out_dtype, void* stream);
/* PositionEmbeddingSineHW (models/dino/position_encoding.py:79-108, normalize=True) + level_embed[C] (may be NULL;
 * deformable_transformer.py:281-282).  mask [B,H,W] uint8, 1 = padding; out as dtlr_groupnorm; C = 2*npf. */
int dtlr_pos_sine(const unsigned char* mask, const float* level_embed, void* out, int B, int H, int W, int npf,
                  float temp_h, float temp_w, long long out_stride_b, int out_dtype, void* stream);
/* y = LayerNorm_256(x (+ res)), eps; optional y2 = y + add2 (the "+pos" query of the next block).
 * models/dino/deformable_transformer.py:813-814,806-807,906-907,956-957,878-879,326,758.  res/add2/y2 may be NULL. */
int dtlr_add_layernorm(const void* x, const void* res, const float* gamma, const float* beta, void* y,
                       const void* add2, void* y2, long long rows, int C, float eps, int dtype, void* stream);
int dtlr_add(const void* a, const void* b, void* out, long long n, int dtype, void* stream);
/* value.masked_fill(padding_mask[..., None], 0) (models/dino/ops/modules/ms_deform_attn.py:95-96) */
int dtlr_zero_masked_rows(void* x, const unsigned char* rowmask, long long rows, int C, int dtype, void* stream);
/* softmax over the L*P attention logits of each head + sampling locations (ms_deform_attn.py:98-108) with the
 * per-level valid-ratio scaling of the reference points folded in (deformable_transformer.py:491, 686-687).
 * proj fp32 [B*Lq, ld]: M*L*P*2 offsets then M*L*P logits; ref fp32 [B*Lq, ref_dim] (2: encoder, 4: decoder boxes);
 * valid_ratios fp32 [B,L,2]; shapes host (L,2); loc [B,Lq,M,L,P,2], attn [B,Lq,M,L,P] fp32 out. */
int dtlr_msda_prep(const float* proj, int ld, const float* ref, int ref_dim, const float* valid_ratios,
                   const int64_t* shapes, int L, float* loc, float* attn, int B, int Lq, int M, int P, void* stream);
/* TransformerEncoder.get_reference_points (deformable_transformer.py:479-490) before the valid-ratio product */
int dtlr_enc_ref_points(const float* valid_ratios, const int64_t* shapes, int L, float* ref, int B, int S, void* stream);
/* gen_encoder_output_proposals (models/dino/utils.py:15-64): proposals fp32 [B,S,4] (logit space, +inf when invalid),
 * out_memory = memory with invalid / padded rows zeroed.  valid_hw int32 [B,L,2] = (valid_H, valid_W). */
int dtlr_encoder_proposals(const void* memory, const unsigned char* pad, const int* valid_hw, const int64_t* shapes,
                           int L, void* out_memory, float* proposals, int B, int S, int C, float default_hw, int dtype,
                           void* stream);
/* max over the first N columns of each row (two-stage class score, deformable_transformer.py:345) */
int dtlr_rowmax(const float* x, int ld, int N, float* out, long long rows, void* stream);
/* gen_sineembed_for_position (models/dino/utils.py:141-167) of ref*valid_ratio[:,0] (deformable_transformer.py:686-691):
 * ref fp32 [B*Q,4] -> out [B*Q,512] in order (y,x,w,h) */
int dtlr_sine_embed(const float* ref, const float* valid_ratios, void* out, int B, int Q, int L, int out_dtype, void* stream);
/* out = sigmoid(delta[:, :4] + inverse_sigmoid(ref, eps=1e-3)) (deformable_transformer.py:734-738, dino.py:343-345) */
int dtlr_box_refine(const float* delta, int ldd, const float* ref, float* out, long long rows, void* stream);
int dtlr_sigmoid(const float* x, float* out, long long n, void* stream);
int dtlr_cast(const void* x, void* out, long long n, int in_dtype, int out_dtype, void* stream);
/* Split-precision A operand of the tensor-core parity mode: x fp32 [M,K] (row pitch ldx floats) -> out 16-bit [M,3K] = [hi | hi | lo],
 * hi = rn16(x), lo = rn16(x - hi).  dtlr_gemm of it against a weight packed [hi | lo | hi] (K' = 3K) is the 3-term split product
 * hi.hi + hi.lo + lo.hi accumulated in fp32; replaces the fp32 nn.Linear / conv contractions of the reference forward
 * (models/dino/deformable_transformer.py:806-814,951-957, backbone.py:109-128) at ~2^-22 operand precision.  K % 8 == 0. */
int dtlr_split_cast(const float* x, long long ldx, void* out, long long M, int K, void* stream);
/* nn.MultiheadAttention(d_model, heads) core of the decoder self-attention (deformable_transformer.py:847, 903-905):
 * softmax(q k^T / sqrt(head_dim) [masked]) v per (batch, head), scores never materialised.
 * q rows: qk[b*Q+i, h*32 ...], k rows: qk[b*Q+j, k_off + h*32 ...], v rows: v[b*Q+j, h*32 ...];
 * attn_mask uint8 [Q,Q], 1 = blocked, or NULL; out [B*Q, ld_o]. */
int dtlr_mha_self_attention(const void* qk, int ld_qk, int k_off, const void* v, int ld_v, const unsigned char* attn_mask,
                            void* out, int ld_o, int B, int Q, int heads, int head_dim, int dtype, void* stream);
/* tuning hook of the 16-bit flash kernel above: warps per CTA (1..20) and query splits per (image, head); (0, 0) = automatic (the
 * partition with the fewest warp-rounds, see csrc/attention.cu) */
int dtlr_attn_config(int warps, int splits);
/* The same attention on the tcgen05 tensor cores (bf16, no mask, head_dim 32, Q <= 1024): K / V^T resident in shared memory via
 * TMA, S = QK^T and O += PV as tcgen05.mma with TMEM accumulators, two-pass softmax between them.  vt_scratch: device buffer of
 * B*heads*32*1024 bf16 (V^T, written by a pre-pass).  Returns DTLR_ERR_UNSUPPORTED for shapes it does not cover. */
int dtlr_mha_tcgen05(const void* qk, int ld_qk, int k_off, const void* v, int ld_v, void* vt_scratch, void* out, int ld_o, int B,
                     int Q, int heads, int head_dim, void* stream);
/* The "CTC view" decode tail, fused (models/dino/dino.py:472-502 transform + engine.py:512-530 argmax; evaluation.py:116-158
 * uses eps = 0.03/C): frames[b,pos] = argmax over [blank, classes] of the pos-th query in cx order (0 = blank, c+1 = class c).
 * logits fp32 [B*Q, ld], boxes fp32 [B*Q,4] (cx first).  perm (int32 [B,Q], may be NULL) receives the sort permutation,
 * new_pred (fp32 [B,Q,C+1], may be NULL; needs perm + scratch_sum) the full new_pred_logits tensor.  scratch_label int32 [B*Q],
 * scratch_sum fp32 [B*Q] (may be NULL when new_pred is NULL). */
int dtlr_ctc_decode(const float* logits, int ld, const float* boxes, int* frames, int* perm, float* new_pred,
                    int* scratch_label, float* scratch_sum, int B, int Q, int C, float eps, void* stream);
/* Same with every class probability multiplied by prob_scale before the blank synthesis: the layout the n-gram rescoring
 * tool feeds to its CTC beam-search decoder (ngram/prediction_helpers.py:5-46, `multiply_pred_logits_by`). */
int dtlr_ctc_decode_scaled(const float* logits, int ld, const float* boxes, int* frames, int* perm, float* new_pred,
                           int* scratch_label, float* scratch_sum, int B, int Q, int C, float eps, float prob_scale,
                           void* stream);

/* The 3-layer box MLP (models/dino/utils.py:110-122 as instantiated for bbox_embed / enc_out_bbox_embed: 256 -> 256 -> 256 -> 4, ReLU
 * between) fused with the iterative box refinement that consumes it (models/dino/deformable_transformer.py:734-738 and the repeated
 * evaluation in models/dino/dino.py:343-345): out4[r] = sigmoid(MLP(X[r]) + inverse_sigmoid(ref[r])) with inverse_sigmoid of
 * util/misc.py:575-579 (eps 1e-3), or the raw MLP output when ref is NULL.  One tcgen05 kernel (the FFN kernel's HEAD variant): the two
 * 256-wide intermediates never reach HBM.  X [M, ldx] / W1 / W2 [256, ld] 16-bit (the library's flavour), b1 / b2 fp32 [256],
 * W3 fp32 [4,256], b3 fp32 [4], ref fp32 [M,4] or NULL, out4 fp32 [M,4]. */
int dtlr_mlp_head(const void* X, int ldx, const void* W1, int ldw1, const float* b1, const void* W2, int ldw2,
                  const float* b2, const float* W3, const float* b3, const float* ref, float* out4, int M, void* stream);

/* Two-stage query selection (models/dino/deformable_transformer.py:345-353): idx int64 [B,K] = torch.topk(scores [B,S], K, dim=1)[1]
 * (descending, ties resolved to the lowest token index); fails like torch.topk when K > S (reference quirk: 40x704 lines have 627 < 900
 * tokens).  dtlr_select_gather then performs what follows it in one pass (:348-353, :676): refpoint [B,K,4] = sigmoid((coord +
 * prop)[idx]) -- coord = enc_out_bbox_embed output, prop = the logit-space anchors (+inf where invalid -> 1.0) --, initbox [B,K,4] =
 * sigmoid(prop[idx]), tgt [B,K,d] = mem[idx] (mem / tgt of `dtype`, the others fp32). */
int dtlr_topk_select(const float* scores, int B, int S, int K, long long* idx, void* stream);
int dtlr_select_gather(const long long* idx, const float* coord, const float* prop, const void* mem, float* refpoint,
                       float* initbox, void* tgt, int B, int S, int K, int d, int dtype, void* stream);
/* PostProcess.forward (models/dino/dino.py:1008-1046) and, with `keep`, its NMS branch plus the reading of evaluation.py:94-115:
 * scores [B,K] / labels int32 [B,K] / boxes_out [B,K,4] = the K largest of the Q*C sigmoid scores of each line (descending; ties ->
 * lowest flat index q*C + c), label = index % C, box of query index / C converted (box_mode 0: cxcywh -> xyxy, 1: as stored,
 * 2: x0,y0,w,h -- the reference's `test` flag) and scaled by sizes [B,2] = (img_h, img_w).  logits fp32 [B*Q, ld], boxes fp32 [B*Q,4].
 * keep (uint8 [B,K], may be NULL): class-agnostic greedy NMS with torchvision.ops.nms semantics at IoU > nms_iou (<= 0: keep all);
 * read_labels int32 [B,K] / read_count int32 [B]: labels of the kept detections with score > score_thr in order of box centre
 * (x0+x1)/2 (-1 padded).  K <= 1024 when keep is given. */
int dtlr_postprocess(const float* logits, int ld, const float* boxes, const float* sizes, int B, int Q, int C, int K,
                     int box_mode, float nms_iou, float score_thr, float* scores, int* labels, float* boxes_out,
                     unsigned char* keep, int* read_labels, int* read_count, void* stream);

/* The CTC loss of the fine-tuning step, fused forward + backward (SetCriterion.loss_CTC, models/dino/dino.py:457-551: cx sort, sigmoid,
 * blank synthesis with eps, one hard-blank frame (1, 1e-5, ...) interleaved after every query (:505-517), log, nn.CTCLoss(blank=0,
 * zero_infinity, reduction 'mean') (:538-544)) WITHOUT the (B,2Q,C+1) tensors: only the blank and target-label probabilities of each
 * frame enter a per-line alpha/beta lattice kernel.
 * logits fp32 [B*Q, ld], boxes fp32 [B*Q,4]; targets dev int32 [B, Lmax] class ids 0..C-1 (padding ignored), target_len dev int32 [B]
 * (Lmax <= 511).  nll fp32 [B] receives -log p(target | line) (+inf when no alignment exists); grad_logits fp32 [B,Q,C] (may be NULL)
 * receives d(mean_b nll_b / max(len_b,1)) / d logits in the ORIGINAL query order (zero rows for infeasible lines when zero_infinity).
 * Scratch (device): perm int32 [B*Q] (the cx permutation, also an output), row_sum fp32 [B*Q], scratch_label / scratch_frames int32
 * [B*Q], lp fp32 [B*Q*(Lmax+1)], alpha and gext fp32 [B*Q*(2*Lmax+1)] each. */
int dtlr_ctc_loss(const float* logits, int ld, const float* boxes, const int* targets, const int* target_len, int Lmax,
                  float eps, int zero_infinity, float* nll, float* grad_logits, int* perm, float* row_sum,
                  int* scratch_label, int* scratch_frames, float* lp, float* alpha, float* gext, int B, int Q, int C,
                  void* stream);

/* GPU input stage for already-resized 8-bit line images: ToTensor (datasets/transforms.py:247-249) + Normalize
 * (datasets/transforms.py:552-558) + nested_tensor_from_tensor_list (util/misc.py:375-397) in one kernel.
 * packed (dev u8): the B images back to back, image b at byte offsets[b] (dev int64 [B]), h x w x channels, row-major, channels
 * interleaved (PIL layout; channels = 1: grayscale replicated to the 3 planes as datasets/IAM.py:86-88 does, 3: RGB).
 * hw (dev int32 [B,2]) = (h, w) of every image, h <= Hmax, w <= Wmax.  out (dev f32 [B,3,Hmax,Wmax]) = ((u8/255) - mean) / std in
 * IEEE fp32, bit-identical to the torch chain, zero on the padding; mask (dev u8 [B,Hmax,Wmax]) = 1 on the padding.
 * mean3/std3 are HOST pointers to 3 floats. */
int dtlr_preprocess_u8(const uint8_t* packed, const long long* offsets, const int* hw, int channels, float* out, uint8_t* mask,
                       int B, int Hmax, int Wmax, const float* mean3_host, const float* std3_host, void* stream);
/* 8-bit bilinear resize of a ragged batch with PIL's arithmetic (torchvision F.resize on PIL images, datasets/transforms.py:107-108;
 * Pillow src/libImaging/Resample.c): horizontal pass in -> tmp, vertical pass tmp -> out, 22-bit fixed-point coefficients.
 * meta (dev int64 [B,12]) per image: byte offsets of the image in `in`, `tmp`, `out`; h, w, oh, ow; int32-element offsets of its
 * horizontal and vertical tables in `tables`; their kernel sizes ksx, ksy; one reserved word.  A table of an axis with n outputs is
 * xmin[n], count[n], coef[n*ksize] (built on the host as precompute_coeffs / normalize_coeffs_8bpc do: dtlr_b200/input.py).
 * Images are h x w x channels, interleaved.  tmp holds h x ow x channels per image, out oh x ow x channels. */
int dtlr_resize_u8_bilinear(const uint8_t* in, const long long* meta, const int* tables, uint8_t* tmp, uint8_t* out, int B,
                            int channels, int max_h, int max_oh, int max_ow, void* stream);

/* Hungarian matcher of the detection loss (models/dino/matcher.py:57-96), two kernels for all P = layers x B problems of a
 * training step at once (the reference runs a cdist/GIoU chain over the full (B*Q) x sum(T) matrix, copies it to the CPU and
 * calls scipy.optimize.linear_sum_assignment per image, once per decoder layer).
 *
 * dtlr_match_cost: block-diagonal matching cost, target-major: cost[p][t][q] fp32 [P, Tmax, Q] =
 *   w_bbox * L1(cxcywh) + w_class * focal_class_cost(alpha, gamma 2) + w_giou * (-GIoU); logits fp32 [P, Q, C], boxes fp32
 *   [P, Q, 4]; targets concatenated over the B images of one layer: tgt_labels int64 [sum T], tgt_boxes fp32 [sum T, 4],
 *   t_off / t_cnt device int32 [B]; problem p uses the targets of image p % B.  Rows t >= t_cnt are left untouched.
 * dtlr_lsap: problem p assigns its t_cnt[p % B] targets (rows of its [Tmax, Q] cost block) to distinct queries with minimal
 *   total cost; q_of_t int32 [P, Tmax] receives the query of each target (-1 pad).  One CTA per problem, shortest augmenting
 *   paths in fp64 like scipy (Tmax <= Q). */
int dtlr_match_cost(const float* logits, const float* boxes, const int64_t* tgt_labels, const float* tgt_boxes, const int* t_off,
                    const int* t_cnt, int P, int B, int Q, int C, int Tmax, float w_class, float w_bbox, float w_giou,
                    float alpha, float* cost, void* stream);
int dtlr_lsap(const float* cost, int P, int B, int Q, const int* t_cnt, int Tmax, int* q_of_t, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Backward half of the fine-tune step (engine.py:172-274 train_one_epoch_CTC: loss.backward(), clip_grad_norm_(0.01), AdamW.step;
 * finetuning.py:211-231).  The reference gets these from torch autograd over cuBLAS / ATen; here (csrc/train.cu):
 *
 * dtlr_wgrad: dW[n,k] += sum_r dY[r,n] * X[r,k] -- the weight gradient of every nn.Linear (and 1x1 conv on NHWC rows).  dY [rows, ldy]
 *   (N used), X [rows, ldx] (K used) of `dtype` (the library's 16-bit type: tcgen05 with both operands MN-major, read in place; F32:
 *   exact SIMT), dW fp32 [N, ldw], ACCUMULATED with atomic reductions (the caller zeroes the gradient arena once per step; shared
 *   heads accumulate over their uses).  The dgrad dX = dY . W is dtlr_gemm(dY, W^T) against the transposed operand copy below.
 * dtlr_colsum: out[c] += sum of column c over nseg segments of seg_rows rows, segment s starting at row s*seg_stride (bias gradients:
 *   one segment; per-level sums for level_embed: nseg = B, seg_stride = S).  x [.., ld] of `dtype`, out fp32 [N].
 * dtlr_layernorm_bwd: nn.LayerNorm(256) backward from the saved pre-norm rows z (`dtype`), dy (+ dy2, may be NULL) fp32; dz32 (fp32)
 *   and / or dz16 (`dtype`) receive dz, dgamma / dbeta fp32 [256] are accumulated (any of the four may be NULL).
 * dtlr_relu_bwd: dh = h > 0 ? dh : 0 in place.   dtlr_add_cast: out = a (+ b) (+ c), fp32 in, fp32 or 16-bit out.
 * dtlr_msda_bwd_glue: grad of the fused projection row [offsets | logits] (ms_deform_attn.py:98-108: softmax + sampling locations)
 *   from dtlr_msda_backward's grad_loc / grad_attn; ref / valid_ratios as dtlr_msda_prep (no gradient: detached in the reference).
 * dtlr_pack_weights: one launch over a device table of fp32 matrices -> 16-bit (or fp32) copy [rows, ld_dst] and / or transposed
 *   copy [cols, ld_dstT]; entry = 8 int64 {src, rows, cols, ld_src, dst|0, ld_dst, dstT|0, ld_dstT}; tile_start int32 [n+1] = prefix
 *   of 32x32 tiles per entry.
 * dtlr_optim_begin / dtlr_grad_sumsq / dtlr_adamw: state fp32[4] on the device ([0] sum of squared gradients, [1] step count);
 *   begin zeroes [0] and advances [1]; sumsq adds one arena; adamw applies the clip coefficient min(1, max_norm / (norm + 1e-6))
 *   (max_norm <= 0: none) and the torch.optim.AdamW update to one arena slice (one learning-rate group). */
int dtlr_wgrad(const void* dY, int ldy, const void* X, int ldx, float* dW, int ldw, int rows, int N, int K, int dtype, void* stream);
int dtlr_colsum(const void* x, long long ld, int N, long long nseg, long long seg_rows, long long seg_stride, float* out, int dtype,
                void* stream);
int dtlr_layernorm_bwd(const void* z, const float* dy, const float* dy2, const float* gamma, float* dz32, void* dz16, float* dgamma,
                       float* dbeta, long long rows, int C, float eps, int dtype, void* stream);
int dtlr_relu_bwd(void* dh, const void* h, long long n, int dtype, void* stream);
int dtlr_add_cast(const float* a, const float* b, const float* c, void* out, long long n, int out_dtype, void* stream);
int dtlr_msda_bwd_glue(const float* grad_loc, const float* grad_attn, const float* attn, const float* ref, int ref_dim,
                       const float* valid_ratios, const int64_t* shapes, int L, void* dproj, int ld, int B, int Lq, int M, int P,
                       int out_dtype, void* stream);
int dtlr_pack_weights(const long long* table, const int* tile_start, int n_entries, int total_tiles, int dtype, void* stream);
int dtlr_optim_begin(float* state, void* stream);
int dtlr_grad_sumsq(const float* g, long long n, float* state, void* stream);
int dtlr_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
               float weight_decay, float max_norm, const float* state, void* stream);

/* Decoder self-attention of the fine-tune step (nn.MultiheadAttention(256, 8) under autograd with the denoising attn_mask:
 * deformable_transformer.py:847, 903-905; dn_components.py:121-141), csrc/attention_train.cu.  16-bit operands of the library's flavour,
 * head_dim 32; q rows qk[b*Q+i, h*32..], k rows qk[b*Q+j, k_off + h*32..], v rows v[b*Q+j, h*32..] as dtlr_mha_self_attention.
 * mask_bits uint32 [Q, KP/32] (KP = Q rounded up to 64; bit j of row i set = query i may not attend key j) or NULL; maskT_bits = the
 * same matrix transposed (row = key, bit = query).  forward: out [B*Q, ld_o], lse2 fp32 [B, heads, Q] = log2 sum_j exp(scale q.k)
 * (kept for the backward).  backward: dout [B*Q, ld_do] -> dqk [B*Q, ld_dqk] (dq at column h*32, dk at k_off + h*32), dv [B*Q, ld_dv];
 * D_scratch fp32 [B, heads, Q].  Probabilities are recomputed from lse2; no atomics (a query-major pass for dq, a key-major for dk, dv). */
int dtlr_mha_train_forward(const void* qk, int ld_qk, int k_off, const void* v, int ld_v, const uint32_t* mask_bits, void* out,
                           int ld_o, float* lse2, int B, int Q, int heads, int head_dim, int dtype, void* stream);
int dtlr_mha_train_backward(const void* qk, int ld_qk, int k_off, const void* v, int ld_v, const void* out, int ld_o,
                            const void* dout, int ld_do, const uint32_t* mask_bits, const uint32_t* maskT_bits, const float* lse2,
                            float* D_scratch, void* dqk, int ld_dqk, void* dv, int ld_dv, int B, int Q, int heads, int head_dim,
                            int dtype, void* stream);

/* Backward pieces of the ResNet-50 layer2-4 / input_proj front of the fine-tune step (reference: torch autograd over cuDNN for
 * models/dino/backbone.py:109-128 and models/dino/dino.py:118-135), csrc/train.cu.  1x1 convolutions on NHWC rows are Linears
 * (dtlr_wgrad / dtlr_gemm); 3x3 wgrad = dtlr_im2col + dtlr_wgrad; stride-1 3x3 dgrad = dtlr_conv2d_nhwc over dY with the flipped
 * weight copy; strided dgrad = dtlr_gemm (dcol = dY . W') + dtlr_col2im.
 * dtlr_relu_bwd_dual: v = y > 0 ? dy32 : 0 written back to dy32 (fp32) and to out16 (`dtype`, may be NULL).
 * dtlr_groupnorm_bwd: nn.GroupNorm(32, 256) backward of one level: x fp32 [B, HW, C] (saved input), dy fp32 rows at
 *   (b * dy_stride_b + hw) * C, dx [B*HW, C] of out_dtype, dgamma / dbeta fp32 [C] accumulated (may be NULL).
 * dtlr_col2im: dx fp32 [B,H,W,C] (+)= gather of dcol fp32 [B*Ho*Wo, ldc] (K order kh, kw, c) -- the transposed convolution.
 * dtlr_pack_conv: FrozenBatchNorm-folded 16-bit (or fp32) operand copies of convolution weights, entry = 10 int64 {src fp32
 *   [Cout,Cin,taps], Cout, Cin, taps, scale|0, fwd dst [Cout, taps*Cin], bwd dst|0, bwd kind (1 transpose, 2 flipped taps), bwd pitch,
 *   first element index}.  dtlr_unpack_conv_grads: grad[co][ci][t] += scale[co] * scratch[co][t*Cin+ci], entry = 6 int64 {scratch, grad,
 *   Cout, Cin, taps, scale|0}, elem_start int64 [n]. */
int dtlr_relu_bwd_dual(float* dy32, const void* y, void* out16, long long n, int dtype, void* stream);
int dtlr_groupnorm_bwd(const float* x, const float* dy, long long dy_stride_b, const float* gamma, void* dx, float* dgamma,
                       float* dbeta, int B, int HW, int C, int G, float eps, int out_dtype, void* stream);
int dtlr_col2im(const float* dcol, int ldc, float* dx, int B, int H, int W, int C, int KH, int KW, int stride, int pad, int Ho, int Wo,
                int accumulate, void* stream);
int dtlr_pack_conv(const long long* table, int n_entries, long long total_elems, int dtype, void* stream);
int dtlr_unpack_conv_grads(const long long* table, const long long* elem_start, int n_entries, long long total_elems, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DTLR_B200_H */
