/* vqa_b200 — C ABI of the B200-native ResnetVQAModel training step.
 *
 * The reference (shiv-vignesh/T5-Resnet-VQA) is pure Python/PyTorch and has no FFI layer; its hot path is
 * the nn.Module surface of `ResnetVQAModel` (model/resnet_vqa_model.py:28-165) driven by
 * `train_one_step` (trainer/faster_rcnn_vqa_trainer.py:391-406).  This header is the boundary a
 * maintainer binds with ctypes: plain pointers and sizes, no torch types.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; `vqa_last_error()` (thread-local)
 *     describes the failure.  Nothing throws, nothing calls exit().
 *   - all pointers are DEVICE pointers owned by the caller unless stated otherwise; the library keeps
 *     no reference after the call returns, except inside a `vqa_step*` handle, which borrows the
 *     parameter / gradient / workspace buffers it was bound to until `vqa_step_destroy`.
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous on that stream; no call
 *     synchronises the device.
 *   - bf16 tensors are row-major with the innermost dimension contiguous; activations are NHWC /
 *     [tokens, features].
 */
#ifndef VQA_B200_H_
#define VQA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* vqa_last_error(void);
int vqa_version(void);
/* bring-up aid: override the MN-major UMMA descriptor strides (bytes); zeros restore defaults */
int vqa_debug_set_umma(int a_lbo, int a_sbo, int b_lbo, int b_sbo);

/* ------------------------------------------------------------------------------------------------
 * tcgen05 GEMM.  out[M,N] = epilogue(alpha * op(A) op(B)^T)
 *   a_mn = 0: A is [M,K] row-major (lda);  a_mn = 1: A is stored [K,M] row-major (lda)
 *   b_mn = 0: B is [N,K] row-major (ldb);  b_mn = 1: B is stored [K,N] row-major (ldb)
 * Epilogue order: +bias[N] -> ReLU -> keep where relu_mask>0 -> dropout(p, rng, sid) -> +residual.
 * Replaces nn.Linear forward / dgrad / wgrad (model/multi_head_vision_text_attn.py:31-34,92-93,
 * model/resnet_vqa_model.py:86-88, hf T5 q/k/v/o/wi/wo).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int M, N, K;
  const void* A; long long lda; int a_mn;
  const void* B; long long ldb; int b_mn;
  void* out; long long ldo; int out_fp32;
  const float* bias;
  int relu;
  const void* relu_mask; long long ldm;          /* bf16 [M, ldm] */
  float drop_p; uint32_t drop_sid; const uint64_t* rng;   /* rng: device {seed, offset} */
  const void* residual; long long ldr; int res_fp32;
  float alpha;
  int bn;        /* output tile width: 64, 128 or 256 */
  int split_k;   /* >1: fp32 out must be zeroed by the caller; partial sums are red.add'ed */
} vqa_gemm_args;
int vqa_gemm_bf16(const vqa_gemm_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution, NHWC bf16, folded-BN bias + residual + ReLU epilogue.
 *   x [N,H,W,Cin] (stem7: [N,H,W+8,8], 3 zero pixels left / 5 right), w [Cout, R*S*Cin]
 *   (stem7: [Cout, 7*8*8]), out [N,Ho,Wo,Cout] bf16 (out_fp32=0) or fp32.
 * Replaces torchvision Conv2d+BatchNorm2d(eval)+ReLU(+identity) (tv resnet.py:89-105,143-163,197-200)
 * and the ConvTranspose2d channel projection (model/resnet_vqa_model.py:64-78,124,135).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int N, H, W, Cin, Cout, R, S, stride, pad, Ho, Wo, stem7;
  const void* x; const void* w; void* out; int out_fp32;
  const float* bias; const void* residual; int relu;
  int bn;
} vqa_conv_args;
int vqa_conv2d_bf16(const vqa_conv_args* a, void* stream);

/* Weight gradient of a stride-1 same-padded RxS convolution: dw[Cout, R*S*Cin] fp32 (zeroed by the
 * caller when split_k > 1) = sum_pixels dy[pix,Cout] * x[pix+tap,Cin].  (ConvTranspose2d wgrad.) */
typedef struct {
  int N, H, W, Cin, Cout, R, S, pad;
  const void* dy; const void* x; float* dw;
  int bn, split_k;
} vqa_conv_wgrad_args;
int vqa_conv2d_wgrad_bf16(const vqa_conv_wgrad_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* VQA_B200_H_ */
