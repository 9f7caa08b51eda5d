// Host-side builders for the GEMM / implicit-GEMM launches: encode the tensor maps once, launch many times.
#pragma once
#include "gemm.cuh"

namespace vqa {

struct Epilogue {
  const float* bias = nullptr;
  int relu = 0;
  const __nv_bfloat16* relu_mask = nullptr;
  long long ldm = 0;
  float drop_p = 0.f;
  uint32_t drop_sid = 0;
  const unsigned long long* rng = nullptr;
  const void* residual = nullptr;
  long long ldr = 0;
  int res_fp32 = 0;
  int res_first = 0;
  float alpha = 1.f;
  int accumulate = 0;  // fp32 out += result (red.add) instead of store
  int ksplit = 1;      // > 1: cluster split-K over `ksplit` CTAs per output tile (needs ks_ws; see gemm.cuh)
  float* ks_ws = nullptr;
  size_t ks_ws_bytes = 0;
  const void* b_lo = nullptr;   // two-term operand split (see vqa_gemm_args.B_lo / a_lo_col)
  long long a_lo_col = 0;
  int max_ctas = 0;             // > 0: cap of the persistent grid (vqa_gemm_args.max_ctas)
};

struct GemmOp {
  CUtensorMap tmA, tmB, tmOut, tmRes;
  GemmParams p;
  int bn = 128;
  int split_k = 1;
  int ctas = 1;        // 2: CTA pairs (cta_group::2)
  bool valid = false;
};

// D[M,N] = epi( op(A) * op(B)^T ).
//   a_mn = 0: A is row-major [M, K] (ld = lda);   a_mn = 1: A is stored row-major [K, M] (ld = lda)
//   b_mn = 0: B is row-major [N, K] (ld = ldb);   b_mn = 1: B is stored row-major [K, N] (ld = ldb)
// out is bf16 or fp32 with row stride ldo.  split_k > 1 requires fp32 out that was zeroed beforehand.
int gemm_op_init(GemmOp* op, int M, int N, int K, const void* A, long long lda, int a_mn,
                 const void* B, long long ldb, int b_mn, void* out, long long ldo, int out_fp32,
                 const Epilogue& epi, int bn, int split_k, int ctas = 1);

struct ConvGeom {
  int Nimg, H, W, Cin;     // input NHWC (bf16); for stem7 the input is the padded 8-channel layout
  int Cout, R, S, stride, pad;
  int Ho, Wo;
  int stem7;               // 1: 7x7/2 stem over [N, H, W+8, 8] input (3 zero pixels left, 5 right)
};
// out[N,Ho,Wo,Cout] = epi( conv(x, w) ), w is bf16 [Cout, R*S*Cin] (tap-major, channel-minor);
// for stem7 w is [Cout, 7*8*8].  residual (bf16 NHWC, same shape as out) and bias go through epi.
int conv_op_init(GemmOp* op, const ConvGeom& g, const void* x, const void* w, void* out, int out_fp32,
                 const Epilogue& epi, int bn, int ctas = 1);

// Weight gradient of a stride-1 "same" RxS convolution: dW[Cout, R*S*Cin] (fp32, ld = R*S*Cin) =
// sum over pixels dY[pix, Cout] * X[pix + tap, Cin].  dy: bf16 [N,Ho,Wo,Cout]; x: bf16 [N,H,W,Cin].
int conv_wgrad_op_init(GemmOp* op, const ConvGeom& g, const void* dy, const void* x, float* dw,
                       int bn, int split_k, int ctas = 1);

int gemm_op_run(const GemmOp* op, cudaStream_t stream);

// bring-up only: override MN-major descriptor strides (bytes); zeros restore the defaults
void gemm_debug_set_umma(int a_lbo, int a_sbo, int b_lbo, int b_sbo);
// bring-up only: GEMM launches issued after this call stamp clock64() of CTA 0's phases into buf (11 warps x 16)
void gemm_debug_set_clock(long long* buf);

}  // namespace vqa
