// extern "C" surface declared in include/vqa_b200.h.
#include "../../include/vqa_b200.h"

#include "common.cuh"
#include "gemm_ops.cuh"
#include "tmap.cuh"

using namespace vqa;

extern "C" {

const char* vqa_last_error(void) { return get_last_error(); }
int vqa_version(void) { return 101; }
int vqa_debug_set_umma(int a_lbo, int a_sbo, int b_lbo, int b_sbo) {
  gemm_debug_set_umma(a_lbo, a_sbo, b_lbo, b_sbo);
  return 0;
}

int vqa_debug_gemm_timing(long long* buf) {
  gemm_debug_set_clock(buf);
  return 0;
}

long long vqa_gemm_ksplit_workspace(int M, int N, int bn, int ksplit) {
  if (ksplit <= 1 || bn <= 0) return 0;
  const int tiles = ((M + 127) / 128) * ((N + bn - 1) / bn);
  return static_cast<long long>(gemm_ksplit_ws_bytes(tiles, ksplit, bn));
}

int vqa_gemm_bf16(void* plan, const vqa_gemm_args* a, void* stream) {
  Epilogue e;
  e.bias = a->bias; e.relu = a->relu;
  e.relu_mask = reinterpret_cast<const __nv_bfloat16*>(a->relu_mask); e.ldm = a->ldm;
  e.drop_p = a->drop_p; e.drop_sid = a->drop_sid;
  e.rng = reinterpret_cast<const unsigned long long*>(a->rng);
  e.residual = a->residual; e.ldr = a->ldr; e.res_fp32 = a->res_fp32; e.res_first = a->res_first;
  e.alpha = a->alpha; e.accumulate = a->accumulate;
  e.ksplit = a->ksplit; e.ks_ws = static_cast<float*>(a->ks_ws); e.ks_ws_bytes = static_cast<size_t>(a->ks_ws_bytes);
  e.b_lo = a->B_lo; e.a_lo_col = a->a_lo_col; e.max_ctas = a->max_ctas;
  GemmOp op;
  int r = gemm_op_init(&op, a->M, a->N, a->K, a->A, a->lda, a->a_mn, a->B, a->ldb, a->b_mn, a->out,
                       a->ldo, a->out_fp32, e, a->bn, a->split_k, a->cta_pair ? 2 : 1);
  if (r) return r;
  note_op("gemm", 2.0 * a->M * a->N * a->K, 0.0);
  return submit(plan, stream, [op](cudaStream_t s) { return gemm_op_run(&op, s); });
}

int vqa_conv2d_bf16(void* plan, const vqa_conv_args* a, void* stream) {
  ConvGeom g;
  g.Nimg = a->N; g.H = a->H; g.W = a->W; g.Cin = a->Cin; g.Cout = a->Cout; g.R = a->R; g.S = a->S;
  g.stride = a->stride; g.pad = a->pad; g.Ho = a->Ho; g.Wo = a->Wo; g.stem7 = a->stem7;
  Epilogue e;
  e.bias = a->bias; e.relu = a->relu; e.residual = a->residual; e.ldr = a->Cout; e.res_fp32 = 0; e.res_first = 1;
  e.ksplit = a->ksplit; e.ks_ws = static_cast<float*>(a->ks_ws); e.ks_ws_bytes = static_cast<size_t>(a->ks_ws_bytes);
  GemmOp op;
  int r = conv_op_init(&op, g, a->x, a->w, a->out, a->out_fp32, e, a->bn, a->cta_pair ? 2 : 1);
  if (r) return r;
  // algorithmic MACs of the convolution (the stem's zero-padded taps/channels are not counted)
  note_op("conv", 2.0 * a->N * a->Ho * a->Wo * a->Cout * a->R * a->S * (a->stem7 ? 3 : a->Cin), 0.0);
  return submit(plan, stream, [op](cudaStream_t s) { return gemm_op_run(&op, s); });
}

int vqa_conv2d_wgrad_bf16(void* plan, const vqa_conv_wgrad_args* a, void* stream) {
  ConvGeom g;
  g.Nimg = a->N; g.H = a->H; g.W = a->W; g.Cin = a->Cin; g.Cout = a->Cout; g.R = a->R; g.S = a->S;
  g.stride = 1; g.pad = a->pad; g.Ho = a->H; g.Wo = a->W; g.stem7 = 0;
  GemmOp op;
  int r = conv_wgrad_op_init(&op, g, a->dy, a->x, a->dw, a->bn, a->split_k, a->cta_pair ? 2 : 1);
  if (r) return r;
  note_op("conv_wgrad", 2.0 * a->N * a->H * a->W * a->Cout * a->R * a->S * a->Cin, 0.0);
  return submit(plan, stream, [op](cudaStream_t s) { return gemm_op_run(&op, s); });
}

}  // extern "C"
