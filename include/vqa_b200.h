/* vqa_b200 — C ABI of the B200-native ResnetVQAModel training step.
 *
 * The reference (shiv-vignesh/T5-Resnet-VQA) is pure Python/PyTorch and has no FFI layer; its hot path is
 * the nn.Module surface of `ResnetVQAModel` (model/resnet_vqa_model.py:28-165) driven by
 * `train_one_step` (trainer/faster_rcnn_vqa_trainer.py:391-406).  This header is the boundary a
 * maintainer binds with ctypes: plain pointers and sizes, no torch types.
 * Each entry point names the reference call site (file:line) whose arithmetic it replaces; `tv:` is
 * torchvision/models/resnet.py and `hf:` is transformers/models/t5/modeling_t5.py, the two third-party
 * files the reference's model is assembled from.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; `vqa_last_error()` (thread-local)
 *     describes the failure.  Nothing throws, nothing calls exit().
 *   - all data pointers are DEVICE pointers owned by the caller (PyTorch); the library allocates no
 *     device memory and keeps no reference after the call returns, except inside a plan (below),
 *     which borrows every pointer recorded into it until `vqa_plan_destroy`.
 *   - `plan`: NULL runs the op immediately on `stream`; otherwise the fully-resolved launch (tensor
 *     maps, pointers, shapes) is appended to the plan and nothing runs until `vqa_plan_run`.
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous on it; no call
 *     synchronises the device.  Safe to call from the autograd engine's worker thread.
 *   - bf16 tensors are row-major with the innermost dimension contiguous; activations are NHWC /
 *     [tokens, features].
 *   - dropout: counter-based Philox4x32-7 keyed by the device pair rng = {seed, offset}; an element
 *     of stream `sid` at flat index i is dropped iff u16(philox(seed, offset, sid, i / 8), i % 8) <
 *     p * 65536, survivors are scaled by 1/(1-p).  Forward and backward regenerate the same mask.
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
/* bring-up aid: GEMM / conv launches issued after this call stamp clock64() of CTA 0's phases into buf (device
 * memory, 11 warps x 16 slots); NULL switches it off.  Immediate-mode launches only. */
int vqa_debug_gemm_timing(long long* buf);

/* ---- launch plans ------------------------------------------------------------------------------ */
void* vqa_plan_create(void);
int vqa_plan_destroy(void* plan);
int vqa_plan_size(void* plan);                       /* number of recorded launches */
int vqa_plan_run(void* plan, void* stream);          /* replay (through the CUDA graph if captured) */
int vqa_plan_capture_graph(void* plan, void* stream);/* capture the recorded launches into a CUDA graph */
/* Two-lane plans: launches recorded after vqa_plan_set_lane(plan, 1) replay on a plan-owned side stream and
 * overlap lane 0 (parallel graph branches).  vqa_plan_fork: lane 1 waits for everything recorded on lane 0 so
 * far; vqa_plan_join: lane 0 waits for lane 1.  A plan is always joined at its end.  The caller must keep the
 * two lanes' buffers disjoint between a fork and the next join. */
int vqa_plan_set_lane(void* plan, int lane);
int vqa_plan_fork(void* plan);
int vqa_plan_join(void* plan);
/* Finer ordering than a full join: vqa_plan_mark returns an id (>= 0) for lane 1's current position;
 * vqa_plan_wait makes lane 0 wait until lane 1 has passed that position (used before lane 0 overwrites a
 * buffer an earlier lane-1 launch reads). */
int vqa_plan_mark(void* plan);
int vqa_plan_wait(void* plan, int mark);
/* measurement aid: eager replay with a CUDA event between launches; ms_out[vqa_plan_size] device durations
 * (synchronises the stream).  op_info: kernel family and the algorithmic flops / HBM bytes of launch i. */
int vqa_plan_profile(void* plan, void* stream, float* ms_out, int spin_us);
/* Back-to-back device time of the launches whose op name is in `names` ("gemm,conv,..."): `reps` passes between one
 * pair of events, no per-launch gaps (bench.py's roofline numerator / denominator).  Ignores data dependencies. */
int vqa_plan_time_ops(void* plan, void* stream, const char* names, int reps, float* ms_per_rep,
                      double* flops_per_rep, int* launches_per_rep);
int vqa_plan_op_info(void* plan, int i, const char** name, double* flops, double* bytes);

/* ------------------------------------------------------------------------------------------------
 * tcgen05 GEMM.  out[M,N] = epilogue(alpha * op(A) op(B)^T)
 *   a_mn = 0: A is [M,K] row-major (lda);  a_mn = 1: A is stored [K,M] row-major (lda)
 *   b_mn = 0: B is [N,K] row-major (ldb);  b_mn = 1: B is stored [K,N] row-major (ldb)
 * Epilogue order: *alpha -> +bias[N] -> (+residual if res_first) -> ReLU -> keep where relu_mask>0
 *                 -> dropout(p, rng, sid) -> (+residual if !res_first).
 * Replaces nn.Linear forward / dgrad / wgrad (model/multi_head_vision_text_attn.py:31-34,92-93,
 * model/resnet_vqa_model.py:86-88, hf:92-103 wi/wo, hf:178-181 q/k/v/o).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int M, N, K;
  const void* A; long long lda; int a_mn;
  const void* B; long long ldb; int b_mn;
  void* out; long long ldo; int out_fp32;
  const float* bias;
  int relu;                                      /* 1: ReLU; 2: exact (erf) GELU instead - bf16 output, no residual / mask /
                                                    dropout / accumulate (the frozen ViT's fc1, vit: ViTIntermediate) */
  const void* relu_mask; long long ldm;          /* bf16 [M, ldm] */
  float drop_p; uint32_t drop_sid; const uint64_t* rng;   /* rng: device {seed, offset} */
  const void* residual; long long ldr; int res_fp32; int res_first;
  float alpha;
  int accumulate; /* 1: fp32 out += result (red.add) instead of a plain store */
  int bn;        /* output tile width: 64, 128 or 256 */
  int split_k;   /* >1: fp32 out must be zeroed by the caller; partial sums are red.add'ed */
  int cta_pair;  /* 1: CTA pairs (cta_group::2): 256-row MMA, B tile split across two SMs; bn must be 128 or 256 */
  int ksplit;    /* >1: cluster split-K: ksplit CTAs (one thread-block cluster) per output tile, each contracting a
                    k-slice; partial sums meet in ks_ws and every epilogue (incl. ReLU / dropout / bf16 output) still
                    applies.  Needs tiles x ksplit <= 148 and K >= 128 * ksplit; excludes split_k, accumulate, cta_pair */
  void* ks_ws;   /* device workspace, >= vqa_gemm_ksplit_workspace(M, N, bn, ksplit) bytes, 16-byte aligned; must not be
                    shared by launches that can run concurrently (other streams / plan lanes) */
  long long ks_ws_bytes;
  /* Two-term operand split (a forward GEMM whose bf16 rounding must not be seen: the early T5 blocks, DESIGN.md
     "Parity").  B_lo: bf16 [N,K] (ldb) holding bf16(W - bf16(W)) of the fp32 weight whose bf16(W) is B; the launch
     then contracts A*B^T + A*B_lo^T in one k-loop (fp32 accumulation in TMEM).  a_lo_col > 0: A carries its own
     low-order half in columns [a_lo_col, a_lo_col + K) of the same rows (lda >= a_lo_col + K) and the term
     A_lo*B^T is added as well.  Needs K % 64 == 0, K-major operands, no split_k / accumulate / ksplit and no bf16
     residual.  NULL / 0: plain bf16 GEMM. */
  const void* B_lo;
  long long a_lo_col;
  int max_ctas;  /* > 0: the persistent grid uses at most this many CTAs (SMs).  A launch on a plan's side lane (weight
                    gradients) can leave the rest of the GPU to the chain on the main lane.  0: all SMs */
} vqa_gemm_args;
long long vqa_gemm_ksplit_workspace(int M, int N, int bn, int ksplit);
int vqa_gemm_bf16(void* plan, const vqa_gemm_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution, NHWC bf16, folded-BN bias + residual + ReLU epilogue.
 *   x [N,H,W,Cin] (stem7: [N,H,W+8,8], 3 zero pixels left / 5 right), w [Cout, R*S*Cin]
 *   (stem7: [Cout, 7*8*8]), out [N,Ho,Wo,Cout] bf16 (out_fp32=0) or fp32.
 * Replaces torchvision Conv2d+BatchNorm2d(eval)+ReLU(+identity) (tv:89-105,143-163,197-200)
 * and the ConvTranspose2d channel projection (model/resnet_vqa_model.py:64-78,124,135).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int N, H, W, Cin, Cout, R, S, stride, pad, Ho, Wo, stem7;
  const void* x; const void* w; void* out; int out_fp32;
  const float* bias; const void* residual; int relu;
  int bn;
  int cta_pair;
  int ksplit; void* ks_ws; long long ks_ws_bytes;   /* cluster split-K as in vqa_gemm_args (M = N*Ho*Wo, N = Cout) */
} vqa_conv_args;
int vqa_conv2d_bf16(void* plan, const vqa_conv_args* a, void* stream);

/* Weight gradient of a stride-1 same-padded RxS convolution: dw[Cout, R*S*Cin] fp32 (zeroed by the
 * caller when split_k > 1) = sum_pixels dy[pix,Cout] * x[pix+tap,Cin].  (ConvTranspose2d wgrad,
 * autograd of model/resnet_vqa_model.py:124,135.) */
typedef struct {
  int N, H, W, Cin, Cout, R, S, pad;
  const void* dy; const void* x; float* dw;
  int bn, split_k;
  int cta_pair;
} vqa_conv_wgrad_args;
int vqa_conv2d_wgrad_bf16(void* plan, const vqa_conv_wgrad_args* a, void* stream);

/* ---- layout / precision plumbing (weight preparation, input formatting) ------------------------- */
int vqa_cast_f32_bf16(void* plan, const float* src, void* dst, long long n, void* stream);
int vqa_cast_bf16_f32(void* plan, const void* src, float* dst, long long n, void* stream);
/* dst[i] = bf16(src[i] - float(bf16(src[i]))): the low-order half of the two-term split of an fp32 weight range */
int vqa_split_lo_bf16(void* plan, const float* src, void* dst, long long n, void* stream);
int vqa_memset_zero(void* plan, void* ptr, long long bytes, void* stream);
/* device-to-device copy as a plan step (the frozen backbone's last feature map is copied out of the backbone's own buffers at
 * the start of the projection plan, so that the NEXT step's backbone may overwrite them while this step's backward still reads
 * the copy: engine.forward, "early backbone") */
int vqa_memcpy_d2d(void* plan, void* dst, const void* src, long long bytes, void* stream);
int vqa_axpy_f32(void* plan, float* y, const float* x, float a, long long n, void* stream); /* y += a*x */
/* Conv2d weight [O,I,R,S] fp32 (+ eval BatchNorm gamma/beta/mean/var, may be NULL) -> bf16
 * [O, R, Sp, Ip] (zero padded) scaled by gamma/sqrt(var+eps), and bias[O] = beta - mean*scale.
 * (tv:197-199 conv1/bn1 and every block's conv/bn pair; BN in eval mode, model/resnet_vqa_model.py:116,127) */
int vqa_fold_conv_bn(void* plan, const float* w, const float* gamma, const float* beta,
                     const float* mean, const float* var, float eps, void* w_out, float* bias_out,
                     int O, int I, int R, int S, int Sp, int Ip, void* stream);
/* ConvTranspose2d weight [Cin,Cout,3,3] fp32 -> equivalent Conv2d weight bf16 [Cout, 3,3, Cin] with the
 * taps flipped; and the inverse mapping of its fp32 gradient (model/resnet_vqa_model.py:64-78). */
int vqa_convT_weight_prep(void* plan, const float* w, void* w_out, int Cin, int Cout, void* stream);
/* same, from the bf16 shadow of the weight (identical result: the shadow is bf16_rn of the fp32 weight) */
int vqa_convT_weight_prep_bf16(void* plan, const void* w_bf16, void* w_out, int Cin, int Cout, void* stream);
int vqa_convT_wgrad_unprep(void* plan, const float* dw_conv, float* dw, int Cin, int Cout, void* stream);
/* images fp32 [N,3,H,W] -> bf16 [N,H,W+8,8] (stem layout); bf16 NHWC -> fp32 NCHW feature map */
int vqa_image_to_stem(void* plan, const float* img, void* out, int N, int H, int W, void* stream);
/* input edge: uint8 RGB [N,H,W,3] as cv2 produces it (dataset_utils/resnet_vqa_daquar_dataset.py:153-171)
 * -> the same stem layout with transforms.ToTensor()'s /255 folded in: a quarter of the fp32 CHW upload */
int vqa_image_u8_to_stem(void* plan, const uint8_t* img, void* out, int N, int H, int W, void* stream);
/* F.interpolate(mode="nearest") on bf16 NHWC [N,h,w,C] -> [N,H,W,C]: the FPN top-down pathway of the Faster R-CNN backbone
 * (torchvision feature_pyramid_network.py; model/faster_rcnn_vqa_model.py:51-53,150-154) */
int vqa_upsample_nearest_nhwc(void* plan, const void* x, void* out, int N, int h, int w, int H, int W, int C, void* stream);
int vqa_nhwc_to_nchw_f32(void* plan, const void* x, float* out, int N, int H, int W, int C, void* stream);
/* MaxPool2d(3, stride 2, pad 1) on bf16 NHWC (tv:200) */
int vqa_maxpool3x3s2(void* plan, const void* x, void* out, int N, int H, int W, int C, void* stream);

/* ---- T5 encoder pieces --------------------------------------------------------------------------- */
/* nn.Embedding gather + dropout (hf:682,734): out fp32 [M,D]; backward scatter-adds into dtable. */
int vqa_embedding_fwd(void* plan, const long long* ids, const float* table, float* out, int M, int D,
                      int vocab, float drop_p, uint32_t sid, const uint64_t* rng, void* stream);
int vqa_embedding_bwd(void* plan, const long long* ids, const float* dout, float* dtable, int M, int D,
                      int vocab, float drop_p, uint32_t sid, const uint64_t* rng, void* stream);
/* Data-parallel form of embedding_bwd: rows[t,:] = scale * dropout(dout[t,:]) leave the rank (all-gather with the ids), and
 * every rank scatters all ranks' rows into its zeroed dtable such that the result does not depend on the order of the
 * additions (ids carried by several tokens are summed as 64-bit fixed-point integers, unit 2^-40), so the replicas'
 * embedding gradients are bit-identical.  first_ws: int[2*vocab] scratch, acc_ws: int64[T*D] scratch (16-byte aligned).
 * (New work: the reference is single-device.) */
int vqa_embedding_bwd_rows(void* plan, const float* dout, float* rows, int M, int D, float drop_p, uint32_t sid,
                           const uint64_t* rng, float scale, void* stream);
int vqa_embedding_scatter_ordered(void* plan, const long long* ids, const float* rows, float* dtable, int* first_ws,
                                  long long* acc_ws, int T, int D, int vocab, void* stream);
/* T5LayerNorm (hf:55-68): y = w * x * rsqrt(mean(x^2) + eps), then optional dropout (hf:768).
 * y_bf16 and/or y_f32 may be NULL.  Backward: dx = (dres?) + d/dx, dw += sum_rows (atomic). */
int vqa_rmsnorm_fwd(void* plan, const float* x, const float* w, void* y_bf16, float* y_f32,
                    float* rstd, int M, int D, float eps, float drop_p, uint32_t sid,
                    const uint64_t* rng, void* stream);
/* Same normalisation written as a two-term bf16 split: y_hilo is bf16 [M, 2*D], columns [0,D) = bf16(y),
 * columns [D,2D) = bf16(y - bf16(y)) (the A operand of a vqa_gemm_bf16 launch with a_lo_col = D). */
int vqa_rmsnorm_fwd_split(void* plan, const float* x, const float* w, void* y_hilo, float* rstd, int M, int D,
                          float eps, void* stream);
/* g_out (bf16 [M,D], may be NULL): dropout-masked copy of dx under stream g_sid / probability g_drop_p, i.e.
 * the gradient entering the next residual branch, produced here instead of by a separate vqa_dropout_cast. */
int vqa_rmsnorm_bwd(void* plan, const void* dy, int dy_fp32, const float* x, const float* w,
                    const float* rstd, const float* dres, float* dx, float* dw, int M, int D,
                    float drop_p, uint32_t sid, const uint64_t* rng, void* g_out, float g_drop_p,
                    uint32_t g_sid, void* stream);
/* relative-position bias (hf:236-251): bias[h,i,j] = table[bucket[i*Lk+j], h]; gradient back to table */
int vqa_t5_bias_build(void* plan, const float* table, const int* bucket, float* bias, int H, int L,
                      int nbuckets, void* stream);
int vqa_t5_bias_grad(void* plan, const float* dbias, const int* bucket, float* dtable, int H, int L,
                     int nbuckets, void* stream);

/* ---- attention (hf:308-334 T5; model/multi_head_vision_text_attn.py:73-86 SGA) ------------------ */
typedef struct {
  int B, H, Lq, Lk, hd;                 /* hd = 64 (T5) or 96 (SGA) */
  const void* q; long long ldq;         /* bf16, element (b, i, h, d) at q[(b*Lq+i)*ldq + h*hd + d] */
  const void* k; long long ldk;
  const void* v; long long ldv;
  void* out; long long ldo;             /* bf16 context, same indexing */
  void* probs;                          /* fp32 [B,H,Lq,Lk] softmax output (pre-dropout), saved for backward; may be NULL */
  const float* bias;                    /* fp32 [H,Lq,Lk] additive or NULL */
  const long long* key_mask;            /* int64 [B,Lk], 0 = masked key, or NULL (hf:323-325) */
  float scale;                          /* 1.0 (T5) or 1/sqrt(hd) (SGA) */
  float drop_p; uint32_t sid; const uint64_t* rng;
  float* stats;                         /* fp32 [B,H,Lq,2] row max and 1/row-sum of the softmax.  Non-NULL with Lq, Lk <= 32
                                           selects the tcgen05 flash kernel (probs is then not written); NULL or longer
                                           sequences run the SIMT kernel, which saves probs instead */
} vqa_attn_fwd_args;
int vqa_attention_fwd(void* plan, const vqa_attn_fwd_args* a, void* stream);
typedef struct {
  int B, H, Lq, Lk, hd;
  const void* q; long long ldq; const void* k; long long ldk; const void* v; long long ldv;
  const void* probs; const void* dout; long long ldo;
  void* dq; long long lddq; void* dk; long long lddk; void* dv; long long lddv;   /* bf16 */
  float* dbias;                         /* fp32 [H,Lq,Lk] accumulated atomically, or NULL */
  float scale; float drop_p; uint32_t sid; const uint64_t* rng;
  const float* stats;                   /* forward's stats: backward recomputes the probabilities (probs unused) */
  const float* bias;                    /* forward's bias / key_mask, needed for that recomputation */
  const long long* key_mask;
} vqa_attn_bwd_args;
int vqa_attention_bwd(void* plan, const vqa_attn_bwd_args* a, void* stream);
/* bring-up aid: tcgen05 attention ops created after this call stamp clock64() of CTA 0's phases into buf
 * (device memory, 4 warps x 16 slots); NULL switches it off */
int vqa_debug_attn_timing(long long* buf);

/* ---- SGA pieces ---------------------------------------------------------------------------------- */
/* nn.LayerNorm(768) (model/multi_head_vision_text_attn.py:120-126) over z = x + dropout(sublayer),
 * z produced by the GEMM epilogue.  Backward: dz = d/dz, dgamma/dbeta accumulated atomically. */
int vqa_layernorm_fwd(void* plan, const float* z, const float* gamma, const float* beta, void* y_bf16,
                      float* y_f32, float* mean, float* rstd, int M, int D, float eps, void* stream);
/* g_out / g_colsum (may be NULL): dropout-masked bf16 copy of dz (stream g_sid) and its column sums += (the
 * bias gradient of the residual branch's last Linear). */
int vqa_layernorm_bwd(void* plan, const float* dy, const float* z, const float* gamma,
                      const float* mean, const float* rstd, float* dz, float* dgamma, float* dbeta,
                      int M, int D, void* g_out, float g_drop_p, uint32_t g_sid, const uint64_t* rng,
                      float* g_colsum, void* stream);
/* g = dropout_mask(x) as bf16 (gradient entering a dropped residual branch) */
int vqa_dropout_cast(void* plan, const float* x, void* out_bf16, long long rows, int N, float drop_p,
                     uint32_t sid, const uint64_t* rng, void* stream);
/* out[n] += sum_m x[m,n]  (bias gradients); x bf16 [M, ld] */
int vqa_colsum_bf16(void* plan, const void* x, long long ld, float* out, int M, int N, void* stream);

/* ---- head: AttentionPooler + classifier + loss (model/resnet_vqa_model.py:14-26,152-160) -------- */
int vqa_pooler_fwd(void* plan, const float* x, const float* a, const float* b, float* w_out,
                   float* pooled_f32, void* pooled_bf16, int B, int L, int D, void* stream);
int vqa_pooler_bwd(void* plan, const float* x, const float* a, const float* w, const float* dpooled,
                   float* dx, float* da, float* db, int B, int L, int D, void* stream);
/* log_softmax + NLLLoss(mean): logp [B,A]; loss += -logp[b,label]/B (loss zeroed by the caller) */
int vqa_logsoftmax_nll_fwd(void* plan, const float* logits, long long ld, const long long* labels,
                           float* logp, float* loss, int B, int A, void* stream);
/* dlogits[b,:] = gscale * (exp(logp) - onehot)/B + glogp terms; gloss: device scalar or NULL (=1),
 * glogp: fp32 [B,A] upstream gradient of the log-probs or NULL.  dlogits bf16 [B, ld] (pad cols zeroed) */
int vqa_logsoftmax_nll_bwd(void* plan, const float* logp, const long long* labels, const float* gloss,
                           const float* glogp, void* dlogits, long long ld, int B, int A, void* stream);

/* ---- VitVQAModel step (model/vit_vqa_model.py:127-227).  `vit:` = transformers/models/vit/modeling_vit.py -- */
/* softmax(scale * Q K^T) V for ONE sequence length L <= 256 on both sides, hd = 64, no mask / bias / dropout, nothing saved:
 * the frozen ViT-B/16's self-attention over 197 tokens (vit: ViTSelfAttention; the reference runs it under torch.no_grad(),
 * model/vit_vqa_model.py:184-186; probs != NULL also writes the softmax, output_attentions=True of :240-243).  tcgen05 kernel, one CTA per (batch, head, 128-query tile).  Indexing as vqa_attn_fwd_args. */
int vqa_attention_long_fwd(void* plan, const void* q, long long ldq, const void* k, long long ldk, const void* v,
                           long long ldv, void* out, long long ldo, int B, int H, int L, int hd, float scale,
                           float* probs /* optional fp32 [B,H,L,L]: the attention maps generate_answers returns */,
                           void* stream);
/* pixel_values fp32 [B,3,H,W] -> bf16 [B*(H/P)*(W/P), 3*P*P] patch rows (column = c*P*P + ky*P + kx): the A operand of the
 * patch-projection GEMM, Conv2d(3,768,P,stride P) as a matrix product (vit: ViTPatchEmbeddings) */
int vqa_vit_patchify(void* plan, const float* img, void* out, int B, int H, int W, int P, void* stream);
/* hidden fp32 [B, NP+1, D]: token 0 = cls + pos[0], token 1+p = patch[b,p,:] + pos[1+p] (vit: ViTEmbeddings.forward) */
int vqa_vit_assemble(void* plan, const float* patch, const float* cls, const float* pos, float* hidden, int B, int NP,
                     int D, void* stream);
/* exact (erf) GELU in place on n bf16 values (vit: ViTIntermediate, hidden_act = "gelu") */
int vqa_gelu_bf16(void* plan, void* x, long long n, void* stream);
/* out bf16 [B, 2D] = [tanh(pooled_pre[b,:]) | enc[b*L, :]]: ViTPooler's activation and torch.cat([pooler_output, encoder
 * token 0]) (model/vit_vqa_model.py:192-198); pooled_out (fp32 [B,D], may be NULL) receives the tanh half */
int vqa_vit_fuse_concat(void* plan, const float* pooled_pre, const float* enc, int L, void* out, float* pooled_out, int B,
                        int D, void* stream);
/* T5 decoder cross-attention onto ONE encoder token (encoder_hidden_states = fused_embedding.unsqueeze(1),
 * model/vit_vqa_model.py:207-212): the softmax over a single key is 1, so ctx[b*Lq+q, h*hd+d] = keep(b,h,q) * v[b, h*hd+d],
 * keep = the dropout HF applies to the attention weights [B,H,Lq,1] (hf:327-334); backward sums dctx over q into dv.
 * The query / key projections receive exactly zero gradient. */
int vqa_xattn1_fwd(void* plan, const void* v, void* ctx, int B, int H, int Lq, int hd, float drop_p, uint32_t sid,
                   const uint64_t* rng, void* stream);
int vqa_xattn1_bwd(void* plan, const void* dctx, void* dv, int B, int H, int Lq, int hd, float drop_p, uint32_t sid,
                   const uint64_t* rng, void* stream);
/* out[b,:] = src[b*L + last(b), :], last(b) = the last position with mask[b,j] == 1 (0 if none or mask == NULL): the answer
 * token gather (model/vit_vqa_model.py:215-219) and, with mask == NULL, encoder_outputs[:,0,:] (:192).  scatter_rows is the
 * backward: dst fp32 [B*L, D] = 0 except row last(b) = src[b,:]. */
int vqa_gather_rows(void* plan, const float* src, const long long* mask, void* out_bf16, float* out_f32, int B, int L,
                    int D, void* stream);
int vqa_scatter_rows(void* plan, const float* src, const long long* mask, float* dst, int B, int L, int D, void* stream);
/* decoder self-attention: bias[h,i,j] = finfo.min for j > i on top of vqa_t5_bias_build's table (HF causal mask) */
int vqa_t5_bias_causal(void* plan, float* bias, int H, int L, void* stream);
/* gradient through Dropout(ReLU(.)) from the layer's OUTPUT y: out bf16 = y > 0 ? dy * scale : 0 (fusing_layer,
 * model/vit_vqa_model.py:150-154; scale = 1/(1-p) in training, 1 in eval) */
int vqa_relu_dropout_bwd(void* plan, const float* dy, const void* y, void* out, float scale, long long n, void* stream);

/* ---- optimizer step (trainer/faster_rcnn_vqa_trainer.py:399-404) -------------------------------- */
/* out[0] += sum x^2 */
int vqa_sumsq_f32(void* plan, const float* x, long long n, float* out, void* stream);
/* torch.nn.utils.clip_grad_norm_'s scaling (trainer/faster_rcnn_vqa_trainer.py:399-400) over a contiguous fp32
 * gradient range: g *= max_norm / (sqrt(gnorm_sq) + 1e-6) when that factor is < 1, untouched otherwise. */
int vqa_clip_scale_f32(void* plan, float* g, long long n, const float* gnorm_sq, float max_norm, void* stream);
/* torch.optim.AdamW(amsgrad) on a contiguous fp32 range (torch/optim/adamw.py single-tensor math:
 * p *= 1-lr*wd; m = lerp(m,g,1-b1); v = b2*v + (1-b2)*g*g; vmax = max(vmax,v);
 * p -= (lr/bc1) * m / (sqrt(vmax)/sqrt(bc2) + eps)); scalar combinations are formed in double on the host.
 * gnorm_sq: device scalar holding the global sum of squared gradients (NULL = no clipping); grads are
 * scaled by min(1, max_norm / (sqrt(gnorm_sq) + 1e-6)) like clip_grad_norm_.  shadow: bf16 copy of the
 * updated parameters (may be NULL). */
int vqa_adamw_amsgrad(void* plan, float* p, const float* g, float* m, float* v, float* vmax,
                      void* shadow, long long n, double lr, double beta1, double beta2, double eps,
                      double weight_decay, double bias_correction1, double bias_correction2,
                      const float* gnorm_sq, float max_norm, int amsgrad, void* stream);
/* rng[1] += 1 (new dropout masks for the next step) */
int vqa_rng_advance(void* plan, uint64_t* rng, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* VQA_B200_H_ */
