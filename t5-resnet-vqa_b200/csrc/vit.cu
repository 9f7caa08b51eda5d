// Small kernels of the VitVQAModel step (model/vit_vqa_model.py:127-227): the frozen ViT-B/16's input
// and activation plumbing (`vit:` = transformers/models/vit/modeling_vit.py), the fusing layer's concat / backward mask,
// the T5 decoder's one-token cross-attention (softmax over ONE key is identically 1, so the attention output is the value
// row of the sample, times the dropout mask HF draws on the attention weights, hf:327-334), the causal part of the
// decoder's position bias, and the gather of the last un-padded decoder position.  All are HBM / latency bound:
// 128-bit accesses, one pass.  The ViT's 197-token attention itself is a tcgen05 kernel (attention_tc.cu).
#include "../../include/vqa_b200.h"
#include "common.cuh"

using namespace vqa;

namespace {

inline int grid_for(long long items, int threads, int max_per_sm = 8) {
  long long b = (items + threads - 1) / threads;
  const long long cap = 148LL * max_per_sm;
  if (b > cap) b = cap;
  return b < 1 ? 1 : static_cast<int>(b);
}

// images fp32 [B,3,H,W] -> bf16 patches [B*(H/P)*(W/P), 3*P*P], column = c*P*P + ky*P + kx: the A operand of the patch
// projection GEMM against Conv2d(3, 768, P, stride P).weight viewed as [768, 3*P*P] (vit: ViTPatchEmbeddings)
__global__ void vit_patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int H, int W,
                                    int P) {
  pdl_grid_sync();
  const int w8 = W >> 3, npw = W / P, nph = H / P;
  const long long total = static_cast<long long>(B) * 3 * H * w8;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int x8 = static_cast<int>(i % w8);
    long long r = i / w8;
    const int y = static_cast<int>(r % H); r /= H;
    const int c = static_cast<int>(r % 3);
    const int b = static_cast<int>(r / 3);
    float f[8];
    load_f32x8(img + ((static_cast<long long>(b) * 3 + c) * H + y) * W + x8 * 8, f);
    const int x = x8 * 8, px = x / P, kx = x - px * P, py = y / P, ky = y - py * P;
    const long long row = (static_cast<long long>(b) * nph + py) * npw + px;
    store_bf16x8(out + row * (3 * P * P) + (c * P + ky) * P + kx, f);
  }
}

// hidden[b, 0, :] = cls + pos[0]; hidden[b, 1 + p, :] = patch[b, p, :] + pos[1 + p]   (vit: ViTEmbeddings.forward)
__global__ void vit_assemble_kernel(const float* __restrict__ patch, const float* __restrict__ cls,
                                    const float* __restrict__ pos, float* __restrict__ hidden, int B, int NP, int D) {
  pdl_grid_sync();
  const int d4 = D >> 2, T = NP + 1;
  const long long total = static_cast<long long>(B) * T * d4;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = static_cast<int>(i % d4);
    const long long r = i / d4;
    const int tkn = static_cast<int>(r % T);
    const int b = static_cast<int>(r / T);
    const float4 pe = reinterpret_cast<const float4*>(pos)[static_cast<long long>(tkn) * d4 + c];
    const float4 v = tkn == 0 ? reinterpret_cast<const float4*>(cls)[c]
                              : reinterpret_cast<const float4*>(patch)[(static_cast<long long>(b) * NP + tkn - 1) * d4 + c];
    reinterpret_cast<float4*>(hidden)[i] = make_float4(v.x + pe.x, v.y + pe.y, v.z + pe.z, v.w + pe.w);
  }
}

// exact GELU in place on bf16 (vit: hidden_act "gelu" = 0.5 x (1 + erf(x / sqrt 2)))
__global__ void gelu_bf16_kernel(__nv_bfloat16* __restrict__ x, long long n8) {
  pdl_grid_sync();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float f[8];
    load_bf16x8(x + i * 8, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = 0.5f * f[k] * (1.f + erff(f[k] * 0.70710678118654752f));
    store_bf16x8(x + i * 8, f);
  }
}

// out[b, 0:D] = tanh(pooled_pre[b, :]) (vit: ViTPooler); out[b, D:2D] = enc[b*L + 0, :]   (model/vit_vqa_model.py:192-198)
__global__ void vit_fuse_concat_kernel(const float* __restrict__ pooled_pre, const float* __restrict__ enc, int L,
                                       __nv_bfloat16* __restrict__ out, float* __restrict__ pooled_out, int B, int D) {
  pdl_grid_sync();
  const int d8 = D >> 3;
  const int total = B * 2 * d8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % (2 * d8), b = i / (2 * d8);
    float f[8];
    if (c < d8) {
      load_f32x8(pooled_pre + static_cast<long long>(b) * D + c * 8, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = tanhf(f[k]);
      if (pooled_out != nullptr) store_f32x8(pooled_out + static_cast<long long>(b) * D + c * 8, f);
    } else {
      load_f32x8(enc + static_cast<long long>(b) * L * D + (c - d8) * 8, f);
    }
    store_bf16x8(out + static_cast<long long>(b) * 2 * D + c * 8, f);
  }
}

__device__ __forceinline__ float keep_mult(const DropCtx& dc, unsigned long long flat) {
  if (!dc.on) return 1.f;
  const Philox8 r = philox8(dc.seed, dc.offset, dc.sid, flat >> 3);
  return r.u16(static_cast<int>(flat & 7)) < dc.thresh ? 0.f : dc.scale;
}

// One-key cross-attention, forward: ctx[b*Lq + q, h*hd + d] = keep(b, h, q) * v[b, h*hd + d]; the dropout element index is
// that of the attention weight [B, H, Lq, 1] (same Philox convention as the attention kernels)
__global__ void xattn1_fwd_kernel(const __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ ctx, int B, int H, int Lq,
                                  int hd, float drop_p, uint32_t sid, const unsigned long long* __restrict__ rng) {
  pdl_grid_sync();
  const DropCtx dc = drop_ctx(drop_p, sid, rng);
  const int D = H * hd, d8 = D >> 3;
  const long long total = static_cast<long long>(B) * Lq * d8;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = static_cast<int>(i % d8);
    const long long row = i / d8;
    const int b = static_cast<int>(row / Lq), q = static_cast<int>(row - static_cast<long long>(b) * Lq);
    const int h = (c * 8) / hd;
    const float m = keep_mult(dc, (static_cast<unsigned long long>(b) * H + h) * Lq + q);
    float f[8];
    load_bf16x8(v + static_cast<long long>(b) * D + c * 8, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] *= m;
    store_bf16x8(ctx + row * D + c * 8, f);
  }
}

// backward: dv[b, c] = sum_q keep(b, h, q) * dctx[b*Lq + q, c]
__global__ void xattn1_bwd_kernel(const __nv_bfloat16* __restrict__ dctx, __nv_bfloat16* __restrict__ dv, int B, int H,
                                  int Lq, int hd, float drop_p, uint32_t sid, const unsigned long long* __restrict__ rng) {
  pdl_grid_sync();
  const DropCtx dc = drop_ctx(drop_p, sid, rng);
  const int D = H * hd, d8 = D >> 3;
  const int total = B * d8;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % d8, b = i / d8;
    const int h = (c * 8) / hd;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int q = 0; q < Lq; ++q) {
      const float m = keep_mult(dc, (static_cast<unsigned long long>(b) * H + h) * Lq + q);
      float f[8];
      load_bf16x8(dctx + (static_cast<long long>(b) * Lq + q) * D + c * 8, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaf(m, f[k], acc[k]);
    }
    store_bf16x8(dv + static_cast<long long>(b) * D + c * 8, acc);
  }
}

// index of the last position whose mask is 1 (0 when there is none, or when mask == NULL): model/vit_vqa_model.py:215
__device__ __forceinline__ int last_one(const long long* mask, int b, int L) {
  int idx = 0;
  if (mask != nullptr)
    for (int j = 0; j < L; ++j)
      if (mask[static_cast<long long>(b) * L + j] == 1) idx = j;
  return idx;
}

// out[b, :] = src[b*L + last_one(b), :]   (fp32 in, bf16 and / or fp32 out)
__global__ void gather_rows_kernel(const float* __restrict__ src, const long long* __restrict__ mask,
                                   __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ out_f32, int B, int L, int D) {
  pdl_grid_sync();
  const int b = blockIdx.x;
  const int idx = last_one(mask, b, L);
  const float* s = src + (static_cast<long long>(b) * L + idx) * D;
  for (int c = threadIdx.x; c < (D >> 3); c += blockDim.x) {
    float f[8];
    load_f32x8(s + c * 8, f);
    if (out_bf16 != nullptr) store_bf16x8(out_bf16 + static_cast<long long>(b) * D + c * 8, f);
    if (out_f32 != nullptr) store_f32x8(out_f32 + static_cast<long long>(b) * D + c * 8, f);
  }
}

// its backward: dst[b*L + j, :] = (j == last_one(b)) ? src[b, :] : 0
__global__ void scatter_rows_kernel(const float* __restrict__ src, const long long* __restrict__ mask,
                                    float* __restrict__ dst, int B, int L, int D) {
  pdl_grid_sync();
  const int row = blockIdx.x, b = row / L, j = row - b * L;
  const bool hit = j == last_one(mask, b, L);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = threadIdx.x; c < (D >> 2); c += blockDim.x)
    reinterpret_cast<float4*>(dst + static_cast<long long>(row) * D)[c] =
        hit ? reinterpret_cast<const float4*>(src + static_cast<long long>(b) * D)[c] : z;
}

// decoder self-attention: keys after the query are invisible (additive finfo.min, as HF's causal mask)
__global__ void t5_bias_causal_kernel(float* __restrict__ bias, int H, int L) {
  pdl_grid_sync();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * L * L) return;
  const int ij = idx % (L * L), i = ij / L, j = ij - i * L;
  if (j > i) bias[idx] = -3.4028234663852886e38f;
}

// gradient through Dropout(ReLU(.)) given the layer's OUTPUT y (zero where dropped or clipped): out = y > 0 ? dy * scale : 0
__global__ void relu_dropout_bwd_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ y,
                                        __nv_bfloat16* __restrict__ out, float scale, long long n8) {
  pdl_grid_sync();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float g[8], f[8];
    load_f32x8(dy + i * 8, g);
    load_bf16x8(y + i * 8, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] = f[k] > 0.f ? g[k] * scale : 0.f;
    store_bf16x8(out + i * 8, g);
  }
}

bool bad_align(const void* p, int bytes) { return (reinterpret_cast<uintptr_t>(p) & (bytes - 1)) != 0; }

}  // namespace

extern "C" {

int vqa_vit_patchify(void* plan, const float* img, void* out, int B, int H, int W, int P, void* stream) {
  if (P < 8 || (P & 7) || H % P || W % P || bad_align(img, 16) || bad_align(out, 16)) {
    set_last_error("vit_patchify: patch size must be a multiple of 8 dividing the image, pointers 16-byte aligned");
    return -1;
  }
  note_op("vit_patchify", 0.0, 6.0 * B * 3 * H * W);
  return submit(plan, stream, [=](cudaStream_t s) {
    launch_pdl(vit_patchify_kernel, dim3(grid_for(static_cast<long long>(B) * 3 * H * (W >> 3), 256)), dim3(256), 0, s, img,
               static_cast<__nv_bfloat16*>(out), B, H, W, P);
    return launch_status("vit_patchify");
  });
}

int vqa_vit_assemble(void* plan, const float* patch, const float* cls, const float* pos, float* hidden, int B, int NP,
                     int D, void* stream) {
  if ((D & 3) || bad_align(patch, 16) || bad_align(cls, 16) || bad_align(pos, 16) || bad_align(hidden, 16)) {
    set_last_error("vit_assemble: D % 4 == 0 and 16-byte aligned pointers required");
    return -1;
  }
  note_op("vit_assemble", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    launch_pdl(vit_assemble_kernel, dim3(grid_for(static_cast<long long>(B) * (NP + 1) * (D >> 2), 256)), dim3(256), 0, s,
               patch, cls, pos, hidden, B, NP, D);
    return launch_status("vit_assemble");
  });
}

int vqa_gelu_bf16(void* plan, void* x, long long n, void* stream) {
  if ((n & 7) || bad_align(x, 16)) { set_last_error("gelu_bf16: n % 8 == 0 and a 16-byte aligned pointer required"); return -1; }
  note_op("gelu_bf16", 0.0, 4.0 * n);
  return submit(plan, stream, [=](cudaStream_t s) {
    launch_pdl(gelu_bf16_kernel, dim3(grid_for(n >> 3, 256)), dim3(256), 0, s, static_cast<__nv_bfloat16*>(x), n >> 3);
    return launch_status("gelu_bf16");
  });
}

int vqa_vit_fuse_concat(void* plan, const float* pooled_pre, const float* enc, int L, void* out, float* pooled_out, int B,
                        int D, void* stream) {
  if ((D & 7) || bad_align(pooled_pre, 16) || bad_align(enc, 16) || bad_align(out, 16)) {
    set_last_error("vit_fuse_concat: D % 8 == 0 and 16-byte aligned pointers required");
    return -1;
  }
  note_op("vit_fuse_concat", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    launch_pdl(vit_fuse_concat_kernel, dim3(grid_for(static_cast<long long>(B) * 2 * (D >> 3), 128)), dim3(128), 0, s,
               pooled_pre, enc, L, static_cast<__nv_bfloat16*>(out), pooled_out, B, D);
    return launch_status("vit_fuse_concat");
  });
}

int vqa_xattn1_fwd(void* plan, const void* v, void* ctx, int B, int H, int Lq, int hd, float drop_p, uint32_t sid,
                   const uint64_t* rng, void* stream) {
  if ((hd & 7) || bad_align(v, 16) || bad_align(ctx, 16)) { set_last_error("xattn1_fwd: hd % 8 == 0, aligned pointers"); return -1; }
  note_op("xattn1_fwd", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    launch_pdl(xattn1_fwd_kernel, dim3(grid_for(static_cast<long long>(B) * Lq * (H * hd >> 3), 256)), dim3(256), 0, s,
               static_cast<const __nv_bfloat16*>(v), static_cast<__nv_bfloat16*>(ctx), B, H, Lq, hd, drop_p, sid,
               reinterpret_cast<const unsigned long long*>(rng));
    return launch_status("xattn1_fwd");
  });
}

int vqa_xattn1_bwd(void* plan, const void* dctx, void* dv, int B, int H, int Lq, int hd, float drop_p, uint32_t sid,
                   const uint64_t* rng, void* stream) {
  if ((hd & 7) || bad_align(dctx, 16) || bad_align(dv, 16)) { set_last_error("xattn1_bwd: hd % 8 == 0, aligned pointers"); return -1; }
  note_op("xattn1_bwd", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    launch_pdl(xattn1_bwd_kernel, dim3(grid_for(static_cast<long long>(B) * (H * hd >> 3), 64)), dim3(64), 0, s,
               static_cast<const __nv_bfloat16*>(dctx), static_cast<__nv_bfloat16*>(dv), B, H, Lq, hd, drop_p, sid,
               reinterpret_cast<const unsigned long long*>(rng));
    return launch_status("xattn1_bwd");
  });
}

int vqa_gather_rows(void* plan, const float* src, const long long* mask, void* out_bf16, float* out_f32, int B, int L,
                    int D, void* stream) {
  if ((D & 7) || bad_align(src, 16)) { set_last_error("gather_rows: D % 8 == 0 and aligned pointers required"); return -1; }
  note_op("gather_rows", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    launch_pdl(gather_rows_kernel, dim3(B), dim3(96), 0, s, src, mask, static_cast<__nv_bfloat16*>(out_bf16), out_f32, B, L, D);
    return launch_status("gather_rows");
  });
}

int vqa_scatter_rows(void* plan, const float* src, const long long* mask, float* dst, int B, int L, int D, void* stream) {
  if ((D & 3) || bad_align(src, 16) || bad_align(dst, 16)) { set_last_error("scatter_rows: D % 4 == 0 and aligned pointers required"); return -1; }
  note_op("scatter_rows", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    launch_pdl(scatter_rows_kernel, dim3(B * L), dim3(192), 0, s, src, mask, dst, B, L, D);
    return launch_status("scatter_rows");
  });
}

int vqa_t5_bias_causal(void* plan, float* bias, int H, int L, void* stream) {
  note_op("t5_bias_causal", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    launch_pdl(t5_bias_causal_kernel, dim3((H * L * L + 255) / 256), dim3(256), 0, s, bias, H, L);
    return launch_status("t5_bias_causal");
  });
}

int vqa_relu_dropout_bwd(void* plan, const float* dy, const void* y, void* out, float scale, long long n, void* stream) {
  if ((n & 7) || bad_align(dy, 16) || bad_align(y, 16) || bad_align(out, 16)) {
    set_last_error("relu_dropout_bwd: n % 8 == 0 and 16-byte aligned pointers required");
    return -1;
  }
  note_op("relu_dropout_bwd", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    launch_pdl(relu_dropout_bwd_kernel, dim3(grid_for(n >> 3, 256)), dim3(256), 0, s, dy, static_cast<const __nv_bfloat16*>(y),
               static_cast<__nv_bfloat16*>(out), scale, n >> 3);
    return launch_status("relu_dropout_bwd");
  });
}

}  // extern "C"
