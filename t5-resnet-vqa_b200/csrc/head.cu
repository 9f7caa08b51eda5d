// Head of the model: AttentionPooler (model/resnet_vqa_model.py:14-26) and log_softmax + NLLLoss(mean)
// (model/resnet_vqa_model.py:154-160), forward and backward.  Tiny, latency-bound kernels: one CTA per
// sample for the pooler, one CTA for the whole loss so the batch mean is summed in a fixed order.
#include "../../include/vqa_b200.h"
#include "common.cuh"

using namespace vqa;

namespace {

constexpr int kPoolThreads = 256;

// x [B, L, D] fp32; scores_l = x_l . a + b; w = softmax_l(scores); pooled = sum_l w_l x_l
__global__ void __launch_bounds__(kPoolThreads)
pooler_fwd_kernel(const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ bptr,
                  float* __restrict__ w_out, float* __restrict__ pooled_f32, __nv_bfloat16* __restrict__ pooled_bf16,
                  int L, int D) {
  pdl_grid_sync();
  extern __shared__ float sc[];  // [L]
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xb = x + static_cast<long long>(b) * L * D;
  const float bias = bptr[0];
  for (int l = warp; l < L; l += kPoolThreads / 32) {
    float s = 0.f;
    for (int d = lane * 4; d < D; d += 128) {
      const float4 xv = *reinterpret_cast<const float4*>(xb + static_cast<long long>(l) * D + d);
      const float4 av = *reinterpret_cast<const float4*>(a + d);
      s += xv.x * av.x + xv.y * av.y + xv.z * av.z + xv.w * av.w;
    }
    s = warp_sum(s);
    if (lane == 0) sc[l] = s + bias;
  }
  __syncthreads();
  if (warp == 0) {
    float m = -INFINITY;
    for (int l = lane; l < L; l += 32) m = fmaxf(m, sc[l]);
    m = warp_max(m);
    float sum = 0.f;
    for (int l = lane; l < L; l += 32) sum += expf(sc[l] - m);
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int l = lane; l < L; l += 32) {
      const float w = expf(sc[l] - m) * inv;
      sc[l] = w;
      w_out[static_cast<long long>(b) * L + l] = w;
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += kPoolThreads) {
    float acc = 0.f;
    for (int l = 0; l < L; ++l) acc += sc[l] * xb[static_cast<long long>(l) * D + d];
    if (pooled_f32 != nullptr) pooled_f32[static_cast<long long>(b) * D + d] = acc;
    if (pooled_bf16 != nullptr) pooled_bf16[static_cast<long long>(b) * D + d] = __float2bfloat16_rn(acc);
  }
}

__global__ void __launch_bounds__(kPoolThreads)
pooler_bwd_kernel(const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ w,
                  const float* __restrict__ dpooled, float* __restrict__ dx, float* __restrict__ da,
                  float* __restrict__ db, int L, int D) {
  pdl_grid_sync();
  extern __shared__ float sm[];  // ds[L]
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xb = x + static_cast<long long>(b) * L * D;
  const float* wb = w + static_cast<long long>(b) * L;
  const float* dp = dpooled + static_cast<long long>(b) * D;
  // dw_l = dpooled . x_l
  for (int l = warp; l < L; l += kPoolThreads / 32) {
    float s = 0.f;
    for (int d = lane * 4; d < D; d += 128) {
      const float4 xv = *reinterpret_cast<const float4*>(xb + static_cast<long long>(l) * D + d);
      const float4 gv = *reinterpret_cast<const float4*>(dp + d);
      s += xv.x * gv.x + xv.y * gv.y + xv.z * gv.z + xv.w * gv.w;
    }
    s = warp_sum(s);
    if (lane == 0) sm[l] = s;
  }
  __syncthreads();
  __shared__ float s_dot, s_dsum;
  if (warp == 0) {
    float dot = 0.f;
    for (int l = lane; l < L; l += 32) dot += wb[l] * sm[l];
    dot = warp_sum(dot);
    if (lane == 0) s_dot = dot;
  }
  __syncthreads();
  if (warp == 0) {
    const float dot = s_dot;
    float dsum = 0.f;
    for (int l = lane; l < L; l += 32) {
      const float ds = wb[l] * (sm[l] - dot);
      sm[l] = ds;
      dsum += ds;
    }
    dsum = warp_sum(dsum);
    if (lane == 0) s_dsum = dsum;
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += kPoolThreads) {
    const float g = dp[d], av = a[d];
    float dacc = 0.f;
    for (int l = 0; l < L; ++l) {
      const float ds = sm[l];
      const float xv = xb[static_cast<long long>(l) * D + d];
      dx[(static_cast<long long>(b) * L + l) * D + d] = wb[l] * g + ds * av;
      dacc += ds * xv;
    }
    atomicAdd(da + d, dacc);
  }
  if (threadIdx.x == 0) atomicAdd(db, s_dsum);
}

// one CTA; warps loop over rows; the batch mean is accumulated in row order by thread 0
__global__ void __launch_bounds__(1024)
logsoftmax_nll_fwd_kernel(const float* __restrict__ logits, long long ld, const long long* __restrict__ labels,
                          float* __restrict__ logp, float* __restrict__ loss, int B, int A) {
  pdl_grid_sync();
  extern __shared__ float nll[];  // [B]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = warp; row < B; row += 32) {
    const float* lr = logits + static_cast<long long>(row) * ld;
    float m = -INFINITY;
    for (int j = lane; j < A; j += 32) m = fmaxf(m, lr[j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < A; j += 32) sum += expf(lr[j] - m);
    sum = warp_sum(sum);
    const float lse = m + logf(sum);
    for (int j = lane; j < A; j += 32) logp[static_cast<long long>(row) * A + j] = lr[j] - lse;
    if (lane == 0 && labels != nullptr) {
      long long y = labels[row];
      if (y < 0) y = 0;
      if (y >= A) y = A - 1;
      nll[row] = lse - lr[y];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && labels != nullptr && loss != nullptr) {
    float s = 0.f;
    for (int r = 0; r < B; ++r) s += nll[r];
    loss[0] = s / static_cast<float>(B);
  }
}

__global__ void logsoftmax_nll_bwd_kernel(const float* __restrict__ logp, const long long* __restrict__ labels,
                                          const float* __restrict__ gloss, const float* __restrict__ glogp,
                                          __nv_bfloat16* __restrict__ dlogits, long long ld, int B, int A) {
  pdl_grid_sync();
  const int row = blockIdx.x;
  const int lane = threadIdx.x;  // one warp per row
  const float gl = (labels != nullptr) ? (gloss != nullptr ? gloss[0] : 1.f) / static_cast<float>(B) : 0.f;
  long long y = labels != nullptr ? labels[row] : -1;
  if (labels != nullptr) { if (y < 0) y = 0; if (y >= A) y = A - 1; }
  float gsum = 0.f;
  if (glogp != nullptr) {
    for (int j = lane; j < A; j += 32) gsum += glogp[static_cast<long long>(row) * A + j];
    gsum = warp_sum(gsum);
  }
  for (int j = lane; j < ld; j += 32) {
    float v = 0.f;
    if (j < A) {
      const float pr = expf(logp[static_cast<long long>(row) * A + j]);
      v = gl * (pr - (j == y ? 1.f : 0.f));
      if (glogp != nullptr) v += glogp[static_cast<long long>(row) * A + j] - pr * gsum;
    }
    dlogits[static_cast<long long>(row) * ld + j] = __float2bfloat16_rn(v);
  }
}

}  // namespace

extern "C" {

int vqa_pooler_fwd(void* plan, const float* x, const float* a, const float* b, float* w_out, float* pooled_f32,
                   void* pooled_bf16, int B, int L, int D, void* stream) {
  if (D % 4) { set_last_error("pooler: D must be a multiple of 4"); return -1; }
  note_op("pooler_fwd", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    launch_pdl(pooler_fwd_kernel, dim3(B), dim3(kPoolThreads), L * sizeof(float), s, x, a, b, w_out, pooled_f32,
                                                                 static_cast<__nv_bfloat16*>(pooled_bf16), L, D);
    return launch_status("pooler_fwd");
  });
}

int vqa_pooler_bwd(void* plan, const float* x, const float* a, const float* w, const float* dpooled, float* dx,
                   float* da, float* db, int B, int L, int D, void* stream) {
  if (D % 4) { set_last_error("pooler: D must be a multiple of 4"); return -1; }
  note_op("pooler_bwd", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    launch_pdl(pooler_bwd_kernel, dim3(B), dim3(kPoolThreads), L * sizeof(float), s, x, a, w, dpooled, dx, da, db, L, D);
    return launch_status("pooler_bwd");
  });
}

int vqa_logsoftmax_nll_fwd(void* plan, const float* logits, long long ld, const long long* labels, float* logp,
                           float* loss, int B, int A, void* stream) {
  if (B > 8192) { set_last_error("logsoftmax_nll: batch > 8192 not supported"); return -1; }
  note_op("logsoftmax_nll_fwd", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    launch_pdl(logsoftmax_nll_fwd_kernel, dim3(1), dim3(1024), B * sizeof(float), s, logits, ld, labels, logp, loss, B, A);
    return launch_status("logsoftmax_nll_fwd");
  });
}

int vqa_logsoftmax_nll_bwd(void* plan, const float* logp, const long long* labels, const float* gloss,
                           const float* glogp, void* dlogits, long long ld, int B, int A, void* stream) {
  note_op("logsoftmax_nll_bwd", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    launch_pdl(logsoftmax_nll_bwd_kernel, dim3(B), dim3(32), 0, s, logp, labels, gloss, glogp, static_cast<__nv_bfloat16*>(dlogits), ld,
                                               B, A);
    return launch_status("logsoftmax_nll_bwd");
  });
}

}  // extern "C"
