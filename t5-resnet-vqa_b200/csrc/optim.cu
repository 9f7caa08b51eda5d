// Optimizer step of the reference trainer (trainer/faster_rcnn_vqa_trainer.py:399-404): global gradient
// norm for clip_grad_norm_, and torch.optim.AdamW(amsgrad=True) as one fused multi-state pass.
// HBM-bound: 36 B/parameter (read p, g, m, v, vmax; write p, m, v, vmax) + 2 B for the bf16 shadow.
#include "../../include/vqa_b200.h"
#include "common.cuh"

#include <math.h>
#include <stdlib.h>

using namespace vqa;

namespace {

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  pdl_grid_sync();
  __shared__ float red[8];
  float acc = 0.f;
  const long long nvec = n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = x[(nvec << 2) + threadIdx.x];
    acc += v * v;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}

// scalar combinations are formed on the host in double, as torch does, then rounded to fp32 once
struct AdamHyper {
  float decay;      // 1 - lr * weight_decay
  float w1;         // 1 - beta1
  float beta2, w2;  // beta2, 1 - beta2
  float eps;
  float step_size;  // lr / (1 - beta1^t)
  float bc2_sqrt;   // sqrt(1 - beta2^t)
  float max_norm;
  int amsgrad;
};

__device__ __forceinline__ float adam_one(float& p, float g, float& m, float& v, float& vmax, const AdamHyper& h,
                                          float clip) {
  g *= clip;
  p *= h.decay;                                   // param.mul_(1 - lr * weight_decay)
  m = m + h.w1 * (g - m);                         // exp_avg.lerp_(grad, 1 - beta1)
  v = v * h.beta2 + h.w2 * g * g;                 // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  float vhat = v;
  if (h.amsgrad) { vmax = fmaxf(vmax, v); vhat = vmax; }
  const float denom = __fdiv_rn(__fsqrt_rn(vhat), h.bc2_sqrt) + h.eps;
  p = p - h.step_size * __fdiv_rn(m, denom);      // param.addcdiv_(exp_avg, denom, value=-step_size)
  return p;
}

__global__ void __launch_bounds__(1024)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             float* __restrict__ vmax, __nv_bfloat16* __restrict__ shadow, long long n, AdamHyper h,
             const float* __restrict__ gnorm_sq) {
  pdl_grid_sync();
  float clip = 1.f;
  if (gnorm_sq != nullptr) {
    const float norm = __fsqrt_rn(gnorm_sq[0]);
    clip = fminf(1.f, __fdiv_rn(h.max_norm, norm + 1e-6f));
  }
  const long long nvec = n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float4 xv = h.amsgrad ? reinterpret_cast<float4*>(vmax)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    adam_one(pv.x, gv.x, mv.x, vv.x, xv.x, h, clip);
    adam_one(pv.y, gv.y, mv.y, vv.y, xv.y, h, clip);
    adam_one(pv.z, gv.z, mv.z, vv.z, xv.z, h, clip);
    adam_one(pv.w, gv.w, mv.w, vv.w, xv.w, h, clip);
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (h.amsgrad) reinterpret_cast<float4*>(vmax)[i] = xv;
    if (shadow != nullptr) {
      uint2 pk;
      pk.x = pack_bf16x2(pv.x, pv.y);
      pk.y = pack_bf16x2(pv.z, pv.w);
      reinterpret_cast<uint2*>(shadow)[i] = pk;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (nvec << 2) + threadIdx.x;
    float pv = p[i], mv = m[i], vv = v[i], xv = h.amsgrad ? vmax[i] : 0.f;
    adam_one(pv, g[i], mv, vv, xv, h, clip);
    p[i] = pv; m[i] = mv; v[i] = vv;
    if (h.amsgrad) vmax[i] = xv;
    if (shadow != nullptr) shadow[i] = __float2bfloat16_rn(pv);
  }
}

// clip_grad_norm_'s in-place scaling: g *= min(1, max_norm / (||g|| + 1e-6)).  When the norm is already below
// max_norm the factor clamps to exactly 1 (torch multiplies by 1.0 there), so the pass is skipped.
__global__ void __launch_bounds__(256)
clip_scale_kernel(float* __restrict__ g, long long n, const float* __restrict__ gnorm_sq, float max_norm) {
  pdl_grid_sync();
  const float norm = __fsqrt_rn(gnorm_sq[0]);
  const float clip = __fdiv_rn(max_norm, norm + 1e-6f);
  if (!(clip < 1.f)) return;
  const long long nvec = n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float4 v = reinterpret_cast<float4*>(g)[i];
    v.x *= clip; v.y *= clip; v.z *= clip; v.w *= clip;
    reinterpret_cast<float4*>(g)[i] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) g[(nvec << 2) + threadIdx.x] *= clip;
}

__global__ void rng_advance_kernel(unsigned long long* rng) {
  pdl_grid_sync(); rng[1] += 1ull; }

inline int stream_grid(long long items, int threads) {
  long long b = (items + threads - 1) / threads;
  if (b > 148 * 16) b = 148 * 16;
  return b < 1 ? 1 : static_cast<int>(b);
}

// Persistent grid of the AdamW pass.  Measured on the whole step (bench.py, B200): 16 / 4 / 3 / 2 CTAs per SM give
// 5.98 / 6.00 / 6.04 / 6.13 ms, i.e. leaving thread slots free for the next step's backbone convolutions (which run
// beside this pass on another stream) buys nothing: both are HBM-bound there.  VQA_B200_ADAMW_CTAS overrides.
inline int adamw_grid(long long items, int threads) {
  static int per_sm = 0;
  if (per_sm == 0) {
    const char* e = getenv("VQA_B200_ADAMW_CTAS");
    per_sm = e ? atoi(e) : 16;
    if (per_sm < 1 || per_sm > 16) per_sm = 16;
  }
  long long b = (items + threads - 1) / threads;
  if (b > 148LL * per_sm) b = 148LL * per_sm;
  return b < 1 ? 1 : static_cast<int>(b);
}

}  // namespace

extern "C" {

int vqa_sumsq_f32(void* plan, const float* x, long long n, float* out, void* stream) {
  if (reinterpret_cast<uintptr_t>(x) & 15) { set_last_error("sumsq: pointer must be 16-byte aligned"); return -1; }
  note_op("sumsq", 0.0, 4.0 * static_cast<double>(n));
  return submit(plan, stream, [=](cudaStream_t s) {
    launch_pdl(sumsq_kernel, dim3(stream_grid((n >> 2) + 4, 256)), dim3(256), 0, s, x, n, out);
    return launch_status("sumsq");
  });
}

int vqa_clip_scale_f32(void* plan, float* g, long long n, const float* gnorm_sq, float max_norm, void* stream) {
  if (reinterpret_cast<uintptr_t>(g) & 15) { set_last_error("clip_scale: pointer must be 16-byte aligned"); return -1; }
  if (gnorm_sq == nullptr) { set_last_error("clip_scale: needs the squared gradient norm"); return -1; }
  note_op("clip_scale", 0.0, 8.0 * static_cast<double>(n));
  return submit(plan, stream, [=](cudaStream_t s) {
    launch_pdl(clip_scale_kernel, dim3(stream_grid((n >> 2) + 4, 256)), dim3(256), 0, s, g, n, gnorm_sq, max_norm);
    return launch_status("clip_scale");
  });
}

int vqa_adamw_amsgrad(void* plan, float* p, const float* g, float* m, float* v, float* vmax, void* shadow,
                      long long n, double lr, double beta1, double beta2, double eps, double weight_decay,
                      double bias_correction1, double bias_correction2, const float* gnorm_sq, float max_norm,
                      int amsgrad, void* stream) {
  const uintptr_t al = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                       reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) |
                       (amsgrad ? reinterpret_cast<uintptr_t>(vmax) : 0);
  if (al & 15) { set_last_error("adamw: pointers must be 16-byte aligned"); return -1; }
  if (shadow != nullptr && (reinterpret_cast<uintptr_t>(shadow) & 7)) {
    set_last_error("adamw: shadow must be 8-byte aligned");
    return -1;
  }
  AdamHyper h;
  h.decay = static_cast<float>(1.0 - lr * weight_decay);
  h.w1 = static_cast<float>(1.0 - beta1);
  h.beta2 = static_cast<float>(beta2);
  h.w2 = static_cast<float>(1.0 - beta2);
  h.eps = static_cast<float>(eps);
  h.step_size = static_cast<float>(lr / bias_correction1);
  h.bc2_sqrt = static_cast<float>(sqrt(bias_correction2));
  h.max_norm = max_norm; h.amsgrad = amsgrad;
  note_op("adamw", 0.0, (amsgrad ? 36.0 : 28.0) * static_cast<double>(n) + (shadow ? 2.0 * n : 0.0));
  return submit(plan, stream, [=](cudaStream_t s) {
    // VQA_B200_ADAMW_GRID=<n>: n CTAs of 1024 threads instead (experiment: confine the HBM-bound pass to a subset of the
    // SMs so that GEMM CTAs, which need a whole SM's register file, can run on the others)
    static int wide = -1;
    if (wide < 0) { const char* e = getenv("VQA_B200_ADAMW_GRID"); wide = e ? atoi(e) : 0; }
    if (wide > 0) {
      launch_pdl(adamw_kernel, dim3(wide), dim3(1024), 0, s, p, g, m, v, vmax, static_cast<__nv_bfloat16*>(shadow), n, h,
                 gnorm_sq);
      return launch_status("adamw");
    }
    launch_pdl(adamw_kernel, dim3(adamw_grid((n >> 2) + 4, 256)), dim3(256), 0, s, p, g, m, v, vmax, static_cast<__nv_bfloat16*>(shadow), n, h,
                                                                gnorm_sq);
    return launch_status("adamw");
  });
}

int vqa_rng_advance(void* plan, uint64_t* rng, void* stream) {
  note_op("rng_advance", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    launch_pdl(rng_advance_kernel, dim3(1), dim3(1), 0, s, reinterpret_cast<unsigned long long*>(rng));
    return launch_status("rng_advance");
  });
}

}  // extern "C"
