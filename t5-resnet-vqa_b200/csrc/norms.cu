// T5LayerNorm (RMS, hf:55-68) and nn.LayerNorm (model/multi_head_vision_text_attn.py:120-126), forward and
// backward.  HBM-bound: one warp per row, 128-bit coalesced accesses (lane l owns columns
// (c*32 + l)*8 .. +8 for chunk c), warp-shuffle row reductions, per-thread column partials for the
// weight gradients reduced through shared memory and finished with one red.add per column per CTA.
#include "../../include/vqa_b200.h"
#include "common.cuh"

using namespace vqa;

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;

inline int norm_grid(int M) {
  int g = (M + kWarps - 1) / kWarps;
  if (g > 148 * 2) g = 148 * 2;
  return g < 1 ? 1 : g;
}

// Column sums of the per-warp partials in shared memory -> one vector red.add (4 columns) per thread into global:
// a quarter of the L2 atomic operations of a per-column atomicAdd (the per-address serialisation at the L2 is what
// bounds these kernels' tails when 256 CTAs add into the same 768 columns).
template <int D>
__device__ __forceinline__ void reduce_cols_atomic(const float (*red)[D], float* __restrict__ dst) {
  for (int c4 = threadIdx.x; c4 < D / 4; c4 += kThreads) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < kWarps; ++j) {
      const float4 v = *reinterpret_cast<const float4*>(&red[j][c4 * 4]);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    atomicAdd(reinterpret_cast<float4*>(dst) + c4, s);
  }
}

// --------------------------------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(kThreads)
rmsnorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, __nv_bfloat16* __restrict__ y_bf16,
                   float* __restrict__ y_f32, float* __restrict__ rstd, int M, float eps, float drop_p,
                   uint32_t sid, const unsigned long long* __restrict__ rng, int split) {
  // parameters are only written by optimizer kernels, which are full (event) dependencies of the plan: they may be
  // read before the programmatic-dependency wait, which only guards the previous kernel's activations
  pdl_launch_dependents();
  constexpr int D = NC * 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float wv[NC][8];
#pragma unroll
  for (int c = 0; c < NC; ++c) load_f32x8(w + (c * 32 + lane) * 8, wv[c]);
  pdl_wait();
  const DropCtx dc = drop_ctx(drop_p, sid, rng);
  for (int row = blockIdx.x * kWarps + warp; row < M; row += gridDim.x * kWarps) {
    const float* xr = x + static_cast<long long>(row) * D;
    float xv[NC][8];
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      load_f32x8(xr + (c * 32 + lane) * 8, xv[c]);
#pragma unroll
      for (int i = 0; i < 8; ++i) ss += xv[c][i] * xv[c][i];
    }
    ss = warp_sum(ss);
    const float r = rsqrtf(ss * (1.f / D) + eps);
    if (lane == 0 && rstd != nullptr) rstd[row] = r;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = wv[c][i] * (xv[c][i] * r);
      const long long col = (c * 32 + lane) * 8;
      drop8(dc, static_cast<unsigned long long>(row) * D + col, o);
      if (split) {
        // two-term split for the GEMMs of the early T5 blocks: [row, 0:D) = bf16(y), [row, D:2D) = bf16(y - bf16(y))
        __nv_bfloat16* yr = y_bf16 + static_cast<long long>(row) * (2 * D) + col;
        float lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) lo[i] = o[i] - __bfloat162float(__float2bfloat16_rn(o[i]));
        store_bf16x8(yr, o);
        store_bf16x8(yr + D, lo);
        continue;
      }
      if (y_bf16 != nullptr) store_bf16x8(y_bf16 + static_cast<long long>(row) * D + col, o);
      if (y_f32 != nullptr) store_f32x8(y_f32 + static_cast<long long>(row) * D + col, o);
    }
  }
}

template <int NC>
__global__ void __launch_bounds__(kThreads)
rmsnorm_bwd_kernel(const void* __restrict__ dy, int dy_fp32, const float* __restrict__ x,
                   const float* __restrict__ w, const float* __restrict__ rstd, const float* __restrict__ dres,
                   float* __restrict__ dx, float* __restrict__ dw, int M, float drop_p, uint32_t sid,
                   const unsigned long long* __restrict__ rng, __nv_bfloat16* __restrict__ g_out, float g_drop_p,
                   uint32_t g_sid) {
  pdl_launch_dependents();
  constexpr int D = NC * 256;
  __shared__ __align__(16) float red[kWarps][D];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float wv[NC][8], dwp[NC][8];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    load_f32x8(w + (c * 32 + lane) * 8, wv[c]);   // parameters: safe before the dependency wait (see rmsnorm_fwd)
#pragma unroll
    for (int i = 0; i < 8; ++i) dwp[c][i] = 0.f;
  }
  pdl_wait();
  const DropCtx dc = drop_ctx(drop_p, sid, rng);
  const DropCtx dcg = drop_ctx(g_out != nullptr ? g_drop_p : 0.f, g_sid, rng);
  for (int row = blockIdx.x * kWarps + warp; row < M; row += gridDim.x * kWarps) {
    const long long base = static_cast<long long>(row) * D;
    const float r = rstd[row];
    float xv[NC][8], g[NC][8], dr[NC][8];
    float dot = 0.f;
    // every load of the row is issued before the first use (one memory round trip per row instead of three)
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const long long col = (c * 32 + lane) * 8;
      load_f32x8(x + base + col, xv[c]);
      if (dy_fp32) load_f32x8(static_cast<const float*>(dy) + base + col, g[c]);
      else load_bf16x8(static_cast<const __nv_bfloat16*>(dy) + base + col, g[c]);
      if (dres != nullptr) load_f32x8(dres + base + col, dr[c]);
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const long long col = (c * 32 + lane) * 8;
      float d[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = g[c][i];
      drop8(dc, static_cast<unsigned long long>(base + col), d);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        g[c][i] = d[i] * wv[c][i];
        dot += g[c][i] * xv[c][i];
        dwp[c][i] += d[i] * xv[c][i] * r;
      }
    }
    dot = warp_sum(dot);
    const float k = r * r * r * dot * (1.f / D);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const long long col = (c * 32 + lane) * 8;
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = r * g[c][i] - xv[c][i] * k;
      if (dres != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += dr[c][i];
      }
      store_f32x8(dx + base + col, o);
      if (g_out != nullptr) {   // dropout-masked bf16 copy: the GEMM operand of the next sub-layer's backward
        drop8(dcg, static_cast<unsigned long long>(base + col), o);
        store_bf16x8(g_out + base + col, o);
      }
    }
  }
  if (dw != nullptr) {
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int i = 0; i < 8; ++i) red[warp][(c * 32 + lane) * 8 + i] = dwp[c][i];
    __syncthreads();
    reduce_cols_atomic<D>(red, dw);
  }
}

// --------------------------------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(kThreads)
layernorm_fwd_kernel(const float* __restrict__ z, const float* __restrict__ gamma, const float* __restrict__ beta,
                     __nv_bfloat16* __restrict__ y_bf16, float* __restrict__ y_f32, float* __restrict__ mean,
                     float* __restrict__ rstd, int M, float eps) {
  pdl_grid_sync();
  constexpr int D = NC * 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float gv[NC][8], bv[NC][8];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    load_f32x8(gamma + (c * 32 + lane) * 8, gv[c]);
    load_f32x8(beta + (c * 32 + lane) * 8, bv[c]);
  }
  for (int row = blockIdx.x * kWarps + warp; row < M; row += gridDim.x * kWarps) {
    const long long base = static_cast<long long>(row) * D;
    float zv[NC][8];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      load_f32x8(z + base + (c * 32 + lane) * 8, zv[c]);
#pragma unroll
      for (int i = 0; i < 8; ++i) s += zv[c][i];
    }
    const float mu = warp_sum(s) * (1.f / D);
    float vs = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = zv[c][i] - mu; vs += d * d; }
    const float r = rsqrtf(warp_sum(vs) * (1.f / D) + eps);
    if (lane == 0) {
      if (mean != nullptr) mean[row] = mu;
      if (rstd != nullptr) rstd[row] = r;
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const long long col = (c * 32 + lane) * 8;
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = (zv[c][i] - mu) * r * gv[c][i] + bv[c][i];
      if (y_bf16 != nullptr) store_bf16x8(y_bf16 + base + col, o);
      if (y_f32 != nullptr) store_f32x8(y_f32 + base + col, o);
    }
  }
}

template <int NC>
__global__ void __launch_bounds__(kThreads)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z, const float* __restrict__ gamma,
                     const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dz,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, int M,
                     __nv_bfloat16* __restrict__ g_out, float g_drop_p, uint32_t g_sid,
                     const unsigned long long* __restrict__ rng, float* __restrict__ g_colsum) {
  pdl_grid_sync();
  constexpr int D = NC * 256;
  __shared__ __align__(16) float red[kWarps][D];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const DropCtx dcg = drop_ctx(g_out != nullptr ? g_drop_p : 0.f, g_sid, rng);
  float gsp[NC][8];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) gsp[c][i] = 0.f;
  float gv[NC][8], dgp[NC][8], dbp[NC][8];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    load_f32x8(gamma + (c * 32 + lane) * 8, gv[c]);
#pragma unroll
    for (int i = 0; i < 8; ++i) { dgp[c][i] = 0.f; dbp[c][i] = 0.f; }
  }
  for (int row = blockIdx.x * kWarps + warp; row < M; row += gridDim.x * kWarps) {
    const long long base = static_cast<long long>(row) * D;
    const float mu = mean[row], r = rstd[row];
    float xh[NC][8], g[NC][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const long long col = (c * 32 + lane) * 8;
      float zv[8], d[8];
      load_f32x8(z + base + col, zv);
      load_f32x8(dy + base + col, d);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xh[c][i] = (zv[i] - mu) * r;
        g[c][i] = d[i] * gv[c][i];
        s1 += g[c][i];
        s2 += g[c][i] * xh[c][i];
        dgp[c][i] += d[i] * xh[c][i];
        dbp[c][i] += d[i];
      }
    }
    s1 = warp_sum(s1) * (1.f / D);
    s2 = warp_sum(s2) * (1.f / D);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = r * (g[c][i] - s1 - xh[c][i] * s2);
      store_f32x8(dz + base + (c * 32 + lane) * 8, o);
      if (g_out != nullptr) {   // masked bf16 copy for the branch GEMMs + its column sum (the branch's bias grad)
        drop8(dcg, static_cast<unsigned long long>(base + (c * 32 + lane) * 8), o);
        store_bf16x8(g_out + base + (c * 32 + lane) * 8, o);
#pragma unroll
        for (int i = 0; i < 8; ++i) gsp[c][i] += o[i];
      }
    }
  }
  // dgamma
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) red[warp][(c * 32 + lane) * 8 + i] = dgp[c][i];
  __syncthreads();
  reduce_cols_atomic<D>(red, dgamma);
  __syncthreads();
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) red[warp][(c * 32 + lane) * 8 + i] = dbp[c][i];
  __syncthreads();
  reduce_cols_atomic<D>(red, dbeta);
  if (g_out != nullptr && g_colsum != nullptr) {
    __syncthreads();
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int i = 0; i < 8; ++i) red[warp][(c * 32 + lane) * 8 + i] = gsp[c][i];
    __syncthreads();
    reduce_cols_atomic<D>(red, g_colsum);
  }
}

#define DISPATCH_NC(D, CALL)                                                         \
  switch ((D) / 256) {                                                               \
    case 1: { constexpr int NC = 1; CALL; break; }                                   \
    case 2: { constexpr int NC = 2; CALL; break; }                                   \
    case 3: { constexpr int NC = 3; CALL; break; }                                   \
    case 4: { constexpr int NC = 4; CALL; break; }                                   \
    default: break;                                                                  \
  }

inline int check_d(int D, const char* what) {
  if (D % 256 != 0 || D < 256 || D > 1024) {
    set_last_error("%s: feature size must be 256/512/768/1024 (got %d)", what, D);
    return -1;
  }
  return 0;
}

}  // namespace

extern "C" {

int vqa_rmsnorm_fwd(void* plan, const float* x, const float* w, void* y_bf16, float* y_f32, float* rstd, int M,
                    int D, float eps, float drop_p, uint32_t sid, const uint64_t* rng, void* stream) {
  if (check_d(D, "rmsnorm_fwd")) return -1;
  note_op("rmsnorm_fwd", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    DISPATCH_NC(D, (launch_pdl(rmsnorm_fwd_kernel<NC>, dim3(norm_grid(M)), dim3(kThreads), 0, s, x, w, static_cast<__nv_bfloat16*>(y_bf16), y_f32, rstd, M, eps, drop_p, sid,
                       reinterpret_cast<const unsigned long long*>(rng), 0)));
    return launch_status("rmsnorm_fwd");
  });
}

int vqa_rmsnorm_fwd_split(void* plan, const float* x, const float* w, void* y_hilo, float* rstd, int M, int D,
                          float eps, void* stream) {
  if (check_d(D, "rmsnorm_fwd_split")) return -1;
  if (y_hilo == nullptr) { set_last_error("rmsnorm_fwd_split: null output"); return -1; }
  note_op("rmsnorm_fwd", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    DISPATCH_NC(D, (launch_pdl(rmsnorm_fwd_kernel<NC>, dim3(norm_grid(M)), dim3(kThreads), 0, s, x, w, static_cast<__nv_bfloat16*>(y_hilo),
                       static_cast<float*>(nullptr), rstd, M, eps, 0.f, 0u, static_cast<const unsigned long long*>(nullptr), 1)));
    return launch_status("rmsnorm_fwd_split");
  });
}

int vqa_rmsnorm_bwd(void* plan, const void* dy, int dy_fp32, const float* x, const float* w, const float* rstd,
                    const float* dres, float* dx, float* dw, int M, int D, float drop_p, uint32_t sid,
                    const uint64_t* rng, void* g_out, float g_drop_p, uint32_t g_sid, void* stream) {
  if (check_d(D, "rmsnorm_bwd")) return -1;
  note_op("rmsnorm_bwd", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    DISPATCH_NC(D, (launch_pdl(rmsnorm_bwd_kernel<NC>, dim3(norm_grid(M)), dim3(kThreads), 0, s, dy, dy_fp32, x, w, rstd, dres, dx, dw, M, drop_p, sid,
                       reinterpret_cast<const unsigned long long*>(rng), static_cast<__nv_bfloat16*>(g_out), g_drop_p,
                       g_sid)));
    return launch_status("rmsnorm_bwd");
  });
}

int vqa_layernorm_fwd(void* plan, const float* z, const float* gamma, const float* beta, void* y_bf16,
                      float* y_f32, float* mean, float* rstd, int M, int D, float eps, void* stream) {
  if (check_d(D, "layernorm_fwd")) return -1;
  note_op("layernorm_fwd", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    DISPATCH_NC(D, (launch_pdl(layernorm_fwd_kernel<NC>, dim3(norm_grid(M)), dim3(kThreads), 0, s, z, gamma, beta, static_cast<__nv_bfloat16*>(y_bf16), y_f32, mean, rstd, M, eps)));
    return launch_status("layernorm_fwd");
  });
}

int vqa_layernorm_bwd(void* plan, const float* dy, const float* z, const float* gamma, const float* mean,
                      const float* rstd, float* dz, float* dgamma, float* dbeta, int M, int D, void* g_out,
                      float g_drop_p, uint32_t g_sid, const uint64_t* rng, float* g_colsum, void* stream) {
  if (check_d(D, "layernorm_bwd")) return -1;
  note_op("layernorm_bwd", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    DISPATCH_NC(D, (launch_pdl(layernorm_bwd_kernel<NC>, dim3(norm_grid(M)), dim3(kThreads), 0, s, dy, z, gamma, mean, rstd, dz, dgamma, dbeta, M,
                       static_cast<__nv_bfloat16*>(g_out), g_drop_p, g_sid,
                       reinterpret_cast<const unsigned long long*>(rng), g_colsum)));
    return launch_status("layernorm_bwd");
  });
}

}  // extern "C"
