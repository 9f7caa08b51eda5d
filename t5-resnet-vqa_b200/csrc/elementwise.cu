// Layout / precision plumbing around the contractions: fp32 -> bf16 shadows, BatchNorm folding,
// ConvTranspose2d <-> Conv2d weight views, stem input packing, max-pool, dropout-mask casts, column sums.
// All HBM-bound: 128-bit accesses, grid-stride loops sized to a multiple of the SM count.
#include "../../include/vqa_b200.h"
#include "common.cuh"

using namespace vqa;

namespace {

constexpr int kSMs = 148;

inline int grid_for(long long work_items, int threads, int max_waves = 8) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(kSMs) * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

// ---- cast -------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                     long long n) {
  pdl_grid_sync();
  const long long nvec = n >> 3;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float f[8];
    load_f32x8(src + i * 8, f);
    store_bf16x8(dst + i * 8, f);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const long long i = (nvec << 3) + threadIdx.x;
    dst[i] = __float2bfloat16_rn(src[i]);
  }
}

__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, long long n) {
  pdl_grid_sync();
  const long long nvec = n >> 3;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float f[8];
    load_bf16x8(src + i * 8, f);
    store_f32x8(dst + i * 8, f);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const long long i = (nvec << 3) + threadIdx.x;
    dst[i] = __bfloat162float(src[i]);
  }
}

// low-order half of the two-term bf16 split of an fp32 range: dst = bf16(src - float(bf16(src)))
__global__ void split_lo_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  pdl_grid_sync();
  const long long nvec = n >> 3;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float f[8];
    load_f32x8(src + i * 8, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] -= __bfloat162float(__float2bfloat16_rn(f[k]));
    store_bf16x8(dst + i * 8, f);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const long long i = (nvec << 3) + threadIdx.x;
    dst[i] = __float2bfloat16_rn(src[i] - __bfloat162float(__float2bfloat16_rn(src[i])));
  }
}

__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, float a, long long n) {
  pdl_grid_sync();
  const long long nvec = n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float4 yv = reinterpret_cast<float4*>(y)[i];
    const float4 xv = reinterpret_cast<const float4*>(x)[i];
    yv.x += a * xv.x; yv.y += a * xv.y; yv.z += a * xv.z; yv.w += a * xv.w;
    reinterpret_cast<float4*>(y)[i] = yv;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (nvec << 2) + threadIdx.x;
    y[i] += a * x[i];
  }
}

// ---- BatchNorm folding --------------------------------------------------------------------------
__global__ void fold_conv_bn_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ mean,
                                    const float* __restrict__ var, float eps,
                                    __nv_bfloat16* __restrict__ w_out, float* __restrict__ bias_out,
                                    int O, int I, int R, int S, int Sp, int Ip) {
  pdl_grid_sync();
  const long long total = static_cast<long long>(O) * R * Sp * Ip;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int ip = static_cast<int>(idx % Ip);
    long long t = idx / Ip;
    const int sp = static_cast<int>(t % Sp); t /= Sp;
    const int r = static_cast<int>(t % R);
    const int o = static_cast<int>(t / R);
    float v = 0.f;
    if (sp < S && ip < I) {
      // 1/sqrt in full precision: the reference divides by sqrt(var + eps) in fp32 (ATen batch_norm)
      const float scale = gamma ? gamma[o] / sqrtf(var[o] + eps) : 1.f;
      v = w[((static_cast<long long>(o) * I + ip) * R + r) * S + sp] * scale;
    }
    w_out[idx] = __float2bfloat16_rn(v);
  }
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (bias_out != nullptr && gid < O) {
    float b = 0.f;
    if (gamma) {
      const float scale = gamma[gid] / sqrtf(var[gid] + eps);
      b = beta[gid] - mean[gid] * scale;
    }
    bias_out[gid] = b;
  }
}

// ---- ConvTranspose2d weight views -------------------------------------------------------------
// src fp32 [R0, C0] (C0 = G*9) -> dst [C0 (with tap flip inside each group of 9), R0]
//   dst[(g*9 + 8 - t), r] = src[r, g*9 + t]
template <typename TOut, typename TIn = float>
__global__ void transpose_flip9_kernel(const TIn* __restrict__ src, TOut* __restrict__ dst, int R0, int C0) {
  pdl_grid_sync();
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < R0 && c < C0) ? static_cast<float>(src[static_cast<long long>(r) * C0 + c]) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (c < C0 && r < R0) {
      const int g = c / 9, t = c - g * 9;
      const long long orow = static_cast<long long>(g) * 9 + (8 - t);
      const float v = tile[threadIdx.x][j];
      if constexpr (sizeof(TOut) == 2) dst[orow * R0 + r] = __float2bfloat16_rn(v);
      else dst[orow * R0 + r] = v;
    }
  }
}

// src fp32 [R0 = G*9, C0] -> dst fp32 [C0, R0] with the tap flip on the source row:
//   dst[c, g*9 + 8 - t] = src[g*9 + t, c]
__global__ void transpose_fliprow9_kernel(const float* __restrict__ src, float* __restrict__ dst, int R0, int C0) {
  pdl_grid_sync();
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < R0 && c < C0) ? src[static_cast<long long>(r) * C0 + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (c < C0 && r < R0) {
      const int g = r / 9, t = r - g * 9;
      dst[static_cast<long long>(c) * R0 + g * 9 + (8 - t)] = tile[threadIdx.x][j];
    }
  }
}

// ---- stem packing -----------------------------------------------------------------------------
__global__ void image_to_stem_kernel(const float* __restrict__ img, uint4* __restrict__ out, int N, int H, int W) {
  pdl_grid_sync();
  const int Wp = W + 8;
  const long long total = static_cast<long long>(N) * H * Wp;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    // 32-bit index arithmetic (the host checks total < 2^31): 64-bit div / mod by run-time values is ~4x the instructions
    const uint32_t i32 = static_cast<uint32_t>(idx);
    const uint32_t t = i32 / static_cast<uint32_t>(Wp);
    const int wp = static_cast<int>(i32 - t * static_cast<uint32_t>(Wp));
    const int n = static_cast<int>(t / static_cast<uint32_t>(H));
    const int h = static_cast<int>(t - static_cast<uint32_t>(n) * static_cast<uint32_t>(H));
    uint4 o = make_uint4(0, 0, 0, 0);
    const int w = wp - 3;
    if (w >= 0 && w < W) {
      const long long plane = static_cast<long long>(H) * W;
      const float* p = img + (static_cast<long long>(n) * 3) * plane + static_cast<long long>(h) * W + w;
      o.x = pack_bf16x2(p[0], p[plane]);
      o.y = pack_bf16x2(p[2 * plane], 0.f);
    }
    out[idx] = o;
  }
}

// uint8 RGB [N, H, W, 3] (what cv2 hands the reference's collate before ToTensor, dataset_utils/resnet_vqa_daquar_dataset.py:
// 153-171) -> the same stem layout; the /255 of transforms.ToTensor() is folded in (IEEE division, so the bf16 values equal
// those of the fp32 path bit for bit)
__global__ void image_u8_to_stem_kernel(const uint8_t* __restrict__ img, uint4* __restrict__ out, int N, int H, int W) {
  pdl_grid_sync();
  const int Wp = W + 8;
  const long long total = static_cast<long long>(N) * H * Wp;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const uint32_t i32 = static_cast<uint32_t>(idx);
    const uint32_t t = i32 / static_cast<uint32_t>(Wp);          // = n * H + h
    const int wp = static_cast<int>(i32 - t * static_cast<uint32_t>(Wp));
    uint4 o = make_uint4(0, 0, 0, 0);
    const int w = wp - 3;
    if (w >= 0 && w < W) {
      const uint8_t* p = img + (static_cast<long long>(t) * W + w) * 3;
      o.x = pack_bf16x2(static_cast<float>(p[0]) / 255.f, static_cast<float>(p[1]) / 255.f);
      o.y = pack_bf16x2(static_cast<float>(p[2]) / 255.f, 0.f);
    }
    out[idx] = o;
  }
}

// F.interpolate(mode="nearest") on bf16 NHWC (the FPN's top-down pathway, torchvision/ops/feature_pyramid_network.py):
// out[n, y, x, :] = in[n, floor(y * h / H), floor(x * w / W), :]; one thread per 8 channels
__global__ void upsample_nearest_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int N, int h, int w, int H,
                                        int W, int C8) {
  pdl_grid_sync();
  const long long total = static_cast<long long>(N) * H * W * C8;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int c = static_cast<int>(idx % C8);
    long long t = idx / C8;
    const int x = static_cast<int>(t % W); t /= W;
    const int y = static_cast<int>(t % H);
    const int n = static_cast<int>(t / H);
    const int sy = min(static_cast<int>((static_cast<long long>(y) * h) / H), h - 1);
    const int sx = min(static_cast<int>((static_cast<long long>(x) * w) / W), w - 1);
    out[idx] = in[((static_cast<long long>(n) * h + sy) * w + sx) * C8 + c];
  }
}

// x bf16 [N, HW, C] -> out fp32 [N, C, HW]
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int HW, int C) {
  pdl_grid_sync();
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int c0 = blockIdx.x * 32, p0 = blockIdx.y * 32;
  const __nv_bfloat16* xs = x + static_cast<long long>(n) * HW * C;
  float* os = out + static_cast<long long>(n) * HW * C;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int p = p0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (p < HW && c < C) ? __bfloat162float(xs[static_cast<long long>(p) * C + c]) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, p = p0 + threadIdx.x;
    if (c < C && p < HW) os[static_cast<long long>(c) * HW + p] = tile[threadIdx.x][j];
  }
}

// ---- max-pool 3x3 / stride 2 / pad 1, NHWC bf16, 8 channels per thread -------------------------
__global__ void maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                    int N, int H, int W, int C, int Ho, int Wo) {
  pdl_grid_sync();
  const int C8 = C >> 3;
  const long long total = static_cast<long long>(N) * Ho * Wo * C8;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const uint32_t i32 = static_cast<uint32_t>(idx);      // total < 2^31 (checked by the host)
    uint32_t t = i32 / static_cast<uint32_t>(C8);
    const int c8 = static_cast<int>(i32 - t * static_cast<uint32_t>(C8));
    uint32_t t2 = t / static_cast<uint32_t>(Wo);
    const int wo = static_cast<int>(t - t2 * static_cast<uint32_t>(Wo));
    const int n = static_cast<int>(t2 / static_cast<uint32_t>(Ho));
    const int ho = static_cast<int>(t2 - static_cast<uint32_t>(n) * static_cast<uint32_t>(Ho));
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int h = ho * 2 - 1 + dy;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int w = wo * 2 - 1 + dx;
        if (w < 0 || w >= W) continue;
        float f[8];
        load_bf16x8(x + ((static_cast<long long>(n) * H + h) * W + w) * C + c8 * 8, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = fmaxf(m[i], f[i]);
      }
    }
    store_bf16x8(out + idx * 8, m);
  }
}

// ---- dropout-mask cast: out_bf16 = dropmask(x) --------------------------------------------------
__global__ void dropout_cast_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long n8,
                                    float drop_p, uint32_t sid, const unsigned long long* __restrict__ rng) {
  pdl_grid_sync();
  const DropCtx dc = drop_ctx(drop_p, sid, rng);
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float f[8];
    load_f32x8(x + i * 8, f);
    drop8(dc, static_cast<unsigned long long>(i) * 8, f);
    store_bf16x8(out + i * 8, f);
  }
}

// ---- column sums: out[n] += sum_m x[m, n] ------------------------------------------------------
__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long ld, float* __restrict__ out,
                                   int M, int N) {
  pdl_grid_sync();
  __shared__ float red[8][32][9];
  const int cg = blockIdx.x * 32 + threadIdx.x;  // 8-column group
  const int n0 = cg * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (n0 < N) {
    const bool full = (n0 + 8 <= N) && ((ld & 7) == 0);
    for (int m = blockIdx.y * 8 + threadIdx.y; m < M; m += gridDim.y * 8) {
      const __nv_bfloat16* p = x + static_cast<long long>(m) * ld + n0;
      if (full) {
        float f[8];
        load_bf16x8(p, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += f[i];
      } else {
        for (int i = 0; i < 8 && n0 + i < N; ++i) acc[i] += __bfloat162float(p[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.y][threadIdx.x][i] = acc[i];
  __syncthreads();
  if (threadIdx.y == 0 && n0 < N) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float s = 0.f;
      for (int j = 0; j < 8; ++j) s += red[j][threadIdx.x][i];
      if (n0 + i < N) atomicAdd(out + n0 + i, s);
    }
  }
}

}  // namespace

extern "C" {

int vqa_cast_f32_bf16(void* plan, const float* src, void* dst, long long n, void* stream) {
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15)) {
    set_last_error("cast: pointers must be 16-byte aligned");
    return -1;
  }
  note_op("cast_f32_bf16", 0.0, 6.0 * static_cast<double>(n));
  return submit(plan, stream, [=](cudaStream_t s) {
    launch_pdl(cast_f32_bf16_kernel, dim3(grid_for((n >> 3) + 8, 256)), dim3(256), 0, s, src, static_cast<__nv_bfloat16*>(dst), n);
    return launch_status("cast_f32_bf16");
  });
}

int vqa_cast_bf16_f32(void* plan, const void* src, float* dst, long long n, void* stream) {
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15)) {
    set_last_error("cast: pointers must be 16-byte aligned");
    return -1;
  }
  note_op("cast_bf16_f32", 0.0, 6.0 * static_cast<double>(n));
  return submit(plan, stream, [=](cudaStream_t s) {
    launch_pdl(cast_bf16_f32_kernel, dim3(grid_for((n >> 3) + 8, 256)), dim3(256), 0, s, static_cast<const __nv_bfloat16*>(src), dst, n);
    return launch_status("cast_bf16_f32");
  });
}

int vqa_split_lo_bf16(void* plan, const float* src, void* dst, long long n, void* stream) {
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15)) {
    set_last_error("split_lo: pointers must be 16-byte aligned");
    return -1;
  }
  note_op("split_lo_bf16", 0.0, 6.0 * static_cast<double>(n));
  return submit(plan, stream, [=](cudaStream_t s) {
    launch_pdl(split_lo_bf16_kernel, dim3(grid_for((n >> 3) + 8, 256)), dim3(256), 0, s, src, static_cast<__nv_bfloat16*>(dst), n);
    return launch_status("split_lo_bf16");
  });
}

int vqa_memset_zero(void* plan, void* ptr, long long bytes, void* stream) {
  note_op("memset", 0.0, static_cast<double>(bytes));
  return submit(plan, stream, [=](cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(ptr, 0, static_cast<size_t>(bytes), s);
    if (e != cudaSuccess) { set_last_error("memset: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
    return 0;
  });
}

int vqa_memcpy_d2d(void* plan, void* dst, const void* src, long long bytes, void* stream) {
  note_op("memcpy", 0.0, 2.0 * static_cast<double>(bytes));
  return submit(plan, stream, [=](cudaStream_t s) {
    cudaError_t e = cudaMemcpyAsync(dst, src, static_cast<size_t>(bytes), cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) { set_last_error("memcpy: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
    return 0;
  });
}

int vqa_axpy_f32(void* plan, float* y, const float* x, float a, long long n, void* stream) {
  note_op("axpy", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    launch_pdl(axpy_kernel, dim3(grid_for((n >> 2) + 4, 256)), dim3(256), 0, s, y, x, a, n);
    return launch_status("axpy");
  });
}

int vqa_fold_conv_bn(void* plan, const float* w, const float* gamma, const float* beta, const float* mean,
                     const float* var, float eps, void* w_out, float* bias_out, int O, int I, int R, int S,
                     int Sp, int Ip, void* stream) {
  note_op("fold_conv_bn", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    const long long total = static_cast<long long>(O) * R * Sp * Ip;
    int grid = grid_for(total, 256);
    if (grid * 256 < O) grid = (O + 255) / 256;
    launch_pdl(fold_conv_bn_kernel, dim3(grid), dim3(256), 0, s, w, gamma, beta, mean, var, eps, static_cast<__nv_bfloat16*>(w_out),
                                             bias_out, O, I, R, S, Sp, Ip);
    return launch_status("fold_conv_bn");
  });
}

int vqa_convT_weight_prep(void* plan, const float* w, void* w_out, int Cin, int Cout, void* stream) {
  note_op("convT_weight_prep", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    dim3 grid((Cout * 9 + 31) / 32, (Cin + 31) / 32), block(32, 8);
    launch_pdl(transpose_flip9_kernel<__nv_bfloat16>, dim3(grid), dim3(block), 0, s, w, static_cast<__nv_bfloat16*>(w_out), Cin, Cout * 9);
    return launch_status("convT_weight_prep");
  });
}

int vqa_convT_weight_prep_bf16(void* plan, const void* w_bf16, void* w_out, int Cin, int Cout, void* stream) {
  note_op("convT_weight_prep", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    dim3 grid((Cout * 9 + 31) / 32, (Cin + 31) / 32), block(32, 8);
    launch_pdl(transpose_flip9_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(grid), dim3(block), 0, s,
               static_cast<const __nv_bfloat16*>(w_bf16), static_cast<__nv_bfloat16*>(w_out), Cin, Cout * 9);
    return launch_status("convT_weight_prep_bf16");
  });
}

int vqa_convT_wgrad_unprep(void* plan, const float* dw_conv, float* dw, int Cin, int Cout, void* stream) {
  // dw[ci, co*9 + t] = dw_conv[co*9 + 8 - t, ci]
  note_op("convT_wgrad_unprep", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    dim3 grid((Cin + 31) / 32, (Cout * 9 + 31) / 32), block(32, 8);
    launch_pdl(transpose_fliprow9_kernel, dim3(grid), dim3(block), 0, s, dw_conv, dw, Cout * 9, Cin);
    return launch_status("convT_wgrad_unprep");
  });
}

int vqa_image_to_stem(void* plan, const float* img, void* out, int N, int H, int W, void* stream) {
  note_op("image_to_stem", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    const long long total = static_cast<long long>(N) * H * (W + 8);
    if (total >= (1LL << 31)) { set_last_error("image_to_stem: more than 2^31 pixels"); return -1; }
    launch_pdl(image_to_stem_kernel, dim3(grid_for(total, 256, 16)), dim3(256), 0, s, img, static_cast<uint4*>(out), N, H, W);
    return launch_status("image_to_stem");
  });
}

int vqa_image_u8_to_stem(void* plan, const uint8_t* img, void* out, int N, int H, int W, void* stream) {
  note_op("image_to_stem", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    const long long total = static_cast<long long>(N) * H * (W + 8);
    if (total >= (1LL << 31)) { set_last_error("image_u8_to_stem: more than 2^31 pixels"); return -1; }
    launch_pdl(image_u8_to_stem_kernel, dim3(grid_for(total, 256, 16)), dim3(256), 0, s, img, static_cast<uint4*>(out), N, H, W);
    return launch_status("image_u8_to_stem");
  });
}

int vqa_upsample_nearest_nhwc(void* plan, const void* x, void* out, int N, int h, int w, int H, int W, int C, void* stream) {
  if (C % 8 || h < 1 || w < 1 || H < 1 || W < 1) { set_last_error("upsample_nearest: C must be a multiple of 8"); return -1; }
  note_op("upsample_nearest", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    const long long total = static_cast<long long>(N) * H * W * (C / 8);
    launch_pdl(upsample_nearest_kernel, dim3(grid_for(total, 256, 16)), dim3(256), 0, s, static_cast<const uint4*>(x),
               static_cast<uint4*>(out), N, h, w, H, W, C / 8);
    return launch_status("upsample_nearest");
  });
}

int vqa_nhwc_to_nchw_f32(void* plan, const void* x, float* out, int N, int H, int W, int C, void* stream) {
  note_op("nhwc_to_nchw", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    dim3 grid((C + 31) / 32, (H * W + 31) / 32, N), block(32, 8);
    launch_pdl(nhwc_to_nchw_kernel, dim3(grid), dim3(block), 0, s, static_cast<const __nv_bfloat16*>(x), out, H * W, C);
    return launch_status("nhwc_to_nchw");
  });
}

int vqa_maxpool3x3s2(void* plan, const void* x, void* out, int N, int H, int W, int C, void* stream) {
  if (C % 8) { set_last_error("maxpool: C must be a multiple of 8"); return -1; }
  note_op("maxpool3x3s2", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    const long long total = static_cast<long long>(N) * Ho * Wo * (C / 8);
    if (total >= (1LL << 31)) { set_last_error("maxpool: more than 2^31 output vectors"); return -1; }
    launch_pdl(maxpool3x3s2_kernel, dim3(grid_for(total, 256, 16)), dim3(256), 0, s, static_cast<const __nv_bfloat16*>(x),
                                                                   static_cast<__nv_bfloat16*>(out), N, H, W, C, Ho, Wo);
    return launch_status("maxpool3x3s2");
  });
}

int vqa_dropout_cast(void* plan, const float* x, void* out_bf16, long long rows, int N, float drop_p, uint32_t sid,
                     const uint64_t* rng, void* stream) {
  if (N % 8) { set_last_error("dropout_cast: N must be a multiple of 8"); return -1; }
  note_op("dropout_cast", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    const long long n8 = rows * N / 8;
    launch_pdl(dropout_cast_kernel, dim3(grid_for(n8, 256)), dim3(256), 0, s, x, static_cast<__nv_bfloat16*>(out_bf16), n8, drop_p, sid,
                                                          reinterpret_cast<const unsigned long long*>(rng));
    return launch_status("dropout_cast");
  });
}

int vqa_colsum_bf16(void* plan, const void* x, long long ld, float* out, int M, int N, void* stream) {
  note_op("colsum_bf16", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    dim3 block(32, 8);
    int gy = (M + 63) / 64;
    if (gy > 64) gy = 64;
    if (gy < 1) gy = 1;
    dim3 grid((N + 255) / 256, gy);
    launch_pdl(colsum_bf16_kernel, dim3(grid), dim3(block), 0, s, static_cast<const __nv_bfloat16*>(x), ld, out, M, N);
    return launch_status("colsum_bf16");
  });
}

}  // extern "C"
