// T5 encoder odds and ends: token-embedding gather / scatter-add (hf:682,734) and the shared
// relative-position bias table expansion and its gradient (hf:236-251).
#include "../../include/vqa_b200.h"
#include "common.cuh"

using namespace vqa;

namespace {

// one warp per token row; D % 8 == 0; each lane moves 8 floats per iteration
__global__ void embedding_fwd_kernel(const long long* __restrict__ ids, const float* __restrict__ table,
                                     float* __restrict__ out, int M, int D, int vocab, float drop_p, uint32_t sid,
                                     const unsigned long long* __restrict__ rng) {
  pdl_grid_sync();
  const DropCtx dc = drop_ctx(drop_p, sid, rng);
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int row = warp; row < M; row += nwarps) {
    long long id = ids[row];
    if (id < 0) id = 0;
    if (id >= vocab) id = vocab - 1;
    const float* src = table + id * D;
    for (int col = lane * 8; col < D; col += 256) {
      float f[8];
      load_f32x8(src + col, f);
      drop8(dc, static_cast<unsigned long long>(row) * D + col, f);
      store_f32x8(out + static_cast<long long>(row) * D + col, f);
    }
  }
}

__global__ void embedding_bwd_kernel(const long long* __restrict__ ids, const float* __restrict__ dout,
                                     float* __restrict__ dtable, int M, int D, int vocab, float drop_p,
                                     uint32_t sid, const unsigned long long* __restrict__ rng) {
  pdl_grid_sync();
  const DropCtx dc = drop_ctx(drop_p, sid, rng);
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int row = warp; row < M; row += nwarps) {
    long long id = ids[row];
    if (id < 0) id = 0;
    if (id >= vocab) id = vocab - 1;
    float* dst = dtable + id * D;
    for (int col = lane * 8; col < D; col += 256) {
      float f[8];
      load_f32x8(dout + static_cast<long long>(row) * D + col, f);
      drop8(dc, static_cast<unsigned long long>(row) * D + col, f);
      atomicAdd(reinterpret_cast<float4*>(dst + col), make_float4(f[0], f[1], f[2], f[3]));
      atomicAdd(reinterpret_cast<float4*>(dst + col + 4), make_float4(f[4], f[5], f[6], f[7]));
    }
  }
}

// ---- data-parallel embedding gradient: rows travel, every rank scatters them with an order-independent sum ---------
// rows[t, :] = scale * dropout_mask(t, :) * dout[t, :]   (what embedding_bwd would have added to dtable[ids[t], :])
__global__ void embedding_bwd_rows_kernel(const float* __restrict__ dout, float* __restrict__ rows, int M, int D, float drop_p,
                                          uint32_t sid, const unsigned long long* __restrict__ rng, float scale) {
  pdl_grid_sync();
  const DropCtx dc = drop_ctx(drop_p, sid, rng);
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int row = warp; row < M; row += nwarps) {
    for (int col = lane * 8; col < D; col += 256) {
      float f[8];
      load_f32x8(dout + static_cast<long long>(row) * D + col, f);
      drop8(dc, static_cast<unsigned long long>(row) * D + col, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] *= scale;
      store_f32x8(rows + static_cast<long long>(row) * D + col, f);
    }
  }
}

__device__ __forceinline__ int clamp_id(long long id, int vocab) {
  return id < 0 ? 0 : (id >= vocab ? vocab - 1 : static_cast<int>(id));
}

// ---- order-independent scatter: identical on every rank, whatever order the hardware adds in -----------------
// Pass 1: first[id] = smallest token index carrying that id (atomicMin) and count[id] (integer atomicAdd): both exact.
__global__ void embedding_first_kernel(const long long* __restrict__ ids, int* __restrict__ first, int* __restrict__ count,
                                       int T, int vocab) {
  pdl_grid_sync();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < T) {
    const int id = clamp_id(ids[t], vocab);
    atomicMin(first + id, t);
    atomicAdd(count + id, 1);
  }
}

constexpr float kFixScale = 1099511627776.f;          // 2^40: fixed-point unit 9.1e-13, |sum| up to 2^23
constexpr float kFixInv = 1.f / 1099511627776.f;

// Pass 2 (warp per token, eight consecutive tokens per CTA).  An id carried by ONE token: its row is copied into the table.
// An id carried by several tokens (padding, frequent words): the rows are added as 64-bit FIXED-POINT integers into the
// accumulator row of the id's first token - integer addition is associative, so thousands of concurrent atomics give the
// same bits on every rank and in every run (fp32 atomics would not).  Tokens of one CTA that share an id (runs of padding)
// are first summed in shared memory, so a run of eight costs one atomic per column instead of eight.
// mode 0: first tokens of shared ids clear their accumulator row; 1: add; 2: first tokens convert the sum into the table row.
constexpr int kScatWarps = 8;
__global__ void __launch_bounds__(kScatWarps * 32)
embedding_scatter_fixed_kernel(const long long* __restrict__ ids, const float* __restrict__ rows,
                               const int* __restrict__ first, const int* __restrict__ count,
                               long long* __restrict__ acc, float* __restrict__ dtable, int T, int D, int vocab, int mode) {
  pdl_grid_sync();
  extern __shared__ __align__(16) long long sfix[];        // mode 1: [kScatWarps][D] fixed-point rows of this CTA's tokens
  __shared__ int skey[kScatWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int base = blockIdx.x * kScatWarps; base < T; base += gridDim.x * kScatWarps) {   // trip count uniform per CTA
    const int t = base + warp;
    const bool valid = t < T;
    const int id = valid ? clamp_id(ids[t], vocab) : 0;
    const int c = valid ? count[id] : 0, f = valid ? first[id] : -1;
    if (mode != 1) {
      if (!valid || c == 1) continue;
      long long* arow = acc + static_cast<long long>(f) * D;
      if (f != t) continue;
      if (mode == 0) {
        for (int col = lane * 2; col < D; col += 64) *reinterpret_cast<longlong2*>(arow + col) = make_longlong2(0ll, 0ll);
      } else {
        float* dst = dtable + static_cast<long long>(id) * D;
        for (int col = lane * 8; col < D; col += 256) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __ll2float_rn(arow[col + i]) * kFixInv;
          store_f32x8(dst + col, v);
        }
      }
      continue;
    }
    // ---- mode 1 ----
    const bool shared_id = valid && c > 1;
    if (lane == 0) skey[warp] = shared_id ? f : -1 - warp;     // unique negative keys never match
    if (valid && c == 1) {
      const float* src = rows + static_cast<long long>(t) * D;
      float* dst = dtable + static_cast<long long>(id) * D;
      for (int col = lane * 8; col < D; col += 256) {
        float v[8];
        load_f32x8(src + col, v);
        store_f32x8(dst + col, v);
      }
    } else if (shared_id) {
      const float* src = rows + static_cast<long long>(t) * D;
      long long* srow = sfix + static_cast<long long>(warp) * D;
      for (int col = lane * 8; col < D; col += 256) {
        float v[8];
        load_f32x8(src + col, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) srow[col + i] = __float2ll_rn(v[i] * kFixScale);
      }
    }
    __syncthreads();
    if (shared_id) {
      bool leader = true;                       // the first warp of the CTA holding this id adds for all of them
      for (int w = 0; w < warp; ++w) leader = leader && skey[w] != f;
      if (leader) {
        long long* arow = acc + static_cast<long long>(f) * D;
        for (int col = lane; col < D; col += 32) {
          long long sum = sfix[static_cast<long long>(warp) * D + col];
          for (int w = warp + 1; w < kScatWarps; ++w)
            if (skey[w] == f) sum += sfix[static_cast<long long>(w) * D + col];
          atomicAdd(reinterpret_cast<unsigned long long*>(arow + col), static_cast<unsigned long long>(sum));
        }
      }
    }
    __syncthreads();
  }
}

__global__ void t5_bias_build_kernel(const float* __restrict__ table, const int* __restrict__ bucket,
                                     float* __restrict__ bias, int H, int LL) {
  pdl_grid_sync();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * LL) return;
  const int h = idx / LL, ij = idx - h * LL;
  bias[idx] = table[bucket[ij] * H + h];
}

__global__ void t5_bias_grad_kernel(const float* __restrict__ dbias, const int* __restrict__ bucket,
                                    float* __restrict__ dtable, int H, int LL) {
  pdl_grid_sync();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= H * LL) return;
  const int h = idx / LL, ij = idx - h * LL;
  atomicAdd(dtable + bucket[ij] * H + h, dbias[idx]);
}

}  // namespace

extern "C" {

int vqa_embedding_fwd(void* plan, const long long* ids, const float* table, float* out, int M, int D, int vocab,
                      float drop_p, uint32_t sid, const uint64_t* rng, void* stream) {
  if (D % 8) { set_last_error("embedding: D must be a multiple of 8"); return -1; }
  note_op("embedding_fwd", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    int grid = (M + 7) / 8;
    if (grid > 148 * 8) grid = 148 * 8;
    launch_pdl(embedding_fwd_kernel, dim3(grid), dim3(256), 0, s, ids, table, out, M, D, vocab, drop_p, sid,
                                              reinterpret_cast<const unsigned long long*>(rng));
    return launch_status("embedding_fwd");
  });
}

int vqa_embedding_bwd(void* plan, const long long* ids, const float* dout, float* dtable, int M, int D, int vocab,
                      float drop_p, uint32_t sid, const uint64_t* rng, void* stream) {
  if (D % 8) { set_last_error("embedding: D must be a multiple of 8"); return -1; }
  note_op("embedding_bwd", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    int grid = (M + 7) / 8;
    if (grid > 148 * 8) grid = 148 * 8;
    launch_pdl(embedding_bwd_kernel, dim3(grid), dim3(256), 0, s, ids, dout, dtable, M, D, vocab, drop_p, sid,
                                              reinterpret_cast<const unsigned long long*>(rng));
    return launch_status("embedding_bwd");
  });
}

int vqa_embedding_bwd_rows(void* plan, const float* dout, float* rows, int M, int D, float drop_p, uint32_t sid,
                           const uint64_t* rng, float scale, void* stream) {
  if (D % 8) { set_last_error("embedding: D must be a multiple of 8"); return -1; }
  note_op("embedding_bwd", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    int grid = (M + 7) / 8;
    if (grid > 148 * 8) grid = 148 * 8;
    launch_pdl(embedding_bwd_rows_kernel, dim3(grid), dim3(256), 0, s, dout, rows, M, D, drop_p, sid,
               reinterpret_cast<const unsigned long long*>(rng), scale);
    return launch_status("embedding_bwd_rows");
  });
}

int vqa_embedding_scatter_ordered(void* plan, const long long* ids, const float* rows, float* dtable, int* first_ws,
                                  long long* acc_ws, int T, int D, int vocab, void* stream) {
  if (D % 8 || D > 1024 || (reinterpret_cast<uintptr_t>(acc_ws) & 15)) {
    set_last_error("embedding_scatter_ordered: D must be a multiple of 8, at most 1024, and the workspace 16-byte aligned");
    return -1;
  }
  note_op("embedding_scatter", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    int* count_ws = first_ws + vocab;
    cudaError_t e = cudaMemsetAsync(first_ws, 0x7f, sizeof(int) * static_cast<size_t>(vocab), s);   // 0x7f7f7f7f > any index
    if (e == cudaSuccess) e = cudaMemsetAsync(count_ws, 0, sizeof(int) * static_cast<size_t>(vocab), s);
    if (e != cudaSuccess) { set_last_error("embedding_scatter_ordered: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
    launch_pdl(embedding_first_kernel, dim3((T + 255) / 256), dim3(256), 0, s, ids, first_ws, count_ws, T, vocab);
    int grid = (T + 7) / 8;
    if (grid > 148 * 8) grid = 148 * 8;
    const size_t smem = static_cast<size_t>(kScatWarps) * D * sizeof(long long);
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(embedding_scatter_fixed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 1024 * 8);
      attr = true;
    }
    for (int mode = 0; mode < 3; ++mode)
      launch_pdl(embedding_scatter_fixed_kernel, dim3(grid), dim3(kScatWarps * 32), mode == 1 ? smem : 0, s, ids, rows,
                 static_cast<const int*>(first_ws), static_cast<const int*>(count_ws), acc_ws, dtable, T, D, vocab, mode);
    return launch_status("embedding_scatter_ordered");
  });
}

int vqa_t5_bias_build(void* plan, const float* table, const int* bucket, float* bias, int H, int L, int nbuckets,
                      void* stream) {
  (void)nbuckets;
  note_op("t5_bias_build", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    const int total = H * L * L;
    launch_pdl(t5_bias_build_kernel, dim3((total + 255) / 256), dim3(256), 0, s, table, bucket, bias, H, L * L);
    return launch_status("t5_bias_build");
  });
}

int vqa_t5_bias_grad(void* plan, const float* dbias, const int* bucket, float* dtable, int H, int L, int nbuckets,
                     void* stream) {
  (void)nbuckets;
  note_op("t5_bias_grad", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    const int total = H * L * L;
    launch_pdl(t5_bias_grad_kernel, dim3((total + 255) / 256), dim3(256), 0, s, dbias, bucket, dtable, H, L * L);
    return launch_status("t5_bias_grad");
  });
}

}  // extern "C"
