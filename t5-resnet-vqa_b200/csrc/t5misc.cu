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

// ---- data-parallel embedding gradient: rows travel, every rank scatters them in the same fixed order ----------------
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

// first[id] = smallest token index carrying that id (atomicMin: the result does not depend on the execution order)
__global__ void embedding_first_kernel(const long long* __restrict__ ids, int* __restrict__ first, int T, int vocab) {
  pdl_grid_sync();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < T) atomicMin(first + clamp_id(ids[t], vocab), t);
}

// One warp per token t.  The FIRST token of each id owns that id's row of the table: it walks the later tokens in
// index order and adds the rows of those with the same id, so the sum is formed in one fixed order on every rank
// (bit-identical replicas; fp32 atomics would add duplicates in arrival order), then stores the row (no atomics: one
// owner per id; untouched rows keep the zeros the backward pass wrote).
__global__ void embedding_scatter_ordered_kernel(const long long* __restrict__ ids, const float* __restrict__ rows,
                                                 const int* __restrict__ first, float* __restrict__ dtable, int T, int D,
                                                 int vocab) {
  pdl_grid_sync();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int t = warp; t < T; t += nwarps) {
    const int id = clamp_id(ids[t], vocab);
    if (first[id] != t) continue;            // warp-uniform
    float acc[32];                            // D <= 1024: lane owns columns lane*8 + 256*j
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    for (int u0 = t; u0 < T; u0 += 32) {
      const int u = u0 + lane;
      const bool hit = u < T && clamp_id(ids[u], vocab) == id;
      unsigned m = __ballot_sync(0xffffffffu, hit);
      while (m) {                              // ascending token index
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const float* src = rows + static_cast<long long>(u0 + b) * D;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = lane * 8 + 256 * j;
          if (col < D) {
            float f[8];
            load_f32x8(src + col, f);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[8 * j + i] += f[i];
          }
        }
      }
    }
    float* dst = dtable + static_cast<long long>(id) * D;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = lane * 8 + 256 * j;
      if (col < D) store_f32x8(dst + col, *reinterpret_cast<float(*)[8]>(&acc[8 * j]));
    }
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

int vqa_embedding_scatter_ordered(void* plan, const long long* ids, const float* rows, float* dtable, int* first_ws, int T,
                                  int D, int vocab, void* stream) {
  if (D % 8 || D > 1024) { set_last_error("embedding_scatter_ordered: D must be a multiple of 8, at most 1024"); return -1; }
  note_op("embedding_scatter", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(first_ws, 0x7f, sizeof(int) * static_cast<size_t>(vocab), s);   // 0x7f7f7f7f > any index
    if (e != cudaSuccess) { set_last_error("embedding_scatter_ordered: %s", cudaGetErrorString(e)); return static_cast<int>(e); }
    launch_pdl(embedding_first_kernel, dim3((T + 255) / 256), dim3(256), 0, s, ids, first_ws, T, vocab);
    int grid = (T + 7) / 8;
    if (grid > 148 * 8) grid = 148 * 8;
    launch_pdl(embedding_scatter_ordered_kernel, dim3(grid), dim3(256), 0, s, ids, rows, static_cast<const int*>(first_ws),
               dtable, T, D, vocab);
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
