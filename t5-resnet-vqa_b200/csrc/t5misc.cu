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
