// Micro-benchmarks of the primitives in the GEMM k-loop (one CTA): cycles per operation.  Diagnostic only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../t5-resnet-vqa_b200/csrc handshake.cu -o handshake
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace vqa;

__global__ void __launch_bounds__(384, 1) k(long long* out, int iters, int nmma_n, int spin_mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = base + 200 * 1024, bar1 = bar0 + 8, bar2 = bar0 + 16, bar3 = bar0 + 24, slot = bar0 + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar0, 1); mbar_init(bar1, 1); mbar_init(bar2, 1); mbar_init(bar3, 1); mbar_fence_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
  long long t0, t1;
  // bystanders: 8 warps waiting for a barrier that completes only when warp 0 is done (like the epilogue warps)
  if (warp >= 4) {
    if (spin_mode == 1) mbar_wait(bar3, 0);            // bounded wait with clock reads (epilogue's mbar_wait)
    else if (spin_mode == 2) mbar_wait_lean(bar3, 0);
    else if (spin_mode == 3) { if (threadIdx.x == 128) mbar_wait_lean(bar3, 0); named_bar_sync(1, 256); }
  }
  if (warp == 0) {
    // 1. arrive + wait on own barrier (whole warp waits, elected lane arrives)
    uint32_t ph = 0;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { mbar_arrive_e(bar0); mbar_wait_lean(bar0, ph); ph ^= 1; }
    t1 = clock64();
    if (lane == 0) out[0] = (t1 - t0) / iters;
    // 2. commit (no MMAs outstanding) + wait
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { umma_commit_e<1>(bar0); mbar_wait_lean(bar0, ph); ph ^= 1; }
    t1 = clock64();
    if (lane == 0) out[1] = (t1 - t0) / iters;
    // 3. try_wait that succeeds immediately (phase already complete: wait on the previous parity)
    t0 = clock64();
    for (int i = 0; i < iters; ++i) mbar_wait_lean(bar0, ph ^ 1);
    t1 = clock64();
    if (lane == 0) out[2] = (t1 - t0) / iters;
    // 4. 4 MMAs (128 x N x 16) + commit + wait: completion latency of a k-block
    const uint32_t idesc = umma_idesc_bf16(128, nmma_n, false, false);
    const uint64_t da = umma_smem_desc(base, 16, 1024), db = umma_smem_desc(base + 16384, 16, 1024);
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) umma_bf16_e<1>(tmem_base, da + 2 * ks, db + 2 * ks, idesc, 1u);
      umma_commit_e<1>(bar0); mbar_wait_lean(bar0, ph); ph ^= 1;
    }
    t1 = clock64();
    if (lane == 0) out[3] = (t1 - t0) / iters;
    // 5. 16 k-blocks of 4 MMAs back to back, one commit + wait at the end: MMA throughput
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll 1
      for (int kb = 0; kb < 16; ++kb) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_bf16_e<1>(tmem_base, da + 2 * ks, db + 2 * ks, idesc, 1u);
      }
      umma_commit_e<1>(bar0); mbar_wait_lean(bar0, ph); ph ^= 1;
    }
    t1 = clock64();
    if (lane == 0) out[4] = (t1 - t0) / iters / 16;
    // 6. same with a commit after every k-block (to different barriers nobody waits for except the last)
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll 1
      for (int kb = 0; kb < 16; ++kb) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_bf16_e<1>(tmem_base, da + 2 * ks, db + 2 * ks, idesc, 1u);
        if (kb < 15) umma_commit_e<1>(bar2);   // count-1 barrier: completes a phase each time, nobody waits
      }
      umma_commit_e<1>(bar0); mbar_wait_lean(bar0, ph); ph ^= 1;
    }
    t1 = clock64();
    if (lane == 0) out[5] = (t1 - t0) / iters / 16;
  }
  if (warp == 0) mbar_arrive_e(bar3);
  __syncthreads();
  // 7. cross-warp ping-pong: warp 1 arrives on bar1 when it sees bar0 and vice versa
  if (warp == 0) {
    uint32_t ph = 0;
    // re-sync parity of bar0: read phase by brute force (bar0 phase parity unknown) -> use bar1/bar2 freshly... skipped
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64 * 8); cudaMemset(d, 0, 64 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  for (int mode = 0; mode < 4; ++mode)
  for (int n : {128}) {
    k<<<1, 384, 220 * 1024>>>(d, 200, n, mode);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("bystanders=%d N=%d err=%d | arrive+wait %lld | commit+wait %lld | passing wait %lld | 4 MMA+commit+wait %lld | kb throughput %lld cyc (no commits) | %lld cyc (commit per kb)\n",
           mode, n, (int)e, h[0], h[1], h[2], h[3], h[4], h[5]);
  }
  return 0;
}
