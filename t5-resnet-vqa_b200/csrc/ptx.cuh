// sm_100a PTX wrappers: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld),
// UMMA shared-memory + instruction descriptors.  Hand-written inline PTX; no CUTLASS dependency.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace vqa {

// ---------------------------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}
// Bounded wait: a mis-programmed pipeline traps (-> launch error the host reports) instead of
// hanging the GPU.  ~4e9 cycles is >2 s at any B200 clock; no legitimate wait is that long.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("vqa: mbarrier timeout block(%d,%d,%d) thread %d bar 0x%x parity %u\n", blockIdx.x,
             blockIdx.y, blockIdx.z, threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// Hot-loop variant (producer / MMA issuer): no clock reads in the spin; try_wait suspends in hardware between polls,
// so an iteration bound is an equally good watchdog and costs one add.
__device__ __forceinline__ void mbar_wait_lean(uint32_t bar, uint32_t parity) {
  uint32_t n = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++n > (1u << 27)) __trap();
  }
}

// ---------------------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];\n" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4, %5, %6}], [%2];\n" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// shared -> global tensor stores (bulk async-group completion); the tensor map clips out-of-bounds parts of the box
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// global[box] += shared[box] (element type of the tensor map: fp32)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
// all but the newest N bulk groups of this thread have finished READING shared memory
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;\n" ::"n"(N) : "memory");
}
// generic-proxy shared-memory writes -> visible to the async proxy (TMA stores, UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}
// whole warp; writes the TMEM base address into *smem_dst
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                   bar)
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets lane (base_lane + t)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): two CTAs of a cluster on the two SMs of one TPC run ONE 256-row MMA.  Each CTA stages
// its own 128 rows of A and HALF of the B tile; the leader (cluster rank 0) issues the MMA, which reads both
// CTAs' shared memory at identical offsets and writes 128 accumulator lanes into each CTA's TMEM.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// arrive on an mbarrier that lives in another CTA of the cluster (address from mapa_shared)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(remote_bar) : "memory");
}
// wait for a phase whose arrivals come from other CTAs of the cluster (pairs with mbar_arrive_cluster's release)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, n = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && ++n > (1u << 27)) __trap();
  }
}
// TMA loads of a CTA pair: data lands in THIS CTA's shared memory, the byte count is credited to `bar`, which may
// live in the peer (leader) CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];\n" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4, %5, %6}], [%2];\n" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at the same offset in every CTA of `cta_mask` once the pair's MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(bar),
      "h"(cta_mask)
      : "memory");
}


// ---------------------------------------------------------------------------------------------
// "_e" (elected) variants for the GEMM hot loops: called by a whole CONVERGED warp with warp-uniform operands; one
// elected lane issues.  Keeping the election inside the asm block (instead of an `if (lane == 0)` region around it)
// lets ptxas keep the operands in uniform registers without the per-instruction ELECT / R2UR.BROADCAST waterfall loop
// it emits for code it has to treat as divergent.
// ---------------------------------------------------------------------------------------------
#define VQA_ELECT_BEGIN "{\n\t.reg .pred pe;\n\telect.sync _|pe, 0xffffffff;\n\t"
#define VQA_ELECT_END "\n\t}\n"
__device__ __forceinline__ void mbar_expect_tx_e(uint32_t bar, uint32_t bytes) {
  asm volatile(VQA_ELECT_BEGIN "@pe mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" VQA_ELECT_END ::"r"(bar),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_e(uint32_t bar) {
  asm volatile(VQA_ELECT_BEGIN "@pe mbarrier.arrive.shared::cta.b64 _, [%0];" VQA_ELECT_END ::"r"(bar) : "memory");
}
template <int CTAS>
__device__ __forceinline__ void tma_load_2d_e(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  if (CTAS == 2) {
    asm volatile(VQA_ELECT_BEGIN
                 "@pe cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
                 "{%3, %4}], [%2];" VQA_ELECT_END ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
  } else {
    asm volatile(VQA_ELECT_BEGIN
                 "@pe cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
                 "[%2];" VQA_ELECT_END ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
  }
}
template <int CTAS>
__device__ __forceinline__ void tma_load_4d_e(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                              int c3) {
  if (CTAS == 2) {
    asm volatile(VQA_ELECT_BEGIN
                 "@pe cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
                 "{%3, %4, %5, %6}], [%2];" VQA_ELECT_END ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
  } else {
    asm volatile(VQA_ELECT_BEGIN
                 "@pe cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
                 "%6}], [%2];" VQA_ELECT_END ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
  }
}
// accumulate = 0 overwrites the accumulator (first MMA of a tile)
template <int CTAS>
__device__ __forceinline__ void umma_bf16_e(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  if (CTAS == 2) {
    asm volatile(VQA_ELECT_BEGIN
                 ".reg .pred pa;\n\tsetp.ne.b32 pa, %4, 0;\n\t"
                 "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, pa;" VQA_ELECT_END ::"r"(tmem_d),
                 "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
                 : "memory");
  } else {
    asm volatile(VQA_ELECT_BEGIN
                 ".reg .pred pa;\n\tsetp.ne.b32 pa, %4, 0;\n\t"
                 "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pa;" VQA_ELECT_END ::"r"(tmem_d),
                 "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
                 : "memory");
  }
}
template <int CTAS>
__device__ __forceinline__ void umma_commit_e(uint32_t bar) {
  if (CTAS == 2) {
    asm volatile(VQA_ELECT_BEGIN
                 "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 VQA_ELECT_END ::"r"(bar),
                 "h"(static_cast<uint16_t>(3))
                 : "memory");
  } else {
    asm volatile(VQA_ELECT_BEGIN "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 VQA_ELECT_END ::"r"(bar)
                 : "memory");
  }
}

// One k-block of the GEMM main loop as a single instruction group: a NON-blocking poll of the next stage's full
// barrier is issued first, then the four MMAs (K = 16 each) and the commit that releases the stage; the poll's result
// is read last, so its latency (a passing mbarrier wait costs ~140 cycles on its own, tools/ubench/handshake.cu)
// overlaps the MMA issue instead of preceding it.  Returns 1 when the next stage is already full.
// Whole converged warp, warp-uniform operands; descriptors advance by (a_step, b_step) >> 4 encoded bytes per MMA.
template <int CTAS>
__device__ __forceinline__ uint32_t umma_kblock4_e(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t a_step,
                                                   uint32_t b_step, uint32_t idesc, uint32_t accumulate,
                                                   uint32_t empty_bar, uint32_t next_full_bar, uint32_t next_parity) {
  uint32_t ready;
  // the three further descriptor pairs are derived inside the block (uniform adds), so only the base pair and the
  // steps have to be moved into uniform registers per k-block
  if (CTAS == 2) {
    asm volatile(
        "{\n\t.reg .pred pe, pa, pt, pw;\n\t.reg .b64 a1, a2, a3, b1, b2, b3, sa, sb;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 pw, [%9], %10;\n\t"
        "cvt.u64.u32 sa, %4;\n\tcvt.u64.u32 sb, %5;\n\t"
        "add.u64 a1, %2, sa;\n\tadd.u64 a2, a1, sa;\n\tadd.u64 a3, a2, sa;\n\t"
        "add.u64 b1, %3, sb;\n\tadd.u64 b2, b1, sb;\n\tadd.u64 b3, b2, sb;\n\t"
        "setp.ne.b32 pa, %7, 0;\n\t"
        "setp.eq.u32 pt, 0, 0;\n\t"
        "@pe tcgen05.mma.cta_group::2.kind::f16 [%1], %2, %3, %6, pa;\n\t"
        "@pe tcgen05.mma.cta_group::2.kind::f16 [%1], a1, b1, %6, pt;\n\t"
        "@pe tcgen05.mma.cta_group::2.kind::f16 [%1], a2, b2, %6, pt;\n\t"
        "@pe tcgen05.mma.cta_group::2.kind::f16 [%1], a3, b3, %6, pt;\n\t"
        "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%8], %11;\n\t"
        "selp.u32 %0, 1, 0, pw;\n\t}\n"
        : "=r"(ready)
        : "r"(tmem_d), "l"(da), "l"(db), "r"(a_step), "r"(b_step), "r"(idesc), "r"(accumulate), "r"(empty_bar),
          "r"(next_full_bar), "r"(next_parity), "h"(static_cast<uint16_t>(3))
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred pe, pa, pt, pw;\n\t.reg .b64 a1, a2, a3, b1, b2, b3, sa, sb;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 pw, [%9], %10;\n\t"
        "cvt.u64.u32 sa, %4;\n\tcvt.u64.u32 sb, %5;\n\t"
        "add.u64 a1, %2, sa;\n\tadd.u64 a2, a1, sa;\n\tadd.u64 a3, a2, sa;\n\t"
        "add.u64 b1, %3, sb;\n\tadd.u64 b2, b1, sb;\n\tadd.u64 b3, b2, sb;\n\t"
        "setp.ne.b32 pa, %7, 0;\n\t"
        "setp.eq.u32 pt, 0, 0;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %3, %6, pa;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%1], a1, b1, %6, pt;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%1], a2, b2, %6, pt;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%1], a3, b3, %6, pt;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
        "selp.u32 %0, 1, 0, pw;\n\t}\n"
        : "=r"(ready)
        : "r"(tmem_d), "l"(da), "l"(db), "r"(a_step), "r"(b_step), "r"(idesc), "r"(accumulate), "r"(empty_bar),
          "r"(next_full_bar), "r"(next_parity)
        : "memory");
  }
  return ready;
}

// ---------------------------------------------------------------------------------------------
// UMMA descriptors (bit layout: PTX ISA "tcgen05 shared memory descriptor" / "instruction
// descriptor"; 128-byte swizzle only)
//   K-major  tile: rows = M/N index, 128 B (64 bf16 of K) per row, 8-row swizzle atoms 1024 B apart.
//   MN-major tile: rows = K index, 128 B (64 bf16 of M/N) per row, 8-row atoms 1024 B apart (SBO),
//                  next 64-wide M/N chunk `lbo_bytes` further (LBO).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);        // [0,14)  start address >> 4
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;   // [16,30) leading byte offset >> 4
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;   // [32,46) stride byte offset >> 4
  d |= static_cast<uint64_t>(1) << 46;                           // [46,48) descriptor version (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                           // [61,64) SWIZZLE_128B
  return d;
}
// bf16 x bf16 -> fp32, dense; a_mn / b_mn = operand is MN-major ("transposed") in shared memory
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4)                         // [4,6)   D format  = F32
         | (1u << 7)                       // [7,10)  A format  = BF16
         | (1u << 10)                      // [10,13) B format  = BF16
         | ((a_mn ? 1u : 0u) << 15)        // [15]    A major   (1 = MN-major)
         | ((b_mn ? 1u : 0u) << 16)        // [16]    B major
         | (static_cast<uint32_t>(N >> 3) << 17)   // [17,23) N >> 3
         | (static_cast<uint32_t>(M >> 4) << 24);  // [24,29) M >> 4
}

}  // namespace vqa
