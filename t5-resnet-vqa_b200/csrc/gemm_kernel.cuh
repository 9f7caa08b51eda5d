// tcgen05 / TMEM / TMA GEMM for sm_100a:  D[M,N] = epilogue( A[M,K] * B[N,K]^T ), bf16 in, fp32 accumulate.
//
// Persistent kernel, one CTA per SM, 384 threads (12 warps), static round-robin tile scheduler over 128 x BN output
// tiles (n-tile fastest so CTAs running side by side share the A tile in L2; split-K slices are extra tiles).  The warp
// index is broadcast from lane 0, which makes the role branches warp-uniform for ptxas: the producer and MMA loops are run
// by the WHOLE warp with their state in uniform registers, and one elected lane issues from inside the asm blocks.
//   warp 0      A-operand TMA producer (STAGES-deep ring of 128x64 A and BNx64 B tiles, runs ahead across tiles;
//                               coordinates advance incrementally: no division or parameter reload per k-block)
//   warp 11     B-operand TMA producer (same ring, same barriers: the A producer announces the bytes of both)
//   warp 1      TMEM allocator + MMA issuer: one instruction group per k-block (non-blocking poll of the next stage,
//                               four tcgen05.mma of K = 16, commit) into one of TWO accumulator stages of BN columns, so
//                               tile i+1 is multiplied while tile i is drained
//   warps 2..9  epilogue       (8 warps; warp w owns TMEM lanes 32*(w%4).. and every other 32-column chunk.  A thread
//                               owns ONE output row: tcgen05.ld gives it 32 consecutive fp32 columns, which it runs
//                               through bias (staged in shared memory during the main loop) / ReLU / ReLU-mask / dropout
//                               / residual in registers, packs, and writes into a 128-byte-swizzled staging panel in
//                               shared memory; one elected thread per panel then issues a TMA tensor store (or fp32
//                               reduce-add for split-K / accumulate), double-buffered so stores drain while the next
//                               panel is computed.  TMA clips ragged M / N edges and maps convolution pixel boxes back to
//                               NHWC.  fp32 residuals / ReLU masks are read directly, one chunk ahead of their use.)
//   warp 10     residual producer (bf16 residual tiles of the ResNet block tails, TMA-loaded panel by panel into a
//                               two-slot ring so the HBM latency of the epilogue's reads is covered by bytes in flight)
// CTAS = 2: a CTA pair (cluster of two, cta_group::2) runs one 256-row MMA, each CTA staging half of the B tile.
// KS: cluster split-K instantiations (a cluster per tile, k-slices per CTA, partial sums through an L2 workspace).
// Operands may be K-major or MN-major in global memory (instruction-descriptor transpose bits), so
// forward (X W^T), dgrad (dY W) and wgrad (dY^T X) all run on this kernel without any transposed copy.
// A may also be an implicit-GEMM convolution operand: NHWC activations read through a 4-D tensor map,
// k-block -> (filter tap, 64-channel chunk), out-of-image taps zero-filled by TMA.
//
// Replaces, on the reference's hot path, every nn.Linear / nn.Conv2d / nn.ConvTranspose2d contraction
// (model/resnet_vqa_model.py:64-78,119-135,154; model/multi_head_vision_text_attn.py:31-34,92-93;
// hf T5 q/k/v/o/wi/wo; torchvision ResNet convs) and their autograd backward.
#pragma once
#include "common.cuh"
#include "gemm.cuh"
#include "ptx.cuh"
#include "rng.cuh"

// Bring-up instrumentation (phase stamps, load-only / MMA-only modes) is compiled only into the diagnostic library
// (python t5-resnet-vqa_b200/build.py --debug -> libvqa_b200_dbg.so, selected with VQA_B200_LIB): it is ~10 % of the
// kernel's instructions, and the step's ~400 short launches pay for every instruction-cache line they touch.
#ifdef VQA_GEMM_DEBUG
#define VQA_DBG_MODE(p) ((p).dbg_mode)
#define VQA_DBG_CLK(p) ((p).dbg_clk)
#define VQA_GSTAMP(k) do { if (p.dbg_clk != nullptr && blockIdx.x < 2 && (threadIdx.x & 31) == 0) p.dbg_clk[(blockIdx.x * 11 + (threadIdx.x >> 5)) * 16 + (k)] = clock64(); } while (0)
#else
#define VQA_DBG_MODE(p) 0
#define VQA_DBG_CLK(p) (static_cast<long long*>(nullptr))
#define VQA_GSTAMP(k) do { } while (0)
#endif

namespace vqa {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (4 + kEpiWarps);   // A producer, MMA, 8 epilogue, residual producer, B producer
constexpr int kChunkBytes = BK * 128;   // one 64-wide MN-major chunk: 64 k-rows x 128 B
constexpr int kPanelBytes = BM * 128;   // staging panel: 128 rows x 128 B (32 fp32 or 64 bf16 columns)

template <int BN, int STAGES, int CTAS>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = (BN / CTAS) * BK * 2;   // a CTA pair splits the B tile between its two CTAs
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_OFF = STAGES * STAGE_BYTES;
  // four panels: fp32 output -> two double-buffered panels per column-half group; bf16 output -> a double-buffered
  // output panel [0, 2) and the two-slot bf16 residual ring [2, 4)
  static constexpr int STG_BYTES = 4 * kPanelBytes;
  static constexpr int BAR_OFF = STG_OFF + STG_BYTES;
  static constexpr int NBARS = 2 * STAGES + 9;  // full[], empty[], tmem_full[2], tmem_empty[2], res_full[2], res_empty[2], ksplit
  static constexpr int BIAS_OFF = (BAR_OFF + NBARS * 8 + 16 + 15) & ~15;   // the tile's bias columns, double-buffered
  static constexpr int SMEM_BYTES = BIAS_OFF + 2 * 256 * 4 + 1024;
  static constexpr int TMEM_COLS = 2 * BN;      // two accumulator stages (128, 256 or 512 columns)
};

struct TileCoord {
  int m0;             // first output row (linear outputs)
  int n0;             // first output column
  int pw0, ph0, pn0;  // pixel-box origin (conv outputs)
  int kb_begin, num_kb;
};

// fd_tiles_m counts the m-tiles of a scheduling unit (a CTA, or a CTA pair covering `ctas` consecutive m-tiles)
__device__ __forceinline__ TileCoord tile_coord(const GemmParams& p, int t, const FastDiv& fd_tiles_m,
                                                const FastDiv& fd_tiles_n, int splits, int bn, int ctas = 1,
                                                int rank = 0) {
  TileCoord tc;
  const int rest = fast_div(t, fd_tiles_n);
  const int nt = t - rest * static_cast<int>(fd_tiles_n.d);
  const int z = fast_div(rest, fd_tiles_m);
  const int mt = (rest - z * static_cast<int>(fd_tiles_m.d)) * ctas + rank;
  tc.n0 = nt * bn;
  tc.m0 = mt * BM;
  tc.pw0 = tc.ph0 = tc.pn0 = 0;
  if (p.a_mode == LOAD_CONV) {
    const int r1 = fast_div(mt, p.fd_tiles_w);
    const int tw = mt - r1 * p.tiles_w;
    const int tn = fast_div(r1, p.fd_tiles_h);
    const int th = r1 - tn * p.tiles_h;
    tc.pw0 = tw * p.bx_w;
    tc.ph0 = th * p.bx_h;
    tc.pn0 = tn * p.bx_n;
  }
  if (splits == 1) {
    tc.kb_begin = 0;
    tc.num_kb = p.kb_total;
  } else {
    tc.kb_begin = static_cast<int>((static_cast<long long>(p.kb_total) * z) / splits);
    const int kb_end = static_cast<int>((static_cast<long long>(p.kb_total) * (z + 1)) / splits);
    tc.num_kb = kb_end - tc.kb_begin;
  }
  return tc;
}

}  // namespace

// EPI: compile-time epilogue variant (bit 0 fp32 output, bits 1-2 residual: 0 none / 1 bf16 through the TMA ring /
// 2 fp32 read directly, bit 3 ReLU-mask, bit 4 dropout, bit 5 fp32 reduce-add output, bit 6 exact GELU instead of ReLU).
__host__ __device__ constexpr int epi_code(bool out_fp32, int res, bool mask, bool drop, bool atomic) {
  return (out_fp32 ? 1 : 0) | (res << 1) | (mask ? 8 : 0) | (drop ? 16 : 0) | (atomic ? 32 : 0);
}

// KS: cluster split-K support compiled in (a separate set of instantiations, so that the default kernels do not carry
// its code: +50 % instructions cost 0.3 ms per training step in instruction-cache misses of ~400 short launches).
template <int BN, int STAGES, int EPI, int CTAS, bool KS>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRes,
                    const __grid_constant__ GemmParams p, const FastDiv fd_tiles_m, const FastDiv fd_tiles_n,
                    const int splits) {
  using C = Cfg<BN, STAGES, CTAS>;
  static_assert(CTAS == 1 || BN >= 128, "a CTA pair needs at least 64 B-tile rows per CTA");
  constexpr bool kOutF32 = (EPI & 1) != 0;
  constexpr int kRes = (EPI >> 1) & 3;
  constexpr bool kMask = (EPI & 8) != 0;
  constexpr bool kDrop = (EPI & 16) != 0;
  constexpr bool kAtomic = (EPI & 32) != 0;
  constexpr bool kGelu = (EPI & 64) != 0;   // exact (erf) GELU after the bias instead of ReLU: the frozen ViT's fc1 (vit: ViTIntermediate)
  static_assert(!(kRes == 1 && kOutF32), "a bf16 residual (TMA ring) is only combined with bf16 output");
  static_assert(!kGelu || (!kOutF32 && kRes == 0 && !kMask && !kDrop && !kAtomic && !KS), "GELU: plain bf16-output launches only");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + C::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  auto res_full_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 4 + a); };
  auto res_empty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 6 + a); };
  const uint32_t ks_bar = bar_base + 8u * (2 * STAGES + 8);   // cluster split-K: the peers' partial sums have landed
  const uint32_t tmem_slot = bar_base + 8u * C::NBARS;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + C::BAR_OFF + 8 * C::NBARS);

  // broadcast from lane 0: tells ptxas the warp index is warp-uniform, so the role branches below are uniform branches
  // and the role loops can keep their state in uniform registers
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  // cluster split-K (CTAS == 1 only): the KSP CTAs of a cluster share one output tile; CTA `krank` contracts k-slice
  // `krank`, hands the partial sums of the columns it does not own to their owners through p.ks_ws, and finishes
  // (bias / activation / residual / store) its own column range.  The cluster is co-scheduled by the hardware, so
  // waiting for a peer can never deadlock, whatever else runs on the GPU.
  const int KSP = (KS && CTAS == 1 && p.ksplit > 1) ? p.ksplit : 1;
  const int krank = KSP > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int out_tiles = static_cast<int>(fd_tiles_m.d * fd_tiles_n.d);
  const int total_tiles = KSP > 1 ? out_tiles : out_tiles * splits;
  // scheduling unit: one CTA, a CTA pair (cluster of 2) that owns two consecutive m-tiles of one n-tile, or a
  // split-K cluster
  const int rank = CTAS == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const bool leader = rank == 0;
  const int unit = static_cast<int>(blockIdx.x) / (CTAS * KSP);
  const int nunits = static_cast<int>(gridDim.x) / (CTAS * KSP);
  // column range this CTA finishes, in 32-column chunks, handed out in 64-column panels: [own_lo, own_hi)
  constexpr int NCH = BN / 32;
  int own_lo = 0, own_hi = NCH;
  if (KS && KSP > 1) {
    const int npan = NCH / 2, base = npan / KSP, rem = npan - base * KSP;
    own_lo = 2 * (krank * base + (krank < rem ? krank : rem));
    own_hi = own_lo + 2 * (base + (krank < rem ? 1 : 0));
  }
  auto coord = [&](int t) {
    return (KS && KSP > 1) ? tile_coord(p, t + krank * out_tiles, fd_tiles_m, fd_tiles_n, KSP, BN, CTAS, rank)
                   : tile_coord(p, t, fd_tiles_m, fd_tiles_n, splits, BN, CTAS, rank);
  };

  // ---- one-time setup (overlaps the previous kernel's tail under programmatic dependent launch) ----
  VQA_GSTAMP(0);
  pdl_launch_dependents();
  if (warp == 0) {
    // one barrier per lane (initialising ~20 barriers from one thread is ~400 cycles of every launch's start-up)
    if (lane < 2 * STAGES) mbar_init(bar_base + 8u * lane, 1);   // full[]: the leader's expect_tx arrival; empty[]: commit
    else if (lane < 2 * STAGES + 2) mbar_init(bar_base + 8u * lane, 1);                   // tmem_full[2]
    else if (lane < 2 * STAGES + 4) mbar_init(bar_base + 8u * lane, kEpiWarps * CTAS);    // tmem_empty[2]: both CTAs' epilogues
    else if (lane < 2 * STAGES + 6) mbar_init(bar_base + 8u * lane, 1);                   // res_full[2]
    else if (lane < 2 * STAGES + 8) mbar_init(bar_base + 8u * lane, kEpiWarps);           // res_empty[2]
    else if (lane == 2 * STAGES + 8) mbar_init(ks_bar, KSP > 1 ? KSP - 1 : 1);
    static_assert(2 * STAGES + 9 <= 32, "one barrier per lane");
    mbar_fence_init();
  } else if (warp == 2 && lane < 4) {
    const CUtensorMap* tm = lane == 0 ? &tmA : (lane == 1 ? &tmB : (lane == 2 ? &tmOut : &tmRes));
    if (lane < 3 || kRes == 1 || p.nseg > 1) tma_prefetch_desc(tm);
  }
  if (warp == 1) {
    if (CTAS == 2) { tmem_alloc_pair(tmem_slot, C::TMEM_COLS); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, C::TMEM_COLS); tmem_relinquish(); }
  }
  // the first tile's coordinates depend on launch parameters only: computed here, while barrier initialisation and
  // the TMEM allocation are in flight, so the (cold) parameter reads are off the path to the first TMA load
  const TileCoord tc_first = coord(unit);
  tc_fence_before();
  if (CTAS == 2 || (KS && KSP > 1)) cluster_sync_all();   // the peers' barriers must be initialised before anyone signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  VQA_GSTAMP(1);
  pdl_wait();   // nothing above touches global memory; everything below may
  VQA_GSTAMP(2);

  if (warp == 0 || warp == 3 + kEpiWarps) {
    // ===================================== TMA producers =====================================
    // A tiles are issued from warp 0, B tiles from the last warp, so the two operands' issue latencies overlap.
    // The k-loop is kept as short as a scalar instruction stream can be: the WHOLE warp runs the (uniform) loop and
    // polls the barrier, lane 0 alone issues; every per-k-block coordinate is carried incrementally (no division,
    // no parameter reload inside the loop).  Measured (tools/gemm_dbg.py): the previous generic loop body cost
    // ~500 cycles per k-block on its own, more than the MMAs of a 128 x 128 x 64 block (256 cycles).
    const bool do_a = warp == 0;
    const int a_mode = p.a_mode, a_mn = p.a_mn, b_mode = p.b_mode, b_mn = p.b_mn;
    const uint32_t tx_bytes = static_cast<uint32_t>(p.stage_tx_bytes) * CTAS;
    // pair: both CTAs' loads are credited to the LEADER's full barrier (the leader issues the MMA)
    const uint32_t fb0 = CTAS == 2 ? mapa_shared(full_bar(0), 0) : full_bar(0);
    const bool arrive = leader && do_a;   // one arrival per stage; everybody's bytes are counted by the same barrier
    auto load2 = [](uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
      tma_load_2d_e<CTAS>(dst, m, bar, c0, c1);
    };
    auto load4 = [](uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
      tma_load_4d_e<CTAS>(dst, m, bar, c0, c1, c2, c3);
    };
    const bool pdbg = VQA_DBG_CLK(p) != nullptr && blockIdx.x < 2 && do_a;
    long long pdbg_wait = 0, pdbg_t0 = pdbg ? clock64() : 0;
    int stage = 0;
    uint32_t phase = 0;
    // one pipeline slot: wait until the MMAs of the previous round have released it, announce the bytes, then `issue`
#define VQA_PRODUCE(ISSUE)                                                            \
    do {                                                                              \
      const long long pc0 = pdbg ? clock64() : 0;                                     \
      mbar_wait_lean(empty_bar(stage), phase ^ 1u);                                   \
      if (pdbg) pdbg_wait += clock64() - pc0;                                         \
      const uint32_t fb = fb0 + 8u * stage;                                           \
      const uint32_t sa = smem_base + stage * C::STAGE_BYTES;                         \
      const uint32_t sb = sa + C::A_BYTES;                                            \
      (void)sa; (void)sb;                                                             \
      if (arrive) mbar_expect_tx_e(full_bar(stage), tx_bytes);                        \
      ISSUE;                                                                          \
      if (++stage == STAGES) { stage = 0; phase ^= 1u; }                              \
    } while (0)
    for (int t = unit; t < total_tiles; t += nunits) {
      const TileCoord tc = t == unit ? tc_first : coord(t);
      const int nkb = tc.num_kb;
      if (VQA_DBG_MODE(p) >= 2 && VQA_DBG_MODE(p) <= 4) {   // bring-up: no loads at all (measures the MMA side alone)
        for (int i = 0; i < nkb; ++i) {
          mbar_wait_lean(empty_bar(stage), phase ^ 1u);
          if (arrive) mbar_arrive_e(full_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        continue;
      }
      if (do_a) {
        if (a_mode == LOAD_2D && !a_mn) {
          if (p.nseg > 1) {   // two-term operand split: segments (A, B), [(A_lo, B),] (A, B_lo) over the same tile
            const int nseg = p.nseg, kseg = p.kseg;
            for (int sgm = 0; sgm < nseg; ++sgm) {
              int k0 = (sgm == 1 && nseg == 3) ? p.a_lo_col : 0;
              for (int i = 0; i < kseg; ++i, k0 += BK) VQA_PRODUCE(load2(sa, &tmA, fb, k0, tc.m0));
            }
          } else {
            int k0 = tc.kb_begin * BK;
            for (int i = 0; i < nkb; ++i, k0 += BK) VQA_PRODUCE(load2(sa, &tmA, fb, k0, tc.m0));
          }
        } else if (a_mode == LOAD_2D) {
          int k0 = tc.kb_begin * BK;
          for (int i = 0; i < nkb; ++i, k0 += BK)
            VQA_PRODUCE(load2(sa, &tmA, fb, tc.m0, k0); load2(sa + kChunkBytes, &tmA, fb, tc.m0 + 64, k0));
        } else if (a_mode == LOAD_CONV) {
          // k-block -> (filter row r, filter column s, 64-channel chunk cc), advanced like an odometer
          const int cchunks = p.cchunks, taps_s = p.taps_s, dil_w = p.dil_w;
          const int tap = tc.kb_begin / cchunks;
          int cc = tc.kb_begin - tap * cchunks;
          int r = tap / taps_s;
          int sx = tap - r * taps_s;
          const int w0 = tc.pw0 * p.stride_w - p.pad_w, h0 = tc.ph0 * p.stride_h - p.pad_h;
          for (int i = 0; i < nkb; ++i) {
            VQA_PRODUCE(load4(sa, &tmA, fb, cc * 64, w0 + sx * dil_w, h0 + r, tc.pn0));
            if (++cc == cchunks) { cc = 0; if (++sx == taps_s) { sx = 0; ++r; } }
          }
        } else {  // LOAD_PIXELS_MN: k-block -> pixel box (tw, th, tn)
          const int tiles_w = p.tiles_w, tiles_h = p.tiles_h, bx_w = p.bx_w, bx_h = p.bx_h, bx_n = p.bx_n;
          int tw = tc.kb_begin % tiles_w;
          int th = (tc.kb_begin / tiles_w) % tiles_h;
          int tn = tc.kb_begin / (tiles_w * tiles_h);
          for (int i = 0; i < nkb; ++i) {
            VQA_PRODUCE(load4(sa, &tmA, fb, tc.m0, tw * bx_w, th * bx_h, tn * bx_n);
                        load4(sa + kChunkBytes, &tmA, fb, tc.m0 + 64, tw * bx_w, th * bx_h, tn * bx_n));
            if (++tw == tiles_w) { tw = 0; if (++th == tiles_h) { th = 0; ++tn; } }
          }
        }
      } else {
        const int nb0 = tc.n0 + rank * (BN / CTAS);   // first B-tile row / column staged by this CTA
        if (b_mode == LOAD_2D && !b_mn) {
          if (p.nseg > 1) {   // the last segment reads the low-order half of B (residual tensor map)
            const int nseg = p.nseg, kseg = p.kseg;
            for (int sgm = 0; sgm < nseg; ++sgm) {
              const CUtensorMap* mb = sgm == nseg - 1 ? &tmRes : &tmB;
              int k0 = 0;
              for (int i = 0; i < kseg; ++i, k0 += BK) VQA_PRODUCE(load2(sb, mb, fb, k0, nb0));
            }
          } else {
            int k0 = tc.kb_begin * BK;
            for (int i = 0; i < nkb; ++i, k0 += BK) VQA_PRODUCE(load2(sb, &tmB, fb, k0, nb0));
          }
        } else if (b_mode == LOAD_2D) {
          int k0 = tc.kb_begin * BK;
          for (int i = 0; i < nkb; ++i, k0 += BK) {
            VQA_PRODUCE(_Pragma("unroll") for (int j = 0; j < BN / CTAS / 64; ++j)
                            load2(sb + j * kChunkBytes, &tmB, fb, nb0 + 64 * j, k0));
          }
        } else {  // LOAD_PIXELS_MN: column n -> (tap, input channel); k-block -> pixel box
          const int tiles_w = p.tiles_w, tiles_h = p.tiles_h, bx_w = p.bx_w, bx_h = p.bx_h, bx_n = p.bx_n;
          const int tap = nb0 / p.b_tap_cin, ci0 = nb0 - tap * p.b_tap_cin;
          const int r = tap / p.b_taps_s, sx = tap - r * p.b_taps_s;
          const int dw = sx - p.pad_w, dh = r - p.pad_h;
          int tw = tc.kb_begin % tiles_w;
          int th = (tc.kb_begin / tiles_w) % tiles_h;
          int tn = tc.kb_begin / (tiles_w * tiles_h);
          for (int i = 0; i < nkb; ++i) {
            VQA_PRODUCE(_Pragma("unroll") for (int j = 0; j < BN / CTAS / 64; ++j)
                            load4(sb + j * kChunkBytes, &tmB, fb, ci0 + 64 * j, tw * bx_w + dw, th * bx_h + dh, tn * bx_n));
            if (++tw == tiles_w) { tw = 0; if (++th == tiles_h) { th = 0; ++tn; } }
          }
        }
      }
    }
#undef VQA_PRODUCE
    if (pdbg && lane == 0) {
      p.dbg_clk[(blockIdx.x * 11 + 0) * 16 + 12] = pdbg_wait;
      p.dbg_clk[(blockIdx.x * 11 + 0) * 16 + 13] = clock64() - pdbg_t0;
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ========================================
    // The whole warp of the leader CTA runs the loop (so every operand is warp-uniform and lives in uniform
    // registers); lane 0 issues.  Descriptors: constant high word, low word = (address >> 4) advanced by constants.
    if (leader) {
      const uint32_t idesc = umma_idesc_bf16(BM * CTAS, BN, p.a_mn != 0, p.b_mn != 0);
      const uint32_t a_lbo = p.a_mn ? kChunkBytes : 16u, b_lbo = p.b_mn ? kChunkBytes : 16u;
      const uint32_t a_k16 = (p.a_mn ? 16u * 128u : 32u) >> 4, b_k16 = (p.b_mn ? 16u * 128u : 32u) >> 4;
      const uint64_t da0 = umma_smem_desc(smem_base, a_lbo, 1024u);
      const uint64_t db0 = umma_smem_desc(smem_base + C::A_BYTES, b_lbo, 1024u);
      const uint32_t da_hi = static_cast<uint32_t>(da0 >> 32), db_hi = static_cast<uint32_t>(db0 >> 32);
      const uint32_t da_lo0 = static_cast<uint32_t>(da0), db_lo0 = static_cast<uint32_t>(db0);
      const int dbg_mode = VQA_DBG_MODE(p);
      const bool dbg_on = VQA_DBG_CLK(p) != nullptr && blockIdx.x < 2;
      long long dbg_wait = 0, dbg_issue = 0;   // bring-up: cycles this warp spent waiting for data / issuing
      uint32_t next_ready = 0;                 // the next stage's full barrier has already been seen complete
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = unit; t < total_tiles; t += nunits) {
        const TileCoord tc = t == unit ? tc_first : coord(t);
        mbar_wait_lean(tmem_empty_bar(acc), acc_phase ^ 1u);   // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        const int nkb = tc.num_kb;
        for (int i = 0; i < nkb; ++i) {
          const long long c0 = dbg_on ? clock64() : 0;
          if (!next_ready) mbar_wait_lean(full_bar(stage), phase);   // else: already seen full by the previous block's poll
          tc_fence_after();
          const long long c1 = dbg_on ? clock64() : 0;
          if (i == 0 && t == unit) VQA_GSTAMP(3);
          const uint32_t soff = static_cast<uint32_t>(stage * (C::STAGE_BYTES >> 4));
          const uint32_t la = da_lo0 + soff, lb = db_lo0 + soff;
          if (dbg_mode == 0) {
            static_assert(BK == 64, "umma_kblock4_e issues four K = 16 MMAs");
            const int ns = stage + 1 == STAGES ? 0 : stage + 1;
            next_ready = umma_kblock4_e<CTAS>(d_tmem, (static_cast<uint64_t>(da_hi) << 32) | la,
                                              (static_cast<uint64_t>(db_hi) << 32) | lb, a_k16, b_k16, idesc, i ? 1u : 0u,
                                              empty_bar(stage), full_bar(ns), ns == 0 ? (phase ^ 1u) : phase);
            next_ready = __shfl_sync(0xffffffffu, next_ready, 0);   // warp-uniform by construction
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            if (dbg_on) { const long long c2 = clock64(); dbg_wait += c1 - c0; dbg_issue += c2 - c1; }
            continue;
          } else if (dbg_mode == 6) {   // bring-up: blocking wait per k-block, four separately elected MMAs (the previous loop)
#pragma unroll
            for (int ks = 0; ks < BK / 16; ++ks) {
              const uint64_t da = (static_cast<uint64_t>(da_hi) << 32) | (la + ks * a_k16);
              const uint64_t db = (static_cast<uint64_t>(db_hi) << 32) | (lb + ks * b_k16);
              umma_bf16_e<CTAS>(d_tmem, da, db, idesc, (i | ks) ? 1u : 0u);
            }
          } else if (dbg_mode == 2 || dbg_mode == 3 || dbg_mode == 5) {   // bring-up: 2 = no loads, 3 = no loads and one MMA per block, 5 = unfused issue
            for (int ks = 0; ks < (dbg_mode == 3 ? 1 : BK / 16); ++ks) {
              const uint64_t da = (static_cast<uint64_t>(da_hi) << 32) | (la + ks * a_k16);
              const uint64_t db = (static_cast<uint64_t>(db_hi) << 32) | (lb + ks * b_k16);
              umma_bf16_e<CTAS>(d_tmem, da, db, idesc, (i | ks) ? 1u : 0u);
            }
          }
          // smem slot reusable (in both CTAs of a pair) once these MMAs retire
          umma_commit_e<CTAS>(empty_bar(stage));
          next_ready = 0;
          if (dbg_on) { const long long c2 = clock64(); dbg_wait += c1 - c0; dbg_issue += c2 - c1; }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        // accumulator complete (fires immediately when num_kb == 0)
        umma_commit_e<CTAS>(tmem_full_bar(acc));
        if (t == unit) VQA_GSTAMP(4);
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
      if (dbg_on && lane == 0) {
        p.dbg_clk[(blockIdx.x * 11 + 1) * 16 + 12] = dbg_wait;
        p.dbg_clk[(blockIdx.x * 11 + 1) * 16 + 13] = dbg_issue;
      }
    }
  } else if (warp == 2 + kEpiWarps) {
    // ================================ bf16 residual producer =================================
    if (kRes == 1 && lane == 0) {
      const uint32_t ring = smem_base + C::STG_OFF + 2 * kPanelBytes;
      int slot = 0;
      uint32_t phase = 0;
      for (int t = unit; t < total_tiles; t += nunits) {
        const TileCoord tc = t == unit ? tc_first : coord(t);
        if (tc.kb_begin != 0 && KSP == 1) continue;   // atomic split-K: the residual is added by the first k-slice only
        for (int pn = own_lo / 2; pn < own_hi / 2; ++pn) {
          const int n = tc.n0 + pn * 64;
          if (n >= p.N) break;
          mbar_wait_lean(res_empty_bar(slot), phase ^ 1u);
          mbar_expect_tx(res_full_bar(slot), static_cast<uint32_t>(p.res_tx_bytes));
          if (p.out_pixels) tma_load_4d(ring + slot * kPanelBytes, &tmRes, res_full_bar(slot), n, tc.pw0, tc.ph0, tc.pn0);
          else tma_load_2d(ring + slot * kPanelBytes, &tmRes, res_full_bar(slot), n, tc.m0);
          if (++slot == 2) { slot = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ===================================== epilogue ==========================================
    const int ew = warp - 2;             // 0..7
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int half = ew >> 2;            // which alternating 32-column chunks this warp drains
    const int row = q * 32 + lane;       // tile row owned by this thread
    const int sw = row & 7;              // 128B-swizzle phase of that row
    uint8_t* stg_gen = smem_gen + C::STG_OFF;
    const uint32_t stg = smem_base + C::STG_OFF;
    const uint32_t row_off = static_cast<uint32_t>((row >> 3) * 1024 + sw * 128);
    // fp32 output: each column-half group (4 warps) owns two panels; bf16 output: all 8 warps share two panels
    const int panel0 = kOutF32 ? 2 * half : 0;
    const bool issuer = (q == 0 && lane == 0 && (kOutF32 || half == 0));
    const uint32_t bar_id = kOutF32 ? 1u + half : 1u;
    const uint32_t bar_threads = kOutF32 ? 128u : 256u;

    unsigned long long seed = 0, offset = 0;
    uint32_t thresh = 0;
    float keep_scale = 1.f;
    if (kDrop) {
      seed = p.rng[0]; offset = p.rng[1];
      thresh = drop_threshold(p.drop_p);
      keep_scale = 1.f / (1.f - p.drop_p);
    }

    int acc = 0;
    uint32_t acc_phase = 0;
    int buf = 0;                          // staging double buffer
    int rslot = 0;                        // residual ring position
    uint32_t rphase = 0;
    int tile_it = 0;
    for (int t = unit; t < total_tiles; t += nunits) {
      const TileCoord tc = t == unit ? tc_first : coord(t);
      const bool first_split = (tc.kb_begin == 0) || KSP > 1;   // cluster split-K: every CTA finishes its own columns
      const bool add_bias = p.bias != nullptr && first_split;
      const bool add_res = kRes != 0 && first_split;
      const long long grow = static_cast<long long>(tc.m0) + row;     // global row (linear outputs)
      const bool row_ok = grow < p.M;
      const uint32_t t_tile = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
      bool waited = false;

      if (KS && KSP > 1) {
        // ---- cluster split-K, phase 1: partial sums of the columns other CTAs finish go to the workspace ----
        mbar_wait_lean(tmem_full_bar(acc), acc_phase);
        tc_fence_after();
        waited = true;
        // workspace layout [cluster][source CTA][chunk][k = 0..7][row] float4: a warp's store / load instruction covers
        // 32 consecutive float4 (512 contiguous bytes)
        float4* wsrc = reinterpret_cast<float4*>(p.ks_ws) + (static_cast<size_t>(unit) * KSP + krank) * (NCH * 8 * 128) + row;
#pragma unroll 1
        for (int c = half; c < NCH; c += 2) {
          if (c >= own_lo && c < own_hi) continue;
          const int n = tc.n0 + c * 32;
          if (!(kOutF32 ? (n < p.N) : ((n & ~63) < p.N))) continue;
          uint32_t accr[32];
          if (tc.num_kb > 0) {
            tmem_ld_32x32(t_tile + static_cast<uint32_t>(c * 32), accr);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) accr[i] = 0u;
          }
          float4* dst = wsrc + c * (8 * 128);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            __stcg(dst + k * 128, make_float4(__uint_as_float(accr[4 * k]) * p.alpha, __uint_as_float(accr[4 * k + 1]) * p.alpha,
                                        __uint_as_float(accr[4 * k + 2]) * p.alpha, __uint_as_float(accr[4 * k + 3]) * p.alpha));
        }
        if (own_hi <= own_lo) {   // this CTA finishes nothing: its TMEM reads are over
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
        }
        VQA_GSTAMP(9);
        asm volatile("fence.acq_rel.cluster;" ::: "memory");   // this thread's partial sums are visible in the cluster
        named_bar_sync(3, 32 * kEpiWarps);
        VQA_GSTAMP(10);
        if (ew == 0 && lane == 0) {
          // release.cluster: our stores are visible to the owner once it has seen the arrival.  Only CTAs that finish
          // columns wait (and are therefore still resident); a CTA that owns nothing may already have exited.
          const int npan = NCH / 2, base = npan / KSP, rem = npan - base * KSP;
          for (int r = 0; r < KSP; ++r)
            if (r != krank && base + (r < rem ? 1 : 0) > 0) mbar_arrive_cluster(mapa_shared(ks_bar, r));
        }
        if (own_hi > own_lo) mbar_wait_cluster(ks_bar, static_cast<uint32_t>(((t - unit) / nunits) & 1));
        VQA_GSTAMP(11);
      }

      // the tile's bias columns go to shared memory while the main loop is still running (a per-chunk global load
      // exposed an L2 round trip per chunk: ~0.7 us of a 5 us launch)
      float* bias_s = reinterpret_cast<float*>(smem_gen + C::BIAS_OFF) + (tile_it & 1) * 256;
      ++tile_it;
      if (add_bias) {
        const int col = static_cast<int>(threadIdx.x) - 64;   // epilogue thread 0..255
        if (col < BN) bias_s[col] = (tc.n0 + col < p.N) ? __ldg(p.bias + tc.n0 + col) : 0.f;
      }
      uint4 rres_n[8];
      uint4 rmsk_n[4];
      auto load_direct = [&](int cc, uint4 (&rr)[8], uint4 (&rm)[4]) {
        const int nn = tc.n0 + cc * 32;
        const bool act = kOutF32 ? (nn < p.N) : ((nn & ~63) < p.N);
        if (kRes == 2 && act && add_res && row_ok) {
          const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.residual) + grow * p.ldr + nn);
#pragma unroll
          for (int k = 0; k < 8; ++k) rr[k] = __ldg(rp + k);
        }
        if (kMask && act && row_ok) {
          const uint4* mp = reinterpret_cast<const uint4*>(p.relu_mask + grow * p.ldm + nn);
#pragma unroll
          for (int k = 0; k < 4; ++k) rm[k] = __ldg(mp + k);
        }
      };
      if (own_lo + half < own_hi) load_direct(own_lo + half, rres_n, rmsk_n);
      // cluster split-K has already waited for the accumulator (phase 1); otherwise the wait sits inside the chunk loop,
      // after the second chunk's operand loads have been requested
      if (waited && add_bias) named_bar_sync(4, 32 * kEpiWarps);
#pragma unroll 1
      for (int c = own_lo + half; c < own_hi; c += 2) {
        const int n = tc.n0 + c * 32;
        const bool last_chunk = (c + 2 >= own_hi);
        // uniform over the barrier group: fp32 panels are one chunk wide, bf16 panels two (one per column half)
        const bool active = kOutF32 ? (n < p.N) : ((n & ~63) < p.N);
        float v[32];
        // ---- directly read operands (fp32 residual, ReLU mask): this chunk's were requested one chunk ago (the first
        // chunk's before the accumulator was waited for); the NEXT chunk's are requested now, so their L2 round trip
        // runs under this chunk's TMEM read, arithmetic and store instead of being waited for ----
        uint4 rres[8];
        uint4 rmsk[4];
#pragma unroll
        for (int k = 0; k < 8; ++k) rres[k] = rres_n[k];
#pragma unroll
        for (int k = 0; k < 4; ++k) rmsk[k] = rmsk_n[k];
        if (c + 2 < own_hi) load_direct(c + 2, rres_n, rmsk_n);
        if (!waited) {
          mbar_wait_lean(tmem_full_bar(acc), acc_phase);
          tc_fence_after();
          waited = true;
          if (t == unit) VQA_GSTAMP(5);
          if (add_bias) named_bar_sync(4, 32 * kEpiWarps);   // every epilogue thread's bias column is in shared memory
        }
        float4 pks[8];
        if (KS && KSP > 1 && active) {   // first peer's partial sums of this chunk
          const float4* src = reinterpret_cast<const float4*>(p.ks_ws) +
                              ((static_cast<size_t>(unit) * KSP + (krank + 1) % KSP) * NCH + c) * (8 * 128) + row;
#pragma unroll
          for (int k = 0; k < 8; ++k) pks[k] = __ldcg(src + k * 128);
        }
        if (active) {
          uint32_t accr[32];
          if (tc.num_kb > 0) {
            tmem_ld_32x32(t_tile + static_cast<uint32_t>(c * 32), accr);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) accr[i] = 0u;
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(accr[i]) * p.alpha;
          if (KS && KSP > 1) {
            // phase 2: add the peers' partial sums of this chunk (in L2, read past L1).  The first peer's loads were
            // issued before the accumulator read; every further peer's loads are issued before the previous peer's
            // values are consumed, so one L2 round trip is exposed per chunk, not one per peer.
            int r = (krank + 1) % KSP;
            for (int i = 1; i < KSP; ++i) {
              float4 cur[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) cur[k] = pks[k];
              const int rn = (r + 1) % KSP;
              if (i + 1 < KSP) {
                const float4* src = reinterpret_cast<const float4*>(p.ks_ws) +
                                    ((static_cast<size_t>(unit) * KSP + rn) * NCH + c) * (8 * 128) + row;
#pragma unroll
                for (int k = 0; k < 8; ++k) pks[k] = __ldcg(src + k * 128);
              }
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                v[4 * k] += cur[k].x; v[4 * k + 1] += cur[k].y; v[4 * k + 2] += cur[k].z; v[4 * k + 3] += cur[k].w;
              }
              r = rn;
            }
          }
        }
        if (last_chunk) {
          // all of this warp's TMEM reads of the tile are done: hand the accumulator stage back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CTAS == 2) mbar_arrive_cluster(mapa_shared(tmem_empty_bar(acc), 0));
            else mbar_arrive(tmem_empty_bar(acc));
          }
        }
        if (active) {
          if (add_bias) {   // staged in shared memory at the top of the tile (zeros beyond N)
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c * 32 + 4 * k);
              v[4 * k] += b4.x; v[4 * k + 1] += b4.y; v[4 * k + 2] += b4.z; v[4 * k + 3] += b4.w;
            }
          }
          if (kRes == 1 && add_res) {
            // bf16 residual panel (64 columns) from the TMA ring; this warp's chunk is one 64-byte half of the row
            mbar_wait_lean(res_full_bar(rslot), rphase);
            const uint8_t* rrow = stg_gen + (2 + rslot) * kPanelBytes + row_off;
            float rs[32];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint4 u = *reinterpret_cast<const uint4*>(rrow + ((((c & 1) * 4 + k) ^ sw) << 4));
              float f8[8];
              unpack_bf16x8(u, f8);
#pragma unroll
              for (int i = 0; i < 8; ++i) rs[8 * k + i] = f8[i];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(res_empty_bar(rslot));
            if (++rslot == 2) { rslot = 0; rphase ^= 1u; }
            if (p.res_first) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] += rs[i];
              if (p.relu) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
              }
            } else {
              if (p.relu) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
              }
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] += rs[i];
            }
          } else if (kGelu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.5f * v[i] * (1.f + erff(v[i] * 0.70710678118654752f));
          } else {
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
            }
          }
          if (kMask) {
            if (row_ok) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float f8[8];
                unpack_bf16x8(rmsk[k], f8);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  if (!(f8[i] > 0.f)) v[8 * k + i] = 0.f;
              }
            }
          }
          if (kDrop) {
            const unsigned long long idx = static_cast<unsigned long long>(grow) * p.N + n;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const Philox8 rnd = philox8(seed, offset, p.drop_sid, (idx >> 3) + k);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[8 * k + i] = (rnd.u16(i) < thresh) ? 0.f : v[8 * k + i] * keep_scale;
            }
          }
          if (kRes == 2 && add_res && row_ok) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              v[4 * k] += __uint_as_float(rres[k].x); v[4 * k + 1] += __uint_as_float(rres[k].y);
              v[4 * k + 2] += __uint_as_float(rres[k].z); v[4 * k + 3] += __uint_as_float(rres[k].w);
            }
          }
          // ---- stage the row into the swizzled panel ----
          if (kOutF32) {
            uint8_t* prow = stg_gen + (panel0 + buf) * kPanelBytes + row_off;
#pragma unroll
            for (int k = 0; k < 8; ++k)
              *reinterpret_cast<float4*>(prow + ((k ^ sw) << 4)) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
          } else {
            uint8_t* prow = stg_gen + (panel0 + buf) * kPanelBytes + row_off;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              uint4 u;
              u.x = pack_bf16x2(v[8 * k + 0], v[8 * k + 1]);
              u.y = pack_bf16x2(v[8 * k + 2], v[8 * k + 3]);
              u.z = pack_bf16x2(v[8 * k + 4], v[8 * k + 5]);
              u.w = pack_bf16x2(v[8 * k + 6], v[8 * k + 7]);
              *reinterpret_cast<uint4*>(prow + ((((c & 1) * 4 + k) ^ sw) << 4)) = u;
            }
          }
          fence_proxy_async_smem();
          // the previous store from the OTHER buffer must have finished reading before anyone refills it
          if (issuer) tma_store_wait_read<0>();
          named_bar_sync(bar_id, bar_threads);
          if (issuer) {
            const uint32_t src = stg + (panel0 + buf) * kPanelBytes;
            const int pn = kOutF32 ? n : (n & ~63);
            if (p.out_pixels) {
              tma_store_4d(&tmOut, src, pn, tc.pw0, tc.ph0, tc.pn0);
            } else if (kAtomic) {
              tma_reduce_add_2d(&tmOut, src, pn, tc.m0);
            } else {
              tma_store_2d(&tmOut, src, pn, tc.m0);
            }
            tma_store_commit();
          }
          buf ^= 1;
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    VQA_GSTAMP(6);
    if (issuer) tma_store_wait_all<0>();
    VQA_GSTAMP(7);
  }

  // ---- teardown --------------------------------------------------------------------------------
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all();   // the leader's MMAs read the peer's shared memory: leave together
  else __syncthreads();
  VQA_GSTAMP(8);
  if (warp == 1) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_pair(tmem_base, C::TMEM_COLS);
    else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

namespace {

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

template <int BN, int STAGES, int EPI, int CTAS, bool KS>
int launch_impl(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const CUtensorMap& tmRes,
                const GemmParams& p, int tiles_m, int tiles_n, int splits, cudaStream_t stream) {
  using C = Cfg<BN, STAGES, CTAS>;
  static_assert(C::SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static bool attr_set = false;  // per-process; all devices share the same kernel image attributes
  cudaError_t e;
  auto kern = gemm_tcgen05_kernel<BN, STAGES, EPI, CTAS, KS>;
  if (!attr_set) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  const int units_m = (tiles_m + CTAS - 1) / CTAS;     // a pair owns two consecutive m-tiles
  const int ksp = (KS && CTAS == 1 && p.ksplit > 1) ? p.ksplit : 1;   // cluster split-K: a cluster of ksp CTAs per tile
  const int total = units_m * tiles_n * (ksp > 1 ? 1 : splits);
  int sms = sm_count();
  if (p.max_ctas > 0 && p.max_ctas < sms) sms = p.max_ctas < CTAS * ksp ? CTAS * ksp : p.max_ctas;
  const int max_units = sms / (CTAS * ksp);
  if (ksp > 1 && total > max_units) return -3;          // one tile per cluster (the workspace is not double-buffered)
  const int grid = (total < max_units ? total : max_units) * CTAS * ksp;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (CTAS == 2 || ksp > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CTAS == 2 ? 2 : ksp; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr; cfg.numAttrs = na;
  e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmOut, tmRes, p, make_fastdiv(units_m), make_fastdiv(tiles_n), splits);
  if (e != cudaSuccess) return static_cast<int>(e);
  return static_cast<int>(cudaGetLastError());
}

// one translation unit per tile width instantiates these variants (compile time)
template <int BN, int STAGES, int CTAS>
int launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const CUtensorMap& tmRes,
              const GemmParams& p, int tiles_m, int tiles_n, int splits, cudaStream_t stream) {
  const int res = p.residual == nullptr ? 0 : (p.res_fp32 ? 2 : 1);
  const int code = epi_code(p.out_fp32 != 0, res, p.relu_mask != nullptr, p.drop_p > 0.f, p.atomic_out != 0) |
                   (p.relu == 2 ? 64 : 0);
#define VQA_EPI_CASE(o, r, m, d, a)                                                                   \
  case epi_code(o, r, m, d, a):                                                                       \
    if (CTAS == 1 && p.ksplit > 1)                                                                    \
      return launch_impl<BN, STAGES, epi_code(o, r, m, d, a), CTAS, CTAS == 1>(tmA, tmB, tmOut, tmRes, p, tiles_m,    \
                                                                               tiles_n, splits, stream);              \
    return launch_impl<BN, STAGES, epi_code(o, r, m, d, a), CTAS, false>(tmA, tmB, tmOut, tmRes, p, tiles_m, tiles_n, \
                                                                         splits, stream);
  switch (code) {
    VQA_EPI_CASE(false, 0, false, false, false)  // bf16 out                      (qkv, convs, plain dgrad)
    VQA_EPI_CASE(false, 1, false, false, false)  // bf16 out + bf16 residual      (ResNet block tails)
    VQA_EPI_CASE(false, 0, false, true, false)   // bf16 out + dropout            (FFN hidden, training)
    VQA_EPI_CASE(false, 0, true, false, false)   // bf16 out + ReLU mask          (dgrad through ReLU, eval)
    VQA_EPI_CASE(false, 0, true, true, false)    // bf16 out + ReLU mask + dropout (training)
    VQA_EPI_CASE(true, 2, false, false, false)   // fp32 out + fp32 residual      (residual streams)
    VQA_EPI_CASE(true, 2, false, true, false)    //   ... + dropout
    VQA_EPI_CASE(true, 2, false, false, true)    //   ... accumulated with reduce-add
    VQA_EPI_CASE(true, 0, false, false, false)   // fp32 out                      (wgrad, logits)
    VQA_EPI_CASE(true, 0, false, true, false)    // fp32 out + dropout
    VQA_EPI_CASE(true, 0, false, false, true)    // fp32 reduce-add               (split-K, accumulate)
    case 64:                                     // bf16 out + bias + exact GELU   (the frozen ViT's fc1)
      if (p.ksplit > 1) return -2;
      return launch_impl<BN, STAGES, 64, CTAS, false>(tmA, tmB, tmOut, tmRes, p, tiles_m, tiles_n, splits, stream);
    default:
      return -2;   // combination not instantiated (gemm_op_run reports it)
  }
#undef VQA_EPI_CASE
}

}  // namespace

}  // namespace vqa
