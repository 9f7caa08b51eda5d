// tcgen05 / TMEM / TMA GEMM for sm_100a:  D[M,N] = epilogue( A[M,K] * B[N,K]^T ), bf16 in, fp32 accumulate.
//
// Persistent kernel, one CTA per SM, 320 threads, static round-robin tile scheduler over 128 x BN output tiles
// (n-tile fastest so CTAs running side by side share the A tile in L2; split-K slices are extra tiles):
//   warp 0      TMA producer   (one elected lane; STAGES-deep ring of 128x64 A and BNx64 B tiles, runs ahead
//                               across tile boundaries)
//   warp 1      TMEM allocator + MMA issuer (one elected lane issues tcgen05.mma into one of TWO accumulator
//                               stages of BN columns, so tile i+1 is multiplied while tile i is drained)
//   warps 2..9  epilogue       (8 warps; warp w owns TMEM lanes 32*(w%4).. and every other 32-column chunk:
//                               tcgen05.ld 32 lanes x 32 columns -> per-warp shared-memory transpose -> each
//                               lane handles 8 consecutive columns of a row, so residual / mask loads and the
//                               output stores are coalesced 128-bit accesses with 4 independent rows in flight
//                               -> bias / ReLU / dropout / residual -> global, or fp32 red.add for split-K)
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

namespace vqa {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kChunkBytes = BK * 128;   // one 64-wide MN-major chunk: 64 k-rows x 128 B
constexpr int kStgStride = 36;          // floats per staged row (32 + 4 pad: conflict-free v4 both ways)
constexpr int kStgFloats = 32 * kStgStride;

template <int BN, int STAGES>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STG_OFF = STAGES * STAGE_BYTES;
  static constexpr int STG_BYTES = kEpiWarps * kStgFloats * 4;
  static constexpr int BAR_OFF = STG_OFF + STG_BYTES;
  static constexpr int NBARS = 2 * STAGES + 4;  // full[], empty[], tmem_full[2], tmem_empty[2]
  static constexpr int SMEM_BYTES = BAR_OFF + NBARS * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = 2 * BN;      // two accumulator stages (128, 256 or 512 columns)
};

__device__ __forceinline__ float bf16_bits_to_float(uint32_t lo16) { return __uint_as_float(lo16 << 16); }

struct TileCoord {
  int m0;             // first output row (linear outputs)
  int n0;             // first output column
  int pw0, ph0, pn0;  // pixel-box origin (conv outputs)
  int kb_begin, num_kb;
};

__device__ __forceinline__ TileCoord tile_coord(const GemmParams& p, int t, const FastDiv& fd_tiles_m,
                                                const FastDiv& fd_tiles_n, int splits, int bn) {
  TileCoord tc;
  const int rest = fast_div(t, fd_tiles_n);
  const int nt = t - rest * static_cast<int>(fd_tiles_n.d);
  const int z = fast_div(rest, fd_tiles_m);
  const int mt = rest - z * static_cast<int>(fd_tiles_m.d);
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

// Linear output row of tile row `row` (pixel boxes for convolutions) and whether it exists.
__device__ __forceinline__ long long output_row(const GemmParams& p, const TileCoord& tc, int row, bool* ok) {
  if (p.out_pixels) {
    const int t1 = fast_div(row, p.fd_bx_w);
    const int wi = row - t1 * p.bx_w;
    const int ni = fast_div(t1, p.fd_bx_h);
    const int hi = t1 - ni * p.bx_h;
    const int ow = tc.pw0 + wi, oh = tc.ph0 + hi, on = tc.pn0 + ni;
    *ok = (ni < p.bx_n) && ow < p.Wo && oh < p.Ho && on < p.Nimg;
    return (static_cast<long long>(on) * p.Ho + oh) * p.Wo + ow;
  }
  *ok = (tc.m0 + row) < p.M;
  return tc.m0 + row;
}

// Scalar, fully run-time epilogue for one row x 8 columns (ragged N tail / unaligned tensors).  Out of line
// on purpose: it is cold, and inlining it four times made the epilogue warps instruction-fetch bound.
__device__ __noinline__ void epilogue_slow_row(const GemmParams& p, const float* sp, long long orow, int n,
                                               bool add_bias, bool add_res, unsigned long long seed,
                                               unsigned long long offset, uint32_t thresh, float keep_scale) {
  const Philox8 rnd = (p.drop_p > 0.f)
                          ? philox8(seed, offset, p.drop_sid, (static_cast<unsigned long long>(orow) * p.N + n) >> 3)
                          : Philox8();
  for (int i = 0; i < 8; ++i) {
    if (n + i >= p.N) break;
    float v = sp[i] * p.alpha;
    if (add_bias) v += __ldg(p.bias + n + i);
    float rs = 0.f;
    if (add_res) {
      rs = p.res_fp32 ? reinterpret_cast<const float*>(p.residual)[orow * p.ldr + n + i]
                      : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.residual)[orow * p.ldr + n + i]);
    }
    if (p.res_first) v += rs;
    if (p.relu) v = fmaxf(v, 0.f);
    if (p.relu_mask != nullptr && !(__bfloat162float(p.relu_mask[orow * p.ldm + n + i]) > 0.f)) v = 0.f;
    if (p.drop_p > 0.f) v = (rnd.u16(i) < thresh) ? 0.f : v * keep_scale;
    if (!p.res_first) v += rs;
    if (p.out_fp32) {
      float* op = reinterpret_cast<float*>(p.out) + orow * p.ldo + n + i;
      if (p.atomic_out) atomicAdd(op, v);
      else *op = v;
    } else {
      reinterpret_cast<__nv_bfloat16*>(p.out)[orow * p.ldo + n + i] = __float2bfloat16_rn(v);
    }
  }
}

}  // namespace

// EPI: compile-time epilogue variant (bit 0 fp32 output, bits 1-2 residual: 0 none / 1 bf16 / 2 fp32, bit 3
// ReLU-mask, bit 4 dropout, bit 5 fp32 red.add output); EPI_GENERIC keeps every switch at run time.  The
// specialised variants exist because the fully general epilogue is ~70 KB of SASS and its eight warps
// were instruction-fetch bound (ncu: stall_no_inst on most epilogue instructions).
constexpr int EPI_GENERIC = 0xFF;
__host__ __device__ constexpr int epi_code(bool out_fp32, int res, bool mask, bool drop, bool atomic) {
  return (out_fp32 ? 1 : 0) | (res << 1) | (mask ? 8 : 0) | (drop ? 16 : 0) | (atomic ? 32 : 0);
}

template <int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ GemmParams p, const FastDiv fd_tiles_m, const FastDiv fd_tiles_n,
                    const int splits) {
  using C = Cfg<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + C::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * C::NBARS;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_gen + C::BAR_OFF + 8 * C::NBARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = static_cast<int>(fd_tiles_m.d * fd_tiles_n.d) * splits;

  // ---- one-time setup (overlaps the previous kernel's tail under programmatic dependent launch) ----
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), kEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();   // nothing above touches global memory; everything below may

  if (warp == 0) {
    // ===================================== TMA producer ======================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const TileCoord tc = tile_coord(p, t, fd_tiles_m, fd_tiles_n, splits, BN);
        for (int i = 0; i < tc.num_kb; ++i) {
          const int kb = tc.kb_begin + i;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + C::A_BYTES;
          const uint32_t fb = full_bar(stage);
          mbar_expect_tx(fb, static_cast<uint32_t>(p.stage_tx_bytes));
          // pixel box visited by this k-block when the contraction runs over pixels
          int kw0 = 0, kh0 = 0, kn0 = 0;
          if (p.a_mode == LOAD_PIXELS_MN || p.b_mode == LOAD_PIXELS_MN) {
            const int tw = kb % p.tiles_w;
            const int th = (kb / p.tiles_w) % p.tiles_h;
            const int tn = kb / (p.tiles_w * p.tiles_h);
            kw0 = tw * p.bx_w; kh0 = th * p.bx_h; kn0 = tn * p.bx_n;
          }
          // ---- A ----
          if (p.a_mode == LOAD_2D) {
            if (!p.a_mn) {
              tma_load_2d(sa, &tmA, fb, kb * BK, tc.m0);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_2d(sa + j * kChunkBytes, &tmA, fb, tc.m0 + 64 * j, kb * BK);
            }
          } else if (p.a_mode == LOAD_CONV) {
            const int tap = kb / p.cchunks, cc = kb - tap * p.cchunks;
            const int r = tap / p.taps_s, s = tap - r * p.taps_s;
            tma_load_4d(sa, &tmA, fb, cc * 64, tc.pw0 * p.stride_w - p.pad_w + s * p.dil_w,
                        tc.ph0 * p.stride_h - p.pad_h + r, tc.pn0);
          } else {  // LOAD_PIXELS_MN
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_4d(sa + j * kChunkBytes, &tmA, fb, tc.m0 + 64 * j, kw0, kh0, kn0);
          }
          // ---- B ----
          if (p.b_mode == LOAD_2D) {
            if (!p.b_mn) {
              tma_load_2d(sb, &tmB, fb, kb * BK, tc.n0);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_2d(sb + j * kChunkBytes, &tmB, fb, tc.n0 + 64 * j, kb * BK);
            }
          } else {  // LOAD_PIXELS_MN: column n -> (tap, input channel)
            const int tap = tc.n0 / p.b_tap_cin, ci0 = tc.n0 - tap * p.b_tap_cin;
            const int r = tap / p.b_taps_s, s = tap - r * p.b_taps_s;
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_4d(sb + j * kChunkBytes, &tmB, fb, ci0 + 64 * j, kw0 + s - p.pad_w,
                          kh0 + r - p.pad_h, kn0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ========================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(BM, BN, p.a_mn != 0, p.b_mn != 0);
      const uint32_t a_lbo = p.a_mn ? kChunkBytes : 16u, b_lbo = p.b_mn ? kChunkBytes : 16u;
      const uint32_t a_kstep = p.a_mn ? 16u * 128u : 32u, b_kstep = p.b_mn ? 16u * 128u : 32u;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const TileCoord tc = tile_coord(p, t, fd_tiles_m, fd_tiles_n, splits, BN);
        mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);   // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int i = 0; i < tc.num_kb; ++i) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
          const uint32_t sb = sa + C::A_BYTES;
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks) {
            const uint64_t da = umma_smem_desc(sa + ks * a_kstep, a_lbo, 1024u);
            const uint64_t db = umma_smem_desc(sb + ks * b_kstep, b_lbo, 1024u);
            umma_bf16(d_tmem, da, db, idesc, (i | ks) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));  // smem slot reusable once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tmem_full_bar(acc));  // accumulator complete (fires immediately when num_kb == 0)
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===================================== epilogue ==========================================
    const int ew = warp - 2;             // 0..7
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int half = ew >> 2;            // which alternating 32-column chunks this warp drains
    float* stg = reinterpret_cast<float*>(smem_gen + C::STG_OFF) + ew * kStgFloats;
    const int cg = lane & 3;             // 8-column group inside the 32-column chunk
    const int rsub = lane >> 2;          // row inside an 8-row slab

    constexpr bool G = (EPI == EPI_GENERIC);
    const bool f_out_fp32 = G ? (p.out_fp32 != 0) : ((EPI & 1) != 0);
    const bool f_has_res = G ? (p.residual != nullptr) : (((EPI >> 1) & 3) != 0);
    const bool f_res_fp32 = G ? (p.res_fp32 != 0) : (((EPI >> 1) & 3) == 2);
    const bool f_mask = G ? (p.relu_mask != nullptr) : ((EPI & 8) != 0);
    const bool f_drop = G ? (p.drop_p > 0.f) : ((EPI & 16) != 0);
    const bool f_atomic = G ? (p.atomic_out != 0) : ((EPI & 32) != 0);
    unsigned long long seed = 0, offset = 0;
    uint32_t thresh = 0;
    float keep_scale = 1.f;
    if (f_drop) {
      seed = p.rng[0]; offset = p.rng[1];
      thresh = drop_threshold(p.drop_p);
      keep_scale = 1.f / (1.f - p.drop_p);
    }
    const bool out_vec = f_out_fp32 ? ((p.ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0)
                                    : ((p.ldo & 7) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    const bool res_vec = !f_has_res ? false
                         : (f_res_fp32 ? ((p.ldr & 3) == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0)
                                       : ((p.ldr & 7) == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0));
    const bool msk_vec = f_mask && (p.ldm & 7) == 0 &&
                         (reinterpret_cast<uintptr_t>(p.relu_mask) & 15) == 0;

    // Prefetch registers: residual (bf16: first uint4; fp32: both) and ReLU-mask vectors of the NEXT chunk this
    // lane will finish.  They are issued one whole chunk ahead (across tile boundaries), so the HBM latency
    // of the epilogue's reads hides behind the TMEM drain, the math and the stores of the current chunk.
    struct Rows { long long row[4]; bool ok[4]; bool add_bias, add_res; int n0, num_kb; };
    uint4 pf_res[4][2];
    uint4 pf_msk[4];
    auto rows_of = [&](int t, Rows& r) {
      const TileCoord tc = tile_coord(p, t, fd_tiles_m, fd_tiles_n, splits, BN);
      const bool first_split = (tc.kb_begin == 0);
      r.add_bias = p.bias != nullptr && first_split;
      r.add_res = f_has_res && first_split;
      r.n0 = tc.n0;
      r.num_kb = tc.num_kb;
#pragma unroll
      for (int it = 0; it < 4; ++it) r.row[it] = output_row(p, tc, q * 32 + it * 8 + rsub, &r.ok[it]);
    };
    auto fast_ok = [&](int n) {
      return (n + 8) <= p.N && out_vec && (!f_has_res || res_vec) && (!f_mask || msk_vec);
    };
    auto prefetch = [&](const Rows& r, int c) {
      const int n = r.n0 + c * 32 + cg * 8;
      if (!(f_has_res || f_mask) || !fast_ok(n)) return;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        if (!r.ok[it]) continue;
        if (r.add_res) {
          if (f_res_fp32) {
            const uint4* rp = reinterpret_cast<const uint4*>(
                reinterpret_cast<const float*>(p.residual) + r.row[it] * p.ldr + n);
            pf_res[it][0] = __ldg(rp);
            pf_res[it][1] = __ldg(rp + 1);
          } else {
            pf_res[it][0] = __ldg(reinterpret_cast<const uint4*>(
                reinterpret_cast<const __nv_bfloat16*>(p.residual) + r.row[it] * p.ldr + n));
          }
        }
        if (f_mask) pf_msk[it] = __ldg(reinterpret_cast<const uint4*>(p.relu_mask + r.row[it] * p.ldm + n));
      }
    };

    // One tile ahead, pull the residual / mask lines this lane will read into L2 (fire-and-forget, no
    // registers): the register prefetch above then hits L2 instead of HBM, i.e. ~3x shorter latency for the
    // same bytes in flight (the epilogue-bound 1x1 convolutions were capped at ~2.2 TB/s by Little's law).
    auto l2_prefetch_tile = [&](const Rows& r) {
      if (!(f_has_res || f_mask)) return;
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        const int n = r.n0 + c * 32 + cg * 8;
        if (!fast_ok(n)) continue;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          if (!r.ok[it]) continue;
          if (r.add_res) {
            const char* a = f_res_fp32
                ? reinterpret_cast<const char*>(reinterpret_cast<const float*>(p.residual) + r.row[it] * p.ldr + n)
                : reinterpret_cast<const char*>(reinterpret_cast<const __nv_bfloat16*>(p.residual) + r.row[it] * p.ldr + n);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
          }
          if (f_mask) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.relu_mask + r.row[it] * p.ldm + n));
        }
      }
    };

    int acc = 0;
    uint32_t acc_phase = 0;
    Rows cur;
    int t = blockIdx.x;
    if (t < total_tiles) {
      rows_of(t, cur);
      prefetch(cur, half);
    }
    for (; t < total_tiles; t += gridDim.x) {
      mbar_wait(tmem_full_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t t_tile = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * BN);
      Rows nxt = cur;

#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        const int nc0 = cur.n0 + c * 32;
        const bool last_chunk = (c + 2 >= BN / 32);
        uint32_t accr[32];
        if (cur.num_kb > 0) {
          tmem_ld_32x32(t_tile + static_cast<uint32_t>(c * 32), accr);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) accr[i] = 0u;
        }
        if (last_chunk) {
          // all of this warp's TMEM reads of the tile are done: hand the accumulator stage back
          tc_fence_before();
          if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
          if (t + static_cast<int>(gridDim.x) < total_tiles) {
            rows_of(t + gridDim.x, nxt);
            l2_prefetch_tile(nxt);
          }
        }
        const int n = nc0 + cg * 8;
        const bool active = nc0 < p.N;            // warp-uniform
        if (active) {
          // ---- transpose through shared memory: lane (= tile row) writes its 32 columns ----
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(stg + lane * kStgStride + 4 * j) =
                make_uint4(accr[4 * j], accr[4 * j + 1], accr[4 * j + 2], accr[4 * j + 3]);
        }
        __syncwarp();
        const bool fast = active && n < p.N && fast_ok(n);
        float v[4][8];
        if (fast) {
          float bias8[8];
          if (cur.add_bias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n) + 1);
            bias8[0] = b0.x; bias8[1] = b0.y; bias8[2] = b0.z; bias8[3] = b0.w;
            bias8[4] = b1.x; bias8[5] = b1.y; bias8[6] = b1.z; bias8[7] = b1.w;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) bias8[i] = 0.f;
          }
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            if (!cur.ok[it]) continue;
            {
              const float* sp = stg + (it * 8 + rsub) * kStgStride + cg * 8;
              const float4 a0 = *reinterpret_cast<const float4*>(sp);
              const float4 a1 = *reinterpret_cast<const float4*>(sp + 4);
              v[it][0] = a0.x; v[it][1] = a0.y; v[it][2] = a0.z; v[it][3] = a0.w;
              v[it][4] = a1.x; v[it][5] = a1.y; v[it][6] = a1.z; v[it][7] = a1.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) v[it][i] = v[it][i] * p.alpha + bias8[i];
            float rs[8];
            if (cur.add_res) {
              if (f_res_fp32) {
                rs[0] = __uint_as_float(pf_res[it][0].x); rs[1] = __uint_as_float(pf_res[it][0].y);
                rs[2] = __uint_as_float(pf_res[it][0].z); rs[3] = __uint_as_float(pf_res[it][0].w);
                rs[4] = __uint_as_float(pf_res[it][1].x); rs[5] = __uint_as_float(pf_res[it][1].y);
                rs[6] = __uint_as_float(pf_res[it][1].z); rs[7] = __uint_as_float(pf_res[it][1].w);
              } else {
                const uint32_t rw[4] = {pf_res[it][0].x, pf_res[it][0].y, pf_res[it][0].z, pf_res[it][0].w};
#pragma unroll
                for (int i = 0; i < 8; ++i) rs[i] = bf16_bits_to_float((rw[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) rs[i] = 0.f;
            }
            if (p.res_first) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[it][i] += rs[i];
            }
            if (p.relu) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[it][i] = fmaxf(v[it][i], 0.f);
            }
            if (f_mask) {
              const uint32_t mw[4] = {pf_msk[it].x, pf_msk[it].y, pf_msk[it].z, pf_msk[it].w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float mf = bf16_bits_to_float((mw[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu);
                if (!(mf > 0.f)) v[it][i] = 0.f;
              }
            }
            if (f_drop) {
              const unsigned long long idx = static_cast<unsigned long long>(cur.row[it]) * p.N + n;
              const Philox8 rnd = philox8(seed, offset, p.drop_sid, idx >> 3);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[it][i] = (rnd.u16(i) < thresh) ? 0.f : v[it][i] * keep_scale;
            }
            if (!p.res_first) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[it][i] += rs[i];
            }
          }
        }
        // the prefetch registers are consumed: issue the next chunk's loads before this chunk's stores
        if (!last_chunk) prefetch(cur, c + 2);
        else if (t + static_cast<int>(gridDim.x) < total_tiles) prefetch(nxt, half);

        if (fast) {
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            if (!cur.ok[it]) continue;
            const long long orow = cur.row[it];
            if (f_out_fp32) {
              float* op = reinterpret_cast<float*>(p.out) + orow * p.ldo + n;
              if (f_atomic) {
#pragma unroll
                for (int i = 0; i < 8; ++i) atomicAdd(op + i, v[it][i]);
              } else {
                reinterpret_cast<float4*>(op)[0] = make_float4(v[it][0], v[it][1], v[it][2], v[it][3]);
                reinterpret_cast<float4*>(op)[1] = make_float4(v[it][4], v[it][5], v[it][6], v[it][7]);
              }
            } else {
              uint4 pk;
              __nv_bfloat162 t2;
              t2 = __floats2bfloat162_rn(v[it][0], v[it][1]); pk.x = *reinterpret_cast<uint32_t*>(&t2);
              t2 = __floats2bfloat162_rn(v[it][2], v[it][3]); pk.y = *reinterpret_cast<uint32_t*>(&t2);
              t2 = __floats2bfloat162_rn(v[it][4], v[it][5]); pk.z = *reinterpret_cast<uint32_t*>(&t2);
              t2 = __floats2bfloat162_rn(v[it][6], v[it][7]); pk.w = *reinterpret_cast<uint32_t*>(&t2);
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + n) = pk;
            }
          }
        } else if (active && n < p.N) {
          // ragged column tail or unaligned tensors: scalar path, out of line (rare: N = 170 classifier)
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            if (!cur.ok[it]) continue;
            epilogue_slow_row(p, stg + (it * 8 + rsub) * kStgStride + cg * 8, cur.row[it], n, cur.add_bias,
                              cur.add_res, seed, offset, thresh, keep_scale);
          }
        }
        __syncwarp();  // staging buffer is rewritten by the next chunk
      }
      cur = nxt;
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }

  // ---- teardown --------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
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

template <int BN, int STAGES, int EPI>
int launch_impl(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int tiles_m, int tiles_n,
                int splits, cudaStream_t stream) {
  using C = Cfg<BN, STAGES>;
  static_assert(C::SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static bool attr_set = false;  // per-process; all devices share the same kernel image attributes
  cudaError_t e;
  if (!attr_set) {
    e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, STAGES, EPI>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  const int total = tiles_m * tiles_n * splits;
  const int grid = total < sm_count() ? total : sm_count();
  launch_pdl(gemm_tcgen05_kernel<BN, STAGES, EPI>, dim3(grid), dim3(kThreads), C::SMEM_BYTES, stream, tmA, tmB, p,
             make_fastdiv(tiles_m), make_fastdiv(tiles_n), splits);
  return static_cast<int>(cudaGetLastError());
}

// one translation unit per tile width instantiates these variants (compile time)
template <int BN, int STAGES>
int launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int tiles_m, int tiles_n,
              int splits, cudaStream_t stream) {
  const int res = p.residual == nullptr ? 0 : (p.res_fp32 ? 2 : 1);
  const int code = epi_code(p.out_fp32 != 0, res, p.relu_mask != nullptr, p.drop_p > 0.f, p.atomic_out != 0);
#define VQA_EPI_CASE(o, r, m, d, a)                                                                   \
  case epi_code(o, r, m, d, a):                                                                       \
    return launch_impl<BN, STAGES, epi_code(o, r, m, d, a)>(tmA, tmB, p, tiles_m, tiles_n, splits, stream);
  switch (code) {
    VQA_EPI_CASE(false, 0, false, false, false)  // bf16 out                      (qkv, convs, plain dgrad)
    VQA_EPI_CASE(false, 1, false, false, false)  // bf16 out + bf16 residual      (ResNet block tails)
    VQA_EPI_CASE(false, 0, false, true, false)   // bf16 out + dropout            (FFN hidden, training)
    VQA_EPI_CASE(false, 0, true, false, false)   // bf16 out + ReLU mask          (dgrad through ReLU, eval)
    VQA_EPI_CASE(false, 0, true, true, false)    // bf16 out + ReLU mask + dropout (training)
    VQA_EPI_CASE(true, 2, false, false, false)   // fp32 out + fp32 residual      (residual streams)
    VQA_EPI_CASE(true, 2, false, true, false)    //   ... + dropout
    VQA_EPI_CASE(true, 2, false, false, true)    //   ... accumulated with red.add
    VQA_EPI_CASE(true, 0, false, false, false)   // fp32 out                      (wgrad, logits)
    VQA_EPI_CASE(true, 0, false, false, true)    // fp32 red.add                  (split-K)
    default:
      return launch_impl<BN, STAGES, EPI_GENERIC>(tmA, tmB, p, tiles_m, tiles_n, splits, stream);
  }
#undef VQA_EPI_CASE
}

}  // namespace

}  // namespace vqa
