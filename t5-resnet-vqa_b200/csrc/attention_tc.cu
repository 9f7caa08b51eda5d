// Flash-style tcgen05 attention for short sequences (Lq, Lk <= 64): T5 self-attention (hf:308-334: no 1/sqrt(d),
// relative-position bias, key-padding mask) and the SGA self / guided attention over the 32 text tokens
// (model/multi_head_vision_text_attn.py:73-86: 1/sqrt(96) scale), forward and backward.
//
// One (batch, head) problem is only 32 x 32 x hd, far below one MMA, so a CTA packs FOUR (b, h) pairs into the
// 128 rows of one tcgen05.mma: the packed Q [128 x hd] times the packed K^T [hd x 128] gives a 128 x 128 score tile
// in TMEM whose four diagonal 32 x 32 blocks are the four problems (the off-diagonal blocks are never read; tensor
// FLOPs are free at this size, launches and bytes are not).  Thread t owns TMEM lane t = one query row: it pulls its
// 32 scores with one tcgen05.ld, does bias + mask + softmax + dropout entirely in registers (no shuffles), and writes
// its probabilities as bf16 into a block-diagonal P tile in shared memory (canonical 128B-swizzled K-major layout,
// off-diagonal blocks zero), which is the A operand of the second MMA  O = P V  (V is read MN-major, straight from
// the TMA tile).  Scores / probabilities never touch HBM: forward saves only the row max and 1/row-sum, backward
// recomputes S = Q K^T and P with one more MMA, then  dP = dO V^T,  dS = P (dP - rowsum(P dP)),
// dV = Pd^T dO,  dQ = scale dS K,  dK = scale dS^T Q  as three more MMAs over the same shared-memory tiles
// (transposes are free: the MN-major descriptor bit).
// Operands arrive by TMA (2-D boxes of 64 columns x SLOT rows per pair and 64-column chunk, 128B swizzle).
// SLOT = 32 (four pairs per CTA) when Lq, Lk <= 32; SLOT = 64 (two pairs per CTA, diagonal blocks of 64 x 64: the guided
// attention over the 49 / 64 vision tokens, model/multi_head_vision_text_attn.py:147-149 with y = vision) otherwise: a thread
// then walks its row's scores in two 32-column halves.
#include "../../include/vqa_b200.h"
#include "common.cuh"
#include "ptx.cuh"
#include "rng.cuh"
#include "tmap.cuh"

using namespace vqa;

namespace {

constexpr int kThreads = 128;
constexpr int kTile = 128 * 128;          // bytes of one 64-column chunk: 128 rows x 128 B
// SLOT = query rows / keys per (batch, head) pair inside the 128-row tile (32 or 64); 128 / SLOT pairs per CTA;
// one TMA box delivers SLOT rows x 64 bf16 = SLOT * 128 bytes
template <int SLOT> struct MaskT { using type = uint32_t; };
template <> struct MaskT<64> { using type = unsigned long long; };
constexpr float kMaskedScore = -3.4028234663852886e38f;  // torch.finfo(float32).min (hf additive mask)

// phase stamps are compiled only into the diagnostic library (build.py --debug), as in gemm_kernel.cuh
#ifdef VQA_GEMM_DEBUG
#define VQA_STAMP(k) do { if (a.dbg != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0) a.dbg[(threadIdx.x >> 5) * 16 + (k)] = clock64(); } while (0)
#else
#define VQA_STAMP(k) do { } while (0)
#endif

long long* g_attn_dbg = nullptr;

__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ uint32_t philox_u16(const Philox8& r, uint32_t e) {
  const uint32_t lo = (e & 2u) ? r.w[1] : r.w[0];
  const uint32_t hi = (e & 2u) ? r.w[3] : r.w[2];
  const uint32_t w = (e & 4u) ? hi : lo;
  return (w >> ((e & 1u) * 16u)) & 0xFFFFu;
}

// Bit j = probability j of the row starting at flat index `base` survives dropout.  Same stream convention as the
// SIMT kernel (attention.cu: philox group = flat index / 8, 16-bit lane = flat index % 8).
template <typename M>
__device__ __forceinline__ M keep_bits(const DropCtx& dc, unsigned long long base, int Lk) {
  if (!dc.on) return ~static_cast<M>(0);
  M keep = 0;
  if ((Lk & 7) == 0) {
#pragma unroll 1
    for (int g = 0; g * 8 < Lk; ++g) {   // rolled on purpose: run-once straight-line code is instruction-fetch bound
      const Philox8 r = philox8(dc.seed, dc.offset, dc.sid, (base >> 3) + g);
      uint32_t m = 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) m |= (r.u16(e) >= dc.thresh ? 1u : 0u) << e;
      keep |= static_cast<M>(m) << (g * 8);
    }
  } else {
    unsigned long long cur = ~0ull;
    Philox8 r;
    r.w[0] = r.w[1] = r.w[2] = r.w[3] = 0;
#pragma unroll 1
    for (int j = 0; j < Lk; ++j) {
      const unsigned long long idx = base + j;
      if ((idx >> 3) != cur) { cur = idx >> 3; r = philox8(dc.seed, dc.offset, dc.sid, cur); }
      keep |= static_cast<M>(philox_u16(r, static_cast<uint32_t>(idx & 7)) >= dc.thresh ? 1u : 0u) << j;
    }
  }
  return keep;
}

// 32 lanes x 8 consecutive fp32 columns (thread t of the warp gets lane base + t)
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

// One TMEM row of NC fp32 columns, scaled by `mul` -> bf16 -> global (16-byte stores).  32 columns per round trip
// (four loads in flight, one wait); rolled over the chunks because every CTA runs this exactly once.
template <int NC>
__device__ __forceinline__ void store_tmem_row(uint32_t taddr, __nv_bfloat16* dst, bool ok, float mul) {
#pragma unroll 1
  for (int c = 0; c < NC / 32; ++c) {
    uint32_t r[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k) tmem_ld_32x8(taddr + c * 32 + k * 8, r[k]);
    tmem_ld_wait();
    if (ok) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(r[k][0]) * mul, __uint_as_float(r[k][1]) * mul);
        u.y = pack_bf16x2(__uint_as_float(r[k][2]) * mul, __uint_as_float(r[k][3]) * mul);
        u.z = pack_bf16x2(__uint_as_float(r[k][4]) * mul, __uint_as_float(r[k][5]) * mul);
        u.w = pack_bf16x2(__uint_as_float(r[k][6]) * mul, __uint_as_float(r[k][7]) * mul);
        reinterpret_cast<uint4*>(dst)[c * 4 + k] = u;
      }
    }
  }
}

// The same row drain, but written out by the whole warp: every thread parks its 32-column segment (64 B of bf16) in a
// warp-private 2 KB scratch area (a dead operand tile), then lanes 4r..4r+3 write row r's segment, so one store instruction
// covers eight rows x 64 contiguous bytes instead of 32 rows x 16 bytes (the thread-per-row stores were ~4 us of the
// backward kernel's 16).  dst / ok are the calling thread's OWN row; the other rows' are fetched by shuffle.
template <int NC>
__device__ __forceinline__ void store_tmem_row_warp(uint32_t taddr, __nv_bfloat16* dst, bool ok, float mul,
                                                    uint8_t* scratch, int lane) {
#pragma unroll 1
  for (int c = 0; c < NC / 32; ++c) {
    uint32_t r[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k) tmem_ld_32x8(taddr + c * 32 + k * 8, r[k]);
    tmem_ld_wait();
    __syncwarp();                                  // the previous chunk has been read out of the scratch area
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint4 u;
      u.x = pack_bf16x2(__uint_as_float(r[k][0]) * mul, __uint_as_float(r[k][1]) * mul);
      u.y = pack_bf16x2(__uint_as_float(r[k][2]) * mul, __uint_as_float(r[k][3]) * mul);
      u.z = pack_bf16x2(__uint_as_float(r[k][4]) * mul, __uint_as_float(r[k][5]) * mul);
      u.w = pack_bf16x2(__uint_as_float(r[k][6]) * mul, __uint_as_float(r[k][7]) * mul);
      *reinterpret_cast<uint4*>(scratch + lane * 64 + ((k ^ ((lane >> 1) & 3)) << 4)) = u;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = lane + 32 * j, row = idx >> 2, k = idx & 3;
      const uint4 v = *reinterpret_cast<const uint4*>(scratch + row * 64 + ((k ^ ((row >> 1) & 3)) << 4));
      const long long p = __shfl_sync(0xffffffffu, static_cast<long long>(reinterpret_cast<uintptr_t>(dst)), row);
      const int okr = __shfl_sync(0xffffffffu, ok ? 1 : 0, row);
      if (okr) reinterpret_cast<uint4*>(static_cast<uintptr_t>(p))[c * 4 + k] = v;
    }
  }
}

// 8 bf16 of row `row`, columns col0 .. col0 + 8 (col0 a multiple of 8), of a [128 x 128] K-major tile (two 64-column chunks
// of 128 rows x 128 B, 128B swizzle).  The block-diagonal P / dS tiles: pair g owns columns SLOT*g .. SLOT*g + SLOT - 1.
__device__ __forceinline__ void store_diag8(uint8_t* tile, int row, int col0, const float (&v)[8]) {
  uint8_t* rowp = tile + (col0 >> 6) * kTile + (row >> 3) * 1024 + (row & 7) * 128;
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(rowp + ((((col0 & 63) >> 3) ^ (row & 7)) << 4)) = u;
}

template <int SLOT>
struct RowCtx {
  using M = typename MaskT<SLOT>::type;
  int b, h, i, g;
  bool pair_ok, row_ok;
  long long prow;        // (b*H + h)*Lq + i
  M in_range;            // bit j: key j < Lk
  M visible;             // bit j: key j not masked by key_mask
};

template <int SLOT>
__device__ __forceinline__ RowCtx<SLOT> row_ctx(int B, int H, int Lq, int Lk, const long long* key_mask) {
  using M = typename MaskT<SLOT>::type;
  RowCtx<SLOT> c;
  const int t = threadIdx.x, lane = t & 31;
  const int npairs = B * H;
  c.g = t / SLOT;
  c.i = t - c.g * SLOT;
  const int pr = blockIdx.x * (128 / SLOT) + c.g;
  c.pair_ok = pr < npairs;
  const int prc = c.pair_ok ? pr : npairs - 1;
  c.b = prc / H;
  c.h = prc - c.b * H;
  c.row_ok = c.pair_ok && c.i < Lq;
  c.prow = (static_cast<long long>(c.b) * H + c.h) * Lq + c.i;
  c.in_range = Lk >= SLOT ? ~static_cast<M>(0) : ((static_cast<M>(1) << Lk) - 1);
  // every warp builds the whole key mask of its pair itself: lane l looks at key l (and key l + 32)
  M vis_all = 0;
#pragma unroll
  for (int hlf = 0; hlf < SLOT / 32; ++hlf) {
    const int key = hlf * 32 + lane;
    bool vis = key < Lk;
    if (vis && key_mask != nullptr) vis = key_mask[static_cast<long long>(c.b) * Lk + key] != 0;
    vis_all |= static_cast<M>(__ballot_sync(0xffffffffu, vis)) << (hlf * 32);
  }
  c.visible = vis_all;
  return c;
}

// 32 of the row's bias values (keys 32*hlf .. 32*hlf + 31); for the first half fetched BEFORE the score MMA is waited for
// (L2 latency off the critical path).
__device__ __forceinline__ void load_bias_row(float4 (&bq)[8], const float* bias, int h, int i, int Lq, int Lk, int hlf = 0) {
#pragma unroll
  for (int q = 0; q < 8; ++q) bq[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias == nullptr) return;
  const float* brow = bias + (static_cast<long long>(h) * Lq + (i < Lq ? i : 0)) * Lk;
  const int j0 = hlf * 32;
  if ((Lk & 3) == 0 && (reinterpret_cast<uintptr_t>(brow) & 15) == 0) {
#pragma unroll
    for (int q = 0; q < 8; ++q)
      if (j0 + q * 4 < Lk) bq[q] = __ldg(reinterpret_cast<const float4*>(brow + j0) + q);
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (j0 + q * 4 + e < Lk) v[e] = __ldg(brow + j0 + q * 4 + e);
      bq[q] = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

// scores of 8 consecutive keys of one row: acc * scale + bias, masked keys -> finfo.min (keys >= Lk: caller ignores).
// vis: bit e = key e of the group is visible.
__device__ __forceinline__ void score8(float (&s)[8], const uint32_t (&acc)[8], uint32_t vis, float scale,
                                       const float4& b0, const float4& b1) {
  const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float v = __uint_as_float(acc[e]) * scale + bv[e];
    s[e] = ((vis >> e) & 1u) ? v : kMaskedScore;
  }
}

struct FwdP {
  int B, H, Lq, Lk;
  __nv_bfloat16* out; long long ldo;
  float* stats;
  const float* bias;
  const long long* key_mask;
  float scale, drop_p;
  uint32_t sid;
  const unsigned long long* rng;
  long long* dbg;   // bring-up: clock64 stamps of CTA 0 (vqa_debug_attn_timing), else null
};

template <int HD, int SLOT>
__global__ void __launch_bounds__(kThreads)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const FwdP a) {
  using M = typename MaskT<SLOT>::type;
  constexpr int kPairs = 128 / SLOT, kBox = SLOT * 128, NH = SLOT / 32;
  constexpr int CH = (HD + 63) / 64;   // 64-column chunks per operand row
  constexpr int KS = HD / 16;          // UMMA k-steps over the head dim
  constexpr uint32_t kCols = 256;      // TMEM: S [0,128), O [128, 128 + HD)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  // P (two tiles) lives in K's buffer when K has two tiles (hd = 96): K is dead once S = Q K^T has completed, which every
  // thread waits for before it touches P.  96 KB instead of 128 KB: two CTAs per SM, so the 256 CTAs of the guided
  // attention over 49 / 64 vision tokens are one wave instead of two.
  constexpr bool kAliasP = CH == 2;
  constexpr int kTiles = 3 * CH + (kAliasP ? 0 : 2);
  const uint32_t sQ = base, sK = sQ + CH * kTile, sV = sK + CH * kTile, sP = kAliasP ? sK : sV + CH * kTile;
  const uint32_t bar_tma = base + kTiles * kTile, bar_mma = bar_tma + 8, slot = bar_mma + 8;
  uint8_t* Pg = gen + (kAliasP ? CH : 3 * CH) * kTile;
  const uint32_t* slot_ptr = reinterpret_cast<const uint32_t*>(gen + kTiles * kTile + 16);
  const int t = threadIdx.x;
  // lane-0 broadcast: ptxas then knows the warp index is warp-uniform (uniform branches / uniform registers below)
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);

  VQA_STAMP(0);
  pdl_launch_dependents();
  if (t == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(slot, kCols);
    tmem_relinquish();
  }
  if (!kAliasP)
    for (int i = t; i < 2 * kTile / 16; i += kThreads) reinterpret_cast<uint4*>(Pg)[i] = make_uint4(0u, 0u, 0u, 0u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  VQA_STAMP(1);
  pdl_wait();
  VQA_STAMP(2);

  const int npairs = a.B * a.H;
  // one TMA box per lane of warp 0 (operand, pair, 64-column chunk): a single warp-level issue instead of 12..24 serial ones
  if (warp == 0) {
    if (t == 0) mbar_expect_tx(bar_tma, 3u * CH * kPairs * kBox);
    __syncwarp();
    if (t < 3 * CH * kPairs) {
      const int opnd = t / (CH * kPairs), rem = t - opnd * (CH * kPairs);
      const int g = rem / CH, c = rem - g * CH;
      int pr = blockIdx.x * kPairs + g;
      if (pr >= npairs) pr = npairs - 1;      // duplicate the last pair: finite data, results discarded
      const int b = pr / a.H, h = pr - b * a.H;
      const uint32_t dst = (opnd == 0 ? sQ : (opnd == 1 ? sK : sV)) + c * kTile + g * kBox;
      const CUtensorMap* tm = opnd == 0 ? &tmQ : (opnd == 1 ? &tmK : &tmV);
      tma_load_2d(dst, tm, bar_tma, h * HD + c * 64, b * (opnd == 0 ? a.Lq : a.Lk));
    }
  }
  __syncwarp();
  // per-row context (mask ballot, dropout bits, bias row) is computed while the TMA loads are in flight
  const RowCtx<SLOT> rc = row_ctx<SLOT>(a.B, a.H, a.Lq, a.Lk, a.key_mask);
  const DropCtx dc = drop_ctx(a.drop_p, a.sid, a.rng);
  const M keep = keep_bits<M>(dc, static_cast<unsigned long long>(rc.prow) * a.Lk, a.Lk);
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  float4 bq[8];
  load_bias_row(bq, a.bias, rc.h, rc.i, a.Lq, a.Lk, 0);
  if (warp == 0) {   // whole warp, one elected lane issues (operands stay in uniform registers)
    mbar_wait_lean(bar_tma, 0);
    tc_fence_after();
    VQA_STAMP(3);
    const uint32_t idesc = umma_idesc_bf16(128, 128, false, false);
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const uint32_t off = (ks >> 2) * kTile + (ks & 3) * 32;
      umma_bf16_e<1>(tmem, umma_smem_desc(sQ + off, 16u, 1024u), umma_smem_desc(sK + off, 16u, 1024u), idesc, ks ? 1u : 0u);
    }
    umma_commit_e<1>(bar_mma);
  }
  __syncwarp();

  VQA_STAMP(4);
  mbar_wait_lean(bar_mma, 0);
  tc_fence_after();
  VQA_STAMP(5);
  const uint32_t t_s = tmem + lane_base + rc.g * SLOT;     // this row's SLOT scores (diagonal block of the pair)
  float scs[SLOT];
  // pass 1: scores (scale, bias, key mask) and the row maximum; a row's scores come out of TMEM 32 columns per load
  float mx = -INFINITY;
#pragma unroll
  for (int hf = 0; hf < NH; ++hf) {
    if (hf * 32 >= a.Lk) break;
    uint32_t accs[32];
    tmem_ld_32x32(t_s + hf * 32, accs);
    if (hf > 0) load_bias_row(bq, a.bias, rc.h, rc.i, a.Lq, a.Lk, hf);
    tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j0 = hf * 32 + k * 8;
      if (j0 >= a.Lk) break;
      score8(*reinterpret_cast<float(*)[8]>(&scs[j0]), *reinterpret_cast<const uint32_t(*)[8]>(&accs[8 * k]),
             static_cast<uint32_t>(rc.visible >> j0), a.scale, bq[2 * k], bq[2 * k + 1]);
      const uint32_t inr = static_cast<uint32_t>(rc.in_range >> j0);
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if ((inr >> e) & 1u) mx = fmaxf(mx, scs[j0 + e]);
    }
  }
  if (kAliasP) {   // P shares K's buffer: clear this thread's row of both tiles now that S is complete
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint4* prow = reinterpret_cast<uint4*>(Pg + c * kTile + (t >> 3) * 1024 + (t & 7) * 128);
#pragma unroll
      for (int j = 0; j < 8; ++j) prow[j] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  // pass 2: exp, row sum, dropout, bf16 P (unnormalised: the 1/sum goes onto O, flash style)
  float sum = 0.f;
#pragma unroll
  for (int kk = 0; kk < SLOT / 8; ++kk) {
    const int j0 = kk * 8;
    if (j0 >= a.Lk) break;
    float sc[8];
    const uint32_t inr = static_cast<uint32_t>(rc.in_range >> j0), kp = static_cast<uint32_t>(keep >> j0);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float ex = ((inr >> e) & 1u) ? __expf(scs[j0 + e] - mx) : 0.f;
      sum += ex;
      sc[e] = ((kp >> e) & 1u) ? ex : 0.f;
    }
    store_diag8(Pg, t, rc.g * SLOT + j0, sc);
  }
  const float inv = 1.f / sum;
  if (rc.row_ok && a.stats != nullptr) {
    *reinterpret_cast<float2*>(a.stats + 2 * rc.prow) = make_float2(mx, inv);
  }
  VQA_STAMP(6);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  VQA_STAMP(7);
  if (warp == 0) {   // whole warp, one elected lane issues (operands stay in uniform registers)
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, HD, false, true);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint64_t da = umma_smem_desc(sP + (ks >> 2) * kTile + (ks & 3) * 32, 16u, 1024u);
      const uint64_t db = umma_smem_desc(sV + ks * 2048, kTile, 1024u);
      umma_bf16_e<1>(tmem + 128, da, db, idesc, ks ? 1u : 0u);
    }
    umma_commit_e<1>(bar_mma);
  }
  __syncwarp();
  mbar_wait_lean(bar_mma, 1);
  tc_fence_after();
  VQA_STAMP(8);
  __nv_bfloat16* op = a.out + (static_cast<long long>(rc.b) * a.Lq + rc.i) * a.ldo + rc.h * HD;
  store_tmem_row_warp<HD>(tmem + lane_base + 128, op, rc.row_ok, inv * dc.scale, gen + warp * 2048, t & 31);   // scratch: Q's tile (dead)

  VQA_STAMP(9);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, kCols);
  }
}

struct BwdP {
  int B, H, Lq, Lk;
  const float* stats;
  const float* bias;
  const long long* key_mask;
  __nv_bfloat16 *dq, *dk, *dv;
  long long lddq, lddk, lddv;
  float* dbias;
  float scale, drop_p;
  uint32_t sid;
  const unsigned long long* rng;
  long long* dbg;
};

template <int HD, int SLOT>
__global__ void __launch_bounds__(kThreads)
attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, const BwdP a) {
  using M = typename MaskT<SLOT>::type;
  constexpr int kPairs = 128 / SLOT, kBox = SLOT * 128, NH = SLOT / 32;
  constexpr int CH = (HD + 63) / 64;
  constexpr int KS = HD / 16;
  // TMEM: S [0,128), dP [128,256); once both are in registers their columns are recycled for the outputs
  constexpr uint32_t kCols = (3 * HD <= 256) ? 256 : 512;
  constexpr uint32_t cDV = (3 * HD <= 256) ? 0 : 256;
  constexpr uint32_t cDQ = (3 * HD <= 256) ? HD : 256 + HD;
  constexpr uint32_t cDK = (3 * HD <= 256) ? 2 * HD : 0;
  // Shared memory: Q, K, dO, Pd (2 tiles), dS (2 tiles), V - where the SECOND dS tile IS V's first tile: V is dead
  // once dP = dO V^T has completed, which every thread has waited for before it writes dS.  (4 CH + 3) tiles = 112 KB
  // for the T5 head size, so two CTAs share an SM and the 192 CTAs of the step's launch are one wave instead of two.
  // The dynamic shared memory of a kernel without static shared memory starts on a 1 KB boundary (checked below);
  // asking for alignment slack would push two CTAs over the SM's 228 KB.
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* gen = smem_raw;
  const uint32_t sQ = base, sK = sQ + CH * kTile, sdO = sK + CH * kTile;
  const uint32_t sPd = sdO + CH * kTile, sdS = sPd + 2 * kTile, sV = sdS + kTile;
  const uint32_t bar_tma = sV + CH * kTile, bar_mma = bar_tma + 8, slot = bar_mma + 8;
  uint8_t* Pdg = gen + 3 * CH * kTile;
  uint8_t* dSg = Pdg + 2 * kTile;
  const uint32_t* slot_ptr = reinterpret_cast<const uint32_t*>(gen + (4 * CH + 3) * kTile + 16);
  if ((base & 1023u) != 0u) __trap();
  const int t = threadIdx.x;
  // lane-0 broadcast: ptxas then knows the warp index is warp-uniform (uniform branches / uniform registers below)
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);

  VQA_STAMP(0);
  pdl_launch_dependents();
  if (t == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(slot, kCols);
    tmem_relinquish();
  }
  // zero Pd (2 tiles) and the first dS tile now; the second dS tile is V's buffer and is zeroed after dP (below)
  for (int i = t; i < 3 * kTile / 16; i += kThreads) reinterpret_cast<uint4*>(Pdg)[i] = make_uint4(0u, 0u, 0u, 0u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  VQA_STAMP(1);
  pdl_wait();
  VQA_STAMP(2);

  const int npairs = a.B * a.H;
  if (warp == 0) {     // one TMA box per lane: (operand, pair, 64-column chunk)
    if (t == 0) mbar_expect_tx(bar_tma, 4u * CH * kPairs * kBox);
    __syncwarp();
    if (t < 4 * CH * kPairs) {
      const int opnd = t / (CH * kPairs), rem = t - opnd * (CH * kPairs);
      const int g = rem / CH, c = rem - g * CH;
      int pr = blockIdx.x * kPairs + g;
      if (pr >= npairs) pr = npairs - 1;
      const int b = pr / a.H, h = pr - b * a.H;
      const uint32_t dst = (opnd == 0 ? sQ : (opnd == 1 ? sdO : (opnd == 2 ? sK : sV))) + c * kTile + g * kBox;
      const CUtensorMap* tm = opnd == 0 ? &tmQ : (opnd == 1 ? &tmdO : (opnd == 2 ? &tmK : &tmV));
      tma_load_2d(dst, tm, bar_tma, h * HD + c * 64, b * (opnd < 2 ? a.Lq : a.Lk));
    }
  }
  __syncwarp();
  const RowCtx<SLOT> rc = row_ctx<SLOT>(a.B, a.H, a.Lq, a.Lk, a.key_mask);
  const DropCtx dc = drop_ctx(a.drop_p, a.sid, a.rng);
  const M keep = keep_bits<M>(dc, static_cast<unsigned long long>(rc.prow) * a.Lk, a.Lk);
  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  float4 bq[8];
  load_bias_row(bq, a.bias, rc.h, rc.i, a.Lq, a.Lk, 0);
  float mx = 0.f, inv = 0.f;
  if (rc.row_ok) {
    const float2 st = *reinterpret_cast<const float2*>(a.stats + 2 * rc.prow);
    mx = st.x; inv = st.y;
  }
  if (warp == 0) {   // whole warp, one elected lane issues (operands stay in uniform registers)
    mbar_wait_lean(bar_tma, 0);
    tc_fence_after();
    VQA_STAMP(3);
    const uint32_t idesc = umma_idesc_bf16(128, 128, false, false);
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {     // S = Q K^T
      const uint32_t off = (ks >> 2) * kTile + (ks & 3) * 32;
      umma_bf16_e<1>(tmem, umma_smem_desc(sQ + off, 16u, 1024u), umma_smem_desc(sK + off, 16u, 1024u), idesc, ks ? 1u : 0u);
    }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {     // dP = dO V^T
      const uint32_t off = (ks >> 2) * kTile + (ks & 3) * 32;
      umma_bf16_e<1>(tmem + 128, umma_smem_desc(sdO + off, 16u, 1024u), umma_smem_desc(sV + off, 16u, 1024u), idesc,
                ks ? 1u : 0u);
    }
    umma_commit_e<1>(bar_mma);
  }
  __syncwarp();

  VQA_STAMP(4);
  mbar_wait_lean(bar_mma, 0);
  tc_fence_after();
  VQA_STAMP(5);
  const uint32_t t_s = tmem + lane_base + rc.g * SLOT;          // S block of this row
  const uint32_t t_dp = tmem + lane_base + 128 + rc.g * SLOT;   // dP block of this row
  // rows that do not exist (i >= Lq, pair beyond B*H) get P = 0 so they add nothing to dK / dV
  // S and dP rows: one 32-column TMEM load each per half; the probabilities and the masked, rescaled dP values then stay in
  // registers for pass 2
  float pjs[SLOT], dps[SLOT];
  // pass 1: rowdot = sum_j P_j dP_j (dP through the dropout mask)
  float rowdot = 0.f;
#pragma unroll
  for (int hf = 0; hf < NH; ++hf) {
    if (hf * 32 >= a.Lk) break;
    uint32_t accs[32], accd[32];
    tmem_ld_32x32(t_s + hf * 32, accs);
    tmem_ld_32x32(t_dp + hf * 32, accd);
    if (hf > 0) load_bias_row(bq, a.bias, rc.h, rc.i, a.Lq, a.Lk, hf);
    tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j0 = hf * 32 + k * 8;
      if (j0 >= a.Lk) break;
      float sc[8];
      score8(sc, *reinterpret_cast<const uint32_t(*)[8]>(&accs[8 * k]), static_cast<uint32_t>(rc.visible >> j0), a.scale,
             bq[2 * k], bq[2 * k + 1]);
      const uint32_t inr = rc.row_ok ? static_cast<uint32_t>(rc.in_range >> j0) : 0u, kp = static_cast<uint32_t>(keep >> j0);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float pj = ((inr >> e) & 1u) ? __expf(sc[e] - mx) * inv : 0.f;
        const float dp = ((kp >> e) & 1u) ? __uint_as_float(accd[8 * k + e]) * dc.scale : 0.f;
        pjs[j0 + e] = pj;
        dps[j0 + e] = dp;
        rowdot = fmaf(pj, dp, rowdot);
      }
    }
  }
  // pass 2: dS = P (dP - rowdot); bias gradient; bf16 operands of the three output MMAs
  {   // the second dS tile lives in V's buffer (dead since dP completed): clear this thread's row of it; its diagonal
      // block, if it has one there, is written below by the same thread
    uint4* vrow = reinterpret_cast<uint4*>(dSg + kTile + (t >> 3) * 1024 + (t & 7) * 128);
#pragma unroll
    for (int j = 0; j < 8; ++j) vrow[j] = make_uint4(0u, 0u, 0u, 0u);
  }
  float* dbrow = (a.dbias != nullptr && rc.row_ok) ? a.dbias + (static_cast<long long>(rc.h) * a.Lq + rc.i) * a.Lk : nullptr;
  const bool dbvec = dbrow != nullptr && (a.Lk & 3) == 0 && (reinterpret_cast<uintptr_t>(dbrow) & 15) == 0;
#pragma unroll
  for (int kk = 0; kk < SLOT / 8; ++kk) {
    const int j0 = kk * 8;
    if (j0 >= a.Lk) break;
    float sc[8], ds[8];
    const uint32_t kp = static_cast<uint32_t>(keep >> j0);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float pj = pjs[j0 + e];
      // dps already carries the dropout mask and its 1/(1-p): dropped keys contribute -P * rowdot, as before
      ds[e] = pj * (dps[j0 + e] - rowdot);
      sc[e] = ((kp >> e) & 1u) ? pj * dc.scale : 0.f;   // dropped probabilities (operand of dV)
    }
    if (dbrow != nullptr) {
      if (dbvec) {
        atomicAdd(reinterpret_cast<float4*>(dbrow) + 2 * kk, make_float4(ds[0], ds[1], ds[2], ds[3]));
        atomicAdd(reinterpret_cast<float4*>(dbrow) + 2 * kk + 1, make_float4(ds[4], ds[5], ds[6], ds[7]));
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (j0 + e < a.Lk) atomicAdd(dbrow + j0 + e, ds[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) ds[e] *= a.scale;   // both dQ and dK carry the score scale
    store_diag8(Pdg, t, rc.g * SLOT + j0, sc);
    store_diag8(dSg, t, rc.g * SLOT + j0, ds);
  }
  VQA_STAMP(6);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  VQA_STAMP(7);
  if (warp == 0) {   // whole warp, one elected lane issues (operands stay in uniform registers)
    tc_fence_after();
    const uint32_t id_tt = umma_idesc_bf16(128, HD, true, true);
    const uint32_t id_nt = umma_idesc_bf16(128, HD, false, true);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {      // dV[key, d] = sum_q Pd[q, key] dO[q, d]
      umma_bf16_e<1>(tmem + cDV, umma_smem_desc(sPd + ks * 2048, kTile, 1024u), umma_smem_desc(sdO + ks * 2048, kTile, 1024u),
                id_tt, ks ? 1u : 0u);
    }
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {      // dQ[q, d] = sum_key dS[q, key] K[key, d]
      umma_bf16_e<1>(tmem + cDQ, umma_smem_desc(sdS + (ks >> 2) * kTile + (ks & 3) * 32, 16u, 1024u),
                umma_smem_desc(sK + ks * 2048, kTile, 1024u), id_nt, ks ? 1u : 0u);
    }
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {      // dK[key, d] = sum_q dS[q, key] Q[q, d]
      umma_bf16_e<1>(tmem + cDK, umma_smem_desc(sdS + ks * 2048, kTile, 1024u), umma_smem_desc(sQ + ks * 2048, kTile, 1024u),
                id_tt, ks ? 1u : 0u);
    }
    umma_commit_e<1>(bar_mma);
  }
  __syncwarp();
  mbar_wait_lean(bar_mma, 1);
  tc_fence_after();
  VQA_STAMP(8);
  {
    __nv_bfloat16* qp = a.dq + (static_cast<long long>(rc.b) * a.Lq + rc.i) * a.lddq + rc.h * HD;
    // scratch: this warp's 2 KB of the Q tile (every operand tile is dead once the three output MMAs have completed)
    uint8_t* scratch = gen + warp * 2048;
    const int lane = t & 31;
    store_tmem_row_warp<HD>(tmem + lane_base + cDQ, qp, rc.row_ok, 1.f, scratch, lane);
    const bool key_ok = rc.pair_ok && rc.i < a.Lk;
    __nv_bfloat16* kp = a.dk + (static_cast<long long>(rc.b) * a.Lk + rc.i) * a.lddk + rc.h * HD;
    store_tmem_row_warp<HD>(tmem + lane_base + cDK, kp, key_ok, 1.f, scratch, lane);
    __nv_bfloat16* vp = a.dv + (static_cast<long long>(rc.b) * a.Lk + rc.i) * a.lddv + rc.h * HD;
    store_tmem_row_warp<HD>(tmem + lane_base + cDV, vp, key_ok, 1.f, scratch, lane);
  }
  VQA_STAMP(9);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, kCols);
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Long-sequence forward (the frozen ViT-B/16 of VitVQAModel, model/vit_vqa_model.py:184-186; vit: ViTSelfAttention):
// 197 tokens, hd = 64, softmax(Q K^T / 8) V, no mask / bias / dropout, nothing saved (the ViT runs under no_grad).
// One CTA per (batch, head, 128-query tile): S[128 x 256] = Q K^T is ONE tcgen05.mma chain (N = 256 keys: two 128-row K
// tiles, rows past the sequence hold the next sample's finite data or TMA zero fill and are masked in the softmax), the
// row softmax runs thread-per-row out of TMEM (two passes of 32-column loads), P goes to shared memory as the bf16 K-major
// A operand [128 x 256] (aliased onto the dead Q / K tiles: 96 KB per CTA, two CTAs per SM), O = P V reads V MN-major
// straight from its TMA tiles and lands in the recycled S columns.
// ---------------------------------------------------------------------------------------------------------------------
struct LongP {
  int B, H, L;
  __nv_bfloat16* out; long long ldo;
  float scale;
  float* probs;      // optional fp32 [B, H, L, L] softmax output (output_attentions=True of generate_answers), else null
};

__global__ void __launch_bounds__(kThreads)
attn_long_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const LongP a) {
  constexpr int HD = 64;
  constexpr uint32_t kCols = 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  // [Q | K0 | K1 | pad] = 4 tiles, re-used as P's four 64-column chunks; then V0 | V1
  const uint32_t sQ = base, sK = base + kTile, sP = base, sV = base + 4 * kTile;
  const uint32_t bar_tma = base + 6 * kTile, bar_mma = bar_tma + 8, slot = bar_mma + 8;
  const uint32_t* slot_ptr = reinterpret_cast<const uint32_t*>(gen + 6 * kTile + 16);
  const int t = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);
  const int pair = blockIdx.x >> 1, qt = blockIdx.x & 1;
  const int b = pair / a.H, h = pair - b * a.H;

  pdl_launch_dependents();
  if (t == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(slot, kCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot_ptr;
  pdl_wait();

  if (warp == 0) {
    if (t == 0) mbar_expect_tx(bar_tma, 5u * kTile);
    __syncwarp();
    if (t < 5) {
      const CUtensorMap* tm = t == 0 ? &tmQ : (t < 3 ? &tmK : &tmV);
      const uint32_t dst = t == 0 ? sQ : (t < 3 ? sK + (t - 1) * kTile : sV + (t - 3) * kTile);
      const int row = b * a.L + (t == 0 ? qt * 128 : ((t - 1) & 1) * 128);
      tma_load_2d(dst, tm, bar_tma, h * HD, row);
    }
    __syncwarp();
    mbar_wait_lean(bar_tma, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, 256, false, false);
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks)
      umma_bf16_e<1>(tmem, umma_smem_desc(sQ + ks * 32, 16u, 1024u), umma_smem_desc(sK + ks * 32, 16u, 1024u), idesc,
                     ks ? 1u : 0u);
    umma_commit_e<1>(bar_mma);
  }
  __syncwarp();
  mbar_wait_lean(bar_mma, 0);
  tc_fence_after();

  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  const int qi = qt * 128 + t;
  const bool row_ok = qi < a.L;
  const int nch = (a.L + 31) >> 5;
  // pass 1: row maximum of the raw scores (scale > 0)
  float mx = -INFINITY;
#pragma unroll 1
  for (int c = 0; c < nch; ++c) {
    uint32_t acc[32];
    tmem_ld_32x32(tmem + lane_base + c * 32, acc);
    tmem_ld_wait();
#pragma unroll
    for (int e = 0; e < 32; ++e)
      if (c * 32 + e < a.L) mx = fmaxf(mx, __uint_as_float(acc[e]));
  }
  // pass 2: exp, row sum, bf16 P (unnormalised; 1/sum goes onto O).  Every thread writes all 256 columns of its row.
  const float sl2 = a.scale * 1.4426950408889634f;
  float sum = 0.f;
  uint8_t* prow = gen + (t >> 3) * 1024 + (t & 7) * 128;
  const int sw = t & 7;
#pragma unroll 1
  for (int c = 0; c < 8; ++c) {
    float p[32];
    if (c < nch) {
      uint32_t acc[32];
      tmem_ld_32x32(tmem + lane_base + c * 32, acc);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const float ex = (c * 32 + e < a.L) ? exp2f((__uint_as_float(acc[e]) - mx) * sl2) : 0.f;
        sum += ex;
        p[e] = ex;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e) p[e] = 0.f;
    }
    uint8_t* tile = prow + (c >> 1) * kTile;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint4 u;
      u.x = pack_bf16x2(p[8 * k + 0], p[8 * k + 1]); u.y = pack_bf16x2(p[8 * k + 2], p[8 * k + 3]);
      u.z = pack_bf16x2(p[8 * k + 4], p[8 * k + 5]); u.w = pack_bf16x2(p[8 * k + 6], p[8 * k + 7]);
      *reinterpret_cast<uint4*>(tile + ((((c & 1) * 4 + k) ^ sw) << 4)) = u;
    }
  }
  if (a.probs != nullptr) {
    // pass 3 (heat-map path only): the normalised probabilities of this row, from the scores still in TMEM.  The TMEM load
    // is warp-collective (.sync.aligned): every lane runs it, only rows inside the sequence store.
    const float inv = 1.f / sum;
    float* pr = a.probs + ((static_cast<long long>(b) * a.H + h) * a.L + (row_ok ? qi : 0)) * a.L;
#pragma unroll 1
    for (int c = 0; c < nch; ++c) {
      uint32_t acc[32];
      tmem_ld_32x32(tmem + lane_base + c * 32, acc);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int e = 0; e < 32; ++e)
          if (c * 32 + e < a.L) pr[c * 32 + e] = exp2f((__uint_as_float(acc[e]) - mx) * sl2) * inv;
      }
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, HD, false, true);
    const int nks = (a.L + 15) >> 4;
#pragma unroll 1
    for (int ks = 0; ks < nks; ++ks) {
      const uint64_t da = umma_smem_desc(sP + (ks >> 2) * kTile + (ks & 3) * 32, 16u, 1024u);
      const uint64_t db = umma_smem_desc(sV + ks * 2048, kTile, 1024u);
      umma_bf16_e<1>(tmem, da, db, idesc, ks ? 1u : 0u);
    }
    umma_commit_e<1>(bar_mma);
  }
  __syncwarp();
  mbar_wait_lean(bar_mma, 1);
  tc_fence_after();
  __nv_bfloat16* op = a.out + (static_cast<long long>(b) * a.L + (row_ok ? qi : 0)) * a.ldo + h * HD;
  store_tmem_row_warp<HD>(tmem + lane_base, op, row_ok, 1.f / sum, gen + warp * 2048, t & 31);   // scratch: P (dead)
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, kCols);
  }
}

constexpr size_t kLongSmem = 6 * kTile + 64 + 1024;

template <int HD> constexpr size_t fwd_smem_bytes() {
  return (3 * ((HD + 63) / 64) + ((HD + 63) / 64 == 2 ? 0 : 2)) * kTile + 64 + 1024;   // hd = 96: P aliases K
}
template <int HD> constexpr size_t bwd_smem_bytes() { return (4 * ((HD + 63) / 64) + 3) * kTile + 64; }

template <typename K>
int raise_smem(K kern, size_t bytes, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  if (e != cudaSuccess) { set_last_error("%s: %s", what, cudaGetErrorString(e)); return static_cast<int>(e); }
  return 0;
}

int operand_map(CUtensorMap* tm, const void* base, long long rows, int H, int hd, long long ld, int slot, const char* what) {
  if (reinterpret_cast<uintptr_t>(base) & 15) { set_last_error("%s: operand pointers must be 16-byte aligned", what); return -1; }
  return make_tmap_2d(tm, base, static_cast<uint64_t>(rows), static_cast<uint64_t>(H) * hd, static_cast<uint64_t>(ld), 64, slot);
}

inline int slot_for(int Lq, int Lk) { return (Lq <= 32 && Lk <= 32) ? 32 : 64; }

template <int HD, int SLOT>
int launch_fwd(int grid, cudaStream_t s, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const FwdP& a) {
  static bool attr = false;
  if (!attr) { if (raise_smem(attn_tc_fwd_kernel<HD, SLOT>, fwd_smem_bytes<HD>(), "attention_fwd")) return -1; attr = true; }
  launch_pdl(attn_tc_fwd_kernel<HD, SLOT>, dim3(grid), dim3(kThreads), fwd_smem_bytes<HD>(), s, tq, tk, tv, a);
  return launch_status("attention_fwd");
}

template <int HD, int SLOT>
int launch_bwd(int grid, cudaStream_t s, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
               const CUtensorMap& tdo, const BwdP& a) {
  static bool attr = false;
  if (!attr) { if (raise_smem(attn_tc_bwd_kernel<HD, SLOT>, bwd_smem_bytes<HD>(), "attention_bwd")) return -1; attr = true; }
  launch_pdl(attn_tc_bwd_kernel<HD, SLOT>, dim3(grid), dim3(kThreads), bwd_smem_bytes<HD>(), s, tq, tk, tv, tdo, a);
  return launch_status("attention_bwd");
}

}  // namespace

namespace vqa {

void attention_tc_debug(long long* buf) { g_attn_dbg = buf; }

bool attention_tc_supported(int Lq, int Lk, int hd) {
  return Lq >= 1 && Lq <= 64 && Lk >= 1 && Lk <= 64 && (hd == 64 || hd == 96);
}

int attention_tc_fwd(void* plan, const vqa_attn_fwd_args* x, void* stream) {
  CUtensorMap tq, tk, tv;
  const int slot = slot_for(x->Lq, x->Lk);
  if (operand_map(&tq, x->q, static_cast<long long>(x->B) * x->Lq, x->H, x->hd, x->ldq, slot, "attention_fwd")) return -1;
  if (operand_map(&tk, x->k, static_cast<long long>(x->B) * x->Lk, x->H, x->hd, x->ldk, slot, "attention_fwd")) return -1;
  if (operand_map(&tv, x->v, static_cast<long long>(x->B) * x->Lk, x->H, x->hd, x->ldv, slot, "attention_fwd")) return -1;
  FwdP a;
  a.B = x->B; a.H = x->H; a.Lq = x->Lq; a.Lk = x->Lk;
  a.out = static_cast<__nv_bfloat16*>(x->out); a.ldo = x->ldo;
  a.stats = x->stats; a.bias = x->bias; a.key_mask = x->key_mask;
  a.scale = x->scale; a.drop_p = x->drop_p; a.sid = x->sid;
  a.rng = reinterpret_cast<const unsigned long long*>(x->rng);
  a.dbg = g_attn_dbg;
  const int hd = x->hd;
  const int grid = (a.B * a.H + 128 / slot - 1) / (128 / slot);
  const double fl = 4.0 * a.B * a.H * a.Lq * a.Lk * hd;
  note_op("attention_fwd", fl, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    if (hd == 64) return slot == 32 ? launch_fwd<64, 32>(grid, s, tq, tk, tv, a) : launch_fwd<64, 64>(grid, s, tq, tk, tv, a);
    return slot == 32 ? launch_fwd<96, 32>(grid, s, tq, tk, tv, a) : launch_fwd<96, 64>(grid, s, tq, tk, tv, a);
  });
}

int attention_tc_bwd(void* plan, const vqa_attn_bwd_args* x, void* stream) {
  CUtensorMap tq, tk, tv, tdo;
  const int slot = slot_for(x->Lq, x->Lk);
  if (operand_map(&tq, x->q, static_cast<long long>(x->B) * x->Lq, x->H, x->hd, x->ldq, slot, "attention_bwd")) return -1;
  if (operand_map(&tk, x->k, static_cast<long long>(x->B) * x->Lk, x->H, x->hd, x->ldk, slot, "attention_bwd")) return -1;
  if (operand_map(&tv, x->v, static_cast<long long>(x->B) * x->Lk, x->H, x->hd, x->ldv, slot, "attention_bwd")) return -1;
  if (operand_map(&tdo, x->dout, static_cast<long long>(x->B) * x->Lq, x->H, x->hd, x->ldo, slot, "attention_bwd")) return -1;
  BwdP a;
  a.B = x->B; a.H = x->H; a.Lq = x->Lq; a.Lk = x->Lk;
  a.stats = x->stats; a.bias = x->bias; a.key_mask = x->key_mask;
  a.dq = static_cast<__nv_bfloat16*>(x->dq); a.dk = static_cast<__nv_bfloat16*>(x->dk);
  a.dv = static_cast<__nv_bfloat16*>(x->dv);
  a.lddq = x->lddq; a.lddk = x->lddk; a.lddv = x->lddv;
  a.dbias = x->dbias; a.scale = x->scale; a.drop_p = x->drop_p; a.sid = x->sid;
  a.rng = reinterpret_cast<const unsigned long long*>(x->rng);
  a.dbg = g_attn_dbg;
  const int hd = x->hd;
  const int grid = (a.B * a.H + 128 / slot - 1) / (128 / slot);
  const double fl = 10.0 * a.B * a.H * a.Lq * a.Lk * hd;
  note_op("attention_bwd", fl, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    if (hd == 64) return slot == 32 ? launch_bwd<64, 32>(grid, s, tq, tk, tv, tdo, a) : launch_bwd<64, 64>(grid, s, tq, tk, tv, tdo, a);
    return slot == 32 ? launch_bwd<96, 32>(grid, s, tq, tk, tv, tdo, a) : launch_bwd<96, 64>(grid, s, tq, tk, tv, tdo, a);
  });
}

}  // namespace vqa

extern "C" int vqa_attention_long_fwd(void* plan, const void* q, long long ldq, const void* k, long long ldk, const void* v,
                                      long long ldv, void* out, long long ldo, int B, int H, int L, int hd, float scale,
                                      float* probs, void* stream) {
  using namespace vqa;
  if (hd != 64 || L < 1 || L > 256 || B < 1 || H < 1 || !(scale > 0.f)) {
    set_last_error("attention_long_fwd: hd must be 64, 1 <= L <= 256, scale > 0");
    return -1;
  }
  if ((ldq | ldk | ldv | ldo) & 7) { set_last_error("attention_long_fwd: strides must be multiples of 8"); return -1; }
  if (reinterpret_cast<uintptr_t>(out) & 15) { set_last_error("attention_long_fwd: out must be 16-byte aligned"); return -1; }
  CUtensorMap tq, tk, tv;
  const long long rows = static_cast<long long>(B) * L;
  if (operand_map(&tq, q, rows, H, hd, ldq, 128, "attention_long_fwd")) return -1;
  if (operand_map(&tk, k, rows, H, hd, ldk, 128, "attention_long_fwd")) return -1;
  if (operand_map(&tv, v, rows, H, hd, ldv, 128, "attention_long_fwd")) return -1;
  LongP a;
  a.B = B; a.H = H; a.L = L; a.out = static_cast<__nv_bfloat16*>(out); a.ldo = ldo; a.scale = scale; a.probs = probs;
  note_op("attention_long_fwd", 4.0 * B * H * static_cast<double>(L) * L * hd, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    static bool attr = false;
    if (!attr) { if (raise_smem(attn_long_fwd_kernel, kLongSmem, "attention_long_fwd")) return -1; attr = true; }
    launch_pdl(attn_long_fwd_kernel, dim3(B * H * 2), dim3(kThreads), kLongSmem, s, tq, tk, tv, a);
    return launch_status("attention_long_fwd");
  });
}
