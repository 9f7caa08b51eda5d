// tcgen05 / TMEM / TMA GEMM for sm_100a:  D[M,N] = epilogue( A[M,K] * B[N,K]^T ), bf16 in, fp32 accumulate.
//
// One 128 x BN output tile per CTA (split-K along gridDim.z), 192 threads:
//   warp 0      TMA producer   (one elected lane; STAGES-deep ring of 128x64 A and BNx64 B tiles)
//   warp 1      TMEM allocator + MMA issuer (one elected lane issues tcgen05.mma, commits free slots)
//   warps 2..5  epilogue       (tcgen05.ld 32 lanes x 32 columns -> bias / ReLU / dropout / residual
//                               -> 128-bit global stores, or fp32 red.add for split-K)
// Operands may be K-major or MN-major in global memory (instruction-descriptor transpose bits), so
// forward (X W^T), dgrad (dY W) and wgrad (dY^T X) all run on this kernel without any transposed copy.
// A may also be an implicit-GEMM convolution operand: NHWC activations read through a 4-D tensor map,
// k-block -> (filter tap, 64-channel chunk), out-of-image taps zero-filled by TMA.
//
// Replaces, on the reference's hot path, every nn.Linear / nn.Conv2d / nn.ConvTranspose2d contraction
// (model/resnet_vqa_model.py:64-78,119-135,154; model/multi_head_vision_text_attn.py:31-34,92-93;
// hf T5 q/k/v/o/wi/wo; torchvision ResNet convs) and their autograd backward.
#include "gemm.cuh"
#include "ptx.cuh"
#include "rng.cuh"

namespace vqa {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kThreads = 192;
constexpr int kChunkBytes = BK * 128;  // one 64-wide MN-major chunk: 64 k-rows x 128 B

template <int BN, int STAGES>
struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = BAR_OFF + (2 * STAGES + 1) * 8 + 16 + 1024;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

__device__ __forceinline__ float bf16_bits_to_float(uint32_t lo16) { return __uint_as_float(lo16 << 16); }

struct TileCoord {
  int m0;          // first output row (linear) or, for pixel boxes, unused
  int n0;          // first output column
  int pw0, ph0, pn0;  // pixel-box origin (output coordinates)
};

}  // namespace

template <int BN, int STAGES>
__global__ void __launch_bounds__(kThreads, (BN <= 128 ? 2 : 1))
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmParams p) {
  using C = Cfg<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + C::BAR_OFF;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 1);
  uint32_t* tmem_slot_ptr =
      reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates ----------------------------------------------------------------------
  TileCoord tc;
  tc.n0 = blockIdx.y * BN;
  tc.m0 = blockIdx.x * BM;
  tc.pw0 = tc.ph0 = tc.pn0 = 0;
  if (p.a_mode == LOAD_CONV) {
    const int tw = blockIdx.x % p.tiles_w;
    const int th = (blockIdx.x / p.tiles_w) % p.tiles_h;
    const int tn = blockIdx.x / (p.tiles_w * p.tiles_h);
    tc.pw0 = tw * p.bx_w;
    tc.ph0 = th * p.bx_h;
    tc.pn0 = tn * p.bx_n;
  }
  // split-K range of k-blocks
  const int splits = gridDim.z;
  const int kb_begin = static_cast<int>((static_cast<long long>(p.kb_total) * blockIdx.z) / splits);
  const int kb_end = static_cast<int>((static_cast<long long>(p.kb_total) * (blockIdx.z + 1)) / splits);
  const int num_kb = kb_end - kb_begin;

  // ---- one-time setup ------------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
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

  if (warp == 0) {
    // ===================================== TMA producer ======================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < num_kb; ++i) {
        const int kb = kb_begin + i;
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
  } else if (warp == 1) {
    // ===================================== MMA issuer ========================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(BM, BN, p.a_mn != 0, p.b_mn != 0);
      uint32_t a_lbo = p.a_mn ? kChunkBytes : 16u, b_lbo = p.b_mn ? kChunkBytes : 16u;
      uint32_t a_sbo = 1024u, b_sbo = 1024u;
      if (p.a_mn && p.dbg_a_lbo) { a_lbo = p.dbg_a_lbo; a_sbo = p.dbg_a_sbo; }
      if (p.b_mn && p.dbg_b_lbo) { b_lbo = p.dbg_b_lbo; b_sbo = p.dbg_b_sbo; }
      const uint32_t a_kstep = p.a_mn ? 16u * 128u : 32u, b_kstep = p.b_mn ? 16u * 128u : 32u;
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < num_kb; ++i) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
        const uint32_t sb = sa + C::A_BYTES;
#pragma unroll
        for (int ks = 0; ks < BK / 16; ++ks) {
          const uint64_t da = umma_smem_desc(sa + ks * a_kstep, a_lbo, a_sbo);
          const uint64_t db = umma_smem_desc(sb + ks * b_kstep, b_lbo, b_sbo);
          umma_bf16(tmem_base, da, db, idesc, (i | ks) ? 1u : 0u);
        }
        umma_commit(empty_bar(stage));  // smem slot reusable once these MMAs retire
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ===================================== epilogue ==========================================
    const int q = warp & 3;              // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;       // row inside the tile
    bool row_ok;
    long long out_row;                   // linear output row index
    if (p.out_pixels) {
      const int wi = row % p.bx_w;
      const int hi = (row / p.bx_w) % p.bx_h;
      const int ni = row / (p.bx_w * p.bx_h);
      const int ow = tc.pw0 + wi, oh = tc.ph0 + hi, on = tc.pn0 + ni;
      row_ok = (ni < p.bx_n) && ow < p.Wo && oh < p.Ho && on < p.Nimg;
      out_row = (static_cast<long long>(on) * p.Ho + oh) * p.Wo + ow;
    } else {
      row_ok = (tc.m0 + row) < p.M;
      out_row = tc.m0 + row;
    }
    const bool add_bias = p.bias != nullptr && blockIdx.z == 0;
    const bool add_res = p.residual != nullptr && blockIdx.z == 0;
    unsigned long long seed = 0, offset = 0;
    uint32_t thresh = 0;
    float keep_scale = 1.f;
    if (p.drop_p > 0.f) {
      seed = p.rng[0]; offset = p.rng[1];
      thresh = drop_threshold(p.drop_p);
      keep_scale = 1.f / (1.f - p.drop_p);
    }
    const bool out_vec = p.out_fp32 ? ((p.ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0)
                                    : ((p.ldo & 7) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    const bool res_vec = p.residual == nullptr ? false
                         : (p.res_fp32 ? ((p.ldr & 3) == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0)
                                       : ((p.ldr & 7) == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0));
    const bool msk_vec = p.relu_mask != nullptr && (p.ldm & 7) == 0 &&
                         (reinterpret_cast<uintptr_t>(p.relu_mask) & 15) == 0;

    if (num_kb > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
      auto add_residual = [&](float (&v)[8], int n, bool full8) {
          if (p.res_fp32) {
            const float* rp = reinterpret_cast<const float*>(p.residual) + out_row * p.ldr + n;
            if (res_vec && full8) {
              const float4 r0 = __ldg(reinterpret_cast<const float4*>(rp));
              const float4 r1 = __ldg(reinterpret_cast<const float4*>(rp) + 1);
              v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
              v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (n + i < p.N) v[i] += rp[i];
            }
          } else {
            const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + out_row * p.ldr + n;
            if (res_vec && full8) {
              const uint4 rv = __ldg(reinterpret_cast<const uint4*>(rp));
              const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int i = 0; i < 8; ++i)
                v[i] += bf16_bits_to_float((rw[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (n + i < p.N) v[i] += __bfloat162float(rp[i]);
            }
          }
      };
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int nc0 = tc.n0 + c * 32;
      if (nc0 >= p.N) break;  // warp-uniform
      uint32_t acc[32];
      if (num_kb > 0) {
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c * 32), acc);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0u;
      }
      if (!row_ok) continue;
#pragma unroll
      for (int g = 0; g < 4; ++g) {  // groups of 8 columns
        const int n = nc0 + g * 8;
        if (n >= p.N) break;
        const bool full8 = (n + 8) <= p.N;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(acc[g * 8 + i]) * p.alpha;
        if (add_bias) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (n + i < p.N) v[i] += __ldg(p.bias + n + i);
        }
        if (p.res_first && add_res) add_residual(v, n, full8);
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (p.relu_mask != nullptr) {
          const __nv_bfloat16* mp = p.relu_mask + out_row * p.ldm + n;
          if (msk_vec && full8) {
            const uint4 mv = __ldg(reinterpret_cast<const uint4*>(mp));
            const uint32_t mw[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float mf = bf16_bits_to_float((mw[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu);
              if (!(mf > 0.f)) v[i] = 0.f;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (n + i < p.N && !(__bfloat162float(mp[i]) > 0.f)) v[i] = 0.f;
          }
        }
        if (p.drop_p > 0.f) {
          const unsigned long long idx = static_cast<unsigned long long>(out_row) * p.N + n;
          const Philox8 rnd = philox8(seed, offset, p.drop_sid, idx >> 3);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = (rnd.u16(i) < thresh) ? 0.f : v[i] * keep_scale;
        }
        if (!p.res_first && add_res) add_residual(v, n, full8);
        // ---- store ----
        if (p.out_fp32) {
          float* op = reinterpret_cast<float*>(p.out) + out_row * p.ldo + n;
          if (p.atomic_out) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (n + i < p.N) atomicAdd(op + i, v[i]);
          } else if (out_vec && full8) {
            reinterpret_cast<float4*>(op)[0] = make_float4(v[0], v[1], v[2], v[3]);
            reinterpret_cast<float4*>(op)[1] = make_float4(v[4], v[5], v[6], v[7]);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (n + i < p.N) op[i] = v[i];
          }
        } else {
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + out_row * p.ldo + n;
          if (out_vec && full8) {
            uint4 pk;
            __nv_bfloat162 t;
            t = __floats2bfloat162_rn(v[0], v[1]); pk.x = *reinterpret_cast<uint32_t*>(&t);
            t = __floats2bfloat162_rn(v[2], v[3]); pk.y = *reinterpret_cast<uint32_t*>(&t);
            t = __floats2bfloat162_rn(v[4], v[5]); pk.z = *reinterpret_cast<uint32_t*>(&t);
            t = __floats2bfloat162_rn(v[6], v[7]); pk.w = *reinterpret_cast<uint32_t*>(&t);
            *reinterpret_cast<uint4*>(op) = pk;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (n + i < p.N) op[i] = __float2bfloat16_rn(v[i]);
          }
        }
      }
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
template <int BN, int STAGES>
int launch_impl(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, dim3 grid,
                cudaStream_t stream) {
  using C = Cfg<BN, STAGES>;
  static bool attr_set = false;  // per-process; all devices share the same kernel image attributes
  cudaError_t e;
  if (!attr_set) {
    e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, STAGES>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_set = true;
  }
  gemm_tcgen05_kernel<BN, STAGES><<<grid, kThreads, C::SMEM_BYTES, stream>>>(tmA, tmB, p);
  return static_cast<int>(cudaGetLastError());
}
}  // namespace

int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, int bn,
                int split_k, cudaStream_t stream) {
  int tiles_m;
  if (p.a_mode == LOAD_CONV) {
    tiles_m = p.tiles_w * p.tiles_h * ((p.Nimg + p.bx_n - 1) / p.bx_n);
  } else {
    tiles_m = (p.M + BM - 1) / BM;
  }
  if (split_k < 1) split_k = 1;
  if (split_k > p.kb_total) split_k = p.kb_total > 0 ? p.kb_total : 1;
  dim3 grid(tiles_m, (p.N + bn - 1) / bn, split_k);
  switch (bn) {
    case 64:  return launch_impl<64, 4>(tmA, tmB, p, grid, stream);
    case 128: return launch_impl<128, 3>(tmA, tmB, p, grid, stream);
    case 256: return launch_impl<256, 4>(tmA, tmB, p, grid, stream);
    default:  return static_cast<int>(cudaErrorInvalidValue);
  }
}

}  // namespace vqa
