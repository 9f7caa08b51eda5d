// Host-visible description of one tcgen05 GEMM / implicit-GEMM launch.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vqa {

enum : int {
  LOAD_2D = 0,         // plain matrix.  K-major: box (64 k, rows); MN-major: box (64 mn, 64 k)
  LOAD_CONV = 1,       // K-major A of an implicit-GEMM convolution: NHWC activations through a 4-D
                       // tensor map; k-block -> (filter tap, 64-channel chunk); tile rows = pixel box
  LOAD_PIXELS_MN = 2,  // MN-major operand whose contraction index runs over pixels (wgrad of a
                       // convolution): k-block -> pixel box (n, th, tw); 64-channel chunks along MN
};

// Division by a run-time constant as a multiply-high + shift (the tile scheduler and the pixel-box row mapping
// divide by launch constants once per tile per thread; a hardware-free `/` costs ~25 instructions each).
// Exact for 0 <= n < 2^31 and 1 <= d < 2^31.
struct FastDiv {
  uint32_t d, mul, shr;
};
inline FastDiv make_fastdiv(int dv) {
  FastDiv f;
  f.d = static_cast<uint32_t>(dv < 1 ? 1 : dv);
  if (f.d == 1) { f.mul = 0; f.shr = 0; return f; }
  uint32_t l = 0;
  while ((1ull << l) < f.d) ++l;                       // ceil(log2 d)
  const unsigned long long m = ((1ull << (31 + l)) + f.d - 1) / f.d;   // ceil(2^(31+l) / d)  (< 2^32)
  f.mul = static_cast<uint32_t>(m);
  f.shr = l - 1;                                        // q = (n * m) >> (31 + l) = mulhi(n, m) >> (l - 1)
  return f;
}
#ifdef __CUDACC__
__device__ __forceinline__ int fast_div(int n, const FastDiv& f) {
  return f.d == 1 ? n : static_cast<int>(__umulhi(static_cast<uint32_t>(n), f.mul) >> f.shr);
}
#endif

struct GemmParams {
  int M, N;        // output extent (rows, cols); for LOAD_CONV M is implied by the pixel boxes
  int kb_total;    // number of 64-deep k-blocks in the whole contraction
  int a_mn, b_mn;  // 1 = operand is MN-major in global memory (contraction index is the row index)
  int a_mode, b_mode;
  int stage_tx_bytes;  // bytes the TMA loads of one pipeline stage deliver (A box + B boxes)
  int res_tx_bytes;    // bytes of one bf16 residual panel load (rows of the tile x 128 B)

  // pixel-box geometry (LOAD_CONV for A; LOAD_PIXELS_MN for A and/or B)
  int bx_w, bx_h, bx_n;        // box extent in output pixels (w, h, images); product = rows per box
  int tiles_w, tiles_h;        // boxes per image along w and h
  FastDiv fd_tiles_w, fd_tiles_h, fd_bx_w, fd_bx_h;   // divisors of the fields above
  int Wo, Ho, Nimg;            // output spatial size and image count
  int taps_s, cchunks;         // LOAD_CONV: filter width S and 64-channel chunks per tap
  int stride_w, stride_h, pad_w, pad_h, dil_w;  // input coord = out*stride - pad + tap*dil
  int b_tap_cin;               // LOAD_PIXELS_MN on B: channels per tap (column n -> tap n / cin)
  int b_taps_s;                //                      filter width used to split the tap into (r, s)

  // epilogue
  void* out;                   // bf16 or fp32, row stride ldo elements
  long long ldo;
  int out_fp32;
  int out_pixels;              // 1: output rows follow the LOAD_CONV pixel boxes (NHWC), else linear
  int atomic_out;              // split-K partial sums: fp32 red.add into a pre-zeroed output
  const float* bias;           // [N] or null (added by split 0 only)
  int relu;
  const __nv_bfloat16* relu_mask;  // [M, ldm]: keep element only where mask > 0 (backward of ReLU)
  long long ldm;
  float drop_p;                // dropout probability (0 = off); mask from (rng seed, drop_sid, index)
  uint32_t drop_sid;
  const unsigned long long* rng;   // device pointer: {seed, offset}
  const void* residual;        // added after dropout; fp32 or bf16, row stride ldr
  long long ldr;
  int res_fp32;
  int res_first;              // 1: add the residual before ReLU (ResNet block), else after dropout
  float alpha;                 // scale applied to the accumulator before everything else
  int ksplit;                  // > 1: cluster split-K: `ksplit` CTAs (one cluster) share an output tile, each contracts a
                               // k-slice, exchanges partial accumulators through ks_ws and finishes a column range
  float* ks_ws;                // workspace: clusters x ksplit x (bn/32) x 128 x 32 fp32
  // two-term operand split (plain K-major GEMMs only): the k-loop runs nseg segments of kseg k-blocks each over the SAME
  // output tile: (A, B), [(A_lo, B),] (A, B_lo).  A_lo = columns [a_lo_col, a_lo_col + K) of A's rows; B_lo is read
  // through the residual tensor map (a split launch has no bf16 residual).  nseg <= 1: off.
  int nseg, kseg, a_lo_col;
  int max_ctas;                // > 0: the persistent grid is capped at this many CTAs
  int dbg_mode;                // bring-up: 1 = skip the MMAs, 2 = skip the loads (VQA_B200_GEMM_DBG)
  long long* dbg_clk;          // bring-up: clock64 stamps of CTA 0 (vqa_debug_gemm_timing), else null
  int dbg_a_lbo, dbg_a_sbo, dbg_b_lbo, dbg_b_sbo;  // bring-up overrides of the MN-major descriptor strides (0 = default)
};

// Launch. tmA / tmB are host-encoded CUtensorMaps (copied into kernel parameter space).
// bn in {64, 128, 256}; split_k >= 1.  Returns cudaError_t as int.
// tmOut: store map of the output (box = 128 rows / pixel box x 128 bytes); tmRes: same geometry over the bf16 residual.
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const CUtensorMap& tmRes,
                const GemmParams& p, int bn, int split_k, int ctas, cudaStream_t stream);
// number of output tiles (128 x bn) of a launch and the workspace bytes cluster split-K needs for it
int gemm_out_tiles(const GemmParams& p, int bn);
inline size_t gemm_ksplit_ws_bytes(int out_tiles, int ksplit, int bn) {
  return static_cast<size_t>(out_tiles) * ksplit * (bn / 32) * 128 * 32 * sizeof(float);
}
// ctas = 2: CTA pairs (cta_group::2, 256-row MMA, the B tile split between the two CTAs); bn must be 128 or 256

}  // namespace vqa
