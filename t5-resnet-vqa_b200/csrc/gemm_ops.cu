#include "gemm_ops.cuh"

#include <stdlib.h>
#include <string.h>

#include "tmap.cuh"

namespace vqa {

namespace {

void fill_epilogue(GemmParams& p, const Epilogue& e) {
  p.bias = e.bias;
  p.relu = e.relu;
  p.relu_mask = e.relu_mask;
  p.ldm = e.ldm;
  p.drop_p = e.drop_p;
  p.drop_sid = e.drop_sid;
  p.rng = e.rng;
  p.residual = e.residual;
  p.ldr = e.ldr;
  p.res_fp32 = e.res_fp32;
  p.res_first = e.res_first;
  p.alpha = e.alpha;
  p.ksplit = e.ksplit > 1 ? e.ksplit : 1;
  p.ks_ws = e.ks_ws;
  p.max_ctas = e.max_ctas;
}

// cluster split-K preconditions (the kernel trusts them)
int check_ksplit(const GemmOp* op, const Epilogue& e, int bn, int split_k, int ctas) {
  const GemmParams& p = op->p;
  if (p.ksplit <= 1) return 0;
  if (ctas != 1 || split_k > 1 || p.atomic_out) { set_last_error("gemm: cluster split-K excludes CTA pairs, split_k and accumulate"); return -1; }
  if (p.ksplit > 8 || p.ksplit > bn / 64 * 4) { set_last_error("gemm: ksplit must be <= 8"); return -1; }
  if (p.kb_total < 2 * p.ksplit) { set_last_error("gemm: cluster split-K needs at least 2 k-blocks per slice"); return -1; }
  const int tiles = gemm_out_tiles(p, bn);
  if (tiles * p.ksplit > 148) { set_last_error("gemm: cluster split-K needs tiles x ksplit <= 148 (%d x %d)", tiles, p.ksplit); return -1; }
  if (e.ks_ws == nullptr || e.ks_ws_bytes < gemm_ksplit_ws_bytes(tiles, p.ksplit, bn) ||
      (reinterpret_cast<uintptr_t>(e.ks_ws) & 15)) {
    set_last_error("gemm: cluster split-K workspace missing or too small (%zu bytes needed)", gemm_ksplit_ws_bytes(tiles, p.ksplit, bn));
    return -1;
  }
  return 0;
}

// Pixel box (w, h, images) with at most `rows` output pixels that covers [Wo, Ho, Nimg] with the
// fewest boxes; ties -> fewer wasted rows inside the box.  exact: product must equal rows.
void choose_box(int Wo, int Ho, int Nimg, int rows, bool exact, int* bw, int* bh, int* bn) {
  auto pow2_ceil = [](int v) { int p = 1; while (p < v) p <<= 1; return p; };
  const int max_w = exact ? pow2_ceil(Wo) : Wo, max_h = exact ? pow2_ceil(Ho) : Ho;
  long long best_tiles = -1;
  int best_area = 0;
  for (int w = 1; w <= rows && w <= max_w && w <= 128; ++w) {
    for (int h = 1; w * h <= rows && h <= max_h; ++h) {
      int n = rows / (w * h);
      if (n < 1) continue;
      if (!exact && n > Nimg) n = Nimg;  // exact boxes may overhang the batch: TMA zero-fills the missing images
      if (exact && w * h * n != rows) continue;
      const long long tiles = static_cast<long long>((Wo + w - 1) / w) * ((Ho + h - 1) / h) *
                              ((Nimg + n - 1) / n);
      const int area = w * h * n;
      if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && area < best_area)) {
        best_tiles = tiles; best_area = area;
        *bw = w; *bh = h; *bn = n;
      }
    }
  }
}

// Epilogue-side tensor maps and argument checks shared by the three builders.  Linear outputs: [M, N] with row
// stride ldo; convolution outputs: NHWC through the same pixel boxes the A operand was loaded with.
int finish_output_maps(GemmOp* op, bool pixels, int Nimg, int Ho, int Wo) {
  GemmParams& p = op->p;
  const int esize = p.out_fp32 ? 4 : 2;
  const uint32_t panel_cols = p.out_fp32 ? 32u : 64u;
  if (reinterpret_cast<uintptr_t>(p.out) & 15) { set_last_error("gemm: output pointer must be 16-byte aligned"); return -1; }
  if ((p.ldo * esize) & 15) { set_last_error("gemm: output row stride must be a multiple of 16 bytes"); return -1; }
  if (p.residual != nullptr) {
    if (p.N % 32) { set_last_error("gemm: a residual needs N %% 32 == 0"); return -1; }
    if (p.res_fp32) {
      if (pixels) { set_last_error("conv: fp32 residuals are not supported (bf16 NHWC only)"); return -1; }
      if ((reinterpret_cast<uintptr_t>(p.residual) & 15) || (p.ldr & 3)) { set_last_error("gemm: fp32 residual must be 16-byte aligned with ldr %% 4 == 0"); return -1; }
    } else {
      if (p.out_fp32) { set_last_error("gemm: a bf16 residual is only supported with bf16 output"); return -1; }
      if ((reinterpret_cast<uintptr_t>(p.residual) & 15) || (p.ldr & 7)) { set_last_error("gemm: bf16 residual must be 16-byte aligned with ldr %% 8 == 0"); return -1; }
    }
  }
  if (p.relu_mask != nullptr) {
    if (pixels || (p.N % 32) || (p.ldm & 7) || (reinterpret_cast<uintptr_t>(p.relu_mask) & 15)) {
      set_last_error("gemm: a ReLU mask needs a linear output, N %% 32 == 0, ldm %% 8 == 0 and 16-byte alignment");
      return -1;
    }
  }
  if (p.drop_p > 0.f && (pixels || (p.N % 8))) { set_last_error("gemm: dropout needs a linear output with N %% 8 == 0"); return -1; }
  if (p.atomic_out && pixels) { set_last_error("conv: accumulate is not supported"); return -1; }
  int r;
  if (!pixels) {
    const uint64_t dims[2] = {static_cast<uint64_t>(p.N), static_cast<uint64_t>(p.M)};
    const uint64_t strides[1] = {static_cast<uint64_t>(p.ldo) * esize};
    const uint32_t box[2] = {panel_cols, 128};
    r = p.out_fp32 ? make_tmap_f32(&op->tmOut, p.out, 2, dims, strides, box, nullptr)
                   : make_tmap_bf16(&op->tmOut, p.out, 2, dims, strides, box, nullptr);
    if (r) return r;
    p.res_tx_bytes = 128 * 128;
    if (p.residual != nullptr && !p.res_fp32) {
      const uint64_t rstrides[1] = {static_cast<uint64_t>(p.ldr) * 2};
      r = make_tmap_bf16(&op->tmRes, p.residual, 2, dims, rstrides, box, nullptr);
      if (r) return r;
    } else {
      op->tmRes = op->tmOut;
    }
  } else {
    const uint64_t dims[4] = {static_cast<uint64_t>(p.N), static_cast<uint64_t>(Wo), static_cast<uint64_t>(Ho),
                              static_cast<uint64_t>(Nimg)};
    const uint64_t strides[3] = {static_cast<uint64_t>(p.ldo) * esize, static_cast<uint64_t>(Wo) * p.ldo * esize,
                                 static_cast<uint64_t>(Ho) * Wo * p.ldo * esize};
    const uint32_t box[4] = {panel_cols, static_cast<uint32_t>(p.bx_w), static_cast<uint32_t>(p.bx_h),
                             static_cast<uint32_t>(p.bx_n)};
    r = p.out_fp32 ? make_tmap_f32(&op->tmOut, p.out, 4, dims, strides, box, nullptr)
                   : make_tmap_bf16(&op->tmOut, p.out, 4, dims, strides, box, nullptr);
    if (r) return r;
    p.res_tx_bytes = p.bx_w * p.bx_h * p.bx_n * 128;
    if (p.residual != nullptr) {
      const uint64_t rstrides[3] = {static_cast<uint64_t>(p.ldr) * 2, static_cast<uint64_t>(Wo) * p.ldr * 2,
                                    static_cast<uint64_t>(Ho) * Wo * p.ldr * 2};
      r = make_tmap_bf16(&op->tmRes, p.residual, 4, dims, rstrides, box, nullptr);
      if (r) return r;
    } else {
      op->tmRes = op->tmOut;
    }
  }
  return 0;
}

}  // namespace

int gemm_op_init(GemmOp* op, int M, int N, int K, const void* A, long long lda, int a_mn,
                 const void* B, long long ldb, int b_mn, void* out, long long ldo, int out_fp32,
                 const Epilogue& epi, int bn, int split_k, int ctas) {
  memset(&op->p, 0, sizeof(op->p));
  op->valid = false;
  if (bn != 64 && bn != 128 && bn != 256) { set_last_error("gemm: bn must be 64/128/256"); return -1; }
  if (ctas != 1 && !(ctas == 2 && bn >= 128)) { set_last_error("gemm: CTA pairs need bn 128 or 256"); return -1; }
  if ((split_k > 1 || epi.accumulate) && !out_fp32) { set_last_error("gemm: split_k / accumulate need fp32 output"); return -1; }
  if ((lda & 7) || (ldb & 7)) { set_last_error("gemm: lda/ldb must be multiples of 8 elements"); return -1; }
  GemmParams& p = op->p;
  p.M = M; p.N = N;
  p.kb_total = (K + 63) / 64;
  p.a_mn = a_mn; p.b_mn = b_mn;
  p.a_mode = LOAD_2D; p.b_mode = LOAD_2D;
  p.stage_tx_bytes = 128 * 64 * 2 + (bn / ctas) * 64 * 2;
  p.out = out; p.ldo = ldo; p.out_fp32 = out_fp32; p.out_pixels = 0;
  p.atomic_out = (split_k > 1 || epi.accumulate) ? 1 : 0;
  fill_epilogue(p, epi);
  const bool split = epi.b_lo != nullptr;
  if (split) {
    if (a_mn || b_mn || split_k > 1 || epi.accumulate || epi.ksplit > 1 || (K % 64) || epi.a_lo_col < 0 ||
        (epi.a_lo_col % 8) || (epi.residual != nullptr && !epi.res_fp32) ||
        (reinterpret_cast<uintptr_t>(epi.b_lo) & 15)) {
      set_last_error("gemm: the two-term operand split needs K-major operands, K %% 64 == 0, no split_k / accumulate / "
                     "ksplit and no bf16 residual");
      return -1;
    }
    if (epi.a_lo_col > 0 && (epi.a_lo_col < K || lda < epi.a_lo_col + K)) {
      set_last_error("gemm: a_lo_col must be >= K with lda >= a_lo_col + K");
      return -1;
    }
    p.nseg = epi.a_lo_col > 0 ? 3 : 2;
    p.kseg = K / 64;
    p.a_lo_col = static_cast<int>(epi.a_lo_col);
    p.kb_total = p.nseg * p.kseg;
  }
  int r;
  if (!a_mn) r = make_tmap_2d(&op->tmA, A, M, split && epi.a_lo_col > 0 ? epi.a_lo_col + K : K, lda, 64, 128);
  else       r = make_tmap_2d(&op->tmA, A, K, M, lda, 64, 64);
  if (r) return r;
  if (!b_mn) r = make_tmap_2d(&op->tmB, B, N, K, ldb, 64, bn / ctas);
  else       r = make_tmap_2d(&op->tmB, B, K, N, ldb, 64, 64);
  if (r) return r;
  r = finish_output_maps(op, false, 0, 0, 0);
  if (r) return r;
  if (split) {   // the low-order half of B travels in the (otherwise unused) residual tensor map
    r = make_tmap_2d(&op->tmRes, epi.b_lo, N, K, ldb, 64, bn / ctas);
    if (r) return r;
  }
  r = check_ksplit(op, epi, bn, split_k, ctas);
  if (r) return r;
  op->bn = bn; op->split_k = split_k < 1 ? 1 : split_k; op->ctas = ctas;
  op->valid = true;
  return 0;
}

int conv_op_init(GemmOp* op, const ConvGeom& g, const void* x, const void* w, void* out, int out_fp32,
                 const Epilogue& epi, int bn, int ctas) {
  memset(&op->p, 0, sizeof(op->p));
  op->valid = false;
  if (bn != 64 && bn != 128 && bn != 256) { set_last_error("conv: bn must be 64/128/256"); return -1; }
  if (ctas != 1 && !(ctas == 2 && bn >= 128)) { set_last_error("conv: CTA pairs need bn 128 or 256"); return -1; }
  GemmParams& p = op->p;
  int bw = 1, bh = 1, bb = 1;
  choose_box(g.Wo, g.Ho, g.Nimg, 128, false, &bw, &bh, &bb);
  p.bx_w = bw; p.bx_h = bh; p.bx_n = bb;
  p.tiles_w = (g.Wo + bw - 1) / bw;
  p.tiles_h = (g.Ho + bh - 1) / bh;
  p.Wo = g.Wo; p.Ho = g.Ho; p.Nimg = g.Nimg;
  p.M = g.Nimg * g.Ho * g.Wo;
  p.N = g.Cout;
  p.a_mn = 0; p.b_mn = 0;
  p.a_mode = LOAD_CONV; p.b_mode = LOAD_2D;
  p.stage_tx_bytes = bw * bh * bb * 128 + (bn / ctas) * 64 * 2;
  p.out = out; p.ldo = g.Cout; p.out_fp32 = out_fp32; p.out_pixels = 1; p.atomic_out = 0;
  fill_epilogue(p, epi);
  int r;
  long long Ktot;
  if (g.stem7) {
    // input [N, H, Wp = W + 8, 8] bf16; window of 8 pixels x 8 channels (128 B) per (row tap, wo)
    const int Wp = g.W + 8;
    const uint64_t dims[4] = {64, static_cast<uint64_t>(g.Wo), static_cast<uint64_t>(g.H),
                              static_cast<uint64_t>(g.Nimg)};
    const uint64_t strides[3] = {16 * 2, static_cast<uint64_t>(Wp) * 8 * 2,
                                 static_cast<uint64_t>(g.H) * Wp * 8 * 2};
    const uint32_t box[4] = {64, static_cast<uint32_t>(bw), static_cast<uint32_t>(bh * 2),
                             static_cast<uint32_t>(bb)};
    const uint32_t es[4] = {1, 1, 2, 1};
    r = make_tmap_bf16(&op->tmA, x, 4, dims, strides, box, es);
    if (r) return r;
    p.taps_s = 1; p.cchunks = 1; p.kb_total = 7;
    p.stride_w = 1; p.pad_w = 0; p.dil_w = 0;
    p.stride_h = 2; p.pad_h = 3;
    Ktot = 7 * 64;
  } else {
    if (g.Cin % 64) { set_last_error("conv: Cin must be a multiple of 64"); return -1; }
    const uint64_t dims[4] = {static_cast<uint64_t>(g.Cin), static_cast<uint64_t>(g.W),
                              static_cast<uint64_t>(g.H), static_cast<uint64_t>(g.Nimg)};
    const uint64_t strides[3] = {static_cast<uint64_t>(g.Cin) * 2, static_cast<uint64_t>(g.W) * g.Cin * 2,
                                 static_cast<uint64_t>(g.H) * g.W * g.Cin * 2};
    const uint32_t s = static_cast<uint32_t>(g.stride);
    const uint32_t box[4] = {64, static_cast<uint32_t>(bw) * s, static_cast<uint32_t>(bh) * s,
                             static_cast<uint32_t>(bb)};
    const uint32_t es[4] = {1, s, s, 1};
    r = make_tmap_bf16(&op->tmA, x, 4, dims, strides, box, es);
    if (r) return r;
    p.taps_s = g.S; p.cchunks = g.Cin / 64; p.kb_total = g.R * g.S * p.cchunks;
    p.stride_w = g.stride; p.stride_h = g.stride; p.pad_w = g.pad; p.pad_h = g.pad; p.dil_w = 1;
    Ktot = static_cast<long long>(g.R) * g.S * g.Cin;
  }
  r = make_tmap_2d(&op->tmB, w, g.Cout, Ktot, Ktot, 64, bn / ctas);
  if (r) return r;
  r = finish_output_maps(op, true, g.Nimg, g.Ho, g.Wo);
  if (r) return r;
  r = check_ksplit(op, epi, bn, 1, ctas);
  if (r) return r;
  op->bn = bn; op->split_k = 1; op->ctas = ctas;
  op->valid = true;
  return 0;
}

int conv_wgrad_op_init(GemmOp* op, const ConvGeom& g, const void* dy, const void* x, float* dw, int bn,
                       int split_k, int ctas) {
  memset(&op->p, 0, sizeof(op->p));
  op->valid = false;
  if (bn != 64 && bn != 128 && bn != 256) { set_last_error("wgrad: bn must be 64/128/256"); return -1; }
  if (ctas != 1 && !(ctas == 2 && bn >= 128)) { set_last_error("wgrad: CTA pairs need bn 128 or 256"); return -1; }
  if (g.stride != 1 || g.Ho != g.H || g.Wo != g.W) { set_last_error("wgrad: stride-1 same conv only"); return -1; }
  if (g.Cin % bn) { set_last_error("wgrad: Cin must be a multiple of bn"); return -1; }
  if (g.Cout % 8 || g.Cin % 8) { set_last_error("wgrad: channels must be multiples of 8"); return -1; }
  GemmParams& p = op->p;
  int bw = 1, bh = 1, bb = 1;
  choose_box(g.Wo, g.Ho, g.Nimg, 64, true, &bw, &bh, &bb);
  if (bw * bh * bb != 64) { set_last_error("wgrad: no 64-pixel box"); return -1; }
  p.bx_w = bw; p.bx_h = bh; p.bx_n = bb;
  p.tiles_w = (g.Wo + bw - 1) / bw;
  p.tiles_h = (g.Ho + bh - 1) / bh;
  p.Wo = g.Wo; p.Ho = g.Ho; p.Nimg = g.Nimg;
  p.M = g.Cout;
  p.N = g.R * g.S * g.Cin;
  p.kb_total = p.tiles_w * p.tiles_h * ((g.Nimg + bb - 1) / bb);
  p.a_mn = 1; p.b_mn = 1;
  p.a_mode = LOAD_PIXELS_MN; p.b_mode = LOAD_PIXELS_MN;
  p.stage_tx_bytes = 2 * 64 * 128 + (bn / ctas / 64) * 64 * 128;
  p.b_tap_cin = g.Cin; p.b_taps_s = g.S; p.pad_w = g.pad; p.pad_h = g.pad;
  p.out = dw; p.ldo = p.N; p.out_fp32 = 1; p.out_pixels = 0;
  p.atomic_out = split_k > 1 ? 1 : 0;
  p.alpha = 1.f;
  p.ksplit = 1;
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(g.Cout), static_cast<uint64_t>(g.Wo),
                              static_cast<uint64_t>(g.Ho), static_cast<uint64_t>(g.Nimg)};
    const uint64_t strides[3] = {static_cast<uint64_t>(g.Cout) * 2, static_cast<uint64_t>(g.Wo) * g.Cout * 2,
                                 static_cast<uint64_t>(g.Ho) * g.Wo * g.Cout * 2};
    const uint32_t box[4] = {64, static_cast<uint32_t>(bw), static_cast<uint32_t>(bh), static_cast<uint32_t>(bb)};
    int r = make_tmap_bf16(&op->tmA, dy, 4, dims, strides, box, nullptr);
    if (r) return r;
  }
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(g.Cin), static_cast<uint64_t>(g.W),
                              static_cast<uint64_t>(g.H), static_cast<uint64_t>(g.Nimg)};
    const uint64_t strides[3] = {static_cast<uint64_t>(g.Cin) * 2, static_cast<uint64_t>(g.W) * g.Cin * 2,
                                 static_cast<uint64_t>(g.H) * g.W * g.Cin * 2};
    const uint32_t box[4] = {64, static_cast<uint32_t>(bw), static_cast<uint32_t>(bh), static_cast<uint32_t>(bb)};
    int r = make_tmap_bf16(&op->tmB, x, 4, dims, strides, box, nullptr);
    if (r) return r;
  }
  {
    int r = finish_output_maps(op, false, 0, 0, 0);
    if (r) return r;
  }
  op->bn = bn; op->split_k = split_k < 1 ? 1 : split_k; op->ctas = ctas;
  op->valid = true;
  return 0;
}

static int g_dbg[4] = {0, 0, 0, 0};
static long long* g_dbg_clk = nullptr;
void gemm_debug_set_clock(long long* buf) { g_dbg_clk = buf; }
void gemm_debug_set_umma(int a_lbo, int a_sbo, int b_lbo, int b_sbo) {
  g_dbg[0] = a_lbo; g_dbg[1] = a_sbo; g_dbg[2] = b_lbo; g_dbg[3] = b_sbo;
}

int gemm_op_run(const GemmOp* op, cudaStream_t stream) {
  if (!op->valid) { set_last_error("gemm_op_run: op not initialised"); return -1; }
  GemmParams p = op->p;
  p.fd_tiles_w = make_fastdiv(p.tiles_w); p.fd_tiles_h = make_fastdiv(p.tiles_h);
  p.fd_bx_w = make_fastdiv(p.bx_w); p.fd_bx_h = make_fastdiv(p.bx_h);
  p.dbg_clk = g_dbg_clk;
  { static int dm = -1; if (dm < 0) { const char* e = getenv("VQA_B200_GEMM_DBG"); dm = e ? atoi(e) : 0; } p.dbg_mode = dm; }
  p.dbg_a_lbo = g_dbg[0]; p.dbg_a_sbo = g_dbg[1]; p.dbg_b_lbo = g_dbg[2]; p.dbg_b_sbo = g_dbg[3];
  int r = launch_gemm(op->tmA, op->tmB, op->tmOut, op->tmRes, p, op->bn, op->split_k, op->ctas, stream);
  if (r == -3) set_last_error("gemm: cluster split-K needs one output tile per cluster");
  else if (r == -2) set_last_error("gemm: this epilogue combination (output type / residual / mask / dropout / accumulate) is not built");
  else if (r) set_last_error("gemm launch failed: %s", cudaGetErrorString(static_cast<cudaError_t>(r)));
  return r;
}

}  // namespace vqa
