// Attention core for both encoders of the hot path, forward and backward:
//   T5 self-attention   (hf:308-334): scores = Q K^T + position_bias + key mask, NO 1/sqrt(d), hd = 64, 12 heads
//   SGA self / guided   (model/multi_head_vision_text_attn.py:73-86): scores = Q K^T / sqrt(96), hd = 96, 8 heads
// followed by fp32 softmax, dropout on the probabilities and P V.
//
// The whole (batch, head) problem is tiny (Lq <= 32, Lk <= ~200, hd <= 96 -> < 1.3 MFLOP), so one CTA owns
// one (b, h): Q, K, V are staged once in shared memory (bf16, rows padded by one word so the per-row
// walks are bank-conflict free), scores and probabilities never leave the SM (flash-style: no [B,H,L,L]
// fp32 score round trip; the normalised fp32 probabilities are saved once for backward), and the backward kernel produces
// dQ, dK, dV and the relative-position-bias gradient in one pass.
// Since round 2 this SIMT kernel only runs for key sequences longer than 64 tokens (the guided attention over the 196 vision
// tokens of 448x448 images); everything up to 64 x 64 runs the tcgen05 kernels of attention_tc.cu.
#include "../../include/vqa_b200.h"
#include "common.cuh"

using namespace vqa;

namespace vqa {
// attention_tc.cu: tcgen05 flash kernels for Lq, Lk <= 64
bool attention_tc_supported(int Lq, int Lk, int hd);
int attention_tc_fwd(void* plan, const vqa_attn_fwd_args* x, void* stream);
int attention_tc_bwd(void* plan, const vqa_attn_bwd_args* x, void* stream);
void attention_tc_debug(long long* buf);
}  // namespace vqa

namespace {

constexpr int kThreads = 128;
constexpr int kMaxLq = 32;
constexpr float kMaskedScore = -3.4028234663852886e38f;  // torch.finfo(float32).min (hf additive mask)

__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }

__device__ __forceinline__ float drop_mult(const DropCtx& c, unsigned long long idx) {
  if (!c.on) return 1.f;
  const Philox8 r = philox8(c.seed, c.offset, c.sid, idx >> 3);
  return (r.u16(static_cast<int>(idx & 7)) < c.thresh) ? 0.f : c.scale;
}

// rows x HD bf16 tile (global row stride ld elements) -> shared rows of HD + 2 elements
template <int HD>
__device__ __forceinline__ void load_tile(uint32_t* dst, const __nv_bfloat16* src, long long ld, int rows) {
  constexpr int V = HD / 8;
  constexpr int RW = (HD + 2) / 2;  // row stride in 32-bit words
  for (int idx = threadIdx.x; idx < rows * V; idx += kThreads) {
    const int r = idx / V, c = idx - r * V;
    const uint4 g = __ldg(reinterpret_cast<const uint4*>(src + static_cast<long long>(r) * ld + c * 8));
    uint32_t* d = dst + r * RW + c * 4;
    d[0] = g.x; d[1] = g.y; d[2] = g.z; d[3] = g.w;
  }
}

template <int HD>
__device__ __forceinline__ float dot_rows(const uint32_t* a, const uint32_t* b) {
  float acc = 0.f;
#pragma unroll
  for (int d = 0; d < HD / 2; ++d) {
    const uint32_t x = a[d], y = b[d];
    acc = fmaf(bf_lo(x), bf_lo(y), acc);
    acc = fmaf(bf_hi(x), bf_hi(y), acc);
  }
  return acc;
}

struct FwdArgs {
  int B, H, Lq, Lk;
  const __nv_bfloat16 *q, *k, *v;
  long long ldq, ldk, ldv, ldo;
  __nv_bfloat16* out;
  float* probs;   // fp32: the backward's P * (dP - rowsum(P dP)) cancels a large common term, which only
                  // works if every row of P sums to 1 at fp32 accuracy
  const float* bias;
  const long long* key_mask;
  float scale, drop_p;
  uint32_t sid;
  const unsigned long long* rng;
};

template <int HD>
__global__ void __launch_bounds__(kThreads)
attention_fwd_kernel(const FwdArgs a) {
  pdl_grid_sync();
  constexpr int RW = (HD + 2) / 2;
  constexpr int DQ = HD / 4;  // head dims per thread in the P V phase
  extern __shared__ uint32_t smem[];
  const int b = blockIdx.x / a.H, h = blockIdx.x - b * a.H;
  const int lds = a.Lk | 1;
  uint32_t* Qs = smem;
  uint32_t* Ks = Qs + kMaxLq * RW;
  uint32_t* Vs = Ks + a.Lk * RW;
  float* S = reinterpret_cast<float*>(Vs + a.Lk * RW);

  load_tile<HD>(Qs, a.q + static_cast<long long>(b) * a.Lq * a.ldq + h * HD, a.ldq, a.Lq);
  load_tile<HD>(Ks, a.k + static_cast<long long>(b) * a.Lk * a.ldk + h * HD, a.ldk, a.Lk);
  load_tile<HD>(Vs, a.v + static_cast<long long>(b) * a.Lk * a.ldv + h * HD, a.ldv, a.Lk);
  __syncthreads();

  const int i = threadIdx.x >> 2, jq = threadIdx.x & 3;
  const bool row_ok = i < a.Lq;
  const DropCtx dc = drop_ctx(a.drop_p, a.sid, a.rng);

  // ---- scores + softmax (4 threads per query row) ----
  float mx = -INFINITY;
  if (row_ok) {
    for (int j = jq; j < a.Lk; j += 4) {
      float s = dot_rows<HD>(Qs + i * RW, Ks + j * RW) * a.scale;
      if (a.bias != nullptr) s += a.bias[(static_cast<long long>(h) * a.Lq + i) * a.Lk + j];
      if (a.key_mask != nullptr && a.key_mask[static_cast<long long>(b) * a.Lk + j] == 0) s = kMaskedScore;
      S[i * lds + j] = s;
      mx = fmaxf(mx, s);
    }
  }
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  float sum = 0.f;
  if (row_ok) {
    for (int j = jq; j < a.Lk; j += 4) {
      const float e = __expf(S[i * lds + j] - mx);
      S[i * lds + j] = e;
      sum += e;
    }
  }
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);
  if (row_ok) {
    const float inv = 1.f / sum;
    const long long prow = (static_cast<long long>(b) * a.H + h) * a.Lq + i;
    for (int j = jq; j < a.Lk; j += 4) {
      const float p = S[i * lds + j] * inv;
      if (a.probs != nullptr) a.probs[prow * a.Lk + j] = p;
      S[i * lds + j] = p * drop_mult(dc, static_cast<unsigned long long>(prow) * a.Lk + j);
    }
  }
  __syncthreads();

  // ---- O = P V (thread = query row i, quarter jq of the head dims) ----
  if (row_ok) {
    float acc[DQ];
#pragma unroll
    for (int t = 0; t < DQ; ++t) acc[t] = 0.f;
    const uint32_t* vbase = Vs + jq * (DQ / 2);
    for (int j = 0; j < a.Lk; ++j) {
      const float p = S[i * lds + j];
      const uint32_t* vr = vbase + j * RW;
#pragma unroll
      for (int t = 0; t < DQ / 2; ++t) {
        const uint32_t u = vr[t];
        acc[2 * t] = fmaf(p, bf_lo(u), acc[2 * t]);
        acc[2 * t + 1] = fmaf(p, bf_hi(u), acc[2 * t + 1]);
      }
    }
    __nv_bfloat16* op = a.out + (static_cast<long long>(b) * a.Lq + i) * a.ldo + h * HD + jq * DQ;
#pragma unroll
    for (int t = 0; t < DQ / 8; ++t) {
      uint4 pk;
      pk.x = pack_bf16x2(acc[8 * t + 0], acc[8 * t + 1]);
      pk.y = pack_bf16x2(acc[8 * t + 2], acc[8 * t + 3]);
      pk.z = pack_bf16x2(acc[8 * t + 4], acc[8 * t + 5]);
      pk.w = pack_bf16x2(acc[8 * t + 6], acc[8 * t + 7]);
      reinterpret_cast<uint4*>(op)[t] = pk;
    }
  }
}

struct BwdArgs {
  int B, H, Lq, Lk;
  const __nv_bfloat16 *q, *k, *v, *dout;
  const float* probs;
  long long ldq, ldk, ldv, ldo, lddq, lddk, lddv;
  __nv_bfloat16 *dq, *dk, *dv;
  float* dbias;
  float scale, drop_p;
  uint32_t sid;
  const unsigned long long* rng;
};

template <int HD>
__global__ void __launch_bounds__(kThreads)
attention_bwd_kernel(const BwdArgs a) {
  pdl_grid_sync();
  constexpr int RW = (HD + 2) / 2;
  constexpr int DQ = HD / 4;
  extern __shared__ uint32_t smem[];
  const int b = blockIdx.x / a.H, h = blockIdx.x - b * a.H;
  const int lds = a.Lk | 1;
  uint32_t* Qs = smem;
  uint32_t* dOs = Qs + kMaxLq * RW;
  uint32_t* Ks = dOs + kMaxLq * RW;
  uint32_t* Vs = Ks + a.Lk * RW;
  float* Pd = reinterpret_cast<float*>(Vs + a.Lk * RW);  // dropped probabilities  [Lq][lds]
  float* dS = Pd + kMaxLq * lds;                          // score gradients        [Lq][lds]

  load_tile<HD>(Qs, a.q + static_cast<long long>(b) * a.Lq * a.ldq + h * HD, a.ldq, a.Lq);
  load_tile<HD>(dOs, a.dout + static_cast<long long>(b) * a.Lq * a.ldo + h * HD, a.ldo, a.Lq);
  load_tile<HD>(Ks, a.k + static_cast<long long>(b) * a.Lk * a.ldk + h * HD, a.ldk, a.Lk);
  load_tile<HD>(Vs, a.v + static_cast<long long>(b) * a.Lk * a.ldv + h * HD, a.ldv, a.Lk);
  __syncthreads();

  const int i = threadIdx.x >> 2, jq = threadIdx.x & 3;
  const bool row_ok = i < a.Lq;
  const DropCtx dc = drop_ctx(a.drop_p, a.sid, a.rng);
  const long long prow = (static_cast<long long>(b) * a.H + h) * a.Lq + i;

  // ---- dP = dO V^T (through the dropout mask), dS = P * (dP - rowsum(P dP)) ----
  float rowdot = 0.f;
  if (row_ok) {
    for (int j = jq; j < a.Lk; j += 4) {
      const float p = a.probs[prow * a.Lk + j];
      const float mult = drop_mult(dc, static_cast<unsigned long long>(prow) * a.Lk + j);
      const float dp = dot_rows<HD>(dOs + i * RW, Vs + j * RW) * mult;
      Pd[i * lds + j] = p * mult;
      dS[i * lds + j] = dp;  // dP for now
      rowdot = fmaf(p, dp, rowdot);
    }
  }
  rowdot += __shfl_xor_sync(0xffffffffu, rowdot, 1);
  rowdot += __shfl_xor_sync(0xffffffffu, rowdot, 2);
  if (row_ok) {
    for (int j = jq; j < a.Lk; j += 4) {
      const float p = a.probs[prow * a.Lk + j];
      const float ds = p * (dS[i * lds + j] - rowdot);
      dS[i * lds + j] = ds;
      if (a.dbias != nullptr) atomicAdd(a.dbias + (static_cast<long long>(h) * a.Lq + i) * a.Lk + j, ds);
    }
  }
  __syncthreads();

  // ---- dQ = scale * dS K ----
  if (row_ok) {
    float acc[DQ];
#pragma unroll
    for (int t = 0; t < DQ; ++t) acc[t] = 0.f;
    const uint32_t* kbase = Ks + jq * (DQ / 2);
    for (int j = 0; j < a.Lk; ++j) {
      const float ds = dS[i * lds + j];
      const uint32_t* kr = kbase + j * RW;
#pragma unroll
      for (int t = 0; t < DQ / 2; ++t) {
        const uint32_t u = kr[t];
        acc[2 * t] = fmaf(ds, bf_lo(u), acc[2 * t]);
        acc[2 * t + 1] = fmaf(ds, bf_hi(u), acc[2 * t + 1]);
      }
    }
    __nv_bfloat16* op = a.dq + (static_cast<long long>(b) * a.Lq + i) * a.lddq + h * HD + jq * DQ;
#pragma unroll
    for (int t = 0; t < DQ / 8; ++t) {
      uint4 pk;
      pk.x = pack_bf16x2(acc[8 * t + 0] * a.scale, acc[8 * t + 1] * a.scale);
      pk.y = pack_bf16x2(acc[8 * t + 2] * a.scale, acc[8 * t + 3] * a.scale);
      pk.z = pack_bf16x2(acc[8 * t + 4] * a.scale, acc[8 * t + 5] * a.scale);
      pk.w = pack_bf16x2(acc[8 * t + 6] * a.scale, acc[8 * t + 7] * a.scale);
      reinterpret_cast<uint4*>(op)[t] = pk;
    }
  }

  // ---- dK = scale * dS^T Q,  dV = Pd^T dO  (thread = key row j, quarter jq of the head dims) ----
  for (int j = threadIdx.x >> 2; j < a.Lk; j += kThreads / 4) {
    float ak[DQ], av[DQ];
#pragma unroll
    for (int t = 0; t < DQ; ++t) { ak[t] = 0.f; av[t] = 0.f; }
    const uint32_t* qbase = Qs + jq * (DQ / 2);
    const uint32_t* obase = dOs + jq * (DQ / 2);
    for (int r = 0; r < a.Lq; ++r) {
      const float ds = dS[r * lds + j];
      const float pd = Pd[r * lds + j];
      const uint32_t* qr = qbase + r * RW;
      const uint32_t* dor = obase + r * RW;
#pragma unroll
      for (int t = 0; t < DQ / 2; ++t) {
        const uint32_t uq = qr[t], uo = dor[t];
        ak[2 * t] = fmaf(ds, bf_lo(uq), ak[2 * t]);
        ak[2 * t + 1] = fmaf(ds, bf_hi(uq), ak[2 * t + 1]);
        av[2 * t] = fmaf(pd, bf_lo(uo), av[2 * t]);
        av[2 * t + 1] = fmaf(pd, bf_hi(uo), av[2 * t + 1]);
      }
    }
    __nv_bfloat16* kp = a.dk + (static_cast<long long>(b) * a.Lk + j) * a.lddk + h * HD + jq * DQ;
    __nv_bfloat16* vp = a.dv + (static_cast<long long>(b) * a.Lk + j) * a.lddv + h * HD + jq * DQ;
#pragma unroll
    for (int t = 0; t < DQ / 8; ++t) {
      uint4 pk;
      pk.x = pack_bf16x2(ak[8 * t + 0] * a.scale, ak[8 * t + 1] * a.scale);
      pk.y = pack_bf16x2(ak[8 * t + 2] * a.scale, ak[8 * t + 3] * a.scale);
      pk.z = pack_bf16x2(ak[8 * t + 4] * a.scale, ak[8 * t + 5] * a.scale);
      pk.w = pack_bf16x2(ak[8 * t + 6] * a.scale, ak[8 * t + 7] * a.scale);
      reinterpret_cast<uint4*>(kp)[t] = pk;
      pk.x = pack_bf16x2(av[8 * t + 0], av[8 * t + 1]);
      pk.y = pack_bf16x2(av[8 * t + 2], av[8 * t + 3]);
      pk.z = pack_bf16x2(av[8 * t + 4], av[8 * t + 5]);
      pk.w = pack_bf16x2(av[8 * t + 6], av[8 * t + 7]);
      reinterpret_cast<uint4*>(vp)[t] = pk;
    }
  }
}

inline size_t fwd_smem(int hd, int Lk) {
  const int RW = (hd + 2) / 2;
  return static_cast<size_t>(kMaxLq + 2 * Lk) * RW * 4 + static_cast<size_t>(kMaxLq) * (Lk | 1) * 4;
}
inline size_t bwd_smem(int hd, int Lk) {
  const int RW = (hd + 2) / 2;
  return static_cast<size_t>(2 * kMaxLq + 2 * Lk) * RW * 4 + static_cast<size_t>(2 * kMaxLq) * (Lk | 1) * 4;
}

// tc: the tcgen05 flash kernels will take the call (they carry up to 64 query rows; the SIMT kernel kMaxLq = 32)
inline int check_common(int Lq, int Lk, int hd, long long l0, long long l1, long long l2, long long l3,
                        const char* what, bool tc) {
  if (hd != 64 && hd != 96) { set_last_error("%s: head dim must be 64 or 96 (got %d)", what, hd); return -1; }
  const int max_lq = tc ? 64 : kMaxLq;
  if (Lq < 1 || Lq > max_lq) { set_last_error("%s: Lq must be in [1, %d] (got %d)", what, max_lq, Lq); return -1; }
  if (Lk < 1 || Lk > 512) { set_last_error("%s: Lk must be in [1, 512] (got %d)", what, Lk); return -1; }
  if ((l0 | l1 | l2 | l3) & 7) { set_last_error("%s: row strides must be multiples of 8 elements", what); return -1; }
  return 0;
}

// Raise the dynamic shared-memory cap of a kernel to the sm_100 maximum once (a cap, not a reservation).
template <auto Kern>
int ensure_smem(size_t bytes, const char* what) {
  static bool done = false;
  if (bytes > 227 * 1024) { set_last_error("%s: needs %zu B of shared memory", what, bytes); return -1; }
  if (!done) {
    cudaError_t e = cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_last_error("%s: %s", what, cudaGetErrorString(e)); return static_cast<int>(e); }
    done = true;
  }
  return 0;
}

}  // namespace

extern "C" {

int vqa_debug_attn_timing(long long* buf) {
  attention_tc_debug(buf);
  return 0;
}

int vqa_attention_fwd(void* plan, const vqa_attn_fwd_args* x, void* stream) {
  const bool tc = x->stats != nullptr && attention_tc_supported(x->Lq, x->Lk, x->hd);
  if (check_common(x->Lq, x->Lk, x->hd, x->ldq, x->ldk, x->ldv, x->ldo, "attention_fwd", tc)) return -1;
  if (tc) return attention_tc_fwd(plan, x, stream);
  FwdArgs a;
  a.B = x->B; a.H = x->H; a.Lq = x->Lq; a.Lk = x->Lk;
  a.q = static_cast<const __nv_bfloat16*>(x->q); a.k = static_cast<const __nv_bfloat16*>(x->k);
  a.v = static_cast<const __nv_bfloat16*>(x->v);
  a.ldq = x->ldq; a.ldk = x->ldk; a.ldv = x->ldv; a.ldo = x->ldo;
  a.out = static_cast<__nv_bfloat16*>(x->out); a.probs = static_cast<float*>(x->probs);
  a.bias = x->bias; a.key_mask = x->key_mask; a.scale = x->scale; a.drop_p = x->drop_p; a.sid = x->sid;
  a.rng = reinterpret_cast<const unsigned long long*>(x->rng);
  const int hd = x->hd;
  const size_t smem = fwd_smem(hd, a.Lk);
  int r = hd == 64 ? ensure_smem<attention_fwd_kernel<64>>(smem, "attention_fwd")
                   : ensure_smem<attention_fwd_kernel<96>>(smem, "attention_fwd");
  if (r) return r;
  note_op("attention_fwd", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    if (hd == 64) launch_pdl(attention_fwd_kernel<64>, dim3(a.B * a.H), dim3(kThreads), smem, s, a);
    else launch_pdl(attention_fwd_kernel<96>, dim3(a.B * a.H), dim3(kThreads), smem, s, a);
    return launch_status("attention_fwd");
  });
}

int vqa_attention_bwd(void* plan, const vqa_attn_bwd_args* x, void* stream) {
  const bool tc = x->stats != nullptr && attention_tc_supported(x->Lq, x->Lk, x->hd);
  if (check_common(x->Lq, x->Lk, x->hd, x->ldq, x->ldk, x->ldv, x->ldo, "attention_bwd", tc)) return -1;
  if ((x->lddq | x->lddk | x->lddv) & 7) { set_last_error("attention_bwd: gradient strides must be multiples of 8"); return -1; }
  if (tc) return attention_tc_bwd(plan, x, stream);
  if (x->probs == nullptr) { set_last_error("attention_bwd: the SIMT kernel needs the saved probabilities"); return -1; }
  BwdArgs a;
  a.B = x->B; a.H = x->H; a.Lq = x->Lq; a.Lk = x->Lk;
  a.q = static_cast<const __nv_bfloat16*>(x->q); a.k = static_cast<const __nv_bfloat16*>(x->k);
  a.v = static_cast<const __nv_bfloat16*>(x->v); a.probs = static_cast<const float*>(x->probs);
  a.dout = static_cast<const __nv_bfloat16*>(x->dout);
  a.ldq = x->ldq; a.ldk = x->ldk; a.ldv = x->ldv; a.ldo = x->ldo;
  a.lddq = x->lddq; a.lddk = x->lddk; a.lddv = x->lddv;
  a.dq = static_cast<__nv_bfloat16*>(x->dq); a.dk = static_cast<__nv_bfloat16*>(x->dk);
  a.dv = static_cast<__nv_bfloat16*>(x->dv);
  a.dbias = x->dbias; a.scale = x->scale; a.drop_p = x->drop_p; a.sid = x->sid;
  a.rng = reinterpret_cast<const unsigned long long*>(x->rng);
  const int hd = x->hd;
  const size_t smem = bwd_smem(hd, a.Lk);
  int r = hd == 64 ? ensure_smem<attention_bwd_kernel<64>>(smem, "attention_bwd")
                   : ensure_smem<attention_bwd_kernel<96>>(smem, "attention_bwd");
  if (r) return r;
  note_op("attention_bwd", 0.0, 0.0);
  return submit(plan, stream, [=](cudaStream_t s) {
    if (hd == 64) launch_pdl(attention_bwd_kernel<64>, dim3(a.B * a.H), dim3(kThreads), smem, s, a);
    else launch_pdl(attention_bwd_kernel<96>, dim3(a.B * a.H), dim3(kThreads), smem, s, a);
    return launch_status("attention_bwd");
  });
}

}  // extern "C"
