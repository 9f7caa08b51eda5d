// Shared device helpers for the bandwidth-bound kernels (vectorised 128-bit access, warp-shuffle
// reductions, bf16 packing) and the launch-plan recorder used by every C-ABI op.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <vector>

#include "rng.cuh"
#include "tmap.cuh"

namespace vqa {

// ---------------------------------------------------------------------------------------------
// Launch plan: an ordered list of pre-built kernel launches.  Every op of the C ABI either runs
// immediately (plan == nullptr) or is appended to a plan; vqa_plan_run replays the list on a stream
// (optionally through an instantiated CUDA graph), so a whole forward or backward pass costs one
// host call.  Launch descriptors (tensor maps, pointers, shapes) are resolved once at record time.
// ---------------------------------------------------------------------------------------------
struct OpNote {
  const char* name = "";   // static string: kernel family
  double flops = 0.0;      // algorithmic floating-point operations of the launch (0 for bandwidth kernels)
  double bytes = 0.0;      // algorithmic HBM bytes of the launch (0 when not stated)
};
// Two lanes: lane 0 replays on the caller's stream, lane 1 on a plan-owned side stream, ordered only by
// explicit fork (lane 1 waits for lane 0's position) and join (lane 0 waits for lane 1) markers.  Independent
// chains - the frozen backbone next to the T5 encoder, weight gradients next to the data-gradient chain - then
// overlap on the GPU; in a captured graph they become parallel branches.
enum : int { PLAN_LAUNCH = 0, PLAN_FORK = 1, PLAN_JOIN = 2, PLAN_MARK = 3, PLAN_WAIT = 4 };
struct Plan {
  std::vector<std::function<int(cudaStream_t)>> ops;   // launches only (notes[i] describes ops[i])
  std::vector<OpNote> notes;
  std::vector<int> lanes;                               // lane of ops[i]
  struct Step { int kind; int op; };                    // replay order: launches interleaved with fork / join
  std::vector<Step> steps;
  int cur_lane = 0;
  bool lane1_open = false;                              // lane 1 has work not yet joined
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::vector<cudaEvent_t> marks;                       // lane-1 positions lane 0 may wait for (PLAN_MARK / PLAN_WAIT)
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
};

// Annotation of the NEXT submitted launch (set by each C entry point just before submit()).
inline OpNote& pending_note() {
  static thread_local OpNote n;
  return n;
}
inline void note_op(const char* name, double flops, double bytes) {
  OpNote& n = pending_note();
  n.name = name; n.flops = flops; n.bytes = bytes;
}

template <class F>
inline int submit(void* plan, void* stream, F&& fn) {
  if (plan != nullptr) {
    Plan* p = static_cast<Plan*>(plan);
    p->ops.emplace_back(std::forward<F>(fn));
    p->notes.push_back(pending_note());
    p->lanes.push_back(p->cur_lane);
    p->steps.push_back({PLAN_LAUNCH, static_cast<int>(p->ops.size()) - 1});
    if (p->cur_lane == 1) p->lane1_open = true;
    pending_note() = OpNote();
    return 0;
  }
  pending_note() = OpNote();
  return fn(static_cast<cudaStream_t>(stream));
}

inline int launch_status(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch.  Every kernel of the library is launched with the programmatic-stream-
// serialization attribute and starts with griddepcontrol.launch_dependents (+ its smem / barrier / TMEM
// prologue, where it has one) followed by griddepcontrol.wait before its first global-memory access: the
// next kernel's launch latency and prologue overlap the tail of the previous one, while `wait` (which
// returns only when the preceding grid has completed and flushed) keeps the stream's data dependencies.
// The edges survive CUDA-graph capture as programmatic dependencies.  VQA_B200_PDL=0 disables it.
// ---------------------------------------------------------------------------------------------
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_grid_sync() {
  pdl_launch_dependents();
  pdl_wait();
}

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ uint4 pack_bf16x8(const float (&f)[8]) {
  uint4 u;
  u.x = pack_bf16x2(f[0], f[1]);
  u.y = pack_bf16x2(f[2], f[3]);
  u.z = pack_bf16x2(f[4], f[5]);
  u.w = pack_bf16x2(f[6], f[7]);
  return u;
}
__device__ __forceinline__ void load_f32x8(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store_f32x8(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ void load_bf16x8(const __nv_bfloat16* p, float (&f)[8]) {
  unpack_bf16x8(*reinterpret_cast<const uint4*>(p), f);
}
__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* p, const float (&f)[8]) {
  *reinterpret_cast<uint4*>(p) = pack_bf16x8(f);
}

// Dropout on 8 consecutive elements whose first flat index is idx (idx % 8 == 0).
struct DropCtx {
  unsigned long long seed, offset;
  uint32_t thresh, sid;
  float scale;
  bool on;
};
__device__ __forceinline__ DropCtx drop_ctx(float p, uint32_t sid, const unsigned long long* rng) {
  DropCtx c;
  c.on = p > 0.f;
  c.sid = sid;
  c.thresh = 0; c.scale = 1.f; c.seed = 0; c.offset = 0;
  if (c.on) {
    c.seed = rng[0]; c.offset = rng[1];
    c.thresh = drop_threshold(p);
    c.scale = 1.f / (1.f - p);
  }
  return c;
}
__device__ __forceinline__ void drop8(const DropCtx& c, unsigned long long idx, float (&v)[8]) {
  if (!c.on) return;
  const Philox8 r = philox8(c.seed, c.offset, c.sid, idx >> 3);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (r.u16(i) < c.thresh) ? 0.f : v[i] * c.scale;
}

}  // namespace vqa
