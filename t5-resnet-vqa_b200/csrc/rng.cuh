// Counter-based dropout masks: Philox4x32-7 keyed by the step's (seed, offset) pair, counter =
// (element index / 8, tensor id).  One call yields eight 16-bit uniforms, i.e. the keep/drop decision
// for eight consecutive elements, so forward and backward regenerate identical masks without storing
// them (reference: nn.Dropout(0.1) in model/multi_head_vision_text_attn.py:36,93,135-141 and the T5
// dropouts hf:734,96,149,375,768; torch's own RNG stream cannot be reproduced).
// Seven rounds: the smallest Philox4x32 variant that passes BigCrush (Salmon et al., SC'11); the mask generation runs
// inside GEMM epilogues (~12 instructions per element at ten rounds), where the three extra safety rounds are not free.
#pragma once
#include <stdint.h>

namespace vqa {

struct Philox8 {
  uint32_t w[4];
  __device__ __forceinline__ uint32_t u16(int j) const { return (w[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu; }
};

__device__ __forceinline__ Philox8 philox8(unsigned long long seed, unsigned long long offset,
                                           uint32_t sid, unsigned long long group) {
  uint32_t c0 = static_cast<uint32_t>(group), c1 = static_cast<uint32_t>(group >> 32);
  uint32_t c2 = sid, c3 = static_cast<uint32_t>(offset);
  uint32_t k0 = static_cast<uint32_t>(seed) ^ static_cast<uint32_t>(offset >> 32);
  uint32_t k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
  for (int r = 0; r < 7; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  Philox8 o; o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
  return o;
}

__device__ __forceinline__ uint32_t drop_threshold(float p) {
  return static_cast<uint32_t>(p * 65536.0f + 0.5f);
}

}  // namespace vqa
