// Counter-based dropout mask of the training attention (documented in include/iefvad.h: iefvad_attention_train_fwd).
#pragma once
#include "common.cuh"

namespace iefvad {

// Philox4x32-10 (Salmon et al., SC'11)
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// keep (1) / drop (0) decision of attention weight (bh, query i, key j): one Philox call serves 4 consecutive keys
__device__ __forceinline__ float keep_scale(unsigned long long seed, int bh, int i, int j, uint32_t thresh, float inv_keep) {
  const uint4 r = philox4x32_10(make_uint4(uint32_t(j >> 2), uint32_t(i), uint32_t(bh), 0u),
                                make_uint2(uint32_t(seed), uint32_t(seed >> 32)));
  const uint32_t v = (j & 3) == 0 ? r.x : (j & 3) == 1 ? r.y : (j & 3) == 2 ? r.z : r.w;
  return v >= thresh ? inv_keep : 0.f;           // P(drop) = thresh / 2^32
}

// the same decision for the two consecutive keys j, j + 1 (j even): both lie in one Philox word group
__device__ __forceinline__ float2 keep_scale2(unsigned long long seed, int bh, int i, int j, uint32_t thresh, float inv_keep) {
  const uint4 r = philox4x32_10(make_uint4(uint32_t(j >> 2), uint32_t(i), uint32_t(bh), 0u),
                                make_uint2(uint32_t(seed), uint32_t(seed >> 32)));
  const uint32_t v0 = (j & 2) ? r.z : r.x, v1 = (j & 2) ? r.w : r.y;
  return make_float2(v0 >= thresh ? inv_keep : 0.f, v1 >= thresh ? inv_keep : 0.f);
}

}  // namespace iefvad
