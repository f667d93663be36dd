// mu / logvar heads of both modalities + uncertainty-weighted fusion in one tcgen05 kernel (model/imf_vad.py:125-144),
// for callers that keep only `fused` (the evaluation forward): heads_fuse.cu.
#pragma once
#include "common.cuh"

namespace iefvad {

constexpr int kHeadsFuseDim = 768;

struct HeadsFuseArgs {
  const void* x_i = nullptr;       // fp16 [M, 768]: whitened encoder output of the image modality
  const void* x_e = nullptr;       // ... of the event modality
  const void* w_i16 = nullptr;     // fp16 [1536, 768]: rows [0, 768) = mu head, [768, 1536) = logvar head (image)
  const void* w_e16 = nullptr;     // (event)
  const float* b_i = nullptr;      // [1536] biases, same row order
  const float* b_e = nullptr;
  float factor = 1.f, eps = 1e-8f; // noise-model factor ((nu + 1) / nu for StudentT) and the epsilon of the weights
  void* out_hi = nullptr;          // fp16 [M, 768]: fp16(fused)
  void* out_lo = nullptr;          // fp16 [M, 768]: fp16(fused - hi)
  long long M = 0;
};

int heads_fuse(const HeadsFuseArgs& a, int num_sms, cudaStream_t stream);

}  // namespace iefvad
