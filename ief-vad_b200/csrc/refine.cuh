// Fused refinement chain (model/imf_vad.py:146-149): all steps x <- x - lambda (W2 relu(W1 x + b1) + b2) in ONE persistent
// launch; the hidden activation of a 256-row tile never leaves the CTA pair that computes it (refine_fused.cu).
#pragma once
#include "common.cuh"

namespace iefvad {

constexpr int kRefineMaxSteps = 16;
constexpr int kRefineDim = 768;

struct RefineChainArgs {
  void* x_hi = nullptr;            // fp16 [M, 768] row-major: fp16(x) - the fusion output; REWRITTEN in place every step
  void* x_lo = nullptr;            // fp16 [M, 768] row-major: fp16(x - hi) - the remainder of the residual stream; ditto
  const void* w16 = nullptr;       // fp16 [2 * steps, 768, 768]: W1 of step 0, W2 of step 0, W1 of step 1, ... ([out, in] each)
  const float* b1[kRefineMaxSteps] = {};
  const float* b2[kRefineMaxSteps] = {};
  int steps = 0;
  float lambda = 0.5f;
  long long M = 0;
  float* out_f32 = nullptr;        // fp32 [M, 768]: x after the last step
  void* lo_scratch = nullptr;      // >= refine_chain_scratch_bytes(M): per-tile step counters (zeroed by the call)
};

size_t refine_chain_scratch_bytes(long long M);
// true when the fused kernel is the better choice for M rows on num_sms SMs (enough 256-row tiles to fill the pairs)
bool refine_chain_preferred(long long M, int num_sms);
int refine_chain(const RefineChainArgs& a, int num_sms, cudaStream_t stream);

}  // namespace iefvad
