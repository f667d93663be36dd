// Out-projection + residual + LayerNorm (+ whitening LayerNorm) of one encoder layer as ONE tcgen05 kernel
// (model/imf_vad.py:115-117 / :121-123: `x = LN_i(x + out_proj(ctx))`, after the last layer `LN_whiten(.)`), outproj_ln.cu.
#pragma once
#include "common.cuh"

namespace iefvad {

constexpr int kOutprojLnDim = 768;

struct OutprojLnArgs {
  const void* ctx = nullptr;       // fp16 [M, 768]: attention context (the A operand)
  const void* w16 = nullptr;       // fp16 [768, 768]: out_proj.weight ([out, in])
  const float* bias = nullptr;     // [768] out_proj.bias
  const void* res_hi = nullptr;    // fp16 [M, 768]: the layer's input x (residual) ...
  const void* res_lo = nullptr;    // ... optionally with its fp16 remainder (x = hi + lo, ~22 bits)
  const float* ln_w = nullptr;     // LayerNorm weight / bias
  const float* ln_b = nullptr;
  const float* ln2_w = nullptr;    // second (whitening) LayerNorm applied to the first one's output, or null
  const float* ln2_b = nullptr;
  float eps = 1e-5f;
  void* out_hi = nullptr;          // fp16 [M or out_rows, 768]: fp16(result) - the next GEMM's operand
  void* out_lo = nullptr;          // optional fp16 remainder (the next layer's residual = hi + lo)
  const int* row_map = nullptr;    // optional [M]: result row r is written to row row_map[r] of out_hi (< 0: dropped);
                                   // out_lo must be null then
  long long M = 0;
  void* scratch = nullptr;         // >= outproj_ln_scratch_bytes(M) (partial row statistics + arrival counters)
  const void* identity = nullptr;  // fp16 [64, 64] identity (outproj_ln_identity): the W operand that adds the residual
};

size_t outproj_ln_scratch_bytes(long long M);
size_t outproj_ln_identity_bytes();
int outproj_ln_identity(void* ident, cudaStream_t stream);   // fills a buffer of outproj_ln_identity_bytes()
int outproj_ln(const OutprojLnArgs& a, int num_sms, cudaStream_t stream);

}  // namespace iefvad
