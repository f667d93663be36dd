// Internal C++ interface of the two GEMM engines (tcgen05 bf16 and fp32 SIMT).
#pragma once
#include "common.cuh"
#include "epilogue.cuh"

namespace iefvad {

// out = act( A.W^T [+ X_hi + X_lo] + bias ), optionally scaled on the first `scale_cols` columns, written either as
// fp32 (one destination, or two split at `split_col`) or as bf16 hi (+ lo = bf16(out - hi)), optionally with the
// per-head column remap of the packed attention in-projection.  Everything leaves the SM through TMA bulk stores.
struct GemmTcArgs {
  const bf16* A_hi = nullptr;  // [M, lda]  activations (K-major)
  const bf16* A_lo = nullptr;  // residual of the bf16 rounding, only for nsplit == 3
  const bf16* W_hi = nullptr;  // [N, ldw]  nn.Linear weight layout (K-major)
  const bf16* W_lo = nullptr;
  int M = 0, N = 0, K = 0;
  int lda = 0, ldw = 0;
  int nsplit = 1;              // 1: plain bf16;  3: hi.hi + hi.lo + lo.hi
  int force_bn = 0;            // 0 = heuristic, else 64 / 128 / 256 (tests, tuning)
  // residual added inside the accumulator by identity MMAs: out[:, n] += X_hi[:, n] (+ X_lo[:, n]); [M, ldx]
  const bf16* X_hi = nullptr;
  const bf16* X_lo = nullptr;
  int ldx = 0;
  const float* bias = nullptr; // [N] or null
  int act = ACT_NONE;
  int scale_cols = 0;          // columns [0, scale_cols) are multiplied by `scale` after bias (q pre-scaling)
  float scale = 1.f;
  // fp32 outputs: columns [0, split_col) -> out_f32, [split_col, N) -> out_f32_b[:, col - split_col]; pitch ld_f32
  float* out_f32 = nullptr;
  float* out_f32_b = nullptr;
  int split_col = 1 << 30;
  int ld_f32 = 0;
  // bf16 outputs (used when out_f32 == nullptr); pitch ld_bf elements
  bf16* out_hi = nullptr;
  bf16* out_lo = nullptr;
  int ld_bf = 0;
  // column remap of the bf16 outputs: dest = (col / remap_dh) * remap_dhp + col % remap_dh  (0 = identity)
  int remap_dh = 0, remap_dhp = 0;
  bool discard = false;        // micro-benchmarks: run the mainloop, drop the result
};

int gemm_tc(const GemmTcArgs& g, int num_sms, cudaStream_t stream);

// fp32 FFMA GEMM with the per-thread epilogue of epilogue.cuh: the "fp32 plan" (1e-5 class) and the on-device
// yardstick the tcgen05 path is debugged against.  A [M, lda] fp32, W [N, ldw] fp32.
int gemm_simt(const float* A, int lda, const float* W, int ldw, int M, int N, int K, const EpiParams& ep,
              cudaStream_t stream);

}  // namespace iefvad
