// Internal C++ interface of the two GEMM engines (tcgen05 bf16 and fp32 SIMT).
#pragma once
#include "common.cuh"
#include "epilogue.cuh"

namespace iefvad {

struct GemmTcArgs {
  const bf16* A_hi = nullptr;  // [M, lda]  activations (K-major)
  const bf16* A_lo = nullptr;  // residual of the bf16 rounding, only for nsplit == 3
  const bf16* W_hi = nullptr;  // [N, ldw]  nn.Linear weight layout (K-major)
  const bf16* W_lo = nullptr;
  int M = 0, N = 0, K = 0;
  int lda = 0, ldw = 0;
  int nsplit = 1;              // 1: plain bf16;  3: hi.hi + hi.lo + lo.hi
  int fp16 = 0;                // A_hi / W_hi hold fp16 (E5M10) instead of bf16: 8x finer operand rounding, same MMA rate
  int force_bn = 0;            // 0 = heuristic, else 64 / 128 / 256 (tests, tuning)
  int force_stages = 0;        // 0 = as many smem stages as fit beside the epilogue buffers (tuning)
  int ksplit = 1;              // split-K (small M x N, long K: the wgrad GEMMs): slice s accumulates into rows [s M, (s + 1) M) of
                               // out_f32, which must hold ksplit x M rows; the caller sums the slices (fixed order)
  int force_cg = 0;            // 0 = heuristic, 1 = one CTA per tile, 2 = CTA pair (cta_group::2, 256-row tiles)
};

// C = epilogue(A . W^T): bf16 operands through TMA, tcgen05.mma into TMEM, and an epilogue whose global traffic
// (fp32 residual in; fp32 / bf16 hi / bf16 lo / q / k out) moves through TMA bulk tensor copies as well.
int gemm_tc(const GemmTcArgs& g, const EpiParams& ep, int num_sms, cudaStream_t stream);

// fp32 FFMA GEMM with the per-thread epilogue of epilogue.cuh: the "fp32 plan" (1e-5 class) and the on-device
// yardstick the tcgen05 path is debugged against.  A [M, lda] fp32, W [N, ldw] fp32.
int gemm_simt(const float* A, int lda, const float* W, int ldw, int M, int N, int K, const EpiParams& ep,
              cudaStream_t stream);

}  // namespace iefvad
