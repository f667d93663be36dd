// One epilogue shared by the tcgen05 GEMM (gemm_tc.cu) and the fp32 SIMT GEMM (gemm_simt.cu), so the
// bias / activation / residual / layout logic is exercised identically by both precision plans.
#pragma once
#include "common.cuh"

namespace iefvad {

enum : int { ACT_NONE = 0, ACT_RELU = 1, ACT_QUICKGELU = 2 };
enum : int { EPI_ROWMAJOR = 0, EPI_QKV = 1, EPI_DISCARD = 2 /* micro-benchmarks: mainloop only */ };

struct EpiParams {
  int mode = EPI_ROWMAJOR;
  const float* bias = nullptr;   // [N] or null
  int act = ACT_NONE;
  // out = (resid ? resid[row, col] : 0) + alpha * act(acc + bias)
  const float* resid = nullptr;  // fp32 [M, ld_resid] or null
  int ld_resid = 0;
  // tcgen05 path only: the residual as fp16 [M, ld_resid] (exact for fp16 inputs), optionally plus an fp16 remainder
  // (resid = hi + lo, ~22 mantissa bits); replaces `resid`
  const void* resid_h16 = nullptr;
  const void* resid_l16 = nullptr;
  float alpha = 1.f;
  // row-major outputs (any subset)
  float* out_f32 = nullptr;      // columns [0, split_col)
  float* out_f32_b = nullptr;    // columns [split_col, N) land in out_f32_b[row, col - split_col]
  int ld_f32 = 0;
  int split_col = 1 << 30;
  bf16* out_hi = nullptr;        // bf16(out)
  bf16* out_lo = nullptr;        // bf16(out - hi), optional
  int ld_bf = 0;
  int hi_fp16 = 0;               // out_hi receives fp16(out) instead of bf16(out) (tcgen05 path only; no out_lo)
  // EPI_QKV: packed in-projection scattered into the attention kernel's operand layouts
  //   q  [B, H, T, dhp]  (scaled by qscale, columns >= dh never written: kept zero by the allocator)
  //   k  [B, H, T, dhp]
  //   vt [B, H, dh, Tpad] (transposed so that P.V is a K-major x K-major UMMA)
  bf16* q = nullptr;
  bf16* k = nullptr;
  bf16* vt = nullptr;
  int T = 0, H = 0, dh = 0, dhp = 0, Tpad = 0, D = 0;
  float qscale = 1.f;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_QUICKGELU) return v / (1.f + __expf(-1.702f * v));
  return v;
}

// Store NC (multiple of 4, <= 32) consecutive columns [col0, col0+NC) of one output row.
// `v` holds the raw accumulators.  Every pointer dereferenced here is 16-byte aligned because
// col0 % 4 == 0 (fp32) / col0 % 8 == 0 (bf16 vectors) and all leading dimensions are multiples of 8.
template <int NC>
__device__ __forceinline__ void epi_store_row(const EpiParams& p, long long row, int col0, float* v) {
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    float x = v[j];
    if (p.bias) x += __ldg(p.bias + col0 + j);
    v[j] = apply_act(x, p.act);
  }
  if (p.mode == EPI_QKV) {
    const int which = col0 / p.D;                 // 0 = q, 1 = k, 2 = v  (a chunk never straddles: D % 32 == 0)
    const int c = col0 - which * p.D;
    const int h = c / p.dh;                       // dh % NC == 0 -> chunk stays inside one head
    const int d0 = c - h * p.dh;
    const long long b = row / p.T;
    const int t = static_cast<int>(row - b * p.T);
    const long long bh = b * p.H + h;
    if (which == 2) {
      bf16* dst = p.vt + (bh * p.dh + d0) * static_cast<long long>(p.Tpad) + t;
#pragma unroll
      for (int j = 0; j < NC; ++j) dst[static_cast<long long>(j) * p.Tpad] = __float2bfloat16_rn(v[j]);
    } else {
      const float s = (which == 0) ? p.qscale : 1.f;
      bf16* dst = (which == 0 ? p.q : p.k) + (bh * p.T + t) * static_cast<long long>(p.dhp) + d0;
#pragma unroll
      for (int j = 0; j < NC; j += 8) {
        uint4 u;
        u.x = pack_bf16x2(v[j + 0] * s, v[j + 1] * s);
        u.y = pack_bf16x2(v[j + 2] * s, v[j + 3] * s);
        u.z = pack_bf16x2(v[j + 4] * s, v[j + 5] * s);
        u.w = pack_bf16x2(v[j + 6] * s, v[j + 7] * s);
        *reinterpret_cast<uint4*>(dst + j) = u;
      }
    }
    return;
  }
  if (p.resid) {
    const float* r = p.resid + row * p.ld_resid + col0;
#pragma unroll
    for (int j = 0; j < NC; j += 4) {
      const float4 rv = *reinterpret_cast<const float4*>(r + j);
      v[j + 0] = rv.x + p.alpha * v[j + 0];
      v[j + 1] = rv.y + p.alpha * v[j + 1];
      v[j + 2] = rv.z + p.alpha * v[j + 2];
      v[j + 3] = rv.w + p.alpha * v[j + 3];
    }
  } else if (p.alpha != 1.f) {
#pragma unroll
    for (int j = 0; j < NC; ++j) v[j] *= p.alpha;
  }
  if (p.out_f32) {
    float* o = (col0 < p.split_col) ? p.out_f32 + row * p.ld_f32 + col0
                                    : p.out_f32_b + row * p.ld_f32 + (col0 - p.split_col);
#pragma unroll
    for (int j = 0; j < NC; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
  }
  if (p.out_hi) {
    bf16* oh = p.out_hi + row * p.ld_bf + col0;
    if (p.out_lo) {
      bf16* ol = p.out_lo + row * p.ld_bf + col0;
#pragma unroll
      for (int j = 0; j < NC; j += 8) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float a = v[j + 2 * q], b = v[j + 2 * q + 1];
          const bf16 ah = __float2bfloat16_rn(a), bh = __float2bfloat16_rn(b);
          hi[q] = pack_bf16x2(__bfloat162float(ah), __bfloat162float(bh));
          lo[q] = pack_bf16x2(a - __bfloat162float(ah), b - __bfloat162float(bh));
        }
        *reinterpret_cast<uint4*>(oh + j) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(ol + j) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < NC; j += 8) {
        uint4 u;
        u.x = pack_bf16x2(v[j + 0], v[j + 1]);
        u.y = pack_bf16x2(v[j + 2], v[j + 3]);
        u.z = pack_bf16x2(v[j + 4], v[j + 5]);
        u.w = pack_bf16x2(v[j + 6], v[j + 7]);
        *reinterpret_cast<uint4*>(oh + j) = u;
      }
    }
  }
}

}  // namespace iefvad
