// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = epilogue( A[M,K] . W[N,K]^T ), bf16 operands, fp32 accumulate.
//
// Covers every dense layer of the IEF-VAD forward (model/imf_vad.py:115,121 in/out projections,
// :125-128 heads, :148 refinement MLPs) - K-major A (activations) and K-major W (nn.Linear stores
// [out, in]) are exactly the layouts UMMA wants, so no transposes anywhere.
//
// Structure (one CTA per SM, persistent over output tiles, 192 threads):
//   warp 0     TMA producer   : cp.async.bulk.tensor 128x64 (A) and BNx64 (W) bf16 boxes, 128B swizzle,
//                               into a STAGES-deep smem ring guarded by full/empty mbarriers
//   warp 1     MMA issuer     : one thread issues 4 x tcgen05.mma (128 x BN x 16) per k-block into one of two
//                               TMEM accumulator buffers; tcgen05.commit releases smem slots / publishes tiles
//   warps 2-5  epilogue       : tcgen05.ld 32x32b.x32 (thread == output row), fused bias / activation /
//                               residual / bf16 hi+lo split / QKV scatter (epilogue.cuh), direct global stores
// The two TMEM buffers let the epilogue of tile i overlap the MMAs of tile i+1.
//
// "split-bf16" (nsplit == 3): A ~= A_hi + A_lo, W ~= W_hi + W_lo and the product is accumulated as
// A_hi.W_hi + A_hi.W_lo + A_lo.W_hi in the same fp32 accumulator - implemented as a 3x longer K loop whose
// k-blocks pick the (A, W) tensor-map pair, so the pipeline is unchanged.
#include "common.cuh"
#include "epilogue.cuh"
#include "gemm.cuh"
#include "tensormap.cuh"

namespace iefvad {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;

template <int BN>
struct TcCfg {
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128) ? 6 : 8;
  static constexpr uint32_t kABytes = BM * BK * 2;
  static constexpr uint32_t kBBytes = BN * BK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128
                                        : (2 * BN <= 256) ? 256 : 512;
  static constexpr size_t kStageOff = size_t(kStages) * kStageBytes;          // epilogue transpose tiles (4 warps)
  static constexpr size_t kBarOff = kStageOff + 4 * EPI_STAGE_WORDS * sizeof(float);
  static constexpr size_t kSmemBytes = 1024 /*align slack*/ + kBarOff + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
               int M, int N, int K, int nsplit, EpiParams ep) {
  using Cfg = TcCfg<BN>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (M + BM - 1) / BM;
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int kb_per_seg = K / BK;
  const int total_kb = kb_per_seg * nsplit;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (nsplit > 1) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmB1);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        for (int kb = 0; kb < total_kb; ++kb) {
          const int seg = kb / kb_per_seg, kk = kb - seg * kb_per_seg;
          const CUtensorMap* ma = (seg == 2) ? &tmA1 : &tmA0;   // hi.hi, hi.lo, lo.hi
          const CUtensorMap* mb = (seg == 1) ? &tmB1 : &tmB0;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = smem + size_t(s) * Cfg::kStageBytes;
          mbar_arrive_expect_tx(&full[s], Cfg::kStageBytes);
          tma_load_2d(ma, &full[s], sa, kk * BK, m_blk * BM);
          tma_load_2d(mb, &full[s], sa + Cfg::kABytes, kk * BK, n_blk * BN);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        mbar_wait(&tempty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(as * BN);
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + size_t(s) * Cfg::kStageBytes);
          const uint64_t da = make_smem_desc_sw128(sa);
          const uint64_t db = make_smem_desc_sw128(sa + Cfg::kABytes);
#pragma unroll
          for (int k4 = 0; k4 < BK / 16; ++k4)
            umma_bf16(d_tmem, da + uint64_t(2 * k4), db + uint64_t(2 * k4), idesc, (kb | k4) != 0 ? 1u : 0u);
          tc_commit(&empty[s]);           // smem slot reusable once these MMAs retire
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        tc_commit(&tfull[as]);            // accumulator complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;         // TMEM lanes [32*quarter, 32*quarter+32) belong to this warp
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      const long long row0 = (long long)m_blk * BM + quarter * 32;
      const uint32_t t0 = tmem_base + uint32_t(as * BN) + (uint32_t(quarter * 32) << 16);
      float* stage = reinterpret_cast<float*>(smem + Cfg::kStageOff) + (warp - 2) * EPI_STAGE_WORDS;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        float v[32];
        tmem_ld32(t0 + uint32_t(c * 32), v);
        tmem_ld_wait();
        const int col0 = n_blk * BN + c * 32;
        if (row0 < M && col0 < N) epi_chunk_warp(ep, row0, col0, v, stage, lane, M);
      }
      tc_fence_before();
      mbar_arrive(&tempty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN>
int launch_bn(const GemmTcArgs& g, const EpiParams& ep, int num_sms, cudaStream_t stream) {
  using Cfg = TcCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    IEF_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(Cfg::kSmemBytes)));
    attr_set = true;
  }
  CUtensorMap a0, a1, b0, b1;
  IEF_TRY(make_tmap_2d(&a0, g.A_hi, g.K, g.M, uint64_t(g.lda) * 2, BK, BM));
  IEF_TRY(make_tmap_2d(&b0, g.W_hi, g.K, g.N, uint64_t(g.ldw) * 2, BK, BN));
  if (g.nsplit == 3) {
    IEF_TRY(make_tmap_2d(&a1, g.A_lo, g.K, g.M, uint64_t(g.lda) * 2, BK, BM));
    IEF_TRY(make_tmap_2d(&b1, g.W_lo, g.K, g.N, uint64_t(g.ldw) * 2, BK, BN));
  } else {
    a1 = a0;
    b1 = b0;
  }
  const int num_m = (g.M + BM - 1) / BM, num_n = (g.N + BN - 1) / BN;
  const int tiles = num_m * num_n;
  const int grid = tiles < num_sms ? tiles : num_sms;
  gemm_tc_kernel<BN><<<grid, 192, Cfg::kSmemBytes, stream>>>(a0, a1, b0, b1, g.M, g.N, g.K, g.nsplit, ep);
  count_launches(1);
  IEF_CUDA(cudaGetLastError());
  return IEFVAD_OK;
}

}  // namespace

int gemm_tc(const GemmTcArgs& g, const EpiParams& ep, int num_sms, cudaStream_t stream) {
  IEF_CHECK(g.M > 0 && g.N > 0 && g.K > 0, "gemm_tc: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
  IEF_CHECK(g.K % BK == 0, "gemm_tc: K=%d must be a multiple of %d", g.K, BK);
  IEF_CHECK(g.N % 32 == 0, "gemm_tc: N=%d must be a multiple of 32", g.N);
  IEF_CHECK(g.lda % 8 == 0 && g.ldw % 8 == 0, "gemm_tc: leading dimensions must be multiples of 8 elements");
  IEF_CHECK(g.nsplit == 1 || g.nsplit == 3, "gemm_tc: nsplit must be 1 or 3");
  IEF_CHECK(g.A_hi && g.W_hi && (g.nsplit == 1 || (g.A_lo && g.W_lo)), "gemm_tc: null operand");
  const int num_m = (g.M + BM - 1) / BM;
  int bn = g.force_bn;
  if (bn == 0) {
    // largest tile that still gives every SM work; small problems trade tile efficiency for parallelism
    if (g.N % 256 == 0 && num_m * (g.N / 256) >= num_sms) bn = 256;
    else if (g.N % 128 == 0 && num_m * (g.N / 128) >= num_sms) bn = 128;
    else bn = 64;
  }
  switch (bn) {
    case 256: return launch_bn<256>(g, ep, num_sms, stream);
    case 128: return launch_bn<128>(g, ep, num_sms, stream);
    case 64: return launch_bn<64>(g, ep, num_sms, stream);
    default: set_error("gemm_tc: unsupported BN=%d", bn); return IEFVAD_ERR_INVALID;
  }
}

}  // namespace iefvad
