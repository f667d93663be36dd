// Device-side input shaping (row N2): process_split / process_feat + uniform_extract of data/tools.py:65-114 for a
// packed list of videos.  Pure HBM-bound gather / segmented-mean kernels: 128-bit accesses along the feature
// dimension, one CTA per output chunk (split) or per output row (feat).
#include "shaping.cuh"

#include "elementwise.cuh"

namespace iefvad {

namespace {

// torch.nan_to_num defaults: nan -> 0, +inf -> max finite, -inf -> lowest finite (of the tensor's dtype)
__device__ __forceinline__ float fix_f32(float v) {
  if (v != v) return 0.f;
  if (isinf(v)) return v > 0 ? 3.402823466e+38f : -3.402823466e+38f;
  return v;
}
__device__ __forceinline__ __half fix_f16(__half h) {
  const float v = __half2float(h);
  if (v != v) return __float2half_rn(0.f);
  if (isinf(v)) return __float2half_rn(v > 0 ? 65504.f : -65504.f);
  return h;
}
__device__ __forceinline__ bf16 fix_bf16(bf16 h) {
  const float v = __bfloat162float(h);
  if (v != v) return __float2bfloat16_rn(0.f);
  if (isinf(v)) return __float2bfloat16_rn(v > 0 ? 3.3895313892515355e+38f : -3.3895313892515355e+38f);
  return h;
}
template <typename T> __device__ __forceinline__ T fix(T v);
template <> __device__ __forceinline__ float fix<float>(float v) { return fix_f32(v); }
template <> __device__ __forceinline__ __half fix<__half>(__half v) { return fix_f16(v); }
template <> __device__ __forceinline__ bf16 fix<bf16>(bf16 v) { return fix_bf16(v); }

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// largest v with off[v] <= c  (off ascending, off[0] = 0, off[V] > c)
__device__ __forceinline__ long long find_segment(const long long* off, long long V, long long c) {
  long long lo = 0, hi = V;
  while (hi - lo > 1) {
    const long long mid = (lo + hi) >> 1;
    if (off[mid] <= c) lo = mid; else hi = mid;
  }
  return lo;
}

template <typename T>
__global__ void __launch_bounds__(256)
process_split_kernel(const T* __restrict__ src, const long long* __restrict__ row_off, long long V, int D, int length,
                     const long long* __restrict__ chunk_off, T* __restrict__ dst, int nan_to_num) {
  constexpr int VEC = 16 / sizeof(T);
  __shared__ long long s_src;
  __shared__ int s_valid;
  const long long c = blockIdx.x;
  if (threadIdx.x == 0) {
    const long long v = find_segment(chunk_off, V, c);
    const long long k = c - chunk_off[v];
    const long long Tv = row_off[v + 1] - row_off[v];
    long long valid = Tv - k * length;
    valid = valid < 0 ? 0 : (valid > length ? length : valid);
    s_src = row_off[v] + k * length;
    s_valid = int(valid);
  }
  __syncthreads();
  const int dv = D / VEC;
  const T* sp = src + s_src * D;
  T* dp = dst + c * (long long)length * D;
  const int nvalid = s_valid * dv, ntotal = length * dv;
  for (int i = threadIdx.x; i < ntotal; i += blockDim.x) {
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (i < nvalid) {
      u = __ldcs(reinterpret_cast<const uint4*>(sp) + i);
      if (nan_to_num) {
        T* e = reinterpret_cast<T*>(&u);
#pragma unroll
        for (int j = 0; j < VEC; ++j) e[j] = fix<T>(e[j]);
      }
    }
    reinterpret_cast<uint4*>(dp)[i] = u;
  }
}

// np.linspace(0, T, num + 1, dtype=int32)[i]: y = i * (T / num) in float64, the last point is T exactly; truncation
__device__ __forceinline__ int linspace_edge(long long T, int num, int i) {
  if (i >= num) return int(T);
  const double step = double(T) / double(num);
  return int(double(i) * step);
}

template <typename T>
__global__ void __launch_bounds__(256)
process_feat_kernel(const T* __restrict__ src, const long long* __restrict__ row_off, int D, int length,
                    float* __restrict__ dst, long long* __restrict__ out_len, int nan_to_num) {
  const long long v = blockIdx.y;
  const int i = blockIdx.x;
  const long long Tv = row_off[v + 1] - row_off[v];
  const T* sp = src + row_off[v] * D;
  float* dp = dst + (v * length + i) * (long long)D;
  if (i == 0 && threadIdx.x == 0) out_len[v] = Tv > length ? length : Tv;
  long long a, b;
  if (Tv > length) {                               // uniform_extract, data/tools.py:65-73
    a = linspace_edge(Tv, length, i);
    b = linspace_edge(Tv, length, i + 1);
    if (a == b) b = a + 1;                         // empty bin -> the single row feat[r[i]]
  } else {                                         // pad, :81-86
    a = i;
    b = (i < Tv) ? i + 1 : i;
  }
  const long long n = b - a;
  for (int col = threadIdx.x; col < D; col += blockDim.x) {
    float acc = 0.f;
    for (long long r = a; r < b; ++r) {            // np.mean(axis=0): rows added in order, float32 accumulator
      T e = sp[r * D + col];
      if (nan_to_num) e = fix<T>(e);
      acc += to_f32<T>(e);
    }
    float out = 0.f;
    if (n == 1) out = acc;
    else if (n > 1) {
      out = acc / float(n);
      if (sizeof(T) == 2) out = to_f32<T>(from_f32<T>(out));   // numpy returns the mean of a 16-bit array in that type
    }
    dp[col] = out;
  }
}

}  // namespace

int process_split(const void* src, int dtype, const long long* row_off, long long V, int D, int length,
                  const long long* chunk_off, long long total_chunks, void* dst, int nan_to_num, cudaStream_t stream) {
  IEF_CHECK(V >= 0 && total_chunks >= 0 && D > 0 && length > 0, "process_split: bad sizes");
  IEF_CHECK(total_chunks < (1LL << 31), "process_split: too many chunks");
  if (V == 0 || total_chunks == 0) return IEFVAD_OK;
  IEF_CHECK(src && row_off && chunk_off && dst, "process_split: null argument");
  const int es = dtype == IEFVAD_DT_F32 ? 4 : 2;
  IEF_CHECK((D * es) % 16 == 0, "process_split: rows must be multiples of 16 bytes (D=%d)", D);
  const unsigned grid = unsigned(total_chunks);
  switch (dtype) {
    case IEFVAD_DT_F32:
      process_split_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(src), row_off, V, D, length,
                                                             chunk_off, static_cast<float*>(dst), nan_to_num);
      break;
    case IEFVAD_DT_F16:
      process_split_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half*>(src), row_off, V, D, length,
                                                              chunk_off, static_cast<__half*>(dst), nan_to_num);
      break;
    case IEFVAD_DT_BF16:
      process_split_kernel<bf16><<<grid, 256, 0, stream>>>(static_cast<const bf16*>(src), row_off, V, D, length,
                                                            chunk_off, static_cast<bf16*>(dst), nan_to_num);
      break;
    default: set_error("process_split: unsupported dtype code %d", dtype); return IEFVAD_ERR_INVALID;
  }
  count_launches(1);
  IEF_CUDA(cudaGetLastError());
  return IEFVAD_OK;
}

int process_feat(const void* src, int dtype, const long long* row_off, long long V, int D, int length, float* dst,
                 long long* out_len, int nan_to_num, cudaStream_t stream) {
  IEF_CHECK(V >= 0 && V <= 65535 && D > 0 && length > 0, "process_feat: bad sizes (V <= 65535)");
  if (V == 0) return IEFVAD_OK;
  IEF_CHECK(src && row_off && dst && out_len, "process_feat: null argument");
  const dim3 grid = dim3(static_cast<unsigned>(length), static_cast<unsigned>(V), 1);
  switch (dtype) {
    case IEFVAD_DT_F32:
      process_feat_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(src), row_off, D, length, dst, out_len, nan_to_num);
      break;
    case IEFVAD_DT_F16:
      process_feat_kernel<__half><<<grid, 256, 0, stream>>>(static_cast<const __half*>(src), row_off, D, length, dst, out_len, nan_to_num);
      break;
    case IEFVAD_DT_BF16:
      process_feat_kernel<bf16><<<grid, 256, 0, stream>>>(static_cast<const bf16*>(src), row_off, D, length, dst, out_len, nan_to_num);
      break;
    default: set_error("process_feat: unsupported dtype code %d", dtype); return IEFVAD_ERR_INVALID;
  }
  count_launches(1);
  IEF_CUDA(cudaGetLastError());
  return IEFVAD_OK;
}

}  // namespace iefvad
