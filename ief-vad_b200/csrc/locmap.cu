// Temporal-localisation mAP (reference: train/metrics.py:19-136 getDetectionMAP / getLocMAP / nms,
// imported by train/ucf_test.py:13 and never called).  Two kernels:
//   locmap_proposals_kernel   one CTA per (video, class): class score = mean of the top T/16 values of the column
//                             (:60-64), threshold at max - 0.6 (max - min) (:75), runs of >= 2 frames above it become
//                             proposals scored max + 0.7 class score (:77-83), sorted by score, greedy NMS at IoU 0.6 (:19-41)
//   locmap_match_kernel       one CTA per class: all proposals sorted by score (:96), greedy matching against the class's
//                             ground-truth segments with deletion of matched ones (:104-122), AP = sum(prec * tp) / #gt (:123-128)
// The arithmetic follows numpy's types: the threshold is float64 (float32 column, float64 0.6), proposal scores float32,
// NMS IoUs float64, matching IoUs ratios of integer set sizes, AP float64.
#include <cmath>

#include "common.cuh"
#include "locmap.cuh"

namespace iefvad {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxT = 4096;                 // longest column one CTA sorts in shared memory
constexpr int kMaxProp = kLocmapMaxProposals;

// in-place bitonic sort of n (power of two) shared-memory keys, descending; idx (optional) permuted alongside
__device__ void bitonic_desc(float* key, int* idx, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int p = i ^ j;
        if (p > i) {
          const bool up = (i & k) == 0;       // descending overall: "up" blocks put the larger key first
          const float a = key[i], b = key[p];
          // ties keep the lower original index first (numpy's argsort(-x) is not stable, but equal scores are measure-zero)
          const bool swap = up ? (a < b || (a == b && idx && idx[i] > idx[p])) : (a > b || (a == b && idx && idx[i] < idx[p]));
          if (swap) {
            key[i] = b; key[p] = a;
            if (idx) { const int t = idx[i]; idx[i] = idx[p]; idx[p] = t; }
          }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(kThreads)
locmap_proposals_kernel(const float* __restrict__ pred, const long long* __restrict__ vid_off, const int* __restrict__ vid_len,
                        int C, int* __restrict__ prop_count, int* __restrict__ prop_se, float* __restrict__ prop_score,
                        float* __restrict__ class_score) {
  __shared__ float col[kMaxT];
  __shared__ float srt[kMaxT];
  __shared__ int seg_s[kMaxProp], seg_e[kMaxProp], seg_i[kMaxProp], keep[kMaxProp];
  __shared__ float seg_sc[kMaxProp];
  __shared__ int n_seg, n_keep;
  __shared__ float red[kThreads];
  const int v = blockIdx.x, c = blockIdx.y;
  const int T = vid_len[v];
  const float* p = pred + vid_off[v] * C + c;
  int n2 = 1;
  while (n2 < T) n2 <<= 1;
  for (int t = threadIdx.x; t < n2; t += blockDim.x) {
    const float x = t < T ? p[(long long)t * C] : -INFINITY;
    if (t < T) col[t] = x;
    srt[t] = x;
  }
  __syncthreads();
  bitonic_desc(srt, nullptr, n2);
  // class score: mean of the top int(T / 16) values (np.mean of a float32 slice: pairwise float32 sum; plain order here)
  const int k = T / 16;
  float part = 0.f;
  for (int t = threadIdx.x; t < k; t += blockDim.x) part += srt[t];
  red[threadIdx.x] = part;
  __syncthreads();
  for (int s = kThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  const float cs = k > 0 ? red[0] / float(k) : NAN;                 // mean of an empty slice is NaN (:62)
  const long long slot = (long long)v * C + c;
  if (threadIdx.x == 0) class_score[slot] = cs;
  const bool ind = cs > 0.f;                                         // :63 (NaN > 0 is False)
  if (threadIdx.x == 0) {
    n_seg = 0;
    n_keep = 0;
    if (ind && T > 0) {
      const double mx = double(srt[0]), mn = double(srt[T - 1]);
      const double thr = mx - (mx - mn) * 0.6;                       // :75, thr_set = [0.6]
      int start = -1;
      for (int t = 0; t <= T; ++t) {
        const bool on = t < T && double(col[t]) > thr;
        if (on && start < 0) start = t;
        if (!on && start >= 0) {
          if (t - start >= 2 && n_seg >= kMaxProp) n_keep = -1;      // more runs than the proposal table holds: reported
          if (t - start >= 2 && n_seg < kMaxProp) {                  // :80
            float m = col[start];
            for (int u = start + 1; u < t; ++u) m = fmaxf(m, col[u]);
            seg_s[n_seg] = start; seg_e[n_seg] = t;
            seg_sc[n_seg] = m + 0.7f * cs;                           // :81 (float32 under NEP 50)
            seg_i[n_seg] = n_seg;
            ++n_seg;
          }
          start = -1;
        }
      }
    }
  }
  __syncthreads();
  const int ns = n_seg;
  if (n_keep < 0) {
    if (threadIdx.x == 0) prop_count[slot] = -1;
    return;
  }
  if (ns > 0) {
    int m2 = 1;
    while (m2 < ns) m2 <<= 1;
    for (int i = ns + threadIdx.x; i < m2; i += blockDim.x) { seg_sc[i] = -INFINITY; seg_i[i] = i; }
    __syncthreads();
    bitonic_desc(seg_sc, seg_i, m2);                                 // :86 argsort(-score)
    if (threadIdx.x == 0) {                                          // greedy NMS, :19-41 (thresh 0.6)
      int alive[kMaxProp];
      for (int i = 0; i < ns; ++i) alive[i] = 1;
      for (int i = 0; i < ns; ++i) {
        if (!alive[i]) continue;
        keep[n_keep++] = i;
        const double s1 = seg_s[seg_i[i]], e1 = seg_e[seg_i[i]];
        for (int j = i + 1; j < ns; ++j) {
          if (!alive[j]) continue;
          const double s2 = seg_s[seg_i[j]], e2 = seg_e[seg_i[j]];
          const double inter = fmax(0.0, fmin(e1, e2) - fmax(s1, s2));
          const double ovr = inter / ((e1 - s1) + (e2 - s2) - inter);
          if (!(ovr <= 0.6)) alive[j] = 0;
        }
      }
    }
    __syncthreads();
  }
  const int nk = n_keep;
  if (threadIdx.x == 0) prop_count[slot] = nk;
  for (int i = threadIdx.x; i < nk; i += blockDim.x) {
    const int s = seg_i[keep[i]];
    prop_se[(slot * kMaxProp + i) * 2] = seg_s[s];
    prop_se[(slot * kMaxProp + i) * 2 + 1] = seg_e[s];
    prop_score[slot * kMaxProp + i] = seg_sc[keep[i]];
  }
}

// one CTA per class; scratch per class: cap proposals (video, s, e, score)
__global__ void __launch_bounds__(kThreads)
locmap_match_kernel(const int* __restrict__ prop_count, const int* __restrict__ prop_se, const float* __restrict__ prop_score,
                    int V, int C, const int* __restrict__ gt, const int* __restrict__ gt_off /* [C + 1] */, double th, int cap,
                    float* __restrict__ w_score, int* __restrict__ w_idx, int* __restrict__ w_alive, double* __restrict__ ap,
                    int* __restrict__ n_pred) {
  __shared__ int total;
  const int c = blockIdx.x;
  float* score = w_score + (long long)c * cap;
  int* idx = w_idx + (long long)c * cap;
  // gather this class's proposals of all videos, in video order (segment_predict.extend, :88)
  if (threadIdx.x == 0) {
    int n = 0;
    for (int v = 0; v < V; ++v) {
      const int cnt = prop_count[(long long)v * C + c];
      for (int i = 0; i < cnt && n < cap / 2; ++i) {      // (the sort pads to a power of two <= cap)
        score[n] = prop_score[((long long)v * C + c) * kMaxProp + i];
        idx[n] = v * kMaxProp + i;
        ++n;
      }
    }
    total = n;
    n_pred[c] = n;
  }
  __syncthreads();
  const int n = total;
  if (n == 0) {
    if (threadIdx.x == 0) ap[c] = 0.0;
    return;
  }
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  for (int i = n + threadIdx.x; i < n2; i += blockDim.x) { score[i] = -INFINITY; idx[i] = 0x7fffffff; }
  __syncthreads();
  // global-memory bitonic sort (a few thousand proposals per class at most), :96
  for (int k = 2; k <= n2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        const int p = i ^ j;
        if (p > i) {
          const bool up = (i & k) == 0;
          const float a = score[i], b = score[p];
          const bool swap = up ? (a < b || (a == b && idx[i] > idx[p])) : (a > b || (a == b && idx[i] < idx[p]));
          if (swap) { score[i] = b; score[p] = a; const int t = idx[i]; idx[i] = idx[p]; idx[p] = t; }
        }
      }
      __syncthreads();
    }
  if (threadIdx.x != 0) return;
  const int g0 = gt_off[c], g1 = gt_off[c + 1];
  int* alive = w_alive + g0;
  for (int j = g0; j < g1; ++j) w_alive[j] = 1;
  double tp_c = 0.0, fp_c = 0.0, acc = 0.0;
  bool any_tp = false;
  for (int i = 0; i < n; ++i) {                                      // :104-122
    const int v = idx[i] / kMaxProp, pi = idx[i] % kMaxProp;
    const long long slot = ((long long)v * C + c) * kMaxProp + pi;
    const int ps = prop_se[slot * 2], pe = prop_se[slot * 2 + 1];
    double best = 0.0;
    int best_j = -1;
    bool flag = false;
    for (int j = g0; j < g1; ++j) {
      if (!alive[j - g0] || gt[j * 3] != v) continue;
      const int gs = gt[j * 3 + 1], ge = gt[j * 3 + 2];
      const int lg = ge > gs ? ge - gs : 0, lp = pe > ps ? pe - ps : 0;       // len(range(...))
      int inter = min(pe, ge) - max(ps, gs);
      if (inter < 0 || lg == 0 || lp == 0) inter = 0;
      const int uni = lg + lp - inter;
      if (uni == 0) continue;                                        // (the reference would raise ZeroDivisionError)
      const double iou = double(inter) / double(uni);
      if (iou >= th) {
        flag = true;
        if (iou > best) { best = iou; best_j = j; }
      }
    }
    if (flag) {
      alive[best_j - g0] = 0;                                        // del segment_gt[best_j]
      tp_c += 1.0;
      acc += tp_c / (fp_c + tp_c);                                   // (tp_c / (fp_c + tp_c)) * tp, summed
      any_tp = true;
    } else {
      fp_c += 1.0;
    }
  }
  ap[c] = any_tp ? acc / double(g1 - g0) : 0.0;                      // :123-127
}

}  // namespace

int locmap_proposals(const float* pred, const long long* vid_off, const int* vid_len, int V, int C, int max_len,
                     int* prop_count, int* prop_se, float* prop_score, float* class_score, cudaStream_t stream) {
  IEF_CHECK(max_len <= kMaxT, "locmap: a video of %d segments exceeds the %d one CTA sorts", max_len, kMaxT);
  if (V == 0 || C == 0) return IEFVAD_OK;
  locmap_proposals_kernel<<<dim3(V, C), kThreads, 0, stream>>>(pred, vid_off, vid_len, C, prop_count, prop_se, prop_score,
                                                                class_score);
  count_launches(1);
  IEF_CUDA(cudaGetLastError());
  return IEFVAD_OK;
}

int locmap_match(const int* prop_count, const int* prop_se, const float* prop_score, int V, int C, const int* gt,
                 const int* gt_off, double th, int cap, float* w_score, int* w_idx, int* w_alive, double* ap, int* n_pred,
                 cudaStream_t stream) {
  if (C == 0) return IEFVAD_OK;
  locmap_match_kernel<<<C, kThreads, 0, stream>>>(prop_count, prop_se, prop_score, V, C, gt, gt_off, th, cap, w_score, w_idx,
                                                  w_alive, ap, n_pred);
  count_launches(1);
  IEF_CUDA(cudaGetLastError());
  return IEFVAD_OK;
}

}  // namespace iefvad
