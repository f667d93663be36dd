// Temporal-localisation mAP kernels (locmap.cu): SURVEY 8f row N5, reference train/metrics.py:19-136.
#pragma once
#include "common.cuh"

namespace iefvad {

constexpr int kLocmapMaxProposals = 512;     // proposals kept per (video, class) after NMS (count -1 = more were found)

// pred: all videos' [T_v, C] fp32 predictions back to back (video v starts at row vid_off[v], has vid_len[v] <= 4096 rows).
// prop_count [V, C], prop_se [V, C, 512, 2], prop_score [V, C, 512], class_score [V, C]
int locmap_proposals(const float* pred, const long long* vid_off, const int* vid_len, int V, int C, int max_len,
                     int* prop_count, int* prop_se, float* prop_score, float* class_score, cudaStream_t stream);
// gt [n_gt, 3] = (video, start, end) grouped by class, class c owning rows [gt_off[c], gt_off[c + 1]); ap [C] float64,
// n_pred [C]; workspaces: w_score / w_idx [C, cap], w_alive [n_gt]
int locmap_match(const int* prop_count, const int* prop_se, const float* prop_score, int V, int C, const int* gt,
                 const int* gt_off, double th, int cap, float* w_score, int* w_idx, int* w_alive, double* ap, int* n_pred,
                 cudaStream_t stream);

}  // namespace iefvad
