#pragma once
#include "common.cuh"

namespace iefvad {

// Synthetic event frames (row N4), extracting/ucf_gen_event.py:21-37 + :91-95.  frames: uint8 [B, C, H, W, 3] (C frames per
// stack, 16 in the reference).  sum_out [B, H, W] = number of frame-to-frame gray differences above `threshold`;
// event_out [B, 3, H, W] = clamp(sum, 0, clamp_max) / (max over the whole batch), replicated over 3 channels.
int event_image(const uint8_t* frames, long long B, int C, int H, int W, float threshold, float clamp_max,
                float* sum_out, float* event_out, float* scratch_cnt, unsigned* scratch_max, int num_sms,
                cudaStream_t stream);

}  // namespace iefvad
