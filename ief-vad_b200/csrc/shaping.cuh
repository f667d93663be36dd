// Row N2: the input shaping that precedes the forward in the reference's data layer (data/tools.py:65-114), on the
// device, for whole video lists at once.  Videos arrive packed back to back: src [sum T_v, D] with row offsets.
#pragma once
#include "common.cuh"

namespace iefvad {

// process_split (data/tools.py:100-114) of V videos: video v becomes int(T_v / length) + 1 chunks of `length` rows
// (one chunk when T_v < length), zero-padded - including the extra all-zero chunk when T_v % length == 0.
// row_off, chunk_off: DEVICE int64 [V + 1] (prefix sums of T_v / of the chunk counts); dst [chunk_off[V], length, D]
// in the input's element type; nan_to_num != 0 applies torch.nan_to_num (train/ucf_test.py:83-88) on the way.
int process_split(const void* src, int dtype, const long long* row_off, long long V, int D, int length,
                  const long long* chunk_off, long long total_chunks, void* dst, int nan_to_num, cudaStream_t stream);

// process_feat (data/tools.py:89-97, non-random branch) of V videos -> dst [V, length, D] fp32 and out_len [V]:
// T_v > length: uniform_extract (:65-78) = mean over np.linspace(0, T_v, length + 1, dtype=int32) bins (a single row when
// a bin is empty), out_len = length; else zero-pad, out_len = T_v.
int process_feat(const void* src, int dtype, const long long* row_off, long long V, int D, int length, float* dst,
                 long long* out_len, int nan_to_num, cudaStream_t stream);

}  // namespace iefvad
