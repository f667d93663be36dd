// Host-side CUtensorMap construction (bf16, 128-byte swizzle) without linking libcuda:
// cuTensorMapEncodeTiled is resolved through the runtime's driver entry point query.
#pragma once
#include "common.cuh"

namespace iefvad {

// inner = contiguous dimension (elements), row_stride_bytes must be a multiple of 16.
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                 uint32_t box_inner, uint32_t box_outer);
// dims {d0 (contiguous), d1, d2}; strides in bytes for d1 and d2.
int make_tmap_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                 uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2);

}  // namespace iefvad
