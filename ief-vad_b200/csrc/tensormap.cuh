// Host-side CUtensorMap construction without linking libcuda: cuTensorMapEncodeTiled is resolved through the
// runtime's driver entry point query.
#pragma once
#include "common.cuh"

namespace iefvad {

enum : int { TM_BF16 = 0, TM_F32 = 1 };
enum : int { TM_SWIZZLE_NONE = 0, TM_SWIZZLE_64B = 2, TM_SWIZZLE_128B = 3 };

// 2-D row-major tensor [outer, inner] with a row pitch of row_stride_bytes (multiple of 16).
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                 uint32_t box_inner, uint32_t box_outer, int dtype = TM_BF16, int swizzle = TM_SWIZZLE_128B);
// dims {d0 (contiguous), d1, d2}; strides in bytes for d1 and d2 (bf16, 128B swizzle).
int make_tmap_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                 uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2, int swizzle = TM_SWIZZLE_128B);

// dims {d0 (contiguous), d1, d2, d3}; strides in bytes for d1..d3.
int make_tmap_4d(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                 const uint32_t box[4], int dtype, int swizzle);

}  // namespace iefvad
