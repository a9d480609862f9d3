// Host-side construction of TMA tensor maps over NHWC bf16 views and packed weight matrices.
// cuTensorMapEncodeTiled is fetched through the runtime's driver entry point, so the library has no link-time
// dependency on libcuda (it must load on the GPU-less build box).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace cvb {

// 4-D map (C, W, H, N) over an activation view; box = (64 channels, bw, bh, bn) pixels, 128-byte swizzle,
// out-of-bounds elements read as zero (this is the convolution's zero padding).
int make_act_tmap(CUtensorMap* out, const cvb_view& v, int box_w, int box_h, int box_n);
// 5-D map (C, W, row parity, H/2, N) over an activation view of even height: a box of `box_pairs` rows of ONE parity,
// i.e. every second image row (the transposed cout = 64 conv kernel reads rows 2i + d). Same swizzle / zero fill.
int make_act_tmap_rowpairs(CUtensorMap* out, const cvb_view& v, int box_w, int box_pairs);
// 2-D map over a row-major bf16 matrix [rows][cols] (cols contiguous); box = (64 cols, box_rows), 128-byte swizzle.
int make_mat_tmap(CUtensorMap* out, const void* ptr, long long rows, long long cols, int box_rows);

}  // namespace cvb
