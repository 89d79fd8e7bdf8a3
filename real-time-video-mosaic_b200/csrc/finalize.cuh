// finalize.cuh -- device side of crop_black_areas + scale_to_screen (finalize.cu)
#pragma once
#include "common.cuh"

// d_bounds[4] = (min x, min y, max x, max y) of the canvas pixels whose BGR2GRAY value exceeds thr; (INT_MAX, INT_MAX, -1, -1) if none
cudaError_t bm_launch_crop_bounds(const uchar4* canvas, int w, int h, int thr, int* d_bounds, cudaStream_t s);
// cv2.resize(canvas[ry:ry+sh, rx:rx+sw], (dw, dh), INTER_LINEAR) -> packed BGR
cudaError_t bm_launch_resize_linear(const uchar4* canvas, int cw, int rx, int ry, int sw, int sh, uint8_t* d_out, int dw, int dh, cudaStream_t s);
