#pragma once
#include "common.cuh"
cudaError_t bm_launch_ingest(const uint8_t* d_bgr, int h, int w, uint8_t* d_gray, uchar4* d_bgrx, cudaStream_t s);
