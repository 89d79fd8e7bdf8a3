// sift.cuh -- SIFT (cv2.SIFT_create(700)) detector object.
#pragma once
#include "common.cuh"
#include "orb.cuh"     // BmKeypoints

struct BmSift;
int bm_sift_create(BmSift** out, int h, int w, int nfeatures, cudaStream_t s);
void bm_sift_destroy(BmSift* o);
// launch == false: only capture + instantiate the graph of this (input, output) pair if it is not cached yet
cudaError_t bm_sift_detect(BmSift* o, const uint8_t* d_gray, BmKeypoints* out, bool launch = true);
const float* bm_sift_level_ptr(BmSift* o, int octave, int level, int dog, int* w, int* h);
int bm_sift_num_octaves(BmSift* o);
// device counters of the last detect (cand, kp, overflow flag, selected, raw, threshold bits, kp after pass A, listed candidates)
void bm_sift_counters(BmSift* o, int out[8]);
// pyramid kernels only, reps times, CUDA-event timed on the detector's stream (bench.py roofline_pyramid)
cudaError_t bm_sift_time_pyramid(BmSift* o, const uint8_t* d_gray, int reps, float* ms_total);
