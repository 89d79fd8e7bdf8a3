// ransac.cuh -- cv2.findHomography(p_cur, p_prev, RANSAC, 2.0) on the device (main.py:856-857)
#pragma once
#include "common.cuh"

struct BmRansacResult {      // written by the kernel, read back by the host (one small D2H per frame)
    double H[9];
    int ok;                  // 1 -> H valid; 0 -> findHomography would return None
    int n_points;
    int iters;               // RANSAC iterations executed
    int n_inliers;           // inliers of the winning hypothesis
    int lm_iters;
    int jacobi_sweeps;
    long long cyc[8];        // SM cycles per phase: 0 subsets, 1 hypotheses, 2 selection, 3 mask+refit sums, 4 jacobi, 5 LM, 6 total
};

// src/dst: device float2[n] (n read from *d_count); thresh = ransacReprojThreshold; scratch: >= n bytes
cudaError_t bm_launch_ransac(const float2* d_src, const float2* d_dst, const int* d_count, double thresh, int max_iters,
                             double confidence, uint8_t* d_mask, BmRansacResult* d_out, cudaStream_t s);
