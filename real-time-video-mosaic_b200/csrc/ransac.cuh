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
    int jacobi_sweeps;       // bits 0-7: inverse-iteration steps of the DLT; 8-15: LM iterations that took the eigen-decomposition route; 16+: sweeps of the last one
    long long cyc[8];        // SM cycles per phase: 0 subsets, 1 hypotheses, 2 selection, 3 mask+refit sums, 4 jacobi, 5 LM, 6 total
};

// src/dst: device float2[n] (n read from *d_count); thresh = ransacReprojThreshold; scratch: >= n bytes
cudaError_t bm_launch_ransac(const float2* d_src, const float2* d_dst, const int* d_count, double thresh, int max_iters,
                             double confidence, uint8_t* d_mask, BmRansacResult* d_out, cudaStream_t s);
// debug / tests: make every LM iteration of k_ransac_refine take the eigen-decomposition route (eig9.h)
cudaError_t bm_lm_force_eig(int on);
// counters of the current device since the last reset: polishes run, LM iterations, iterations that took the eigen-decomposition route
cudaError_t bm_lm_stats(unsigned long long out[3], int reset);
