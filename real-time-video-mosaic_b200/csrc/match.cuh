// match.cuh -- brute-force matchers of VideMosaic.match (main.py:676-698) + point gathering of findHomography (:849-854)
#pragma once
#include "common.cuh"
#include "orb.cuh"

struct BmMatches {
    int* q; int* t; float* dist;      // sorted by distance (stable), capacity BM_KP_CAP
    float2* src; float2* dst;         // kp_cur[q].pt, kp_prev[t].pt
    int* count;                       // device scalar
    // scratch
    int *nn_q2t, *nn_t2q, *tq, *tt; float *d_q2t, *d_t2q, *td; float* d2_q2t;
    int* nn2_q2t;
    int4* l2_part;                    // per-split (best, second) pairs of the tensor-core matcher
};
int bm_matches_alloc(BmMatches* m);
void bm_matches_free(BmMatches* m);
// ORB: BFMatcher(NORM_HAMMING, crossCheck=True).match(des_cur, des_prev) + sorted(key=distance)
cudaError_t bm_match_hamming(const BmKeypoints& cur, const BmKeypoints& prev, BmMatches& m, cudaStream_t s);
// SIFT: BFMatcher().knnMatch(des_cur, des_prev, k=2) + ratio 0.7 + sorted(key=distance); descriptors are u8 (exact integers)
cudaError_t bm_match_l2_ratio(const BmKeypoints& cur, const BmKeypoints& prev, BmMatches& m, double ratio, cudaStream_t s);
// tensor-core kNN (match_tc.cu): two nearest rows of B per row of A under exact integer L2, ties -> lower index
// part: scratch, BM_L2_SPLIT * BM_KP_CAP int4
cudaError_t bm_launch_l2_knn2_tc(const uint8_t* A, const int* nA, const uint8_t* B, const int* nB, int4* part, int* nn1, float* d1, int* nn2,
                                 float* d2, cudaStream_t s);
#define BM_L2_SPLIT 8
