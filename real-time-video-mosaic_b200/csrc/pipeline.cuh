// pipeline.cuh -- per-frame feature pipeline (detect -> match -> RANSAC) behind bm_process_frame.
#pragma once
#include "../../include/b200mosaic.h"
#include "common.cuh"

struct BmPipeline;
bm_status bm_pipeline_create(BmPipeline** out, const bm_config& cfg, cudaStream_t stream);
void bm_pipeline_destroy(BmPipeline* p);
// features of frame 0 become "prev" (main.py:104-112)
bm_status bm_pipeline_first_frame(BmPipeline* p, const uint8_t* d_gray);
// features of the current frame, matches against prev, RANSAC homography cur->prev (main.py:717-727)
bm_status bm_pipeline_estimate(BmPipeline* p, const uint8_t* d_gray, bm_frame_info* info, double H_rel[9], int* have_h);
// split form of bm_pipeline_estimate: enqueue (no wait) / wait + read back
bm_status bm_pipeline_estimate_begin(BmPipeline* p, const uint8_t* d_gray);
#define BM_AHEAD_MAX 3           // frames whose features may be computed ahead of the current one
#define BM_KP_SLOTS (BM_AHEAD_MAX + 2)   // previous, current and the frames detected ahead
#define BM_NDET 3                // detector instances (one per detect that can be in flight)
bm_status bm_pipeline_detect_ahead(BmPipeline* p, const uint8_t* d_gray, int* done);
void bm_pipeline_drop_ahead(BmPipeline* p, const uint8_t* d_gray);
bm_status bm_pipeline_estimate_end(BmPipeline* p, bm_frame_info* info, double H_rel[9], int* have_h);
// cur -> prev (main.py:756-759)
void bm_pipeline_advance(BmPipeline* p);
cudaError_t bm_pipeline_sync_est(BmPipeline* p);      // detect (BM_NDET streams, round robin) and match + RANSAC run on the pipeline's own streams
bm_status bm_pipeline_warm_up(BmPipeline* p, const uint8_t* const* d_gray, int n_gray);
cudaEvent_t bm_pipeline_last_detect_event(BmPipeline* p);
cudaError_t bm_pipeline_record_after_last_detect(BmPipeline* p, cudaEvent_t ev);   // `ev` completes when the most recently queued detect has   // completion of the most recently queued detect (owned by the pipeline)
struct BmKeypoints; struct BmMatches;
BmKeypoints* bm_pipeline_keypoints(BmPipeline* p, int which /*0 prev, 1 cur*/);
BmMatches* bm_pipeline_matches(BmPipeline* p);
