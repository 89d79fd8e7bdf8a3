// pipeline.cu -- detect / match / RANSAC orchestration (placeholder until the detector kernels land).
#include "pipeline.cuh"
#include <new>

struct BmPipeline { bm_config cfg; cudaStream_t stream; };

bm_status bm_pipeline_create(BmPipeline** out, const bm_config& cfg, cudaStream_t stream) {
    BmPipeline* p = new (std::nothrow) BmPipeline();
    if (!p) return BM_ERR_ARG;
    p->cfg = cfg; p->stream = stream;
    *out = p;
    return BM_OK;
}
void bm_pipeline_destroy(BmPipeline* p) { delete p; }
bm_status bm_pipeline_first_frame(BmPipeline*, const uint8_t*) { return BM_OK; }
bm_status bm_pipeline_estimate(BmPipeline*, const uint8_t*, bm_frame_info*, double*, int*) {
    bm_set_error("feature pipeline not built yet");
    return BM_ERR_UNSUPPORTED;
}
void bm_pipeline_advance(BmPipeline*) {}
