// pipeline.cu -- per-frame feature pipeline: detectAndCompute -> match -> findHomography (main.py:717-727) on the device,
// one small D2H read (counts + RANSAC result) per frame for the reference's host-side control flow.
#include "pipeline.cuh"
#include "orb.cuh"
#include "sift.cuh"
#include "match.cuh"
#include "ransac.cuh"
#include <new>
#include <string.h>
#include <stdlib.h>
#include <time.h>

struct BmHostReadback {      // pinned + mapped: written by k_publish_readback straight from the device, polled by the host
    BmRansacResult r;
    int n_cur, n_prev, n_matches, overflow_cur, overflow_prev;
    unsigned seq;            // sequence number of the estimate whose results are above (written last, after a system-wide fence)
};

// Last kernel of match + RANSAC: the frame's results go to the host's pinned buffer in ONE step (a store over PCIe + a flag the host
// polls) instead of six small D2H copies and an event -- those were ~15 us of serialised stream operations per frame on the path
// the frame rate is bound by (one frame period = match + RANSAC latency + host turnaround).
__global__ void __launch_bounds__(32) k_publish_readback(const BmRansacResult* __restrict__ r, const int* __restrict__ n_cur, const int* __restrict__ n_prev,
                                                         const int* __restrict__ n_matches, const int* __restrict__ of_cur, const int* __restrict__ of_prev,
                                                         BmHostReadback* __restrict__ h, unsigned seq) {
    BM_PDL_WAIT();
    const int* src = reinterpret_cast<const int*>(r);
    int* dst = reinterpret_cast<int*>(&h->r);
    for (int i = threadIdx.x; i < (int)(sizeof(BmRansacResult) / sizeof(int)); i += 32) dst[i] = src[i];
    if (threadIdx.x == 0) { h->n_cur = *n_cur; h->n_prev = *n_prev; h->n_matches = *n_matches; h->overflow_cur = *of_cur; h->overflow_prev = *of_prev; }
    __threadfence_system();
    __syncwarp();
    if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned*>(&h->seq) = seq;
}

struct BmPipeline {
    bm_config cfg;
    cudaStream_t stream;
    // BM_NDET detector instances on their own streams, used round robin: consecutive frames' detectAndCompute (all but the first
    // queued as detect-aheads) run concurrently, so the latency-bound tail of one frame (selection, orientation, descriptors -- a few
    // small or one-CTA kernels) is covered by the pyramids of the next ones.  `stream` only carries the orderings the caller sets up;
    // every detect forks from it (ev_fork) and publishes ev_det[slot].
    BmOrb* orb[BM_NDET] = {};
    BmSift* sift[BM_NDET] = {};
    cudaStream_t s_det[BM_NDET] = {};
    cudaEvent_t ev_fork = nullptr;
    int det_next = 0, det_last = 0, last_det_slot = 0;
    bool is_orb = false;
    BmKeypoints kp[BM_KP_SLOTS];   // previous / current / detected ahead (the next frames' features may be computed before
    int prev = 0, cur = 1;         // the host knows whether the current frame becomes "previous")
    const uint8_t* ahead_gray[BM_AHEAD_MAX] = {};   // gray buffers whose features were enqueued into kp[ahead_slot[i]] by bm_pipeline_detect_ahead
    int ahead_slot[BM_AHEAD_MAX] = {};
    cudaEvent_t ev_done = nullptr;           // RANSAC result + counts of the current frame are in the pinned readback
    // match + RANSAC run on their own stream: a handful of small, latency-bound launches (one-CTA RANSAC stages, selection sort) that
    // would otherwise sit between two detects on the detect stream; with a detect-ahead queued they overlap the next frame's pyramid
    cudaStream_t s_est = nullptr;
    cudaEvent_t ev_det[BM_KP_SLOTS] = {};    // features of the slot are complete (recorded on the detect stream)
    cudaEvent_t ev_est = nullptr;            // last match that read the keypoint slots finished (a detect may overwrite a slot)
    BmMatches m[2];          // double buffered: the next frame may be matched while the last one's matches are still readable
    int mcur = 0, mdone = 0;
    uint8_t* d_mask = nullptr;
    BmRansacResult* d_res = nullptr;
    BmHostReadback* h_rb = nullptr;
    unsigned seq = 0;        // estimates issued so far (k_publish_readback stamps the readback with it)
    bool have_prev = false;
    // BM_PROFILE=1: in-pipeline latency of the detect graph and of match + RANSAC (CUDA events on their own streams), printed at destroy
    bool prof = false;
    cudaEvent_t pd0[BM_NDET] = {}, pd1[BM_NDET] = {}, pe0 = nullptr, pe1 = nullptr;
    bool pd_pending[BM_NDET] = {}, pe_pending = false;
    double pd_ms = 0.0, pe_ms = 0.0; long pd_n = 0, pe_n = 0;
    double wait_us = 0.0; long wait_n = 0;    // host time spent waiting for the result in estimate_end
};

bm_status bm_pipeline_create(BmPipeline** out, const bm_config& cfg, cudaStream_t stream) {
    BmPipeline* p = new (std::nothrow) BmPipeline();
    if (!p) return BM_ERR_ARG;
    p->cfg = cfg; p->stream = stream;
    memset(p->kp, 0, sizeof(p->kp)); memset(p->m, 0, sizeof(p->m));
    const int desc_bytes = cfg.detector == BM_DET_ORB ? 32 : 128;
    bool ok = cudaEventCreateWithFlags(&p->ev_done, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&p->ev_est, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < BM_KP_SLOTS && ok; ++i)
        ok = bm_kp_alloc(&p->kp[i], desc_bytes) == 0 && cudaEventCreateWithFlags(&p->ev_det[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok &&
              bm_stream_create(&p->s_est, 2) == cudaSuccess && bm_matches_alloc(&p->m[0]) == 0 && bm_matches_alloc(&p->m[1]) == 0 &&
              cudaMalloc(&p->d_mask, BM_KP_CAP) == cudaSuccess && cudaMalloc(&p->d_res, sizeof(BmRansacResult)) == cudaSuccess &&
              cudaHostAlloc(&p->h_rb, sizeof(BmHostReadback), cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess;
    if (ok) memset(p->h_rb, 0, sizeof(BmHostReadback));
    p->is_orb = cfg.detector == BM_DET_ORB;
    { const char* e = getenv("BM_PROFILE"); p->prof = e && e[0] != '0'; }
    if (p->prof) { for (int i = 0; i < BM_NDET; ++i) { cudaEventCreate(&p->pd0[i]); cudaEventCreate(&p->pd1[i]); } cudaEventCreate(&p->pe0); cudaEventCreate(&p->pe1); }
    ok = ok && cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < BM_NDET && ok; ++i) {
        ok = bm_stream_create(&p->s_det[i], 1) == cudaSuccess;
        if (!ok) break;
        if (p->is_orb) ok = bm_orb_create(&p->orb[i], cfg.frame_h, cfg.frame_w, cfg.nfeatures, p->s_det[i]) == 0;
        else ok = bm_sift_create(&p->sift[i], cfg.frame_h, cfg.frame_w, cfg.nfeatures, p->s_det[i]) == 0;
    }
    if (!ok) { bm_set_error("bm_pipeline_create: allocation failed: %s", cudaGetErrorString(cudaGetLastError())); bm_pipeline_destroy(p); return BM_ERR_CUDA; }
    *out = p;
    return BM_OK;
}

void bm_pipeline_destroy(BmPipeline* p) {
    if (!p) return;
    if (p->s_est) cudaStreamSynchronize(p->s_est);
    if (p->stream) cudaStreamSynchronize(p->stream);
    if (p->prof && p->pd_n > 0)
        fprintf(stderr, "[bm profile] detect graph: %.1f us avg over %ld (in pipeline), match + RANSAC: %.1f us avg over %ld; host waited %.1f us per frame for the result\n",
                1e3 * p->pd_ms / p->pd_n, p->pd_n, p->pe_n ? 1e3 * p->pe_ms / p->pe_n : 0.0, p->pe_n, p->wait_n ? p->wait_us / p->wait_n : 0.0);
    for (int i = 0; i < BM_NDET; ++i) {
        if (p->s_det[i]) cudaStreamSynchronize(p->s_det[i]);
        bm_orb_destroy(p->orb[i]); bm_sift_destroy(p->sift[i]);
        if (p->s_det[i]) cudaStreamDestroy(p->s_det[i]);
    }
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    for (int i = 0; i < BM_KP_SLOTS; ++i) bm_kp_free(&p->kp[i]);
    if (p->ev_done) cudaEventDestroy(p->ev_done);
    if (p->ev_est) cudaEventDestroy(p->ev_est);
    for (int i = 0; i < BM_KP_SLOTS; ++i) if (p->ev_det[i]) cudaEventDestroy(p->ev_det[i]);
    if (p->s_est) cudaStreamDestroy(p->s_est);
    bm_matches_free(&p->m[0]); bm_matches_free(&p->m[1]);
    cudaFree(p->d_mask); cudaFree(p->d_res); cudaFreeHost(p->h_rb);
    delete p;
}

static cudaError_t detect(BmPipeline* p, const uint8_t* d_gray, BmKeypoints* out) {
    BM_NVTX("bm:detectAndCompute");
    const int slot = (int)(out - p->kp);
    const int i = p->det_next;
    p->det_next = (i + 1) % BM_NDET; p->det_last = i;
    cudaError_t e = cudaEventRecord(p->ev_fork, p->stream);                  // everything the caller ordered on `stream` so far
    if (e == cudaSuccess) e = cudaStreamWaitEvent(p->s_det[i], p->ev_fork, 0);
    if (e != cudaSuccess) return e;
    if (p->prof) {
        if (p->pd_pending[i] && cudaEventSynchronize(p->pd1[i]) == cudaSuccess) { float ms = 0.f; cudaEventElapsedTime(&ms, p->pd0[i], p->pd1[i]); p->pd_ms += ms; p->pd_n++; }
        cudaEventRecord(p->pd0[i], p->s_det[i]);
    }
    e = p->is_orb ? bm_orb_detect(p->orb[i], d_gray, out) : bm_sift_detect(p->sift[i], d_gray, out);
    if (e != cudaSuccess) return e;
    if (p->prof) { cudaEventRecord(p->pd1[i], p->s_det[i]); p->pd_pending[i] = true; }
    p->last_det_slot = slot;
    return cudaEventRecord(p->ev_det[slot], p->s_det[i]);
}

bm_status bm_pipeline_first_frame(BmPipeline* p, const uint8_t* d_gray) {
    p->prev = 0; p->cur = 1;
    for (int i = 0; i < BM_AHEAD_MAX; ++i) p->ahead_gray[i] = nullptr;
    BM_CUDA_OK(detect(p, d_gray, &p->kp[0]));
    p->have_prev = true;
    return BM_OK;
}

// a keypoint slot that holds neither the previous frame's features, nor the current frame's (while `cur_active`), nor features detected ahead
static int free_kp_slot(const BmPipeline* p, bool cur_active) {
    for (int s = 0; s < BM_KP_SLOTS; ++s) {
        if (s == p->prev || (cur_active && s == p->cur)) continue;
        bool taken = false;
        for (int i = 0; i < BM_AHEAD_MAX; ++i) taken |= p->ahead_gray[i] && p->ahead_slot[i] == s;
        if (!taken) return s;
    }
    return -1;
}

bm_status bm_pipeline_estimate_begin(BmPipeline* p, const uint8_t* d_gray) {
    if (!p->have_prev) { bm_set_error("process_frame before first frame"); return BM_ERR_ARG; }
    cudaStream_t s = p->s_est;
    int hit = -1;
    for (int i = 0; i < BM_AHEAD_MAX; ++i) if (p->ahead_gray[i] == d_gray && p->ahead_slot[i] != p->prev) hit = i;
    if (hit >= 0) {                                        // features already enqueued (detect_ahead)
        p->cur = p->ahead_slot[hit];
        p->ahead_gray[hit] = nullptr;
    } else {
        p->cur = free_kp_slot(p, false);
        if (p->cur < 0) { for (int i = 0; i < BM_AHEAD_MAX; ++i) p->ahead_gray[i] = nullptr; p->cur = free_kp_slot(p, false); }
        // the slot may still be read by the match of a frame the caller abandoned (a detect-ahead never needs this: its slot is
        // neither operand of the match in flight, and everything older has been waited for by the host)
        BM_CUDA_OK(cudaStreamWaitEvent(p->stream, p->ev_est, 0));
        BM_CUDA_OK(detect(p, d_gray, &p->kp[p->cur]));
    }
    BmKeypoints& cur = p->kp[p->cur];
    BmKeypoints& prev = p->kp[p->prev];
    BM_CUDA_OK(cudaStreamWaitEvent(s, p->ev_det[p->cur], 0));
    BM_CUDA_OK(cudaStreamWaitEvent(s, p->ev_det[p->prev], 0));
    p->mcur ^= 1;
    BmMatches& mm = p->m[p->mcur];
    if (p->prof) {
        if (p->pe_pending && cudaEventSynchronize(p->pe1) == cudaSuccess) { float ms = 0.f; cudaEventElapsedTime(&ms, p->pe0, p->pe1); p->pe_ms += ms; p->pe_n++; }
        cudaEventRecord(p->pe0, s);
    }
    if (p->is_orb) BM_CUDA_OK(bm_match_hamming(cur, prev, mm, s));
    else BM_CUDA_OK(bm_match_l2_ratio(cur, prev, mm, 0.7, s));                           // main.py:691
    BM_CUDA_OK(bm_launch_ransac(mm.src, mm.dst, mm.count, 2.0, 2000, 0.995, p->d_mask, p->d_res, s));   // main.py:857
    BM_COUNT_LAUNCHES(1);
    BM_CUDA_OK(bm_launch_pdl(k_publish_readback, dim3(1), dim3(32), 0, s, (const BmRansacResult*)p->d_res, (const int*)cur.count, (const int*)prev.count,
                             (const int*)mm.count, (const int*)cur.flags, (const int*)prev.flags, p->h_rb, ++p->seq));
    if (p->prof) { cudaEventRecord(p->pe1, s); p->pe_pending = true; }
    BM_CUDA_OK(cudaEventRecord(p->ev_done, s));
    BM_CUDA_OK(cudaEventRecord(p->ev_est, s));
    return BM_OK;
}

// detectAndCompute of the NEXT frame, enqueued behind the current frame's RANSAC before the host has read its result: whichever way
// the skip / accept decision goes (main.py:722-731), the features of the next frame are needed and depend on nothing else.  They go
// to the third keypoint slot; the following estimate_begin with the same gray buffer only adds match + RANSAC.
bm_status bm_pipeline_detect_ahead(BmPipeline* p, const uint8_t* d_gray, int* done) {
    if (done) *done = 0;
    if (!p->have_prev) return BM_OK;
    int e = -1;
    for (int i = 0; i < BM_AHEAD_MAX; ++i) {
        if (p->ahead_gray[i] == d_gray) { if (done) *done = 1; return BM_OK; }
        if (e < 0 && p->ahead_gray[i] == nullptr) e = i;
    }
    const int slot = free_kp_slot(p, true);
    if (e < 0 || slot < 0) return BM_OK;                   // BM_AHEAD_MAX frames are already detected ahead
    // (the free slot is neither operand of the match in flight, and every older match has been waited for by the host)
    BM_CUDA_OK(detect(p, d_gray, &p->kp[slot]));
    p->ahead_gray[e] = d_gray; p->ahead_slot[e] = slot;
    if (done) *done = 1;
    return BM_OK;
}
// the buffer is about to be overwritten: features detected ahead from it no longer describe its contents
void bm_pipeline_drop_ahead(BmPipeline* p, const uint8_t* d_gray) {
    if (!p) return;
    for (int i = 0; i < BM_AHEAD_MAX; ++i) if (p->ahead_gray[i] == d_gray) p->ahead_gray[i] = nullptr;
}

bm_status bm_pipeline_estimate_end(BmPipeline* p, bm_frame_info* info, double H_rel[9], int* have_h) {
    BM_NVTX("bm:wait match+RANSAC");
    BmHostReadback* rb = p->h_rb;
    timespec tw0; if (p->prof) clock_gettime(CLOCK_MONOTONIC, &tw0);
    {   // spin on the flag k_publish_readback stores last; the event is only consulted now and then, to surface a device error
        const volatile unsigned* flag = &rb->seq;
        for (unsigned spins = 1; *flag != p->seq; ++spins) {
            if ((spins & 0xfffu) == 0) {
                const cudaError_t q = cudaEventQuery(p->ev_done);
                if (q == cudaSuccess) { if (*flag != p->seq) { bm_set_error("match + RANSAC finished without publishing its result"); return BM_ERR_CUDA; } break; }
                if (q != cudaErrorNotReady) BM_CUDA_OK(q);
            }
#if defined(__x86_64__) || defined(__i386__)
            __builtin_ia32_pause();
#endif
        }
        __atomic_thread_fence(__ATOMIC_ACQUIRE);
    }
    if (p->prof) { timespec tw1; clock_gettime(CLOCK_MONOTONIC, &tw1); p->wait_us += 1e6 * (tw1.tv_sec - tw0.tv_sec) + 1e-3 * (tw1.tv_nsec - tw0.tv_nsec); p->wait_n++; }
    p->mdone = p->mcur;
    if (rb->overflow_cur || rb->overflow_prev) {
        // a detector list ran out of capacity: which candidates were kept depends on atomic order, the features are not cv2's
        bm_set_error("%s detector: candidate / keypoint list capacity exceeded on the %s frame", p->is_orb ? "ORB" : "SIFT", rb->overflow_cur ? "current" : "previous");
        return BM_ERR_UNSUPPORTED;
    }
    info->n_kp_cur = rb->n_cur; info->n_kp_prev = rb->n_prev; info->n_matches = rb->n_matches;
    info->ransac_iters = rb->r.iters; info->n_inliers = rb->r.n_inliers;
    *have_h = rb->r.ok == 1;
    if (*have_h) memcpy(H_rel, rb->r.H, 72);
    return BM_OK;
}

bm_status bm_pipeline_estimate(BmPipeline* p, const uint8_t* d_gray, bm_frame_info* info, double H_rel[9], int* have_h) {
    bm_status st = bm_pipeline_estimate_begin(p, d_gray);
    if (st != BM_OK) return st;
    return bm_pipeline_estimate_end(p, info, H_rel, have_h);
}

void bm_pipeline_advance(BmPipeline* p) { p->prev = p->cur; }
cudaError_t bm_pipeline_sync_est(BmPipeline* p) {       // everything the pipeline has queued on its own streams
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < BM_NDET && e == cudaSuccess; ++i) e = cudaStreamSynchronize(p->s_det[i]);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->s_est);
    return e;
}
cudaEvent_t bm_pipeline_last_detect_event(BmPipeline* p) { return p->ev_det[p->last_det_slot]; }
// record `ev` behind the most recently queued detect, on that detect's own stream (the caller owns the event: unlike ev_det[slot] it is
// not re-recorded when the keypoint slot is reused, so waiting on it never picks up a LATER detect)
cudaError_t bm_pipeline_record_after_last_detect(BmPipeline* p, cudaEvent_t ev) { return cudaEventRecord(ev, p->s_det[p->det_last]); }

// Capture the detector graph of every (detector instance, gray buffer, keypoint slot) combination now, so that no capture /
// instantiation (milliseconds for the ~90-node SIFT graph) lands in the first frames of a stream.  Nothing is executed.
bm_status bm_pipeline_warm_up(BmPipeline* p, const uint8_t* const* d_gray, int n_gray) {
    for (int i = 0; i < BM_NDET; ++i)
        for (int g = 0; g < n_gray; ++g)
            for (int k = 0; k < BM_KP_SLOTS; ++k)
                BM_CUDA_OK(p->is_orb ? bm_orb_detect(p->orb[i], d_gray[g], &p->kp[k], false) : bm_sift_detect(p->sift[i], d_gray[g], &p->kp[k], false));
    return BM_OK;
}

BmKeypoints* bm_pipeline_keypoints(BmPipeline* p, int which) { return &p->kp[which ? p->cur : p->prev]; }
BmMatches* bm_pipeline_matches(BmPipeline* p) { return &p->m[p->mdone]; }      // matches of the last frame that was waited for
