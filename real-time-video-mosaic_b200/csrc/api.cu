// api.cu -- C ABI of libb200mosaic.so (see include/b200mosaic.h) and the per-mosaic device state.
// Host-side control flow mirrors VideMosaic (/root/reference/main.py:17-112, 710-859); all pixel / feature work is
// CUDA.  There is deliberately no CPU fallback: every entry point fails with BM_ERR_CUDA if the device is missing.
#include "../../include/b200mosaic.h"
#include "common.cuh"
#include "ingest.cuh"
#include "warp_blend.cuh"
#include "pipeline.cuh"
#include "finalize.cuh"
#include "preview.cuh"
#include "jpeg.cuh"
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
#include <new>

static thread_local char g_err[512] = "";
bool bm_pdl_enabled() {
    static const bool on = [] { const char* e = getenv("BM_NO_PDL"); return !(e && e[0] == '1'); }();
    return on;
}

void bm_set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
extern "C" const char* bm_last_error(void) { return g_err; }
extern "C" int bm_version(void) { return 100; }
long long g_bm_launches = 0;
thread_local long long* t_bm_launch_sink = nullptr;
extern "C" long long bm_kernel_launches(void) { return __atomic_load_n(&g_bm_launches, __ATOMIC_RELAXED); }

#define BM_TRY(expr) do { bm_status _s = (expr); if (_s < 0) return _s; } while (0)

// ------------------------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------------------------
static void free_blend(BmBlendBufs& b) {
    cudaFree(b.canvas); cudaFree(b.wbuf); cudaFree(b.wno); cudaFree(b.flags);
    bm_dt_free_plane(&b.dt.p[0]); bm_dt_free_plane(&b.dt.p[1]);
    memset(&b, 0, sizeof(b));
}

static bm_status alloc_blend(BmBlendBufs& b, int canvas_h, int canvas_w, size_t scratch_px) {
    memset(&b, 0, sizeof(b));
    b.canvas_h = canvas_h; b.canvas_w = canvas_w;
    const size_t n = (size_t)canvas_h * canvas_w;
    if (scratch_px > n || scratch_px == 0) scratch_px = n;
    b.scratch_px = scratch_px;
    const size_t pad = (size_t)8 * canvas_h + 64;          // row strides of the scratch planes are padded to 8 / 4 columns
    BM_CUDA_OK(cudaMalloc(&b.canvas, n * sizeof(uchar4)));
    BM_CUDA_OK(bm_dt_alloc_plane(&b.dt.p[0], canvas_w, canvas_h, n));
    BM_CUDA_OK(bm_dt_alloc_plane(&b.dt.p[1], canvas_w, canvas_h, scratch_px));
    BM_CUDA_OK(cudaMalloc(&b.wbuf, (scratch_px + pad) * sizeof(uchar4)));
    BM_CUDA_OK(cudaMalloc(&b.wno, (scratch_px + pad) * sizeof(float2)));
    BM_CUDA_OK(cudaMalloc(&b.flags, 16 * sizeof(int)));
    BM_CUDA_OK(cudaMemset(b.canvas, 0, n * sizeof(uchar4)));
    BM_CUDA_OK(cudaMemset(b.flags, 0, 16 * sizeof(int)));
    return BM_OK;
}

// Three streams per handle:
//   stream   : detect / match / RANSAC of frame t (the host waits on it once per frame for the homography)
//   s_chain  : warp / blend chain of frame t -- runs concurrently with detect of frame t+1 (it only touches the canvas)
//   s_copy   : H2D + ingest of the next frame (bm_prefetch_frame) while the current one is being processed
// Frame slots rotate over three buffers; events order  upload(slot) -> detect / chain(slot) -> next upload(slot).  Three, not two:
// the upload of frame t+1 must not wait for the chain of frame t-1, which is still reading that frame's BGRX copy when the caller
// stages t+1 (with two slots the H2D copy of every frame started a whole chain late and sat on the critical path of the e2e rate).
#define BM_LOOKAHEAD BM_AHEAD_MAX       // frames that may be staged (and detected) ahead of the current one
#define BM_SLOTS (BM_LOOKAHEAD + 2)
#define BM_CANVAS_STAGE_BYTES ((size_t)4 << 20)
#define BM_SLOT_NEXT(c) (((c) + 1) % BM_SLOTS)
#define BM_SLOT_PREV(c) (((c) + BM_SLOTS - 1) % BM_SLOTS)
struct bm_mosaic_s {
    bm_config cfg;
    cudaStream_t stream = nullptr, s_chain = nullptr, s_copy = nullptr;
    cudaEvent_t ev_up[BM_SLOTS] = {};          // upload + ingest of the slot finished
    cudaEvent_t ev_chain[BM_SLOTS] = {};       // last chain that read the slot's BGRX finished
    cudaEvent_t ev_spec[BM_SLOTS] = {};        // last detect-ahead that read the slot's gray plane finished (recorded on that detect's stream)
    bool spec_used[BM_SLOTS] = {};
    int overlap = 1;                                    // 0: detect waits for the previous chain (clean chain timing)
    // frames staged ahead by bm_prefetch_frame, in the order the caller will process them: q[0] is the next frame.  Their H2D copy +
    // ingest run on the copy stream, their detectAndCompute is queued at the start of the next _end ("detect-ahead"): with two frames
    // ahead three detects are in flight and the frame period is no longer tied to the detect latency.
    struct Staged { const uint8_t* ptr; int slot; bool detected; } q[BM_LOOKAHEAD] = {};
    int nq = 0;
    const uint8_t* begun = nullptr;                     // frame whose detect / match / RANSAC was already enqueued by the previous _end
    BmBlendBufs blend;
    // frame staging: double-buffered pinned host + device buffers
    uint8_t* h_stage[BM_SLOTS] = {};
    uint8_t* d_bgr[BM_SLOTS] = {};
    uchar4* d_bgrx[BM_SLOTS] = {};
    uint8_t* d_gray[BM_SLOTS] = {};
    cudaEvent_t ev_h2d[BM_SLOTS] = {};
    int cur = 0;
    // canvas export
    uint8_t* d_canvas_bgr = nullptr;
    uint32_t* d_ghost[2] = {nullptr, nullptr};             // row-tile mode: carries entering the canvas plane from the tile above / below, [3][ts]
    uint8_t* d_final = nullptr; size_t final_cap = 0;      // finalisation result (screen sized), allocated on first use
    int* d_bounds = nullptr;
    uint8_t* h_cstage[2] = {nullptr, nullptr};              // pinned staging of bm_get_canvas for pageable destinations
    cudaEvent_t ev_cstage[2] = {nullptr, nullptr};
    BmPreviewPlan preview;                                  // thumbnail tables / buffers, built on first use
    BmJpeg jpeg;                                            // scratch of the mosaic.jpg encoder, built on first use
    // stitcher state (main.py:92-102)
    double H_old[9];
    double history[5][9];
    int history_len = 0;
    int stabilization_enabled = 1, history_size = 5;
    double translation_threshold = 50.0, scale_threshold = 0.3;
    int w_offset = 0, h_offset = 0;
    BmPipeline* pipe = nullptr;       // detector / matcher / RANSAC state (pipeline.cu)
    // optional CUDA-event timing of the warp/blend chain
    int timing = 0;
    static const int kEvRing = 64;
    cudaEvent_t ev0[kEvRing] = {}, ev1[kEvRing] = {};
    int ev_pending = 0;
    double t_ms = 0.0, t_bytes = 0.0; int t_frames = 0;
};

static void cancel_early_begin(bm_mosaic_s* m);
extern "C" bm_status bm_destroy(bm_handle m);
static size_t frame_bytes(const bm_config& c) { return (size_t)c.frame_h * c.frame_w * 3; }

extern "C" bm_status bm_create(const bm_config* cfg, bm_handle* out) {
    if (!cfg || !out || cfg->frame_h <= 0 || cfg->frame_w <= 0 || cfg->canvas_h < cfg->frame_h || cfg->canvas_w < cfg->frame_w) {
        bm_set_error("bm_create: bad config"); return BM_ERR_ARG;
    }
    if (cfg->detector != BM_DET_SIFT && cfg->detector != BM_DET_ORB) { bm_set_error("bm_create: detector must be sift|orb"); return BM_ERR_ARG; }
    *out = nullptr;
    int ndev = 0;
    BM_CUDA_OK(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) { bm_set_error("bm_create: no CUDA device %d", cfg->device); return BM_ERR_CUDA; }
    BM_CUDA_OK(cudaSetDevice(cfg->device));
    // every failure below releases what was created so far (bm_destroy tolerates a half-built handle): OOM at creation is the
    // expected failure for config-5 sized canvases and many-stream runs
#define BM_CREATE_OK(expr)                                                                 \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            bm_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            bm_destroy(m);                                                                 \
            cudaGetLastError();                                                            \
            return BM_ERR_CUDA;                                                            \
        }                                                                                  \
    } while (0)
    bm_mosaic_s* m = new (std::nothrow) bm_mosaic_s();
    if (!m) return BM_ERR_ARG;
    m->cfg = *cfg;
    if (m->cfg.nfeatures <= 0) m->cfg.nfeatures = 700;
    BM_CREATE_OK(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    BM_CREATE_OK(cudaStreamCreateWithFlags(&m->s_chain, cudaStreamNonBlocking));
    BM_CREATE_OK(cudaStreamCreateWithFlags(&m->s_copy, cudaStreamNonBlocking));
    for (int i = 0; i < BM_SLOTS; ++i) {
        BM_CREATE_OK(cudaEventCreateWithFlags(&m->ev_up[i], cudaEventDisableTiming));
        BM_CREATE_OK(cudaEventCreateWithFlags(&m->ev_chain[i], cudaEventDisableTiming));
        BM_CREATE_OK(cudaEventCreateWithFlags(&m->ev_spec[i], cudaEventDisableTiming));
    }
    for (int k = 0; k < 2; ++k) {            // page-locking is slow (tens of ms): at creation, not in the first bm_get_canvas
        BM_CREATE_OK(cudaHostAlloc(&m->h_cstage[k], BM_CANVAS_STAGE_BYTES, cudaHostAllocDefault));
        BM_CREATE_OK(cudaEventCreateWithFlags(&m->ev_cstage[k], cudaEventDisableTiming));
    }
    // scratch: the window of a frame is at most the canvas; typical is frame-sized.  Size for the whole canvas when
    // it is small (<= 64 Mpx), otherwise for 4x the frame area plus margins (config 5: 32768^2 canvas, 4K frames).
    const size_t canvas_px = (size_t)cfg->canvas_h * cfg->canvas_w;
    size_t scratch = canvas_px;
    if (canvas_px > ((size_t)64 << 20)) scratch = (size_t)4 * (cfg->frame_h + 64) * (cfg->frame_w + 64);
    bm_status st = alloc_blend(m->blend, cfg->canvas_h, cfg->canvas_w, scratch);
    if (st != BM_OK) { bm_destroy(m); return st; }
    const size_t fb = frame_bytes(*cfg), fpx = (size_t)cfg->frame_h * cfg->frame_w;
    for (int i = 0; i < BM_SLOTS; ++i) {
        BM_CREATE_OK(cudaHostAlloc(&m->h_stage[i], fb, cudaHostAllocDefault));
        BM_CREATE_OK(cudaMalloc(&m->d_bgr[i], fb + 16));
        BM_CREATE_OK(cudaMalloc(&m->d_bgrx[i], fpx * sizeof(uchar4)));
        BM_CREATE_OK(cudaMalloc(&m->d_gray[i], fpx + 16));
        BM_CREATE_OK(cudaEventCreateWithFlags(&m->ev_h2d[i], cudaEventDisableTiming));
    }
    BM_CREATE_OK(cudaMalloc(&m->d_canvas_bgr, canvas_px * 3 + 16));
    for (int i = 0; i < 9; ++i) m->H_old[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int i = 0; i < bm_mosaic_s::kEvRing; ++i) { BM_CREATE_OK(cudaEventCreate(&m->ev0[i])); BM_CREATE_OK(cudaEventCreate(&m->ev1[i])); }
    st = bm_pipeline_create(&m->pipe, m->cfg, m->stream);
    if (st != BM_OK) { bm_destroy(m); return st; }
    *out = m;
    return BM_OK;
}

#undef BM_CREATE_OK

extern "C" bm_status bm_destroy(bm_handle m) {
    if (!m) return BM_OK;
    cudaSetDevice(m->cfg.device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    if (m->s_chain) cudaStreamSynchronize(m->s_chain);
    if (m->s_copy) cudaStreamSynchronize(m->s_copy);
    if (m->pipe) bm_pipeline_sync_est(m->pipe);
    bm_pipeline_destroy(m->pipe);
    free_blend(m->blend);
    for (int i = 0; i < BM_SLOTS; ++i) {
        cudaFreeHost(m->h_stage[i]); cudaFree(m->d_bgr[i]); cudaFree(m->d_bgrx[i]); cudaFree(m->d_gray[i]);
        if (m->ev_h2d[i]) cudaEventDestroy(m->ev_h2d[i]);
    }
    cudaFree(m->d_canvas_bgr); cudaFree(m->d_final); cudaFree(m->d_bounds); cudaFree(m->d_ghost[0]); cudaFree(m->d_ghost[1]);
    bm_preview_free(&m->preview);
    bm_jpeg_free(&m->jpeg);
    for (int k = 0; k < 2; ++k) { if (m->h_cstage[k]) cudaFreeHost(m->h_cstage[k]); if (m->ev_cstage[k]) cudaEventDestroy(m->ev_cstage[k]); }
    for (int i = 0; i < bm_mosaic_s::kEvRing; ++i) { if (m->ev0[i]) cudaEventDestroy(m->ev0[i]); if (m->ev1[i]) cudaEventDestroy(m->ev1[i]); }
    for (int i = 0; i < BM_SLOTS; ++i) { if (m->ev_up[i]) cudaEventDestroy(m->ev_up[i]); if (m->ev_chain[i]) cudaEventDestroy(m->ev_chain[i]); if (m->ev_spec[i]) cudaEventDestroy(m->ev_spec[i]); }
    if (m->stream) cudaStreamDestroy(m->stream);
    if (m->s_chain) cudaStreamDestroy(m->s_chain);
    if (m->s_copy) cudaStreamDestroy(m->s_copy);
    delete m;
    return BM_OK;
}

// host frame -> device (BGR packed, BGRX, gray).  Pinned sources are copied directly; pageable ones are staged.
static bm_status upload(bm_mosaic_s* m, const uint8_t* h_bgr, size_t stride, int slot) {
    BM_NVTX("bm:upload+ingest");
    const int fh = m->cfg.frame_h, fw = m->cfg.frame_w;
    const size_t rowb = (size_t)fw * 3, fb = rowb * fh;
    if (stride == 0) stride = rowb;
    const uint8_t* src = h_bgr;
    cudaPointerAttributes attr;
    bool pinned = false;
    if (stride == rowb && cudaPointerGetAttributes(&attr, h_bgr) == cudaSuccess) pinned = (attr.type == cudaMemoryTypeHost);
    else cudaGetLastError();
    if (!pinned) {
        // the staging slot may still be in flight from two frames ago
        BM_CUDA_OK(cudaEventSynchronize(m->ev_h2d[slot]));
        if (stride == rowb) memcpy(m->h_stage[slot], h_bgr, fb);
        else for (int y = 0; y < fh; ++y) memcpy(m->h_stage[slot] + (size_t)y * rowb, h_bgr + (size_t)y * stride, rowb);
        src = m->h_stage[slot];
    }
    // the slot's previous tenant: its detect finished (the host waited for it), its chain may still be reading the BGRX copy
    // (a detect-ahead of a frame the caller did not continue with may also still be reading the gray plane)
    BM_CUDA_OK(cudaStreamWaitEvent(m->s_copy, m->ev_chain[slot], 0));
    if (m->spec_used[slot]) BM_CUDA_OK(cudaStreamWaitEvent(m->s_copy, m->ev_spec[slot], 0));
    bm_pipeline_drop_ahead(m->pipe, m->d_gray[slot]);
    BM_CUDA_OK(cudaMemcpyAsync(m->d_bgr[slot], src, fb, cudaMemcpyHostToDevice, m->s_copy));
    BM_CUDA_OK(cudaEventRecord(m->ev_h2d[slot], m->s_copy));
    BM_CUDA_OK(bm_launch_ingest(m->d_bgr[slot], fh, fw, m->d_gray[slot], m->d_bgrx[slot], m->s_copy));
    BM_CUDA_OK(cudaEventRecord(m->ev_up[slot], m->s_copy));
    return BM_OK;
}

static void q_clear(bm_mosaic_s* m) {
    for (int i = 0; i < m->nq; ++i) bm_pipeline_drop_ahead(m->pipe, m->d_gray[m->q[i].slot]);
    m->nq = 0;
}
static int q_find(const bm_mosaic_s* m, const uint8_t* ptr) {
    for (int i = 0; i < m->nq; ++i) if (m->q[i].ptr == ptr) return i;
    return -1;
}
static void q_pop_front(bm_mosaic_s* m) {
    for (int i = 1; i < m->nq; ++i) m->q[i - 1] = m->q[i];
    m->nq--;
}

// device frame (packed BGR) -> slot, ingest on the copy stream
static bm_status ingest_device(bm_mosaic_s* m, const uint8_t* d_bgr, int slot, cudaStream_t s) {
    BM_CUDA_OK(cudaStreamWaitEvent(s, m->ev_chain[slot], 0));              // the slot's previous chain still reads its BGRX copy
    if (m->spec_used[slot]) BM_CUDA_OK(cudaStreamWaitEvent(s, m->ev_spec[slot], 0));   // ... or an abandoned detect-ahead its gray plane
    bm_pipeline_drop_ahead(m->pipe, m->d_gray[slot]);
    BM_CUDA_OK(bm_launch_ingest(d_bgr, m->cfg.frame_h, m->cfg.frame_w, m->d_gray[slot], m->d_bgrx[slot], s));
    BM_CUDA_OK(cudaEventRecord(m->ev_up[slot], s));
    return BM_OK;
}

// Make `ptr` the current frame (m->cur): the staged copy if it is the next staged frame, otherwise every staged frame is stale (the
// caller continued with another frame) and the frame is uploaded / ingested now.  `consumer` then waits for the slot's upload.
static bm_status take_frame(bm_mosaic_s* m, const uint8_t* ptr, size_t stride, bool device, cudaStream_t consumer) {
    if (m->nq > 0 && m->q[0].ptr == ptr) {
        m->cur = m->q[0].slot;
        q_pop_front(m);
    } else {
        q_clear(m);
        m->cur = BM_SLOT_NEXT(m->cur);
        if (device) BM_TRY(ingest_device(m, ptr, m->cur, m->s_copy));
        else BM_TRY(upload(m, ptr, stride, m->cur));
    }
    BM_CUDA_OK(cudaStreamWaitEvent(consumer, m->ev_up[m->cur], 0));
    return BM_OK;
}

static bm_status prefetch(bm_mosaic_s* m, const uint8_t* ptr, size_t stride, bool device) {
    if (m->begun == ptr || q_find(m, ptr) >= 0 || m->nq >= BM_LOOKAHEAD) return BM_OK;     // already staged / no room: nothing to do
    const int slot = (m->cur + 1 + m->nq) % BM_SLOTS;
    if (device) BM_TRY(ingest_device(m, ptr, slot, m->s_copy));
    else BM_TRY(upload(m, ptr, stride, slot));
    m->q[m->nq].ptr = ptr; m->q[m->nq].slot = slot; m->q[m->nq].detected = false;
    m->nq++;
    return BM_OK;
}

extern "C" bm_status bm_prefetch_frame(bm_handle m, const uint8_t* h_bgr, size_t stride) {
    if (!m || !h_bgr) { bm_set_error("bm_prefetch_frame: null"); return BM_ERR_ARG; }
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    return prefetch(m, h_bgr, stride, false);
}

// same for a frame that already lives in device memory (packed BGR): ingest into the free slot on the copy stream
extern "C" bm_status bm_prefetch_frame_device(bm_handle m, const uint8_t* d_bgr) {
    if (!m || !d_bgr) { bm_set_error("bm_prefetch_frame_device: null"); return BM_ERR_ARG; }
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    return prefetch(m, d_bgr, 0, true);
}

extern "C" bm_status bm_set_overlap(bm_handle m, int on) { if (!m) return BM_ERR_ARG; m->overlap = on ? 1 : 0; return BM_OK; }

extern "C" bm_status bm_first_frame(bm_handle m, const uint8_t* h_bgr, size_t stride) {
    if (!m || !h_bgr) { bm_set_error("bm_first_frame: null"); return BM_ERR_ARG; }
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    const int fh = m->cfg.frame_h, fw = m->cfg.frame_w, ch = m->cfg.canvas_h, cw = m->cfg.canvas_w;
    m->cur = 0; m->nq = 0; m->begun = nullptr;
    BM_TRY(upload(m, h_bgr, stride, 0));
    BM_CUDA_OK(cudaStreamWaitEvent(m->stream, m->ev_up[0], 0));
    BM_CUDA_OK(cudaStreamWaitEvent(m->s_chain, m->ev_up[0], 0));
    // main.py:86-87 (the reference's names are swapped: w_offset is the ROW offset)
    m->w_offset = (int)((double)ch / 1 - (double)fh / 1);
    m->h_offset = (int)((double)cw / 2 - (double)fw / 2);
    BM_CUDA_OK(cudaMemsetAsync(m->blend.canvas, 0, (size_t)ch * cw * sizeof(uchar4), m->s_chain));
    BM_CUDA_OK(bm_launch_paste(m->blend.canvas, cw, m->d_bgrx[0], fw, fh, m->h_offset, m->w_offset, m->s_chain));   // main.py:89-90
    BM_CUDA_OK(bm_launch_full_rowscan(m->blend, m->s_chain));
    BM_CUDA_OK(cudaEventRecord(m->ev_chain[0], m->s_chain));
    for (int i = 0; i < 9; ++i) m->H_old[i] = (i % 4 == 0) ? 1.0 : 0.0;      // main.py:92-94
    m->H_old[2] = m->h_offset; m->H_old[5] = m->w_offset;
    m->history_len = 0;
    BM_TRY(bm_pipeline_first_frame(m->pipe, m->d_gray[0]));                  // main.py:104-112
    BM_CUDA_OK(cudaStreamSynchronize(m->stream));
    BM_CUDA_OK(bm_pipeline_sync_est(m->pipe));
    BM_CUDA_OK(cudaStreamSynchronize(m->s_chain));
    return BM_OK;
}

static void fill_info_plan(bm_frame_info* info, const BmFramePlan& p) {
    if (!info) return;
    info->win[0] = p.win.x0; info->win[1] = p.win.y0; info->win[2] = p.win.x1; info->win[3] = p.win.y1;
}

static bm_status timing_drain(bm_mosaic_s* m) {
    for (int i = 0; i < m->ev_pending; ++i) {
        float ms = 0.f;
        BM_CUDA_OK(cudaEventSynchronize(m->ev1[i]));
        BM_CUDA_OK(cudaEventElapsedTime(&ms, m->ev0[i], m->ev1[i]));
        m->t_ms += ms;
    }
    m->ev_pending = 0;
    return BM_OK;
}

extern "C" bm_status bm_timing_enable(bm_handle m, int on) { if (!m) return BM_ERR_ARG; m->timing = on; return BM_OK; }
extern "C" bm_status bm_timing_read(bm_handle m, double* ms, double* bytes, int* frames, int reset) {
    if (!m) return BM_ERR_ARG;
    BM_TRY(timing_drain(m));
    if (ms) *ms = m->t_ms;
    if (bytes) *bytes = m->t_bytes;
    if (frames) *frames = m->t_frames;
    if (reset) { m->t_ms = 0; m->t_bytes = 0; m->t_frames = 0; }
    return BM_OK;
}

static bm_status warp_device(bm_mosaic_s* m, const uchar4* d_bgrx, const double H[9], bm_frame_info* info, bool want_flag, int slot = -1) {
    BM_NVTX("bm:warp+blend chain");
    BmFramePlan plan;
    bm_make_plan(H, m->cfg.frame_w, m->cfg.frame_h, m->cfg.canvas_w, m->cfg.canvas_h, &plan);
    const size_t need = (size_t)bm_win_w(plan.reg) * bm_win_h(plan.reg);
    if (plan.valid && need > m->blend.scratch_px) { bm_set_error("warp window %zu px exceeds scratch %zu px", need, m->blend.scratch_px); return BM_ERR_UNSUPPORTED; }
    if (m->timing && m->ev_pending == bm_mosaic_s::kEvRing) BM_TRY(timing_drain(m));
    if (m->timing) BM_CUDA_OK(cudaEventRecord(m->ev0[m->ev_pending], m->s_chain));
    BM_CUDA_OK(bm_launch_warp_blend(m->blend, d_bgrx, plan, m->s_chain));
    if (slot >= 0) BM_CUDA_OK(cudaEventRecord(m->ev_chain[slot], m->s_chain));
    if (m->timing) {
        BM_CUDA_OK(cudaEventRecord(m->ev1[m->ev_pending], m->s_chain));
        m->ev_pending++;
        const double N = (double)m->cfg.frame_w * m->cfg.frame_h, A = plan.valid ? (double)bm_win_w(plan.win) * bm_win_h(plan.win) : 0.0;
        m->t_bytes += 3.0 * N + 6.0 * A; m->t_frames++;
    }
    fill_info_plan(info, plan);
    if (info && want_flag) {
        int f = 0;
        if (plan.valid) BM_CUDA_OK(cudaMemcpyAsync(&f, m->blend.flags, sizeof(int), cudaMemcpyDeviceToHost, m->s_chain));
        BM_CUDA_OK(cudaStreamSynchronize(m->s_chain));
        info->any_overlap = f;
    }
    return BM_OK;
}

extern "C" bm_status bm_warp_frame(bm_handle m, const uint8_t* h_bgr, size_t stride, const double H[9], bm_frame_info* info) {
    if (!m || !h_bgr || !H) { bm_set_error("bm_warp_frame: null"); return BM_ERR_ARG; }
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    cancel_early_begin(m);
    BM_TRY(take_frame(m, h_bgr, stride, false, m->s_chain));
    if (info) { memset(info, 0, sizeof(*info)); memcpy(info->H, H, 9 * sizeof(double)); }
    return warp_device(m, m->d_bgrx[m->cur], H, info, true, m->cur);
}

extern "C" bm_status bm_warp_frame_async(bm_handle m, const uint8_t* h_bgr, size_t stride, const double H[9]) {
    if (!m || !h_bgr || !H) { bm_set_error("bm_warp_frame_async: null"); return BM_ERR_ARG; }
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    cancel_early_begin(m);
    BM_TRY(take_frame(m, h_bgr, stride, false, m->s_chain));
    return warp_device(m, m->d_bgrx[m->cur], H, nullptr, false, m->cur);
}

extern "C" bm_status bm_warp_frame_device(bm_handle m, const uint8_t* d_bgrx, const double H[9], bm_frame_info* info) {
    if (!m || !d_bgrx || !H) { bm_set_error("bm_warp_frame_device: null"); return BM_ERR_ARG; }
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    return warp_device(m, reinterpret_cast<const uchar4*>(d_bgrx), H, info, false);
}

extern "C" bm_status bm_upload_frame(bm_handle m, const uint8_t* h_bgr, size_t stride, const uint8_t** d_out) {
    if (!m || !h_bgr) return BM_ERR_ARG;
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    cancel_early_begin(m);
    BM_TRY(take_frame(m, h_bgr, stride, false, m->s_chain));
    if (d_out) *d_out = reinterpret_cast<const uint8_t*>(m->d_bgrx[m->cur]);
    return BM_OK;
}

// Optional: build every CUDA graph the per-frame path will replay (detector instance x frame slot x keypoint slot) up front, so a
// real-time caller sees no capture / instantiation hiccup in its first frames.  Executes nothing; results are unaffected.
extern "C" bm_status bm_warm_up(bm_handle m) {
    if (!m) return BM_ERR_ARG;
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    const uint8_t* grays[BM_SLOTS];
    for (int i = 0; i < BM_SLOTS; ++i) grays[i] = m->d_gray[i];
    return bm_pipeline_warm_up(m->pipe, grays, BM_SLOTS);
}

extern "C" bm_status bm_sync(bm_handle m) {
    if (!m) return BM_ERR_ARG;
    BM_CUDA_OK(cudaStreamSynchronize(m->s_copy));
    BM_CUDA_OK(cudaStreamSynchronize(m->stream));
    if (m->pipe) BM_CUDA_OK(bm_pipeline_sync_est(m->pipe));
    BM_CUDA_OK(cudaStreamSynchronize(m->s_chain));
    return BM_OK;
}
extern "C" void* bm_stream(bm_handle m) { return m ? (void*)m->stream : nullptr; }

extern "C" bm_status bm_get_canvas(bm_handle m, uint8_t* h_out) {
    if (!m || !h_out) return BM_ERR_ARG;
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    const size_t n = (size_t)m->cfg.canvas_h * m->cfg.canvas_w;
    BM_CUDA_OK(bm_launch_unpack_canvas(m->blend.canvas, m->d_canvas_bgr, (int)n, m->s_chain));
    const size_t bytes = n * 3;
    cudaPointerAttributes attr;
    bool pinned = false;
    if (cudaPointerGetAttributes(&attr, h_out) == cudaSuccess) pinned = attr.type == cudaMemoryTypeHost;
    else cudaGetLastError();
    if (pinned) {
        BM_CUDA_OK(cudaMemcpyAsync(h_out, m->d_canvas_bgr, bytes, cudaMemcpyDeviceToHost, m->s_chain));
        BM_CUDA_OK(cudaStreamSynchronize(m->s_chain));
        return BM_OK;
    }
    // pageable destination (a NumPy array): the driver's own staging runs at a few GB/s; DMA into two pinned 4 MB buffers
    // instead and copy chunk i to the caller while chunk i + 1 is in flight
    const size_t CH = BM_CANVAS_STAGE_BYTES;
    size_t prev_off = 0, prev_n = 0;
    int idx = 0;
    for (size_t off = 0; off < bytes; off += CH, ++idx) {
        const int k = idx & 1;
        const size_t cn = bytes - off < CH ? bytes - off : CH;
        BM_CUDA_OK(cudaMemcpyAsync(m->h_cstage[k], m->d_canvas_bgr + off, cn, cudaMemcpyDeviceToHost, m->s_chain));
        BM_CUDA_OK(cudaEventRecord(m->ev_cstage[k], m->s_chain));
        if (prev_n) {                                    // the other buffer: its DMA was issued one iteration ago
            BM_CUDA_OK(cudaEventSynchronize(m->ev_cstage[k ^ 1]));
            memcpy(h_out + prev_off, m->h_cstage[k ^ 1], prev_n);
        }
        prev_off = off; prev_n = cn;
    }
    if (prev_n) {
        BM_CUDA_OK(cudaEventSynchronize(m->ev_cstage[(idx - 1) & 1]));
        memcpy(h_out + prev_off, m->h_cstage[(idx - 1) & 1], prev_n);
    }
    return BM_OK;
}

// crop_black_areas(output_img, threshold, margin) + scale_to_screen(cropped, target_w, target_h) (main.py:980-1038, as called
// at :1647-1659) on the device canvas.  h_out == NULL: only the sizes are computed (out_wh, rect), so that the caller can allocate.
// crop_black_areas + scale_to_screen of the live canvas into m->d_final (packed BGR, out_wh[0] x out_wh[1]); run = false only sizes it
static bm_status bm_finalize_device(bm_handle m, int threshold, int margin, int target_w, int target_h, bool run, int out_wh[2], int rect[4]) {
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    const int cw = m->cfg.canvas_w, ch = m->cfg.canvas_h;
    if (!m->d_bounds) BM_CUDA_OK(cudaMalloc(&m->d_bounds, 4 * sizeof(int)));
    int b[4];
    BM_CUDA_OK(bm_launch_crop_bounds(m->blend.canvas, cw, ch, threshold, m->d_bounds, m->s_chain));
    BM_CUDA_OK(cudaMemcpyAsync(b, m->d_bounds, sizeof(b), cudaMemcpyDeviceToHost, m->s_chain));
    BM_CUDA_OK(cudaStreamSynchronize(m->s_chain));
    int x = 0, y = 0, w = cw, h = ch;
    if (b[2] >= 0) {                                         // main.py:996-1003 (coords is None -> the image itself)
        const int bx = b[0], by = b[1], bw = b[2] - b[0] + 1, bh = b[3] - b[1] + 1;
        x = bx + margin > 0 ? bx + margin : 0;
        y = by + margin > 0 ? by + margin : 0;
        w = cw - x < bw - 2 * margin ? cw - x : bw - 2 * margin;
        h = ch - y < bh - 2 * margin ? ch - y : bh - 2 * margin;
    }
    if (rect) { rect[0] = x; rect[1] = y; rect[2] = w; rect[3] = h; }
    if (w <= 0 || h <= 0) { bm_set_error("bm_finalize: the crop is empty (%d x %d)", w, h); return BM_ERR_UNSUPPORTED; }
    const double sw_ = target_w > 0 && target_h > 0 ? target_w : 1920, sh_ = target_w > 0 && target_h > 0 ? target_h : 1080;   // main.py:1013-1024 off Windows
    double scale = sw_ / (double)w < sh_ / (double)h ? sw_ / (double)w : sh_ / (double)h;
    if (scale <= 0) scale = 1.0;
    int nw = (int)(w * scale), nh = (int)(h * scale);
    if (nw < 1) nw = 1;
    if (nh < 1) nh = 1;
    out_wh[0] = nw; out_wh[1] = nh;
    if (!run) return BM_OK;
    const size_t need = (size_t)nw * nh * 3;
    if (m->final_cap < need) {
        cudaFree(m->d_final); m->d_final = nullptr; m->final_cap = 0;
        BM_CUDA_OK(cudaMalloc(&m->d_final, need + 16));
        m->final_cap = need;
    }
    BM_CUDA_OK(bm_launch_resize_linear(m->blend.canvas, cw, x, y, w, h, m->d_final, nw, nh, m->s_chain));
    return BM_OK;
}

extern "C" bm_status bm_finalize(bm_handle m, int threshold, int margin, int target_w, int target_h, uint8_t* h_out, size_t cap_bytes,
                                 int out_wh[2], int rect[4]) {
    if (!m || !out_wh) return BM_ERR_ARG;
    if (!h_out) return bm_finalize_device(m, threshold, margin, target_w, target_h, false, out_wh, rect);
    bm_status st = bm_finalize_device(m, threshold, margin, target_w, target_h, false, out_wh, rect);
    if (st != BM_OK) return st;
    const size_t need = (size_t)out_wh[0] * out_wh[1] * 3;
    if (cap_bytes < need) { bm_set_error("bm_finalize: output buffer too small (%zu < %zu)", cap_bytes, need); return BM_ERR_ARG; }
    st = bm_finalize_device(m, threshold, margin, target_w, target_h, true, out_wh, rect);
    if (st != BM_OK) return st;
    BM_CUDA_OK(cudaMemcpyAsync(h_out, m->d_final, need, cudaMemcpyDeviceToHost, m->s_chain));
    BM_CUDA_OK(cudaStreamSynchronize(m->s_chain));
    return BM_OK;
}

// header + stuffed scan (already in j->out) + EOI -> host
static bm_status bm_jpeg_assemble(BmJpeg* j, int w, int h, int quality, size_t scan_bytes, uint8_t* h_out, size_t cap, size_t* nbytes, cudaStream_t s) {
    uint8_t hdr[BM_JPEG_HEADER_MAX];
    const size_t nh = bm_jpeg_write_header(hdr, w, h, quality), total = nh + scan_bytes + 2;
    *nbytes = total;
    if (cap < total) { bm_set_error("jpeg: output buffer too small (%zu < %zu)", cap, total); return BM_ERR_ARG; }
    memcpy(h_out, hdr, nh);
    if (scan_bytes) BM_CUDA_OK(cudaMemcpyAsync(h_out + nh, j->out, scan_bytes, cudaMemcpyDeviceToHost, s));
    BM_CUDA_OK(cudaStreamSynchronize(s));
    h_out[nh + scan_bytes] = 0xFF; h_out[nh + scan_bytes + 1] = 0xD9;
    return BM_OK;
}

extern "C" size_t bm_jpeg_bound(int w, int h) { return BM_JPEG_HEADER_MAX + bm_jpeg_scan_bound(w, h) + 2; }

// cv2.imwrite(path, img) / cv2.imencode('.jpg', img) of a host BGR image, encoded on `device` (stage entry point; main.py:1664-1665)
extern "C" bm_status bm_jpeg_encode(const uint8_t* h_bgr, int w, int h, int quality, int device, uint8_t* h_out, size_t cap_bytes, size_t* nbytes) {
    if (!h_bgr || !h_out || !nbytes || w < 1 || h < 1 || w > 65535 || h > 65535) { bm_set_error("bm_jpeg_encode: bad args"); return BM_ERR_ARG; }
    BM_CUDA_OK(cudaSetDevice(device));
    BmJpeg j; uint8_t* d_img = nullptr; cudaStream_t s = nullptr; size_t scan = 0;
    bm_status st = BM_ERR_CUDA;
    cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&d_img, (size_t)w * h * 3);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_img, h_bgr, (size_t)w * h * 3, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = bm_jpeg_encode_scan(&j, d_img, w, h, (size_t)w * 3, quality, &scan, s);
    if (e == cudaSuccess) st = bm_jpeg_assemble(&j, w, h, quality, scan, h_out, cap_bytes, nbytes, s);
    else bm_set_error("bm_jpeg_encode: %s", cudaGetErrorString(e));
    bm_jpeg_free(&j); cudaFree(d_img);
    if (s) cudaStreamDestroy(s);
    return st;
}

// main.py:1647-1666 in one call: crop_black_areas + scale_to_screen + the bytes of cv2.imwrite('mosaic.jpg', scaled), all on the device
extern "C" bm_status bm_finalize_jpeg(bm_handle m, int threshold, int margin, int target_w, int target_h, int quality, uint8_t* h_jpeg_out,
                                      size_t cap_bytes, size_t* nbytes, int out_wh[2], int rect[4]) {
    if (!m || !out_wh || !h_jpeg_out || !nbytes) return BM_ERR_ARG;
    bm_status st = bm_finalize_device(m, threshold, margin, target_w, target_h, true, out_wh, rect);
    if (st != BM_OK) return st;
    size_t scan = 0;
    BM_CUDA_OK(bm_jpeg_encode_scan(&m->jpeg, m->d_final, out_wh[0], out_wh[1], (size_t)out_wh[0] * 3, quality, &scan, m->s_chain));
    return bm_jpeg_assemble(&m->jpeg, out_wh[0], out_wh[1], quality, scan, h_jpeg_out, cap_bytes, nbytes, m->s_chain);
}

// Thumbnail of the live canvas for progress callbacks: cv2.cvtColor(BGR2RGB) + PIL Image.resize((out_w, out_h)) of output_img
// (gui.py:143-158 on the copy main.py:1630-1632 hands over), made on the device; ordered after the frames issued so far.
extern "C" bm_status bm_preview(bm_handle m, int out_w, int out_h, int rgb, uint8_t* h_out, size_t cap_bytes) {
    if (!m || !h_out || out_w < 1 || out_h < 1) return BM_ERR_ARG;
    const size_t need = (size_t)out_w * out_h * 3;
    if (cap_bytes < need) { bm_set_error("bm_preview: output buffer too small (%zu < %zu)", cap_bytes, need); return BM_ERR_ARG; }
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    BM_CUDA_OK(bm_preview_prepare(&m->preview, m->cfg.canvas_w, m->cfg.canvas_h, out_w, out_h, m->s_chain));
    BM_CUDA_OK(bm_launch_preview(m->preview, m->blend.canvas, rgb, m->s_chain));
    BM_CUDA_OK(cudaMemcpyAsync(h_out, m->preview.d_out, need, cudaMemcpyDeviceToHost, m->s_chain));
    BM_CUDA_OK(cudaStreamSynchronize(m->s_chain));
    return BM_OK;
}

extern "C" bm_status bm_get_state(bm_handle m, double H_old[9], int* history_len, double* history) {
    if (!m) return BM_ERR_ARG;
    if (H_old) memcpy(H_old, m->H_old, sizeof(m->H_old));
    if (history_len) *history_len = m->history_len;
    if (history) memcpy(history, m->history, sizeof(double) * 9 * m->history_len);
    return BM_OK;
}

extern "C" bm_status bm_set_stabilization(bm_handle m, int enabled, int history_size, double tt, double st) {
    if (!m || history_size < 1 || history_size > 5) return BM_ERR_ARG;
    m->stabilization_enabled = enabled; m->history_size = history_size; m->translation_threshold = tt; m->scale_threshold = st;
    return BM_OK;
}

extern "C" bm_status bm_alloc_pinned(size_t bytes, void** out) {
    if (!out) return BM_ERR_ARG;
    BM_CUDA_OK(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return BM_OK;
}
extern "C" bm_status bm_free_pinned(void* p) { BM_CUDA_OK(cudaFreeHost(p)); return BM_OK; }

// ------------------------------------------------------------------------------------------------------------------
// host control flow of process_frame: validate (main.py:761-801), smooth (:803-834), compose (:746)
// ------------------------------------------------------------------------------------------------------------------
// The ONE implementation of validate_homography (main.py:761-801): the frame loop below, the Python mirror's method and the
// offline pair-sharding chain (sharding.py) all call it.  Pure host code, usable without a device.
extern "C" int bm_validate_homography(const double H[9], double translation_threshold, double scale_threshold, double* value) {
    if (value) *value = 0.0;
    if (!H) return BM_VAL_NAN;
    for (int i = 0; i < 9; ++i) if (isnan(H[i]) || isinf(H[i])) return BM_VAL_NAN;
    const double tr = sqrt(H[2] * H[2] + H[5] * H[5]);
    const double sc = sqrt(H[0] * H[4] - H[1] * H[3]);          // NaN for det < 0: both comparisons false (quirk A.11)
    if (tr > translation_threshold) { if (value) *value = tr; return BM_VAL_TRANSLATION; }
    if (fabs(sc - 1.0) > scale_threshold) { if (value) *value = sc; return BM_VAL_SCALE; }
    if (fabs(H[6]) > 0.001 || fabs(H[7]) > 0.001) return BM_VAL_PERSPECTIVE;
    return BM_VAL_OK;
}
static int validate_h(const bm_mosaic_s* m, const double* H, double* value) {
    return bm_validate_homography(H, m->translation_threshold, m->scale_threshold, value);
}

static void smooth_h(bm_mosaic_s* m, const double* H, double* out) {
    if (!m->stabilization_enabled) { memcpy(out, H, 72); return; }
    if (m->history_len == m->history_size) { memmove(m->history[0], m->history[1], sizeof(double) * 9 * (m->history_size - 1)); m->history_len--; }
    memcpy(m->history[m->history_len++], H, 72);
    const int n = m->history_len;
    if (n < 2) { memcpy(out, H, 72); return; }
    double w[5], sum = 0.0;
    for (int i = 0; i < n; ++i) {                                // np.linspace(0.5, 1.0, n): start + i*step, last = stop
        const double step = (1.0 - 0.5) / (n - 1);
        w[i] = (i == n - 1) ? 1.0 : 0.5 + i * step;
        sum += w[i];
    }
    for (int k = 0; k < 9; ++k) out[k] = 0.0;
    for (int i = 0; i < n; ++i) { const double wi = w[i] / sum; for (int k = 0; k < 9; ++k) out[k] += wi * m->history[i][k]; }
}

static void matmul3(const double* A, const double* B, double* C) {
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) {
        double s = 0.0;
        for (int k = 0; k < 3; ++k) s += A[3 * r + k] * B[3 * k + c];
        C[3 * r + c] = s;
    }
}

// If the next frame is staged, start its match / RANSAC now (its features were detected ahead); the following
// bm_process_frame_begin with the same pointer is then a no-op.  Only with stream overlap on.
static bm_status early_begin(bm_mosaic_s* m) {
    if (!m->overlap || m->nq == 0) return BM_OK;
    const bm_mosaic_s::Staged e = m->q[0];
    const int prev_cur = m->cur;
    m->cur = e.slot;
    q_pop_front(m);
    BM_CUDA_OK(cudaStreamWaitEvent(m->stream, m->ev_up[m->cur], 0));
    bm_status st = bm_pipeline_estimate_begin(m->pipe, m->d_gray[m->cur]);
    if (st < 0) { m->cur = prev_cur; return st; }
    m->begun = e.ptr;
    return BM_OK;
}

// The current frame's detect / match / RANSAC are queued: queue the detectAndCompute of every staged frame behind them NOW, before the
// host blocks on the RANSAC result -- the detect streams then run straight on while the host takes the skip / validate / smooth
// decision and issues the chain.  With two frames staged, three detects are in flight.
static bm_status detect_ahead(bm_mosaic_s* m) {
    if (!m->overlap) return BM_OK;
    for (int i = 0; i < m->nq; ++i) {
        if (m->q[i].detected) continue;
        const int slot = m->q[i].slot;
        BM_CUDA_OK(cudaStreamWaitEvent(m->stream, m->ev_up[slot], 0));
        int done = 0;
        BM_TRY(bm_pipeline_detect_ahead(m->pipe, m->d_gray[slot], &done));
        if (!done) break;                                  // no free keypoint slot yet: try again at the next frame
        BM_CUDA_OK(bm_pipeline_record_after_last_detect(m->pipe, m->ev_spec[slot]));
        m->spec_used[slot] = true;
        m->q[i].detected = true;
    }
    return BM_OK;
}

// an early-begun frame that is not the one the caller continues with: its enqueued work is simply ignored
static void cancel_early_begin(bm_mosaic_s* m) {
    if (m->begun) { m->begun = nullptr; m->cur = BM_SLOT_PREV(m->cur); }
}

static bm_status finish_frame(bm_mosaic_s* m, int slot, bm_frame_info* info_out) {
    BM_NVTX("bm:process_frame_end");
    bm_frame_info info; memset(&info, 0, sizeof(info));
    // one small D2H read of (n_matches, H_rel): the reference's control flow (skip / reject prints) needs them on the host
    double H_rel[9]; int have_h = 0;
    BM_TRY(detect_ahead(m));
    bm_status st = bm_pipeline_estimate_end(m->pipe, &info, H_rel, &have_h);
    if (st < 0) return st;
    if (info.n_matches < 4) { info.status = BM_SKIP_FEW_MATCHES; BM_TRY(early_begin(m)); if (info_out) *info_out = info; return BM_SKIP_FEW_MATCHES; }
    if (!have_h) { info.status = BM_SKIP_NO_H; BM_TRY(early_begin(m)); if (info_out) *info_out = info; return BM_SKIP_NO_H; }
    memcpy(info.H_rel, H_rel, 72);
    double Hv[9]; memcpy(Hv, H_rel, 72);
    info.validate_reason = validate_h(m, H_rel, &info.validate_value);
    bm_status ret = BM_OK;
    if (info.validate_reason != BM_VAL_OK) { for (int i = 0; i < 9; ++i) Hv[i] = (i % 4 == 0) ? 1.0 : 0.0; ret = BM_REJECTED_IDENTITY; }
    double Hs[9], Habs[9];
    double hist_save[5][9]; const int hist_len_save = m->history_len;
    memcpy(hist_save, m->history, sizeof(hist_save));
    smooth_h(m, Hv, Hs);
    matmul3(m->H_old, Hs, Habs);
    memcpy(info.H, Habs, 72);
    {   // nothing is committed before the warp is known to be executable: a window larger than the scratch planes (canvases over
        // 64 Mpx) must leave H_old, the smoothing history and the "previous" features as they were, like a skipped frame
        BmFramePlan plan;
        bm_make_plan(Habs, m->cfg.frame_w, m->cfg.frame_h, m->cfg.canvas_w, m->cfg.canvas_h, &plan);
        const size_t need = (size_t)bm_win_w(plan.reg) * bm_win_h(plan.reg);
        if (plan.valid && need > m->blend.scratch_px) {
            memcpy(m->history, hist_save, sizeof(hist_save)); m->history_len = hist_len_save;
            bm_set_error("warp window %zu px exceeds scratch %zu px", need, m->blend.scratch_px);
            BM_TRY(early_begin(m));
            return BM_ERR_UNSUPPORTED;
        }
    }
    bm_pipeline_advance(m->pipe);                                     // kp_prev/des_prev <- cur (main.py:756-759)
    // the host decision is final: the staged next frame (if any) goes to the detect stream BEFORE this frame's chain is enqueued,
    // so the device never waits for the ~10 launch calls of the chain
    BM_TRY(early_begin(m));
    BM_TRY(warp_device(m, m->d_bgrx[slot], Habs, &info, false, slot));
    memcpy(m->H_old, Habs, 72);
    info.status = ret;
    if (info_out) *info_out = info;
    return ret;
}

// overlap == 0: detect of this frame starts only after the previous frame's chain (used for clean chain timing)
static bm_status order_after_chain(bm_mosaic_s* m) {
    if (m->overlap) return BM_OK;
    for (int i = 0; i < BM_SLOTS; ++i) BM_CUDA_OK(cudaStreamWaitEvent(m->stream, m->ev_chain[i], 0));
    return BM_OK;
}

extern "C" bm_status bm_process_frame_begin(bm_handle m, const uint8_t* h_bgr, size_t stride) {
    if (!m || !h_bgr) { bm_set_error("bm_process_frame_begin: null"); return BM_ERR_ARG; }
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    if (m->begun == h_bgr) { m->begun = nullptr; return BM_OK; }      // already enqueued by the previous frame's _end
    cancel_early_begin(m);
    BM_TRY(take_frame(m, h_bgr, stride, false, m->stream));
    BM_TRY(order_after_chain(m));
    return bm_pipeline_estimate_begin(m->pipe, m->d_gray[m->cur]);
}

extern "C" bm_status bm_process_frame_begin_device(bm_handle m, const uint8_t* d_bgr) {
    if (!m || !d_bgr) { bm_set_error("bm_process_frame_begin_device: null"); return BM_ERR_ARG; }
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    if (m->begun == d_bgr) { m->begun = nullptr; return BM_OK; }      // already enqueued by the previous frame's _end
    cancel_early_begin(m);
    BM_TRY(take_frame(m, d_bgr, 0, true, m->stream));
    BM_TRY(order_after_chain(m));
    return bm_pipeline_estimate_begin(m->pipe, m->d_gray[m->cur]);
}

extern "C" bm_status bm_process_frame_end(bm_handle m, bm_frame_info* info_out) {
    if (!m) return BM_ERR_ARG;
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    return finish_frame(m, m->cur, info_out);
}

extern "C" bm_status bm_process_frame(bm_handle m, const uint8_t* h_bgr, size_t stride, bm_frame_info* info_out) {
    bm_status st = bm_process_frame_begin(m, h_bgr, stride);
    if (st < 0) return st;
    return bm_process_frame_end(m, info_out);
}

extern "C" bm_status bm_process_frame_device(bm_handle m, const uint8_t* d_bgr, bm_frame_info* info_out) {
    bm_status st = bm_process_frame_begin_device(m, d_bgr);
    if (st < 0) return st;
    return bm_process_frame_end(m, info_out);
}

extern "C" bm_status bm_estimate_frame(bm_handle m, const uint8_t* h_bgr, size_t stride, const uint8_t* h_next, bm_frame_info* info_out) {
    if (!m || !h_bgr) { bm_set_error("bm_estimate_frame: null"); return BM_ERR_ARG; }
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    cancel_early_begin(m);
    BM_TRY(take_frame(m, h_bgr, stride, false, m->stream));
    bm_frame_info info; memset(&info, 0, sizeof(info));
    double H_rel[9]; int have_h = 0;
    bm_status st = bm_pipeline_estimate_begin(m->pipe, m->d_gray[m->cur]);
    if (st < 0) return st;
    if (h_next) {                                  // the next frame's H2D + ingest overlap this pair's estimation
        BM_TRY(prefetch(m, h_next, stride, false));
        BM_TRY(detect_ahead(m));                   // and its features are computed while the host waits for this pair
    }
    st = bm_pipeline_estimate_end(m->pipe, &info, H_rel, &have_h);
    if (st < 0) return st;
    bm_pipeline_advance(m->pipe);                  // pair (t-1, t): the frame always becomes "previous"
    bm_status ret = BM_OK;
    if (info.n_matches < 4) ret = BM_SKIP_FEW_MATCHES;
    else if (!have_h) ret = BM_SKIP_NO_H;
    else memcpy(info.H_rel, H_rel, 72);
    info.status = ret;
    if (info_out) *info_out = info;
    return ret;
}

extern "C" bm_status bm_clear_canvas(bm_handle m) {
    if (!m) return BM_ERR_ARG;
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    BM_CUDA_OK(cudaMemsetAsync(m->blend.canvas, 0, (size_t)m->cfg.canvas_h * m->cfg.canvas_w * sizeof(uchar4), m->s_chain));
    BM_CUDA_OK(bm_launch_full_rowscan(m->blend, m->s_chain));
    BM_CUDA_OK(cudaStreamSynchronize(m->s_chain));
    return BM_OK;
}

// `video_mosaic.output_img = img` of a reference caller: replace the device canvas by a host image (packed BGR, Hc x Wc x 3)
extern "C" bm_status bm_set_canvas(bm_handle m, const uint8_t* h_bgr) {
    if (!m || !h_bgr) return BM_ERR_ARG;
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    const size_t n = (size_t)m->cfg.canvas_h * m->cfg.canvas_w;
    BM_CUDA_OK(cudaStreamSynchronize(m->s_chain));
    BM_CUDA_OK(cudaMemcpy(m->d_canvas_bgr, h_bgr, n * 3, cudaMemcpyHostToDevice));
    BM_CUDA_OK(bm_launch_pack_canvas(m->d_canvas_bgr, m->blend.canvas, (int)n, m->s_chain));
    BM_CUDA_OK(bm_launch_full_rowscan(m->blend, m->s_chain));
    BM_CUDA_OK(cudaStreamSynchronize(m->s_chain));
    return BM_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// canvas row tiles (config 5): the handle's canvas is one tile (+ halo rows) of a larger canvas; what crosses tile boundaries is
// (a) the distance-transform sweep state at the tile's top / bottom edge and (b) the pixels of halo rows a neighbour changed
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_copy_rect(const uchar4* __restrict__ src, int src_stride, uchar4* __restrict__ dst, int dst_stride, int w, int h) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x < w && y < h) dst[(size_t)y * dst_stride + x] = src[(size_t)y * src_stride + x];
}

// side 0: the rows above the tile (downward sweeps enter through the top edge), 1: below.  d_rows: [3][canvas_w] u32 (E1, E2, V) as
// exported by the neighbour, or NULL for "canvas border" (cv2's semantics: nothing beyond).
extern "C" bm_status bm_tile_set_ghost(bm_handle m, int side, const uint32_t* d_rows) {
    if (!m || side < 0 || side > 1) return BM_ERR_ARG;
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    BmDtPlane& p = m->blend.dt.p[0];
    if (!d_rows) { (side ? p.gh_bot : p.gh_top) = nullptr; return BM_OK; }
    if (!m->d_ghost[side]) BM_CUDA_OK(cudaMalloc(&m->d_ghost[side], (size_t)3 * p.ts * sizeof(uint32_t)));
    for (int i = 0; i < 3; ++i)
        BM_CUDA_OK(cudaMemcpyAsync(m->d_ghost[side] + (size_t)i * p.ts, d_rows + (size_t)i * p.W, (size_t)p.W * sizeof(uint32_t), cudaMemcpyDeviceToDevice, m->s_chain));
    (side ? p.gh_bot : p.gh_top) = m->d_ghost[side];
    BM_CUDA_OK(cudaStreamSynchronize(m->s_chain));          // the caller may reuse d_rows
    return BM_OK;
}

// the carries a neighbouring tile needs: up == 0 -> downward sweeps at the LAST row of 16-row block `block` (for the tile below),
// up == 1 -> upward sweeps at the FIRST row of `block` (for the tile above).  d_out: [3][canvas_w] u32.  Enqueued on the chain stream
// (after every frame blended so far); bm_sync / a stream wait makes the result visible.
extern "C" bm_status bm_tile_export_carries(bm_handle m, int up, int block, uint32_t* d_out) {
    if (!m || !d_out) return BM_ERR_ARG;
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    BM_CUDA_OK(bm_launch_dt_export_carries(m->blend.dt.p[0], up ? 1 : 0, block, d_out, m->s_chain));
    BM_CUDA_OK(cudaStreamSynchronize(m->s_chain));
    return BM_OK;
}

// canvas rectangle (x0, y0, w, h) <-> a dense device buffer of w * h uchar4 (B, G, R, mask).  Import also refreshes the persistent
// row seeds and block-local sweeps of the rows it touched, exactly like a blend does.
extern "C" bm_status bm_tile_export_rect(bm_handle m, int x0, int y0, int w, int h, uint8_t* d_out_bgrx) {
    if (!m || !d_out_bgrx || x0 < 0 || y0 < 0 || w <= 0 || h <= 0 || x0 + w > m->cfg.canvas_w || y0 + h > m->cfg.canvas_h) return BM_ERR_ARG;
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    BM_COUNT_LAUNCHES(1), k_copy_rect<<<dim3(bm_div_up(w, 256), h), 256, 0, m->s_chain>>>(m->blend.canvas + (size_t)y0 * m->cfg.canvas_w + x0, m->cfg.canvas_w,
                                                                                        reinterpret_cast<uchar4*>(d_out_bgrx), w, w, h);
    BM_CUDA_OK(cudaGetLastError());
    BM_CUDA_OK(cudaStreamSynchronize(m->s_chain));
    return BM_OK;
}
extern "C" bm_status bm_tile_import_rect(bm_handle m, int x0, int y0, int w, int h, const uint8_t* d_in_bgrx) {
    if (!m || !d_in_bgrx || x0 < 0 || y0 < 0 || w <= 0 || h <= 0 || x0 + w > m->cfg.canvas_w || y0 + h > m->cfg.canvas_h) return BM_ERR_ARG;
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    BM_COUNT_LAUNCHES(1), k_copy_rect<<<dim3(bm_div_up(w, 256), h), 256, 0, m->s_chain>>>(reinterpret_cast<const uchar4*>(d_in_bgrx), w,
                                                                                        m->blend.canvas + (size_t)y0 * m->cfg.canvas_w + x0, m->cfg.canvas_w, w, h);
    BM_CUDA_OK(cudaGetLastError());
    const BmDtPlane& po = m->blend.dt.p[0];
    BM_CUDA_OK(bm_launch_rowscan_bgrx(m->blend.canvas, m->cfg.canvas_w, 0, y0, po, y0, h, m->blend.flags, 0, m->s_chain));
    BM_CUDA_OK(bm_launch_dt_local(po, y0 / BM_BLK_ROWS, bm_div_up(y0 + h, BM_BLK_ROWS), m->blend.flags, 0, m->s_chain));
    BM_CUDA_OK(cudaStreamSynchronize(m->s_chain));
    return BM_OK;
}

extern "C" bm_status bm_get_canvas_device(bm_handle m, uint8_t* d_out) {
    if (!m || !d_out) return BM_ERR_ARG;
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    BM_CUDA_OK(bm_launch_unpack_canvas(m->blend.canvas, d_out, (int)((size_t)m->cfg.canvas_h * m->cfg.canvas_w), m->s_chain));
    BM_CUDA_OK(cudaStreamSynchronize(m->s_chain));
    return BM_OK;
}

// features / matches of the last frame
#include "orb.cuh"
#include "match.cuh"
bm_status bm_download_keypoints(const BmKeypoints& k, int desc_bytes, float* h_kp, uint8_t* h_desc, int cap, int* n_out, cudaStream_t s);

extern "C" bm_status bm_get_keypoints(bm_handle m, int which, float* h_kp, uint8_t* h_desc, int cap, int* n_out) {
    if (!m) return BM_ERR_ARG;
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    BM_CUDA_OK(bm_pipeline_sync_est(m->pipe));                // the features are written on the pipeline's detector streams
    return bm_download_keypoints(*bm_pipeline_keypoints(m->pipe, which), m->cfg.detector == BM_DET_ORB ? 32 : 128, h_kp, h_desc, cap, n_out, m->stream);
}

extern "C" bm_status bm_get_matches(bm_handle m, int* h_q, int* h_t, float* h_dist, int cap, int* m_out) {
    if (!m) return BM_ERR_ARG;
    BM_CUDA_OK(cudaSetDevice(m->cfg.device));
    BmMatches* mm = bm_pipeline_matches(m->pipe);
    int M = 0;
    BM_CUDA_OK(cudaMemcpyAsync(&M, mm->count, 4, cudaMemcpyDeviceToHost, m->stream));
    BM_CUDA_OK(cudaStreamSynchronize(m->stream));
    if (M > cap) { bm_set_error("match buffer too small"); return BM_ERR_ARG; }
    if (M > 0) {
        BM_CUDA_OK(cudaMemcpyAsync(h_q, mm->q, M * 4, cudaMemcpyDeviceToHost, m->stream));
        BM_CUDA_OK(cudaMemcpyAsync(h_t, mm->t, M * 4, cudaMemcpyDeviceToHost, m->stream));
        BM_CUDA_OK(cudaMemcpyAsync(h_dist, mm->dist, M * 4, cudaMemcpyDeviceToHost, m->stream));
        BM_CUDA_OK(cudaStreamSynchronize(m->stream));
    }
    if (m_out) *m_out = M;
    return BM_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// stage entry points
// ------------------------------------------------------------------------------------------------------------------
extern "C" bm_status bm_ingest_bgr(const uint8_t* d_bgr, int h, int w, uint8_t* d_gray, uint8_t* d_bgrx, void* stream) {
    if (!d_bgr || h <= 0 || w <= 0) return BM_ERR_ARG;
    BM_CUDA_OK(bm_launch_ingest(d_bgr, h, w, d_gray, reinterpret_cast<uchar4*>(d_bgrx), (cudaStream_t)stream));
    return BM_OK;
}

extern "C" bm_status bm_warp_perspective_bgr(const uint8_t* d_src, int sh, int sw, const double H[9], uint8_t* d_dst, int dh, int dw, void* stream) {
    if (!d_src || !d_dst || !H) return BM_ERR_ARG;
    BmFramePlan plan;
    bm_make_plan(H, sw, sh, dw, dh, &plan);
    BM_CUDA_OK(bm_launch_warp_full_bgr(d_src, sh, sw, plan, d_dst, (cudaStream_t)stream));
    return BM_OK;
}

extern "C" bm_status bm_distance_transform(const uint8_t* d_mask, int h, int w, float* d_out, void* stream) {
    if (!d_mask || !d_out || h <= 0 || w <= 0) return BM_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    BmDtPlane p;
    cudaError_t e = bm_dt_alloc_plane(&p, w, h, (size_t)w * h);
    if (e == cudaSuccess) e = bm_launch_rowscan_u8(d_mask, w, p, s);
    if (e == cudaSuccess) e = bm_launch_dt_map(p, d_out, s);
    cudaStreamSynchronize(s);
    bm_dt_free_plane(&p);
    BM_CUDA_OK(e);
    return BM_OK;
}

extern "C" bm_status bm_gaussian_blur31(const float* d_in, int h, int w, float* d_out, void* stream) {
    if (!d_in || !d_out || h < 16 || w < 16) return BM_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    float* tmp = nullptr;
    BM_CUDA_OK(cudaMalloc(&tmp, (size_t)h * w * 4));
    cudaError_t e = bm_launch_blur31(d_in, h, w, tmp, d_out, s);
    cudaStreamSynchronize(s);
    cudaFree(tmp);
    BM_CUDA_OK(e);
    return BM_OK;
}

extern "C" bm_status bm_blend_step_bgr(uint8_t* d_canvas, const uint8_t* d_warped, int dh, int dw, const int* win, int* any_overlap, void* stream) {
    if (!d_canvas || !d_warped || dh < 32 || dw < 32) return BM_ERR_ARG;
    cudaStream_t s = (cudaStream_t)stream;
    BmBlendBufs b;
    bm_status st = alloc_blend(b, dh, dw, 0);
    if (st != BM_OK) { free_blend(b); return st; }
    BmFramePlan plan; memset(&plan, 0, sizeof(plan));
    plan.canvas_w = dw; plan.canvas_h = dh; plan.valid = 1; plan.block_w = 64;
    BmWin w = {0, 0, dw, dh};
    if (win) {   // grow by one pixel so that the window keeps a ring of zero pixels (see k_dt_weights)
        w.x0 = win[0] - 1 < 0 ? 0 : win[0] - 1; w.y0 = win[1] - 1 < 0 ? 0 : win[1] - 1;
        w.x1 = win[2] + 1 > dw ? dw : win[2] + 1; w.y1 = win[3] + 1 > dh ? dh : win[3] + 1;
    }
    plan.win = w;
    bm_finish_plan(&plan);
    cudaError_t e = cudaSuccess;
    if (plan.valid) {
        e = bm_launch_pack_canvas(d_canvas, b.canvas, dh * dw, s);
        if (e == cudaSuccess) e = bm_launch_full_rowscan(b, s);
        if (e == cudaSuccess) e = bm_launch_extract_wbuf(d_warped, plan, b, s);
        if (e == cudaSuccess) e = bm_launch_blend_from_wbuf(b, plan, s);
        if (e == cudaSuccess) e = bm_launch_unpack_canvas(b.canvas, d_canvas, dh * dw, s);
        int f = 0;
        if (e == cudaSuccess) e = cudaMemcpyAsync(&f, b.flags, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (any_overlap) *any_overlap = f;
    }
    free_blend(b);
    BM_CUDA_OK(e);
    return BM_OK;
}
