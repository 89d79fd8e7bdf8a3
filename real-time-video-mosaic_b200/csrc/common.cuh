// Shared declarations for the b200mosaic CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define BM_CHAMFER_A 62587          // cvRound(0.955f  * 65536), OpenCV distanceTransform DIST_L2 3x3
#define BM_CHAMFER_B 89738          // cvRound(1.3693f * 65536)
#define BM_DT_INIT 4294877557u      // OpenCV 4.x distanceTransform_3x3: DIST_MAX = UINT_MAX - DIAG_DIST (saturation value)
#define BM_G_INF 0xFFFF             // "no zero pixel in this row" (row-scan distances are clamped to it)
#define BM_CHAMFER_SEED(gv) ((gv) == (unsigned)BM_G_INF ? BM_DT_INIT : (unsigned)BM_CHAMFER_A * (gv))   // g <= 65534 keeps a * g < DIST_MAX
#define BM_BLK_ROWS 16              // rows per block of the distance-transform sweep tables
#define BM_BLUR_R 15                // 31-tap Gaussian radius

struct BmWin { int x0, y0, x1, y1; };   // half-open pixel rectangle in canvas coordinates
static inline __host__ __device__ int bm_win_w(const BmWin& w) { return w.x1 - w.x0; }
static inline __host__ __device__ int bm_win_h(const BmWin& w) { return w.y1 - w.y0; }

// Per-frame parameters every kernel of the warp/blend chain reads from device memory, so that the chain can be
// enqueued without a host round trip (the host mirrors them only to size the grids).
struct BmFramePlan {
    double M[9];        // inverse homography (canvas -> frame), cofactor form
    BmWin win;          // W: clipped bounding box of the warped frame (+ zero ring)
    BmWin reg;          // R: W dilated by the blur radius, clipped
    int src_w, src_h;
    int canvas_w, canvas_h;
    int block_w;        // OpenCV's warpPerspective evaluation block width (64 for canvases >= 64 px wide)
    int valid;          // 0 -> nothing to do this frame
    int ws;             // row stride of the window scratch (warped frame, g_new): win width padded to 8
    int rx0, rws;       // weight planes over R: column origin (reg.x0 rounded down to 4) and row stride
};
// win -> aligned win (x0 to 8 columns, y0 to 16 rows), reg, strides.  Host side (warp_blend.cu).
void bm_finish_plan(BmFramePlan* p);

extern "C" const char* bm_last_error(void);
void bm_set_error(const char* fmt, ...);

#define BM_CUDA_OK(expr)                                                                   \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            bm_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return BM_ERR_CUDA;                                                            \
        }                                                                                  \
    } while (0)

extern long long g_bm_launches;          // kernels launched by this library (bench.py reports it); handles may be driven from several threads
extern thread_local long long* t_bm_launch_sink;   // non-null while this thread captures a CUDA graph: the launches are counted per replay instead
#define BM_COUNT_LAUNCHES(n) ((void)__atomic_fetch_add(t_bm_launch_sink ? t_bm_launch_sink : &g_bm_launches, (long long)(n), __ATOMIC_RELAXED))

// NVTX ranges around the stages of the per-frame path (visible in Nsight Systems / ncu --nvtx; header-only, no library needed)
#include <nvtx3/nvToolsExt.h>
struct BmNvtxRange {
    explicit BmNvtxRange(const char* name) { nvtxRangePushA(name); }
    ~BmNvtxRange() { nvtxRangePop(); }
};
#define BM_NVTX(name) BmNvtxRange bm_nvtx_range_##__LINE__(name)

// Opt a kernel into more than 48 KB of dynamic shared memory, once per (call site, device): the attribute is per device, and one
// process may drive handles on several devices (bm_config.device).  `err` receives the CUDA status.
#define BM_SMEM_OPTIN(kernel, bytes, err)                                                                                   \
    do {                                                                                                                    \
        static unsigned long long bm_optin_done_ = 0ull;                                                                    \
        int bm_dev_ = 0;                                                                                                    \
        (err) = cudaGetDevice(&bm_dev_);                                                                                    \
        if ((err) == cudaSuccess && !((bm_optin_done_ >> (bm_dev_ & 63)) & 1ull)) {                                         \
            (err) = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes));                 \
            if ((err) == cudaSuccess) bm_optin_done_ |= 1ull << (bm_dev_ & 63);                                             \
        }                                                                                                                   \
    } while (0)

static inline int bm_div_up(int a, int b) { return (a + b - 1) / b; }
#define BM_DET_MAX_GRAPHS 32     // detector graph cache per instance: (frame slot, keypoint slot) pairs, 5 x 5 with three frames of look-ahead

// Programmatic dependent launch for back-to-back kernels of one stream (the warp / blend chain): a kernel launched through
// bm_launch_pdl is set up (launch processing, CTA scheduling) while its predecessor drains instead of after it; it starts with
// BM_PDL_WAIT (griddepcontrol.wait: returns once the predecessor grid has completed and its writes are visible) before it touches
// global memory.  Both macros are no-ops for a launch without the attribute; BM_NO_PDL=1 falls back to plain launches.
// Measured (1080p, chain alone / whole pipeline): plain launches 143 us / SIFT 1815 frames/s; PDL with the implicit trigger at CTA
// exit 125 us / 1819; PDL with an EARLY trigger (griddepcontrol.launch_dependents at the top of every kernel, -DBM_PDL_EARLY_TRIGGER)
// 129 us / 1765 -- the successor's CTAs then sit on registers and shared memory that the detector kernels of the other streams would
// have used.  Hence no explicit trigger.
#ifdef BM_PDL_EARLY_TRIGGER
#define BM_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#else
#define BM_PDL_TRIGGER() ((void)0)
#endif
#define BM_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
bool bm_pdl_enabled();
template <typename... KA, typename... A>
static inline cudaError_t bm_launch_pdl(void (*kernel)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, A&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = bm_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KA>(args)...);
}

// Stream priorities: the per-frame rate is bound by the LATENCY of detect -> match -> RANSAC (the caller hands frames over one ahead, so
// a frame period is (L_detect + L_estimate) / 2), while the warp/blend chain only has to keep up.  level 2 = latency critical
// (match + RANSAC: a few one-CTA kernels), 1 = detectors, 0 = chain / copies.
static inline cudaError_t bm_stream_create(cudaStream_t* s, int level) {
    int lo = 0, hi = 0;                                   // lo = least urgent (numerically largest), hi = most urgent
    cudaError_t e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (e != cudaSuccess) return e;
    int prio = lo - level;
    if (prio < hi) prio = hi;
    return cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, prio);
}
