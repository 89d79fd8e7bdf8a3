/* b200mosaic.h -- C ABI of libb200mosaic.so (hand-written sm_100a CUDA; no torch / C++ types cross this line).
 *
 * The reference (PROcessorI/Real-Time-Video-Mosaic) has no FFI: its boundary for the stitching hot path is the
 * Python class `VideMosaic` (/root/reference/main.py:15-112, 676-977) whose methods call OpenCV.  Each entry
 * point below replaces one of those calls (cited per function); the Python mirror of the class
 * (`real-time-video-mosaic_b200/mosaic.py`, ctypes) binds exactly these symbols.  See INTEGRATION.md.
 *
 * Conventions: all functions return bm_status (0 = OK, >0 = soft outcome of the reference's control flow,
 * <0 = error; bm_last_error() gives the text).  No exceptions cross the ABI.  Pointers named d_* are DEVICE
 * pointers in the current CUDA primary context (e.g. torch tensors' data_ptr()); h_* are HOST pointers.
 * `stream` is a cudaStream_t passed as void* (NULL = default stream).  Images are row-major, tightly packed
 * unless a stride is given; colour order is OpenCV's BGR.
 */
#ifndef B200MOSAIC_H
#define B200MOSAIC_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    BM_OK = 0,
    BM_SKIP_FEW_MATCHES = 1,   /* main.py:722-724  (<4 matches: frame skipped, state not advanced)        */
    BM_SKIP_NO_H = 2,          /* main.py:729-731  (findHomography returned None)                          */
    BM_REJECTED_IDENTITY = 3,  /* main.py:734-737  (validate_homography failed -> identity substituted)    */
    BM_ERR_CUDA = -1,
    BM_ERR_ARG = -2,
    BM_ERR_UNSUPPORTED = -3
} bm_status;

enum { BM_DET_SIFT = 0, BM_DET_ORB = 1 };

/* reasons reported by validate_homography (main.py:761-801), for the Python shim's prints */
enum { BM_VAL_OK = 0, BM_VAL_NAN = 1, BM_VAL_TRANSLATION = 2, BM_VAL_SCALE = 3, BM_VAL_PERSPECTIVE = 4 };

typedef struct bm_mosaic_s* bm_handle;

typedef struct {
    int frame_h, frame_w;      /* first_image.shape[:2]                       main.py:17          */
    int canvas_h, canvas_w;    /* int(oht*H), int(owt*W) or an explicit size  main.py:80-81       */
    int detector;              /* BM_DET_SIFT | BM_DET_ORB                    main.py:32-37       */
    int nfeatures;             /* 700                                         main.py:33,36       */
    int device;                /* CUDA device ordinal                                             */
} bm_config;

typedef struct {
    int status;                /* bm_status of the frame                                          */
    int n_kp_cur, n_kp_prev;
    int n_matches;
    int ransac_iters;
    int n_inliers;
    int validate_reason;       /* BM_VAL_*                                                        */
    int any_overlap;           /* np.any(overlap) of main.py:885                                  */
    int win[4];                /* x0,y0,x1,y1 of the warped frame's bounding window on the canvas */
    double validate_value;     /* translation px or scale, for the reference's warning text       */
    double H_rel[9];           /* raw RANSAC result (before validation), main.py:727              */
    double H[9];               /* absolute homography used for the warp,  main.py:746             */
} bm_frame_info;

const char* bm_last_error(void);
int bm_version(void);

/* ---- whole-path handle: VideMosaic.__init__ / process_frame / output_img ------------------------------ */
bm_status bm_create(const bm_config* cfg, bm_handle* out);                       /* main.py:17-102 (minus YOLO)  */
bm_status bm_destroy(bm_handle h);
/* frame 0: features + paste at rows [Hc-H,Hc), cols [Wc/2-W/2,+W); H_old = translation.   main.py:78-94,104-112 */
bm_status bm_first_frame(bm_handle h, const uint8_t* h_bgr, size_t stride_bytes);
/* one process_frame (main.py:710-759): H2D, gray, detect, match, RANSAC, validate, smooth, compose, warp, blend */
bm_status bm_process_frame(bm_handle h, const uint8_t* h_bgr, size_t stride_bytes, bm_frame_info* info);
/* same, with the frame already resident in device memory as packed BGR (kernel-only timing leg of bench.py) */
bm_status bm_process_frame_device(bm_handle h, const uint8_t* d_bgr, bm_frame_info* info);
/* split form for many concurrent handles (config 4: 64 streams): _begin enqueues H2D + detect + match + RANSAC + the small
 * D2H on the handle's stream and returns without waiting; _end waits for that read-back, runs the host control flow and
 * enqueues the warp/blend chain.  Calling _begin on every handle and then _end on every handle overlaps the streams. */
bm_status bm_process_frame_begin(bm_handle h, const uint8_t* h_bgr, size_t stride_bytes);
bm_status bm_process_frame_begin_device(bm_handle h, const uint8_t* d_bgr);
bm_status bm_process_frame_end(bm_handle h, bm_frame_info* info);
/* double-buffered ingest (north_star: "pinned, double-buffered H2D copies"): start the H2D + BGR->gray/BGRX of the NEXT frame on
 * the copy stream while the current one is processed; the following bm_process_frame / _begin / bm_warp_frame / bm_estimate_frame
 * call with the same host pointer consumes the staged copy instead of uploading again.  The buffer must stay unchanged in
 * between (pageable buffers are copied to pinned staging memory immediately).  May be called for up to THREE frames ahead, in the order
 * they will be processed: each staged frame's detectAndCompute is queued while the current frame is still in flight. */
bm_status bm_prefetch_frame(bm_handle h, const uint8_t* h_bgr, size_t stride_bytes);
bm_status bm_prefetch_frame_device(bm_handle h, const uint8_t* d_bgr);     /* frame already in device memory (packed BGR) */
/* 1 (default): the warp/blend chain of frame t runs concurrently with detect/match/RANSAC of frame t+1 (separate streams);
 * 0: strictly one after the other (used to time the chain alone) */
bm_status bm_set_overlap(bm_handle h, int on);
/* offline pair-sharded mode (config 3 at N GPUs): features of this frame, matches and RANSAC against the previous frame of
 * this handle, NO validation / warp; the frame always becomes the new "previous".  info->H_rel, n_matches, status
 * (BM_OK | BM_SKIP_FEW_MATCHES | BM_SKIP_NO_H) are filled.  h_next (optional): the frame of the next call, staged (H2D + ingest) while
 * this pair is estimated. */
bm_status bm_estimate_frame(bm_handle h, const uint8_t* h_bgr, size_t stride_bytes, const uint8_t* h_next_or_null, bm_frame_info* info);
/* canvas row-tile mode (config 5): empty the canvas (the caller then warps frame 0 with its own homography) */
bm_status bm_clear_canvas(bm_handle h);
/* `output_img = image` (callers of the reference may assign the attribute): replace the canvas by a host image, Hc x Wc x 3 BGR */
bm_status bm_set_canvas(bm_handle h, const uint8_t* h_bgr);
/* ---- canvas row tiles (config 5: a canvas sharded into row tiles over several GPUs).  The reference has no such mode (its blend,
 * main.py:878-927, is global over the canvas); these four calls carry across tile boundaries exactly what that global blend needs, so
 * that tiles reproduce the untiled canvas bit for bit: the sweep state of cv2.distanceTransform(mask_old) (main.py:889) entering a
 * tile from the rows above / below, and the pixels of halo rows.  See real-time-video-mosaic_b200/sharding.py (TileGroup). ---- */
bm_status bm_tile_set_ghost(bm_handle h, int side /*0 above, 1 below*/, const uint32_t* d_rows /*[3][canvas_w] or NULL = canvas border*/);
bm_status bm_tile_export_carries(bm_handle h, int up, int block, uint32_t* d_out /*[3][canvas_w]*/);
bm_status bm_tile_export_rect(bm_handle h, int x0, int y0, int w, int hgt, uint8_t* d_out_bgrx /* w * hgt * 4 bytes */);
bm_status bm_tile_import_rect(bm_handle h, int x0, int y0, int w, int hgt, const uint8_t* d_in_bgrx);
/* canvas as packed BGR into a DEVICE buffer (e.g. a torch tensor that NCCL then gathers) */
bm_status bm_get_canvas_device(bm_handle h, uint8_t* d_bgr_out);
/* output_img (uint8, Hc x Wc x 3): lazy D2H of the device canvas                     main.py:1632,1649 */
bm_status bm_get_canvas(bm_handle h, uint8_t* h_bgr_out);
/* mosaic finalisation on the device: crop_black_areas(output_img, threshold, margin) (main.py:980-1003) followed by
 * scale_to_screen(cropped, target_w, target_h) (main.py:1006-1038; target <= 0 -> 1920 x 1080 like the reference off Windows), as
 * main() does before writing mosaic.jpg (main.py:1647-1659).  out_wh = (width, height) of the result, rect = (x, y, w, h) of the
 * crop.  h_out == NULL only computes the sizes; otherwise h_out receives height x width x 3 BGR bytes. */
bm_status bm_finalize(bm_handle h, int threshold, int margin, int target_w, int target_h, uint8_t* h_bgr_out, size_t cap_bytes,
                      int out_wh[2], int rect[4]);
/* the file cv2.imwrite(os.path.join(output_dir, 'mosaic.jpg'), scaled_mosaic) writes (main.py:1664-1665), encoded on the device:
 * baseline JPEG, 4:2:0, Annex K Huffman tables, quality as cv2's IMWRITE_JPEG_QUALITY (the reference uses the default, 95); byte for
 * byte the output of cv2 4.13's bundled libjpeg.  bm_finalize_jpeg = bm_finalize followed by that encoder without leaving the device:
 * only the compressed file crosses PCIe.  *nbytes = size of the file; BM_ERR_ARG with *nbytes set when cap_bytes is too small
 * (bm_jpeg_bound(w, h) is always enough).  bm_jpeg_encode is the stage entry point for a host BGR image. */
size_t bm_jpeg_bound(int w, int h);
bm_status bm_jpeg_encode(const uint8_t* h_bgr, int w, int h, int quality, int device, uint8_t* h_out, size_t cap_bytes, size_t* nbytes);
bm_status bm_finalize_jpeg(bm_handle h, int threshold, int margin, int target_w, int target_h, int quality, uint8_t* h_jpeg_out,
                           size_t cap_bytes, size_t* nbytes, int out_wh[2], int rect[4]);
/* live-preview thumbnail made on the device: what the GUI's progress callback computes from a full-canvas copy --
 * cv2.cvtColor(BGR2RGB), Image.fromarray, Image.resize((out_w, out_h)) with Pillow's default bicubic filter (gui.py:143-158 on the
 * copy handed over at main.py:1630-1632).  h_out receives out_h x out_w x 3 bytes, RGB when rgb != 0 (the GUI's order) else BGR. */
bm_status bm_preview(bm_handle h, int out_w, int out_h, int rgb, uint8_t* h_out, size_t cap_bytes);
/* optional: capture every CUDA graph the per-frame path replays now instead of lazily during the first frames of a stream (the
 * reference has no equivalent; its first process_frame is as slow as any other, main.py:711).  Executes nothing. */
bm_status bm_warm_up(bm_handle h);
bm_status bm_get_state(bm_handle h, double H_old[9], int* history_len, double* history /* <=5*9 */);
bm_status bm_set_stabilization(bm_handle h, int enabled, int history_size, double translation_threshold,
                               double scale_threshold);                          /* main.py:97-101 */
/* validate_homography(H) (main.py:761-801): returns BM_VAL_OK or the reason of the rejection; *value = the translation (px) or
 * scale the reference prints.  Reproduces the quirk that a negative 2x2 determinant gives sqrt -> NaN and therefore PASSES.
 * Host-only (no device needed); the frame loop, the Python mirror and the pair-sharding chain all use this one implementation. */
int bm_validate_homography(const double H[9], double translation_threshold, double scale_threshold, double* value);
/* pinned staging memory a decoder can write into directly (cv2.VideoCapture stays on the host)           */
bm_status bm_alloc_pinned(size_t bytes, void** out);
bm_status bm_free_pinned(void* p);
/* VideMosaic.warp(frame, H) on the handle's canvas (main.py:861-927): warpPerspective + blend, device canvas. */
bm_status bm_warp_frame(bm_handle h, const uint8_t* h_bgr, size_t stride_bytes, const double H[9], bm_frame_info* info);
/* same with the frame already resident on the device as packed BGRX (uchar4); used by the kernel-only bench leg */
/* same as bm_warp_frame but only enqueues (no any_overlap read-back, no wait): canvas-tile mode warps many frames back to back */
bm_status bm_warp_frame_async(bm_handle h, const uint8_t* h_bgr, size_t stride_bytes, const double H[9]);
bm_status bm_warp_frame_device(bm_handle h, const uint8_t* d_bgrx, const double H[9], bm_frame_info* info);
/* timing helpers: last warp/blend chain duration measured with CUDA events on the handle's stream (ms) */
bm_status bm_sync(bm_handle h);
/* CUDA-event timing of the warp/blend chain on the handle's stream: enable, then read the accumulated milliseconds,
 * algorithmic bytes (3N + 6A per frame) and frame count since the last reset */
bm_status bm_timing_enable(bm_handle h, int on);
bm_status bm_timing_read(bm_handle h, double* warp_blend_ms, double* algorithmic_bytes, int* frames, int reset);
/* number of kernels this library has launched in the calling process (all handles) */
long long bm_kernel_launches(void);
void* bm_stream(bm_handle h);
/* device pointer of the handle's current BGRX frame buffer after an upload (for the kernel-only bench leg) */
bm_status bm_upload_frame(bm_handle h, const uint8_t* h_bgr, size_t stride_bytes, const uint8_t** d_bgrx_out);

/* ---- stage entry points (parity tests call these with device buffers) --------------------------------- */
/* cv2.cvtColor(BGR2GRAY)  main.py:111,717 ; also emits the BGRX copy the warp kernel samples (either may be NULL) */
bm_status bm_ingest_bgr(const uint8_t* d_bgr, int h, int w, uint8_t* d_gray, uint8_t* d_bgrx, void* stream);
/* cv2.warpPerspective(frame, H, (Wc,Hc), INTER_LINEAR)  main.py:871 ; d_dst is Hc x Wc x 3, fully written */
bm_status bm_warp_perspective_bgr(const uint8_t* d_src_bgr, int sh, int sw, const double H[9],
                                  uint8_t* d_dst_bgr, int dh, int dw, void* stream);
/* cv2.distanceTransform(mask, DIST_L2, 3)  main.py:888-889 ; mask u8 (0 = zero pixel), out float32 */
bm_status bm_distance_transform(const uint8_t* d_mask, int h, int w, float* d_out, void* stream);
/* cv2.GaussianBlur(w, (31,31), 0) on float32  main.py:897-898 */
bm_status bm_gaussian_blur31(const float* d_in, int h, int w, float* d_out, void* stream);
/* the blend of VideMosaic.warp  main.py:878-927 : canvas (Hc x Wc x 3 u8, in/out) <- warped (Hc x Wc x 3 u8).
 * win = optional x0,y0,x1,y1 bounding window of the non-zero part of `warped` (NULL = whole canvas).      */
bm_status bm_blend_step_bgr(uint8_t* d_canvas_bgr, const uint8_t* d_warped_bgr, int dh, int dw,
                            const int* win, int* any_overlap_out, void* stream);

/* ---- feature / matching / RANSAC stage entry points (small host arrays; parity tests + the Python mirror) ---------- */
/* cv2.ORB_create(n).detectAndCompute(gray, None)  main.py:36,112,718.  h_kp rows: x, y, size, angle, response, octave
 * (float32 x 6); h_desc: n x 32 uint8.  Order: cv2's own (level-major; inside a level the order KeyPointsFilter::retainBest,
 * i.e. libstdc++'s nth_element + partition, leaves -- emulated on the device, see bm_cv_retain_best). */
bm_status bm_orb_detect_and_compute(const uint8_t* d_gray, int h, int w, int nfeatures, float* h_kp, uint8_t* h_desc,
                                    int cap, int* n_out);
/* cv2.SIFT_create(n).detectAndCompute(gray, None)  main.py:33,112,718.  h_desc: n x 128 float32 (integer valued). */
bm_status bm_sift_detect_and_compute(const uint8_t* d_gray, int h, int w, int nfeatures, float* h_kp, float* h_desc,
                                     int cap, int* n_out);
/* cv::KeyPointsFilter::retainBest(keypoints, n_points) as the detectors above apply it (inside detectAndCompute, main.py:112,718):
 * which of the n items (responses h_resp, input order) survive and IN WHICH ORDER -- std::nth_element(begin, begin + n_points - 1,
 * end, response >) + std::partition(begin + n_points, end, response >= boundary) of libstdc++, reproduced element for element by
 * parallel pairing passes on one CTA.  as_u8 != 0: keys are integers 0..255 (FAST scores) handled as bytes.  h_idx_out (capacity n)
 * receives the surviving input indices in output order, *m_out their number. */
bm_status bm_cv_retain_best(const float* h_resp, int n, int n_points, int as_u8, int* h_idx_out, int* m_out);
/* BFMatcher(NORM_HAMMING, crossCheck=True).match + sorted(key=distance)  main.py:694-698 */
bm_status bm_match_hamming_crosscheck(const uint8_t* h_des_q, int nq, const uint8_t* h_des_t, int nt,
                                      int* h_q, int* h_t, float* h_dist, int* m_out);
/* BFMatcher().knnMatch(k=2) + ratio test (double compare) + sorted(key=distance)  main.py:687-698 */
bm_status bm_match_l2_knn2_ratio(const float* h_des_q, int nq, const float* h_des_t, int nt, double ratio,
                                 int* h_q, int* h_t, float* h_dist, int* m_out);
/* cv2.findHomography(src, dst, cv2.RANSAC, thresh) (maxIters, confidence explicit)  main.py:856-857.
 * h_src/h_dst: n x 2 float32.  *ok = 0 means the reference would get None. */
bm_status bm_ransac_homography(const float* h_src, const float* h_dst, int n, double thresh, int max_iters,
                               double confidence, double H[9], int* ok, int* iters, int* n_inliers);
/* profiling variant: also returns SM cycle counts per phase (subsets, hypotheses, selection, refit sums, DLT, LM, total, -), the number of
 * LM iterations and, in *jacobi_sweeps, bits 8-15: how many of them solved their system by eigen-decomposition (see below) */
bm_status bm_ransac_profile(const float* h_src, const float* h_dst, int n, double thresh, int max_iters, double confidence,
                            double H[9], long long cycles[8], int* lm_iters, int* jacobi_sweeps);
/* debug / parity tests: cv2 4.13's LM polish (nine parameters, cv::solve / cv::invert with DECOMP_EIG = eigen-decomposition + truncated
 * back substitution) is evaluated through one SPD elimination per iteration whenever a bound proves that only the scale gauge is
 * truncated, and through the literal eigen-decomposition otherwise; on != 0 forces the literal route for every iteration */
bm_status bm_debug_lm_force_eig(int on);
/* counters of the current CUDA device since the last reset (synchronises the device): out[0] polishes run, out[1] their LM iterations,
 * out[2] the iterations among them that were solved by eigen-decomposition */
bm_status bm_debug_lm_stats(unsigned long long out[3], int reset);
/* features / matches of the handle's last frame: which = 0 -> prev, 1 -> cur.  h_desc: uint8 n x 32 (ORB) or n x 128 (SIFT) */
bm_status bm_get_keypoints(bm_handle h, int which, float* h_kp, uint8_t* h_desc, int cap, int* n_out);
bm_status bm_get_matches(bm_handle h, int* h_q, int* h_t, float* h_dist, int cap, int* m_out);
int bm_keypoint_capacity(void);
/* debug / parity: level `level` of the ORB pyramid (INTER_LINEAR_EXACT chain) and its FAST score map (either may be NULL) */
bm_status bm_orb_debug_level(const uint8_t* d_gray, int h, int w, int level, uint8_t* h_img, uint8_t* h_score, int* lw, int* lh);
/* measurement: the SIFT matcher of main.py:687-698 (tensor-core kNN + merge + ratio + sort) on nq x nt random descriptors, CUDA-event
 * timed on its launching stream; *flops_per_pair = 2 * nq * nt * 128 (the dense contraction north_star asks the tensor-pipe figure for) */
bm_status bm_match_l2_ms(int nq, int nt, int reps, double* ms_per_pair, double* flops_per_pair);
/* measurement: the SIFT Gaussian + DoG pyramid kernels alone (inside detectAndCompute, main.py:718), `reps` times between two CUDA
 * events on their launching stream; *algorithmic_bytes = 256 * h * w (six f32 levels per octave written and read once) */
bm_status bm_sift_pyramid_ms(const uint8_t* d_gray, int h, int w, int reps, double* ms_per_frame, double* algorithmic_bytes);
/* debug / parity: one Gaussian (dog=0, level 0..5) or DoG (dog=1, level 0..4) image of the SIFT pyramid; returns the number of octaves in *noct */
bm_status bm_sift_debug_level(const uint8_t* d_gray, int h, int w, int octave, int level, int dog, float* h_out, int* lw, int* lh, int* noct);

#ifdef __cplusplus
}
#endif
#endif /* B200MOSAIC_H */
