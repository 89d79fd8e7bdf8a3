// Launchers of the warp / distance-transform / blur / blend chain (warp_blend.cu).
#pragma once
#include "common.cuh"
#include "dt.cuh"

// Scratch + persistent buffers of one canvas.  All device memory.
struct BmBlendBufs {
    uchar4* canvas;        // BGR + mask byte (255 iff any channel > 0)       [canvas_h * canvas_w]
    BmDtPair dt;           // sweep tables: [0] canvas plane (mask_old, persistent), [1] window plane (mask_new, per frame)
    // per-frame scratch, sized for the largest window
    uchar4* wbuf;          // warped frame over W, .w = mask_new               [win_h][plan.ws]
    float2* wno;           // (dn/s, do/s) over R                               [reg_h][plan.rws]
    int* flags;            // [0] any_overlap
    size_t scratch_px;     // capacity of the scratch buffers in pixels (reg_h*reg_w <= scratch_px)
    int canvas_w, canvas_h;
};

// host-side: plan for one frame (window from H).  Returns plan.valid.
void bm_make_plan(const double H[9], int src_w, int src_h, int canvas_w, int canvas_h, BmFramePlan* plan);
void bm_invert3x3(const double H[9], double M[9]);

cudaError_t bm_launch_full_rowscan(const BmBlendBufs& b, cudaStream_t s);          // rebuild g_old for all rows
cudaError_t bm_launch_warp_blend(BmBlendBufs& b, const uchar4* d_src_bgrx, const BmFramePlan& plan,
                                 cudaStream_t s);                                    // whole per-frame chain
cudaError_t bm_launch_blend_from_wbuf(BmBlendBufs& b, const BmFramePlan& plan, cudaStream_t s);
cudaError_t bm_launch_warp_full_bgr(const uint8_t* d_src_bgr, int sh, int sw, const BmFramePlan& plan,
                                    uint8_t* d_dst_bgr, cudaStream_t s);
cudaError_t bm_launch_pack_canvas(const uint8_t* d_bgr, uchar4* d_canvas, int n_px, cudaStream_t s);
cudaError_t bm_launch_unpack_canvas(const uchar4* d_canvas, uint8_t* d_bgr, int n_px, cudaStream_t s);
cudaError_t bm_launch_extract_wbuf(const uint8_t* d_warped_bgr, const BmFramePlan& plan, const BmBlendBufs& b,
                                   cudaStream_t s);
cudaError_t bm_launch_paste(uchar4* canvas, int canvas_w, const uchar4* src, int sw, int sh, int ox, int oy,
                            cudaStream_t s);
cudaError_t bm_launch_blur31(const float* d_in, int h, int w, float* d_tmp, float* d_out, cudaStream_t s);
