// preview.cuh -- live-preview thumbnail of the device canvas (preview.cu)
#pragma once
#include "common.cuh"

struct BmPreviewPlan {
    int in_w = 0, in_h = 0, out_w = 0, out_h = 0;
    int kx = 0, ky = 0;                 // taps per output column / row
    int* d_tab = nullptr;               // [bounds_x 2*out_w][kk_x out_w*kx][bounds_y 2*out_h][kk_y out_h*ky]
    size_t tab_cap = 0;
    uchar4* d_tmp = nullptr;            // in_h x out_w, horizontally resampled (u8 like Pillow's intermediate image)
    size_t tmp_cap = 0;
    uint8_t* d_out = nullptr;           // out_h x out_w x 3
    size_t out_cap = 0;
};

// (re)builds the coefficient tables when the sizes change; host arithmetic in double exactly as Pillow's precompute_coeffs
cudaError_t bm_preview_prepare(BmPreviewPlan* p, int in_w, int in_h, int out_w, int out_h, cudaStream_t s);
cudaError_t bm_launch_preview(const BmPreviewPlan& p, const uchar4* canvas, int rgb, cudaStream_t s);
void bm_preview_free(BmPreviewPlan* p);
