// dt.cuh -- exact 3x3 chamfer distance transform (cv2.distanceTransform(mask, DIST_L2, 3), main.py:888-889) as
// separable min-plus sweeps.  See dt.cu for the derivation.
#pragma once
#include "common.cuh"

#define BM_DT_INF BM_DT_INIT                          // "no zero pixel reachable": cv2's DIST_MAX; INF + b == UINT_MAX, no wrap
#define BM_DT_TILE_VALID 96                           // a warp sweeps 128 columns, 16 on each side are halo
#define BM_ROWSCAN_CHUNK 2048                         // pixels per CTA pass of the row scan (256 threads x 8 px)
#define BM_ROWSCAN_MAX_CHUNKS 16                      // rows up to 32768 px

static inline __host__ __device__ int bm_pad4(int v) { return (v + 3) & ~3; }
static inline __host__ __device__ int bm_pad8(int v) { return (v + 7) & ~7; }

// Sweep tables of one mask plane.  "Plane" = the pixel grid the mask lives on: the canvas for mask_old, the window for
// mask_new.  All tables are indexed [table][block][column] with row stride ts; a block is BM_BLK_ROWS plane rows.
struct BmDtPlane {
    uint32_t* g;       // sweep seeds a * (horizontal distance to the nearest zero pixel of the row) in cv2's 16.16 units, BM_DT_INF
                       // if the row has no zero pixel and in the padding columns [W, gs): ready-made for the sweeps      [H][gs]
    int gs;            // row stride of g (multiple of 8)
    int W, H;          // plane size in pixels
    int nb;            // ceil(H / 16)
    int ts;            // row stride of the tables (multiple of 4)
    size_t tsz;        // elements per table (multiple of 4, >= nb * ts)
    uint32_t* LE;       // [4][tsz] block-local diagonal sweeps: 0/1 = downward (value at the block's last row) flowing
                       //          right / left, 2/3 = upward (value at the block's first row) flowing right / left
    uint32_t* CE;       // [4][tsz] the same with the carries of all blocks above / below
    uint32_t* CV;       // [2][tsz] vertical sweep at the block's last row (down) / first row (up); local, then with carries
    size_t g_cap;      // capacity of g in elements (for bm_dt_shape_plane)
    // Row tiles of a larger canvas: sweep state entering the plane from the rows above / below it, i.e. the
    // carries a neighbouring tile's plane holds at the adjoining block boundary.  [3][ts]: E1, E2, V.  nullptr = image border (no
    // zero pixel beyond: BM_DT_INF, cv2's semantics at the canvas edge).
    const uint32_t* gh_top;   // downward sweeps: values at the row just above row 0
    const uint32_t* gh_bot;   // upward sweeps: values at the row just below row H - 1
};

struct BmDtPair { BmDtPlane p[2]; };   // [0] = canvas (mask_old), [1] = window (mask_new)

cudaError_t bm_dt_alloc_plane(BmDtPlane* p, int Wcap, int Hcap, size_t px_cap);
void bm_dt_free_plane(BmDtPlane* p);
// set the live size of a plane whose buffers were allocated for a larger capacity; false if it does not fit
bool bm_dt_shape_plane(BmDtPlane* p, int W, int H);

// g rows [row0, row0+nrows) of a plane from a BGRX image (zero pixel <=> .w == 0); img is addressed
// img[(img_row0 + r) * img_stride + img_col0 + i], i in [0, p.W)
cudaError_t bm_launch_rowscan_bgrx(const uchar4* img, int img_stride, int img_col0, int img_row0, const BmDtPlane& p, int row0, int nrows,
                                   const int* flags, int need_flag, cudaStream_t s);
cudaError_t bm_launch_rowscan_u8(const uint8_t* mask, int stride, const BmDtPlane& p, cudaStream_t s);
// block-local diagonal sweeps (table LE) for blocks [kb0, kb1)
cudaError_t bm_launch_dt_local(const BmDtPlane& p, int kb0, int kb1, const int* flags, int need_flag, cudaStream_t s);
// diagonal carries + local vertical sweeps + vertical carries; column range [xa, xb) per plane (xa multiple of 4)
cudaError_t bm_launch_dt_carries(const BmDtPair& pp, int nplanes, const int xa[2], const int xb[2], const int* flags, int need_flag,
                                 cudaStream_t s);
// (dn/s, do/s) over R (main.py:888-894) -> one plane of (w_new, w_old) pairs with origin (plan.rx0, plan.reg.y0) and row stride plan.rws
cudaError_t bm_launch_dt_weights(const BmDtPair& pp, const BmFramePlan& plan, float2* wno, const int* flags, cudaStream_t s);
// canvas-plane carries over the full width, then the three carry rows (E1, E2, V) a neighbouring tile needs: downward sweeps at the
// last row of block `block` (up == 0) or upward sweeps at the first row of block `block` (up == 1) -> d_out [3][p.W]
cudaError_t bm_launch_dt_export_carries(const BmDtPlane& p, int up, int block, uint32_t* d_out, cudaStream_t s);
// plain distance map of one plane (stage entry bm_distance_transform)
cudaError_t bm_launch_dt_map(const BmDtPlane& p, float* d_out, cudaStream_t s);
