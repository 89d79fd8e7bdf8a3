// warp_blend.cu -- VideMosaic.warp on the device (reference: /root/reference/main.py:861-927).
//
//   warped = cv2.warpPerspective(frame, H, canvas_size, INTER_LINEAR)            main.py:871
//   mask_new / mask_old / overlap                                               main.py:878-882
//   distanceTransform x2, weight normalisation, GaussianBlur(31) x2, blend      main.py:885-924
//   else-branch channel-wise overwrite                                          main.py:925-927
//
// B200 design (not a translation of OpenCV's CPU code):
//  * The reference does ~15 full-canvas passes per frame.  Here all work is confined to the window W (clipped
//    bounding box of the warped quad) and R = W (+) 15 px; the only canvas-global quantity, the chamfer distance to
//    the nearest uncovered canvas pixel, is obtained exactly from a PERSISTENT per-row structure: g_old(x,y) =
//    horizontal distance to the nearest zero pixel of row y, refreshed only for rows the frame touched, plus a 16-row
//    block-min table used to prune the vertical search.  D(x,y) = min_y' N(g(x,y'), |y-y'|) with
//    N(u,v) = a*max(u,v) + (b-a)*min(u,v) is the closed form of OpenCV's 3x3 integer chamfer.
//  * canvas is stored as uchar4 (B,G,R,mask) so every access is a coalesced 32-bit word and mask_old is free.
//  * integer / fixed-point arithmetic of cv2 is reproduced bit for bit (INTER_BITS=5 weights, 64-column block
//    evaluation of the homography in double without FMA contraction, 16.16 chamfer, float32 weights with the FMA
//    order OpenCV's AVX2 sepFilter2D uses).
#include "warp_blend.cuh"
#include <math.h>
#include <string.h>

// cv::getGaussianKernel(31, 5.0, CV_32F)  (sigma = 0.3*((31-1)*0.5-1)+0.8), taps 0..15; tap 30-k == tap k.
__constant__ float c_gk[16] = {
    8.880585083e-04f, 1.586106606e-03f, 2.721769968e-03f, 4.487439990e-03f, 7.108436897e-03f, 1.081876736e-02f,
    1.582011767e-02f, 2.222643606e-02f, 3.000254929e-02f, 3.891120851e-02f, 4.848635197e-02f, 5.804870278e-02f,
    6.677190214e-02f, 7.379436493e-02f, 7.835755497e-02f, 7.994048297e-02f};

// ------------------------------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------------------------------
void bm_invert3x3(const double a[9], double t[9]) {
    // closed-form cofactor inverse in double (what cv::invert does for 3x3; main.py:871 passes H, cv2 inverts it)
    double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    if (det == 0.0) { for (int i = 0; i < 9; ++i) t[i] = 0.0; return; }
    double d = 1.0 / det;
    t[0] = (a[4] * a[8] - a[5] * a[7]) * d;
    t[1] = (a[2] * a[7] - a[1] * a[8]) * d;
    t[2] = (a[1] * a[5] - a[2] * a[4]) * d;
    t[3] = (a[5] * a[6] - a[3] * a[8]) * d;
    t[4] = (a[0] * a[8] - a[2] * a[6]) * d;
    t[5] = (a[2] * a[3] - a[0] * a[5]) * d;
    t[6] = (a[3] * a[7] - a[4] * a[6]) * d;
    t[7] = (a[1] * a[6] - a[0] * a[7]) * d;
    t[8] = (a[0] * a[4] - a[1] * a[3]) * d;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

void bm_make_plan(const double H[9], int src_w, int src_h, int canvas_w, int canvas_h, BmFramePlan* p) {
    memset(p, 0, sizeof(*p));
    bm_invert3x3(H, p->M);
    p->src_w = src_w; p->src_h = src_h; p->canvas_w = canvas_w; p->canvas_h = canvas_h;
    {   // OpenCV warpPerspectiveInvoker block geometry: BLOCK_SZ = 32
        int bh0 = canvas_h < 16 ? canvas_h : 16;
        int bw0 = 1024 / (bh0 > 0 ? bh0 : 1);
        p->block_w = bw0 < canvas_w ? bw0 : canvas_w;
        if (p->block_w < 1) p->block_w = 1;
    }
    // forward-map the rectangle of source coordinates that can produce a non-zero bilinear sample: (-1,W) x (-1,H)
    const double cx[4] = {-1.0, (double)src_w, (double)src_w, -1.0};
    const double cy[4] = {-1.0, -1.0, (double)src_h, (double)src_h};
    double minx = 1e300, miny = 1e300, maxx = -1e300, maxy = -1e300;
    bool full = false;
    double wmin = 1e300, wmax = -1e300;
    for (int i = 0; i < 4; ++i) {
        double w = H[6] * cx[i] + H[7] * cy[i] + H[8];
        wmin = fmin(wmin, w); wmax = fmax(wmax, w);
    }
    bool finite = true;
    for (int i = 0; i < 9; ++i) if (!isfinite(H[i])) finite = false;
    if (!finite || wmin * wmax <= 0.0 || fabs(wmin) < 1e-9 * fabs(wmax)) full = true;   // horizon crosses the frame
    if (!full) {
        for (int i = 0; i < 4; ++i) {
            double w = H[6] * cx[i] + H[7] * cy[i] + H[8];
            double x = (H[0] * cx[i] + H[1] * cy[i] + H[2]) / w;
            double y = (H[3] * cx[i] + H[4] * cy[i] + H[5]) / w;
            minx = fmin(minx, x); maxx = fmax(maxx, x); miny = fmin(miny, y); maxy = fmax(maxy, y);
        }
        if (!(isfinite(minx) && isfinite(maxx) && isfinite(miny) && isfinite(maxy))) full = true;
    }
    BmWin w;
    if (full) { w.x0 = 0; w.y0 = 0; w.x1 = canvas_w; w.y1 = canvas_h; }
    else {
        // +-3: 1 px for the zero ring the DT of mask_new relies on, 2 px of slack for rounding
        double fx0 = floor(minx) - 3.0, fy0 = floor(miny) - 3.0, fx1 = ceil(maxx) + 4.0, fy1 = ceil(maxy) + 4.0;
        w.x0 = (int)fmax(0.0, fmin((double)canvas_w, fx0));
        w.y0 = (int)fmax(0.0, fmin((double)canvas_h, fy0));
        w.x1 = (int)fmax(0.0, fmin((double)canvas_w, fx1));
        w.y1 = (int)fmax(0.0, fmin((double)canvas_h, fy1));
    }
    w.y0 = (w.y0 / BM_BLK_ROWS) * BM_BLK_ROWS;      // window rows start on the 16-row block grid (k_dt_weights shares one grid for both masks)
    p->win = w;
    p->reg.x0 = clampi(w.x0 - BM_BLUR_R, 0, canvas_w); p->reg.x1 = clampi(w.x1 + BM_BLUR_R, 0, canvas_w);
    p->reg.y0 = clampi(w.y0 - BM_BLUR_R, 0, canvas_h); p->reg.y1 = clampi(w.y1 + BM_BLUR_R, 0, canvas_h);
    p->valid = (w.x1 > w.x0 && w.y1 > w.y0) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------------------------
// warpPerspective INTER_LINEAR: exact fixed-point coordinates and weights
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_coords(const double* __restrict__ M, int block_w, int x, int y, int& X, int& Y) {
    const int bx = (x / block_w) * block_w;
    const double x1 = (double)(x - bx), dbx = (double)bx, dy = (double)y;
    // no FMA contraction: OpenCV evaluates (M0*bx + M1*y) + M2 once per 64-column block, then adds M0*x1
    const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(M[0], dbx), __dmul_rn(M[1], dy)), M[2]);
    const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(M[3], dbx), __dmul_rn(M[4], dy)), M[5]);
    const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(M[6], dbx), __dmul_rn(M[7], dy)), M[8]);
    double W = __dadd_rn(W0, __dmul_rn(M[6], x1));
    W = (W != 0.0) ? __ddiv_rn(32.0, W) : 0.0;
    double fX = __dmul_rn(__dadd_rn(X0, __dmul_rn(M[0], x1)), W);
    double fY = __dmul_rn(__dadd_rn(Y0, __dmul_rn(M[3], x1)), W);
    fX = fmax(-2147483648.0, fmin(2147483647.0, fX));
    fY = fmax(-2147483648.0, fmin(2147483647.0, fY));
    X = __double2int_rn(fX);       // cvRound: round half to even
    Y = __double2int_rn(fY);
}

struct SrcBGRX {   // source stored as uchar4 words
    const uchar4* p; int w, h;
    __device__ __forceinline__ uchar4 at(int y, int x) const { return __ldg(p + (size_t)y * w + x); }
};
struct SrcBGR {    // source stored as packed 3-byte pixels
    const uint8_t* p; int w, h;
    __device__ __forceinline__ uchar4 at(int y, int x) const {
        const uint8_t* q = p + ((size_t)y * w + x) * 3;
        return make_uchar4(__ldg(q), __ldg(q + 1), __ldg(q + 2), 0);
    }
};

template <class Src>
__device__ __forceinline__ uchar4 warp_sample(const Src& s, int X, int Y) {
    int sx = X >> 5, sy = Y >> 5;
    sx = max(-32768, min(32767, sx));     // OpenCV keeps integer source coordinates as saturated int16
    sy = max(-32768, min(32767, sy));
    const int ax = X & 31, ay = Y & 31;
    const int w00 = (32 - ax) * (32 - ay), w01 = ax * (32 - ay), w10 = (32 - ax) * ay, w11 = ax * ay;
    uchar4 p00 = make_uchar4(0, 0, 0, 0), p01 = p00, p10 = p00, p11 = p00;
    if ((unsigned)sx < (unsigned)(s.w - 1) && (unsigned)sy < (unsigned)(s.h - 1)) {
        p00 = s.at(sy, sx); p01 = s.at(sy, sx + 1); p10 = s.at(sy + 1, sx); p11 = s.at(sy + 1, sx + 1);
    } else {
        if (sx < -1 || sy < -1 || sx >= s.w || sy >= s.h) return make_uchar4(0, 0, 0, 0);
        const bool x0ok = sx >= 0, x1ok = sx + 1 < s.w, y0ok = sy >= 0, y1ok = sy + 1 < s.h;
        if (y0ok && x0ok) p00 = s.at(sy, sx);
        if (y0ok && x1ok) p01 = s.at(sy, sx + 1);
        if (y1ok && x0ok) p10 = s.at(sy + 1, sx);
        if (y1ok && x1ok) p11 = s.at(sy + 1, sx + 1);
    }
    uchar4 o;
    o.x = (unsigned char)((p00.x * w00 + p01.x * w01 + p10.x * w10 + p11.x * w11 + 512) >> 10);
    o.y = (unsigned char)((p00.y * w00 + p01.y * w01 + p10.y * w10 + p11.y * w11 + 512) >> 10);
    o.z = (unsigned char)((p00.z * w00 + p01.z * w01 + p10.z * w10 + p11.z * w11 + 512) >> 10);
    o.w = (o.x | o.y | o.z) ? 255 : 0;
    return o;
}

// K1: warp the frame into the window scratch, record mask_new, detect np.any(overlap)
__global__ void __launch_bounds__(256) k_warp_window(const uchar4* __restrict__ src, const BmFramePlan* __restrict__ planp,
                                                     const uchar4* __restrict__ canvas, uchar4* __restrict__ wbuf,
                                                     int* __restrict__ flags) {
    __shared__ BmFramePlan plan;
    if (threadIdx.x == 0 && threadIdx.y == 0) plan = *planp;
    __syncthreads();
    const int ww = bm_win_w(plan.win), wh = bm_win_h(plan.win);
    const int lx = blockIdx.x * blockDim.x + threadIdx.x, ly = blockIdx.y * blockDim.y + threadIdx.y;
    bool ov = false;
    if (lx < ww && ly < wh) {
        const int x = plan.win.x0 + lx, y = plan.win.y0 + ly;
        int X, Y;
        warp_coords(plan.M, plan.block_w, x, y, X, Y);
        SrcBGRX s{src, plan.src_w, plan.src_h};
        const uchar4 o = warp_sample(s, X, Y);
        wbuf[(size_t)ly * ww + lx] = o;
        if (o.w) ov = canvas[(size_t)y * plan.canvas_w + x].w != 0;
    }
    if (__syncthreads_or(ov) && threadIdx.x == 0 && threadIdx.y == 0) atomicOr(flags, 1);
}

// stage entry (parity vs cv2.warpPerspective): full-canvas packed BGR output; dst is pre-zeroed by the launcher
__global__ void __launch_bounds__(256) k_warp_full_bgr(const uint8_t* __restrict__ src, BmFramePlan plan, uint8_t* __restrict__ dst) {
    const int lx = blockIdx.x * blockDim.x + threadIdx.x, ly = blockIdx.y * blockDim.y + threadIdx.y;
    if (lx >= bm_win_w(plan.win) || ly >= bm_win_h(plan.win)) return;
    const int x = plan.win.x0 + lx, y = plan.win.y0 + ly;
    int X, Y;
    warp_coords(plan.M, plan.block_w, x, y, X, Y);
    SrcBGR s{src, plan.src_w, plan.src_h};
    const uchar4 o = warp_sample(s, X, Y);
    uint8_t* q = dst + ((size_t)y * plan.canvas_w + x) * 3;
    q[0] = o.x; q[1] = o.y; q[2] = o.z;
}

// ------------------------------------------------------------------------------------------------------------------
// row scan: g(x) = distance to the nearest pixel of the row whose mask byte is 0 (one warp per row)
// img is addressed as img[(row0 + r) * stride + col0 + i], i in [0,n); g is addressed g[r_out * n + i]
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rowscan(const uchar4* __restrict__ img, int stride, int col0, int n, int row0, int nrows,
                                                 uint16_t* __restrict__ g, int g_row0, const int* __restrict__ flags, int need_flag) {
    // one CTA per row; warp w owns the contiguous chunk [w*chunk, (w+1)*chunk).  Phase A: first / last zero of every chunk;
    // phase B: forward (nearest zero at or before x) and backward (at or after x) ballot scans inside the chunk with the
    // carries of the neighbouring chunks.  3 passes over n/8 pixels per warp instead of 2 passes over n.
    if (need_flag && flags[0] == 0) return;
    __shared__ int s_first[8], s_last[8];
    const int r = blockIdx.x;
    if (r >= nrows) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uchar4* row = img + (size_t)(row0 + r) * stride + col0;
    uint16_t* grow = g + (size_t)(g_row0 + r) * n;
    const int chunk = (((n + 7) / 8) + 31) & ~31;
    const int c0 = warp * chunk, c1 = min(n, c0 + chunk);
    const int NOZ = -(1 << 28), FAR = 1 << 28;
    int first = FAR, last = NOZ;
    for (int c = c0; c < c1; c += 32) {
        const int x = c + lane;
        const unsigned b = __ballot_sync(0xffffffffu, (x < c1) && (row[x].w == 0));
        if (b) { if (first == FAR) first = c + __ffs(b) - 1; last = c + 31 - __clz(b); }
    }
    if (lane == 0) { s_first[warp] = first; s_last[warp] = last; }
    __syncthreads();
    int carry = NOZ;
    for (int w = 0; w < warp; ++w) carry = max(carry, s_last[w]);
    for (int c = c0; c < c1; c += 32) {
        const int x = c + lane;
        const unsigned b = __ballot_sync(0xffffffffu, (x < c1) && (row[x].w == 0));
        const unsigned m = b & (0xffffffffu >> (31 - lane));
        const int lastz = m ? (c + 31 - __clz(m)) : carry;
        if (x < c1) grow[x] = (uint16_t)min(x - lastz, (int)BM_G_INF);
        if (b) carry = c + 31 - __clz(b);
    }
    carry = FAR;
    for (int w = 7; w > warp; --w) carry = min(carry, s_first[w]);
    for (int c = c0 + ((max(c1 - c0, 1) - 1) / 32) * 32; c >= c0; c -= 32) {
        const int x = c + lane;
        const unsigned b = __ballot_sync(0xffffffffu, (x < c1) && (row[x].w == 0));
        const unsigned m = b & (0xffffffffu << lane);
        const int nextz = m ? (c + __ffs(m) - 1) : carry;
        if (x < c1) grow[x] = (uint16_t)min((int)grow[x], min(nextz - x, (int)BM_G_INF));
        if (b) carry = c + __ffs(b) - 1;
    }
}

// same scan for a plain u8 mask (stage entry bm_distance_transform)
__global__ void __launch_bounds__(256) k_rowscan_u8(const uint8_t* __restrict__ mask, int n, int nrows, uint16_t* __restrict__ g) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nrows) return;
    const uint8_t* row = mask + (size_t)warp * n;
    uint16_t* grow = g + (size_t)warp * n;
    int carry = -(1 << 28);
    for (int c = 0; c < n; c += 32) {
        const int x = c + lane;
        const bool z = (x < n) && (row[x] == 0);
        const unsigned b = __ballot_sync(0xffffffffu, z);
        const unsigned m = b & (0xffffffffu >> (31 - lane));
        const int lastz = m ? (c + 31 - __clz(m)) : carry;
        if (x < n) grow[x] = (uint16_t)min(x - lastz, (int)BM_G_INF);
        if (b) carry = c + 31 - __clz(b);
    }
    carry = 1 << 28;
    for (int c = ((n - 1) / 32) * 32; c >= 0; c -= 32) {
        const int x = c + lane;
        const bool z = (x < n) && (row[x] == 0);
        const unsigned b = __ballot_sync(0xffffffffu, z);
        const unsigned m = b & (0xffffffffu << lane);
        const int nextz = m ? (c + __ffs(m) - 1) : carry;
        if (x < n) grow[x] = (uint16_t)min((int)grow[x], min(nextz - x, (int)BM_G_INF));
        if (b) carry = c + __ffs(b) - 1;
    }
}

// block-min table: gblk[Y*n + x] = min over rows [16Y, 16Y+16) of g[row*n + x], for blocks Y in [Y0, Y1)
__global__ void __launch_bounds__(256) k_blockmin(const uint16_t* __restrict__ g, int n, int nrows, int Y0, int Y1,
                                                  uint16_t* __restrict__ gblk, const int* __restrict__ flags, int need_flag) {
    if (need_flag && flags[0] == 0) return;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int Y = Y0 + blockIdx.y;
    if (x >= n || Y >= Y1) return;
    const int r0 = Y * BM_BLK_ROWS, r1 = min(nrows, r0 + BM_BLK_ROWS);
    int m = BM_G_INF;
    for (int r = r0; r < r1; ++r) m = min(m, (int)g[(size_t)r * n + x]);
    gblk[(size_t)Y * n + x] = (uint16_t)m;
}

// ------------------------------------------------------------------------------------------------------------------
// exact chamfer distance from the row structure (SURVEY A.9 decomposition)
// ------------------------------------------------------------------------------------------------------------------
// ---- block-parallel exact search: one thread owns a 16-row run of one column -------------------------------------
// N(u,v) = a*max(u,v) + (b-a)*min(u,v) = max(a*u + (b-a)*v, (b-a)*u + a*v) because a >= b-a.
#define BM_U_CAP 16000     // a*16000 > DIST_MAX: "no zero reachable"; keeps every product inside int32
__device__ __forceinline__ int ncost(int u, int v) {
    return max(BM_CHAMFER_A * u + (BM_CHAMFER_B - BM_CHAMFER_A) * v, (BM_CHAMFER_B - BM_CHAMFER_A) * u + BM_CHAMFER_A * v);
}

// visit one 16-row block of candidate rows for the 16 pixels of the thread.  voff: v(i, r) = voff + SI*i + SR*r
template <int SI, int SR>
__device__ __forceinline__ void scan_block16(const uint16_t* __restrict__ g, int n, int nrows, int x, int r0, int voff, int vmin_base,
                                             int bmax, int (&best)[BM_BLK_ROWS]) {
    int gr[BM_BLK_ROWS];
#pragma unroll
    for (int r = 0; r < BM_BLK_ROWS; ++r) gr[r] = (r0 + r < nrows) ? min((int)__ldg(&g[(size_t)(r0 + r) * n + x]), BM_U_CAP) : BM_U_CAP;
#pragma unroll
    for (int r = 0; r < BM_BLK_ROWS; ++r) {
        const int u = gr[r];
        // smallest vertical distance from this row to any of the 16 pixels
        const int vmin = vmin_base + (SR < 0 ? (BM_BLK_ROWS - 1 - r) : r);
        if (ncost(u, vmin) < bmax) {
#pragma unroll
            for (int i = 0; i < BM_BLK_ROWS; ++i) best[i] = min(best[i], ncost(u, voff + SI * i + SR * r));
        }
    }
}

__device__ __forceinline__ void search16(const uint16_t* __restrict__ g, const uint16_t* __restrict__ gblk, int n, int nrows, int x, int Yb,
                                         int (&best)[BM_BLK_ROWS]) {
    const int nb = (nrows + BM_BLK_ROWS - 1) / BM_BLK_ROWS;
#pragma unroll
    for (int i = 0; i < BM_BLK_ROWS; ++i) best[i] = BM_DT_INIT;
    {   // own block: v = |r - i|
        int gr[BM_BLK_ROWS];
        const int r0 = Yb * BM_BLK_ROWS;
#pragma unroll
        for (int r = 0; r < BM_BLK_ROWS; ++r) gr[r] = (r0 + r < nrows) ? min((int)__ldg(&g[(size_t)(r0 + r) * n + x]), BM_U_CAP) : BM_U_CAP;
#pragma unroll
        for (int i = 0; i < BM_BLK_ROWS; ++i) best[i] = min(best[i], BM_CHAMFER_A * gr[i]);
        int bmax = 0;
#pragma unroll
        for (int i = 0; i < BM_BLK_ROWS; ++i) bmax = max(bmax, best[i]);
#pragma unroll
        for (int r = 0; r < BM_BLK_ROWS; ++r) {
            const int u = gr[r];
            if (BM_CHAMFER_A * u < bmax) {
#pragma unroll
                for (int i = 0; i < BM_BLK_ROWS; ++i) best[i] = min(best[i], ncost(u, i > r ? i - r : r - i));
            }
        }
    }
    // blocks above: rows r0..r0+15 with r0 = (Yb-k)*16; v(i,r) = 16k + i - r, nearest pair (i=0, r=15): 16k - 15
    for (int k = 1; Yb - k >= 0; ++k) {
        const int vnear = BM_BLK_ROWS * k - (BM_BLK_ROWS - 1);
        int bmax = 0;
#pragma unroll
        for (int i = 0; i < BM_BLK_ROWS; ++i) bmax = max(bmax, best[i]);
        if (vnear > 8578 || BM_CHAMFER_A * vnear >= bmax) break;
        const int gm = min((int)__ldg(&gblk[(size_t)(Yb - k) * n + x]), BM_U_CAP);
        if (ncost(gm, vnear) < bmax) scan_block16<1, -1>(g, n, nrows, x, (Yb - k) * BM_BLK_ROWS, BM_BLK_ROWS * k, vnear, bmax, best);
    }
    // blocks below: v(i,r) = 16k + r - i, nearest pair (i=15, r=0)
    for (int k = 1; Yb + k < nb; ++k) {
        const int vnear = BM_BLK_ROWS * k - (BM_BLK_ROWS - 1);
        int bmax = 0;
#pragma unroll
        for (int i = 0; i < BM_BLK_ROWS; ++i) bmax = max(bmax, best[i]);
        if (vnear > 8578 || BM_CHAMFER_A * vnear >= bmax) break;
        const int gm = min((int)__ldg(&gblk[(size_t)(Yb + k) * n + x]), BM_U_CAP);
        if (ncost(gm, vnear) < bmax) scan_block16<-1, 1>(g, n, nrows, x, (Yb + k) * BM_BLK_ROWS, BM_BLK_ROWS * k, vnear, bmax, best);
    }
}

// K3: over R, dn and do -> (dn/s, do/s) in float32 exactly as NumPy does it (main.py:892-894).
// thread = (column x, 16-row block Yb on the canvas block grid); plan.win.y0 is a multiple of 16 so the window-local
// block grid of g_new coincides with the canvas grid.
__global__ void __launch_bounds__(128) k_dt_weights(const BmFramePlan* __restrict__ planp, const uint16_t* __restrict__ g_old,
                                                    const uint16_t* __restrict__ gblk_old, const uint16_t* __restrict__ g_new,
                                                    const uint16_t* __restrict__ gblk_new, float2* __restrict__ rbuf,
                                                    const int* __restrict__ flags) {
    if (flags[0] == 0) return;
    __shared__ BmFramePlan plan;
    if (threadIdx.x == 0 && threadIdx.y == 0) plan = *planp;
    __syncthreads();
    const int rw = bm_win_w(plan.reg);
    const int lx = blockIdx.x * blockDim.x + threadIdx.x;
    const int Yb = plan.reg.y0 / BM_BLK_ROWS + blockIdx.y * blockDim.y + threadIdx.y;
    if (lx >= rw || Yb * BM_BLK_ROWS >= plan.reg.y1) return;
    const int x = plan.reg.x0 + lx;
    int bo[BM_BLK_ROWS], bn[BM_BLK_ROWS];
    search16(g_old, gblk_old, plan.canvas_w, plan.canvas_h, x, Yb, bo);
    const int xl = x - plan.win.x0, Ybl = Yb - plan.win.y0 / BM_BLK_ROWS;
    const int ww = bm_win_w(plan.win), wh = bm_win_h(plan.win);
    const bool in_win = xl >= 0 && xl < ww && Ybl >= 0 && Ybl * BM_BLK_ROWS < wh;
    if (in_win) search16(g_new, gblk_new, ww, wh, xl, Ybl, bn);
    const float scale = 1.0f / 65536.0f;
#pragma unroll
    for (int i = 0; i < BM_BLK_ROWS; ++i) {
        const int y = Yb * BM_BLK_ROWS + i;
        if (y < plan.reg.y0 || y >= plan.reg.y1) continue;
        const int d_new = (in_win && y < plan.win.y1) ? min(bn[i], BM_DT_INIT) : 0;
        const int d_old = min(bo[i], BM_DT_INIT);
        const float dn = __fmul_rn(__int2float_rn(d_new), scale);
        const float dold = __fmul_rn(__int2float_rn(d_old), scale);
        const float s = __fadd_rn(__fadd_rn(dn, dold), 1e-6f);
        rbuf[(size_t)(y - plan.reg.y0) * rw + lx] = make_float2(__fdiv_rn(dn, s), __fdiv_rn(dold, s));
    }
}

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

// K4: GaussianBlur(31) of both weight planes + blend + canvas update, one kernel.  CTA = 32x32 output pixels of W.
//  stage 0: (dn/s, do/s) tile with a 15 px halo (reflect-101 at the canvas border) -> shared memory
//  stage 1: row pass, 8 consecutive outputs per thread from a 38-value register window (cv2 order: tap 0 product, then
//           FMAs left to right)
//  stage 2: column pass, 4 consecutive outputs per thread from a 34-value register window (cv2's symmetric FMA form),
//           then the blend of main.py:905-927 and the canvas write.
#define FB_T 32
#define FB_S (FB_T + 2 * BM_BLUR_R)      // 62
__global__ void __launch_bounds__(256) k_blur_blend(const BmFramePlan* __restrict__ planp, const float2* __restrict__ rbuf,
                                                    const uchar4* __restrict__ wbuf, uchar4* __restrict__ canvas,
                                                    const int* __restrict__ flags) {
    __shared__ BmFramePlan plan;
    __shared__ float2 tile[FB_S][FB_S + 1];
    __shared__ float2 hrow[FB_S][FB_T + 1];
    const int tid = threadIdx.y * 32 + threadIdx.x;
    if (tid == 0) plan = *planp;
    __syncthreads();
    const int ww = bm_win_w(plan.win), wh = bm_win_h(plan.win);
    const int bx = plan.win.x0 + blockIdx.x * FB_T, by = plan.win.y0 + blockIdx.y * FB_T;     // canvas coords of the tile origin
    if (flags[0] == 0) {                                   // main.py:925-927: channel-wise overwrite, no weights needed
        for (int i = tid; i < FB_T * FB_T; i += 256) {
            const int x = bx + (i & 31), y = by + (i >> 5);
            if (x >= plan.win.x1 || y >= plan.win.y1) continue;
            const uchar4 w = wbuf[(size_t)(y - plan.win.y0) * ww + (x - plan.win.x0)];
            if (!w.w) continue;
            uchar4* cp = canvas + (size_t)y * plan.canvas_w + x;
            uchar4 c = *cp;
            if (w.x) c.x = w.x;
            if (w.y) c.y = w.y;
            if (w.z) c.z = w.z;
            c.w = 255;
            *cp = c;
        }
        return;
    }
    const int rw = bm_win_w(plan.reg);
    for (int i = tid; i < FB_S * FB_S; i += 256) {
        const int ty = i / FB_S, tx = i % FB_S;
        const int gx = reflect101(bx + tx - BM_BLUR_R, plan.canvas_w), gy = reflect101(by + ty - BM_BLUR_R, plan.canvas_h);
        float2 v = make_float2(0.f, 0.f);
        // pixels of the tile that lie beyond the window (only for partial edge tiles) are never used by valid outputs
        if (gx >= plan.reg.x0 && gx < plan.reg.x1 && gy >= plan.reg.y0 && gy < plan.reg.y1)
            v = rbuf[(size_t)(gy - plan.reg.y0) * rw + (gx - plan.reg.x0)];
        tile[ty][tx] = v;
    }
    __syncthreads();
    if (tid < FB_S * 4) {                                  // 62 rows x 4 segments of 8 outputs
        const int r = tid >> 2, c0 = (tid & 3) * 8;
        float ax[8], ay[8];
#pragma unroll
        for (int t = 0; t < 38; ++t) {
            const float2 v = tile[r][c0 + t];
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                const int k = t - o;
                if (k == 0) { ax[o] = __fmul_rn(v.x, c_gk[0]); ay[o] = __fmul_rn(v.y, c_gk[0]); }
                else if (k > 0 && k < 31) { const float g = c_gk[k < 16 ? k : 30 - k]; ax[o] = __fmaf_rn(v.x, g, ax[o]); ay[o] = __fmaf_rn(v.y, g, ay[o]); }
            }
        }
#pragma unroll
        for (int o = 0; o < 8; ++o) hrow[r][c0 + o] = make_float2(ax[o], ay[o]);
    }
    __syncthreads();
    {
        const int c = tid & 31, r0 = (tid >> 5) * 4;       // 32 columns x 8 groups of 4 rows
        float2 h[34];
#pragma unroll
        for (int t = 0; t < 34; ++t) h[t] = hrow[r0 + t][c];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            const int x = bx + c, y = by + r0 + o;
            if (x >= plan.win.x1 || y >= plan.win.y1) continue;
            const uchar4 w = wbuf[(size_t)(y - plan.win.y0) * ww + (x - plan.win.x0)];
            if (!w.w) continue;                            // canvas keeps its value where mask_new == 0
            uchar4* cp = canvas + (size_t)y * plan.canvas_w + x;
            const uchar4 cv = *cp;
            if (!cv.w) { *cp = w; continue; }              // non-overlap new: pixel copy (main.py:922-924)
            float wn = __fmul_rn(h[o + 15].x, c_gk[15]), wo = __fmul_rn(h[o + 15].y, c_gk[15]);
#pragma unroll
            for (int t = 1; t <= BM_BLUR_R; ++t) {
                wn = __fmaf_rn(__fadd_rn(h[o + 15 + t].x, h[o + 15 - t].x), c_gk[15 - t], wn);
                wo = __fmaf_rn(__fadd_rn(h[o + 15 + t].y, h[o + 15 - t].y), c_gk[15 - t], wo);
            }
            uchar4 ov;
            ov.x = (unsigned char)__float2int_rz(__fadd_rn(__fmul_rn((float)cv.x, wo), __fmul_rn((float)w.x, wn)));
            ov.y = (unsigned char)__float2int_rz(__fadd_rn(__fmul_rn((float)cv.y, wo), __fmul_rn((float)w.y, wn)));
            ov.z = (unsigned char)__float2int_rz(__fadd_rn(__fmul_rn((float)cv.z, wo), __fmul_rn((float)w.z, wn)));
            ov.w = (ov.x | ov.y | ov.z) ? 255 : 0;
            *cp = ov;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// format helpers
// ------------------------------------------------------------------------------------------------------------------
__global__ void k_pack_canvas(const uint8_t* __restrict__ bgr, uchar4* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t b = bgr[3 * (size_t)i], g = bgr[3 * (size_t)i + 1], r = bgr[3 * (size_t)i + 2];
    out[i] = make_uchar4(b, g, r, (b | g | r) ? 255 : 0);
}
__global__ void k_unpack_canvas(const uchar4* __restrict__ in, uint8_t* __restrict__ bgr, int n) {
    // 4 pixels (16 B in, 12 B out) per thread so both sides move whole words
    const int i4 = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = i4 * 4;
    if (i >= n) return;
    if (i + 3 < n && ((reinterpret_cast<uintptr_t>(bgr) & 3) == 0)) {
        const uint4 v = *reinterpret_cast<const uint4*>(in + i);
        const unsigned p0 = v.x & 0xffffff, p1 = v.y & 0xffffff, p2 = v.z & 0xffffff, p3 = v.w & 0xffffff;
        unsigned* o = reinterpret_cast<unsigned*>(bgr + 3 * (size_t)i);
        o[0] = p0 | (p1 << 24);
        o[1] = (p1 >> 8) | (p2 << 16);
        o[2] = (p2 >> 16) | (p3 << 8);
    } else {
        for (int k = i; k < n && k < i + 4; ++k) {
            const uchar4 c = in[k];
            bgr[3 * (size_t)k] = c.x; bgr[3 * (size_t)k + 1] = c.y; bgr[3 * (size_t)k + 2] = c.z;
        }
    }
}
__global__ void k_extract_wbuf(const uint8_t* __restrict__ warped, BmFramePlan plan, const uchar4* __restrict__ canvas,
                               uchar4* __restrict__ wbuf, int* __restrict__ flags) {
    const int ww = bm_win_w(plan.win), wh = bm_win_h(plan.win);
    const int lx = blockIdx.x * blockDim.x + threadIdx.x, ly = blockIdx.y * blockDim.y + threadIdx.y;
    bool ov = false;
    if (lx < ww && ly < wh) {
        const int x = plan.win.x0 + lx, y = plan.win.y0 + ly;
        const uint8_t* q = warped + ((size_t)y * plan.canvas_w + x) * 3;
        uchar4 o = make_uchar4(q[0], q[1], q[2], 0);
        o.w = (o.x | o.y | o.z) ? 255 : 0;
        wbuf[(size_t)ly * ww + lx] = o;
        if (o.w) ov = canvas[(size_t)y * plan.canvas_w + x].w != 0;
    }
    if (__syncthreads_or(ov) && threadIdx.x == 0 && threadIdx.y == 0) atomicOr(flags, 1);
}
__global__ void k_paste(uchar4* __restrict__ canvas, int canvas_w, const uchar4* __restrict__ src, int sw, int sh, int ox, int oy) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= sw || y >= sh) return;
    uchar4 v = src[(size_t)y * sw + x];
    v.w = (v.x | v.y | v.z) ? 255 : 0;
    canvas[(size_t)(oy + y) * canvas_w + ox + x] = v;
}
__global__ void __launch_bounds__(128) k_dt_from_g(const uint16_t* __restrict__ g, const uint16_t* __restrict__ gblk, int n, int nrows, float* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, Yb = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= n || Yb * BM_BLK_ROWS >= nrows) return;
    int best[BM_BLK_ROWS];
    search16(g, gblk, n, nrows, x, Yb, best);
#pragma unroll
    for (int i = 0; i < BM_BLK_ROWS; ++i) {
        const int y = Yb * BM_BLK_ROWS + i;
        if (y < nrows) out[(size_t)y * n + x] = __fmul_rn(__int2float_rn(min(best[i], BM_DT_INIT)), 1.0f / 65536.0f);
    }
}
__global__ void k_blur31_rows_plain(const float* __restrict__ in, int h, int w, float* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const float* row = in + (size_t)y * w;
    float a = __fmul_rn(row[reflect101(x - 15, w)], c_gk[0]);
#pragma unroll
    for (int k = 1; k < 31; ++k) a = __fmaf_rn(row[reflect101(x + k - 15, w)], c_gk[k < 16 ? k : 30 - k], a);
    out[(size_t)y * w + x] = a;
}
__global__ void k_blur31_cols_plain(const float* __restrict__ in, int h, int w, float* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    float a = __fmul_rn(in[(size_t)y * w + x], c_gk[15]);
#pragma unroll
    for (int t = 1; t <= 15; ++t)
        a = __fmaf_rn(__fadd_rn(in[(size_t)reflect101(y + t, h) * w + x], in[(size_t)reflect101(y - t, h) * w + x]), c_gk[15 - t], a);
    out[(size_t)y * w + x] = a;
}

// ------------------------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------------------------
static inline dim3 grid2(int w, int h, dim3 b) { return dim3((w + b.x - 1) / b.x, (h + b.y - 1) / b.y); }

cudaError_t bm_launch_full_rowscan(const BmBlendBufs& b, cudaStream_t s) {
    const int rows = b.canvas_h;
    BM_COUNT_LAUNCHES(1), k_rowscan<<<rows, 256, 0, s>>>(b.canvas, b.canvas_w, 0, b.canvas_w, 0, rows, b.g_old, 0, b.flags, 0);
    const int nb = bm_div_up(rows, BM_BLK_ROWS);
    BM_COUNT_LAUNCHES(1), k_blockmin<<<dim3(bm_div_up(b.canvas_w, 256), nb), 256, 0, s>>>(b.g_old, b.canvas_w, rows, 0, nb, b.gblk_old, b.flags, 0);
    return cudaGetLastError();
}

// chain after wbuf + flags[0] are filled
cudaError_t bm_launch_blend_from_wbuf(const BmBlendBufs& b, const BmFramePlan& plan, cudaStream_t s) {
    const int ww = bm_win_w(plan.win), wh = bm_win_h(plan.win);
    const int rw = bm_win_w(plan.reg), rh = bm_win_h(plan.reg);
    const dim3 blk(32, 8);
    // mask_new row structure over W (only needed when there is overlap: kernels exit early on flags[0]==0)
    BM_COUNT_LAUNCHES(1), k_rowscan<<<wh, 256, 0, s>>>(b.wbuf, ww, 0, ww, 0, wh, b.g_new, 0, b.flags, 1);
    const int nbw = bm_div_up(wh, BM_BLK_ROWS);
    BM_COUNT_LAUNCHES(1), k_blockmin<<<dim3(bm_div_up(ww, 256), nbw), 256, 0, s>>>(b.g_new, ww, wh, 0, nbw, b.gblk_new, b.flags, 1);
    {
        const int nyb = (plan.reg.y1 - 1) / BM_BLK_ROWS - plan.reg.y0 / BM_BLK_ROWS + 1;
        const dim3 b2(32, 4);
        BM_COUNT_LAUNCHES(1), k_dt_weights<<<dim3(bm_div_up(rw, 32), bm_div_up(nyb, 4)), b2, 0, s>>>(b.plan, b.g_old, b.gblk_old, b.g_new, b.gblk_new, b.rbuf, b.flags);
    }
    BM_COUNT_LAUNCHES(1), k_blur_blend<<<dim3(bm_div_up(ww, FB_T), bm_div_up(wh, FB_T)), blk, 0, s>>>(b.plan, b.rbuf, b.wbuf, b.canvas, b.flags);
    // refresh the persistent row structure for the rows the frame touched
    BM_COUNT_LAUNCHES(1), k_rowscan<<<wh, 256, 0, s>>>(b.canvas, b.canvas_w, 0, b.canvas_w, plan.win.y0, wh, b.g_old, plan.win.y0, b.flags, 0);
    const int Y0 = plan.win.y0 / BM_BLK_ROWS, Y1 = bm_div_up(plan.win.y1, BM_BLK_ROWS);
    BM_COUNT_LAUNCHES(1), k_blockmin<<<dim3(bm_div_up(b.canvas_w, 256), Y1 - Y0), 256, 0, s>>>(b.g_old, b.canvas_w, b.canvas_h, Y0, Y1, b.gblk_old, b.flags, 0);
    return cudaGetLastError();
}

cudaError_t bm_launch_warp_blend(const BmBlendBufs& b, const uchar4* d_src, const BmFramePlan& plan, cudaStream_t s) {
    if (!plan.valid) return cudaSuccess;
    cudaError_t e = cudaMemcpyAsync(b.plan, &plan, sizeof(plan), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(b.flags, 0, 4 * sizeof(int), s);
    if (e != cudaSuccess) return e;
    const dim3 blk(32, 8);
    BM_COUNT_LAUNCHES(1), k_warp_window<<<grid2(bm_win_w(plan.win), bm_win_h(plan.win), blk), blk, 0, s>>>(d_src, b.plan, b.canvas, b.wbuf, b.flags);
    return bm_launch_blend_from_wbuf(b, plan, s);
}

cudaError_t bm_launch_warp_full_bgr(const uint8_t* d_src, int sh, int sw, const BmFramePlan& plan, uint8_t* d_dst, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(d_dst, 0, (size_t)plan.canvas_w * plan.canvas_h * 3, s);
    if (e != cudaSuccess) return e;
    if (!plan.valid) return cudaSuccess;
    const dim3 blk(32, 8);
    BM_COUNT_LAUNCHES(1), k_warp_full_bgr<<<grid2(bm_win_w(plan.win), bm_win_h(plan.win), blk), blk, 0, s>>>(d_src, plan, d_dst);
    return cudaGetLastError();
}

cudaError_t bm_launch_pack_canvas(const uint8_t* d_bgr, uchar4* d_canvas, int n, cudaStream_t s) {
    BM_COUNT_LAUNCHES(1), k_pack_canvas<<<bm_div_up(n, 256), 256, 0, s>>>(d_bgr, d_canvas, n);
    return cudaGetLastError();
}
cudaError_t bm_launch_unpack_canvas(const uchar4* d_canvas, uint8_t* d_bgr, int n, cudaStream_t s) {
    BM_COUNT_LAUNCHES(1), k_unpack_canvas<<<bm_div_up(bm_div_up(n, 4), 256), 256, 0, s>>>(d_canvas, d_bgr, n);
    return cudaGetLastError();
}
cudaError_t bm_launch_extract_wbuf(const uint8_t* d_warped, const BmFramePlan& plan, const BmBlendBufs& b, cudaStream_t s) {
    cudaError_t e = cudaMemcpyAsync(b.plan, &plan, sizeof(plan), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(b.flags, 0, 4 * sizeof(int), s);
    if (e != cudaSuccess) return e;
    const dim3 blk(32, 8);
    BM_COUNT_LAUNCHES(1), k_extract_wbuf<<<grid2(bm_win_w(plan.win), bm_win_h(plan.win), blk), blk, 0, s>>>(d_warped, plan, b.canvas, b.wbuf, b.flags);
    return cudaGetLastError();
}
cudaError_t bm_launch_paste(uchar4* canvas, int canvas_w, const uchar4* src, int sw, int sh, int ox, int oy, cudaStream_t s) {
    const dim3 blk(32, 8);
    BM_COUNT_LAUNCHES(1), k_paste<<<grid2(sw, sh, blk), blk, 0, s>>>(canvas, canvas_w, src, sw, sh, ox, oy);
    return cudaGetLastError();
}
cudaError_t bm_launch_dt_mask(const uint8_t* d_mask, int h, int w, float* d_out, uint16_t* g, uint16_t* gblk, cudaStream_t s) {
    BM_COUNT_LAUNCHES(1), k_rowscan_u8<<<bm_div_up(h * 32, 256), 256, 0, s>>>(d_mask, w, h, g);
    const int nb = bm_div_up(h, BM_BLK_ROWS);
    BM_COUNT_LAUNCHES(1), k_blockmin<<<dim3(bm_div_up(w, 256), nb), 256, 0, s>>>(g, w, h, 0, nb, gblk, nullptr, 0);
    const dim3 blk(32, 8);
    BM_COUNT_LAUNCHES(1), k_dt_from_g<<<dim3(bm_div_up(w, 32), bm_div_up(bm_div_up(h, BM_BLK_ROWS), 4)), dim3(32, 4), 0, s>>>(g, gblk, w, h, d_out);
    return cudaGetLastError();
}
cudaError_t bm_launch_blur31(const float* d_in, int h, int w, float* d_tmp, float* d_out, cudaStream_t s) {
    const dim3 blk(32, 8);
    BM_COUNT_LAUNCHES(1), k_blur31_rows_plain<<<grid2(w, h, blk), blk, 0, s>>>(d_in, h, w, d_tmp);
    BM_COUNT_LAUNCHES(1), k_blur31_cols_plain<<<grid2(w, h, blk), blk, 0, s>>>(d_tmp, h, w, d_out);
    return cudaGetLastError();
}
