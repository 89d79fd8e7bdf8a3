// warp_blend.cu -- VideMosaic.warp on the device (reference: /root/reference/main.py:861-927).
//
//   warped = cv2.warpPerspective(frame, H, canvas_size, INTER_LINEAR)            main.py:871
//   mask_new / mask_old / overlap                                               main.py:878-882
//   distanceTransform x2, weight normalisation, GaussianBlur(31) x2, blend      main.py:885-924
//   else-branch channel-wise overwrite                                          main.py:925-927
//
// B200 design (not a translation of OpenCV's CPU code):
//  * The reference does ~15 full-canvas passes per frame.  Here all pixel work is confined to the window W (clipped
//    bounding box of the warped quad) and R = W (+) 15 px; the only canvas-global quantity, the chamfer distance to
//    the nearest uncovered canvas pixel, comes from the sweep tables of dt.cu: per-row nearest-zero distances g_old
//    and block-local diagonal sweeps are PERSISTENT and refreshed only for rows the frame touched; the per-line
//    carry chains over the canvas cost O(canvas / 16) per frame.
//  * canvas is stored as uchar4 (B,G,R,mask) so every access is a coalesced 32-bit word and mask_old is free.
//  * integer / fixed-point arithmetic of cv2 is reproduced bit for bit (INTER_BITS=5 weights, 64-column block
//    evaluation of the homography in double without FMA contraction, 16.16 chamfer, float32 weights with the FMA
//    order OpenCV's AVX2 sepFilter2D uses).
#include "warp_blend.cuh"
#include "rowscan.cuh"
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
    p->win = w;
    bm_finish_plan(p);
}

void bm_finish_plan(BmFramePlan* p) {
    BmWin& w = p->win;
    w.y0 = (w.y0 / BM_BLK_ROWS) * BM_BLK_ROWS;      // window rows start on the 16-row block grid (both masks share one grid)
    w.x0 &= ~7;                                     // 8-pixel groups of the row kernels never straddle a 64-column warp block
    p->reg.x0 = clampi(w.x0 - BM_BLUR_R, 0, p->canvas_w); p->reg.x1 = clampi(w.x1 + BM_BLUR_R, 0, p->canvas_w);
    p->reg.y0 = clampi(w.y0 - BM_BLUR_R, 0, p->canvas_h); p->reg.y1 = clampi(w.y1 + BM_BLUR_R, 0, p->canvas_h);
    p->valid = (w.x1 > w.x0 && w.y1 > w.y0) ? 1 : 0;
    p->ws = bm_pad8(bm_win_w(w));
    p->rx0 = p->reg.x0 & ~3;
    p->rws = bm_pad4(p->reg.x1 - p->rx0);
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

// border / outside pixels: per-tap path, kept out of line (rare; keeps the hot loop small)
__device__ __noinline__ uchar4 warp_sample_border(const uchar4* __restrict__ src, int w, int h, int X, int Y) {
    const SrcBGRX s{src, w, h};
    return warp_sample(s, X, Y);
}

// BGRX fast path: the 4 taps as 32-bit words, B and R interpolated together in 16-bit lanes.  Integer arithmetic is exact and
// distributive: sum_ij p_ij*wx_i*wy_j = wy0*(p00*wx0 + p01*wx1) + wy1*(p10*wx0 + p11*wx1); the horizontal sums are <= 255*32.
__device__ __forceinline__ uchar4 warp_sample_bgrx(const uchar4* __restrict__ src, int w, int h, int X, int Y) {
    int sx = X >> 5, sy = Y >> 5;
    sx = max(-32768, min(32767, sx));
    sy = max(-32768, min(32767, sy));
    if (!((unsigned)sx < (unsigned)(w - 1) && (unsigned)sy < (unsigned)(h - 1))) return warp_sample_border(src, w, h, X, Y);
    const unsigned ax = X & 31, ay = Y & 31, bxw = 32u - ax, byw = 32u - ay;
    const unsigned* p = reinterpret_cast<const unsigned*>(src) + (size_t)sy * w + sx;
    const unsigned a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + w), d = __ldg(p + w + 1);
    const unsigned h0br = (a & 0x00ff00ffu) * bxw + (b & 0x00ff00ffu) * ax, h1br = (c & 0x00ff00ffu) * bxw + (d & 0x00ff00ffu) * ax;
    const unsigned h0g = ((a >> 8) & 0xffu) * bxw + ((b >> 8) & 0xffu) * ax, h1g = ((c >> 8) & 0xffu) * bxw + ((d >> 8) & 0xffu) * ax;
    const unsigned B = ((h0br & 0xffffu) * byw + (h1br & 0xffffu) * ay + 512u) >> 10;
    const unsigned R = ((h0br >> 16) * byw + (h1br >> 16) * ay + 512u) >> 10;
    const unsigned G = (h0g * byw + h1g * ay + 512u) >> 10;
    uchar4 o;
    o.x = (unsigned char)B; o.y = (unsigned char)G; o.z = (unsigned char)R;
    o.w = (B | G | R) ? 255 : 0;
    return o;
}

// K1: one CTA per window row: warp the frame into the window scratch (8 consecutive pixels per thread: the homography
// terms of OpenCV's 64-column evaluation block are shared), detect np.any(overlap), and produce the nearest-zero row scan
// of mask_new from the mask bits still in registers.
__global__ void __launch_bounds__(256) k_warp_rows(const uchar4* __restrict__ src, BmFramePlan plan, const uchar4* __restrict__ canvas,
                                                   uchar4* __restrict__ wbuf, uint32_t* __restrict__ g_new, int gs, int* __restrict__ flags) {
    BM_PDL_TRIGGER();                                           // first kernel of the chain (follows a memset): nothing to wait for
    const int ly = blockIdx.x, tid = threadIdx.x;
    const int ww = bm_win_w(plan.win), y = plan.win.y0 + ly;
    const int nch = (ww + BM_ROWSCAN_CHUNK - 1) / BM_ROWSCAN_CHUNK;
    const uchar4* crow = canvas + (size_t)y * plan.canvas_w + plan.win.x0;
    uchar4* wrow = wbuf + (size_t)ly * plan.ws;
    const double* M = plan.M;
    const double dy = (double)y;
    BmZeroBits zb; zb.clear();
    bool ov = false;
    for (int c = 0; c < nch; ++c) {
        const int base = c * BM_ROWSCAN_CHUNK + 8 * tid;
        if (base >= ww) continue;
        const int x0 = plan.win.x0 + base;
        uchar4 o[8];
        {
            const int bx = (x0 / plan.block_w) * plan.block_w;
            const double dbx = (double)bx;
            const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(M[0], dbx), __dmul_rn(M[1], dy)), M[2]);
            const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(M[3], dbx), __dmul_rn(M[4], dy)), M[5]);
            const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(M[6], dbx), __dmul_rn(M[7], dy)), M[8]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const double x1 = (double)(x0 + i - bx);
                double W = __dadd_rn(W0, __dmul_rn(M[6], x1));
                W = (W != 0.0) ? __ddiv_rn(32.0, W) : 0.0;
                // cv2 clamps to [INT_MIN, INT_MAX] and then rounds (cvRound); cvt.rni.s32.f64 saturates to the same values
                const double fX = __dmul_rn(__dadd_rn(X0, __dmul_rn(M[0], x1)), W);
                const double fY = __dmul_rn(__dadd_rn(Y0, __dmul_rn(M[3], x1)), W);
                o[i] = warp_sample_bgrx(src, plan.src_w, plan.src_h, __double2int_rn(fX), __double2int_rn(fY));
            }
        }
        unsigned b = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (base + i >= ww) o[i] = make_uchar4(0, 0, 0, 0);
            else if (o[i].w == 0) b |= 1u << i;
            else ov |= crow[base + i].w != 0;
        }
        zb.set(c, b);
        unsigned pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pk[i] = (unsigned)o[i].x | ((unsigned)o[i].y << 8) | ((unsigned)o[i].z << 16) | ((unsigned)o[i].w << 24);
        uint4* wp = reinterpret_cast<uint4*>(wrow + base);
        wp[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        wp[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
    bm_rowscan_block(zb, nch, ww, g_new + (size_t)ly * gs);
    if (__syncthreads_or(ov) && tid == 0) atomicOr(flags, 1);
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

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

// K4: GaussianBlur(31) of both weight planes + blend + canvas update, one kernel.  CTA = 64 x 64 output pixels of W (512 threads).
// The two planes are blurred TOGETHER with Blackwell's packed fp32 instructions (FFMA2 / FADD2 / FMUL2, PTX fma.rn.f32x2): the tile
// holds (w_new, w_old) pairs, one LDS.64 fetches both and one FFMA2 applies the tap to both -- per lane the IEEE-rn result of the
// scalar instruction, so the arithmetic is unchanged, at half the FP and shared-memory instructions.  Tile with a 15 px halo
// (reflect-101 at the canvas border) -> shared memory, one 8-byte cp.async per pair (k_dt_weights writes the pairs interleaved); row pass, 16 consecutive outputs per thread from a 46-pair register
// window (cv2 order: tap 0 product, then FMAs left to right), written back IN PLACE over the tile (one extra barrier; 71 KB per CTA);
// column pass, 8 consecutive outputs per thread (cv2's symmetric FMA form).  Then the blend of main.py:905-927.
// Rows are 95 pairs apart: lanes walking rows hit distinct banks within each half warp of a 64-bit access; lanes walking columns are
// contiguous.
#define FB_TW 64
#define FB_TH 64
#define FB_NT 512                         // threads: 64 columns x 8 groups of 8 rows in the column pass
#define FB_SW (FB_TW + 2 * BM_BLUR_R)      // 94
#define FB_SH (FB_TH + 2 * BM_BLUR_R)      // 94
struct FbSmem {
    float2 tile[FB_SH][FB_SW + 1];       // (w_new, w_old) with halo; columns [0, 64) are overwritten by the row-filtered pairs
};
typedef unsigned long long fb_u64;
__constant__ float2 c_gk2[16] = {          // (k, k) pairs of c_gk: the multiplier operand of the packed instructions
    {8.880585083e-04f, 8.880585083e-04f}, {1.586106606e-03f, 1.586106606e-03f}, {2.721769968e-03f, 2.721769968e-03f}, {4.487439990e-03f, 4.487439990e-03f},
    {7.108436897e-03f, 7.108436897e-03f}, {1.081876736e-02f, 1.081876736e-02f}, {1.582011767e-02f, 1.582011767e-02f}, {2.222643606e-02f, 2.222643606e-02f},
    {3.000254929e-02f, 3.000254929e-02f}, {3.891120851e-02f, 3.891120851e-02f}, {4.848635197e-02f, 4.848635197e-02f}, {5.804870278e-02f, 5.804870278e-02f},
    {6.677190214e-02f, 6.677190214e-02f}, {7.379436493e-02f, 7.379436493e-02f}, {7.835755497e-02f, 7.835755497e-02f}, {7.994048297e-02f, 7.994048297e-02f}};
__device__ __forceinline__ fb_u64 fb_fma2(fb_u64 a, fb_u64 b, fb_u64 c) { fb_u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ fb_u64 fb_mul2(fb_u64 a, fb_u64 b) { fb_u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ fb_u64 fb_add2(fb_u64 a, fb_u64 b) { fb_u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ fb_u64 fb_gk2(int k) { return *reinterpret_cast<const fb_u64*>(&c_gk2[k]); }
__device__ __forceinline__ void cp_async8(float2* smem_dst, const float2* gmem_src, bool valid) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;        // src-size 0: the 8 bytes are zero-filled, nothing is read
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(gmem_src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One channel of the blend without the conversion unit (I2F.U8 / F2I.TRUNC run at a quarter of the FP rate and were the busiest
// pipe of the kernel): byte -> float as the exact difference (2^23 + b) - 2^23 with the byte dropped into the mantissa by one PRMT,
// float -> byte as the low mantissa bits of f + 2^23 rounded TOWARD ZERO (= 2^23 + floor(f) for 0 <= f < 2^23: the truncation of astype).
template <int CH>
__device__ __forceinline__ unsigned fb_blend_channel(unsigned cv, unsigned w, float wo, float wn) {
    const float fc = __fsub_rn(__uint_as_float(__byte_perm(cv, 0x4B000000u, 0x7440 | CH)), 8388608.0f);
    const float fw = __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440 | CH)), 8388608.0f);
    const float f = __fadd_rn(__fmul_rn(fc, wo), __fmul_rn(fw, wn));
    return __float_as_uint(__fadd_rz(f, 8388608.0f)) & 255u;
}

__global__ void __launch_bounds__(FB_NT, 2) k_blur_blend(BmFramePlan plan, const float2* __restrict__ wnop,
                                                       const uchar4* __restrict__ wbuf, uchar4* __restrict__ canvas,
                                                       const int* __restrict__ flags) {
    BM_PDL_TRIGGER(); BM_PDL_WAIT();
    extern __shared__ __align__(16) unsigned char fb_smem_raw[];
    FbSmem& sm = *reinterpret_cast<FbSmem*>(fb_smem_raw);
    const int tid = threadIdx.x;
    const int bx = plan.win.x0 + blockIdx.x * FB_TW, by = plan.win.y0 + blockIdx.y * FB_TH;     // canvas coords of the tile origin
    const int c = tid & (FB_TW - 1), r0 = (tid >> 6) * 8;       // column pass / blend: 64 columns x 8 groups of 8 rows
    const int x = bx + c;
    const bool xin = x < plan.win.x1;
    const uchar4* __restrict__ wcol = wbuf + (size_t)(by + r0 - plan.win.y0) * plan.ws + (x - plan.win.x0);
    uchar4* __restrict__ ccol = canvas + (size_t)(by + r0) * plan.canvas_w + x;
    const int nrow = xin ? min(8, plan.win.y1 - (by + r0)) : 0;     // rows of this thread inside the window
    if (flags[0] == 0) {                                   // main.py:925-927: channel-wise overwrite, no weights needed
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            if (o >= nrow) break;
            const uchar4 w = wcol[(size_t)o * plan.ws];
            if (!w.w) continue;
            uchar4 q = ccol[(size_t)o * plan.canvas_w];
            if (w.x) q.x = w.x;
            if (w.y) q.y = w.y;
            if (w.z) q.z = w.z;
            q.w = 255;
            ccol[(size_t)o * plan.canvas_w] = q;
        }
        return;
    }
    {
        const int warp = tid >> 5, lane = tid & 31;
        // both weight planes (tile + 15 px halo) are requested first: the copies fly while the overlap test below loads wbuf / canvas.
        // tile columns of this lane (3 per row) and their source columns, fixed for the whole tile
        int gxo[3]; bool gxok[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int tx = lane + 32 * j;
            const int gx = reflect101(bx + tx - BM_BLUR_R, plan.canvas_w);
            gxok[j] = gx >= plan.reg.x0 && gx < plan.reg.x1;     // beyond R only in partial edge tiles: never used
            gxo[j] = gxok[j] ? gx - plan.rx0 : 0;
        }
        for (int ty = warp; ty < FB_SH; ty += FB_NT / 32) {
            const int gy = reflect101(by + ty - BM_BLUR_R, plan.canvas_h);
            const bool yok = gy >= plan.reg.y0 && gy < plan.reg.y1;
            const size_t ro = yok ? (size_t)(gy - plan.reg.y0) * plan.rws : 0;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int tx = lane + 32 * j;
                if (tx < FB_SW) cp_async8(&sm.tile[ty][tx], wnop + ro + gxo[j], yok && gxok[j]);
            }
        }
        cp_async_commit();
    }
    // the thread's 8 warped and 8 canvas pixels: 16 independent loads issued together (a short-circuit test would chain them), kept
    // in registers for the blend at the end -- this CTA is the only writer of these canvas pixels
    unsigned wv[8], cv8[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        wv[o] = 0u; cv8[o] = 0u;
        if (o < nrow) {
            wv[o] = __ldg(reinterpret_cast<const unsigned*>(wcol + (size_t)o * plan.ws));
            cv8[o] = *reinterpret_cast<const unsigned*>(ccol + (size_t)o * plan.canvas_w);
        }
    }
    bool need = false;
#pragma unroll
    for (int o = 0; o < 8; ++o) need |= (wv[o] >> 24) && (cv8[o] >> 24);
    fb_u64 wgt[8];                                         // (w_new, w_old) of the thread's 8 pixels
    cp_async_wait<0>();                                    // nothing may be in flight into shared memory when the CTA exits
    if (__syncthreads_or(need)) {                          // tiles without a single overlap pixel need no weights; the tile is complete
        const bool rowt = tid < FB_SH * (FB_TW / 16);      // 94 rows x 4 segments of 16 outputs; lanes walk rows
        const int seg = tid / FB_SH, r = tid - seg * FB_SH, c0 = seg * 16;
        fb_u64 acc[16];
        if (rowt) {
            const fb_u64* __restrict__ trow = reinterpret_cast<const fb_u64*>(&sm.tile[r][c0]);
#pragma unroll
            for (int t = 0; t < 46; ++t) {
                const fb_u64 v = trow[t];
#pragma unroll
                for (int o = 0; o < 16; ++o) {
                    const int k = t - o;
                    if (k == 0) acc[o] = fb_mul2(v, fb_gk2(0));
                    else if (k > 0 && k < 31) acc[o] = fb_fma2(v, fb_gk2(k < 16 ? k : 30 - k), acc[o]);
                }
            }
        }
        __syncthreads();                                   // every window has been read: the row-filtered pairs go back in place
        if (rowt) {
            fb_u64* __restrict__ hrow = reinterpret_cast<fb_u64*>(&sm.tile[r][c0]);
#pragma unroll
            for (int o = 0; o < 16; ++o) hrow[o] = acc[o];
        }
        __syncthreads();
        fb_u64 h[38];
#pragma unroll
        for (int t = 0; t < 38; ++t) h[t] = *reinterpret_cast<const fb_u64*>(&sm.tile[r0 + t][c]);
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            fb_u64 a = fb_mul2(h[o + 15], fb_gk2(15));
#pragma unroll
            for (int t = 1; t <= BM_BLUR_R; ++t) a = fb_fma2(fb_add2(h[o + 15 + t], h[o + 15 - t]), fb_gk2(15 - t), a);
            wgt[o] = a;
        }
    } else {
#pragma unroll
        for (int o = 0; o < 8; ++o) wgt[o] = 0ull;
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
        if (o >= nrow) break;
        const unsigned w = wv[o], cv = cv8[o];
        if (!(w >> 24)) continue;                          // canvas keeps its value where mask_new == 0 (nrow rows only: wv = 0 beyond)
        unsigned* cp = reinterpret_cast<unsigned*>(ccol + (size_t)o * plan.canvas_w);
        if (!(cv >> 24)) { *cp = w; continue; }            // non-overlap new: pixel copy (main.py:922-924)
        const float wn = __uint_as_float((unsigned)wgt[o]), wo = __uint_as_float((unsigned)(wgt[o] >> 32));
        // float32(canvas)*w_old + float32(warped)*w_new, then astype(uint8) = truncation (main.py:905-910)
        const unsigned ox = fb_blend_channel<0>(cv, w, wo, wn), oy = fb_blend_channel<1>(cv, w, wo, wn), oz = fb_blend_channel<2>(cv, w, wo, wn);
        *cp = ox | (oy << 8) | (oz << 16) | ((ox | oy | oz) ? 0xff000000u : 0u);
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
        wbuf[(size_t)ly * plan.ws + lx] = o;
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

// g_old for all canvas rows + block-local sweep tables of all blocks
cudaError_t bm_launch_full_rowscan(const BmBlendBufs& b, cudaStream_t s) {
    const BmDtPlane& po = b.dt.p[0];
    cudaError_t e = bm_launch_rowscan_bgrx(b.canvas, b.canvas_w, 0, 0, po, 0, b.canvas_h, b.flags, 0, s);
    if (e != cudaSuccess) return e;
    return bm_launch_dt_local(po, 0, po.nb, b.flags, 0, s);
}

// chain after wbuf, g_new (rows of the window plane) and flags[0] are filled.  b.dt.p[1] is shaped for the window.
static cudaError_t blend_tail(const BmBlendBufs& b, const BmFramePlan& plan, cudaStream_t s) {
    cudaError_t attr;
    BM_SMEM_OPTIN(k_blur_blend, sizeof(FbSmem), attr);
    if (attr != cudaSuccess) return attr;
    const int ww = bm_win_w(plan.win), wh = bm_win_h(plan.win);
    const BmDtPlane& po = b.dt.p[0];
    const BmDtPlane& pn = b.dt.p[1];
    cudaError_t e;
    // everything up to the weights is only needed when there is overlap: the kernels exit early on flags[0] == 0
    if ((e = bm_launch_dt_local(pn, 0, pn.nb, b.flags, 1, s)) != cudaSuccess) return e;
    const int xa[2] = {plan.rx0, 0}, xb[2] = {plan.reg.x1, ww};
    if ((e = bm_launch_dt_carries(b.dt, 2, xa, xb, b.flags, 1, s)) != cudaSuccess) return e;
    if ((e = bm_launch_dt_weights(b.dt, plan, b.wno, b.flags, s)) != cudaSuccess) return e;
    BM_COUNT_LAUNCHES(1);
    if ((e = bm_launch_pdl(k_blur_blend, dim3(bm_div_up(ww, FB_TW), bm_div_up(wh, FB_TH)), dim3(FB_NT), sizeof(FbSmem), s, plan, (const float2*)b.wno,
                           (const uchar4*)b.wbuf, b.canvas, (const int*)b.flags)) != cudaSuccess) return e;
    // refresh the persistent tables of the canvas plane for the rows the frame touched
    if ((e = bm_launch_rowscan_bgrx(b.canvas, b.canvas_w, 0, plan.win.y0, po, plan.win.y0, wh, b.flags, 0, s)) != cudaSuccess) return e;
    return bm_launch_dt_local(po, plan.win.y0 / BM_BLK_ROWS, bm_div_up(plan.win.y1, BM_BLK_ROWS), b.flags, 0, s);
}

cudaError_t bm_launch_blend_from_wbuf(BmBlendBufs& b, const BmFramePlan& plan, cudaStream_t s) {
    const int ww = bm_win_w(plan.win), wh = bm_win_h(plan.win);
    if (!bm_dt_shape_plane(&b.dt.p[1], ww, wh)) return cudaErrorInvalidValue;
    cudaError_t e = bm_launch_rowscan_bgrx(b.wbuf, plan.ws, 0, 0, b.dt.p[1], 0, wh, b.flags, 1, s);
    if (e != cudaSuccess) return e;
    return blend_tail(b, plan, s);
}

cudaError_t bm_launch_warp_blend(BmBlendBufs& b, const uchar4* d_src, const BmFramePlan& plan, cudaStream_t s) {
    if (!plan.valid) return cudaSuccess;
    const int ww = bm_win_w(plan.win), wh = bm_win_h(plan.win);
    // (block_w is 64 for every canvas >= 64 px wide; the row kernel shares the block terms between 8 consecutive pixels)
    if ((plan.block_w & 7) || ww > BM_ROWSCAN_CHUNK * BM_ROWSCAN_MAX_CHUNKS || !bm_dt_shape_plane(&b.dt.p[1], ww, wh)) return cudaErrorInvalidValue;
    cudaError_t e = cudaMemsetAsync(b.flags, 0, 4 * sizeof(int), s);
    if (e != cudaSuccess) return e;
    BM_COUNT_LAUNCHES(1), k_warp_rows<<<wh, 256, 0, s>>>(d_src, plan, b.canvas, b.wbuf, b.dt.p[1].g, b.dt.p[1].gs, b.flags);
    return blend_tail(b, plan, s);
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
    cudaError_t e = cudaMemsetAsync(b.flags, 0, 4 * sizeof(int), s);
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
cudaError_t bm_launch_blur31(const float* d_in, int h, int w, float* d_tmp, float* d_out, cudaStream_t s) {
    const dim3 blk(32, 8);
    BM_COUNT_LAUNCHES(1), k_blur31_rows_plain<<<grid2(w, h, blk), blk, 0, s>>>(d_in, h, w, d_tmp);
    BM_COUNT_LAUNCHES(1), k_blur31_cols_plain<<<grid2(w, h, blk), blk, 0, s>>>(d_tmp, h, w, d_out);
    return cudaGetLastError();
}
