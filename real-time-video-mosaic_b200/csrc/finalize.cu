// finalize.cu -- mosaic finalisation on the device: crop_black_areas (/root/reference/main.py:980-1003)
// and scale_to_screen (:1006-1038) as main() applies them to the final canvas (:1647-1659).  Only the few-megabyte screen-sized
// result crosses PCIe instead of the whole canvas (3.2 GB in config 5).
//   k_crop_bounds   : BGR2GRAY + THRESH_BINARY + boundingRect(findNonZero) as one min/max reduction over the canvas
//   k_resize_linear : cv2.resize(INTER_LINEAR) of the cropped window, 8-bit fixed point exactly as OpenCV's SIMD path
//                     (weights = saturate_cast<short>(w * 2048), rows in int32, columns ((b*(S>>4))>>16 ... +2)>>2); exact 2x2
//                     decimation takes OpenCV's INTER_AREA route.  Arithmetic pinned in oracle/finalize.py.
#include "finalize.cuh"
#include <limits.h>

__global__ void k_bounds_init(int* b) { if (threadIdx.x < 4) b[threadIdx.x] = threadIdx.x < 2 ? INT_MAX : -1; }

__global__ void __launch_bounds__(256) k_crop_bounds(const uchar4* __restrict__ canvas, int w, int h, int thr, int* __restrict__ bounds) {
    __shared__ int s[4][8];
    int x0 = INT_MAX, y0 = INT_MAX, x1 = -1, y1 = -1;
    for (int y = blockIdx.y; y < h; y += gridDim.y) {
        const uchar4* row = canvas + (size_t)y * w;
        for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < w; x += gridDim.x * blockDim.x) {
            const uchar4 p = __ldg(row + x);
            const int gray = (3735 * p.x + 19235 * p.y + 9798 * p.z + 16384) >> 15;          // cv2 BGR2GRAY
            if (gray > thr) { x0 = min(x0, x); x1 = max(x1, x); y0 = min(y0, y); y1 = max(y1, y); }
        }
    }
    x0 = __reduce_min_sync(0xffffffffu, x0); y0 = __reduce_min_sync(0xffffffffu, y0);
    x1 = __reduce_max_sync(0xffffffffu, x1); y1 = __reduce_max_sync(0xffffffffu, y1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s[0][warp] = x0; s[1][warp] = y0; s[2][warp] = x1; s[3][warp] = y1; }
    __syncthreads();
    if (threadIdx.x == 0) {                                  // one atomic per CTA and bound
        for (int k = 1; k < 8; ++k) { x0 = min(x0, s[0][k]); y0 = min(y0, s[1][k]); x1 = max(x1, s[2][k]); y1 = max(y1, s[3][k]); }
        if (x1 >= 0) { atomicMin(&bounds[0], x0); atomicMin(&bounds[1], y0); atomicMax(&bounds[2], x1); atomicMax(&bounds[3], y1); }
    }
}

struct ResizeAxis { int s0, s1, a0, a1; };

// cv2's coefficient computation for one destination index (see oracle/finalize.py::_axis)
__device__ __forceinline__ ResizeAxis resize_axis(int d, double scale, int sn, bool vertical) {
    float f = __double2float_rn(__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5));
    int s = __float2int_rd(f);
    f = __fsub_rn(f, (float)s);
    if (!vertical) {
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= sn - 1) { s = sn - 1; f = 0.f; }
    }
    ResizeAxis r;
    r.a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    r.a1 = __float2int_rn(__fmul_rn(f, 2048.f));
    r.s0 = min(max(s, 0), sn - 1);
    r.s1 = min(max(s + 1, 0), sn - 1);
    return r;
}

__global__ void __launch_bounds__(256) k_resize_linear(const uchar4* __restrict__ canvas, int cw, int rx, int ry, int sw, int sh,
                                                       uint8_t* __restrict__ out, int dw, int dh, double scale_x, double scale_y, int area2) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y * blockDim.y + threadIdx.y;
    if (dx >= dw || dy >= dh) return;
    const uchar4* src = canvas + (size_t)ry * cw + rx;
    int B, G, R;
    if (area2) {                                             // cv2 routes exact 2x2 decimation to INTER_AREA
        const uchar4 a = __ldg(src + (size_t)(2 * dy) * cw + 2 * dx), b = __ldg(src + (size_t)(2 * dy) * cw + 2 * dx + 1);
        const uchar4 c = __ldg(src + (size_t)(2 * dy + 1) * cw + 2 * dx), d = __ldg(src + (size_t)(2 * dy + 1) * cw + 2 * dx + 1);
        B = (a.x + b.x + c.x + d.x + 2) >> 2; G = (a.y + b.y + c.y + d.y + 2) >> 2; R = (a.z + b.z + c.z + d.z + 2) >> 2;
    } else {
        const ResizeAxis X = resize_axis(dx, scale_x, sw, false), Y = resize_axis(dy, scale_y, sh, true);
        const uchar4 p00 = __ldg(src + (size_t)Y.s0 * cw + X.s0), p01 = __ldg(src + (size_t)Y.s0 * cw + X.s1);
        const uchar4 p10 = __ldg(src + (size_t)Y.s1 * cw + X.s0), p11 = __ldg(src + (size_t)Y.s1 * cw + X.s1);
        auto chan = [&](int c00, int c01, int c10, int c11) {
            const int S0 = c00 * X.a0 + c01 * X.a1, S1 = c10 * X.a0 + c11 * X.a1;
            return (((Y.a0 * (S0 >> 4)) >> 16) + ((Y.a1 * (S1 >> 4)) >> 16) + 2) >> 2;
        };
        B = chan(p00.x, p01.x, p10.x, p11.x); G = chan(p00.y, p01.y, p10.y, p11.y); R = chan(p00.z, p01.z, p10.z, p11.z);
    }
    uint8_t* q = out + ((size_t)dy * dw + dx) * 3;
    q[0] = (uint8_t)min(max(B, 0), 255); q[1] = (uint8_t)min(max(G, 0), 255); q[2] = (uint8_t)min(max(R, 0), 255);
}

cudaError_t bm_launch_crop_bounds(const uchar4* canvas, int w, int h, int thr, int* d_bounds, cudaStream_t s) {
    BM_COUNT_LAUNCHES(1), k_bounds_init<<<1, 32, 0, s>>>(d_bounds);
    const int gy = h < 592 ? h : 592;
    BM_COUNT_LAUNCHES(1), k_crop_bounds<<<dim3(bm_div_up(w, 1024) < 1 ? 1 : bm_div_up(w, 1024), gy), 256, 0, s>>>(canvas, w, h, thr, d_bounds);
    return cudaGetLastError();
}

cudaError_t bm_launch_resize_linear(const uchar4* canvas, int cw, int rx, int ry, int sw, int sh, uint8_t* d_out, int dw, int dh, cudaStream_t s) {
    const double scale_x = 1.0 / ((double)dw / (double)sw), scale_y = 1.0 / ((double)dh / (double)sh);   // cv2: scale = 1 / inv_scale
    const int area2 = (sw == 2 * dw && sh == 2 * dh) ? 1 : 0;
    const dim3 blk(32, 8);
    BM_COUNT_LAUNCHES(1), k_resize_linear<<<dim3(bm_div_up(dw, 32), bm_div_up(dh, 8)), blk, 0, s>>>(canvas, cw, rx, ry, sw, sh, d_out, dw, dh, scale_x, scale_y, area2);
    return cudaGetLastError();
}
