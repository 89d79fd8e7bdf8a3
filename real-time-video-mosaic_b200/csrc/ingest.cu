// ingest.cu -- frame ingest: packed BGR (as cv2.VideoCapture delivers it) -> gray (cv2.cvtColor BGR2GRAY,
// /root/reference/main.py:111,717) + BGRX words for the warp sampler, one pass over the frame.
//   Y = (3735*B + 19235*G + 9798*R + 16384) >> 15      (bit-exact with cv2 4.13)
#include "common.cuh"
#include "ingest.cuh"

__device__ __forceinline__ unsigned gray_of(unsigned b, unsigned g, unsigned r) {
    return (3735u * b + 19235u * g + 9798u * r + 16384u) >> 15;
}

// 4 pixels per thread: 12 B in (3 words), 4 B gray out, 16 B BGRX out.  Requires (w % 4 == 0) and 4-byte aligned rows.
__global__ void __launch_bounds__(256) k_ingest4(const uint32_t* __restrict__ bgr, int n4, uint8_t* __restrict__ gray,
                                                 uchar4* __restrict__ bgrx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const uint32_t w0 = __ldg(bgr + 3 * (size_t)i), w1 = __ldg(bgr + 3 * (size_t)i + 1), w2 = __ldg(bgr + 3 * (size_t)i + 2);
    const unsigned b0 = w0 & 255, g0 = (w0 >> 8) & 255, r0 = (w0 >> 16) & 255;
    const unsigned b1 = w0 >> 24, g1 = w1 & 255, r1 = (w1 >> 8) & 255;
    const unsigned b2 = (w1 >> 16) & 255, g2 = w1 >> 24, r2 = w2 & 255;
    const unsigned b3 = (w2 >> 8) & 255, g3 = (w2 >> 16) & 255, r3 = w2 >> 24;
    if (gray) {
        const unsigned y = gray_of(b0, g0, r0) | (gray_of(b1, g1, r1) << 8) | (gray_of(b2, g2, r2) << 16) | (gray_of(b3, g3, r3) << 24);
        reinterpret_cast<uint32_t*>(gray)[i] = y;
    }
    if (bgrx) {
        uint4 o;
        o.x = b0 | (g0 << 8) | (r0 << 16);
        o.y = b1 | (g1 << 8) | (r1 << 16);
        o.z = b2 | (g2 << 8) | (r2 << 16);
        o.w = b3 | (g3 << 8) | (r3 << 16);
        reinterpret_cast<uint4*>(bgrx)[i] = o;
    }
}

__global__ void __launch_bounds__(256) k_ingest1(const uint8_t* __restrict__ bgr, int n, uint8_t* __restrict__ gray, uchar4* __restrict__ bgrx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned b = bgr[3 * (size_t)i], g = bgr[3 * (size_t)i + 1], r = bgr[3 * (size_t)i + 2];
    if (gray) gray[i] = (uint8_t)gray_of(b, g, r);
    if (bgrx) bgrx[i] = make_uchar4(b, g, r, 0);
}

cudaError_t bm_launch_ingest(const uint8_t* d_bgr, int h, int w, uint8_t* d_gray, uchar4* d_bgrx, cudaStream_t s) {
    const size_t n = (size_t)h * w;
    const bool vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_bgr) & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(d_gray) & 3) == 0) && ((reinterpret_cast<uintptr_t>(d_bgrx) & 15) == 0);
    if (vec) BM_COUNT_LAUNCHES(1), k_ingest4<<<(unsigned)((n / 4 + 255) / 256), 256, 0, s>>>(reinterpret_cast<const uint32_t*>(d_bgr), (int)(n / 4), d_gray, d_bgrx);
    else BM_COUNT_LAUNCHES(1), k_ingest1<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_bgr, (int)n, d_gray, d_bgrx);
    return cudaGetLastError();
}
