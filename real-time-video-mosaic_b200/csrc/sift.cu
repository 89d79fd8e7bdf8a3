// sift.cu -- cv2.SIFT_create(700).detectAndCompute on the device (reference call sites: main.py:33,112,718).
// Spec: OpenCV 4.x features2d/sift (nOctaveLayers 3, contrastThreshold 0.04, edgeThreshold 10, sigma 1.6, CV_32F
// descriptors, first octave -1), pinned in SURVEY.md A.5 and oracle/sift.py.
//
// B200 design:
//  * whole Gaussian + DoG pyramid resident in HBM (~0.5 GB at 1080p); each Gaussian level is produced by ONE fused kernel:
//    shared-memory tile (+halo, reflect-101), separable row pass then column pass in OpenCV's float accumulation order
//    (rows: FMA left to right; columns: symmetric FMA form), and the DoG level is written in the same pass
//    (out - in), so the DoG pyramid costs no extra reads.
//  * extrema: one thread per DoG pixel of layers 1..3 (short-circuited 26-neighbour test), sub-pixel refinement inline
//    (rare), survivors appended with per-warp ballot/popc compaction; duplicates (two start pixels converging to the
//    same cell) are removed with an atomic claim bitmap instead of cv2's sort.
//  * orientation histograms: one warp per candidate; descriptors: one CTA per keypoint, trilinear scatter into a
//    shared 6x6x10 histogram with 64-bit fixed-point atomics (deterministic, order independent).
//  * retainBest(700): 3-pass radix select on the float response bits (ties kept), final order = cv2's
//    KeyPoint_LessThan order (x, y, size desc, angle, response desc, octave desc) by rank counting.
#include "sift.cuh"
#include "cvorder.cuh"
#include <cuda.h>          // CUtensorMap / cuTensorMapEncodeTiled (TMA descriptors of the pyramid levels)
#include <stdlib.h>
#include <float.h>
#include <math.h>
#include <string.h>
#include <new>
#include <vector>

#define SIFT_MAX_OCT 12
#define SIFT_LAYERS 3
#define SIFT_BORDER 5
// List capacities scale with the frame (SiftLayout::raw_cap / cand_cap / kp_cap, set in bm_sift_create): a scale-space extremum is a
// strict 3x3x3 optimum, so a layer holds at most one per 2x2 block -- raw extrema <= 3 layers * (16 N / 3) / 4 = 4 N for an N-pixel
// frame: raw_cap = 4 N cannot overflow; cand_cap = kp_cap = max(2^17, N / 2) is ~10x what richly textured frames produce (1080p
// synthetic sweep: 55 k candidates) and an overflow is reported through BmKeypoints::flags -> BM_ERR_UNSUPPORTED, never silently.

struct SiftOct { int w, h; long long g[6]; long long d[5]; long long claim; };   // float offsets; claim: bit offset / 32
struct SiftLayout { int noct; int raw_cap, cand_cap, kp_cap, order_cand_cap; SiftOct o[SIFT_MAX_OCT]; };

struct SiftCand {            // refined extremum (adjustLocalExtrema output)
    int o, layer, r, c;
    float ptx, pty, size, response;
    int octave_packed;
};

struct SiftGraph { const uint8_t* gray; const void* out_pt; cudaGraphExec_t exec; int launches; };

struct BmSift {
    int w, h, nfeatures;
    SiftLayout lay;
    float* pyr;              // all Gaussian + DoG levels
    float* up;               // 2x upsampled gray (float)
    unsigned* claim;         // one bit per (octave, layer, pixel)
    size_t claim_words;
    SiftCand* cand; int* ctr;            // ctr[0] = #cand, ctr[1] = #kp (pre-select), ctr[2] = overflow, ctr[3] = #selected, ctr[4] = #raw
    unsigned* raw;           // raw extrema before refinement: o<<28 | (layer-1)<<26 | r<<13 | c
    int* order_idx;          // [SIFT_ORDER_MAX_KP] permutation scratch of k_sift_order
    float* cresp; int* csel; // candidate responses, candidates that can survive retainBest (ctr[5] = threshold bits, ctr[7] = count, ctr[6] = #kp after pass A)
    float2* kpt; float* ksize; float* kangle; float* kresp; int* koct;     // pre-select keypoint list (lay.kp_cap)
    int* sel;                // indices of the selected keypoints
    unsigned* hist;          // radix-select scratch
    float* kernels_dev;      // 6 kernels x 32 taps
    CUtensorMap tmap[SIFT_MAX_OCT][6];   // TMA descriptor of the SOURCE image of blur level l in octave o (box 32 x 64|128 floats, SWIZZLE_128B)
    bool tma_ok[SIFT_MAX_OCT][6];
    bool use_tma;
    bool separate_upsample;  // BM_SIFT_SEPARATE_UPSAMPLE=1: keep k_sift_upsample + the float plane (A/B of the fused base level)
    cudaStream_t stream;
    // fork / join streams + events of the captured detect graph, and the graph cache
    cudaStream_t s2, s3;
    cudaEvent_t ev_fork, ev_l3[SIFT_MAX_OCT], ev_l5[SIFT_MAX_OCT], ev_join2, ev_join3;
    static const int kMaxGraphs = BM_DET_MAX_GRAPHS;   // 5 frame slots x 5 keypoint slots
    SiftGraph graphs[kMaxGraphs];
    int ngraphs;
    bool graphs_enabled;
};

__constant__ float c_sift_k[6][32];     // [0] = base blur (sigma 1.249), [1..5] = level blurs; taps 0..K-1
static const int h_sift_ksize[6] = {11, 11, 13, 17, 21, 27};

// ------------------------------------------------------------------------------------------------------------------
// 2x bilinear upsample of the gray image to float (cv2.resize INTER_LINEAR: weights .25/.75, edge replicate; exact)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sift_upsample(const uint8_t* __restrict__ gray, int w, int h, float* __restrict__ up) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int W = 2 * w, H = 2 * h;
    if (x >= W) return;
    // fx = x/2 - 0.25: even x -> (.25, .75) on (k-1, k); odd x -> (.75, .25) on (k, k+1)
    const int kx = x >> 1;
    int x0, x1; float ax;      // weight of the SECOND tap
    if (x & 1) { x0 = kx; x1 = min(kx + 1, w - 1); ax = 0.25f; } else { x0 = max(kx - 1, 0); x1 = kx; ax = 0.75f; }
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {                      // 4 rows per thread (32 x 32 pixels per CTA)
        const int y = blockIdx.y * 32 + threadIdx.y + 8 * rr;
        if (y >= H) break;
        const int ky = y >> 1;
        int y0, y1; float ay;
        if (y & 1) { y0 = ky; y1 = min(ky + 1, h - 1); ay = 0.25f; } else { y0 = max(ky - 1, 0); y1 = ky; ay = 0.75f; }
        const float r0 = (float)gray[(size_t)y0 * w + x0] * (1.f - ax) + (float)gray[(size_t)y0 * w + x1] * ax;
        const float r1 = (float)gray[(size_t)y1 * w + x0] * (1.f - ax) + (float)gray[(size_t)y1 * w + x1] * ax;
        up[(size_t)y * W + x] = r0 * (1.f - ay) + r1 * ay;
    }
}

__device__ __forceinline__ int refl101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    if (i < 0) i = 0;                      // degenerate tiny images
    return i;
}

// ------------------------------------------------------------------------------------------------------------------
// fused separable Gaussian + DoG (+ the decimated copy that seeds the next octave).  One instantiation per pyramid level
// (kernel taps are compile-time constant-bank operands).  CTA = 256 threads, tile 64 x TH outputs with TH = 64 - 2R
// (rounded to 4) so that the row pass has <= 256 (row, 16-output segment) units:
//   load : cp.async of the tile + halo (reflect-101 coordinates are always inside the image, so no zero fill)
//   rows : one thread = 16 consecutive outputs from a (16 + 2R)-value register window, cv2 order (tap 0 product, FMAs
//          left to right); lanes walk rows, shared-memory strides are odd -> conflict free
//   cols : one thread = TH/4 consecutive outputs of one column from a register window, cv2's symmetric FMA form
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sift_cp_async4(float* smem_dst, const float* gmem_src) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem_src) : "memory");
}
__host__ __device__ constexpr int sift_level_radius(int level) { return level <= 1 ? 5 : level == 2 ? 6 : level == 3 ? 8 : level == 4 ? 10 : 13; }
// a CTA stages SIFT_BLUR_SH rows: the row pass runs on all of them, the column pass keeps SH - 2R (rounded down to the number of
// row groups) -- the taller the tile, the smaller the share of row-pass work spent on halo rows (R = 13: 1.33x at 128 rows, 1.72x at 64)
// Measured on octave 0 (3840 x 2160): 128 rows win for the two widest kernels (level 5: 47.4 -> 41.6 us, level 4: 37.3 -> 35.5 us) and
// lose 1-3 us on levels 0-3 (two 80 KB CTAs per SM hide less latency than five 40 KB ones), hence the split.
// SH (64 or 128 staged rows) is a template parameter: the tall tile is used for levels 4 / 5 of the LARGE octaves only -- on the small
// ones fewer, bigger CTAs fill the SMs worse (aggregate over all octaves, ncu: level 4 +11 %, level 5 +3 % with tall tiles everywhere).
__host__ __device__ constexpr int sift_blur_nt(int sh) { return 4 * sh; }                           // one thread per (row, 16-output segment)
__host__ __device__ constexpr int sift_blur_ng(int sh) { return sift_blur_nt(sh) / 64; }            // row groups of the column pass
__host__ __device__ constexpr int sift_tile_h(int level, int sh) {
    return ((sh - 2 * sift_level_radius(level)) / sift_blur_ng(sh)) * sift_blur_ng(sh);
}
#define SIFT_BLUR_TSTRIDE (64 + 2 * 13 + 1)
__host__ __device__ constexpr size_t sift_blur_smem(int sh) { return (size_t)sh * (SIFT_BLUR_TSTRIDE + 65) * sizeof(float); }

// UP (level 0 of octave 0 only): the source is the 8-bit gray frame and the tile fill evaluates cv2.resize(2x, INTER_LINEAR) on the fly
// instead of reading a float plane a separate kernel wrote (k_sift_upsample: 21 us and 66 MB of traffic per 1080p frame).  The
// interpolation weights are 1/4 and 3/4 on both axes, so every intermediate of the float evaluation is an exact multiple of 1/16
// below 256: the integer form  (wy0 (wx0 g00 + wx1 g01) + wy1 (wx0 g10 + wx1 g11)) / 16  is the same float, bit for bit.
template <int LEVEL, int SHT, bool UP = false>
__global__ void __launch_bounds__(sift_blur_nt(SHT), 2) k_sift_blur(const float* __restrict__ in, float* __restrict__ out, float* __restrict__ dog,
                                                               float* __restrict__ dec, int w, int h, const uint8_t* __restrict__ gray = nullptr) {
    constexpr int R = sift_level_radius(LEVEL), K = 2 * R + 1, TW = 64, TH = sift_tile_h(LEVEL, SHT), SW = TW + 2 * R, SH = TH + 2 * R,
                  NO = TH / sift_blur_ng(SHT), NT = sift_blur_nt(SHT);
    extern __shared__ __align__(16) unsigned char sift_blur_smem_raw[];
    float (*tile)[SIFT_BLUR_TSTRIDE] = reinterpret_cast<float (*)[SIFT_BLUR_TSTRIDE]>(sift_blur_smem_raw);
    float (*rowf)[64 + 1] = reinterpret_cast<float (*)[64 + 1]>(sift_blur_smem_raw + (size_t)SHT * SIFT_BLUR_TSTRIDE * sizeof(float));
    const int bx = blockIdx.x * TW, by = blockIdx.y * TH;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (UP) {
        // 1. the gray pixels the tile needs -- rows [r_lo, r_hi] x columns [c_lo, c_hi], at most 34 x 39 -- go to shared memory (the row-pass
        //    buffer is still free); the bounds are the min / max of the (reflected) source coordinates, the same in every warp
        // 2. every tile element is interpolated from four shared-memory bytes
        const int gw = w >> 1, gh = h >> 1;
        uint8_t* gt = reinterpret_cast<uint8_t*>(&rowf[0][0]);
        constexpr int GTS = 48;                            // row stride of the staged gray tile
        int x0[3], x1[3], wx0[3];                          // per lane column: the two source columns and the weight (in quarters) of the first
        int c_lo = gw, c_hi = 0, r_lo = gh, r_hi = 0;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int X = refl101(bx + min(lane + 32 * j, SW - 1) - R, w), kx = X >> 1;
            if (X & 1) { x0[j] = kx; x1[j] = min(kx + 1, gw - 1); wx0[j] = 3; } else { x0[j] = max(kx - 1, 0); x1[j] = kx; wx0[j] = 1; }
            c_lo = min(c_lo, x0[j]); c_hi = max(c_hi, x1[j]);
        }
#pragma unroll
        for (int j = 0; j < (SH + 31) / 32; ++j) {
            const int Y = refl101(by + min(lane + 32 * j, SH - 1) - R, h), ky = Y >> 1;
            r_lo = min(r_lo, (Y & 1) ? ky : max(ky - 1, 0)); r_hi = max(r_hi, (Y & 1) ? min(ky + 1, gh - 1) : ky);
        }
        c_lo = __reduce_min_sync(0xffffffffu, c_lo); c_hi = __reduce_max_sync(0xffffffffu, c_hi);
        r_lo = __reduce_min_sync(0xffffffffu, r_lo); r_hi = __reduce_max_sync(0xffffffffu, r_hi);
        const int nc = c_hi - c_lo + 1, nr = r_hi - r_lo + 1;
        for (int i = tid; i < nr * nc; i += NT) {
            const int r = i / nc, c = i - r * nc;
            gt[r * GTS + c] = __ldg(gray + (size_t)(r_lo + r) * gw + c_lo + c);
        }
        __syncthreads();
        for (int ty = warp; ty < SH; ty += NT / 32) {
            const int Y = refl101(by + ty - R, h), ky = Y >> 1;
            int y0, y1, wy0;
            if (Y & 1) { y0 = ky; y1 = min(ky + 1, gh - 1); wy0 = 3; } else { y0 = max(ky - 1, 0); y1 = ky; wy0 = 1; }
            const uint8_t* r0 = gt + (y0 - r_lo) * GTS - c_lo;
            const uint8_t* r1 = gt + (y1 - r_lo) * GTS - c_lo;
#pragma unroll
            for (int j = 0; j < 3; ++j)
                if (lane + 32 * j < SW) {
                    const int a = wx0[j], b = 4 - a;
                    const int v = wy0 * (a * r0[x0[j]] + b * r0[x1[j]]) + (4 - wy0) * (a * r1[x0[j]] + b * r1[x1[j]]);
                    tile[ty][lane + 32 * j] = __fmul_rn((float)v, 0.0625f);
                }
        }
    } else {
        int gx[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) gx[j] = refl101(bx + lane + 32 * j - R, w);
        for (int ty = warp; ty < SH; ty += NT / 32) {
            const float* __restrict__ rowp = in + (size_t)refl101(by + ty - R, h) * w;
#pragma unroll
            for (int j = 0; j < 3; ++j)
                if (lane + 32 * j < SW) sift_cp_async4(&tile[ty][lane + 32 * j], rowp + gx[j]);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (tid < SH * 4) {
        const int seg = tid / SH, r = tid - seg * SH, c0 = seg * 16;
        float acc[16];
#pragma unroll
        for (int t = 0; t < 16 + 2 * R; ++t) {
            const float v = tile[r][c0 + t];
#pragma unroll
            for (int o = 0; o < 16; ++o) {
                const int k = t - o;
                if (k == 0) acc[o] = __fmul_rn(c_sift_k[LEVEL][0], v);
                else if (k > 0 && k < K) acc[o] = __fmaf_rn(c_sift_k[LEVEL][k], v, acc[o]);
            }
        }
#pragma unroll
        for (int o = 0; o < 16; ++o) rowf[r][c0 + o] = acc[o];
    }
    __syncthreads();
    {
        const int c = tid & 63, r0 = (tid >> 6) * NO;
        const int gx = bx + c;
        float hh[NO + 2 * R];
#pragma unroll
        for (int t = 0; t < NO + 2 * R; ++t) hh[t] = rowf[r0 + t][c];
        if (gx < w) {
            // 32-bit element offsets (a plane holds < 2^31 pixels) advanced by w per row: the three stores share one index
            const int gy0 = by + r0, nrows = h - gy0;
            unsigned idx = (unsigned)gy0 * (unsigned)w + (unsigned)gx;
            const bool decx = LEVEL == 3 && dec != nullptr && !(gx & 1) && (gx >> 1) < (w >> 1);
#pragma unroll
            for (int o = 0; o < NO; ++o) {
                if (o < nrows) {
                    float a = __fmul_rn(c_sift_k[LEVEL][R], hh[o + R]);
#pragma unroll
                    for (int t = 1; t <= R; ++t) a = __fmaf_rn(c_sift_k[LEVEL][R + t], __fadd_rn(hh[o + R + t], hh[o + R - t]), a);
                    out[idx] = a;
                    if (LEVEL >= 1) dog[idx] = __fsub_rn(a, tile[r0 + o + R][c + R]);
                    if (LEVEL == 3) {
                        // next octave base = this level decimated by 2 (INTER_NEAREST: dst(x,y) = src(2x,2y))
                        const int gy = gy0 + o;
                        if (decx && !(gy & 1) && (gy >> 1) < (h >> 1)) dec[(unsigned)(gy >> 1) * (unsigned)(w >> 1) + (unsigned)(gx >> 1)] = a;
                    }
                }
                idx += (unsigned)w;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// TMA form of the same kernel (same arithmetic, same outputs).  The tile + halo is fetched by the TMA unit instead of one 4-byte
// cp.async per element with reflect-101 index math: three 2-D box loads (32 floats x SHT rows each, SWIZZLE_128B) issued by ONE
// thread land the 96-column tile in shared memory and complete on an mbarrier.  SWIZZLE_128B XORs the 16-byte chunk index of every
// 128-byte line with the line number mod 8, so the row pass -- lanes walk ROWS -- reads its window as LDS.128: the 8 lanes of a
// quarter warp hit 8 different chunks (conflict free) and a thread needs (16 + 2R) / 4 loads instead of 16 + 2R.  The TMA unit
// zero-fills outside the image, OpenCV wants BORDER_REFLECT_101, and a box must start on a 16-byte boundary of its row (it begins
// RA = R rounded up to 4 columns left of the tile; an unaligned start faults with "illegal instruction"): tiles whose box leaves the image (the frame's rim, ~12 % of octave
// 0) fill the same swizzled layout with the per-element path.
// ------------------------------------------------------------------------------------------------------------------
__host__ __device__ constexpr size_t sift_blur_tma_smem(int sh) { return (size_t)sh * (3 * 128 + 65 * sizeof(float)) + 1024; }

__device__ __forceinline__ unsigned sift_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// byte offset of tile element (row, col) in the three-box swizzled layout; col in [0, 96)
template <int SHT>
__device__ __forceinline__ unsigned sift_swz(int row, int col) {
    const int q = col >> 2;                                   // 16-byte chunk column
    return (unsigned)((q >> 3) * (SHT * 128) + row * 128 + (((q & 7) ^ (row & 7)) << 4) + ((col & 3) << 2));
}

template <int LEVEL, int SHT>
__global__ void __launch_bounds__(sift_blur_nt(SHT), 2) k_sift_blur_tma(const __grid_constant__ CUtensorMap tm, const float* __restrict__ in,
                                                                   float* __restrict__ out, float* __restrict__ dog, float* __restrict__ dec,
                                                                   int w, int h) {
    constexpr int R = sift_level_radius(LEVEL), K = 2 * R + 1, TW = 64, TH = sift_tile_h(LEVEL, SHT), SH = TH + 2 * R,
                  NO = TH / sift_blur_ng(SHT), NT = sift_blur_nt(SHT),
                  RA = (R + 3) & ~3,          // the box must start on a 16-byte boundary of its row: it begins RA >= R columns left of the tile
                  D = RA - R, NQ = (D + 16 + 2 * R + 3) / 4;
    static_assert(64 + R + RA <= 96, "tile + halo must fit the three 32-column boxes");
    extern __shared__ unsigned char sift_blur_smem_raw2[];
    __shared__ __align__(8) unsigned long long mbar;
    unsigned char* tile = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sift_blur_smem_raw2) + 1023) & ~(uintptr_t)1023);
    float (*rowf)[64 + 1] = reinterpret_cast<float (*)[64 + 1]>(tile + (size_t)SHT * 384);
    const int bx = blockIdx.x * TW, by = blockIdx.y * TH;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool interior = bx - RA >= 0 && bx - RA + 96 <= w && by - R >= 0 && by - R + SHT <= h;      // (CTA uniform)
    if (interior) {
        if (tid == 0) {
            const unsigned mb = sift_smem_u32(&mbar);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((unsigned)(3 * SHT * 128)) : "memory");
#pragma unroll
            for (int b = 0; b < 3; ++b)
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(sift_smem_u32(tile + (size_t)b * SHT * 128)), "l"(&tm), "r"(bx - RA + 32 * b), "r"(by - R), "r"(mb)
                             : "memory");
        }
        __syncthreads();                                       // the barrier is initialised for everyone
        {
            const unsigned mb = sift_smem_u32(&mbar);
            unsigned done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(mb), "r"(0u) : "memory");
        }
    } else {
        int gx[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) gx[j] = refl101(bx + lane + 32 * j - RA, w);
        for (int ty = warp; ty < SH; ty += NT / 32) {
            const float* __restrict__ rowp = in + (size_t)refl101(by + ty - R, h) * w;
#pragma unroll
            for (int j = 0; j < 3; ++j)
                sift_cp_async4(reinterpret_cast<float*>(tile + sift_swz<SHT>(ty, lane + 32 * j)), rowp + gx[j]);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    }
    if (tid < SH * 4) {
        const int seg = tid / SH, r = tid - seg * SH, c0 = seg * 16;
        float acc[16];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const float4 v4 = *reinterpret_cast<const float4*>(tile + sift_swz<SHT>(r, c0 + 4 * q));
            const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int t = 4 * q + e - D;                          // window index of tile column c0 + 4 q + e
                if (t >= 0 && t < 16 + 2 * R) {
#pragma unroll
                    for (int o = 0; o < 16; ++o) {
                        const int k = t - o;
                        if (k == 0) acc[o] = __fmul_rn(c_sift_k[LEVEL][0], vv[e]);
                        else if (k > 0 && k < K) acc[o] = __fmaf_rn(c_sift_k[LEVEL][k], vv[e], acc[o]);
                    }
                }
            }
        }
#pragma unroll
        for (int o = 0; o < 16; ++o) rowf[r][c0 + o] = acc[o];
    }
    __syncthreads();
    {
        const int c = tid & 63, r0 = (tid >> 6) * NO;
        const int gx = bx + c;
        float hh[NO + 2 * R];
#pragma unroll
        for (int t = 0; t < NO + 2 * R; ++t) hh[t] = rowf[r0 + t][c];
        if (gx < w) {
            const int gy0 = by + r0, nrows = h - gy0;
            unsigned idx = (unsigned)gy0 * (unsigned)w + (unsigned)gx;
            const bool decx = LEVEL == 3 && dec != nullptr && !(gx & 1) && (gx >> 1) < (w >> 1);
#pragma unroll
            for (int o = 0; o < NO; ++o) {
                if (o < nrows) {
                    float a = __fmul_rn(c_sift_k[LEVEL][R], hh[o + R]);
#pragma unroll
                    for (int t = 1; t <= R; ++t) a = __fmaf_rn(c_sift_k[LEVEL][R + t], __fadd_rn(hh[o + R + t], hh[o + R - t]), a);
                    out[idx] = a;
                    if (LEVEL >= 1) dog[idx] = __fsub_rn(a, *reinterpret_cast<const float*>(tile + sift_swz<SHT>(r0 + o + R, c + RA)));
                    if (LEVEL == 3) {
                        const int gy = gy0 + o;
                        if (decx && !(gy & 1) && (gy >> 1) < (h >> 1)) dec[(unsigned)(gy >> 1) * (unsigned)(w >> 1) + (unsigned)(gx >> 1)] = a;
                    }
                }
                idx += (unsigned)w;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// scale-space extrema + adjustLocalExtrema
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool solve3_cramer(const float a[3][3], const float b[3], float x[3]) {
    // Matx33f::solve(DECOMP_LU) fast path: explicit determinant / Cramer in float
    float d = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
              a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
    if (d == 0.f) { x[0] = x[1] = x[2] = 0.f; return false; }
    d = 1.f / d;
    x[0] = d * (b[0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (b[1] * a[2][2] - a[1][2] * b[2]) + a[0][2] * (b[1] * a[2][1] - a[1][1] * b[2]));
    x[1] = d * (a[0][0] * (b[1] * a[2][2] - a[1][2] * b[2]) - b[0] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) + a[0][2] * (a[1][0] * b[2] - b[1] * a[2][0]));
    x[2] = d * (a[0][0] * (a[1][1] * b[2] - b[1] * a[2][1]) - a[0][1] * (a[1][0] * b[2] - b[1] * a[2][0]) + b[0] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]));
    return true;
}

__device__ bool adjust_local_extrema(const float* __restrict__ pyr, const SiftOct& O, int o, int& layer, int& r, int& c, SiftCand& out) {
    const float img_scale = 1.f / 255.f, deriv_scale = img_scale * 0.5f, second_deriv_scale = img_scale, cross_deriv_scale = img_scale * 0.25f;
    const int w = O.w, h = O.h;
    float xi = 0, xr = 0, xc = 0;
    int i = 0;
    for (; i < 5; ++i) {
        const float* img = pyr + O.d[layer]; const float* prv = pyr + O.d[layer - 1]; const float* nxt = pyr + O.d[layer + 1];
        const size_t p = (size_t)r * w + c;
        const float dD[3] = {(img[p + 1] - img[p - 1]) * deriv_scale, (img[p + w] - img[p - w]) * deriv_scale, (nxt[p] - prv[p]) * deriv_scale};
        const float v2 = img[p] * 2.f;
        const float dxx = (img[p + 1] + img[p - 1] - v2) * second_deriv_scale;
        const float dyy = (img[p + w] + img[p - w] - v2) * second_deriv_scale;
        const float dss = (nxt[p] + prv[p] - v2) * second_deriv_scale;
        const float dxy = (img[p + w + 1] - img[p + w - 1] - img[p - w + 1] + img[p - w - 1]) * cross_deriv_scale;
        const float dxs = (nxt[p + 1] - nxt[p - 1] - prv[p + 1] + prv[p - 1]) * cross_deriv_scale;
        const float dys = (nxt[p + w] - nxt[p - w] - prv[p + w] + prv[p - w]) * cross_deriv_scale;
        const float Hm[3][3] = {{dxx, dxy, dxs}, {dxy, dyy, dys}, {dxs, dys, dss}};
        float X[3];
        solve3_cramer(Hm, dD, X);
        xi = -X[2]; xr = -X[1]; xc = -X[0];
        if (fabsf(xi) < 0.5f && fabsf(xr) < 0.5f && fabsf(xc) < 0.5f) break;
        if (fabsf(xi) > (float)(INT_MAX / 3) || fabsf(xr) > (float)(INT_MAX / 3) || fabsf(xc) > (float)(INT_MAX / 3)) return false;
        c += __float2int_rn(xc); r += __float2int_rn(xr); layer += __float2int_rn(xi);
        if (layer < 1 || layer > SIFT_LAYERS || c < SIFT_BORDER || c >= w - SIFT_BORDER || r < SIFT_BORDER || r >= h - SIFT_BORDER) return false;
    }
    if (i >= 5) return false;
    {
        const float* img = pyr + O.d[layer]; const float* prv = pyr + O.d[layer - 1]; const float* nxt = pyr + O.d[layer + 1];
        const size_t p = (size_t)r * w + c;
        const float d0 = (img[p + 1] - img[p - 1]) * deriv_scale, d1 = (img[p + w] - img[p - w]) * deriv_scale, d2 = (nxt[p] - prv[p]) * deriv_scale;
        const float t = d0 * xc + d1 * xr + d2 * xi;
        const float contr = img[p] * img_scale + t * 0.5f;
        if (fabsf(contr) * SIFT_LAYERS < 0.04f) return false;
        const float v2 = img[p] * 2.f;
        const float dxx = (img[p + 1] + img[p - 1] - v2) * second_deriv_scale;
        const float dyy = (img[p + w] + img[p - w] - v2) * second_deriv_scale;
        const float dxy = (img[p + w + 1] - img[p + w - 1] - img[p - w + 1] + img[p - w - 1]) * cross_deriv_scale;
        const float tr = dxx + dyy, det = dxx * dyy - dxy * dxy;
        if (det <= 0 || tr * tr * 10.f >= 11.f * 11.f * det) return false;
        out.o = o; out.layer = layer; out.r = r; out.c = c;
        out.ptx = (c + xc) * (float)(1 << o);
        out.pty = (r + xr) * (float)(1 << o);
        out.octave_packed = o + (layer << 8) + (__double2int_rn(((double)xi + 0.5) * 255.0) << 16);
        out.size = 1.6f * powf(2.f, (layer + xi) / (float)SIFT_LAYERS) * (float)(1 << o) * 2.f;
        out.response = fabsf(contr);
    }
    return true;
}

// 26-neighbour test of one DoG sample; prv / cur / nxt point at the sample in the three layers
__device__ __forceinline__ bool sift_is_extremum(const float* __restrict__ prv, const float* __restrict__ cur, const float* __restrict__ nxt,
                                                 int w, float v) {
    if (v > 0) {
        return v >= cur[-1] && v >= cur[1] && v >= cur[-w - 1] && v >= cur[-w] && v >= cur[-w + 1] && v >= cur[w - 1] && v >= cur[w] && v >= cur[w + 1] &&
               v >= nxt[-1] && v >= nxt[1] && v >= nxt[-w - 1] && v >= nxt[-w] && v >= nxt[-w + 1] && v >= nxt[w - 1] && v >= nxt[w] && v >= nxt[w + 1] &&
               v >= prv[-1] && v >= prv[1] && v >= prv[-w - 1] && v >= prv[-w] && v >= prv[-w + 1] && v >= prv[w - 1] && v >= prv[w] && v >= prv[w + 1];
    }
    return v <= cur[-1] && v <= cur[1] && v <= cur[-w - 1] && v <= cur[-w] && v <= cur[-w + 1] && v <= cur[w - 1] && v <= cur[w] && v <= cur[w + 1] &&
           v <= nxt[-1] && v <= nxt[1] && v <= nxt[-w - 1] && v <= nxt[-w] && v <= nxt[-w + 1] && v <= nxt[w - 1] && v <= nxt[w] && v <= nxt[w + 1] &&
           v <= prv[-1] && v <= prv[1] && v <= prv[-w - 1] && v <= prv[-w] && v <= prv[-w + 1] && v <= prv[w - 1] && v <= prv[w] && v <= prv[w + 1];
}

// Phase 1 -- one thread per pixel of octave `o`: the 5 DoG samples of the pixel are read once (coalesced); a sample of
// layers 1..3 is a candidate only if it beats the threshold and the samples directly above / below it -- which rejects
// almost everything before any neighbour is touched.  The conjunction is the same 26-neighbour test as cv2's, only its
// order differs.  Raw extrema are compacted (ballot / popc) into a list; the divergent sub-pixel refinement runs densely
// over that list in phase 2 instead of stalling 31 lanes of the detecting warp.
__global__ void __launch_bounds__(256) k_sift_extrema(SiftLayout lay, int o, const float* __restrict__ pyr, unsigned* __restrict__ raw,
                                                      int* __restrict__ ctr) {
    const SiftOct O = lay.o[o];
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y * blockDim.y + threadIdx.y;
    const int w = O.w, h = O.h;
    const bool inside = c >= SIFT_BORDER && c < w - SIFT_BORDER && r >= SIFT_BORDER && r < h - SIFT_BORDER;
    const size_t p = (size_t)r * w + c;
    float d[5];
#pragma unroll
    for (int l = 0; l < 5; ++l) d[l] = inside ? __ldg(pyr + O.d[l] + p) : 0.f;
#pragma unroll
    for (int l0 = 1; l0 <= SIFT_LAYERS; ++l0) {
        const float v = d[l0];
        // threshold = cvFloor(0.5*0.04/3*255) = 1
        const bool found = inside && fabsf(v) > 1.0f && (v > 0 ? (v >= d[l0 - 1] && v >= d[l0 + 1]) : (v <= d[l0 - 1] && v <= d[l0 + 1])) &&
                           sift_is_extremum(pyr + O.d[l0 - 1] + p, pyr + O.d[l0] + p, pyr + O.d[l0 + 1] + p, w, v);
        const unsigned bal = __ballot_sync(0xffffffffu, found);
        if (bal) {
            const int lane = threadIdx.x & 31, leader = __ffs(bal) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&ctr[4], __popc(bal));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (found) {
                const int idx = base + __popc(bal & ((1u << lane) - 1u));
                if (idx < lay.raw_cap) raw[idx] = ((unsigned)o << 28) | ((unsigned)(l0 - 1) << 26) | ((unsigned)r << 13) | (unsigned)c;
                else ctr[2] = 1;
            }
        }
    }
}

// Tiled form of phase 1 for octaves whose rows are 16-byte aligned (w % 4 == 0, w >= 64): 12 % of the samples pass the cheap
// centre-column test, but 82 % of all warps hold at least one of them, so in the per-pixel kernel nearly every warp pays for the
// (serialised, short-circuit) 26-neighbour test of every layer.  Here a CTA stages a 64 x 16 tile (+1 halo) of the 5 DoG planes in
// shared memory with whole-row float4 loads; every WARP compacts the samples of its two tile rows that pass the centre-column test into
// its own list and runs the neighbour test densely over it (24 independent shared loads + a max / min tree) -- two barriers per tile
// (tile staged / tile free again), no CTA-wide list.  Found extrema are collected per CTA over all its tiles: one atomic on the global
// list counter per CTA.  Same conjunction as cv2's test, evaluated in another order.
#define EX_TW 64
#define EX_TH 14                       // interior rows; 16 rows are staged (one float4 per thread and plane)
#define EX_STRIDE 72                   // floats per tile row: [3] left halo, [4, 68) interior (16-byte aligned), [68] right halo
#define EX_FOUND_CAP 256
__global__ void __launch_bounds__(256) k_sift_extrema_tile(SiftLayout lay, int o, int kt, const float* __restrict__ pyr,
                                                           unsigned* __restrict__ raw, int* __restrict__ ctr) {
    __shared__ __align__(16) float sm[5][EX_TH + 2][EX_STRIDE];
    __shared__ unsigned short s_list[8 * 2 * EX_TW * 3];           // per warp: its two tile rows x 64 columns x 3 layers
    __shared__ unsigned s_found[EX_FOUND_CAP];
    __shared__ int s_nf, s_base;
    const SiftOct& O = lay.o[o];                                     // stays in the parameter bank
    const int w = O.w, h = O.h;
    const size_t pn = (size_t)w * h;                                 // the 5 DoG planes of an octave are consecutive (checked by the launcher)
    const float* __restrict__ dog = pyr + O.d[0];
    const int tid = threadIdx.x, lane = tid & 31, q = tid & 15, rr = tid >> 4;
    const int x0 = blockIdx.x * EX_TW, gx = x0 + 4 * q;
    const int hl = tid >> 5, hr = (tid & 31) >> 1, side = tid & 1;   // halo duty of threads 0..159: plane, tile row, left / right
    const int hx = side ? x0 + EX_TW : x0 - 1;
    float* const my_sm = &sm[0][rr][4 + 4 * q];
    float* const my_halo = &sm[tid < 160 ? hl : 0][hr][side ? 4 + EX_TW : 3];
    const unsigned ebase = ((unsigned)rr << 8) | ((unsigned)(4 * q) << 2);
    if (tid == 0) s_nf = 0;
    // a CTA walks kt vertically adjacent tiles: the per-thread address set-up is paid once
    for (int it = 0; it < kt; ++it) {
        const int y0 = (blockIdx.y * kt + it) * EX_TH - 1;           // tile row rr <-> image row y0 + rr
        if (y0 + 1 >= h) break;
        const int r = y0 + rr;
        // stage the tile: the thread's own 4 pixels of the 5 planes stay in registers for the centre-column test
        float4 d4[5];
        {
            const bool ok = gx < w && r >= 0 && r < h;
            const float* __restrict__ src = dog + (size_t)(ok ? r : 0) * w + (ok ? gx : 0);
#pragma unroll
            for (int l = 0; l < 5; ++l) d4[l] = ok ? __ldg(reinterpret_cast<const float4*>(src + l * pn)) : make_float4(0.f, 0.f, 0.f, 0.f);
            float hv = 0.f;
            if (tid < 160) {
                const int hy = y0 + hr;
                if (hx >= 0 && hx < w && hy >= 0 && hy < h) hv = __ldg(dog + hl * pn + (size_t)hy * w + hx);
            }
#pragma unroll
            for (int l = 0; l < 5; ++l) *reinterpret_cast<float4*>(my_sm + l * (EX_TH + 2) * EX_STRIDE) = d4[l];
            if (tid < 160) *my_halo = hv;
        }
        // centre-column test of the thread's 4 pixels x 3 layers -> 12-bit mask (bit 3 j + l0 - 1); threshold = cvFloor(0.5*0.04/3*255) = 1
        unsigned mask = 0;
        if (rr >= 1 && rr <= EX_TH && r >= SIFT_BORDER && r < h - SIFT_BORDER) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = gx + j;
                const bool inside = c >= SIFT_BORDER && c < w - SIFT_BORDER;
                float d[5];
#pragma unroll
                for (int l = 0; l < 5; ++l) d[l] = j == 0 ? d4[l].x : j == 1 ? d4[l].y : j == 2 ? d4[l].z : d4[l].w;
#pragma unroll
                for (int l0 = 1; l0 <= SIFT_LAYERS; ++l0) {
                    // v > 1 && v >= both neighbours  <=>  v >= max3(below, above, nextafter(1));  likewise for minima: 5 instructions
                    const float v = d[l0];
                    const float hi = fmaxf(fmaxf(d[l0 - 1], d[l0 + 1]), __int_as_float(0x3F800001));
                    const float lo = fminf(fminf(d[l0 - 1], d[l0 + 1]), __int_as_float(0xBF800001));
                    if (inside && (v >= hi || v <= lo)) mask |= 1u << (3 * j + l0 - 1);
                }
            }
        }
        // warp list of the survivors (a warp owns two tile rows = 384 samples): warp scan of the counts, no CTA-wide step
        unsigned short* const wl = s_list + (tid >> 5) * (2 * EX_TW * 3);
        const int cnt = __popc(mask);
        int inc = cnt;
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, dlt); if (lane >= dlt) inc += u; }
        const int n = __shfl_sync(0xffffffffu, inc, 31);
        {
            int at = inc - cnt;
            while (mask) {
                const unsigned b = __ffs(mask) - 1; mask &= mask - 1;
                const unsigned j = (b * 11u) >> 5;                     // b / 3 for b < 12
                wl[at++] = (unsigned short)(ebase + (j << 2) + (b - 3u * j));
            }
        }
        __syncthreads();                                               // the tile is complete (and this warp's list)
        for (int i = lane; i < n; i += 32) {
            const unsigned e = wl[i];
            const int l0 = (e & 3) + 1, col = (e >> 2) & 63, row = e >> 8;
            const float* __restrict__ ctrp = &sm[l0][row][4 + col];
            const float v = *ctrp;
            float mx = v, mn = v;
#pragma unroll
            for (int dl = -1; dl <= 1; ++dl)
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx) {
                        if (dy == 0 && dx == 0) continue;          // the column itself was tested above
                        const float t = ctrp[(dl * (EX_TH + 2) + dy) * EX_STRIDE + dx];
                        mx = fmaxf(mx, t); mn = fminf(mn, t);
                    }
            if (v > 0 ? mx <= v : mn >= v) {
                const unsigned word = ((unsigned)o << 28) | ((unsigned)(l0 - 1) << 26) | ((unsigned)(y0 + row) << 13) | (unsigned)(x0 + col);
                const int k = atomicAdd(&s_nf, 1);
                if (k < EX_FOUND_CAP) s_found[k] = word;
                else {                                              // more than the CTA buffer holds: straight to the global list
                    const int idx = atomicAdd(&ctr[4], 1);
                    if (idx < lay.raw_cap) raw[idx] = word; else ctr[2] = 1;
                }
            }
        }
        if (it + 1 < kt) __syncthreads();                           // the tile may be overwritten
    }
    // the extrema of all kt tiles leave the CTA together: one atomic on the global list counter per CTA
    __syncthreads();
    const int nf = min(s_nf, EX_FOUND_CAP);
    if (nf) {
        if (tid == 0) s_base = atomicAdd(&ctr[4], nf);
        __syncthreads();
        for (int i = tid; i < nf; i += 256) {
            const int idx = s_base + i;
            if (idx < lay.raw_cap) raw[idx] = s_found[i]; else ctr[2] = 1;
        }
    }
}

// Phase 2 -- adjustLocalExtrema: one thread per raw extremum
__global__ void __launch_bounds__(128) k_sift_refine(SiftLayout lay, const float* __restrict__ pyr, const unsigned* __restrict__ raw,
                                                     unsigned* __restrict__ claim, SiftCand* __restrict__ cand, float* __restrict__ cresp,
                                                     int* __restrict__ ctr) {
    int n = ctr[4]; if (n > lay.raw_cap) n = lay.raw_cap;
  for (int i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {                   // fixed grid, CTA-uniform trip count
    const int i = i0 + threadIdx.x;
    bool found = false;
    SiftCand cd;
    if (i < n) {
        const unsigned u = raw[i];
        const int o = u >> 28;
        int layer = ((u >> 26) & 3) + 1, r1 = (u >> 13) & 8191, c1 = u & 8191;
        const SiftOct O = lay.o[o];
        if (adjust_local_extrema(pyr, O, o, layer, r1, c1, cd)) {
            // duplicate removal: the first start pixel that reaches a cell claims it
            const long long bit = O.claim + ((long long)(layer - 1) * O.h + r1) * O.w + c1;
            const unsigned m = 1u << (bit & 31);
            const unsigned old = atomicOr(&claim[bit >> 5], m);
            found = (old & m) == 0;
        }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, found);
    if (bal) {
        const int lane = threadIdx.x & 31, leader = __ffs(bal) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&ctr[0], __popc(bal));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (found) {
            const int idx = base + __popc(bal & ((1u << lane) - 1u));
            if (idx < lay.cand_cap) { cand[idx] = cd; cresp[idx] = cd.response; } else ctr[2] = 1;
        }
    }
  }
}

// hal::fastAtan2 (vector path: FMA Horner), degrees
__device__ __forceinline__ float sift_atan2_deg(float y, float x) {
    const float k = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * k, p3 = -0.3258083974640975f * k, p5 = 0.1555786518463281f * k, p7 = -0.04432655554792128f * k;
    const float ax = fabsf(x), ay = fabsf(y);
    const float c = __fdiv_rn(fminf(ax, ay), __fadd_rn(fmaxf(ax, ay), 2.220446049250313e-16f));
    const float c2 = __fmul_rn(c, c);
    float a = __fmul_rn(__fmaf_rn(__fmaf_rn(__fmaf_rn(c2, p7, p5), c2, p3), c2, p1), c);
    if (ay > ax) a = 90.f - a;
    if (x < 0.f) a = 180.f - a;
    if (y < 0.f) a = 360.f - a;
    return a;
}

// calcOrientationHist + peak extraction: one warp per refined candidate.  Every lane accumulates its pixels into a private
// column of the warp's histogram (2^-24 fixed point: integer sums are order independent, so the result is deterministic and
// needs no atomics), the 32 columns are then summed with a skewed, conflict-free read.
#define SIFT_ORI_WARPS 4
// retainBest(nfeatures) keeps keypoints by response, and a keypoint's response is its candidate's: with T' = the nfeatures-th
// largest CANDIDATE response, pass A (rest == 0) builds histograms only for the candidates listed in csel (response >= T').
// If they yield >= nfeatures keypoints, the nfeatures-th largest keypoint response is >= T' and no other candidate can
// survive; otherwise (a candidate without any peak -- practically never) pass B (rest == 1) processes the remaining ones.
__global__ void __launch_bounds__(32 * SIFT_ORI_WARPS) k_sift_orient(SiftLayout lay, const float* __restrict__ pyr, const SiftCand* __restrict__ cand,
                                                     const int* __restrict__ csel, int rest, int nfeatures,
                                                     int* __restrict__ ctr, float2* __restrict__ kpt, float* __restrict__ ksize,
                                                     float* __restrict__ kangle, float* __restrict__ kresp, int* __restrict__ koct) {
    __shared__ unsigned long long sh_priv[SIFT_ORI_WARPS][36][32];
    __shared__ float sh_f[SIFT_ORI_WARPS][40];
    const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int ncand = ctr[0]; if (ncand > lay.cand_cap) ncand = lay.cand_cap;
    // rest == 2: EVERY candidate, for frames whose keypoints are emitted in cv2's retainBest order (that order depends on the number
    // of orientations of every candidate); taken iff the frame has at most order_cand_cap candidates.  Passes A / B otherwise.
    const bool exact = ncand <= lay.order_cand_cap;
    if ((rest == 2) != exact) return;
    if (rest == 1 && (ctr[6] >= nfeatures || ctr[7] >= ncand)) return;      // pass A was enough / had everything
  for (int ci = blockIdx.x * SIFT_ORI_WARPS + wi; ci < (rest ? ncand : ctr[7]); ci += gridDim.x * SIFT_ORI_WARPS) {
    if (rest == 1 && __float_as_uint(cand[ci].response) >= (unsigned)ctr[5]) continue;     // done in pass A
    const SiftCand cd = cand[rest ? ci : csel[ci]];
    const SiftOct O = lay.o[cd.o];
    const float* img = pyr + O.g[cd.layer];
    const int w = O.w, h = O.h;
    const float scl_octv = cd.size * 0.5f / (float)(1 << cd.o);
    const int radius = __float2int_rn(4.5f * scl_octv);
    const float sigma = 1.5f * scl_octv;
    const float expf_scale = -1.f / (2.f * sigma * sigma);
#pragma unroll
    for (int i = 0; i < 36; ++i) sh_priv[wi][i][lane] = 0ull;
    const int side = 2 * radius + 1, len = side * side;
    for (int k = lane; k < len; k += 32) {
        const int i = k / side - radius, j = k % side - radius;
        const int y = cd.r + i, x = cd.c + j;
        if (y <= 0 || y >= h - 1 || x <= 0 || x >= w - 1) continue;
        const float dx = img[(size_t)y * w + x + 1] - img[(size_t)y * w + x - 1];
        const float dy = img[(size_t)(y - 1) * w + x] - img[(size_t)(y + 1) * w + x];
        const float wgt = expf((float)(i * i + j * j) * expf_scale);
        const float ori = sift_atan2_deg(dy, dx);
        const float mag = sqrtf(dx * dx + dy * dy);
        int bin = __float2int_rn((36.f / 360.f) * ori);
        if (bin >= 36) bin -= 36;
        if (bin < 0) bin += 36;
        sh_priv[wi][bin][lane] += (unsigned long long)__float2ll_rn(wgt * mag * 16777216.f);
    }
    __syncwarp();
    for (int b = lane; b < 36; b += 32) {
        unsigned long long t = 0ull;
#pragma unroll
        for (int l = 0; l < 32; ++l) t += sh_priv[wi][b][(l + lane) & 31];
        sh_f[wi][b + 2] = (float)((double)t * (1.0 / 16777216.0));
    }
    __syncwarp();
    if (lane == 0) { sh_f[wi][0] = sh_f[wi][36]; sh_f[wi][1] = sh_f[wi][37]; sh_f[wi][38] = sh_f[wi][2]; sh_f[wi][39] = sh_f[wi][3]; }
    __syncwarp();
    float hv[2];
    float mx = 0.f;
    for (int q = 0; q < 2; ++q) {
        const int i = lane + 32 * q;
        hv[q] = 0.f;
        if (i < 36) {
            const float* t = &sh_f[wi][i + 2];
            hv[q] = (t[-2] + t[2]) * (1.f / 16.f) + (t[-1] + t[1]) * (4.f / 16.f) + t[0] * (6.f / 16.f);
            mx = fmaxf(mx, hv[q]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __syncwarp();
    for (int q = 0; q < 2; ++q) { const int i = lane + 32 * q; if (i < 36) sh_f[wi][i] = hv[q]; }     // smoothed hist in [0,36)
    __syncwarp();
    const float mag_thr = mx * 0.8f;
    for (int q = 0; q < 2; ++q) {
        const int j = lane + 32 * q;
        bool peak = false; float angle = 0.f;
        if (j < 36) {
            const int l = j > 0 ? j - 1 : 35, r2 = j < 35 ? j + 1 : 0;
            const float hj = sh_f[wi][j], hl = sh_f[wi][l], hr = sh_f[wi][r2];
            if (hj > hl && hj > hr && hj >= mag_thr) {
                float bin = j + 0.5f * (hl - hr) / (hl - 2 * hj + hr);
                bin = bin < 0 ? 36 + bin : (bin >= 36 ? bin - 36 : bin);
                angle = 360.f - (360.f / 36.f) * bin;
                if (fabsf(angle - 360.f) < FLT_EPSILON) angle = 0.f;
                peak = true;
            }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, peak);
        if (bal) {
            const int leader = __ffs(bal) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&ctr[1], __popc(bal));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (peak) {
                const int idx = base + __popc(bal & ((1u << lane) - 1u));
                if (idx < lay.kp_cap) {
                    kpt[idx] = make_float2(cd.ptx, cd.pty); ksize[idx] = cd.size; kangle[idx] = angle; kresp[idx] = cd.response; koct[idx] = cd.octave_packed;
                } else ctr[2] = 1;
            }
        }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------------------------
// retainBest(nfeatures) (ties kept) via radix select on the response bits, then cv2's KeyPoint_LessThan order
// ------------------------------------------------------------------------------------------------------------------
// Generic form: keep the values >= the K-th largest (ties kept).  n = min(*n_ptr, ncap) values; indices of the kept values go
// to sel (any order), their number to *out_count (capped at BM_KP_CAP, overflow flagged in ctr[2]), the threshold bits to
// *out_thr.  Used twice: on the candidates' responses BEFORE orientation assignment (only candidates that can survive
// retainBest get an orientation histogram) and on the final keypoint list.
__global__ void __launch_bounds__(1024) k_sift_select(int nfeatures, const int* __restrict__ n_ptr, int ncap, int* __restrict__ ctr,
                                                      const float* __restrict__ kresp, int* __restrict__ sel, int* __restrict__ out_count,
                                                      unsigned* __restrict__ out_thr) {
    __shared__ unsigned hist[2048];
    __shared__ unsigned s_prefix, s_mask, s_remaining, s_wsum[32];
    __shared__ int s_count;
    if (out_thr == nullptr && ctr[8]) return;            // the keypoint list was already selected in cv2's order (k_sift_order)
    int n = *n_ptr; if (n > ncap) n = ncap;
    const int tid = threadIdx.x;
    unsigned thr_bits = 0;      // keep response bits >= thr_bits
    if (n > nfeatures && n <= 1024) {
        // short lists (the final keypoint list: ~900 entries): one value per thread, rank by counting -- a value is kept iff fewer
        // than nfeatures values are strictly greater, which is exactly ">= the nfeatures-th largest, ties kept"
        unsigned* vals = hist;
        const unsigned mine = tid < n ? __float_as_uint(__ldg(kresp + tid)) : 0u;
        vals[tid] = mine;
        if (tid == 0) s_count = 0;
        __syncthreads();
        int greater = 0;
        for (int j = 0; j < n; ++j) greater += vals[j] > mine ? 1 : 0;
        const bool keep = tid < n && greater < nfeatures;
        if (keep) { const int p = atomicAdd(&s_count, 1); if (p < BM_KP_CAP) sel[p] = tid; }
        // threshold bits = the smallest kept value
        unsigned kmin = keep ? mine : 0xffffffffu;
        kmin = __reduce_min_sync(0xffffffffu, kmin);
        if ((tid & 31) == 0) s_wsum[tid >> 5] = kmin;
        __syncthreads();
        if (tid == 0) {
            unsigned t = 0xffffffffu;
            for (int w = 0; w < 32; ++w) t = min(t, s_wsum[w]);
            *out_count = s_count;
            if (out_thr) *out_thr = t;
        }
        return;
    }
    if (n > nfeatures) {
        if (tid == 0) { s_prefix = 0; s_mask = 0; s_remaining = (unsigned)nfeatures; }
        __syncthreads();
        const int shifts[3] = {21, 10, 0}; const int bitsn[3] = {11, 11, 10};
        for (int pass = 0; pass < 3; ++pass) {
            for (int i = tid; i < 2048; i += blockDim.x) hist[i] = 0;
            __syncthreads();
            const unsigned prefix = s_prefix, mask = s_mask;
            for (int i0 = tid; i0 < n; i0 += 8 * blockDim.x) {          // 8 independent loads in flight per thread
                unsigned b[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { const int i = i0 + u * blockDim.x; b[u] = i < n ? __float_as_uint(__ldg(kresp + i)) : 0xffffffffu; }
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (i0 + u * (int)blockDim.x < n && (b[u] & mask) == prefix) atomicAdd(&hist[(b[u] >> shifts[pass]) & ((1u << bitsn[pass]) - 1u)], 1u);
            }
            __syncthreads();
            {   // the bin holding the rem-th largest key: suffix sums over the bins by a block-wide scan (2 bins per thread,
                // bins in descending order), exactly one bin satisfies cum_before < rem <= cum_before + hist[bin]
                const int nbins = 1 << bitsn[pass];
                const int j0 = 2 * tid, b0 = nbins - 1 - j0, b1 = b0 - 1;            // descending bin order
                const unsigned h0 = b0 >= 0 ? hist[b0] : 0u, h1 = b1 >= 0 ? hist[b1] : 0u;
                unsigned incl = h0 + h1;
                const int lane = tid & 31, wp = tid >> 5;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += u; }
                if (lane == 31) s_wsum[wp] = incl;
                __syncthreads();
                if (wp == 0) {
                    unsigned v = s_wsum[lane];
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += u; }
                    s_wsum[lane] = v;
                }
                __syncthreads();
                const unsigned before = incl - (h0 + h1) + (wp ? s_wsum[wp - 1] : 0u);   // keys in bins above b0
                const unsigned rem = s_remaining, total = s_wsum[31];
                __syncthreads();
                int bin = -1; unsigned cum = 0;
                if (b0 >= 0 && before < rem && before + h0 >= rem) { bin = b0; cum = before; }
                else if (b1 >= 0 && before + h0 < rem && before + h0 + h1 >= rem) { bin = b1; cum = before + h0; }
                if (total < rem && tid == 0) { bin = 0; cum = total - hist[0]; }       // cannot happen (n > nfeatures), kept for safety
                if (bin >= 0) {
                    s_remaining = rem - cum;
                    s_prefix = prefix | ((unsigned)bin << shifts[pass]);
                    s_mask = mask | (((1u << bitsn[pass]) - 1u) << shifts[pass]);
                }
            }
            __syncthreads();
        }
        thr_bits = s_prefix;     // exact bits of the nfeatures-th largest response
    }
    if (tid == 0) s_count = 0;
    __syncthreads();
    for (int i0 = tid; i0 < n; i0 += 8 * blockDim.x) {
        unsigned b[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int i = i0 + u * blockDim.x; b[u] = i < n ? __float_as_uint(__ldg(kresp + i)) : 0u; }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + u * blockDim.x;
            if (i < n && b[u] >= thr_bits) { const int p = atomicAdd(&s_count, 1); if (p < BM_KP_CAP) sel[p] = i; }
        }
    }
    __syncthreads();
    if (tid == 0) { int m = s_count; if (m > BM_KP_CAP) { m = BM_KP_CAP; ctr[2] = 1; } *out_count = m; if (out_thr) *out_thr = thr_bits; }
}

// ------------------------------------------------------------------------------------------------------------------
// cv2's SIFT output order (frames with <= order_cand_cap candidates, i.e. every real clip we have; richer frames keep the
// KeyPoint_LessThan order below).  SIFT_Impl::detectAndCompute does, on ALL keypoints of the frame:
//   KeyPointsFilter::removeDuplicatedSorted  -> std::sort by (x, y, size desc, angle, response desc, octave desc)
//   KeyPointsFilter::retainBest(nfeatures)   -> std::nth_element + std::partition by response (cvorder.cuh)
// One CTA: bitonic sort of (x, index) pairs in shared memory with the full comparator on ties, then the retainBest emulation on the
// responses in that order.  Duplicates cannot occur here (the claim bitmap of k_sift_refine removed them), so the "remove" half of
// removeDuplicatedSorted is a no-op.
// ------------------------------------------------------------------------------------------------------------------
#define SIFT_ORDER_MAX_KP 16384
#define SIFT_ORDER_SMEM (SIFT_ORDER_MAX_KP * 8 + 64 * 1024)
__device__ __forceinline__ bool sift_kp_before(const float2* __restrict__ kpt, const float* __restrict__ ksize, const float* __restrict__ kangle,
                                               const float* __restrict__ kresp, const int* __restrict__ koct, float xa, int a, float xb, int b) {
    if (xa != xb) return xa < xb;
    if (a == b) return false;
    if (a < 0 || b < 0) return b < 0 && a >= 0;                      // padding sorts last
    const float ya = kpt[a].y, yb = kpt[b].y;
    if (ya != yb) return ya < yb;
    if (ksize[a] != ksize[b]) return ksize[a] > ksize[b];
    if (kangle[a] != kangle[b]) return kangle[a] < kangle[b];
    if (kresp[a] != kresp[b]) return kresp[a] > kresp[b];
    if (koct[a] != koct[b]) return koct[a] > koct[b];
    return a < b;
}

__global__ void __launch_bounds__(CVO_THREADS) k_sift_order(SiftLayout lay, int nfeatures, int* __restrict__ ctr, const float2* __restrict__ kpt,
                                                            const float* __restrict__ ksize, const float* __restrict__ kangle,
                                                            const float* __restrict__ kresp, const int* __restrict__ koct, int* __restrict__ sel,
                                                            int* __restrict__ scratch_idx) {
    extern __shared__ __align__(16) unsigned char ord_smem[];
    __shared__ CvoShared sh;
    int ncand = ctr[0]; if (ncand > lay.cand_cap) ncand = lay.cand_cap;
    const int n = ctr[1];
    if (ncand > lay.order_cand_cap || n > SIFT_ORDER_MAX_KP || n > lay.kp_cap) return;      // k_sift_select + the sorting emit take over
    const int tid = threadIdx.x;
    float* sx = reinterpret_cast<float*>(ord_smem);                      // [np]
    int* si = reinterpret_cast<int*>(ord_smem) + SIFT_ORDER_MAX_KP;       // [np]
    int np = 1; while (np < n) np <<= 1;
    for (int i = tid; i < np; i += CVO_THREADS) { sx[i] = i < n ? kpt[i].x : __int_as_float(0x7f800000); si[i] = i < n ? i : -1; }
    __syncthreads();
    for (int k = 2; k <= np; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < np / 2; t += CVO_THREADS) {
                const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;
                const bool up = (lo & k) == 0;
                const float xa = sx[lo], xb = sx[hi]; const int a = si[lo], b = si[hi];
                const bool swap = up ? sift_kp_before(kpt, ksize, kangle, kresp, koct, xb, b, xa, a) : sift_kp_before(kpt, ksize, kangle, kresp, koct, xa, a, xb, b);
                if (swap) { sx[lo] = xb; sx[hi] = xa; si[lo] = b; si[hi] = a; }
            }
            __syncthreads();
        }
    // sorted order -> keys (responses) for retainBest; the selection permutes scratch_idx (positions in the sorted list)
    float* keys = sx;                                                    // reuse: x is no longer needed
    for (int i = tid; i < n; i += CVO_THREADS) { keys[i] = kresp[si[i]]; scratch_idx[i] = i; }
    __syncthreads();
    CvoRows rows;
    rows.bind(ord_smem + (size_t)SIFT_ORDER_MAX_KP * 8, cvo_rows_needed(n));
    const int m = cvo_retain_best<float>(keys, scratch_idx, n, nfeatures, sh, rows);
    const int mm = m > BM_KP_CAP ? BM_KP_CAP : m;
    for (int i = tid; i < mm; i += CVO_THREADS) sel[i] = si[scratch_idx[i]];
    if (tid == 0) { ctr[3] = mm; ctr[8] = 1; if (m > BM_KP_CAP) ctr[2] = 1; }
}

__device__ __forceinline__ bool kp_less(float ax, float ay, float as, float aa, float ar, int ao, int ai,
                                        float bx, float by, float bs, float ba, float br, int bo, int bi) {
    if (ax != bx) return ax < bx;
    if (ay != by) return ay < by;
    if (as != bs) return as > bs;
    if (aa != ba) return aa < ba;
    if (ar != br) return ar > br;
    if (ao != bo) return ao > bo;
    return ai < bi;
}

// order the selected keypoints (cv2's KeyPoint_LessThan), rescale to the input image (first octave -1), emit
__global__ void __launch_bounds__(256) k_sift_emit(const int* __restrict__ ctr, const int* __restrict__ sel, const float2* __restrict__ kpt,
                                                    const float* __restrict__ ksize, const float* __restrict__ kangle, const float* __restrict__ kresp,
                                                    const int* __restrict__ koct, BmKeypoints out) {
    const int m = ctr[3];
    const bool ordered = ctr[8] != 0;                          // sel[] is already in cv2's retainBest order (k_sift_order)
    __shared__ float s_x[1024], s_y[1024], s_s[1024], s_a[1024], s_r[1024];
    __shared__ int s_o[1024], s_i[1024];
    if ((int)(blockIdx.x * blockDim.x) >= m && blockIdx.x != 0) return;      // CTA b ranks keypoints [256 b, 256 b + 256)
    {
        const int a = blockIdx.x * blockDim.x + threadIdx.x;
        const int i = a < m ? sel[a] : 0;
        float2 pi = make_float2(0.f, 0.f); float si = 0.f, ai = 0.f, ri = 0.f; int oi = 0;
        if (a < m) { pi = kpt[i]; si = ksize[i]; ai = kangle[i]; ri = kresp[i]; oi = koct[i]; }
        int rank = ordered ? a : 0;
        for (int b0 = 0; b0 < (ordered ? 0 : m); b0 += 1024) {  // rank = number of selected keypoints ordered before this one
            __syncthreads();
            for (int b = b0 + threadIdx.x; b < min(m, b0 + 1024); b += blockDim.x) {
                const int j = sel[b], q = b - b0;
                const float2 pj = kpt[j];
                s_x[q] = pj.x; s_y[q] = pj.y; s_s[q] = ksize[j]; s_a[q] = kangle[j]; s_r[q] = kresp[j]; s_o[q] = koct[j]; s_i[q] = j;
            }
            __syncthreads();
            const int nb = min(1024, m - b0);
            if (a < m) {
                // the primary key (x) almost never ties: count on it alone, and run the full comparison only over the ties
                int ties = 0;
                for (int q = 0; q < nb; ++q) { const float xq = s_x[q]; rank += xq < pi.x ? 1 : 0; ties += xq == pi.x ? 1 : 0; }
                if (ties > (b0 <= a && a < b0 + nb ? 1 : 0))
                    for (int q = 0; q < nb; ++q)
                        if (s_x[q] == pi.x) rank += kp_less(s_x[q], s_y[q], s_s[q], s_a[q], s_r[q], s_o[q], s_i[q], pi.x, pi.y, si, ai, ri, oi, i) ? 1 : 0;
            }
        }
        if (a < m) {
            // firstOctave = -1: octave byte -1, pt and size halved
            const int oc = (oi & ~255) | ((oi - 1) & 255);
            out.pt[rank] = make_float2(pi.x * 0.5f, pi.y * 0.5f);
            out.size[rank] = si * 0.5f;
            out.angle[rank] = ai;
            out.response[rank] = ri;
            out.octave[rank] = oc;
            out.lxy[rank] = make_int2(0, 0);
        }
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) { *out.count = m; out.flags[0] = ctr[2]; }   // ctr[2]: a list overflowed somewhere upstream
}

// ------------------------------------------------------------------------------------------------------------------
// calcSIFTDescriptor: one CTA per keypoint
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sift_describe(SiftLayout lay, const float* __restrict__ pyr, BmKeypoints kp) {
    // 2^-24 fixed-point histogram (integer sums are order independent -> deterministic).  64-bit shared atomics compile to a
    // compare-and-swap spin loop, so a bin is two 32-bit words updated with native 32-bit adds and an explicit carry.
    __shared__ unsigned hist_lo[6 * 6 * 10], hist_hi[6 * 6 * 10];
    __shared__ float raw[128];
    __shared__ float red[8];
    const int tid = threadIdx.x;
  for (int k = blockIdx.x; k < *kp.count; k += gridDim.x) {
    __syncthreads();                                       // previous keypoint of this CTA is done with the shared arrays
    for (int i = tid; i < 360; i += 256) { hist_lo[i] = 0u; hist_hi[i] = 0u; }
    const int packed = kp.octave[k];
    int octave = packed & 255; const int layer = (packed >> 8) & 255;
    octave = octave < 128 ? octave : (-128 | octave);
    const float scale = octave >= 0 ? 1.f / (float)(1 << octave) : (float)(1 << -octave);
    const float size = kp.size[k] * scale;
    const float2 p0 = kp.pt[k];
    const float ptfx = p0.x * scale, ptfy = p0.y * scale;
    const SiftOct O = lay.o[octave + 1];
    const float* img = pyr + O.g[layer];
    const int cols = O.w, rows = O.h;
    float ori = 360.f - kp.angle[k];
    if (fabsf(ori - 360.f) < FLT_EPSILON) ori = 0.f;
    const float scl = size * 0.5f;
    const int ptx = __float2int_rn(ptfx), pty = __float2int_rn(ptfy);
    float cos_t = cosf(ori * (float)(3.14159265358979323846 / 180.0)), sin_t = sinf(ori * (float)(3.14159265358979323846 / 180.0));
    const float bins_per_rad = 8.f / 360.f, exp_scale = -1.f / (4.f * 4.f * 0.5f), hist_width = 3.f * scl;
    int radius = __float2int_rn(hist_width * 1.4142135623730951f * 5.f * 0.5f);
    radius = min(radius, (int)sqrt((double)cols * cols + (double)rows * rows));
    cos_t /= hist_width; sin_t /= hist_width;
    __syncthreads();
    const int side = 2 * radius + 1, len = side * side;
    for (int q = tid; q < len; q += 256) {
        const int i = q / side - radius, j = q % side - radius;
        const float c_rot = j * cos_t - i * sin_t, r_rot = j * sin_t + i * cos_t;
        float rbin = r_rot + 2.f - 0.5f, cbin = c_rot + 2.f - 0.5f;
        const int r = pty + i, c = ptx + j;
        if (!(rbin > -1 && rbin < 4 && cbin > -1 && cbin < 4 && r > 0 && r < rows - 1 && c > 0 && c < cols - 1)) continue;
        const float dx = img[(size_t)r * cols + c + 1] - img[(size_t)r * cols + c - 1];
        const float dy = img[(size_t)(r - 1) * cols + c] - img[(size_t)(r + 1) * cols + c];
        const float wgt = expf((c_rot * c_rot + r_rot * r_rot) * exp_scale);
        const float o_deg = sift_atan2_deg(dy, dx);
        const float mag = sqrtf(dx * dx + dy * dy) * wgt;
        float obin = (o_deg - ori) * bins_per_rad;
        const int r0 = (int)floorf(rbin), c0 = (int)floorf(cbin); int o0 = (int)floorf(obin);
        rbin -= r0; cbin -= c0; obin -= o0;
        if (o0 < 0) o0 += 8;
        if (o0 >= 8) o0 -= 8;
        const float v_r1 = mag * rbin, v_r0 = mag - v_r1;
        const float v_rc11 = v_r1 * cbin, v_rc10 = v_r1 - v_rc11, v_rc01 = v_r0 * cbin, v_rc00 = v_r0 - v_rc01;
        const float v111 = v_rc11 * obin, v110 = v_rc11 - v111, v101 = v_rc10 * obin, v100 = v_rc10 - v101;
        const float v011 = v_rc01 * obin, v010 = v_rc01 - v011, v001 = v_rc00 * obin, v000 = v_rc00 - v001;
        const int idx = ((r0 + 1) * 6 + c0 + 1) * 10 + o0;
        const float FX = 16777216.f;
#define HADD(off, v) do { const unsigned long long _q = (unsigned long long)__float2ll_rn((v) * FX); const unsigned _l = (unsigned)_q; \
                          const unsigned _o = atomicAdd(&hist_lo[idx + (off)], _l); const unsigned _h = (unsigned)(_q >> 32) + ((_o + _l < _o) ? 1u : 0u); \
                          if (_h) atomicAdd(&hist_hi[idx + (off)], _h); } while (0)
        HADD(0, v000); HADD(1, v001); HADD(10, v010); HADD(11, v011); HADD(60, v100); HADD(61, v101); HADD(70, v110); HADD(71, v111);
#undef HADD
    }
    __syncthreads();
    // fold the circular orientation bins, gather the 4x4x8 core
    if (tid < 128) {
        const int i = tid >> 5, j = (tid >> 3) & 3, o = tid & 7;
        const int idx = ((i + 1) * 6 + (j + 1)) * 10;
        auto bin = [&](int q) { return (long long)(((unsigned long long)hist_hi[q] << 32) | hist_lo[q]); };
        long long v = bin(idx + o);
        if (o == 0) v += bin(idx + 8);
        if (o == 1) v += bin(idx + 9);
        raw[tid] = (float)((double)v * (1.0 / 16777216.0));
    }
    __syncthreads();
    // norm -> clip at 0.2*norm -> renormalise to 512 -> saturate_cast<uchar>
    float v = tid < 128 ? raw[tid] : 0.f;
    float s = v * v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    const float nrm2 = red[0] + red[1] + red[2] + red[3];
    const float thr = sqrtf(nrm2) * 0.2f;
    __syncthreads();
    v = fminf(v, thr);
    s = tid < 128 ? v * v : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((tid & 31) == 0) red[tid >> 5] = s;
    __syncthreads();
    const float n2 = 512.f / fmaxf(sqrtf(red[0] + red[1] + red[2] + red[3]), FLT_EPSILON);
    if (tid < 128) {
        const int q = __float2int_rn(v * n2);
        kp.desc[(size_t)k * 128 + tid] = (uint8_t)max(0, min(255, q));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------------------------
static void gaussian_kernel(int ksize, double sigma, float* out) {
    // cv::getGaussianKernel(ksize, sigma, CV_32F): exp and normalisation in double, cast to float
    std::vector<double> k(ksize);
    double sum = 0;
    for (int i = 0; i < ksize; ++i) { const double x = i - (ksize - 1) * 0.5; k[i] = exp(-0.5 * x * x / (sigma * sigma)); sum += k[i]; }
    for (int i = 0; i < ksize; ++i) out[i] = (float)(k[i] / sum);
}

static void sift_make_tensor_maps(BmSift* o);

int bm_sift_create(BmSift** out, int h, int w, int nfeatures, cudaStream_t s) {
    BmSift* o = new (std::nothrow) BmSift();
    if (!o) return -1;
    memset(o, 0, sizeof(*o));
    o->w = w; o->h = h; o->nfeatures = nfeatures; o->stream = s;
    const int bw = 2 * w, bh = 2 * h;
    if (bw > 8191 || bh > 8191) { bm_set_error("SIFT: frames larger than 4095 px are not supported"); delete o; return -1; }   // raw extrema pack r, c in 13 bits
    int noct = (int)nearbyint(log((double)(bw < bh ? bw : bh)) / log(2.0) - 2.0) + 1;
    if (noct < 1) noct = 1;
    if (noct > SIFT_MAX_OCT) noct = SIFT_MAX_OCT;
    o->lay.noct = noct;
    {
        const long long N = (long long)w * h;
        o->lay.raw_cap = (int)(4 * N < (1LL << 21) ? (1LL << 21) : 4 * N);
        o->lay.cand_cap = o->lay.kp_cap = (int)(N / 2 < (1LL << 17) ? (1LL << 17) : N / 2);
        // frames with at most this many candidates get cv2's exact retainBest order (needs an orientation histogram for every
        // candidate instead of the ~nfeatures best: ~3.5 us per thousand); BM_SIFT_ORDER_CAP overrides, 0 = never
        o->lay.order_cand_cap = 12288;
        if (const char* e = getenv("BM_SIFT_ORDER_CAP")) o->lay.order_cand_cap = atoi(e);
    }
    long long off = 0, cbits = 0;
    int ow = bw, oh = bh;
    for (int i = 0; i < noct; ++i) {
        SiftOct& O = o->lay.o[i];
        O.w = ow; O.h = oh;
        const long long n = ((long long)ow * oh + 63) & ~63LL;
        for (int l = 0; l < 6; ++l) { O.g[l] = off; off += n; }
        for (int l = 0; l < 5; ++l) { O.d[l] = off; off += n; }
        O.claim = cbits; cbits += 3LL * ow * oh;
        ow /= 2; oh /= 2;
        if (ow < 1 || oh < 1) { o->lay.noct = i + 1; break; }
    }
    o->claim_words = (size_t)((cbits + 31) / 32 + 1);
    float hk[6][32]; memset(hk, 0, sizeof(hk));
    {
        // createInitialImage: sig_diff = sqrtf(max(sigma^2 - (2*0.5)^2, 0.01f)); buildGaussianPyramid: incremental sigmas
        const float sig_diff = sqrtf(fmaxf(1.6f * 1.6f - 0.5f * 0.5f * 4.f, 0.01f));
        gaussian_kernel(h_sift_ksize[0], (double)sig_diff, hk[0]);
        const double k = pow(2.0, 1.0 / 3.0);
        for (int i = 1; i <= 5; ++i) {
            const double sig_prev = pow(k, (double)(i - 1)) * 1.6, sig_total = sig_prev * k;
            const double sg = sqrt(sig_total * sig_total - sig_prev * sig_prev);
            int ks = (int)nearbyint(sg * 8 + 1) | 1;
            if (ks != h_sift_ksize[i]) { bm_set_error("unexpected SIFT kernel size %d at level %d", ks, i); delete o; return -1; }
            gaussian_kernel(ks, sg, hk[i]);
        }
    }
    bool ok = cudaMalloc(&o->pyr, (size_t)off * sizeof(float)) == cudaSuccess && cudaMalloc(&o->up, (size_t)bw * bh * sizeof(float)) == cudaSuccess &&
              cudaMalloc(&o->claim, o->claim_words * 4) == cudaSuccess && cudaMalloc(&o->cand, (size_t)o->lay.cand_cap * sizeof(SiftCand)) == cudaSuccess &&
              cudaMalloc(&o->ctr, 16 * sizeof(int)) == cudaSuccess && cudaMalloc(&o->kpt, (size_t)o->lay.kp_cap * sizeof(float2)) == cudaSuccess &&
              cudaMalloc(&o->ksize, (size_t)o->lay.kp_cap * 4) == cudaSuccess && cudaMalloc(&o->kangle, (size_t)o->lay.kp_cap * 4) == cudaSuccess &&
              cudaMalloc(&o->kresp, (size_t)o->lay.kp_cap * 4) == cudaSuccess && cudaMalloc(&o->koct, (size_t)o->lay.kp_cap * 4) == cudaSuccess &&
              cudaMalloc(&o->sel, BM_KP_CAP * 4) == cudaSuccess && cudaMalloc(&o->raw, (size_t)o->lay.raw_cap * 4) == cudaSuccess &&
              cudaMalloc(&o->order_idx, (size_t)SIFT_ORDER_MAX_KP * 4) == cudaSuccess && cudaMalloc(&o->cresp, (size_t)o->lay.cand_cap * 4) == cudaSuccess && cudaMalloc(&o->csel, BM_KP_CAP * 4) == cudaSuccess;
    if (ok) ok = cudaMemcpyToSymbol(c_sift_k, hk, sizeof(hk)) == cudaSuccess;
    if (ok) ok = bm_stream_create(&o->s2, 1) == cudaSuccess && bm_stream_create(&o->s3, 1) == cudaSuccess &&
                 cudaEventCreateWithFlags(&o->ev_fork, cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&o->ev_join2, cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&o->ev_join3, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; ok && i < SIFT_MAX_OCT; ++i)
        ok = cudaEventCreateWithFlags(&o->ev_l3[i], cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&o->ev_l5[i], cudaEventDisableTiming) == cudaSuccess;
    o->graphs_enabled = true;
    if (ok) sift_make_tensor_maps(o);
    if (ok) { cudaError_t ea; BM_SMEM_OPTIN(k_sift_order, SIFT_ORDER_SMEM, ea); ok = ea == cudaSuccess; }
    if (!ok) { bm_set_error("bm_sift_create: %s", cudaGetErrorString(cudaGetLastError())); bm_sift_destroy(o); return -1; }
    *out = o;
    return 0;
}

void bm_sift_destroy(BmSift* o) {
    if (!o) return;
    for (int i = 0; i < o->ngraphs; ++i) cudaGraphExecDestroy(o->graphs[i].exec);
    if (o->s2) cudaStreamDestroy(o->s2);
    if (o->s3) cudaStreamDestroy(o->s3);
    if (o->ev_fork) cudaEventDestroy(o->ev_fork);
    if (o->ev_join2) cudaEventDestroy(o->ev_join2);
    if (o->ev_join3) cudaEventDestroy(o->ev_join3);
    for (int i = 0; i < SIFT_MAX_OCT; ++i) { if (o->ev_l3[i]) cudaEventDestroy(o->ev_l3[i]); if (o->ev_l5[i]) cudaEventDestroy(o->ev_l5[i]); }
    cudaFree(o->pyr); cudaFree(o->up); cudaFree(o->claim); cudaFree(o->cand); cudaFree(o->ctr); cudaFree(o->kpt); cudaFree(o->ksize);
    cudaFree(o->kangle); cudaFree(o->kresp); cudaFree(o->koct); cudaFree(o->sel); cudaFree(o->raw); cudaFree(o->cresp); cudaFree(o->csel); cudaFree(o->order_idx);
    delete o;
}

__host__ __device__ constexpr int sift_blur_sht(int level, long long px) { return (level >= 4 && px >= (1 << 21)) ? 128 : 64; }   // octave 0 of a 1080p frame (3840 x 2160)

template <int LEVEL, int SHT>
static void launch_blur_sh(const CUtensorMap* tm, const float* in, float* out, float* dog, float* dec, int w, int h, cudaStream_t s) {
    cudaError_t attr;
    constexpr int TH = sift_tile_h(LEVEL, SHT);
    const dim3 grid((w + 63) / 64, (h + TH - 1) / TH);
    if (tm) {
        BM_SMEM_OPTIN((k_sift_blur_tma<LEVEL, SHT>), sift_blur_tma_smem(SHT), attr);
        (void)attr;
        BM_COUNT_LAUNCHES(1), k_sift_blur_tma<LEVEL, SHT><<<grid, sift_blur_nt(SHT), sift_blur_tma_smem(SHT), s>>>(*tm, in, out, dog, dec, w, h);
        return;
    }
    BM_SMEM_OPTIN((k_sift_blur<LEVEL, SHT>), sift_blur_smem(SHT), attr);
    (void)attr;                                    // a failure surfaces as the launch error picked up by the caller
    BM_COUNT_LAUNCHES(1), k_sift_blur<LEVEL, SHT><<<grid, sift_blur_nt(SHT), sift_blur_smem(SHT), s>>>(in, out, dog, dec, w, h);
}
template <int LEVEL>
static void launch_blur(const CUtensorMap* tm, const float* in, float* out, float* dog, float* dec, int w, int h, cudaStream_t s) {
    if (sift_blur_sht(LEVEL, (long long)w * h) == 128) launch_blur_sh<LEVEL, 128>(tm, in, out, dog, dec, w, h, s);
    else launch_blur_sh<LEVEL, 64>(tm, in, out, dog, dec, w, h, s);
}

// level 0 of octave 0 straight from the gray frame (k_sift_blur<0, 64, true>)
static void blur_base_from_gray(const uint8_t* d_gray, float* out, int w, int h, cudaStream_t s) {
    cudaError_t attr;
    constexpr int TH = sift_tile_h(0, 64);
    BM_SMEM_OPTIN((k_sift_blur<0, 64, true>), sift_blur_smem(64), attr);
    (void)attr;
    BM_COUNT_LAUNCHES(1), k_sift_blur<0, 64, true><<<dim3((w + 63) / 64, (h + TH - 1) / TH), sift_blur_nt(64), sift_blur_smem(64), s>>>(nullptr, out, nullptr, nullptr, w, h, d_gray);
}

static void blur_level(const BmSift* o, int oc, int level, const float* in, float* out, float* dog, float* dec, int w, int h, cudaStream_t s) {
    const CUtensorMap* tm = (o->use_tma && o->tma_ok[oc][level]) ? &o->tmap[oc][level] : nullptr;
    switch (level) {
        case 0: launch_blur<0>(tm, in, out, dog, dec, w, h, s); break;
        case 1: launch_blur<1>(tm, in, out, dog, dec, w, h, s); break;
        case 2: launch_blur<2>(tm, in, out, dog, dec, w, h, s); break;
        case 3: launch_blur<3>(tm, in, out, dog, dec, w, h, s); break;
        case 4: launch_blur<4>(tm, in, out, dog, dec, w, h, s); break;
        default: launch_blur<5>(tm, in, out, dog, dec, w, h, s); break;
    }
}

// TMA descriptors of every blur source: 2-D float tensor (w, h), box 32 x SHT, SWIZZLE_128B, zero fill outside (only tiles whose box
// lies inside the image use it).  A level whose rows are not 16-byte multiples keeps the per-element path.
// cuTensorMapEncodeTiled is resolved through the runtime (cudaGetDriverEntryPoint) instead of linking libcuda: the library must load
// on a machine without a driver (the CPU-side tests check its exports there)
typedef CUresult (*bm_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                       const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bm_encode_tiled_fn sift_encode_tiled() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) { cudaGetLastError(); return nullptr; }
    return reinterpret_cast<bm_encode_tiled_fn>(fn);
}

static void sift_make_tensor_maps(BmSift* o) {
    // A/B (profiles/r02_pyramid_tma_ab.md): the TMA form is not faster than the cp.async form at 1080p (the kernel is bound by the
    // FMA / LDS issue of its two passes, not by the tile fill), so it is opt-in: BM_SIFT_TMA=1
    o->use_tma = getenv("BM_SIFT_TMA") != nullptr && getenv("BM_SIFT_NO_TMA") == nullptr;
    { const char* e = getenv("BM_SIFT_SEPARATE_UPSAMPLE"); o->separate_upsample = e && e[0] == '1'; }
    const bm_encode_tiled_fn encode = o->use_tma ? sift_encode_tiled() : nullptr;
    if (!encode) o->use_tma = false;
    for (int oc = 0; oc < o->lay.noct; ++oc) {
        const SiftOct& O = o->lay.o[oc];
        for (int l = 0; l < 6; ++l) {
            o->tma_ok[oc][l] = false;
            if (!encode) continue;
            const float* src = l == 0 ? (oc == 0 ? o->up : nullptr) : o->pyr + O.g[l - 1];
            const int sht = sift_blur_sht(l, (long long)O.w * O.h);
            if (!src || (O.w & 3) || O.w < 96 || O.h < sht || (reinterpret_cast<uintptr_t>(src) & 15)) continue;
            const cuuint64_t gdim[2] = {(cuuint64_t)O.w, (cuuint64_t)O.h};
            const cuuint64_t gstr[1] = {(cuuint64_t)O.w * sizeof(float)};
            const cuuint32_t box[2] = {32u, (cuuint32_t)sht};
            const cuuint32_t estr[2] = {1u, 1u};
            const CUresult r = encode(&o->tmap[oc][l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(src), gdim, gstr, box, estr,
                                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            o->tma_ok[oc][l] = r == CUDA_SUCCESS;
            if (const char* only = getenv("BM_SIFT_TMA_ONLY")) {               // debug: "oc,level" enables the TMA form for one launch only
                int a = -1, b = -1;
                if (sscanf(only, "%d,%d", &a, &b) == 2 && (a != oc || b != l)) o->tma_ok[oc][l] = false;
            }
        }
    }
}

const float* bm_sift_level_ptr(BmSift* o, int octave, int level, int dog, int* w, int* h) {
    const SiftOct& O = o->lay.o[octave];
    if (w) *w = O.w;
    if (h) *h = O.h;
    return o->pyr + (dog ? O.d[level] : O.g[level]);
}
int bm_sift_num_octaves(BmSift* o) { return o->lay.noct; }
void bm_sift_counters(BmSift* o, int out[8]) { cudaStreamSynchronize(o->stream); cudaMemcpy(out, o->ctr, 8 * sizeof(int), cudaMemcpyDeviceToHost); }

// The whole detectAndCompute as one enqueue.  With `forked`, levels 4-5 of every octave run on a second stream and the
// extrema scans on a third, so that the dependency chain is only  level 1 -> 2 -> 3 (-> next octave's base)  per octave:
// the seven smallest octaves are one-CTA kernels whose ~4 us latencies would otherwise add up serially.
static cudaError_t sift_enqueue(BmSift* o, const uint8_t* d_gray, BmKeypoints* out, bool forked) {
    cudaStream_t s = o->stream, s2 = forked ? o->s2 : s, s3 = forked ? o->s3 : s;
    const SiftLayout& L = o->lay;
    cudaError_t e;
#define SIFT_OK(x) do { if ((e = (x)) != cudaSuccess) return e; } while (0)
    SIFT_OK(cudaMemsetAsync(o->ctr, 0, 16 * sizeof(int), s));
    if (forked) { SIFT_OK(cudaEventRecord(o->ev_fork, s)); SIFT_OK(cudaStreamWaitEvent(s3, o->ev_fork, 0)); }
    SIFT_OK(cudaMemsetAsync(o->claim, 0, o->claim_words * 4, s3));
    const dim3 blk(32, 8);
    const int bw = 2 * o->w, bh = 2 * o->h;
    const bool fuse_up = !(o->use_tma && o->tma_ok[0][0]) && !o->separate_upsample;
    if (!fuse_up) BM_COUNT_LAUNCHES(1), k_sift_upsample<<<dim3((bw + 31) / 32, (bh + 31) / 32), blk, 0, s>>>(d_gray, o->w, o->h, o->up);
    for (int oc = 0; oc < L.noct; ++oc) {
        const SiftOct& O = L.o[oc];
        if (oc == 0 && fuse_up) blur_base_from_gray(d_gray, o->pyr + O.g[0], O.w, O.h, s);
        else if (oc == 0) blur_level(o, oc, 0, o->up, o->pyr + O.g[0], nullptr, nullptr, O.w, O.h, s);
        // level 3 also writes the next octave's base (its 2x decimation)
        for (int l = 1; l <= 3; ++l)
            blur_level(o, oc, l, o->pyr + O.g[l - 1], o->pyr + O.g[l], o->pyr + O.d[l - 1], (l == 3 && oc + 1 < L.noct) ? o->pyr + L.o[oc + 1].g[0] : nullptr, O.w, O.h, s);
        if (forked) { SIFT_OK(cudaEventRecord(o->ev_l3[oc], s)); SIFT_OK(cudaStreamWaitEvent(s2, o->ev_l3[oc], 0)); }
        for (int l = 4; l <= 5; ++l) blur_level(o, oc, l, o->pyr + O.g[l - 1], o->pyr + O.g[l], o->pyr + O.d[l - 1], nullptr, O.w, O.h, s2);
        if (O.w <= 2 * SIFT_BORDER || O.h <= 2 * SIFT_BORDER) continue;
        if (forked) { SIFT_OK(cudaEventRecord(o->ev_l5[oc], s2)); SIFT_OK(cudaStreamWaitEvent(s3, o->ev_l5[oc], 0)); }
        bool tiled = O.w >= EX_TW && (O.w & 3) == 0 && (reinterpret_cast<uintptr_t>(o->pyr) & 15) == 0;
        for (int l = 0; l < 5; ++l) tiled = tiled && (O.d[l] & 3) == 0 && O.d[l] == O.d[0] + (long long)l * O.w * O.h;
        if (tiled) {
            const int tx = (O.w + EX_TW - 1) / EX_TW, ty = (O.h + EX_TH - 1) / EX_TH;
            const int kt = tx * ty >= 8000 ? 4 : tx * ty >= 2000 ? 2 : 1;      // tiles walked by one CTA (keeps >= 2 waves of CTAs)
            BM_COUNT_LAUNCHES(1), k_sift_extrema_tile<<<dim3(tx, (ty + kt - 1) / kt), 256, 0, s3>>>(L, oc, kt, o->pyr, o->raw, o->ctr);
        }
        else BM_COUNT_LAUNCHES(1), k_sift_extrema<<<dim3((O.w + 31) / 32, (O.h + 7) / 8), blk, 0, s3>>>(L, oc, o->pyr, o->raw, o->ctr);
    }
    if (forked) {
        SIFT_OK(cudaEventRecord(o->ev_join2, s2)); SIFT_OK(cudaEventRecord(o->ev_join3, s3));
        SIFT_OK(cudaStreamWaitEvent(s, o->ev_join2, 0)); SIFT_OK(cudaStreamWaitEvent(s, o->ev_join3, 0));
    }
    BM_COUNT_LAUNCHES(1), k_sift_refine<<<148 * 8, 128, 0, s>>>(L, o->pyr, o->raw, o->claim, o->cand, o->cresp, o->ctr);
    BM_COUNT_LAUNCHES(1), k_sift_select<<<1, 1024, 0, s>>>(o->nfeatures, o->ctr + 0, L.cand_cap, o->ctr, o->cresp, o->csel, o->ctr + 7, (unsigned*)(o->ctr + 5));
    BM_COUNT_LAUNCHES(1), k_sift_orient<<<512, 32 * SIFT_ORI_WARPS, 0, s>>>(L, o->pyr, o->cand, o->csel, 0, o->nfeatures, o->ctr, o->kpt, o->ksize, o->kangle, o->kresp, o->koct);
    SIFT_OK(cudaMemcpyAsync(o->ctr + 6, o->ctr + 1, sizeof(int), cudaMemcpyDeviceToDevice, s));
    BM_COUNT_LAUNCHES(1), k_sift_orient<<<1024, 32 * SIFT_ORI_WARPS, 0, s>>>(L, o->pyr, o->cand, o->csel, 1, o->nfeatures, o->ctr, o->kpt, o->ksize, o->kangle, o->kresp, o->koct);
    BM_COUNT_LAUNCHES(1), k_sift_orient<<<1024, 32 * SIFT_ORI_WARPS, 0, s>>>(L, o->pyr, o->cand, o->csel, 2, o->nfeatures, o->ctr, o->kpt, o->ksize, o->kangle, o->kresp, o->koct);
    BM_COUNT_LAUNCHES(1), k_sift_order<<<1, CVO_THREADS, SIFT_ORDER_SMEM, s>>>(L, o->nfeatures, o->ctr, o->kpt, o->ksize, o->kangle, o->kresp, o->koct, o->sel, o->order_idx);
    BM_COUNT_LAUNCHES(1), k_sift_select<<<1, 1024, 0, s>>>(o->nfeatures, o->ctr + 1, L.kp_cap, o->ctr, o->kresp, o->sel, o->ctr + 3, nullptr);
    BM_COUNT_LAUNCHES(1), k_sift_emit<<<BM_KP_CAP / 256, 256, 0, s>>>(o->ctr, o->sel, o->kpt, o->ksize, o->kangle, o->kresp, o->koct, *out);
    BM_COUNT_LAUNCHES(1), k_sift_describe<<<1024, 256, 0, s>>>(L, o->pyr, *out);
#undef SIFT_OK
    return cudaGetLastError();
}

// Gaussian + DoG pyramid alone (upsample + all blur levels, serially on the detector's stream), `reps` times between two CUDA events:
// the figure behind bench.py's `roofline_pyramid` (256 N algorithmic bytes per frame).
cudaError_t bm_sift_time_pyramid(BmSift* o, const uint8_t* d_gray, int reps, float* ms_total) {
    cudaStream_t s = o->stream;
    const SiftLayout& L = o->lay;
    cudaEvent_t e0, e1;
    cudaError_t e = cudaEventCreate(&e0);
    if (e != cudaSuccess) return e;
    if ((e = cudaEventCreate(&e1)) != cudaSuccess) { cudaEventDestroy(e0); return e; }
    const dim3 blk(32, 8);
    const int bw = 2 * o->w, bh = 2 * o->h;
    for (int r = -1; r < reps; ++r) {                      // r == -1: warm-up
        if (r == 0) cudaEventRecord(e0, s);
        const bool fuse_up = !(o->use_tma && o->tma_ok[0][0]) && !o->separate_upsample;
        if (!fuse_up) BM_COUNT_LAUNCHES(1), k_sift_upsample<<<dim3((bw + 31) / 32, (bh + 31) / 32), blk, 0, s>>>(d_gray, o->w, o->h, o->up);
        for (int oc = 0; oc < L.noct; ++oc) {
            const SiftOct& O = L.o[oc];
            if (oc == 0 && fuse_up) blur_base_from_gray(d_gray, o->pyr + O.g[0], O.w, O.h, s);
            else if (oc == 0) blur_level(o, oc, 0, o->up, o->pyr + O.g[0], nullptr, nullptr, O.w, O.h, s);
            for (int l = 1; l <= 5; ++l)
                blur_level(o, oc, l, o->pyr + O.g[l - 1], o->pyr + O.g[l], o->pyr + O.d[l - 1], (l == 3 && oc + 1 < L.noct) ? o->pyr + L.o[oc + 1].g[0] : nullptr, O.w, O.h, s);
        }
    }
    cudaEventRecord(e1, s);
    e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaEventElapsedTime(ms_total, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e == cudaSuccess) e = cudaGetLastError();
    return e;
}

// detectAndCompute is a fixed launch sequence per (input buffer, output buffer): it is captured once into a CUDA graph with the
// fork / join structure above and replayed -- one launch call per frame instead of ~90, and the independent branches overlap.
cudaError_t bm_sift_detect(BmSift* o, const uint8_t* d_gray, BmKeypoints* out, bool launch) {
    if (o->stream == nullptr || !o->graphs_enabled) return sift_enqueue(o, d_gray, out, false);   // legacy stream cannot be captured
    for (int i = 0; i < o->ngraphs; ++i)
        if (o->graphs[i].gray == d_gray && o->graphs[i].out_pt == (const void*)out->pt) {
            if (!launch) return cudaSuccess;
            BM_COUNT_LAUNCHES(o->graphs[i].launches);
            return cudaGraphLaunch(o->graphs[i].exec, o->stream);
        }
    if (o->ngraphs >= BmSift::kMaxGraphs) return sift_enqueue(o, d_gray, out, true);
    long long captured = 0;
    t_bm_launch_sink = &captured;                            // nothing runs during capture: count the nodes, not launches
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(o->stream, cudaStreamCaptureModeRelaxed);
    if (e != cudaSuccess) return e;
    e = sift_enqueue(o, d_gray, out, true);
    const cudaError_t e2 = cudaStreamEndCapture(o->stream, &graph);
    t_bm_launch_sink = nullptr;
    const int launches = (int)captured;
    cudaGraphExec_t exec = nullptr;
    if (e == cudaSuccess && e2 == cudaSuccess) e = cudaGraphInstantiate(&exec, graph, 0);
    else if (e == cudaSuccess) e = e2;
    if (graph) cudaGraphDestroy(graph);
    if (e != cudaSuccess) {                                    // fall back to plain stream launches from now on
        cudaGetLastError();
        o->graphs_enabled = false;
        return sift_enqueue(o, d_gray, out, false);
    }
    SiftGraph& g = o->graphs[o->ngraphs++];
    g.gray = d_gray; g.out_pt = out->pt; g.exec = exec; g.launches = launches;
    if (!launch) return cudaSuccess;                          // capture only (bm_warm_up)
    BM_COUNT_LAUNCHES(launches);
    return cudaGraphLaunch(exec, o->stream);
}
