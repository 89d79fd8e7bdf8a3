// jpeg.cu -- baseline JPEG encoder on the device (the last step of the finalisation): the file
// cv2.imwrite(os.path.join(output_dir, 'mosaic.jpg'), scaled_mosaic) writes (/root/reference/main.py:1664-1665), byte for byte, so that
// only the compressed file crosses PCIe.  cv2 4.13 drives its bundled libjpeg-turbo with the defaults (quality 95, 4:2:0, ISLOW DCT,
// the Annex K Huffman tables, no restart markers); the arithmetic is restated in oracle/jpeg.py and pinned there against cv2.imencode.
//   k_jpeg_dct    : one CTA per 16 x 16 MCU -- BGR -> YCbCr (16-bit fixed point), h2v2 chroma box with the alternating bias, edge
//                   replication exactly as jcprepct.c / jcsample.c do it, 13-bit integer LLM forward DCT, round-half-up quantisation;
//                   dummy blocks (outside the component) carry the previous block's DC.  Output: int16 coefficients in zigzag order.
//   k_jpeg_code<0>: one warp per block -- length of its Huffman code (DC difference against the previous block of the component,
//                   run/size symbols with ZRL, EOB); exclusive scan -> bit offset of every block
//   k_jpeg_code<1>: the same walk, now writing the bits: a warp scan of the item lengths places every coefficient's code, atomicOr
//                   into a zeroed MSB-first word stream
//   k_jpeg_ffcount / k_jpeg_stuff : 0xFF -> 0xFF 0x00 byte stuffing as count + scan + scatter
// Every step is data parallel; the only host round trips are the two totals (bits, stuffed bytes) that size the next launches.
#include "jpeg.cuh"
#include <mutex>
#include <string.h>

namespace {

const uint8_t kZigzag[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                             35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
const uint8_t kLumaQ[64] = {16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29, 51, 87, 80, 62,
                            18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
const uint8_t kChromaQ[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                              99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
// Annex K.3 Huffman tables: bits[1..16], then the values in code order
const uint8_t kDcBits[2][16] = {{0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0}, {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0}};
const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t kAcBits[2][16] = {{0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d}, {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77}};
const uint8_t kAcVals[2][162] = {
    {0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1,
     0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26,
     0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56,
     0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85,
     0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa,
     0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6,
     0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
     0xfa},
    {0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42,
     0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19,
     0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55,
     0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83,
     0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8,
     0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4,
     0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9,
     0xfa}};

struct JpegTables {                       // canonical codes (jchuff.c jpeg_make_c_derived_tbl), [0] luma, [1] chroma
    unsigned short dc_code[2][12]; unsigned char dc_size[2][12];
    unsigned short ac_code[2][256]; unsigned char ac_size[2][256];
    unsigned char zigzag[64];
};

void derive(const uint8_t* bits, const uint8_t* vals, unsigned short* code, unsigned char* size) {
    unsigned c = 0; int k = 0;
    for (int len = 1; len <= 16; ++len) {
        for (int i = 0; i < bits[len - 1]; ++i, ++k, ++c) { code[vals[k]] = (unsigned short)c; size[vals[k]] = (unsigned char)len; }
        c <<= 1;
    }
}

void quality_table(const uint8_t* base, int quality, uint8_t* q /*natural order*/) {      // jcparam.c jpeg_quality_scaling + jpeg_add_quant_table
    quality = quality < 1 ? 1 : quality > 100 ? 100 : quality;
    const int scale = quality < 50 ? 5000 / quality : 200 - 2 * quality;
    for (int i = 0; i < 64; ++i) {
        int v = (base[i] * scale + 50) / 100;
        q[i] = (uint8_t)(v < 1 ? 1 : v > 255 ? 255 : v);
    }
}

}  // namespace

__constant__ JpegTables c_jpeg;
struct JpegQ { unsigned short d[2][64]; };          // quantisation divisors 8 * q (the ISLOW DCT leaves a factor 8), zigzag order

static cudaError_t jpeg_upload_tables() {           // once per device
    static std::mutex mu; static unsigned long long done = 0ull;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    if ((done >> (dev & 63)) & 1ull) return cudaSuccess;
    JpegTables t; memset(&t, 0, sizeof t);
    for (int c = 0; c < 2; ++c) {
        derive(kDcBits[c], kDcVals, t.dc_code[c], t.dc_size[c]);
        derive(kAcBits[c], kAcVals[c], t.ac_code[c], t.ac_size[c]);
    }
    memcpy(t.zigzag, kZigzag, 64);
    e = cudaMemcpyToSymbol(c_jpeg, &t, sizeof t);
    if (e == cudaSuccess) done |= 1ull << (dev & 63);
    return e;
}

// one pass of jfdctint.c over 8 values; FIRST = the row pass (results scaled up by 4), else the column pass
template <bool FIRST>
__device__ __forceinline__ void jpeg_dct8(int* d) {
    const int t0 = d[0] + d[7], t7 = d[0] - d[7], t1 = d[1] + d[6], t6 = d[1] - d[6];
    const int t2 = d[2] + d[5], t5 = d[2] - d[5], t3 = d[3] + d[4], t4 = d[3] - d[4];
    const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    constexpr int N = FIRST ? 11 : 15, RND = 1 << (N - 1);
    if (FIRST) { d[0] = (t10 + t11) << 2; d[4] = (t10 - t11) << 2; }
    else { d[0] = (t10 + t11 + 2) >> 2; d[4] = (t10 - t11 + 2) >> 2; }
    int z1 = (t12 + t13) * 4433;
    d[2] = (z1 + t13 * 6270 + RND) >> N;
    d[6] = (z1 - t12 * 15137 + RND) >> N;
    z1 = t4 + t7;
    int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
    const int z5 = (z3 + z4) * 9633;
    const int a4 = t4 * 2446, a5 = t5 * 16819, a6 = t6 * 25172, a7 = t7 * 12299;
    z1 *= -7373; z2 *= -20995; z3 = z3 * -16069 + z5; z4 = z4 * -3196 + z5;
    d[7] = (a4 + z1 + z3 + RND) >> N;
    d[5] = (a5 + z2 + z4 + RND) >> N;
    d[3] = (a6 + z2 + z3 + RND) >> N;
    d[1] = (a7 + z1 + z4 + RND) >> N;
}

__global__ void __launch_bounds__(64) k_jpeg_dct(const uint8_t* __restrict__ bgr, int w, int h, size_t stride, JpegQ q, int16_t* __restrict__ coef) {
    __shared__ int blk[6][64];
    const int t = threadIdx.x, mx = blockIdx.x, my = blockIdx.y;
    {
        const int qx = t & 7, qy = t >> 3, gx0 = 16 * mx + 2 * qx, gy0 = 16 * my + 2 * qy;
        const int c0 = min(gx0, w - 1), c1 = min(gx0 + 1, w - 1);
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const uint8_t* row = bgr + (size_t)min(gy0 + dy, h - 1) * stride;
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const uint8_t* p = row + 3 * (dx ? c1 : c0);
                const int Y = (19595 * p[2] + 38470 * p[1] + 7471 * p[0] + 32768) >> 16;
                const int yy = 2 * qy + dy, xx = 2 * qx + dx;
                blk[(yy >> 3) * 2 + (xx >> 3)][(yy & 7) * 8 + (xx & 7)] = Y - 128;
            }
        }
        // chroma: the input is replicated to an even height only; below that the DOWNSAMPLED row is replicated (jcprepct.c)
        const int dsr = min(gy0 >> 1, ((h + 1) >> 1) - 1), r0 = 2 * dsr, r1 = min(r0 + 1, h - 1);
        int cb = 0, cr = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint8_t* p = bgr + (size_t)(k & 2 ? r1 : r0) * stride + 3 * (k & 1 ? c1 : c0);
            const int B = p[0], G = p[1], R = p[2];
            cb += (-11059 * R - 21709 * G + 32768 * B + (128 << 16) + 32767) >> 16;
            cr += (32768 * R - 27439 * G - 5329 * B + (128 << 16) + 32767) >> 16;
        }
        const int bias = (qx & 1) ? 2 : 1;
        blk[4][qy * 8 + qx] = ((cb + bias) >> 2) - 128;
        blk[5][qy * 8 + qx] = ((cr + bias) >> 2) - 128;
    }
    __syncthreads();
    if (t < 48) {
        int d[8]; int* r = &blk[t >> 3][(t & 7) * 8];
#pragma unroll
        for (int k = 0; k < 8; ++k) d[k] = r[k];
        jpeg_dct8<true>(d);
#pragma unroll
        for (int k = 0; k < 8; ++k) r[k] = d[k];
    }
    __syncthreads();
    if (t < 48) {
        int d[8]; int* c = &blk[t >> 3][t & 7];
#pragma unroll
        for (int k = 0; k < 8; ++k) d[k] = c[8 * k];
        jpeg_dct8<false>(d);
#pragma unroll
        for (int k = 0; k < 8; ++k) c[8 * k] = d[k];
    }
    __syncthreads();
    const int ybw = (w + 7) >> 3, ybh = (h + 7) >> 3, nat = c_jpeg.zigzag[t];
    int16_t* out = coef + ((size_t)my * gridDim.x + mx) * 6 * 64 + t;
    int prev_dc = 0;                                         // meaningful in thread 0 only
#pragma unroll
    for (int b = 0; b < 6; ++b) {
        const bool real = b >= 4 || (2 * mx + (b & 1) < ybw && 2 * my + (b >> 1) < ybh);
        int v;
        if (real) {
            const int c = blk[b][nat], dv = q.d[b >> 2][t];
            const int a = (abs(c) + (dv >> 1)) / dv;
            v = c < 0 ? -a : a;
        } else {
            v = t == 0 ? prev_dc : 0;                        // jccoefct.c: dummy block = DC of the block before it, no AC
        }
        prev_dc = v;
        out[b * 64] = (int16_t)v;
    }
}

__device__ __forceinline__ void jpeg_append(unsigned long long& code, int& len, unsigned c, int n) { code = (code << n) | c; len += n; }

// code of coefficient k of a block (k = 0: the DC difference `v`; k = 63 with v == 0: the end-of-block symbol)
__device__ __forceinline__ void jpeg_item(int k, int v, unsigned long long nzmask, int cls, unsigned long long& code, int& len) {
    code = 0ull; len = 0;
    if (k == 0) {
        int t = v, t2 = v;
        if (t < 0) { t = -t; --t2; }
        const int nb = 32 - __clz(t);
        jpeg_append(code, len, c_jpeg.dc_code[cls][nb], c_jpeg.dc_size[cls][nb]);
        if (nb) jpeg_append(code, len, (unsigned)t2 & ((1u << nb) - 1u), nb);
        return;
    }
    if (v == 0) {
        if (k == 63) jpeg_append(code, len, c_jpeg.ac_code[cls][0], c_jpeg.ac_size[cls][0]);
        return;
    }
    const unsigned long long below = nzmask & ((1ull << k) - 1ull);
    const int prev = below ? 63 - __clzll((long long)below) : 0;
    int run = k - prev - 1;
    for (; run > 15; run -= 16) jpeg_append(code, len, c_jpeg.ac_code[cls][0xF0], c_jpeg.ac_size[cls][0xF0]);
    int t = v, t2 = v;
    if (t < 0) { t = -t; --t2; }
    const int nb = 32 - __clz(t), sym = (run << 4) | nb;
    jpeg_append(code, len, c_jpeg.ac_code[cls][sym], c_jpeg.ac_size[cls][sym]);
    jpeg_append(code, len, (unsigned)t2 & ((1u << nb) - 1u), nb);
}

// `len` bits of `code` at bit position `pos` of an MSB-first stream of zeroed 32-bit words
__device__ __forceinline__ void jpeg_put(unsigned* __restrict__ words, unsigned long long pos, unsigned long long code, int len) {
    if (len == 0) return;
    const unsigned sh = (unsigned)(pos & 31ull);
    unsigned* wp = words + (pos >> 5);
    const unsigned long long left = code << (64 - len);       // left aligned
    const unsigned long long hi = left >> sh;
    const unsigned w0 = (unsigned)(hi >> 32), w1 = (unsigned)hi, w2 = sh ? (unsigned)((left << (64 - sh)) >> 32) : 0u;
    if (w0) atomicOr(wp, w0);
    if (w1) atomicOr(wp + 1, w1);
    if (w2) atomicOr(wp + 2, w2);
}

template <int EMIT>
__global__ void __launch_bounds__(256) k_jpeg_code(const int16_t* __restrict__ coef, int nblocks, unsigned* __restrict__ bits, unsigned* __restrict__ words) {
    const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= nblocks) return;
    const int r = b % 6, cls = r >= 4;
    int v0 = coef[(size_t)b * 64 + lane];
    const int v1 = coef[(size_t)b * 64 + 32 + lane];
    if (lane == 0) {                                          // DC prediction: the previous block of the same component in scan order
        const int pb = r == 0 ? b - 3 : r < 4 ? b - 1 : b - 6;
        v0 -= pb >= 0 ? coef[(size_t)pb * 64] : 0;
    }
    const unsigned long long nz = (unsigned long long)(__ballot_sync(0xffffffffu, lane > 0 && v0 != 0)) |
                                  ((unsigned long long)__ballot_sync(0xffffffffu, v1 != 0) << 32);
    unsigned long long c0, c1; int l0, l1;
    jpeg_item(lane, v0, nz, cls, c0, l0);
    jpeg_item(lane + 32, v1, nz, cls, c1, l1);
    int s0 = l0, s1 = l1;                                     // inclusive warp scans of the two halves
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, s0, d), c = __shfl_up_sync(0xffffffffu, s1, d);
        if (lane >= d) { s0 += a; s1 += c; }
    }
    const int tot0 = __shfl_sync(0xffffffffu, s0, 31), tot1 = __shfl_sync(0xffffffffu, s1, 31);
    if (!EMIT) {
        if (lane == 0) bits[b] = (unsigned)(tot0 + tot1);
    } else {
        const unsigned long long base = bits[b];
        jpeg_put(words, base + (unsigned)(s0 - l0), c0, l0);
        jpeg_put(words, base + (unsigned)(tot0 + s1 - l1), c1, l1);
    }
}

// exclusive prefix sum of a[0..n) in place by ONE CTA; *total = the sum
__global__ void __launch_bounds__(1024) k_jpeg_scan(unsigned* __restrict__ a, int n, unsigned* __restrict__ total) {
    __shared__ unsigned wsum[32];
    const int t = threadIdx.x, per = (n + 1023) / 1024, lo = min(t * per, n), hi = min(lo + per, n);
    unsigned s = 0;
    for (int i = lo; i < hi; ++i) s += a[i];
    unsigned inc = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, inc, d); if ((t & 31) >= d) inc += o; }
    if ((t & 31) == 31) wsum[t >> 5] = inc;
    __syncthreads();
    if (t < 32) {
        unsigned v = wsum[t], w = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, w, d); if (t >= d) w += o; }
        wsum[t] = w - v;
        if (t == 31) *total = w;
    }
    __syncthreads();
    unsigned run = wsum[t >> 5] + inc - s;
    for (int i = lo; i < hi; ++i) { const unsigned v = a[i]; a[i] = run; run += v; }
}

// jchuff.c flush_bits: the last byte is filled with 1-bits
__global__ void k_jpeg_pad(unsigned* __restrict__ words, const unsigned* __restrict__ totals) {
    const unsigned T = totals[0], r = T & 7u;
    if (threadIdx.x == 0 && r) jpeg_put(words, T, (1ull << (8 - r)) - 1ull, 8 - r);
}

__device__ __forceinline__ unsigned jpeg_ff_bytes(unsigned w, int valid) {         // number of 0xFF bytes among the first `valid` (MSB first)
    unsigned n = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) n += (i < valid && ((w >> (24 - 8 * i)) & 0xFFu) == 0xFFu) ? 1u : 0u;
    return n;
}

// 1024 stream bytes (256 words) per CTA
template <int WRITE>
__global__ void __launch_bounds__(256) k_jpeg_stuff(const unsigned* __restrict__ words, unsigned nraw, unsigned* __restrict__ ffcount, uint8_t* __restrict__ out) {
    __shared__ unsigned wsum[8];
    const int t = threadIdx.x;
    const unsigned g = blockIdx.x * 1024u + 4u * t;
    const int valid = g >= nraw ? 0 : (int)min(4u, nraw - g);
    const unsigned w = valid ? words[g >> 2] : 0u;
    const unsigned n = jpeg_ff_bytes(w, valid);
    unsigned inc = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(0xffffffffu, inc, d); if ((t & 31) >= d) inc += o; }
    if ((t & 31) == 31) wsum[t >> 5] = inc;
    __syncthreads();
    unsigned before = 0, total = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { if (k < (t >> 5)) before += wsum[k]; total += wsum[k]; }
    if (!WRITE) {
        if (t == 0) ffcount[blockIdx.x] = total;
    } else {
        unsigned o = g + ffcount[blockIdx.x] + before + inc - n;
        for (int i = 0; i < valid; ++i) {
            const uint8_t b = (uint8_t)(w >> (24 - 8 * i));
            out[o++] = b;
            if (b == 0xFF) out[o++] = 0;
        }
    }
}

size_t bm_jpeg_scan_bound(int w, int h) {
    const size_t nblocks = (size_t)((w + 15) / 16) * ((h + 15) / 16) * 6;
    return nblocks * 2 * 224 + 64;             // <= 64 coefficients x 26 bits (16-bit code + 10 value bits) = 208 bytes, every byte stuffed
}

void bm_jpeg_free(BmJpeg* j) {
    cudaFree(j->coef); cudaFree(j->bits); cudaFree(j->words); cudaFree(j->ffcount); cudaFree(j->totals); cudaFree(j->out);
    *j = BmJpeg();
}

#define JPEG_OK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return e_; } while (0)

cudaError_t bm_jpeg_encode_scan(BmJpeg* j, const uint8_t* d_bgr, int w, int h, size_t stride, int quality, size_t* scan_bytes, cudaStream_t s) {
    JPEG_OK(jpeg_upload_tables());
    const int mw = (w + 15) / 16, mh = (h + 15) / 16;
    const size_t nblocks = (size_t)mw * mh * 6;
    if (nblocks > (size_t)2400000 || mh > 65535) return cudaErrorInvalidValue;   // bit offsets are 32-bit: 2.4 M blocks x 1664 bits < 2^32 (~100 Mpixel)
    if (j->cap_blocks < nblocks) {
        cudaFree(j->coef); cudaFree(j->bits); j->coef = nullptr; j->bits = nullptr; j->cap_blocks = 0;
        JPEG_OK(cudaMalloc(&j->coef, nblocks * 64 * sizeof(int16_t)));
        JPEG_OK(cudaMalloc(&j->bits, nblocks * sizeof(unsigned)));
        j->cap_blocks = nblocks;
    }
    if (!j->totals) JPEG_OK(cudaMalloc(&j->totals, 2 * sizeof(unsigned)));
    JpegQ q; uint8_t ql[64], qc[64];
    quality_table(kLumaQ, quality, ql); quality_table(kChromaQ, quality, qc);
    for (int i = 0; i < 64; ++i) { q.d[0][i] = (unsigned short)(8 * ql[kZigzag[i]]); q.d[1][i] = (unsigned short)(8 * qc[kZigzag[i]]); }
    BM_COUNT_LAUNCHES(3);
    k_jpeg_dct<<<dim3(mw, mh), 64, 0, s>>>(d_bgr, w, h, stride, q, j->coef);
    k_jpeg_code<0><<<(unsigned)((nblocks + 7) / 8), 256, 0, s>>>(j->coef, (int)nblocks, j->bits, nullptr);
    k_jpeg_scan<<<1, 1024, 0, s>>>(j->bits, (int)nblocks, j->totals);
    unsigned T = 0;
    JPEG_OK(cudaMemcpyAsync(&T, j->totals, sizeof T, cudaMemcpyDeviceToHost, s));
    JPEG_OK(cudaStreamSynchronize(s));
    const unsigned nraw = (unsigned)(((unsigned long long)T + 7ull) / 8ull);
    const size_t nwords = (size_t)nraw / 4 + 4, nchunks = ((size_t)nraw + 1023) / 1024;
    if (j->cap_words < nwords) {
        cudaFree(j->words); cudaFree(j->ffcount); j->words = nullptr; j->ffcount = nullptr; j->cap_words = 0;
        JPEG_OK(cudaMalloc(&j->words, nwords * sizeof(unsigned)));
        JPEG_OK(cudaMalloc(&j->ffcount, (nwords / 256 + 2) * sizeof(unsigned)));
        j->cap_words = nwords;
    }
    if (j->cap_out < 2 * (size_t)nraw + 16) {
        cudaFree(j->out); j->out = nullptr; j->cap_out = 0;
        JPEG_OK(cudaMalloc(&j->out, 2 * (size_t)nraw + 16));
        j->cap_out = 2 * (size_t)nraw + 16;
    }
    JPEG_OK(cudaMemsetAsync(j->words, 0, nwords * sizeof(unsigned), s));
    BM_COUNT_LAUNCHES(2);
    k_jpeg_code<1><<<(unsigned)((nblocks + 7) / 8), 256, 0, s>>>(j->coef, (int)nblocks, j->bits, j->words);
    k_jpeg_pad<<<1, 32, 0, s>>>(j->words, j->totals);
    unsigned nff = 0;
    if (nchunks) {
        BM_COUNT_LAUNCHES(3);
        k_jpeg_stuff<0><<<(unsigned)nchunks, 256, 0, s>>>(j->words, nraw, j->ffcount, nullptr);
        k_jpeg_scan<<<1, 1024, 0, s>>>(j->ffcount, (int)nchunks, j->totals + 1);
        k_jpeg_stuff<1><<<(unsigned)nchunks, 256, 0, s>>>(j->words, nraw, j->ffcount, j->out);
        JPEG_OK(cudaMemcpyAsync(&nff, j->totals + 1, sizeof nff, cudaMemcpyDeviceToHost, s));
    }
    JPEG_OK(cudaGetLastError());
    JPEG_OK(cudaStreamSynchronize(s));
    *scan_bytes = (size_t)nraw + nff;
    return cudaSuccess;
}

// jcmarker.c: write_file_header (SOI, JFIF APP0), write_frame_header (DQT per table, SOF0), write_scan_header (DHT per table, SOS)
size_t bm_jpeg_write_header(uint8_t* dst, int w, int h, int quality) {
    uint8_t* p = dst;
    auto put = [&](std::initializer_list<int> v) { for (int b : v) *p++ = (uint8_t)b; };
    put({0xFF, 0xD8});
    put({0xFF, 0xE0, 0, 16, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0});
    for (int tb = 0; tb < 2; ++tb) {
        uint8_t q[64];
        quality_table(tb ? kChromaQ : kLumaQ, quality, q);
        put({0xFF, 0xDB, 0, 67, tb});
        for (int i = 0; i < 64; ++i) *p++ = q[kZigzag[i]];
    }
    put({0xFF, 0xC0, 0, 17, 8, h >> 8, h & 255, w >> 8, w & 255, 3, 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1});
    for (int tb = 0; tb < 2; ++tb)
        for (int ac = 0; ac < 2; ++ac) {
            const uint8_t* bits = ac ? kAcBits[tb] : kDcBits[tb];
            const uint8_t* vals = ac ? kAcVals[tb] : kDcVals;
            int n = 0;
            for (int i = 0; i < 16; ++i) n += bits[i];
            put({0xFF, 0xC4, (n + 19) >> 8, (n + 19) & 255, (ac << 4) | tb});
            memcpy(p, bits, 16); p += 16;
            memcpy(p, vals, n); p += n;
        }
    put({0xFF, 0xDA, 0, 12, 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0});
    return (size_t)(p - dst);
}
