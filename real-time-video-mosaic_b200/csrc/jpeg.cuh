// jpeg.cuh -- baseline JPEG encoder on the device (jpeg.cu): the file cv2.imwrite('mosaic.jpg', img) writes, byte for byte
#pragma once
#include "common.cuh"

struct BmJpeg {
    int16_t* coef = nullptr;         // [nblocks][64] quantised coefficients, zigzag order, scan order (Y00 Y01 Y10 Y11 Cb Cr per MCU)
    unsigned* bits = nullptr;        // [nblocks] code length of each block, then its exclusive prefix (bit offset)
    unsigned* words = nullptr;       // entropy-coded bit stream before byte stuffing, MSB first in 32-bit words
    unsigned* ffcount = nullptr;     // [nchunks] 0xFF bytes per 1024-byte chunk, then the exclusive prefix
    unsigned* totals = nullptr;      // [0] = bits in the scan, [1] = 0xFF bytes in it
    uint8_t* out = nullptr;          // stuffed scan bytes
    size_t cap_blocks = 0, cap_words = 0, cap_out = 0;
};

#define BM_JPEG_HEADER_MAX 1024
// worst case of the entropy-coded segment for w x h pixels (every coefficient at its longest code, every byte stuffed)
size_t bm_jpeg_scan_bound(int w, int h);
void bm_jpeg_free(BmJpeg* j);
// d_bgr: packed BGR, `stride` bytes per row.  Leaves the stuffed scan in j->out and its length in *scan_bytes (host, after a stream sync).
cudaError_t bm_jpeg_encode_scan(BmJpeg* j, const uint8_t* d_bgr, int w, int h, size_t stride, int quality, size_t* scan_bytes, cudaStream_t s);
// SOI ... SOS of the file (jcmarker.c order); returns the number of bytes written (<= BM_JPEG_HEADER_MAX)
size_t bm_jpeg_write_header(uint8_t* dst, int w, int h, int quality);
