"""oracle/jpeg.py -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the baseline JPEG file `cv2.imwrite('mosaic.jpg', img)` writes
(/root/reference/main.py:1664-1665, default parameters: quality 95, 4:2:0, Huffman tables of Annex K, no restart markers).

The encoder is third-party code absent from /root/reference: OpenCV 4.13.0's bundled libjpeg-turbo driven by `grfmt_jpeg.cpp`
(jpeg_set_defaults, in_color_space BGR, jpeg_set_quality(95, TRUE), JDCT_ISLOW, optimize_coding off).  Restated from the published
algorithm of the IJG library; pinned byte for byte against live `cv2.imencode('.jpg', img)` (tests/test_oracle_jpeg_cpu.py):

  * colour (jccolor.c rgb_ycc_convert): 16-bit fixed point, Y = (19595 R + 38470 G + 7471 B + 32768) >> 16,
    Cb = (-11059 R - 21709 G + 32768 B + (128 << 16) + 32767) >> 16, Cr = (32768 R - 27439 G - 5329 B + (128 << 16) + 32767) >> 16;
  * edges (jcprepct.c / jcsample.c): the last column is replicated up to the padded width and the last row to an even height, chroma =
    h2v2 box (a + b + c + d + bias) >> 2 with bias 1, 2, 1, 2, ... along a row, then the last (downsampled) row is replicated down to
    the MCU row;
  * blocks wholly outside the component (the MCU grid is 16 x 16) are dummies: AC = 0, DC = the previous block's quantised DC
    (jccoefct.c compress_data);
  * forward DCT (jfdctint.c, the 13-bit Loeffler-Ligtenberg-Moschytz integer transform, output scaled by 8), samples - 128;
  * quantisation (jcdctmgr.c): round-half-up division of |c| by 8 q, sign restored; q = clamp((base * 10 + 50) / 100, 1, 255);
  * entropy coding (jchuff.c encode_one_block), 0xFF byte stuffing, final pad with 1-bits;
  * markers (jcmarker.c): SOI, APP0 JFIF 1.01 (density 1:1, no unit), DQT 0, DQT 1, SOF0, DHT DC0, AC0, DC1, AC1, SOS, data, EOI.
"""
import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21,
                   28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61,
                   54, 47, 55, 62, 63], np.int32)

STD_LUMA_Q = np.array([16, 11, 10, 16, 24, 40, 51, 61, 12, 12, 14, 19, 26, 58, 60, 55, 14, 13, 16, 24, 40, 57, 69, 56, 14, 17, 22, 29,
                       51, 87, 80, 62, 18, 22, 37, 56, 68, 109, 103, 77, 24, 35, 55, 64, 81, 104, 113, 92, 49, 64, 78, 87, 103, 121,
                       120, 101, 72, 92, 95, 98, 112, 100, 103, 99], np.int32)
STD_CHROMA_Q = np.array([17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99, 24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99,
                         99, 99, 99, 99, 99] + [99] * 32, np.int32)

DC_LUMA_BITS = [0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0]
DC_CHROMA_BITS = [0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0]
DC_VALS = list(range(12))
AC_LUMA_BITS = [0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d]
AC_LUMA_VALS = [0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32,
                0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16,
                0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45,
                0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69,
                0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94,
                0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6,
                0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8,
                0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
                0xf9, 0xfa]
AC_CHROMA_BITS = [0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77]
AC_CHROMA_VALS = [0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81,
                  0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34,
                  0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44,
                  0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68,
                  0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92,
                  0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4,
                  0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6,
                  0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
                  0xf9, 0xfa]


def quant_table(base, quality):
    """jcparam.c jpeg_quality_scaling + jpeg_add_quant_table(force_baseline): natural (row-major) order."""
    quality = min(max(int(quality), 1), 100)
    scale = 5000 // quality if quality < 50 else 200 - 2 * quality
    return np.clip((base * scale + 50) // 100, 1, 255).astype(np.int32)


def huff_codes(bits, vals):
    """jchuff.c jpeg_make_c_derived_tbl: canonical codes -> (code[256], size[256])."""
    code = np.zeros(256, np.int64)
    size = np.zeros(256, np.int32)
    c, k = 0, 0
    for ln in range(1, 17):
        for _ in range(bits[ln - 1]):
            code[vals[k]], size[vals[k]] = c, ln
            c += 1
            k += 1
        c <<= 1
    return code, size


def ycc_planes(bgr):
    b = bgr[..., 0].astype(np.int64)
    g = bgr[..., 1].astype(np.int64)
    r = bgr[..., 2].astype(np.int64)
    y = (19595 * r + 38470 * g + 7471 * b + 32768) >> 16
    cb = (-11059 * r - 21709 * g + 32768 * b + (128 << 16) + 32767) >> 16
    cr = (32768 * r - 27439 * g - 5329 * b + (128 << 16) + 32767) >> 16
    return y, cb, cr


def _pad_edge(p, hh, ww):
    return np.pad(p, ((0, hh - p.shape[0]), (0, ww - p.shape[1])), mode="edge")


def component_planes(bgr):
    """the three component planes the DCT reads, padded to whole MCUs by edge replication (only the blocks inside the component
    are used; the rest become dummy blocks)."""
    h, w = bgr.shape[:2]
    mw, mh = (w + 15) // 16, (h + 15) // 16
    y, cb, cr = ycc_planes(bgr)
    yp = _pad_edge(y, mh * 16, mw * 16)
    out = [yp]
    for c in (cb, cr):
        # columns: the INPUT row is replicated out to the padded width; rows: the input is replicated to an even height only, and
        # the DOWNSAMPLED rows are replicated down to the MCU row (jcprepct.c pre_process_data)
        cp = _pad_edge(c, h + (h & 1), mw * 16)
        bias = np.tile(np.array([1, 2], np.int64), mw * 4)[None, :]
        ds = (cp[0::2, 0::2] + cp[0::2, 1::2] + cp[1::2, 0::2] + cp[1::2, 1::2] + bias) >> 2
        out.append(_pad_edge(ds, mh * 8, mw * 8))
    return out, mw, mh


def _dct_1d(d, first):
    """one pass of jfdctint.c over the LAST axis of d[..., 8] (int64)."""
    t0, t7 = d[..., 0] + d[..., 7], d[..., 0] - d[..., 7]
    t1, t6 = d[..., 1] + d[..., 6], d[..., 1] - d[..., 6]
    t2, t5 = d[..., 2] + d[..., 5], d[..., 2] - d[..., 5]
    t3, t4 = d[..., 3] + d[..., 4], d[..., 3] - d[..., 4]
    t10, t13, t11, t12 = t0 + t3, t0 - t3, t1 + t2, t1 - t2
    o = np.empty_like(d)
    if first:
        o[..., 0] = (t10 + t11) << 2
        o[..., 4] = (t10 - t11) << 2
        n = 11
    else:
        o[..., 0] = (t10 + t11 + 2) >> 2
        o[..., 4] = (t10 - t11 + 2) >> 2
        n = 15
    rnd = 1 << (n - 1)
    z1 = (t12 + t13) * 4433
    o[..., 2] = (z1 + t13 * 6270 + rnd) >> n
    o[..., 6] = (z1 - t12 * 15137 + rnd) >> n
    z1, z2, z3, z4 = t4 + t7, t5 + t6, t4 + t6, t5 + t7
    z5 = (z3 + z4) * 9633
    t4, t5, t6, t7 = t4 * 2446, t5 * 16819, t6 * 25172, t7 * 12299
    z1, z2, z3, z4 = z1 * -7373, z2 * -20995, z3 * -16069 + z5, z4 * -3196 + z5
    o[..., 7] = (t4 + z1 + z3 + rnd) >> n
    o[..., 5] = (t5 + z2 + z4 + rnd) >> n
    o[..., 3] = (t6 + z2 + z3 + rnd) >> n
    o[..., 1] = (t7 + z1 + z4 + rnd) >> n
    return o


def fdct_quant(plane, q):
    """plane [8 * bh, 8 * bw] -> quantised coefficients [bh, bw, 64] in natural order."""
    bh, bw = plane.shape[0] // 8, plane.shape[1] // 8
    blk = plane.reshape(bh, 8, bw, 8).transpose(0, 2, 1, 3).astype(np.int64) - 128
    blk = _dct_1d(blk, True)                                   # rows
    blk = _dct_1d(blk.swapaxes(-1, -2), False).swapaxes(-1, -2)  # columns
    d = (q.reshape(8, 8) * 8).astype(np.int64)
    a = (np.abs(blk) + (d >> 1)) // d
    return (np.sign(blk) * a).reshape(bh, bw, 64).astype(np.int32)


def quantised_blocks(bgr, quality=95):
    """every block of the scan in coding order: (coef[n, 64] natural order, comp[n])."""
    h, w = bgr.shape[:2]
    planes, mw, mh = component_planes(bgr)
    ql, qc = quant_table(STD_LUMA_Q, quality), quant_table(STD_CHROMA_Q, quality)
    cy = fdct_quant(planes[0], ql)
    cc = [fdct_quant(planes[1], qc), fdct_quant(planes[2], qc)]
    ybw, ybh = (w + 7) // 8, (h + 7) // 8                      # blocks that belong to the component; the others are dummies
    cbw, cbh = ((w + 1) // 2 + 7) // 8, ((h + 1) // 2 + 7) // 8
    out = np.zeros((mh * mw * 6, 64), np.int32)
    comp = np.zeros(mh * mw * 6, np.int32)
    n = 0
    for my in range(mh):
        for mx in range(mw):
            for by in range(2):
                for bx in range(2):
                    yy, xx = 2 * my + by, 2 * mx + bx
                    if yy < ybh and xx < ybw:
                        out[n] = cy[yy, xx]
                    else:                                       # dummy: DC of the block before it in the MCU buffer
                        out[n, 0] = out[n - 1, 0] if (bx or by) else 0
                        if not (bx or by):
                            raise AssertionError("the first block of an MCU is always inside the image")
                    comp[n] = 0
                    n += 1
            for ci in range(2):
                if my < cbh and mx < cbw:
                    out[n] = cc[ci][my, mx]
                else:
                    raise AssertionError("chroma has one block per MCU: never a dummy")
                comp[n] = 1 + ci
                n += 1
    return out, comp


class _Bits:
    def __init__(self):
        self.acc, self.n, self.out = 0, 0, bytearray()

    def put(self, code, size):
        self.acc = (self.acc << size) | (int(code) & ((1 << size) - 1))
        self.n += size
        while self.n >= 8:
            b = (self.acc >> (self.n - 8)) & 0xFF
            self.out.append(b)
            if b == 0xFF:
                self.out.append(0)
            self.n -= 8
        self.acc &= (1 << self.n) - 1

    def flush(self):
        if self.n:
            self.put(0x7F, 8 - self.n)       # jchuff.c flush_bits: fill the last byte with 1-bits


def entropy_code(coefs, comp):
    dcl, acl = huff_codes(DC_LUMA_BITS, DC_VALS), huff_codes(AC_LUMA_BITS, AC_LUMA_VALS)
    dcc, acc = huff_codes(DC_CHROMA_BITS, DC_VALS), huff_codes(AC_CHROMA_BITS, AC_CHROMA_VALS)
    bw = _Bits()
    last = [0, 0, 0]
    for blk, c in zip(coefs, comp):
        dc, ac = (dcl, acl) if c == 0 else (dcc, acc)
        z = blk[ZIGZAG]
        t = int(z[0]) - last[c]
        last[c] = int(z[0])
        t2 = t
        if t < 0:
            t, t2 = -t, t - 1
        nb = t.bit_length()
        bw.put(dc[0][nb], dc[1][nb])
        if nb:
            bw.put(t2, nb)
        r = 0
        for k in range(1, 64):
            t = int(z[k])
            if t == 0:
                r += 1
                continue
            while r > 15:
                bw.put(ac[0][0xF0], ac[1][0xF0])
                r -= 16
            t2 = t
            if t < 0:
                t, t2 = -t, t - 1
            nb = t.bit_length()
            bw.put(ac[0][(r << 4) + nb], ac[1][(r << 4) + nb])
            bw.put(t2, nb)
            r = 0
        if r > 0:
            bw.put(ac[0][0], ac[1][0])
    bw.flush()
    return bytes(bw.out)


def _marker(m, payload):
    return bytes([0xFF, m]) + (len(payload) + 2).to_bytes(2, "big") + payload


def header(w, h, quality=95):
    ql, qc = quant_table(STD_LUMA_Q, quality), quant_table(STD_CHROMA_Q, quality)
    out = b"\xff\xd8" + _marker(0xE0, b"JFIF\0\x01\x01\x00\x00\x01\x00\x01\x00\x00")
    out += _marker(0xDB, bytes([0]) + bytes(int(v) for v in ql[ZIGZAG]))
    out += _marker(0xDB, bytes([1]) + bytes(int(v) for v in qc[ZIGZAG]))
    out += _marker(0xC0, bytes([8]) + h.to_bytes(2, "big") + w.to_bytes(2, "big") + bytes([3, 1, 0x22, 0, 2, 0x11, 1, 3, 0x11, 1]))
    for idx, bits, vals in ((0x00, DC_LUMA_BITS, DC_VALS), (0x10, AC_LUMA_BITS, AC_LUMA_VALS),
                            (0x01, DC_CHROMA_BITS, DC_VALS), (0x11, AC_CHROMA_BITS, AC_CHROMA_VALS)):
        out += _marker(0xC4, bytes([idx]) + bytes(bits) + bytes(vals))
    out += _marker(0xDA, bytes([3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0]))
    return out


def encode(bgr, quality=95):
    """the bytes of cv2.imencode('.jpg', bgr) / cv2.imwrite(path, bgr) for a 3-channel 8-bit image."""
    assert bgr.dtype == np.uint8 and bgr.ndim == 3 and bgr.shape[2] == 3
    h, w = bgr.shape[:2]
    coefs, comp = quantised_blocks(bgr, quality)
    return header(w, h, quality) + entropy_code(coefs, comp) + b"\xff\xd9"
