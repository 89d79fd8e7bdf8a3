// rowscan.cuh -- nearest-zero scan of one image row by one 256-thread CTA (shared by dt.cu and the warp kernel).
#pragma once
#include "dt.cuh"

// zero-pixel bits of a thread: one byte per 2048-pixel chunk, chunks 0..7 in lo, 8..15 in hi
struct BmZeroBits {
    unsigned long long lo, hi;
    __device__ __forceinline__ void clear() { lo = 0ull; hi = 0ull; }
    __device__ __forceinline__ void set(int c, unsigned b) {
        if (c < 8) lo |= (unsigned long long)b << (8 * c); else hi |= (unsigned long long)b << (8 * (c - 8));
    }
    __device__ __forceinline__ unsigned get(int c) const {
        return (unsigned)((c < 8 ? lo >> (8 * c) : hi >> (8 * (c - 8))) & 0xffull);
    }
};

// Thread t owns pixels [8t, 8t+8) of every 2048-pixel chunk of the row; bit i of zb.get(c) is set iff pixel i of its
// group in chunk c is a zero pixel (pixels beyond n must be reported as non-zero).  With d(x) = min(distance to the nearest zero
// pixel of the row, 0xFFFF) it writes the sweep seed g[x] = a * d(x) (cv2's 16.16 axial step; BM_DT_INF for d = 0xFFFF = "no zero
// in this row") for x in [0, n) and BM_DT_INF for the padding columns -- in whole groups of 8, the row buffer is padded to 8.
// Every thread of the CTA must call it (it synchronises).
static __device__ __forceinline__ void bm_rowscan_block(const BmZeroBits& zb, int nch, int n, uint32_t* __restrict__ grow) {
    __shared__ int s_wl[BM_ROWSCAN_MAX_CHUNKS][8], s_wf[BM_ROWSCAN_MAX_CHUNKS][8];
    __shared__ int s_cl[BM_ROWSCAN_MAX_CHUNKS], s_cf[BM_ROWSCAN_MAX_CHUNKS];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NEG = -(1 << 28), POS = 1 << 28;
    for (int c = 0; c < nch; ++c) {
        const int base = c * BM_ROWSCAN_CHUNK + 8 * tid;
        const unsigned b = zb.get(c);
        const int lz = b ? base + 31 - __clz(b) : NEG, fz = b ? base + __ffs(b) - 1 : POS;
        const int wl = __reduce_max_sync(0xffffffffu, lz), wf = __reduce_min_sync(0xffffffffu, fz);
        if (lane == 0) { s_wl[c][warp] = wl; s_wf[c][warp] = wf; }
    }
    __syncthreads();
    if (tid < nch) {           // exclusive prefix (last zero of earlier chunks) / suffix (first zero of later chunks)
        int l = NEG, f = POS;
        for (int c = 0; c < tid; ++c) for (int w = 0; w < 8; ++w) l = max(l, s_wl[c][w]);
        for (int c = tid + 1; c < nch; ++c) for (int w = 0; w < 8; ++w) f = min(f, s_wf[c][w]);
        s_cl[tid] = l; s_cf[tid] = f;
    }
    __syncthreads();
    for (int c = 0; c < nch; ++c) {
        const int base = c * BM_ROWSCAN_CHUNK + 8 * tid;
        const unsigned b = zb.get(c);
        int lz = b ? base + 31 - __clz(b) : NEG, fz = b ? base + __ffs(b) - 1 : POS;
        // inclusive warp scans: last zero in lanes <= lane, first zero in lanes >= lane
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, lz, d), v = __shfl_down_sync(0xffffffffu, fz, d);
            if (lane >= d) lz = max(lz, u);
            if (lane + d < 32) fz = min(fz, v);
        }
        int P = __shfl_up_sync(0xffffffffu, lz, 1), S = __shfl_down_sync(0xffffffffu, fz, 1);
        if (lane == 0) P = NEG;
        if (lane == 31) S = POS;
        P = max(P, s_cl[c]); S = min(S, s_cf[c]);
        for (int w = 0; w < warp; ++w) P = max(P, s_wl[c][w]);
        for (int w = warp + 1; w < 8; ++w) S = min(S, s_wf[c][w]);
        if (base < n) {
            unsigned o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const unsigned ml = b & ((2u << i) - 1u), mr = b >> i;
                const int p = base + i;
                const int left = ml ? base + 31 - __clz(ml) : P;
                const int right = mr ? p + __ffs(mr) - 1 : S;
                const unsigned gv = (unsigned)min(min(p - left, right - p), (int)BM_G_INF);
                o[i] = p < n ? BM_CHAMFER_SEED(gv) : BM_DT_INIT;          // padding columns never seed a sweep
            }
            *reinterpret_cast<uint4*>(grow + base) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4*>(grow + base + 4) = make_uint4(o[4], o[5], o[6], o[7]);
        }
    }
    __syncthreads();           // the shared tables may be reused by the next row of the same CTA
}
