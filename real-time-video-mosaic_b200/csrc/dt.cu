// dt.cu -- cv2.distanceTransform(mask, DIST_L2, 3) (reference: /root/reference/main.py:888-889), exact, O(1) per pixel.
//
// OpenCV's 3x3 chamfer transform is the shortest-path distance on the 8-connected pixel grid with integer 16.16 step
// costs a (axial) and b (diagonal), a <= b <= 2a.  A shortest path to a zero pixel can always be ordered as
//        k vertical steps, then j steps along ONE diagonal direction, then a horizontal run,
// and steps commute, so with  s(x,y) = a * g(x,y)  (g = horizontal distance to the nearest zero pixel of row y):
//        E1(x,y) = min(s(x,y), E1(x-1,y-1) + b)        E2(x,y) = min(s(x,y), E2(x+1,y-1) + b)
//        V (x,y) = min(min(E1,E2)(x,y), V(x,y-1) + a)                                   ("downward" sweep)
// plus the mirrored upward sweep; D = min(V_down, V_up).  Every recurrence is a one-directional linear-cost min-scan
// along a line (diagonal or column): cutting the lines into blocks of 16 rows, a block's result is
//        min(block-local scan, carry of the previous block + 16 * step)
// and the carries form a scalar chain PER LINE (no coupling between lines).  That gives five fully parallel phases:
//   1. k_dt_local       block-local E1/E2 (down and up) at the block boundary rows              -> LE
//   2. k_dt_diag_chain  one thread per diagonal line walks the blocks                           -> CE
//   3. k_dt_vert_local  E1/E2 with their carries, min, block-local V at the boundary rows       -> CV (local)
//   4. k_dt_vert_chain  one thread per column walks the blocks                                  -> CV (in place)
//   5. k_dt_weights / k_dt_map   both sweeps of a block with all carries -> D (-> blend weights)
// For the canvas plane phase 1 is persistent: only blocks whose rows changed are recomputed (after each blend).
// All arithmetic is integer and order-free (min / +), so the result equals OpenCV's two-pass raster scan bit for bit.
// A warp sweeps a tile of 128 columns x 16 rows, 4 consecutive columns per lane (8-byte loads of g), exchanging only
// the two edge columns per row with its neighbours by shuffle; 16 columns on each side are halo.
#include "dt.cuh"
#include "rowscan.cuh"

#define A_ ((unsigned)BM_CHAMFER_A)
#define B_ ((unsigned)BM_CHAMFER_B)
typedef unsigned int u32;
struct DtRange { int xa[2], xb[2]; };
#define DT_CHAIN_BATCH 32
#define FULL 0xffffffffu

// ------------------------------------------------------------------------------------------------------------------
// row scans
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rowscan_bgrx(const uchar4* __restrict__ img, int img_stride, int img_col0, int img_row0,
                                                      uint32_t* __restrict__ g, int gs, int n, int row0, const int* __restrict__ flags,
                                                      int need_flag) {
    BM_PDL_TRIGGER(); BM_PDL_WAIT();
    if (need_flag && flags[0] == 0) return;
    const int r = blockIdx.x, tid = threadIdx.x;
    const uchar4* row = img + (size_t)(img_row0 + r) * img_stride + img_col0;
    const int nch = (n + BM_ROWSCAN_CHUNK - 1) / BM_ROWSCAN_CHUNK;
    const bool vec = (reinterpret_cast<uintptr_t>(row) & 15) == 0;
    BmZeroBits zb; zb.clear();
    for (int c = 0; c < nch; ++c) {
        const int base = c * BM_ROWSCAN_CHUNK + 8 * tid;
        unsigned b = 0;
        if (vec && base + 7 < n) {
            const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(row + base)), v1 = __ldg(reinterpret_cast<const uint4*>(row + base) + 1);
            b = ((v0.x >> 24) == 0) | (((v0.y >> 24) == 0) << 1) | (((v0.z >> 24) == 0) << 2) | (((v0.w >> 24) == 0) << 3) |
                (((v1.x >> 24) == 0) << 4) | (((v1.y >> 24) == 0) << 5) | (((v1.z >> 24) == 0) << 6) | (((v1.w >> 24) == 0) << 7);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) if (base + i < n && row[base + i].w == 0) b |= 1u << i;
        }
        zb.set(c, b);
    }
    bm_rowscan_block(zb, nch, n, g + (size_t)(row0 + r) * gs);
}

__global__ void __launch_bounds__(256) k_rowscan_mask(const uint8_t* __restrict__ mask, int stride, uint32_t* __restrict__ g, int gs, int n) {
    const int r = blockIdx.x, tid = threadIdx.x;
    const uint8_t* row = mask + (size_t)r * stride;
    const int nch = (n + BM_ROWSCAN_CHUNK - 1) / BM_ROWSCAN_CHUNK;
    BmZeroBits zb; zb.clear();
    for (int c = 0; c < nch; ++c) {
        const int base = c * BM_ROWSCAN_CHUNK + 8 * tid;
        unsigned b = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) if (base + i < n && row[base + i] == 0) b |= 1u << i;
        zb.set(c, b);
    }
    bm_rowscan_block(zb, nch, n, g + (size_t)r * gs);
}

// ------------------------------------------------------------------------------------------------------------------
// warp-tile sweep primitives
// ------------------------------------------------------------------------------------------------------------------
struct DtRows { uint4 g[BM_BLK_ROWS]; };

// the 16 rows of block k for the lane's 4 columns [cx, cx+4) (cx is a multiple of 4, may lie outside the plane), in
// PROCESSING order: R.g[i] is row i of the block for a downward sweep, row 15-i for an upward one -- one code path
// serves both directions (and both planes), which keeps the unrolled sweeps small in the instruction cache
__device__ __forceinline__ void dt_load_rows(const BmDtPlane& p, int k, int cx, bool up, DtRows& R) {
    const bool ld = cx >= 0 && cx < p.W;                  // columns [W, gs) of a row hold BM_DT_INF (row scan), so no per-column mask
    const int r0 = k * BM_BLK_ROWS + (up ? BM_BLK_ROWS - 1 : 0), dr = up ? -1 : 1;
    const uint32_t* base = p.g + (size_t)r0 * p.gs + cx;
    const ptrdiff_t step = (ptrdiff_t)dr * p.gs;
#pragma unroll
    for (int i = 0; i < BM_BLK_ROWS; ++i) {
        R.g[i] = make_uint4(BM_DT_INF, BM_DT_INF, BM_DT_INF, BM_DT_INF);
        if (ld && r0 + dr * i < p.H) R.g[i] = __ldg(reinterpret_cast<const uint4*>(base + i * step));
    }
}

// seeds of row r: stored ready-made by the row scan (BmDtPlane::g)
__device__ __forceinline__ void dt_seeds(const DtRows& R, int r, u32 (&s)[4]) {
    const uint4 v = R.g[r];
    s[0] = v.x; s[1] = v.y; s[2] = v.z; s[3] = v.w;
}

// one row step of both diagonal scans: E1 flows to the right (from column x-1 of the previous row), E2 to the left
// (every value is <= DIST_MAX = UINT_MAX - b, so "+ b" cannot wrap; "+ a" neither since a < b)
__device__ __forceinline__ void dt_step(u32 (&E1)[4], u32 (&E2)[4], const u32 (&s)[4], int lane) {
    u32 l = __shfl_up_sync(FULL, E1[3], 1), r = __shfl_down_sync(FULL, E2[0], 1);
    if (lane == 0) l = BM_DT_INF;
    if (lane == 31) r = BM_DT_INF;
    E1[3] = min(s[3], E1[2] + B_); E1[2] = min(s[2], E1[1] + B_); E1[1] = min(s[1], E1[0] + B_); E1[0] = min(s[0], l + B_);
    E2[0] = min(s[0], E2[1] + B_); E2[1] = min(s[1], E2[2] + B_); E2[2] = min(s[2], E2[3] + B_); E2[3] = min(s[3], r + B_);
}

__device__ __forceinline__ void dt_fill(u32 (&v)[4], u32 x) { v[0] = v[1] = v[2] = v[3] = x; }

// 4 carry values of table row `k` at columns [cx, cx+4); BM_DT_INF outside the plane / block range
// `kind`: 0 = E1, 1 = E2, 2 = V -- selects the row of the plane's ghost carries used when k is the block just outside the plane
__device__ __forceinline__ void dt_carry4(const u32* __restrict__ tab, const BmDtPlane& p, int k, int cx, u32 (&o)[4], int kind) {
    dt_fill(o, BM_DT_INF);
    if (cx < 0 || cx >= p.W) return;
    const u32* row;
    if (k >= 0 && k < p.nb) row = tab + (size_t)k * p.ts;
    else {
        const u32* gh = k < 0 ? p.gh_top : p.gh_bot;
        if (gh == nullptr || (k != -1 && k != p.nb)) return;
        row = gh + (size_t)kind * p.ts;
    }
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(row + cx));
    o[0] = v.x;
    if (cx + 1 < p.W) o[1] = v.y;
    if (cx + 2 < p.W) o[2] = v.z;
    if (cx + 3 < p.W) o[3] = v.w;
}

__device__ __forceinline__ void dt_store4(u32* __restrict__ tab, const BmDtPlane& p, int k, int cx, const u32 (&v)[4]) {
    *reinterpret_cast<uint4*>(tab + (size_t)k * p.ts + cx) = make_uint4(v[0], v[1], v[2], v[3]);
}

// lanes whose 4 columns are exact after 16 row steps (the outer 16 columns of the 128 miss contributions)
__device__ __forceinline__ bool dt_lane_valid(int lane) { return lane >= 4 && lane < 28; }

// one sweep over the 16 rows of a block in processing order.  row(i, V) is called after row step i.
template <bool WITH_V, class RowFn>
__device__ __forceinline__ void dt_sweep(const DtRows& R, u32 (&E1)[4], u32 (&E2)[4], u32 (&V)[4], int lane, RowFn row) {
#pragma unroll
    for (int i = 0; i < BM_BLK_ROWS; ++i) {
        u32 s[4];
        dt_seeds(R, i, s);
        dt_step(E1, E2, s, lane);
        if (WITH_V) {
#pragma unroll
            for (int c = 0; c < 4; ++c) V[c] = min(min(E1[c], E2[c]), V[c] + A_);
        }
        row(i, V);
    }
}
struct DtNoRow { __device__ __forceinline__ void operator()(int, const u32 (&)[4]) const {} };

// ------------------------------------------------------------------------------------------------------------------
// phase 1: block-local diagonal sweeps.  CTA = 2 tiles x {down, up}: one warp per (tile, direction)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_dt_local(BmDtPlane p, int kb0, const int* __restrict__ flags, int need_flag) {
    BM_PDL_TRIGGER(); BM_PDL_WAIT();
    if (need_flag && flags[0] == 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * 2 + (warp >> 1), k = kb0 + blockIdx.y;
    const bool up = warp & 1;
    if (tile * BM_DT_TILE_VALID >= p.W) return;
    const int cx = tile * BM_DT_TILE_VALID - 16 + 4 * lane;
    DtRows R;
    dt_load_rows(p, k, cx, up, R);
    u32 E1[4], E2[4], V[4];
    dt_fill(E1, BM_DT_INF); dt_fill(E2, BM_DT_INF); dt_fill(V, BM_DT_INF);
    dt_sweep<false>(R, E1, E2, V, lane, DtNoRow());
    if (dt_lane_valid(lane) && cx < p.W) {
        dt_store4(p.LE + (up ? 2 : 0) * p.tsz, p, k, cx, E1);
        dt_store4(p.LE + (up ? 3 : 1) * p.tsz, p, k, cx, E2);
    }
}

__device__ __forceinline__ u32 dt_add_sat(u32 c, u32 d) { return c > BM_DT_INF - d ? BM_DT_INF : c + d; }

// carry entering a diagonal line from outside the plane: the line's first block (j = 0) sits at column x0, its predecessor in the
// neighbouring tile at column x0 - dx of the ghost row (E1 for the lines flowing right, E2 for those flowing left)
__device__ __forceinline__ u32 dt_ghost_diag(const BmDtPlane& p, int type, int x0, int dx) {
    const u32* gh = type < 2 ? p.gh_top : p.gh_bot;
    const int xg = x0 - dx;
    if (gh == nullptr || x0 < 0 || x0 >= p.W || xg < 0 || xg >= p.W) return BM_DT_INF;
    return __ldg(gh + (size_t)(type & 1) * p.ts + xg);
}
__device__ __forceinline__ u32 dt_ghost_vert(const BmDtPlane& p, bool down, int x) {
    const u32* gh = down ? p.gh_top : p.gh_bot;
    if (gh == nullptr || x < 0 || x >= p.W) return BM_DT_INF;
    return __ldg(gh + (size_t)2 * p.ts + x);
}

// ------------------------------------------------------------------------------------------------------------------
// phase 2: diagonal carries.  type 0: down, from x-16; 1: down, from x+16; 2: up, from x-16; 3: up, from x+16
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_dt_diag_chain(BmDtPair pp, const int* __restrict__ flags, int need_flag) {
    BM_PDL_TRIGGER(); BM_PDL_WAIT();
    if (need_flag && flags[0] == 0) return;
    const BmDtPlane& p = pp.p[blockIdx.z];
    const int type = blockIdx.y;
    const int nlines = p.W + 16 * (p.nb - 1);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nlines) return;
    const bool right = (type & 1) == 0, down = type < 2;
    const int x0 = right ? t - 16 * (p.nb - 1) : t, dx = right ? 16 : -16;
    const u32* __restrict__ L = p.LE + (size_t)type * p.tsz;
    u32* __restrict__ C = p.CE + (size_t)type * p.tsz;
    u32 c = dt_ghost_diag(p, type, x0, dx);
    // the loads do not depend on the chain: issue DT_CHAIN_BATCH of them at once, then walk
    for (int j0 = 0; j0 < p.nb; j0 += DT_CHAIN_BATCH) {
        u32 v[DT_CHAIN_BATCH];
#pragma unroll
        for (int u = 0; u < DT_CHAIN_BATCH; ++u) {
            const int j = j0 + u, x = x0 + dx * j, k = down ? j : p.nb - 1 - j;
            v[u] = (j < p.nb && x >= 0 && x < p.W) ? __ldg(L + (size_t)k * p.ts + x) : 0u;
        }
#pragma unroll
        for (int u = 0; u < DT_CHAIN_BATCH; ++u) {
            const int j = j0 + u, x = x0 + dx * j, k = down ? j : p.nb - 1 - j;
            if (j < p.nb && x >= 0 && x < p.W) { c = min(v[u], dt_add_sat(c, 16u * B_)); C[(size_t)k * p.ts + x] = c; }
            else c = BM_DT_INF;
        }
    }
}

// Parallel form of the same chains for nb <= 256 blocks: c_j = min_i<=j (L_i + (j - i) * K) is a prefix minimum, so a
// line is cut into segments of 8 blocks handled by one thread each: local chain, segment ends exchanged through shared
// memory, carry of all earlier segments applied.  CTA = 16 lines x nseg segments (lanes walk lines: 64-byte rows).
// A line's in-plane elements are contiguous in j, so a finite segment end always reaches the following elements.
// SEG = 8 blocks per thread serves planes of up to 256 blocks (4096 rows), SEG = 32 up to 1024 blocks (the extended row tiles of config 5)
template <int DT_SEG, class Addr>
__device__ __forceinline__ void dt_chain_seg(const u32* __restrict__ L, u32* __restrict__ C, int nb, u32 K, Addr addr, u32 gin) {
    __shared__ u32 E[32][17];
    const int ll = threadIdx.x & 15, sg = threadIdx.x >> 4;
    const int j0 = sg * DT_SEG;
    int idx[DT_SEG];
    u32 loc[DT_SEG];
#pragma unroll
    for (int u = 0; u < DT_SEG; ++u) {
        idx[u] = j0 + u < nb ? addr(j0 + u) : -1;
        loc[u] = idx[u] >= 0 ? __ldg(L + idx[u]) : BM_DT_INF;
    }
    u32 c = BM_DT_INF;
#pragma unroll
    for (int u = 0; u < DT_SEG; ++u) { c = idx[u] >= 0 ? min(loc[u], dt_add_sat(c, K)) : BM_DT_INF; loc[u] = c; }
    E[sg][ll] = c;
    __syncthreads();
    u32 cin = gin;                                         // carry at the last element of segment sg-1 (gin: entering from outside the plane)
    for (int s2 = 0; s2 < sg; ++s2) cin = min(E[s2][ll], dt_add_sat(cin, DT_SEG * K));      // Horner form of the prefix minimum
#pragma unroll
    for (int u = 0; u < DT_SEG; ++u)
        if (idx[u] >= 0) C[idx[u]] = min(loc[u], dt_add_sat(cin, (u32)(u + 1) * K));
}

template <int DT_SEG>
__global__ void __launch_bounds__(512) k_dt_diag_chain16(BmDtPair pp, const int* __restrict__ flags, int need_flag) {
    BM_PDL_TRIGGER(); BM_PDL_WAIT();
    if (need_flag && flags[0] == 0) return;
    const BmDtPlane& p = pp.p[blockIdx.z];
    const int type = blockIdx.y;
    const int nlines = p.W + 16 * (p.nb - 1);
    if (blockIdx.x * 16 >= nlines) return;
    const int t = blockIdx.x * 16 + (threadIdx.x & 15);
    const bool right = (type & 1) == 0, down = type < 2;
    const int x0 = right ? t - 16 * (p.nb - 1) : t, dx = right ? 16 : -16;
    const int nb = p.nb, W = p.W, ts = p.ts;
    dt_chain_seg<DT_SEG>(p.LE + (size_t)type * p.tsz, p.CE + (size_t)type * p.tsz, nb, 16u * B_, [=](int j) -> int {
        const int x = x0 + dx * j, k = down ? j : nb - 1 - j;
        return (t < nlines && x >= 0 && x < W) ? k * ts + x : -1;
    }, t < nlines ? dt_ghost_diag(p, type, x0, dx) : BM_DT_INF);
}

template <int DT_SEG>
__global__ void __launch_bounds__(512) k_dt_vert_chain16(BmDtPair pp, DtRange rg, const int* __restrict__ flags, int need_flag) {
    BM_PDL_TRIGGER(); BM_PDL_WAIT();
    if (need_flag && flags[0] == 0) return;
    const int pl = blockIdx.z;
    const BmDtPlane& p = pp.p[pl];
    if (rg.xa[pl] + blockIdx.x * 16 >= rg.xb[pl]) return;
    const int x = rg.xa[pl] + blockIdx.x * 16 + (threadIdx.x & 15);
    const bool ok = x < rg.xb[pl] && x < p.W, down = blockIdx.y == 0;
    const int nb = p.nb, ts = p.ts;
    u32* C = p.CV + (size_t)blockIdx.y * p.tsz;
    dt_chain_seg<DT_SEG>(C, C, nb, 16u * A_, [=](int j) -> int { return ok ? (down ? j : nb - 1 - j) * ts + x : -1; }, ok ? dt_ghost_vert(p, down, x) : BM_DT_INF);
}

// ------------------------------------------------------------------------------------------------------------------
// phase 3: diagonal sweeps with carries -> block-local vertical sweep
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_dt_vert_local(BmDtPair pp, DtRange rg, const int* __restrict__ flags, int need_flag) {
    BM_PDL_TRIGGER(); BM_PDL_WAIT();
    if (need_flag && flags[0] == 0) return;
    const int pl = blockIdx.z;
    const BmDtPlane& p = pp.p[pl];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * 2 + (warp >> 1), k = blockIdx.y;
    const bool up = warp & 1;
    const int xa = rg.xa[pl], xb = rg.xb[pl];
    if (k >= p.nb || xa + tile * BM_DT_TILE_VALID >= xb) return;
    const int cx = xa + tile * BM_DT_TILE_VALID - 16 + 4 * lane;
    DtRows R;
    dt_load_rows(p, k, cx, up, R);
    u32 E1[4], E2[4], V[4];
    dt_carry4(p.CE + (up ? 2 : 0) * p.tsz, p, up ? k + 1 : k - 1, cx, E1, 0);
    dt_carry4(p.CE + (up ? 3 : 1) * p.tsz, p, up ? k + 1 : k - 1, cx, E2, 1);
    dt_fill(V, BM_DT_INF);
    dt_sweep<true>(R, E1, E2, V, lane, DtNoRow());
    if (dt_lane_valid(lane) && cx < xb && cx < p.W) dt_store4(p.CV + (up ? 1 : 0) * p.tsz, p, k, cx, V);
}

// ------------------------------------------------------------------------------------------------------------------
// phase 4: vertical carries, in place.  type 0: down, 1: up
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_dt_vert_chain(BmDtPair pp, DtRange rg, const int* __restrict__ flags, int need_flag) {
    BM_PDL_TRIGGER(); BM_PDL_WAIT();
    if (need_flag && flags[0] == 0) return;
    const int pl = blockIdx.z;
    const BmDtPlane& p = pp.p[pl];
    const int x = rg.xa[pl] + blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= rg.xb[pl] || x >= p.W) return;
    const bool down = blockIdx.y == 0;
    u32* __restrict__ C = p.CV + (size_t)blockIdx.y * p.tsz;
    u32 c = dt_ghost_vert(p, down, x);
    for (int j0 = 0; j0 < p.nb; j0 += DT_CHAIN_BATCH) {
        u32 v[DT_CHAIN_BATCH];
#pragma unroll
        for (int u = 0; u < DT_CHAIN_BATCH; ++u) {
            const int j = j0 + u, k = down ? j : p.nb - 1 - j;
            v[u] = j < p.nb ? C[(size_t)k * p.ts + x] : BM_DT_INF;
        }
#pragma unroll
        for (int u = 0; u < DT_CHAIN_BATCH; ++u) {
            const int j = j0 + u, k = down ? j : p.nb - 1 - j;
            if (j < p.nb) { c = min(v[u], dt_add_sat(c, 16u * A_)); C[(size_t)k * p.ts + x] = c; }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// phase 5: sweeps of one block with all carries.  One warp per (plane, direction) writes its 16 x 128 values to shared
// memory; the CTA then combines them.
// ------------------------------------------------------------------------------------------------------------------
#define DT_TW 128
__device__ __forceinline__ void dt_sweep_to_smem(const BmDtPlane& p, int k, int cx, bool up, int lane, u32 (*T)[DT_TW]) {
    DtRows R;
    dt_load_rows(p, k, cx, up, R);
    u32 E1[4], E2[4], V[4];
    dt_carry4(p.CE + (up ? 2 : 0) * p.tsz, p, up ? k + 1 : k - 1, cx, E1, 0);
    dt_carry4(p.CE + (up ? 3 : 1) * p.tsz, p, up ? k + 1 : k - 1, cx, E2, 1);
    dt_carry4(p.CV + (up ? 1 : 0) * p.tsz, p, up ? k + 1 : k - 1, cx, V, 2);
    auto put = [&](int i, const u32 (&v)[4]) {
        const int r = up ? BM_BLK_ROWS - 1 - i : i;
        *reinterpret_cast<uint4*>(&T[r][4 * lane]) = make_uint4(v[0], v[1], v[2], v[3]);
    };
    dt_sweep<true>(R, E1, E2, V, lane, put);
}

__global__ void __launch_bounds__(64) k_dt_map(BmDtPlane p, float* __restrict__ out) {
    __shared__ __align__(16) u32 T[2][BM_BLK_ROWS][DT_TW];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x, k = blockIdx.y;
    const int tx0 = tile * BM_DT_TILE_VALID - 16;
    dt_sweep_to_smem(p, k, tx0 + 4 * lane, warp == 1, lane, T[warp]);
    __syncthreads();
    for (int i = threadIdx.x; i < BM_BLK_ROWS * BM_DT_TILE_VALID; i += 64) {
        const int r = i / BM_DT_TILE_VALID, c = 16 + i - r * BM_DT_TILE_VALID;
        const int x = tx0 + c, y = k * BM_BLK_ROWS + r;
        if (x < p.W && y < p.H) out[(size_t)y * p.W + x] = __fmul_rn(__uint2float_rn(min(T[0][r][c], T[1][r][c])), 1.0f / 65536.0f);
    }
}

// (dn/s, do/s) in float32 exactly as NumPy does it (main.py:892-894): dist = uint * 2^-16, s = (dn + do) + 1e-6f
__global__ void __launch_bounds__(128) k_dt_weights(BmDtPair pp, BmFramePlan plan, float2* __restrict__ wno,
                                                    const int* __restrict__ flags) {
    BM_PDL_TRIGGER(); BM_PDL_WAIT();
    if (flags[0] == 0) return;
    __shared__ __align__(16) u32 T[4][BM_BLK_ROWS][DT_TW];      // old down, old up, new down, new up
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x;
    const int k = plan.reg.y0 / BM_BLK_ROWS + blockIdx.y;
    const int tx0 = plan.rx0 + tile * BM_DT_TILE_VALID - 16;      // canvas column of the tile's first (halo) column
    const int kn = k - plan.win.y0 / BM_BLK_ROWS;
    const bool has_new = kn >= 0 && kn < pp.p[1].nb && tx0 + DT_TW > plan.win.x0 && tx0 < plan.win.x1;
    if (warp < 2 || has_new)                                     // one code path: (plane, direction) are data
        dt_sweep_to_smem(pp.p[warp >> 1], warp < 2 ? k : kn, tx0 - (warp < 2 ? 0 : plan.win.x0) + 4 * lane, warp & 1, lane, T[warp]);
    __syncthreads();
    const float scale = 1.0f / 65536.0f;
    for (int i = threadIdx.x; i < BM_BLK_ROWS * BM_DT_TILE_VALID; i += 128) {
        const int r = i / BM_DT_TILE_VALID, c = 16 + i - r * BM_DT_TILE_VALID;
        const int x = tx0 + c, y = k * BM_BLK_ROWS + r;
        if (x < plan.reg.x0 || x >= plan.reg.x1 || y < plan.reg.y0 || y >= plan.reg.y1) continue;
        const u32 d_old = min(T[0][r][c], T[1][r][c]);
        u32 d_new = 0u;
        if (has_new && x >= plan.win.x0 && x < plan.win.x1 && y >= plan.win.y0 && y < plan.win.y1) d_new = min(T[2][r][c], T[3][r][c]);
        const float dn = __fmul_rn(__uint2float_rn(d_new), scale);
        const float dold = __fmul_rn(__uint2float_rn(d_old), scale);
        const float s = __fadd_rn(__fadd_rn(dn, dold), 1e-6f);
        const size_t o = (size_t)(y - plan.reg.y0) * plan.rws + (x - plan.rx0);
        wno[o] = make_float2(__fdiv_rn(dn, s), __fdiv_rn(dold, s));
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
cudaError_t bm_dt_alloc_plane(BmDtPlane* p, int Wcap, int Hcap, size_t px_cap) {
    memset(p, 0, sizeof(*p));                             // (gh_top / gh_bot = nullptr: image border)
    // any live shape (W, H) with W <= Wcap, H <= Hcap, W*H <= px_cap must fit
    p->g_cap = px_cap + (size_t)8 * Hcap + 64;
    const size_t tsz = ((px_cap / BM_BLK_ROWS + (size_t)2 * (Wcap + 8) + (size_t)Hcap + 64) + 3) & ~(size_t)3;
    p->tsz = tsz;
    cudaError_t e;
    if ((e = cudaMalloc(&p->g, p->g_cap * sizeof(uint32_t))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&p->LE, 4 * tsz * sizeof(uint32_t))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&p->CE, 4 * tsz * sizeof(uint32_t))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&p->CV, 2 * tsz * sizeof(uint32_t))) != cudaSuccess) return e;
    bm_dt_shape_plane(p, Wcap, Hcap);
    return cudaSuccess;
}

void bm_dt_free_plane(BmDtPlane* p) {
    cudaFree(p->g); cudaFree(p->LE); cudaFree(p->CE); cudaFree(p->CV);
    memset(p, 0, sizeof(*p));
}

bool bm_dt_shape_plane(BmDtPlane* p, int W, int H) {
    p->W = W; p->H = H;
    p->gs = bm_pad8(W); p->ts = bm_pad4(W);
    p->nb = bm_div_up(H, BM_BLK_ROWS);
    return (size_t)p->gs * H <= p->g_cap && (size_t)p->nb * p->ts <= p->tsz;
}

cudaError_t bm_launch_rowscan_bgrx(const uchar4* img, int img_stride, int img_col0, int img_row0, const BmDtPlane& p, int row0, int nrows,
                                   const int* flags, int need_flag, cudaStream_t s) {
    if (nrows <= 0) return cudaSuccess;
    if (p.W > BM_ROWSCAN_CHUNK * BM_ROWSCAN_MAX_CHUNKS) return cudaErrorInvalidValue;
    BM_COUNT_LAUNCHES(1);
    return bm_launch_pdl(k_rowscan_bgrx, dim3(nrows), dim3(256), 0, s, img, img_stride, img_col0, img_row0, p.g, p.gs, p.W, row0, flags, need_flag);
}

cudaError_t bm_launch_rowscan_u8(const uint8_t* mask, int stride, const BmDtPlane& p, cudaStream_t s) {
    if (p.W > BM_ROWSCAN_CHUNK * BM_ROWSCAN_MAX_CHUNKS) return cudaErrorInvalidValue;
    BM_COUNT_LAUNCHES(1), k_rowscan_mask<<<p.H, 256, 0, s>>>(mask, stride, p.g, p.gs, p.W);
    return cudaGetLastError();
}

static inline int dt_tiles(int width) { return bm_div_up(width, BM_DT_TILE_VALID); }

cudaError_t bm_launch_dt_local(const BmDtPlane& p, int kb0, int kb1, const int* flags, int need_flag, cudaStream_t s) {
    if (kb1 <= kb0) return cudaSuccess;
    BM_COUNT_LAUNCHES(1);
    return bm_launch_pdl(k_dt_local, dim3(bm_div_up(dt_tiles(p.W), 2), kb1 - kb0), dim3(128), 0, s, p, kb0, flags, need_flag);
}

cudaError_t bm_launch_dt_carries(const BmDtPair& pp, int nplanes, const int xa[2], const int xb[2], const int* flags, int need_flag,
                                 cudaStream_t s) {
    DtRange rg;
    int max_lines = 0, max_tiles = 0, max_nb = 0, max_cols = 0;
    for (int i = 0; i < 2; ++i) {
        const int j = i < nplanes ? i : 0;
        rg.xa[i] = xa[j]; rg.xb[i] = xb[j];
        if (i >= nplanes) continue;
        const BmDtPlane& p = pp.p[i];
        max_lines = max_lines > p.W + 16 * (p.nb - 1) ? max_lines : p.W + 16 * (p.nb - 1);
        max_tiles = max_tiles > dt_tiles(xb[i] - xa[i]) ? max_tiles : dt_tiles(xb[i] - xa[i]);
        max_nb = max_nb > p.nb ? max_nb : p.nb;
        max_cols = max_cols > xb[i] - xa[i] ? max_cols : xb[i] - xa[i];
    }
    if (max_nb == 0 || max_cols <= 0) return cudaSuccess;
    // segment-parallel chains (table index fits 32 bits); longer lines fall back to one thread per line
    const bool par = max_nb <= 1024 && pp.p[0].tsz < ((size_t)1 << 31) && pp.p[1].tsz < ((size_t)1 << 31);
    const int seg = max_nb <= 256 ? 8 : 32;
    const int chain_threads = 32 * bm_div_up(bm_div_up(max_nb, seg), 2);        // 16 lines x ceil(nb / seg) segments
    BM_COUNT_LAUNCHES(3);
    cudaError_t e;
    const dim3 gd(bm_div_up(max_lines, 16), 4, nplanes), gv(bm_div_up(max_cols, 16), 2, nplanes);
    if (par && seg == 8) e = bm_launch_pdl(k_dt_diag_chain16<8>, gd, dim3(chain_threads), 0, s, pp, flags, need_flag);
    else if (par) e = bm_launch_pdl(k_dt_diag_chain16<32>, gd, dim3(chain_threads), 0, s, pp, flags, need_flag);
    else e = bm_launch_pdl(k_dt_diag_chain, dim3(bm_div_up(max_lines, 128), 4, nplanes), dim3(128), 0, s, pp, flags, need_flag);
    if (e != cudaSuccess) return e;
    if ((e = bm_launch_pdl(k_dt_vert_local, dim3(bm_div_up(max_tiles, 2), max_nb, nplanes), dim3(128), 0, s, pp, rg, flags, need_flag)) != cudaSuccess) return e;
    if (par && seg == 8) e = bm_launch_pdl(k_dt_vert_chain16<8>, gv, dim3(chain_threads), 0, s, pp, rg, flags, need_flag);
    else if (par) e = bm_launch_pdl(k_dt_vert_chain16<32>, gv, dim3(chain_threads), 0, s, pp, rg, flags, need_flag);
    else e = bm_launch_pdl(k_dt_vert_chain, dim3(bm_div_up(max_cols, 128), 2, nplanes), dim3(128), 0, s, pp, rg, flags, need_flag);
    return e;
}

cudaError_t bm_launch_dt_weights(const BmDtPair& pp, const BmFramePlan& plan, float2* wno, const int* flags, cudaStream_t s) {
    const int nyb = (plan.reg.y1 - 1) / BM_BLK_ROWS - plan.reg.y0 / BM_BLK_ROWS + 1;
    BM_COUNT_LAUNCHES(1);
    return bm_launch_pdl(k_dt_weights, dim3(dt_tiles(plan.reg.x1 - plan.rx0), nyb), dim3(128), 0, s, pp, plan, wno, flags);
}

cudaError_t bm_launch_dt_export_carries(const BmDtPlane& p, int up, int block, uint32_t* d_out, cudaStream_t s) {
    if (block < 0 || block >= p.nb) return cudaErrorInvalidValue;
    BmDtPair pp; pp.p[0] = p; pp.p[1] = p;
    const int xa[2] = {0, 0}, xb[2] = {p.W, p.W};
    cudaError_t e = bm_launch_dt_carries(pp, 1, xa, xb, nullptr, 0, s);
    if (e != cudaSuccess) return e;
    const uint32_t* src[3] = {p.CE + (size_t)(up ? 2 : 0) * p.tsz, p.CE + (size_t)(up ? 3 : 1) * p.tsz, p.CV + (size_t)(up ? 1 : 0) * p.tsz};
    for (int i = 0; i < 3 && e == cudaSuccess; ++i)
        e = cudaMemcpyAsync(d_out + (size_t)i * p.W, src[i] + (size_t)block * p.ts, (size_t)p.W * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s);
    return e;
}

cudaError_t bm_launch_dt_map(const BmDtPlane& p, float* d_out, cudaStream_t s) {
    BmDtPair pp; pp.p[0] = p; pp.p[1] = p;
    cudaError_t e = bm_launch_dt_local(p, 0, p.nb, nullptr, 0, s);
    if (e != cudaSuccess) return e;
    const int xa[2] = {0, 0}, xb[2] = {p.W, p.W};
    e = bm_launch_dt_carries(pp, 1, xa, xb, nullptr, 0, s);
    if (e != cudaSuccess) return e;
    BM_COUNT_LAUNCHES(1), k_dt_map<<<dim3(dt_tiles(p.W), p.nb), 64, 0, s>>>(p, d_out);
    return cudaGetLastError();
}
