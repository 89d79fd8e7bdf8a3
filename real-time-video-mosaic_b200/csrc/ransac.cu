// ransac.cu -- batched RANSAC homography + LS refit + Levenberg-Marquardt, one persistent CTA per point set.
// Reference: cv2.findHomography(p_cur, p_prev, cv2.RANSAC, 2.0) at main.py:856-857; algorithm = OpenCV 4.x
// RANSACPointSetRegistrator::run / HomographyEstimatorCallback / LMSolver (restated in oracle/ransac.py).
//
// Sequential semantics kept on a parallel evaluator: the cv::RNG stream, the subset rejection rules, "first strictly
// better hypothesis wins" and the adaptive iteration cap are evaluated in iteration order by one thread, while the
// expensive parts -- 4-point solves and inlier counting of a batch of hypotheses (one per warp), the 45-term normal
// equation sums of the refit and the 55 sums of each (nine-parameter, cv2 4.13) LM step -- run across the CTA.
//
// Code-size note (measured, see profiles/): the kernel is executed by 16 warps once per frame, so every instruction is
// fetched cold.  A first version with fully unrolled register-resident 8x8 / 9x9 algebra compiled to 34k SASS
// instructions (550 KB) and ran at ~80 cycles per instruction.  All small dense algebra therefore goes through ONE
// non-inlined, non-unrolled routine (warp_solve) on a per-warp matrix in shared memory, and helpers are __noinline__.
#include "ransac.cuh"
#include "eig9.h"
#include <float.h>
#include <math.h>

#define RS_BATCH 32
#define RS_THREADS 512
#define RS_WARPS (RS_THREADS / 32)
#define RS_SMEM_PTS 2048      // point sets up to this size are staged in shared memory

struct RsShared {
    int idx[RS_BATCH][4];
    int valid[RS_BATCH];
    int good[RS_BATCH];
    double Hb[RS_BATCH][9];
    double bestH[9];
    double red[RS_WARPS][46];    // cross-warp reduction scratch
    double sums[46];
    double M[RS_WARPS][9][10];   // per-warp augmented matrices for warp_solve
    double X[RS_WARPS][9];       // per-warp solutions
    double P[RS_WARPS][9];       // per-warp reciprocal pivots
    int niters, iter, best_good, stop, batch, fail_at;
};

__device__ __forceinline__ unsigned rng_next(unsigned long long& st) {
    st = (unsigned long long)(unsigned)st * 4164903690ULL + (st >> 32);
    return (unsigned)st;
}

__device__ __noinline__ bool collinear4(const float2* p) {       // haveCollinearPoints(m, 4): point 3 against pairs of 0..2
    const int i = 3;
    for (int j = 0; j < i; ++j) {
        const double dx1 = (double)p[j].x - (double)p[i].x, dy1 = (double)p[j].y - (double)p[i].y;
        for (int k = 0; k < j; ++k) {
            const double dx2 = (double)p[k].x - (double)p[i].x, dy2 = (double)p[k].y - (double)p[i].y;
            if (fabs(dx2 * dy1 - dy2 * dx1) <= (double)FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2))) return true;
        }
    }
    return false;
}
__device__ __forceinline__ double det3pts(const float2& a, const float2& b, const float2& c) {
    const double a0 = a.x, a1 = a.y, b0 = b.x, b1 = b.y, c0 = c.x, c1 = c.y;
    return a0 * (b1 - c1) - a1 * (b0 - c0) + (b0 * c1 - b1 * c0);
}
__device__ __noinline__ bool check_subset(const float2* s, const float2* d) {
    if (collinear4(s) || collinear4(d)) return false;
    int neg = 0;
    for (int i = 0; i < 4; ++i) {
        const int t0 = (i == 1) ? 1 : 0, t1 = (i < 2) ? i + 1 : (i == 2 ? 2 : 1), t2 = (i == 0) ? 2 : 3;   // {0,1,2},{1,2,3},{0,2,3},{0,1,3}
        neg += (det3pts(s[t0], s[t1], s[t2]) * det3pts(d[t0], d[t1], d[t2]) < 0.0) ? 1 : 0;
    }
    return neg == 0 || neg == 4;
}

// Gaussian elimination with partial pivoting of the NxN system stored in M[r][0..N) | M[r][N], by one warp.
// Solution in X[0..N).  Returns false if a pivot magnitude is <= tiny.  Not unrolled on purpose (see header).
__device__ __noinline__ bool warp_solve(double (*M)[10], double* X, double* P, int N, double tiny) {
    const int lane = threadIdx.x & 31;
    bool ok = true;
    for (int c = 0; c < N; ++c) {
        double best = (lane >= c && lane < N) ? fabs(M[lane][c]) : -1.0;
        int bi = lane;
        for (int o = 8; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        best = __shfl_sync(0xffffffffu, best, 0); bi = __shfl_sync(0xffffffffu, bi, 0);
        if (!(best > tiny)) ok = false;
        if (bi != c && lane <= N) { const double t = M[c][lane]; M[c][lane] = M[bi][lane]; M[bi][lane] = t; }
        __syncwarp();
        const double inv = __drcp_rn(M[c][c]);
        if (lane == 0) P[c] = inv;
        const int k = c + 1 + lane;                     // lane <-> column; loop over the rows below the pivot
        for (int r = c + 1; r < N; ++r) {
            const double f = M[r][c] * inv;
            if (k <= N) M[r][k] -= f * M[c][k];
        }
        __syncwarp();
    }
    for (int r = N - 1; r >= 0; --r) {
        double part = (lane > r && lane < N) ? M[r][lane] * X[lane] : 0.0;
        for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) X[r] = (M[r][N] - part) * P[r];
        __syncwarp();
    }
    return ok;
}

// 4-point homography by one warp: normalise like HomographyEstimatorCallback::runKernel, solve the 8x8 system (h33 = 1 in
// normalised coordinates), denormalise, scale so that H[8] = 1.  Result on every lane.
__device__ __noinline__ bool solve4_warp(RsShared& sh, const float2* M, const float2* m, double* H) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double cMx = 0, cMy = 0, cmx = 0, cmy = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { cMx += M[i].x; cMy += M[i].y; cmx += m[i].x; cmy += m[i].y; }
    cMx /= 4; cMy /= 4; cmx /= 4; cmy /= 4;
    double sMx = 0, sMy = 0, smx = 0, smy = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { sMx += fabs(M[i].x - cMx); sMy += fabs(M[i].y - cMy); smx += fabs(m[i].x - cmx); smy += fabs(m[i].y - cmy); }
    if (fabs(sMx) < DBL_EPSILON || fabs(sMy) < DBL_EPSILON || fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON) return false;
    sMx = 4 / sMx; sMy = 4 / sMy; smx = 4 / smx; smy = 4 / smy;
    double (*A)[10] = sh.M[warp];
    if (lane < 8) {
        const int i = lane >> 1;
        const double X = (M[i].x - cMx) * sMx, Y = (M[i].y - cMy) * sMy, x = (m[i].x - cmx) * smx, y = (m[i].y - cmy) * smy;
        double* r = A[lane];
        if ((lane & 1) == 0) { r[0] = X; r[1] = Y; r[2] = 1; r[3] = 0; r[4] = 0; r[5] = 0; r[6] = -x * X; r[7] = -x * Y; r[8] = x; }
        else { r[0] = 0; r[1] = 0; r[2] = 0; r[3] = X; r[4] = Y; r[5] = 1; r[6] = -y * X; r[7] = -y * Y; r[8] = y; }
    }
    __syncwarp();
    const bool ok = warp_solve(A, sh.X[warp], sh.P[warp], 8, 1e-12);
    double h[9];
#pragma unroll
    for (int k = 0; k < 8; ++k) h[k] = sh.X[warp][k];
    h[8] = 1.0;
    __syncwarp();
    const double inv[9] = {1.0 / smx, 0, cmx, 0, 1.0 / smy, cmy, 0, 0, 1};
    const double nrm[9] = {sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1};
    double t[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) t[3 * r + c] = inv[3 * r] * h[c] + inv[3 * r + 1] * h[3 + c] + inv[3 * r + 2] * h[6 + c];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) H[3 * r + c] = t[3 * r] * nrm[c] + t[3 * r + 1] * nrm[3 + c] + t[3 * r + 2] * nrm[6 + c];
    const double sc = 1.0 / H[8];
    if (!ok || !isfinite(sc)) return false;
#pragma unroll
    for (int i = 0; i < 9; ++i) H[i] *= sc;
    return true;
}

// HomographyEstimatorCallback::computeError: float32, the model cast to float, err <= thresh^2
__device__ __forceinline__ bool is_inlier(const float* Hf, float2 M, float2 m, float t2) {
    const float ww = __fdiv_rn(1.f, __fadd_rn(__fadd_rn(__fmul_rn(Hf[6], M.x), __fmul_rn(Hf[7], M.y)), 1.f));
    const float dx = __fsub_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(Hf[0], M.x), __fmul_rn(Hf[1], M.y)), Hf[2]), ww), m.x);
    const float dy = __fsub_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(Hf[3], M.x), __fmul_rn(Hf[4], M.y)), Hf[5]), ww), m.y);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) <= t2;
}

__device__ __noinline__ int update_num_iters(double p, double ep, int model_points, int max_iters) {
    p = fmin(fmax(p, 0.0), 1.0); ep = fmin(fmax(ep, 0.0), 1.0);
    double num = fmax(1.0 - p, DBL_MIN);
    double denom = 1.0 - pow(1.0 - ep, (double)model_points);
    if (denom < DBL_MIN) return 0;
    num = log(num); denom = log(denom);
    return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : __double2int_rn(num / denom);
}

// CTA-wide sum of NV doubles per thread -> sh.sums[0..NV).  Template + full unroll so that the caller's accumulators stay
// in registers (passing them by pointer put them in local memory: hundreds of cycles per access on this one-CTA kernel).
template <int NV>
__device__ __forceinline__ void block_sum(RsShared& sh, const double (&v)[NV]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sh.red[warp][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0;
        for (int w = 0; w < RS_WARPS; ++w) s += sh.red[w][threadIdx.x];
        sh.sums[threadIdx.x] = s;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(RS_THREADS, 1) k_ransac_homography(const float2* __restrict__ gsrc, const float2* __restrict__ gdst,
                                                                     const int* __restrict__ countp, double thresh, int max_iters, double confidence,
                                                                     uint8_t* __restrict__ mask, BmRansacResult* __restrict__ out) {
    BM_PDL_WAIT();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RsShared& sh = *reinterpret_cast<RsShared*>(smem_raw);
    float2* spts = reinterpret_cast<float2*>(smem_raw + ((sizeof(RsShared) + 15) & ~(size_t)15));
    const int n = *countp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float t2 = (float)(thresh * thresh);
    __shared__ unsigned long long rng;
    __shared__ long long cyc[8];
    if (tid < 8) cyc[tid] = 0;
    const long long t_start = clock64();
    if (tid == 0) {
        out->ok = 0; out->n_points = n; out->iters = 0; out->n_inliers = 0; out->lm_iters = 0; out->jacobi_sweeps = 0;
        for (int i = 0; i < 9; ++i) out->H[i] = 0.0;
        for (int i = 0; i < 8; ++i) out->cyc[i] = 0;
    }
    if (n < 4) return;
    const float2* src = gsrc; const float2* dst = gdst;
    if (n <= RS_SMEM_PTS) {
        for (int i = tid; i < n; i += blockDim.x) { spts[i] = gsrc[i]; spts[RS_SMEM_PTS + i] = gdst[i]; }
        src = spts; dst = spts + RS_SMEM_PTS;
    }
    __syncthreads();
    if (n == 4) {       // findHomography: npoints == 4 -> runKernel only, no refinement
        if (warp == 0) {
            double H[9];
            float2 s4[4], d4[4];
            for (int i = 0; i < 4; ++i) { s4[i] = src[i]; d4[i] = dst[i]; }
            const bool ok = solve4_warp(sh, s4, d4, H);
            if (lane == 0 && ok) { for (int i = 0; i < 9; ++i) out->H[i] = H[i]; out->ok = 1; out->n_inliers = 4; }
        }
        return;
    }
    if (tid == 0) { rng = 0xFFFFFFFFFFFFFFFFULL; sh.niters = max_iters > 1 ? max_iters : 1; sh.iter = 0; sh.best_good = 0; sh.stop = 0; }
    __syncthreads();
    int batch_cap = 8;            // most frames stop after a handful of iterations: small first batch, then full batches
    while (true) {
        long long t0 = clock64();
        // ---- one thread: next batch of accepted subsets, in cv::RNG order (getSubset, 10000 attempts each) ----
        if (tid == 0) {
            int B = sh.niters - sh.iter; if (B > batch_cap) B = batch_cap;
            sh.batch = B; sh.fail_at = -1;
            unsigned long long st = rng;
            for (int b = 0; b < B; ++b) {
                bool found = false;
                for (int att = 0; att < 10000 && !found; ++att) {
                    int id[4];
                    for (int i = 0; i < 4; ++i) {
                        int v;
                        bool dup;
                        do {
                            v = (int)(rng_next(st) % (unsigned)n);
                            dup = false;
                            for (int k = 0; k < i; ++k) dup |= (id[k] == v);
                        } while (dup);
                        id[i] = v;
                    }
                    float2 s4[4], d4[4];
                    for (int i = 0; i < 4; ++i) { s4[i] = src[id[i]]; d4[i] = dst[id[i]]; }
                    if (check_subset(s4, d4)) { found = true; for (int i = 0; i < 4; ++i) sh.idx[b][i] = id[i]; }
                }
                if (!found) { sh.fail_at = b; break; }
            }
            rng = st;
        }
        __syncthreads();
        if (tid == 0) { const long long t1 = clock64(); cyc[0] += t1 - t0; t0 = t1; }
        const int B = sh.batch, fail_at = sh.fail_at;
        const int nb = fail_at >= 0 ? fail_at : B;
        // ---- warp hb: hypothesis hb: 4-point solve, inlier count by all lanes ----
        for (int hb = warp; hb < nb; hb += RS_WARPS) {
            float2 s4[4], d4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { s4[i] = src[sh.idx[hb][i]]; d4[i] = dst[sh.idx[hb][i]]; }
            double H[9];
            const bool ok = solve4_warp(sh, s4, d4, H);
            int cnt = 0;
            if (ok) {
                float Hf[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) Hf[i] = (float)H[i];
                for (int i = lane; i < n; i += 32) cnt += is_inlier(Hf, src[i], dst[i], t2) ? 1 : 0;
                for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            }
            if (lane == 0) {
                sh.valid[hb] = ok ? 1 : 0; sh.good[hb] = cnt;
                if (ok) {
#pragma unroll
                    for (int i = 0; i < 9; ++i) sh.Hb[hb][i] = H[i];
                }
            }
        }
        __syncthreads();
        // ---- one thread: sequential winner selection / adaptive iteration cap ----
        if (tid == 0) {
            { const long long t1 = clock64(); cyc[1] += t1 - t0; t0 = t1; }
            for (int b = 0; b < nb; ++b) {
                if (sh.iter >= sh.niters) break;
                if (sh.valid[b]) {
                    const int good = sh.good[b];
                    const int lim = sh.best_good > 3 ? sh.best_good : 3;
                    if (good > lim) {
                        sh.best_good = good;
                        for (int i = 0; i < 9; ++i) sh.bestH[i] = sh.Hb[b][i];
                        sh.niters = update_num_iters(confidence, (double)(n - good) / n, 4, sh.niters);
                    }
                }
                sh.iter++;
            }
            if (sh.iter >= sh.niters || fail_at >= 0) sh.stop = 1;
            cyc[2] += clock64() - t0;
        }
        __syncthreads();
        if (sh.stop) break;
        batch_cap = RS_BATCH;
    }
    if (sh.best_good <= 0) { if (tid == 0) { out->iters = sh.iter; } return; }
    long long tp = clock64();
    // ---- inlier mask of the winning hypothesis ----
    {
        float Hf[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) Hf[i] = (float)sh.bestH[i];
        for (int i = tid; i < n; i += blockDim.x) mask[i] = is_inlier(Hf, src[i], dst[i], t2) ? 1 : 0;
    }
    __syncthreads();
    if (tid == 0) {
        // stage-1 result: the winning 4-point model; k_ransac_refine (one warp) does the LS refit + LM on the inliers
        for (int i = 0; i < 9; ++i) out->H[i] = sh.bestH[i];
        out->ok = 2; out->iters = sh.iter; out->n_inliers = sh.best_good;
        cyc[3] = clock64() - tp; cyc[6] = clock64() - t_start;
        for (int i = 0; i < 8; ++i) out->cyc[i] = cyc[i];
    }
}

// ------------------------------------------------------------------------------------------------------------------
// stage 2: LS refit (normalised DLT) + LMSolver on the inliers, ONE WARP.
// Measured on B200: inside the 16-warp CTA of stage 1 the shuffle-heavy parts of warp_solve ran 12x slower than the same
// code in a one-warp kernel (34.7k vs 2.9k cycles for the pivot searches of a 9x9 solve), so the serial algebra lives in
// its own 32-thread launch; the 45-term normal-equation sums are strided over the lanes and reduced with shuffles.
// ------------------------------------------------------------------------------------------------------------------
#define RF_WARPS 4            // warp 0 runs the serial algebra; all four warps share the sums over the points (8 warps measured no faster)
struct RfShared {
    double A[9][9], V[9];
    double lmA[9][9], lmv[9], lmd[9], lmx[9], lmxd[9], lmD[9], lmW[9];
    double X[9], P[9];              // LM: unit gauge vector h / |h|, diagonal of the pseudo-inverse
    double G[9 * 18], Ainv[9][9];   // [A | I] elimination workspace, explicit inverse (shifted LtL of the DLT; damped / deflated JtJ of the LM)
    double sums[56];                // 0..44 upper triangle of a 9x9, 45..53 J^T r, 54 |r|^2, 55 max |r|
    double wsum[RF_WARPS][56];      // per-warp partial sums of rf_accumulate
    const double* cmd_h;            // command block for the helper warps: parameter vector, Jacobian wanted, 0 = exit
    int cmd, cmd_want_j;
};

// Gauss-Jordan elimination WITHOUT pivoting of a symmetric positive definite N x N matrix by one warp: every lane owns a few entries
// of the N x NC augmented matrix (row-major, NC > N) and all rows are eliminated at once, so a pivot step costs one round of
// shared-memory reads instead of a serial loop over the rows (the stage-2 systems -- LtL + shift*I of the inverse iteration, the damped /
// deflated JtJ of Levenberg-Marquardt -- are SPD, for which elimination without pivoting is as stable as Cholesky).  With [A | I] the
// right half becomes inv(A) once every row is divided by its pivot.  Columns evolve independently, so column N + c equals the
// right-hand side of a separate solve of A x = e_c bit for bit.  Returns false on a non-positive / non-finite pivot.
template <int N, int NC>
__device__ __noinline__ bool warp_gj_spd(double* M) {
    constexpr int EPL = (N * NC + 31) / 32;
    const int lane = threadIdx.x & 31;
    int er[EPL], ek[EPL];
#pragma unroll
    for (int j = 0; j < EPL; ++j) { const int e = lane + 32 * j; er[j] = e < N * NC ? e / NC : -1; ek[j] = e < N * NC ? e - er[j] * NC : 0; }
    bool ok = true;
    for (int c = 0; c < N; ++c) {
        const double piv = M[c * NC + c];
        if (!(piv > 0.0) || !isfinite(piv)) ok = false;
        const double inv = __drcp_rn(piv);
        double nv[EPL];
#pragma unroll
        for (int j = 0; j < EPL; ++j) {
            nv[j] = 0.0;
            if (er[j] >= 0 && er[j] != c && ek[j] > c) nv[j] = M[er[j] * NC + ek[j]] - (M[er[j] * NC + c] * inv) * M[c * NC + ek[j]];
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < EPL; ++j) if (er[j] >= 0 && er[j] != c && ek[j] > c) M[er[j] * NC + ek[j]] = nv[j];
        __syncwarp();
    }
    return ok;
}

template <int NV>
__device__ __forceinline__ void warp_sum(RfShared& sh, const double (&v)[NV]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sh.sums[k] = x;
    }
    __syncwarp();
}

template <int HALF>
__device__ __forceinline__ void rf_halve(double (&v)[64], int hi, int xor_mask) {
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
        const double a = v[i], b = v[i + HALF];
        const double send = hi ? a : b, keep = hi ? b : a;
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, xor_mask);
    }
}

// HomographyRefineCallback sums over this warp's share of the points (i = tid, tid + 128, ...): JtJ (45), JtR (9), |r|^2, max |r|.
// cv2 4.13 refines ALL NINE elements of H (fundam.cpp: `J.cols == 9`; the classic callback fixed h33 = 1): with ww = 1 / (h6 Mx + h7 My + h8)
// the Jacobian rows are j0 = (Mx*ww, My*ww, ww, 0, 0, 0, -Mx*ww*xi, -My*ww*xi, -ww*xi), j1 = (0, 0, 0, Mx*ww, My*ww, ww, -Mx*ww*yi,
// -My*ww*yi, -ww*yi): products with a structural zero are skipped; the sums use fused multiply-adds (cv2's own gemm order is not
// reproducible either -- the refined H is compared within 0.5 px, tests/test_features_gpu.py).
__device__ __noinline__ void rf_partial(RfShared& sh, const float2* src, const float2* dst, const uint8_t* mask, int n, const double* h, bool want_j) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr bool nz0[9] = {true, true, true, false, false, false, true, true, true}, nz1[9] = {false, false, false, true, true, true, true, true, true};
    double acc[55];
#pragma unroll
    for (int k = 0; k < 55; ++k) acc[k] = 0.0;
    double rmax = 0.0;
    for (int i = threadIdx.x; i < n; i += 32 * RF_WARPS) {
        if (!mask[i]) continue;
        const double Mx = src[i].x, My = src[i].y;
        const double ww = __drcp_rn(h[6] * Mx + h[7] * My + h[8]);
        const double xi = (h[0] * Mx + h[1] * My + h[2]) * ww, yi = (h[3] * Mx + h[4] * My + h[5]) * ww;
        const double rx = xi - (double)dst[i].x, ry = yi - (double)dst[i].y;
        acc[54] += rx * rx + ry * ry;
        rmax = fmax(rmax, fmax(fabs(rx), fabs(ry)));
        if (want_j) {
            const double j0[9] = {Mx * ww, My * ww, ww, 0, 0, 0, -Mx * ww * xi, -My * ww * xi, -ww * xi};
            const double j1[9] = {0, 0, 0, Mx * ww, My * ww, ww, -Mx * ww * yi, -My * ww * yi, -ww * yi};
            int k = 0;
#pragma unroll
            for (int a = 0; a < 9; ++a) {
#pragma unroll
                for (int b = a; b < 9; ++b) {
                    if (nz0[a] && nz0[b] && nz1[a] && nz1[b]) acc[k] = __fma_rn(j1[a], j1[b], __fma_rn(j0[a], j0[b], acc[k]));
                    else if (nz0[a] && nz0[b]) acc[k] = __fma_rn(j0[a], j0[b], acc[k]);
                    else if (nz1[a] && nz1[b]) acc[k] = __fma_rn(j1[a], j1[b], acc[k]);
                    ++k;
                }
            }
#pragma unroll
            for (int a = 0; a < 9; ++a) {
                if (nz0[a] && nz1[a]) acc[45 + a] = __fma_rn(j1[a], ry, __fma_rn(j0[a], rx, acc[45 + a]));
                else if (nz0[a]) acc[45 + a] = __fma_rn(j0[a], rx, acc[45 + a]);
                else acc[45 + a] = __fma_rn(j1[a], ry, acc[45 + a]);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
    if (want_j) {
        // 55 sums over 32 lanes by recursive halving: at every step a lane hands one half of its values to its partner and adds
        // the partner's copy of the other half -- 62 exchanges instead of 55 x 5 (the shuffle unit is the bottleneck of this kernel)
        double v[64];
#pragma unroll
        for (int k = 0; k < 64; ++k) v[k] = k < 55 ? acc[k] : 0.0;
        rf_halve<32>(v, lane & 16, 16); rf_halve<16>(v, lane & 8, 8); rf_halve<8>(v, lane & 4, 4); rf_halve<4>(v, lane & 2, 2);
        rf_halve<2>(v, lane & 1, 1);
        if (2 * lane < 55) sh.wsum[warp][2 * lane] = v[0];            // lane L ends up with the totals of values 2L and 2L + 1
        if (2 * lane + 1 < 55) sh.wsum[warp][2 * lane + 1] = v[1];
    } else {
        double x = acc[54];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) sh.wsum[warp][54] = x;
    }
    if (lane == 0) sh.wsum[warp][55] = rmax;
}

// called by warp 0: wake the helper warps, take a share, combine the per-warp partials (fixed order: deterministic)
__device__ __noinline__ void rf_accumulate(RfShared& sh, const float2* src, const float2* dst, const uint8_t* mask, int n, const double* h, bool want_j) {
    const int lane = threadIdx.x & 31;
    if (lane == 0) { sh.cmd_h = h; sh.cmd_want_j = want_j ? 1 : 0; sh.cmd = 1; }
    __syncthreads();
    rf_partial(sh, src, dst, mask, n, h, want_j);
    __syncthreads();
    for (int k = lane; k < 56; k += 32) {
        if (!want_j && k < 54) continue;
        double t = sh.wsum[0][k];
#pragma unroll
        for (int w = 1; w < RF_WARPS; ++w) t = (k == 55) ? fmax(t, sh.wsum[w][k]) : t + sh.wsum[w][k];
        sh.sums[k] = t;
    }
    __syncwarp();
}

// warp 0: the sums of the last rf_accumulate become the LM's normal matrix and gradient
__device__ __forceinline__ void lm_adopt(RfShared& sh) {
    const int lane = threadIdx.x & 31;
    for (int t = lane; t < 45; t += 32) {          // unpack the upper triangle of JtJ
        int a = 0, k = t;
        while (k >= 9 - a) { k -= 9 - a; ++a; }
        const int b = a + k;
        sh.lmA[a][b] = sh.sums[t]; sh.lmA[b][a] = sh.sums[t];
    }
    if (lane < 9) sh.lmv[lane] = sh.sums[45 + lane];
    __syncwarp();
}

// bm_jacobi9 (eig9.h) with the two 9-element loops of every rotation spread over nine lanes: lane k rotates row / column k of the matrix and
// row k of the eigenvector matrix.  An element (k, p) or (k, q) is read and written by lane k only, so the values are those of the serial
// loop bit for bit (--fmad=false: no contraction on either side); the rotation parameters are derived by every lane from the same three
// shared-memory words.  a, v: shared memory, row-major 9x9.  ~3x faster than the serial form on one lane (-DBM_JACOBI_SERIAL).
__device__ __noinline__ int jacobi9_warp(double* a, double* w, double* v) {
    const int lane = threadIdx.x & 31;
    constexpr int n = 9;
    for (int i = lane; i < n * n; i += 32) v[i] = (i / n == i % n) ? 1.0 : 0.0;
    __syncwarp();
    int sweep = 0;
    for (; sweep < 40; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < n - 1; ++p) {
            for (int q = p + 1; q < n; ++q) {
                const double apq = a[p * n + q];
                if (apq == 0.0) continue;
                const double app = a[p * n + p], aqq = a[q * n + q];
                __syncwarp();                       // every lane has read the three words before anybody writes them
                if (fabs(apq) <= 1e-300 + 2.220446049250313e-19 * sqrt(fabs(app) * fabs(aqq))) {
                    if (lane == 0) { a[p * n + q] = 0.0; a[q * n + p] = 0.0; }
                    __syncwarp();
                    continue;
                }
                rotated = true;
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                double np_ = 0, nq_ = 0, vp_ = 0, vq_ = 0;
                if (lane < n) {
                    const double akp = a[lane * n + p], akq = a[lane * n + q];
                    np_ = c * akp - sn * akq; nq_ = sn * akp + c * akq;
                    const double vp = v[lane * n + p], vq = v[lane * n + q];
                    vp_ = c * vp - sn * vq; vq_ = sn * vp + c * vq;
                }
                __syncwarp();
                if (lane < n) {
                    if (lane != p && lane != q) { a[lane * n + p] = np_; a[p * n + lane] = np_; a[lane * n + q] = nq_; a[q * n + lane] = nq_; }
                    v[lane * n + p] = vp_; v[lane * n + q] = vq_;
                }
                if (lane == 0) { a[p * n + p] = app - t * apq; a[q * n + q] = aqq + t * apq; a[p * n + q] = 0.0; a[q * n + p] = 0.0; }
                __syncwarp();
            }
        }
        if (!rotated) break;
    }
    if (lane < n) w[lane] = a[lane * n + lane];
    __syncwarp();
    return sweep;
}

__device__ unsigned long long g_lm_stats[3];   // polishes run, LM iterations, iterations solved by eigen-decomposition (bm_debug_lm_stats)
__device__ int g_lm_force_eig = 0;              // debug / tests (bm_debug_lm_force_eig): always take the eigen-decomposition route below

__global__ void __launch_bounds__(32 * RF_WARPS, 1) k_ransac_refine(const float2* __restrict__ gsrc, const float2* __restrict__ gdst, const int* __restrict__ countp,
                                                         const uint8_t* __restrict__ mask, BmRansacResult* __restrict__ out) {
    BM_PDL_WAIT();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RfShared& sh = *reinterpret_cast<RfShared*>(smem_raw);
    float2* spts = reinterpret_cast<float2*>(smem_raw + ((sizeof(RfShared) + 15) & ~(size_t)15));
    const int lane = threadIdx.x & 31;
    if (out->ok != 2) return;                   // nothing to refine (no model, or the 4-point special case)
    const long long t_start = clock64();
    const int n = *countp, cnt = out->n_inliers;
    const float2* src = gsrc; const float2* dst = gdst;
    if (n <= RS_SMEM_PTS) {
        for (int i = threadIdx.x; i < n; i += 32 * RF_WARPS) { spts[i] = gsrc[i]; spts[RS_SMEM_PTS + i] = gdst[i]; }
        src = spts; dst = spts + RS_SMEM_PTS;
    }
    __syncthreads();
    if (threadIdx.x >= 32) {                    // helper warps: serve rf_accumulate requests of warp 0 until told to exit
        while (true) {
            __syncthreads();
            if (sh.cmd == 0) return;
            rf_partial(sh, src, dst, mask, n, sh.cmd_h, sh.cmd_want_j != 0);
            __syncthreads();
        }
    }
    if (lane < 9) sh.V[lane] = out->H[lane];     // bestH of stage 1
    __syncwarp();
    // ---- LS refit: normalised DLT (runKernel), 9x9 LtL, smallest eigenvector ----
    {
        double v[4] = {0, 0, 0, 0};
        for (int i = lane; i < n; i += 32) if (mask[i]) { v[0] += dst[i].x; v[1] += dst[i].y; v[2] += src[i].x; v[3] += src[i].y; }
        warp_sum<4>(sh, v);
    }
    const double cmx = sh.sums[0] / cnt, cmy = sh.sums[1] / cnt, cMx = sh.sums[2] / cnt, cMy = sh.sums[3] / cnt;
    __syncwarp();
    {
        double v[4] = {0, 0, 0, 0};
        for (int i = lane; i < n; i += 32) if (mask[i]) {
            v[0] += fabs(dst[i].x - cmx); v[1] += fabs(dst[i].y - cmy); v[2] += fabs(src[i].x - cMx); v[3] += fabs(src[i].y - cMy);
        }
        warp_sum<4>(sh, v);
    }
    double smx = sh.sums[0], smy = sh.sums[1], sMx = sh.sums[2], sMy = sh.sums[3];
    __syncwarp();
    double bestH[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) bestH[i] = sh.V[i];
    int eig_iters = 0;
    const bool refit_ok = !(fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON || fabs(sMx) < DBL_EPSILON || fabs(sMy) < DBL_EPSILON);
    if (refit_ok) {
        smx = cnt / smx; smy = cnt / smy; sMx = cnt / sMx; sMy = cnt / sMy;
        double acc[45];
#pragma unroll
        for (int k = 0; k < 45; ++k) acc[k] = 0.0;
        for (int i = lane; i < n; i += 32) {
            if (!mask[i]) continue;
            const double x = (dst[i].x - cmx) * smx, y = (dst[i].y - cmy) * smy, X = (src[i].x - cMx) * sMx, Y = (src[i].y - cMy) * sMy;
            const double Lx[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x};
            const double Ly[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
            int k = 0;
#pragma unroll
            for (int a = 0; a < 9; ++a) {
#pragma unroll
                for (int b = a; b < 9; ++b) { acc[k] = __fma_rn(Ly[a], Ly[b], __fma_rn(Lx[a], Lx[b], acc[k])); ++k; }
            }
        }
        warp_sum<45>(sh, acc);
        for (int t = lane; t < 45; t += 32) {     // unpack the upper triangle
            int a = 0, k = t;
            while (k >= 9 - a) { k -= 9 - a; ++a; }
            const int b = a + k;
            sh.A[a][b] = sh.sums[t]; sh.A[b][a] = sh.sums[t];
        }
        __syncwarp();
        // smallest eigenvector by shifted inverse iteration (cv::eigen runs a Jacobi; see header of smallest... note)
        double tr = 0;
        for (int k = 0; k < 9; ++k) tr += sh.A[k][k];
        const double shift = 1e-13 * tr + 1e-300;
        if (lane < 9) sh.V[lane] = 1.0 + 0.37 * lane;
        __syncwarp();
        // (A + shift I) is inverted once; every iteration is then a 9x9 matrix-vector product
        for (int e = lane; e < 9 * 18; e += 32) {
            const int r = e / 18, k = e - r * 18;
            sh.G[e] = k < 9 ? sh.A[r][k] + (k == r ? shift : 0.0) : (k - 9 == r ? 1.0 : 0.0);
        }
        __syncwarp();
        warp_gj_spd<9, 18>(sh.G);
        for (int e = lane; e < 81; e += 32) { const int r = e / 9, k = e - r * 9; sh.Ainv[r][k] = sh.G[r * 18 + 9 + k] / sh.G[r * 18 + r]; }
        __syncwarp();
        // six inverse-iteration steps, normalised once at the end: the growth per step is bounded by 1 / shift <= 1e13 / trace, far
        // from overflow, and the sign of the vector is irrelevant (H is scaled by 1 / h33 below)
        double x = 0;
        for (int it = 0; it < 6; ++it) {
            x = 0;
            if (lane < 9) for (int k = 0; k < 9; ++k) x = __fma_rn(sh.Ainv[lane][k], sh.V[k], x);
            __syncwarp();
            if (lane < 9) sh.V[lane] = x;
            __syncwarp();
        }
        {
            double nrm = 0;
            const double xx = lane < 9 ? x * x : 0.0;
#pragma unroll
            for (int k = 0; k < 9; ++k) nrm += __shfl_sync(0xffffffffu, xx, k);
            const double inv = rsqrt(nrm);
            if (lane < 9) sh.V[lane] = x * inv;
            __syncwarp();
            eig_iters = 6;
        }
        double h0[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) h0[i] = sh.V[i];
        const double inv[9] = {1.0 / smx, 0, cmx, 0, 1.0 / smy, cmy, 0, 0, 1};
        const double nrm[9] = {sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1};
        double t[9], H[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) t[3 * r + c] = inv[3 * r] * h0[c] + inv[3 * r + 1] * h0[3 + c] + inv[3 * r + 2] * h0[6 + c];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) H[3 * r + c] = t[3 * r] * nrm[c] + t[3 * r + 1] * nrm[3 + c] + t[3 * r + 2] * nrm[6 + c];
        const double sc = 1.0 / H[8];
        if (isfinite(sc)) {
#pragma unroll
            for (int i = 0; i < 9; ++i) bestH[i] = H[i] * sc;
        }
    }
    const long long t_eig = clock64();
    // ---- LMSolver(HomographyRefineCallback, maxIters = 10) over all nine elements of H, scaled by 1 / h33 afterwards (cv2 4.13) ----
    // J^T J is singular along the scale gauge h -> (1 + e) h.  cv2 solves (A + lambda diag(D)) d = v with cv::solve(DECOMP_EIG) and takes
    // max |diag| of cv::invert(A, DECOMP_EIG): eigen-decomposition + a back substitution that drops every eigenvalue
    // |w| <= 2 eps sum(w) = 2 eps trace -- a truncated pseudo-inverse.  Here: ONE elimination of [B | I] per iteration, B = the damped
    // matrix (lambda > 0) or A + mu h^ h^T (lambda == 0: the gauge direction h^ = h / |h| is the exact null vector of J, deflated with
    // mu = trace / 9 and projected out of the step again, which IS the pseudo-inverse with only the gauge dropped).  1 / |inv(B)|_F is a
    // lower bound of B's smallest eigenvalue: above twice cv2's threshold nothing else can have been dropped and the step is exact;
    // otherwise (ill-conditioned consensus sets: inliers in a corner of the frame) lane 0 runs the eigen-decomposition and applies
    // cv2's rule literally (eig9.h).  No frame of the reference clip takes the second route; tests force it.
    __syncwarp();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) sh.lmx[i] = bestH[i];
    }
    __syncwarp();
    rf_accumulate(sh, src, dst, mask, n, sh.lmx, true);
    lm_adopt(sh);
    if (lane < 9) sh.lmD[lane] = sh.lmA[lane][lane];
    __syncwarp();
    double sumD = 0;
    for (int k = 0; k < 9; ++k) sumD += sh.lmD[k];
    double S = sh.sums[54], rinf = sh.sums[55];
    double lambda = 1.0, lc = 0.75;
    const bool force_eig = g_lm_force_eig != 0;
    int it = 0, n_eig = 0, last_sweeps = 0;
    while (true) {
        double tr = 0;
        for (int k = 0; k < 9; ++k) tr += sh.lmA[k][k];
        const double thr = (tr + lambda * sumD) * (2.0 * DBL_EPSILON);      // cv2's truncation threshold for this system
        const bool gauge = lambda == 0.0;
        const double mu = tr / 9.0;
        if (gauge) {
            double nrm2 = 0;
            for (int k = 0; k < 9; ++k) nrm2 += sh.lmx[k] * sh.lmx[k];
            if (lane < 9) sh.X[lane] = sh.lmx[lane] / sqrt(nrm2);
        }
        __syncwarp();
        for (int e = lane; e < 9 * 18; e += 32) {
            const int r = e / 18, k = e - r * 18;
            double val;
            if (k < 9) {
                val = sh.lmA[r][k];
                if (k == r) val += lambda * sh.lmD[k];
                if (gauge) val += mu * sh.X[r] * sh.X[k];
            } else val = (k - 9 == r) ? 1.0 : 0.0;
            sh.G[e] = val;
        }
        __syncwarp();
        bool fast = warp_gj_spd<9, 18>(sh.G);
        double fro2 = 0;
        for (int e = lane; e < 81; e += 32) {
            const int r = e / 9, k = e - r * 9;
            const double bi = sh.G[r * 18 + 9 + k] / sh.G[r * 18 + r];
            sh.Ainv[r][k] = bi; fro2 += bi * bi;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) fro2 += __shfl_xor_sync(0xffffffffu, fro2, o);
        __syncwarp();
        fast = fast && !force_eig && isfinite(fro2) && fro2 > 0.0 && (1.0 / sqrt(fro2) > 2.0 * thr);
        if (fast) {
            double d = 0;
            if (lane < 9) for (int k = 0; k < 9; ++k) d = __fma_rn(sh.Ainv[lane][k], sh.lmv[k], d);
            if (gauge) {
                const double t = lane < 9 ? sh.X[lane] * d : 0.0;
                double dot = 0;
#pragma unroll
                for (int a = 0; a < 9; ++a) dot += __shfl_sync(0xffffffffu, t, a);
                if (lane < 9) { d -= sh.X[lane] * dot; sh.P[lane] = sh.Ainv[lane][lane] - sh.X[lane] * sh.X[lane] / mu; }
            } else if (lane < 9) sh.P[lane] = sh.Ainv[lane][lane];
            if (lane < 9) sh.lmd[lane] = d;
        } else {
#ifdef BM_JACOBI_SERIAL
            if (lane == 0) {
                for (int r = 0; r < 9; ++r) for (int k = 0; k < 9; ++k) sh.G[r * 9 + k] = sh.lmA[r][k] + (k == r ? lambda * sh.lmD[k] : 0.0);
                last_sweeps = bm_jacobi9(sh.G, sh.lmW, &sh.A[0][0]);
                bm_eig_pinv9(sh.lmW, &sh.A[0][0], sh.lmv, sh.lmd, sh.P);
            }
#else
            for (int e = lane; e < 81; e += 32) { const int r = e / 9, k = e - r * 9; sh.G[e] = sh.lmA[r][k] + (k == r ? lambda * sh.lmD[k] : 0.0); }
            __syncwarp();
            last_sweeps = jacobi9_warp(sh.G, sh.lmW, &sh.A[0][0]);
            if (lane == 0) bm_eig_pinv9(sh.lmW, &sh.A[0][0], sh.lmv, sh.lmd, sh.P);
#endif
            ++n_eig;
        }
        __syncwarp();
        if (lane < 9) sh.lmxd[lane] = sh.lmx[lane] - sh.lmd[lane];
        __syncwarp();
        // error AND normal equations at the trial point in one pass over the points: when the step is accepted (the usual case) the
        // sums are simply adopted below instead of being recomputed (same code, same order: bit-identical to a second pass)
        rf_accumulate(sh, src, dst, mask, n, sh.lmxd, true);
        const double Sd = sh.sums[54];
        // dS = d . (-A d + 2 v), tdv = d . v: lane a forms row a, the nine terms are added in index order (as a serial loop would)
        double ta = 0, ua = 0;
        if (lane < 9) {
            double Ad = 0;
            for (int b = 0; b < 9; ++b) Ad += sh.lmA[lane][b] * sh.lmd[b];
            ta = sh.lmd[lane] * (-Ad + 2.0 * sh.lmv[lane]);
            ua = sh.lmd[lane] * sh.lmv[lane];
        }
        double dS = 0, tdv = 0;
#pragma unroll
        for (int a = 0; a < 9; ++a) { dS += __shfl_sync(0xffffffffu, ta, a); tdv += __shfl_sync(0xffffffffu, ua, a); }
        const double R = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1.0);
        if (R > 0.75) { lambda *= 0.5; if (lambda < lc) lambda = 0.0; }
        else if (R < 0.25) {
            double nu = (Sd - S) / (fabs(tdv) > DBL_EPSILON ? tdv : 1.0) + 2.0;
            nu = fmin(fmax(nu, 2.0), 10.0);
            if (lambda == 0.0) {
                double maxval = DBL_EPSILON;       // max |diag(pinv(A))|: lambda == 0, so this iteration's decomposition is A's
                for (int c = 0; c < 9; ++c) maxval = fmax(maxval, fabs(sh.P[c]));
                lambda = lc = 1.0 / maxval;
                nu *= 0.5;
            }
            lambda *= nu;
        }
        const bool accept = Sd < S;
        __syncwarp();
        if (accept) {
            S = Sd;
            if (lane < 9) sh.lmx[lane] = sh.lmxd[lane];
            lm_adopt(sh);
            rinf = sh.sums[55];
        }
        ++it;
        double dinf = lane < 9 ? fabs(sh.lmd[lane]) : 0.0;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) dinf = fmax(dinf, __shfl_xor_sync(0xffffffffu, dinf, o));
        dinf = __shfl_sync(0xffffffffu, dinf, 0);
        if (!(it < 10 && dinf >= (double)FLT_EPSILON && rinf >= (double)FLT_EPSILON)) break;
    }
    if (lane == 0) {
        for (int i = 0; i < 8; ++i) out->H[i] = sh.lmx[i] / sh.lmx[8];
        out->H[8] = 1.0;
        out->ok = 1; out->lm_iters = it; out->jacobi_sweeps = eig_iters | (n_eig << 8) | (last_sweeps << 16);
        out->cyc[4] = t_eig - t_start; out->cyc[5] = clock64() - t_eig; out->cyc[7] = clock64() - t_start;
        atomicAdd(&g_lm_stats[0], 1ull); atomicAdd(&g_lm_stats[1], (unsigned long long)it); atomicAdd(&g_lm_stats[2], (unsigned long long)n_eig);
        sh.cmd = 0;
    }
    __syncwarp();
    __syncthreads();                            // releases the helper warps (cmd == 0)
}

cudaError_t bm_lm_force_eig(int on) { return cudaMemcpyToSymbol(g_lm_force_eig, &on, sizeof(int)); }

cudaError_t bm_lm_stats(unsigned long long out[3], int reset) {
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpyFromSymbol(out, g_lm_stats, 3 * sizeof(unsigned long long));
    if (e == cudaSuccess && reset) { const unsigned long long z[3] = {0, 0, 0}; e = cudaMemcpyToSymbol(g_lm_stats, z, sizeof(z)); }
    return e;
}

cudaError_t bm_launch_ransac(const float2* d_src, const float2* d_dst, const int* d_count, double thresh, int max_iters, double confidence,
                             uint8_t* d_mask, BmRansacResult* d_out, cudaStream_t s) {
    const size_t smem = ((sizeof(RsShared) + 15) & ~(size_t)15) + 2 * RS_SMEM_PTS * sizeof(float2);
    const size_t smem2 = ((sizeof(RfShared) + 15) & ~(size_t)15) + 2 * RS_SMEM_PTS * sizeof(float2);
    cudaError_t e;
    BM_SMEM_OPTIN(k_ransac_homography, smem, e);
    if (e != cudaSuccess) return e;
    BM_SMEM_OPTIN(k_ransac_refine, smem2, e);
    if (e != cudaSuccess) return e;
    BM_COUNT_LAUNCHES(2);
    if ((e = bm_launch_pdl(k_ransac_homography, dim3(1), dim3(RS_THREADS), smem, s, d_src, d_dst, d_count, thresh, max_iters, confidence, d_mask, d_out)) != cudaSuccess) return e;
    return bm_launch_pdl(k_ransac_refine, dim3(1), dim3(32 * RF_WARPS), smem2, s, d_src, d_dst, d_count, (const uint8_t*)d_mask, d_out);
}
