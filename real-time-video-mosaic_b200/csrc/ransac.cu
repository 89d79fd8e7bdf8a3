// ransac.cu -- batched RANSAC homography + LS refit + Levenberg-Marquardt, one persistent CTA per point set.
// Reference: cv2.findHomography(p_cur, p_prev, cv2.RANSAC, 2.0) at main.py:856-857; algorithm = OpenCV 4.x
// RANSACPointSetRegistrator::run / HomographyEstimatorCallback / LMSolver (restated in oracle/ransac.py).
//
// Sequential semantics kept on a parallel evaluator: the cv::RNG stream, the subset rejection rules, "first strictly
// better hypothesis wins" and the adaptive iteration cap are evaluated in iteration order by one thread, while the
// expensive parts -- 4-point solves and inlier counting of a batch of 32 hypotheses (one per warp), the 45-term normal
// equation sums of the refit and of each LM step -- run across the CTA.
#include "ransac.cuh"
#include <float.h>
#include <math.h>

#define RS_BATCH 32
#define RS_THREADS 512

struct RsShared {
    int idx[RS_BATCH][4];
    int valid[RS_BATCH];
    int good[RS_BATCH];
    double Hb[RS_BATCH][9];
    double bestH[9];
    double red[32][46];          // cross-warp reduction scratch
    double sums[46];
    double A[9][9], V[9][9];     // Jacobi workspace
    double lmA[8][8], lmv[8], lmd[8], lmx[8], lmxd[8], lmD[8];
    int niters, iter, best_good, stop, batch, fail_at;
    int n_in;
    double S, Sd;
};

__device__ __forceinline__ unsigned rng_next(unsigned long long& st) {
    st = (unsigned long long)(unsigned)st * 4164903690ULL + (st >> 32);
    return (unsigned)st;
}

__device__ bool collinear4(const float2* p) {       // haveCollinearPoints(m, 4): point 3 against pairs of 0..2
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
__device__ double det3pts(const float2& a, const float2& b, const float2& c) {
    const double a0 = a.x, a1 = a.y, b0 = b.x, b1 = b.y, c0 = c.x, c1 = c.y;
    return a0 * (b1 * 1.0 - 1.0 * c1) - a1 * (b0 * 1.0 - 1.0 * c0) + 1.0 * (b0 * c1 - b1 * c0);
}
__device__ bool check_subset(const float2* s, const float2* d) {
    if (collinear4(s) || collinear4(d)) return false;
    const int tt[4][3] = {{0, 1, 2}, {1, 2, 3}, {0, 2, 3}, {0, 1, 3}};
    int neg = 0;
    for (int i = 0; i < 4; ++i)
        neg += (det3pts(s[tt[i][0]], s[tt[i][1]], s[tt[i][2]]) * det3pts(d[tt[i][0]], d[tt[i][1]], d[tt[i][2]]) < 0.0) ? 1 : 0;
    return neg == 0 || neg == 4;
}

// 4-point homography: normalise like HomographyEstimatorCallback::runKernel, solve the 8x8 system (h33 = 1 in normalised
// coordinates) by Gaussian elimination with partial pivoting, denormalise, scale so that H[8] = 1.
__device__ bool solve4(const float2* M, const float2* m, double* H) {
    double cMx = 0, cMy = 0, cmx = 0, cmy = 0;
    for (int i = 0; i < 4; ++i) { cMx += M[i].x; cMy += M[i].y; cmx += m[i].x; cmy += m[i].y; }
    cMx /= 4; cMy /= 4; cmx /= 4; cmy /= 4;
    double sMx = 0, sMy = 0, smx = 0, smy = 0;
    for (int i = 0; i < 4; ++i) { sMx += fabs(M[i].x - cMx); sMy += fabs(M[i].y - cMy); smx += fabs(m[i].x - cmx); smy += fabs(m[i].y - cmy); }
    if (fabs(sMx) < DBL_EPSILON || fabs(sMy) < DBL_EPSILON || fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON) return false;
    sMx = 4 / sMx; sMy = 4 / sMy; smx = 4 / smx; smy = 4 / smy;
    double a[8][9];
    for (int i = 0; i < 4; ++i) {
        const double X = (M[i].x - cMx) * sMx, Y = (M[i].y - cMy) * sMy, x = (m[i].x - cmx) * smx, y = (m[i].y - cmy) * smy;
        double* r0 = a[2 * i]; double* r1 = a[2 * i + 1];
        r0[0] = X; r0[1] = Y; r0[2] = 1; r0[3] = 0; r0[4] = 0; r0[5] = 0; r0[6] = -x * X; r0[7] = -x * Y; r0[8] = x;
        r1[0] = 0; r1[1] = 0; r1[2] = 0; r1[3] = X; r1[4] = Y; r1[5] = 1; r1[6] = -y * X; r1[7] = -y * Y; r1[8] = y;
    }
    for (int c = 0; c < 8; ++c) {
        int piv = c; double best = fabs(a[c][c]);
        for (int r = c + 1; r < 8; ++r) if (fabs(a[r][c]) > best) { best = fabs(a[r][c]); piv = r; }
        if (best < 1e-12) return false;
        if (piv != c) for (int k = c; k < 9; ++k) { const double t = a[c][k]; a[c][k] = a[piv][k]; a[piv][k] = t; }
        const double inv = 1.0 / a[c][c];
        for (int r = c + 1; r < 8; ++r) {
            const double f = a[r][c] * inv;
            if (f != 0.0) for (int k = c; k < 9; ++k) a[r][k] -= f * a[c][k];
        }
    }
    double h[9];
    for (int r = 7; r >= 0; --r) {
        double s = a[r][8];
        for (int k = r + 1; k < 8; ++k) s -= a[r][k] * h[k];
        h[r] = s / a[r][r];
    }
    h[8] = 1.0;
    // H = invHnorm * H0 * Hnorm2
    const double inv[9] = {1.0 / smx, 0, cmx, 0, 1.0 / smy, cmy, 0, 0, 1};
    const double nrm[9] = {sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1};
    double t[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) t[3 * r + c] = inv[3 * r] * h[c] + inv[3 * r + 1] * h[3 + c] + inv[3 * r + 2] * h[6 + c];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) H[3 * r + c] = t[3 * r] * nrm[c] + t[3 * r + 1] * nrm[3 + c] + t[3 * r + 2] * nrm[6 + c];
    const double s = 1.0 / H[8];
    if (!isfinite(s)) return false;
    for (int i = 0; i < 9; ++i) H[i] *= s;
    return true;
}

// HomographyEstimatorCallback::computeError: float32, the model cast to float, err <= thresh^2
__device__ __forceinline__ bool is_inlier(const float* Hf, float2 M, float2 m, float t2) {
    const float ww = __fdiv_rn(1.f, __fadd_rn(__fadd_rn(__fmul_rn(Hf[6], M.x), __fmul_rn(Hf[7], M.y)), 1.f));
    const float dx = __fsub_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(Hf[0], M.x), __fmul_rn(Hf[1], M.y)), Hf[2]), ww), m.x);
    const float dy = __fsub_rn(__fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(Hf[3], M.x), __fmul_rn(Hf[4], M.y)), Hf[5]), ww), m.y);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) <= t2;
}

__device__ int update_num_iters(double p, double ep, int model_points, int max_iters) {
    p = fmin(fmax(p, 0.0), 1.0); ep = fmin(fmax(ep, 0.0), 1.0);
    double num = fmax(1.0 - p, DBL_MIN);
    double denom = 1.0 - pow(1.0 - ep, (double)model_points);
    if (denom < DBL_MIN) return 0;
    num = log(num); denom = log(denom);
    return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : __double2int_rn(num / denom);
}

// CTA-wide sum of NV doubles per thread -> sh.sums[0..NV)
template <int NV>
__device__ void block_sum(RsShared& sh, const double* v) {
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
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh.red[w][threadIdx.x];
        sh.sums[threadIdx.x] = s;
    }
    __syncthreads();
}

// cyclic Jacobi on the symmetric 9x9 sh.A; eigenvectors in the columns of sh.V; returns the index of the smallest eigenvalue
__device__ int jacobi9(RsShared& sh) {
    const int n = 9;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) sh.V[i][j] = (i == j) ? 1.0 : 0.0;
    sh.n_in = 0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        sh.n_in = sweep;
        double off = 0, diag = 0;
        for (int i = 0; i < n; ++i) { diag += sh.A[i][i] * sh.A[i][i]; for (int j = i + 1; j < n; ++j) off += sh.A[i][j] * sh.A[i][j]; }
        if (off <= 1e-32 * diag || off == 0.0) break;
        for (int p = 0; p < n - 1; ++p) for (int q = p + 1; q < n; ++q) {
            const double apq = sh.A[p][q];
            if (fabs(apq) < 1e-300) continue;
            const double theta = (sh.A[q][q] - sh.A[p][p]) / (2.0 * apq);
            const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
            for (int k = 0; k < n; ++k) { const double akp = sh.A[k][p], akq = sh.A[k][q]; sh.A[k][p] = c * akp - s * akq; sh.A[k][q] = s * akp + c * akq; }
            for (int k = 0; k < n; ++k) { const double apk = sh.A[p][k], aqk = sh.A[q][k]; sh.A[p][k] = c * apk - s * aqk; sh.A[q][k] = s * apk + c * aqk; }
            for (int k = 0; k < n; ++k) { const double vkp = sh.V[k][p], vkq = sh.V[k][q]; sh.V[k][p] = c * vkp - s * vkq; sh.V[k][q] = s * vkp + c * vkq; }
        }
    }
    int best = 0;
    for (int i = 1; i < n; ++i) if (sh.A[i][i] < sh.A[best][best]) best = i;
    return best;
}

// 8x8 solve (Gaussian elimination, partial pivoting) of a symmetric positive system; returns false if singular
__device__ bool solve8(const double A[8][8], const double* b, double* x) {
    double a[8][9];
    for (int i = 0; i < 8; ++i) { for (int j = 0; j < 8; ++j) a[i][j] = A[i][j]; a[i][8] = b[i]; }
    for (int c = 0; c < 8; ++c) {
        int piv = c; double best = fabs(a[c][c]);
        for (int r = c + 1; r < 8; ++r) if (fabs(a[r][c]) > best) { best = fabs(a[r][c]); piv = r; }
        if (best == 0.0) return false;
        if (piv != c) for (int k = c; k < 9; ++k) { const double t = a[c][k]; a[c][k] = a[piv][k]; a[piv][k] = t; }
        const double inv = 1.0 / a[c][c];
        for (int r = c + 1; r < 8; ++r) { const double f = a[r][c] * inv; for (int k = c; k < 9; ++k) a[r][k] -= f * a[c][k]; }
    }
    for (int r = 7; r >= 0; --r) { double s = a[r][8]; for (int k = r + 1; k < 8; ++k) s -= a[r][k] * x[k]; x[r] = s / a[r][r]; }
    return true;
}

// residuals / Jacobian sums of HomographyRefineCallback over the inliers: 36 (JtJ upper) + 8 (Jtr) + 1 (|r|^2) + 1 (max|r|)
__device__ void lm_accumulate(RsShared& sh, const float2* src, const float2* dst, const uint8_t* mask, int n, const double* h, bool want_j) {
    double acc[46];
#pragma unroll
    for (int k = 0; k < 46; ++k) acc[k] = 0.0;
    double rmax = 0.0;
    if (threadIdx.x < 256) {
        for (int i = threadIdx.x; i < n; i += 256) {
            if (!mask[i]) continue;
            const double Mx = src[i].x, My = src[i].y;
            const double ww = 1.0 / (h[6] * Mx + h[7] * My + 1.0);
            const double xi = (h[0] * Mx + h[1] * My + h[2]) * ww, yi = (h[3] * Mx + h[4] * My + h[5]) * ww;
            const double rx = xi - (double)dst[i].x, ry = yi - (double)dst[i].y;
            acc[44] += rx * rx + ry * ry;
            rmax = fmax(rmax, fmax(fabs(rx), fabs(ry)));
            if (want_j) {
                const double j0[8] = {Mx * ww, My * ww, ww, 0, 0, 0, -Mx * ww * xi, -My * ww * xi};
                const double j1[8] = {0, 0, 0, Mx * ww, My * ww, ww, -Mx * ww * yi, -My * ww * yi};
                int k = 0;
#pragma unroll
                for (int a = 0; a < 8; ++a) {
#pragma unroll
                    for (int b = a; b < 8; ++b) acc[k++] += j0[a] * j0[b] + j1[a] * j1[b];
                }
#pragma unroll
                for (int a = 0; a < 8; ++a) acc[36 + a] += j0[a] * rx + j1[a] * ry;
            }
        }
    }
    // max-reduce rmax separately through slot 45 (as a sum of per-warp maxima would be wrong): use shuffles + shared
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, o));
    acc[45] = 0.0;
    block_sum<45>(sh, acc);
    if ((threadIdx.x & 31) == 0) sh.red[threadIdx.x >> 5][45] = rmax;
    __syncthreads();
    if (threadIdx.x == 0) { double mx = 0; for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mx = fmax(mx, sh.red[w][45]); sh.sums[45] = mx; }
    __syncthreads();
}

__global__ void __launch_bounds__(RS_THREADS, 1) k_ransac_homography(const float2* __restrict__ src, const float2* __restrict__ dst,
                                                                     const int* __restrict__ countp, double thresh, int max_iters, double confidence,
                                                                     uint8_t* __restrict__ mask, BmRansacResult* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RsShared& sh = *reinterpret_cast<RsShared*>(smem_raw);
    const int n = *countp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float t2 = (float)(thresh * thresh);
    if (tid == 0) {
        out->ok = 0; out->n_points = n; out->iters = 0; out->n_inliers = 0; out->lm_iters = 0;
        for (int i = 0; i < 9; ++i) out->H[i] = 0.0;
    }
    if (n < 4) return;
    if (n == 4) {       // findHomography: npoints == 4 -> runKernel only, no refinement
        if (tid == 0) {
            double H[9];
            float2 s4[4], d4[4];
            for (int i = 0; i < 4; ++i) { s4[i] = src[i]; d4[i] = dst[i]; }
            if (solve4(s4, d4, H)) { for (int i = 0; i < 9; ++i) out->H[i] = H[i]; out->ok = 1; out->n_inliers = 4; }
        }
        return;
    }
    __shared__ unsigned long long rng;
    __shared__ long long cyc[8];
    if (tid < 8) cyc[tid] = 0;
    const long long t_start = clock64();
    if (tid == 0) { rng = 0xFFFFFFFFFFFFFFFFULL; sh.niters = max_iters > 1 ? max_iters : 1; sh.iter = 0; sh.best_good = 0; sh.stop = 0; }
    __syncthreads();
    while (true) {
        // ---- one thread: next batch of accepted subsets, in cv::RNG order (getSubset, 10000 attempts each) ----
        long long t0 = clock64();
        if (tid == 0) {
            int B = sh.niters - sh.iter; if (B > RS_BATCH) B = RS_BATCH;
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
        // ---- warp b: hypothesis b: 4-point solve by lane 0, inlier count by all lanes ----
        for (int hb = warp; hb < nb; hb += (int)(blockDim.x >> 5)) {
            if (lane == 0) {
                float2 s4[4], d4[4];
                for (int i = 0; i < 4; ++i) { s4[i] = src[sh.idx[hb][i]]; d4[i] = dst[sh.idx[hb][i]]; }
                double H[9];
                const bool ok = solve4(s4, d4, H);
                sh.valid[hb] = ok ? 1 : 0;
                if (ok) for (int i = 0; i < 9; ++i) sh.Hb[hb][i] = H[i];
            }
            __syncwarp();
            int cnt = 0;
            if (sh.valid[hb]) {
                float Hf[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) Hf[i] = (float)sh.Hb[hb][i];
                for (int i = lane; i < n; i += 32) cnt += is_inlier(Hf, src[i], dst[i], t2) ? 1 : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            }
            if (lane == 0) sh.good[hb] = cnt;
            __syncwarp();
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
    // ---- LS refit on the inliers: normalised DLT (runKernel), 9x9 LtL, smallest eigenvector ----
    const int cnt = sh.best_good;
    {
        double v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = 0.0;
        for (int i = tid; i < n; i += blockDim.x) if (mask[i]) { v[0] += dst[i].x; v[1] += dst[i].y; v[2] += src[i].x; v[3] += src[i].y; }
        block_sum<4>(sh, v);
    }
    const double cmx = sh.sums[0] / cnt, cmy = sh.sums[1] / cnt, cMx = sh.sums[2] / cnt, cMy = sh.sums[3] / cnt;
    __syncthreads();
    {
        double v[4] = {0, 0, 0, 0};
        for (int i = tid; i < n; i += blockDim.x) if (mask[i]) {
            v[0] += fabs(dst[i].x - cmx); v[1] += fabs(dst[i].y - cmy); v[2] += fabs(src[i].x - cMx); v[3] += fabs(src[i].y - cMy);
        }
        block_sum<4>(sh, v);
    }
    double smx = sh.sums[0], smy = sh.sums[1], sMx = sh.sums[2], sMy = sh.sums[3];
    __syncthreads();
    bool refit_ok = !(fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON || fabs(sMx) < DBL_EPSILON || fabs(sMy) < DBL_EPSILON);
    if (refit_ok) {
        smx = cnt / smx; smy = cnt / smy; sMx = cnt / sMx; sMy = cnt / sMy;
        double acc[45];
#pragma unroll
        for (int k = 0; k < 45; ++k) acc[k] = 0.0;
        if (tid < 256) {
            for (int i = tid; i < n; i += 256) {
                if (!mask[i]) continue;
                const double x = (dst[i].x - cmx) * smx, y = (dst[i].y - cmy) * smy, X = (src[i].x - cMx) * sMx, Y = (src[i].y - cMy) * sMy;
                const double Lx[9] = {X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x};
                const double Ly[9] = {0, 0, 0, X, Y, 1, -y * X, -y * Y, -y};
                int k = 0;
#pragma unroll
                for (int a = 0; a < 9; ++a) {
#pragma unroll
                    for (int b = a; b < 9; ++b) acc[k++] += Lx[a] * Lx[b] + Ly[a] * Ly[b];
                }
            }
        }
        block_sum<45>(sh, acc);
        if (tid == 0) {
            { const long long t1 = clock64(); cyc[3] += t1 - tp; tp = t1; }
            int k = 0;
            for (int a = 0; a < 9; ++a) for (int b = a; b < 9; ++b) { sh.A[a][b] = sh.sums[k]; sh.A[b][a] = sh.sums[k]; ++k; }
            const int e = jacobi9(sh);
            { const long long t1 = clock64(); cyc[4] += t1 - tp; tp = t1; }
            double h0[9];
            for (int i = 0; i < 9; ++i) h0[i] = sh.V[i][e];
            const double inv[9] = {1.0 / smx, 0, cmx, 0, 1.0 / smy, cmy, 0, 0, 1};
            const double nrm[9] = {sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1};
            double t[9], H[9];
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) t[3 * r + c] = inv[3 * r] * h0[c] + inv[3 * r + 1] * h0[3 + c] + inv[3 * r + 2] * h0[6 + c];
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) H[3 * r + c] = t[3 * r] * nrm[c] + t[3 * r + 1] * nrm[3 + c] + t[3 * r + 2] * nrm[6 + c];
            const double s = 1.0 / H[8];
            if (isfinite(s)) for (int i = 0; i < 9; ++i) sh.bestH[i] = H[i] * s;
        }
        __syncthreads();
    }
    // ---- LMSolver(HomographyRefineCallback, maxIters = 10) on the 8 free parameters ----
    if (tid < 8) sh.lmx[tid] = sh.bestH[tid];
    __syncthreads();
    lm_accumulate(sh, src, dst, mask, n, sh.lmx, true);
    if (tid == 0) {
        int k = 0;
        for (int a = 0; a < 8; ++a) for (int b = a; b < 8; ++b) { sh.lmA[a][b] = sh.sums[k]; sh.lmA[b][a] = sh.sums[k]; ++k; }
        for (int a = 0; a < 8; ++a) { sh.lmv[a] = sh.sums[36 + a]; sh.lmD[a] = sh.lmA[a][a]; }
        sh.S = sh.sums[44];
    }
    __syncthreads();
    double lambda = 1.0, lc = 0.75;       // only thread 0's copies matter
    int it = 0;
    __shared__ int lm_go, lm_accept;
    __shared__ double rinf;
    if (tid == 0) rinf = sh.sums[45];
    __syncthreads();
    while (true) {
        if (tid == 0) {
            double Ap[8][8];
            for (int a = 0; a < 8; ++a) for (int b = 0; b < 8; ++b) Ap[a][b] = sh.lmA[a][b] + (a == b ? lambda * sh.lmD[a] : 0.0);
            if (!solve8(Ap, sh.lmv, sh.lmd)) for (int a = 0; a < 8; ++a) sh.lmd[a] = 0.0;
            for (int a = 0; a < 8; ++a) sh.lmxd[a] = sh.lmx[a] - sh.lmd[a];
        }
        __syncthreads();
        lm_accumulate(sh, src, dst, mask, n, sh.lmxd, false);
        if (tid == 0) {
            const double Sd = sh.sums[44], S = sh.S;
            double dS = 0, tdv = 0;
            for (int a = 0; a < 8; ++a) {
                double Ad = 0;
                for (int b = 0; b < 8; ++b) Ad += sh.lmA[a][b] * sh.lmd[b];
                dS += sh.lmd[a] * (-Ad + 2.0 * sh.lmv[a]);
                tdv += sh.lmd[a] * sh.lmv[a];
            }
            const double R = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1.0);
            if (R > 0.75) { lambda *= 0.5; if (lambda < lc) lambda = 0.0; }
            else if (R < 0.25) {
                double nu = (Sd - S) / (fabs(tdv) > DBL_EPSILON ? tdv : 1.0) + 2.0;
                nu = fmin(fmax(nu, 2.0), 10.0);
                if (lambda == 0.0) {
                    double maxval = DBL_EPSILON;       // max |diag(inv(A))|
                    for (int c = 0; c < 8; ++c) {
                        double e[8], x[8];
                        for (int a = 0; a < 8; ++a) e[a] = (a == c) ? 1.0 : 0.0;
                        if (solve8(sh.lmA, e, x)) maxval = fmax(maxval, fabs(x[c]));
                    }
                    lambda = lc = 1.0 / maxval;
                    nu *= 0.5;
                }
                lambda *= nu;
            }
            lm_accept = (Sd < S) ? 1 : 0;
            if (lm_accept) { sh.S = Sd; for (int a = 0; a < 8; ++a) sh.lmx[a] = sh.lmxd[a]; }
        }
        __syncthreads();
        if (lm_accept) {
            lm_accumulate(sh, src, dst, mask, n, sh.lmx, true);
            if (tid == 0) {
                int k = 0;
                for (int a = 0; a < 8; ++a) for (int b = a; b < 8; ++b) { sh.lmA[a][b] = sh.sums[k]; sh.lmA[b][a] = sh.sums[k]; ++k; }
                for (int a = 0; a < 8; ++a) sh.lmv[a] = sh.sums[36 + a];
                rinf = sh.sums[45];
            }
        }
        if (tid == 0) {
            ++it;
            double dinf = 0;
            for (int a = 0; a < 8; ++a) dinf = fmax(dinf, fabs(sh.lmd[a]));
            lm_go = (it < 10 && dinf >= (double)FLT_EPSILON && rinf >= (double)FLT_EPSILON) ? 1 : 0;
        }
        __syncthreads();
        if (!lm_go) break;
    }
    if (tid == 0) {
        for (int i = 0; i < 8; ++i) out->H[i] = sh.lmx[i];
        out->H[8] = 1.0;
        out->ok = 1; out->iters = sh.iter; out->n_inliers = sh.best_good; out->lm_iters = it; out->jacobi_sweeps = sh.n_in;
        cyc[5] = clock64() - tp; cyc[6] = clock64() - t_start;
        for (int i = 0; i < 8; ++i) out->cyc[i] = cyc[i];
    }
}

cudaError_t bm_launch_ransac(const float2* d_src, const float2* d_dst, const int* d_count, double thresh, int max_iters, double confidence,
                             uint8_t* d_mask, BmRansacResult* d_out, cudaStream_t s) {
    static bool attr_set = false;
    const size_t smem = sizeof(RsShared);
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k_ransac_homography, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    BM_COUNT_LAUNCHES(1), k_ransac_homography<<<1, RS_THREADS, smem, s>>>(d_src, d_dst, d_count, thresh, max_iters, confidence, d_mask, d_out);
    return cudaGetLastError();
}
