// eig9.h -- symmetric eigen-decomposition of the 9x9 normal matrix of the homography polish (ransac.cu, k_ransac_refine).
//
// cv2 4.13's LMSolver solves its damped normal equations with cv::solve(..., DECOMP_EIG) / cv::invert(..., DECOMP_EIG): a Jacobi
// eigen-decomposition followed by SVBkSb, which DROPS every eigenvalue with |w_i| <= 2 * DBL_EPSILON * sum(w) (a truncated
// pseudo-inverse).  The polish runs over all nine elements of H, so J^T J is singular along the scale gauge and that truncation is
// what defines the step.  The kernel takes this route only when a cheap bound cannot rule out that a second eigenvalue is near the
// threshold (ill-conditioned consensus sets); plain C++ so that the same code is unit-tested on the host (tests/test_eig9_cpu.py).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define BM_HD __host__ __device__
#else
#define BM_HD
#endif

// Cyclic Jacobi.  a: row-major 9x9 symmetric, destroyed.  w[9]: eigenvalues (unsorted).  v: row-major 9x9, eigenvectors in COLUMNS.
// An off-diagonal element is annihilated exactly once it is negligible against the geometric mean of its two diagonal elements, which
// keeps the relative accuracy of the small eigenvalues (the ones the truncation rule looks at).  Returns the number of sweeps.
BM_HD inline int bm_jacobi9(double* a, double* w, double* v) {
    const int n = 9;
    for (int i = 0; i < n * n; ++i) v[i] = 0.0;
    for (int i = 0; i < n; ++i) v[i * n + i] = 1.0;
    int sweep = 0;
    for (; sweep < 40; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < n - 1; ++p) {
            for (int q = p + 1; q < n; ++q) {
                const double apq = a[p * n + q];
                if (apq == 0.0) continue;
                const double app = a[p * n + p], aqq = a[q * n + q];
                if (fabs(apq) <= 1e-300 + 2.220446049250313e-19 * sqrt(fabs(app) * fabs(aqq))) { a[p * n + q] = 0.0; a[q * n + p] = 0.0; continue; }
                rotated = true;
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                a[p * n + p] = app - t * apq; a[q * n + q] = aqq + t * apq; a[p * n + q] = 0.0; a[q * n + p] = 0.0;
                for (int k = 0; k < n; ++k) {
                    if (k != p && k != q) {
                        const double akp = a[k * n + p], akq = a[k * n + q];
                        const double np_ = c * akp - s * akq, nq_ = s * akp + c * akq;
                        a[k * n + p] = np_; a[p * n + k] = np_; a[k * n + q] = nq_; a[q * n + k] = nq_;
                    }
                    const double vp = v[k * n + p], vq = v[k * n + q];
                    v[k * n + p] = c * vp - s * vq; v[k * n + q] = s * vp + c * vq;
                }
            }
        }
        if (!rotated) break;
    }
    for (int i = 0; i < n; ++i) w[i] = a[i * n + i];
    return sweep;
}

// What cv::solve(A, b, x, DECOMP_EIG) and the diagonal of cv::invert(A, Ai, DECOMP_EIG) return, from the decomposition above:
// x = sum_{kept k} v_k (v_k . b) / w_k,  diag_pinv[i] = sum_{kept k} v_ik^2 / w_k,  kept: |w_k| > 2 eps sum(w).  Returns the number kept.
BM_HD inline int bm_eig_pinv9(const double* w, const double* v, const double* b, double* x, double* diag_pinv) {
    const int n = 9;
    double sum = 0.0;
    for (int k = 0; k < n; ++k) sum += w[k];
    const double thr = sum * (2.0 * 2.220446049250313e-16);
    for (int i = 0; i < n; ++i) { x[i] = 0.0; diag_pinv[i] = 0.0; }
    int kept = 0;
    for (int k = 0; k < n; ++k) {
        if (fabs(w[k]) <= thr) continue;
        ++kept;
        double dot = 0.0;
        for (int j = 0; j < n; ++j) dot += v[j * n + k] * b[j];
        const double wi = 1.0 / w[k];
        dot *= wi;
        for (int i = 0; i < n; ++i) { x[i] += v[i * n + k] * dot; diag_pinv[i] += v[i * n + k] * v[i * n + k] * wi; }
    }
    return kept;
}
