// stages.cu -- feature / matching / RANSAC stage entry points of the C ABI (small host arrays in, host arrays out).
#include "../../include/b200mosaic.h"
#include "common.cuh"
#include "orb.cuh"
#include "sift.cuh"
#include "match.cuh"
#include "ransac.cuh"
#include <vector>
#include <string.h>
#include <math.h>
#include <stdlib.h>

extern "C" int bm_keypoint_capacity(void) { return BM_KP_CAP; }

// device keypoints -> host rows (x, y, size, angle, response, octave) + descriptors
bm_status bm_download_keypoints(const BmKeypoints& k, int desc_bytes, float* h_kp, uint8_t* h_desc, int cap, int* n_out, cudaStream_t s) {
    int n = 0;
    BM_CUDA_OK(cudaMemcpyAsync(&n, k.count, sizeof(int), cudaMemcpyDeviceToHost, s));
    BM_CUDA_OK(cudaStreamSynchronize(s));
    if (n > cap) { bm_set_error("keypoint buffer too small: %d > %d", n, cap); return BM_ERR_ARG; }
    if (n_out) *n_out = n;
    if (n == 0) return BM_OK;
    std::vector<float2> pt(n); std::vector<float> sz(n), an(n), rs(n); std::vector<int> oc(n);
    BM_CUDA_OK(cudaMemcpyAsync(pt.data(), k.pt, n * sizeof(float2), cudaMemcpyDeviceToHost, s));
    BM_CUDA_OK(cudaMemcpyAsync(sz.data(), k.size, n * 4, cudaMemcpyDeviceToHost, s));
    BM_CUDA_OK(cudaMemcpyAsync(an.data(), k.angle, n * 4, cudaMemcpyDeviceToHost, s));
    BM_CUDA_OK(cudaMemcpyAsync(rs.data(), k.response, n * 4, cudaMemcpyDeviceToHost, s));
    BM_CUDA_OK(cudaMemcpyAsync(oc.data(), k.octave, n * 4, cudaMemcpyDeviceToHost, s));
    if (h_desc) BM_CUDA_OK(cudaMemcpyAsync(h_desc, k.desc, (size_t)n * desc_bytes, cudaMemcpyDeviceToHost, s));
    BM_CUDA_OK(cudaStreamSynchronize(s));
    if (h_kp) for (int i = 0; i < n; ++i) {
        float* r = h_kp + 6 * (size_t)i;
        r[0] = pt[i].x; r[1] = pt[i].y; r[2] = sz[i]; r[3] = an[i]; r[4] = rs[i]; r[5] = (float)oc[i];
    }
    return BM_OK;
}

extern "C" bm_status bm_orb_detect_and_compute(const uint8_t* d_gray, int h, int w, int nfeatures, float* h_kp, uint8_t* h_desc, int cap, int* n_out) {
    if (!d_gray || h < 64 || w < 64) { bm_set_error("bm_orb_detect_and_compute: bad args"); return BM_ERR_ARG; }
    BmOrb* o = nullptr; BmKeypoints k;
    if (bm_orb_create(&o, h, w, nfeatures, nullptr) != 0 || bm_kp_alloc(&k, 32) != 0) { bm_set_error("orb alloc: %s", cudaGetErrorString(cudaGetLastError())); return BM_ERR_CUDA; }
    cudaError_t e = bm_orb_detect(o, d_gray, &k);
    bm_status st = BM_OK;
    if (e != cudaSuccess) { bm_set_error("orb detect: %s", cudaGetErrorString(e)); st = BM_ERR_CUDA; }
    if (st == BM_OK) st = bm_download_keypoints(k, 32, h_kp, h_desc, cap, n_out, nullptr);
    if (st == BM_OK) {
        int ov = 0;
        cudaMemcpy(&ov, o->ctr + 32, 4, cudaMemcpyDeviceToHost);
        if (ov) { bm_set_error("ORB candidate / keypoint capacity exceeded"); st = BM_ERR_UNSUPPORTED; }
    }
    bm_kp_free(&k); bm_orb_destroy(o);
    return st;
}

extern "C" bm_status bm_sift_detect_and_compute(const uint8_t* d_gray, int h, int w, int nfeatures, float* h_kp, float* h_desc, int cap, int* n_out) {
    if (!d_gray || h < 32 || w < 32) { bm_set_error("bm_sift_detect_and_compute: bad args"); return BM_ERR_ARG; }
    BmSift* o = nullptr; BmKeypoints k;
    if (bm_sift_create(&o, h, w, nfeatures, nullptr) != 0) return BM_ERR_UNSUPPORTED;
    if (bm_kp_alloc(&k, 128) != 0) { bm_sift_destroy(o); return BM_ERR_CUDA; }
    cudaError_t e = bm_sift_detect(o, d_gray, &k);
    bm_status st = BM_OK;
    if (e != cudaSuccess) { bm_set_error("sift detect: %s", cudaGetErrorString(e)); st = BM_ERR_CUDA; }
    std::vector<uint8_t> d8((size_t)cap * 128);
    int n = 0;
    if (st == BM_OK) st = bm_download_keypoints(k, 128, h_kp, d8.data(), cap, &n, nullptr);
    if (st == BM_OK) {
        int c[8];
        bm_sift_counters(o, c);
        if (getenv("BM_SIFT_DEBUG")) fprintf(stderr, "sift counters: cand %d kp %d overflow %d selected %d raw %d thr %08x kpA %d listed %d\n", c[0], c[1], c[2], c[3], c[4], (unsigned)c[5], c[6], c[7]);
        if (c[2]) { bm_set_error("SIFT candidate / keypoint capacity exceeded"); st = BM_ERR_UNSUPPORTED; }
    }
    if (st == BM_OK) { if (h_desc) for (size_t i = 0; i < (size_t)n * 128; ++i) h_desc[i] = (float)d8[i]; if (n_out) *n_out = n; }
    bm_kp_free(&k); bm_sift_destroy(o);
    return st;
}

static bm_status run_match(int mode, const uint8_t* h_q, int nq, const uint8_t* h_t, int nt, int desc_bytes, double ratio,
                           int* oq, int* ot, float* od, int* m_out) {
    if (nq < 0 || nt < 0 || nq > BM_KP_CAP || nt > BM_KP_CAP) { bm_set_error("match: too many descriptors"); return BM_ERR_ARG; }
    BmKeypoints a, b; BmMatches m;
    if (bm_kp_alloc(&a, desc_bytes) != 0 || bm_kp_alloc(&b, desc_bytes) != 0 || bm_matches_alloc(&m) != 0) return BM_ERR_CUDA;
    cudaMemcpy(a.desc, h_q, (size_t)nq * desc_bytes, cudaMemcpyHostToDevice);
    cudaMemcpy(b.desc, h_t, (size_t)nt * desc_bytes, cudaMemcpyHostToDevice);
    cudaMemcpy(a.count, &nq, 4, cudaMemcpyHostToDevice);
    cudaMemcpy(b.count, &nt, 4, cudaMemcpyHostToDevice);
    cudaMemset(a.pt, 0, BM_KP_CAP * sizeof(float2)); cudaMemset(b.pt, 0, BM_KP_CAP * sizeof(float2));
    cudaError_t e = mode == 0 ? bm_match_hamming(a, b, m, nullptr) : bm_match_l2_ratio(a, b, m, ratio, nullptr);
    int M = 0;
    if (e == cudaSuccess) e = cudaMemcpy(&M, m.count, 4, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && M > 0) {
        cudaMemcpy(oq, m.q, M * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(ot, m.t, M * 4, cudaMemcpyDeviceToHost);
        e = cudaMemcpy(od, m.dist, M * 4, cudaMemcpyDeviceToHost);
    }
    if (m_out) *m_out = M;
    bm_kp_free(&a); bm_kp_free(&b); bm_matches_free(&m);
    BM_CUDA_OK(e);
    return BM_OK;
}

// measurement: the SIFT matcher (tcgen05 kNN + merge + ratio + sort) on nq x nt random descriptors, `reps` times between two CUDA events
// on its launching stream; *flops = 2 * nq * nt * 128 per pair (the dense contraction)
extern "C" bm_status bm_match_l2_ms(int nq, int nt, int reps, double* ms_per_pair, double* flops_per_pair) {
    if (nq < 2 || nt < 2 || nq > BM_KP_CAP || nt > BM_KP_CAP || reps < 1 || !ms_per_pair) return BM_ERR_ARG;
    BmKeypoints a, b; BmMatches m;
    if (bm_kp_alloc(&a, 128) != 0 || bm_kp_alloc(&b, 128) != 0 || bm_matches_alloc(&m) != 0) return BM_ERR_CUDA;
    std::vector<uint8_t> ha((size_t)nq * 128), hb((size_t)nt * 128);
    unsigned st = 12345u;
    for (auto* v : {&ha, &hb}) for (auto& x : *v) { st = st * 1664525u + 1013904223u; x = (uint8_t)((st >> 24) & 0x7f); }
    cudaMemcpy(a.desc, ha.data(), ha.size(), cudaMemcpyHostToDevice); cudaMemcpy(b.desc, hb.data(), hb.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(a.count, &nq, 4, cudaMemcpyHostToDevice); cudaMemcpy(b.count, &nt, 4, cudaMemcpyHostToDevice);
    cudaMemset(a.pt, 0, BM_KP_CAP * sizeof(float2)); cudaMemset(b.pt, 0, BM_KP_CAP * sizeof(float2));
    cudaStream_t s = nullptr; cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    for (int r = -2; r < reps && e == cudaSuccess; ++r) {
        if (r == 0) e = cudaEventRecord(e0, s);
        if (e == cudaSuccess) e = bm_match_l2_ratio(a, b, m, 0.7, s);
    }
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventRecord(e1, s);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (s) cudaStreamDestroy(s);
    bm_kp_free(&a); bm_kp_free(&b); bm_matches_free(&m);
    BM_CUDA_OK(e);
    *ms_per_pair = (double)ms / reps;
    if (flops_per_pair) *flops_per_pair = 2.0 * nq * nt * 128.0;
    return BM_OK;
}

extern "C" bm_status bm_match_hamming_crosscheck(const uint8_t* q, int nq, const uint8_t* t, int nt, int* oq, int* ot, float* od, int* m_out) {
    return run_match(0, q, nq, t, nt, 32, 0.0, oq, ot, od, m_out);
}

extern "C" bm_status bm_match_l2_knn2_ratio(const float* q, int nq, const float* t, int nt, double ratio, int* oq, int* ot, float* od, int* m_out) {
    std::vector<uint8_t> q8((size_t)nq * 128), t8((size_t)nt * 128);
    for (size_t i = 0; i < q8.size(); ++i) { const float v = q[i]; if (v < 0.f || v > 255.f || v != floorf(v)) { bm_set_error("L2 matcher expects SIFT descriptors (integers 0..255)"); return BM_ERR_ARG; } q8[i] = (uint8_t)v; }
    for (size_t i = 0; i < t8.size(); ++i) { const float v = t[i]; if (v < 0.f || v > 255.f || v != floorf(v)) { bm_set_error("L2 matcher expects SIFT descriptors (integers 0..255)"); return BM_ERR_ARG; } t8[i] = (uint8_t)v; }
    return run_match(1, q8.data(), nq, t8.data(), nt, 128, ratio, oq, ot, od, m_out);
}

static bm_status ransac_host(const float* h_src, const float* h_dst, int n, double thresh, int max_iters, double confidence, BmRansacResult* res);

extern "C" bm_status bm_ransac_profile(const float* h_src, const float* h_dst, int n, double thresh, int max_iters, double confidence,
                                       double H[9], long long cycles[8], int* lm_iters, int* jacobi_sweeps) {
    BmRansacResult r;
    bm_status st = ransac_host(h_src, h_dst, n, thresh, max_iters, confidence, &r);
    if (st != BM_OK) return st;
    if (H) memcpy(H, r.H, 72);
    if (cycles) memcpy(cycles, r.cyc, sizeof(r.cyc));
    if (lm_iters) *lm_iters = r.lm_iters;
    if (jacobi_sweeps) *jacobi_sweeps = r.jacobi_sweeps;
    return BM_OK;
}

extern "C" bm_status bm_debug_lm_force_eig(int on) {
    const cudaError_t e = bm_lm_force_eig(on);
    if (e != cudaSuccess) { bm_set_error("bm_debug_lm_force_eig: %s", cudaGetErrorString(e)); return BM_ERR_CUDA; }
    return BM_OK;
}

extern "C" bm_status bm_debug_lm_stats(unsigned long long out[3], int reset) {
    if (!out) return BM_ERR_ARG;
    const cudaError_t e = bm_lm_stats(out, reset);
    if (e != cudaSuccess) { bm_set_error("bm_debug_lm_stats: %s", cudaGetErrorString(e)); return BM_ERR_CUDA; }
    return BM_OK;
}

extern "C" bm_status bm_ransac_homography(const float* h_src, const float* h_dst, int n, double thresh, int max_iters, double confidence,
                                          double H[9], int* ok, int* iters, int* n_inliers) {
    if (!H) return BM_ERR_ARG;
    BmRansacResult r;
    bm_status st = ransac_host(h_src, h_dst, n, thresh, max_iters, confidence, &r);
    if (st != BM_OK) return st;
    memcpy(H, r.H, 72);
    if (ok) *ok = r.ok;
    if (iters) *iters = r.iters;
    if (n_inliers) *n_inliers = r.n_inliers;
    return BM_OK;
}

static bm_status ransac_host(const float* h_src, const float* h_dst, int n, double thresh, int max_iters, double confidence, BmRansacResult* res) {
    double Hd[9]; double* H = Hd; int* ok = nullptr; int* iters = nullptr; int* n_inliers = nullptr;
    if (n < 0 || n > BM_KP_CAP || !h_src || !h_dst || !H) { bm_set_error("bm_ransac_homography: bad args"); return BM_ERR_ARG; }
    float2 *ds = nullptr, *dd = nullptr; int* dc = nullptr; uint8_t* dm = nullptr; BmRansacResult* dr = nullptr;
    BM_CUDA_OK(cudaMalloc(&ds, (n + 1) * sizeof(float2))); BM_CUDA_OK(cudaMalloc(&dd, (n + 1) * sizeof(float2)));
    BM_CUDA_OK(cudaMalloc(&dc, 4)); BM_CUDA_OK(cudaMalloc(&dm, n + 1)); BM_CUDA_OK(cudaMalloc(&dr, sizeof(BmRansacResult)));
    cudaMemcpy(ds, h_src, n * sizeof(float2), cudaMemcpyHostToDevice);
    cudaMemcpy(dd, h_dst, n * sizeof(float2), cudaMemcpyHostToDevice);
    cudaMemcpy(dc, &n, 4, cudaMemcpyHostToDevice);
    cudaError_t e = bm_launch_ransac(ds, dd, dc, thresh, max_iters, confidence, dm, dr, nullptr);
    BmRansacResult r; memset(&r, 0, sizeof(r));
    if (e == cudaSuccess) e = cudaMemcpy(&r, dr, sizeof(r), cudaMemcpyDeviceToHost);
    cudaFree(ds); cudaFree(dd); cudaFree(dc); cudaFree(dm); cudaFree(dr);
    BM_CUDA_OK(e);
    (void)H; (void)ok; (void)iters; (void)n_inliers;
    *res = r;
    return BM_OK;
}

extern "C" bm_status bm_orb_debug_level(const uint8_t* d_gray, int h, int w, int level, uint8_t* h_img, uint8_t* h_score, int* lw, int* lh) {
    if (!d_gray || level < 0 || level >= BM_ORB_LEVELS) return BM_ERR_ARG;
    BmOrb* o = nullptr; BmKeypoints k;
    if (bm_orb_create(&o, h, w, 700, nullptr) != 0 || bm_kp_alloc(&k, 32) != 0) return BM_ERR_CUDA;
    cudaError_t e = bm_orb_detect(o, d_gray, &k);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    const BmOrbLevel& L = o->lv.l[level];
    if (lw) *lw = L.w;
    if (lh) *lh = L.h;
    if (e == cudaSuccess && h_img) e = cudaMemcpy(h_img, o->pyr + L.off, (size_t)L.w * L.h, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && h_score) e = cudaMemcpy(h_score, o->score + L.off, (size_t)L.w * L.h, cudaMemcpyDeviceToHost);
    bm_kp_free(&k); bm_orb_destroy(o);
    BM_CUDA_OK(e);
    return BM_OK;
}

extern "C" bm_status bm_sift_pyramid_ms(const uint8_t* d_gray, int h, int w, int reps, double* ms_per_frame, double* algorithmic_bytes) {
    if (!d_gray || reps < 1 || !ms_per_frame) return BM_ERR_ARG;
    BmSift* o = nullptr;
    cudaStream_t s = nullptr;
    BM_CUDA_OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    if (bm_sift_create(&o, h, w, 700, s) != 0) { cudaStreamDestroy(s); return BM_ERR_CUDA; }
    float ms = 0.f;
    cudaError_t e = bm_sift_time_pyramid(o, d_gray, reps, &ms);
    bm_sift_destroy(o); cudaStreamDestroy(s);
    BM_CUDA_OK(e);
    *ms_per_frame = (double)ms / reps;
    if (algorithmic_bytes) *algorithmic_bytes = 256.0 * (double)h * (double)w;
    return BM_OK;
}

extern "C" bm_status bm_sift_debug_level(const uint8_t* d_gray, int h, int w, int octave, int level, int dog, float* h_out, int* lw, int* lh, int* noct) {
    if (!d_gray) return BM_ERR_ARG;
    BmSift* o = nullptr; BmKeypoints k;
    if (bm_sift_create(&o, h, w, 700, nullptr) != 0) return BM_ERR_CUDA;
    if (bm_kp_alloc(&k, 128) != 0) { bm_sift_destroy(o); return BM_ERR_CUDA; }
    if (noct) *noct = bm_sift_num_octaves(o);
    bm_status st = BM_OK;
    if (octave < 0 || octave >= bm_sift_num_octaves(o) || level < 0 || level > (dog ? 4 : 5)) st = BM_ERR_ARG;
    cudaError_t e = cudaSuccess;
    if (st == BM_OK) {
        e = bm_sift_detect(o, d_gray, &k);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        int ww = 0, hh = 0;
        const float* p = bm_sift_level_ptr(o, octave, level, dog, &ww, &hh);
        if (lw) *lw = ww;
        if (lh) *lh = hh;
        if (e == cudaSuccess && h_out) e = cudaMemcpy(h_out, p, (size_t)ww * hh * sizeof(float), cudaMemcpyDeviceToHost);
    }
    bm_kp_free(&k); bm_sift_destroy(o);
    if (st != BM_OK) return st;
    BM_CUDA_OK(e);
    return BM_OK;
}
