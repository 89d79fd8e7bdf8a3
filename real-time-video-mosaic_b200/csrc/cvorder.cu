// cvorder.cu -- stage entry point for the cv2-order retainBest emulation (cvorder.cuh), used by the parity tests.
#include "../../include/b200mosaic.h"
#include "cvorder.cuh"
#include <vector>

// shared memory: per-row scratch of the pairing passes first, then the keys when they fit (the passes re-read them ~2 lg n times)
#define CVO_SMEM_BYTES (220 * 1024)

template <typename KeyT>
__global__ void __launch_bounds__(CVO_THREADS) k_cvo_retain_best(KeyT* __restrict__ g_keys, int* __restrict__ idx, int n, int n_points,
                                                                 int* __restrict__ m_out) {
    extern __shared__ __align__(16) unsigned char cvo_smem[];
    __shared__ CvoShared sh;
    CvoRows rows;
    const int nrows = cvo_rows_needed(n);
    rows.bind(cvo_smem, nrows);
    const size_t used = CvoRows::bytes(nrows);
    KeyT* keys = g_keys;
    if (used + (size_t)n * sizeof(KeyT) <= CVO_SMEM_BYTES) {
        keys = reinterpret_cast<KeyT*>(cvo_smem + used);
        for (int i = threadIdx.x; i < n; i += CVO_THREADS) keys[i] = g_keys[i];
    }
    for (int i = threadIdx.x; i < n; i += CVO_THREADS) idx[i] = i;
    __syncthreads();
    const int m = cvo_retain_best<KeyT>(keys, idx, n, n_points, sh, rows);
    if (threadIdx.x == 0) *m_out = m;
}

extern "C" bm_status bm_cv_retain_best(const float* h_resp, int n, int n_points, int as_u8, int* h_idx_out, int* m_out) {
    if (!h_resp || n < 0 || !h_idx_out || !m_out) { bm_set_error("bm_cv_retain_best: bad args"); return BM_ERR_ARG; }
    if (n == 0) { *m_out = 0; return BM_OK; }
    if (CvoRows::bytes(cvo_rows_needed(n)) > CVO_SMEM_BYTES) { bm_set_error("bm_cv_retain_best: n = %d exceeds the %d elements one CTA can rank", n, CVO_SMEM_BYTES / 16 / 32 * 1024); return BM_ERR_UNSUPPORTED; }
    void* d_keys = nullptr; int *d_idx = nullptr, *d_m = nullptr;
    std::vector<uint8_t> k8;
    const size_t kbytes = as_u8 ? (size_t)n : (size_t)n * 4;
    if (as_u8) {
        k8.resize(n);
        for (int i = 0; i < n; ++i) {
            if (h_resp[i] < 0.f || h_resp[i] > 255.f || h_resp[i] != (float)(int)h_resp[i]) { bm_set_error("bm_cv_retain_best: u8 keys must be integers 0..255"); return BM_ERR_ARG; }
            k8[i] = (uint8_t)h_resp[i];
        }
    }
    BM_CUDA_OK(cudaMalloc(&d_keys, kbytes));
    BM_CUDA_OK(cudaMalloc(&d_idx, (size_t)n * 4));
    BM_CUDA_OK(cudaMalloc(&d_m, 4));
    BM_CUDA_OK(cudaMemcpy(d_keys, as_u8 ? (const void*)k8.data() : (const void*)h_resp, kbytes, cudaMemcpyHostToDevice));
    cudaError_t e;
    if (as_u8) {
        BM_SMEM_OPTIN(k_cvo_retain_best<uint8_t>, CVO_SMEM_BYTES, e);
        BM_CUDA_OK(e);
        BM_COUNT_LAUNCHES(1), k_cvo_retain_best<uint8_t><<<1, CVO_THREADS, CVO_SMEM_BYTES>>>((uint8_t*)d_keys, d_idx, n, n_points, d_m);
    } else {
        BM_SMEM_OPTIN(k_cvo_retain_best<float>, CVO_SMEM_BYTES, e);
        BM_CUDA_OK(e);
        BM_COUNT_LAUNCHES(1), k_cvo_retain_best<float><<<1, CVO_THREADS, CVO_SMEM_BYTES>>>((float*)d_keys, d_idx, n, n_points, d_m);
    }
    BM_CUDA_OK(cudaGetLastError());
    BM_CUDA_OK(cudaMemcpy(m_out, d_m, 4, cudaMemcpyDeviceToHost));
    if (*m_out > 0) BM_CUDA_OK(cudaMemcpy(h_idx_out, d_idx, (size_t)*m_out * 4, cudaMemcpyDeviceToHost));
    cudaFree(d_keys); cudaFree(d_idx); cudaFree(d_m);
    return BM_OK;
}
