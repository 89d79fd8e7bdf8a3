// match.cu -- VideMosaic.match on the device (reference: main.py:676-698; tie rules: SURVEY.md A.6, oracle/matching.py).
//   ORB : 256-bit Hamming on CUDA cores (__popc), warp-shuffle arg-min, both directions, mutual-NN filter.
//   SIFT: exact integer L2 (descriptors are integers 0..255): sum(a-b)^2 = |a|^2+|b|^2-2ab with the dot products on the tensor
//         cores (tcgen05 bf16 -> fp32 in TMEM, match_tc.cu); top-2 per query with (distance, trainIdx) lexicographic ties; Lowe
//         ratio in double on float32 distances as the reference does.
//   Both: stable sort by distance (the reference's sorted(key=distance)) as an O(m^2) rank count in one block.
#include "match.cuh"
#include <string.h>

int bm_matches_alloc(BmMatches* m) {
    memset(m, 0, sizeof(*m));
    const size_t n = BM_KP_CAP;
    bool ok = cudaMalloc(&m->q, n * 4) == cudaSuccess && cudaMalloc(&m->t, n * 4) == cudaSuccess && cudaMalloc(&m->dist, n * 4) == cudaSuccess &&
              cudaMalloc(&m->src, n * 8) == cudaSuccess && cudaMalloc(&m->dst, n * 8) == cudaSuccess && cudaMalloc(&m->count, 4) == cudaSuccess &&
              cudaMalloc(&m->nn_q2t, n * 4) == cudaSuccess && cudaMalloc(&m->nn_t2q, n * 4) == cudaSuccess && cudaMalloc(&m->tq, n * 4) == cudaSuccess &&
              cudaMalloc(&m->tt, n * 4) == cudaSuccess && cudaMalloc(&m->d_q2t, n * 4) == cudaSuccess && cudaMalloc(&m->d_t2q, n * 4) == cudaSuccess &&
              cudaMalloc(&m->td, n * 4) == cudaSuccess && cudaMalloc(&m->d2_q2t, n * 4) == cudaSuccess && cudaMalloc(&m->nn2_q2t, n * 4) == cudaSuccess &&
              cudaMalloc(&m->l2_part, n * BM_L2_SPLIT * sizeof(int4)) == cudaSuccess;
    if (ok) cudaMemset(m->count, 0, 4);
    return ok ? 0 : -1;
}
void bm_matches_free(BmMatches* m) {
    cudaFree(m->q); cudaFree(m->t); cudaFree(m->dist); cudaFree(m->src); cudaFree(m->dst); cudaFree(m->count); cudaFree(m->nn_q2t);
    cudaFree(m->nn_t2q); cudaFree(m->tq); cudaFree(m->tt); cudaFree(m->d_q2t); cudaFree(m->d_t2q); cudaFree(m->td); cudaFree(m->d2_q2t);
    cudaFree(m->nn2_q2t); cudaFree(m->l2_part);
    memset(m, 0, sizeof(*m));
}

// one warp per row of A: nearest row of B under Hamming, first index on ties
// blockIdx.y = 0: rows of A against B; 1: rows of B against A (both directions of the cross check in one launch)
__global__ void __launch_bounds__(256) k_hamming_nn(const uint8_t* __restrict__ A0, const int* __restrict__ nA0p, const uint8_t* __restrict__ B0,
                                                    const int* __restrict__ nB0p, int* __restrict__ nn0, float* __restrict__ nd0,
                                                    int* __restrict__ nn1, float* __restrict__ nd1) {
    const bool rev = blockIdx.y != 0;
    const uint8_t* __restrict__ A = rev ? B0 : A0; const uint8_t* __restrict__ B = rev ? A0 : B0;
    const int* nAp = rev ? nB0p : nA0p; const int* nBp = rev ? nA0p : nB0p;
    int* __restrict__ nn = rev ? nn1 : nn0; float* __restrict__ nd = rev ? nd1 : nd0;
    const int a = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nA = *nAp, nB = *nBp;
    if (a >= nA) return;
    const uint4* ap = reinterpret_cast<const uint4*>(A + (size_t)a * 32);
    const uint4 a0 = __ldg(ap), a1 = __ldg(ap + 1);
    int best = 1 << 30, bi = 1 << 30;
    for (int b = lane; b < nB; b += 32) {
        const uint4* bp = reinterpret_cast<const uint4*>(B + (size_t)b * 32);
        const uint4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
        const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                      __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
        if (d < best) { best = d; bi = b; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int od = __shfl_xor_sync(0xffffffffu, best, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (od < best || (od == best && oi < bi)) { best = od; bi = oi; }
    }
    if (lane == 0) { nn[a] = (nB > 0) ? bi : -1; nd[a] = (float)best; }
}

// block-wide ordered compaction helper: returns the exclusive prefix of `flag` over the block and the block total
__device__ __forceinline__ int block_prefix(bool flag, int* total, int* warp_sums) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) warp_sums[warp] = __popc(bal);
    __syncthreads();
    int before = 0, tot = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { const int v = warp_sums[w]; if (w < warp) before += v; tot += v; }
    __syncthreads();
    *total = tot;
    return before + __popc(bal & ((1u << lane) - 1u));
}

// select (mode 0: mutual NN; mode 1: Lowe ratio), keep query order, then stable sort by distance and gather the points
__global__ void __launch_bounds__(1024) k_select_sort(int mode, double ratio, const int* __restrict__ nqp, const int* __restrict__ ntp,
                                                      BmMatches m, const float2* __restrict__ pt_cur, const float2* __restrict__ pt_prev) {
    BM_PDL_WAIT();
    __shared__ int warp_sums[32];
    __shared__ int s_count;
    const int nq = *nqp, nt = *ntp;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    for (int base = 0; base < nq; base += blockDim.x) {
        const int q = base + threadIdx.x;
        bool ok = false;
        int t = -1; float d = 0.f;
        if (q < nq && nt > 0) {
            t = m.nn_q2t[q]; d = m.d_q2t[q];
            if (mode == 0) ok = t >= 0 && m.nn_t2q[t] == q;
            else ok = nt >= 2 && t >= 0 && m.nn2_q2t[q] >= 0 && ((double)d < ratio * (double)m.d2_q2t[q]);
        }
        int tot;
        const int pos = block_prefix(ok, &tot, warp_sums);
        const int off = s_count;
        if (ok) { m.tq[off + pos] = q; m.tt[off + pos] = t; m.td[off + pos] = d; }
        __syncthreads();
        if (threadIdx.x == 0) s_count = off + tot;
        __syncthreads();
    }
    const int M = s_count;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        const float di = m.td[i];
        int rank = 0;
        for (int j = 0; j < M; ++j) { const float dj = m.td[j]; rank += (dj < di || (dj == di && j < i)) ? 1 : 0; }
        const int q = m.tq[i], t = m.tt[i];
        m.q[rank] = q; m.t[rank] = t; m.dist[rank] = di;
        m.src[rank] = pt_cur[q]; m.dst[rank] = pt_prev[t];
    }
    if (threadIdx.x == 0) *m.count = M;
}

cudaError_t bm_match_hamming(const BmKeypoints& cur, const BmKeypoints& prev, BmMatches& m, cudaStream_t s) {
    const int blocks = (BM_KP_CAP * 32) / 256;
    BM_COUNT_LAUNCHES(1), k_hamming_nn<<<dim3(blocks, 2), 256, 0, s>>>(cur.desc, cur.count, prev.desc, prev.count, m.nn_q2t, m.d_q2t, m.nn_t2q, m.d_t2q);
    BM_COUNT_LAUNCHES(1);
    return bm_launch_pdl(k_select_sort, dim3(1), dim3(1024), 0, s, 0, 0.0, (const int*)cur.count, (const int*)prev.count, m, (const float2*)cur.pt, (const float2*)prev.pt);
}

cudaError_t bm_match_l2_ratio(const BmKeypoints& cur, const BmKeypoints& prev, BmMatches& m, double ratio, cudaStream_t s) {
    cudaError_t e = bm_launch_l2_knn2_tc(cur.desc, cur.count, prev.desc, prev.count, m.l2_part, m.nn_q2t, m.d_q2t, m.nn2_q2t, m.d2_q2t, s);
    if (e != cudaSuccess) return e;
    BM_COUNT_LAUNCHES(1);
    return bm_launch_pdl(k_select_sort, dim3(1), dim3(1024), 0, s, 1, ratio, (const int*)cur.count, (const int*)prev.count, m, (const float2*)cur.pt, (const float2*)prev.pt);
}
