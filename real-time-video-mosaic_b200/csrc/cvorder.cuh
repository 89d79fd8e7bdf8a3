// cvorder.cuh -- cv::KeyPointsFilter::retainBest with cv2's exact output ORDER, as a block-wide device routine.
//
// The reference keeps the keypoints in whatever order `detectAndCompute` returns them (/root/reference/main.py:112,718) and
// that order decides the match list, the cv::RNG subsets of findHomography and therefore the homography.  In OpenCV 4.x the
// order is made by   std::nth_element(begin, begin + n - 1, end, response >)   followed by
// std::partition(begin + n, end, response >= kp[n - 1].response)   (features2d/src/keypoint.cpp), i.e. by libstdc++'s
// introselect (median of {first + 1, mid, last - 1} moved to first, unguarded Hoare partition, depth limit 2 lg n with a
// heap-select fallback, insertion sort below 4 elements).
//
// Neither algorithm is inherently serial.  A Hoare partition step scans for "stoppers" from both ends and swaps the k-th
// stopper from the left with the k-th stopper from the right while they have not crossed; swapped elements are never looked
// at again, so stopper ranks can be taken on the ORIGINAL array: with cL(p) = #left-stoppers at positions <= p and
// cR(p) = #right-stoppers at positions > p, the number of swaps is K = max_p min(cL(p), cR(p)), pair k <= K swaps
// L_k <-> R_k, and the cut is min(L_{K+1}, R_K).  One step is therefore two ballot/popc passes over the active range by one
// CTA; introselect needs ~lg n steps over geometrically shrinking ranges.  The result is the same permutation libstdc++
// produces, element for element (pinned against the real std::nth_element in tests/test_order_gpu.py).
#pragma once
#include "common.cuh"

#define CVO_THREADS 1024

struct CvoShared {
    int first, last, depth, cut, K, nL, nR, heap_done;
    int wL[32], wR[32];
};

template <typename KeyT>
__device__ __forceinline__ void cvo_swap(KeyT* keys, int* idx, int i, int j) {
    const KeyT k = keys[i]; keys[i] = keys[j]; keys[j] = k;
    const int t = idx[i]; idx[i] = idx[j]; idx[j] = t;
}

// One pairing pass over [lo, hi): left-stoppers = isL(key), right-stoppers = isR(key).  Performs the K swaps; leaves
// sh.K, sh.nL, sh.nR and sh.cut = min(L_{K+1}, R_K) (R_0 = hi, L_{nL+1} = INT_MAX).  All threads of the CTA must call it.
template <typename KeyT, class FL, class FR>
__device__ void cvo_pair_pass(KeyT* keys, int* idx, int* listL, int* listR, int lo, int hi, FL isL, FR isR, CvoShared& sh) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int len = hi - lo;
    const int seg = ((len + CVO_THREADS - 1) / CVO_THREADS) * 32;          // positions per warp, a multiple of 32
    const int s0 = lo + warp * seg, s1 = min(s0 + seg, hi);
    const unsigned le = 0xffffffffu >> (31 - lane);
    int cL = 0, cR = 0;
    for (int r0 = s0; r0 < s1; r0 += 32) {
        const int p = r0 + lane;
        bool l = false, r = false;
        if (p < s1) { const KeyT k = keys[p]; l = isL(k); r = isR(k); }
        cL += __popc(__ballot_sync(0xffffffffu, l));
        cR += __popc(__ballot_sync(0xffffffffu, r));
    }
    if (threadIdx.x == 0) sh.K = 0;
    if (lane == 0) { sh.wL[warp] = cL; sh.wR[warp] = cR; }
    __syncthreads();
    const int vL = sh.wL[lane], vR = sh.wR[lane];
    int preL = lane < warp ? vL : 0, sufR = lane > warp ? vR : 0, nL = vL, nR = vR;
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        preL += __shfl_xor_sync(0xffffffffu, preL, d);
        sufR += __shfl_xor_sync(0xffffffffu, sufR, d);
        nL += __shfl_xor_sync(0xffffffffu, nL, d);
        nR += __shfl_xor_sync(0xffffffffu, nR, d);
    }
    int runL = preL, runRge = sufR + cR, best = 0;
    for (int r0 = s0; r0 < s1; r0 += 32) {
        const int p = r0 + lane;
        bool l = false, r = false;
        if (p < s1) { const KeyT k = keys[p]; l = isL(k); r = isR(k); }
        const unsigned bl = __ballot_sync(0xffffffffu, l), br = __ballot_sync(0xffffffffu, r);
        const int cLle = runL + __popc(bl & le), cRgt = runRge - __popc(br & le);
        if (p < s1) best = max(best, min(cLle, cRgt));
        if (l) listL[cLle] = p;                                              // 1-based rank from the left
        if (r) listR[cRgt + 1] = p;                                          // 1-based rank from the right
        runL += __popc(bl);
        runRge -= __popc(br);
    }
    best = __reduce_max_sync(0xffffffffu, best);
    if (lane == 0 && best > 0) atomicMax(&sh.K, best);
    __syncthreads();
    const int K = sh.K;
    for (int k = threadIdx.x + 1; k <= K; k += CVO_THREADS) cvo_swap(keys, idx, listL[k], listR[k]);
    if (threadIdx.x == 0) {
        int cut = K + 1 <= nL ? listL[K + 1] : 0x7fffffff;
        cut = min(cut, K >= 1 ? listR[K] : hi);
        sh.cut = cut; sh.nL = nL; sh.nR = nR;
    }
    __syncthreads();
}

// libstdc++ __adjust_heap + __push_heap on [first, first + len) with comparator "a > b" (a min-heap of the largest values)
template <typename KeyT>
__device__ void cvo_adjust_heap(KeyT* keys, int* idx, int first, int hole, int len, KeyT vk, int vi) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (keys[first + child] > keys[first + child - 1]) --child;
        keys[first + hole] = keys[first + child]; idx[first + hole] = idx[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        keys[first + hole] = keys[first + child - 1]; idx[first + hole] = idx[first + child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && keys[first + parent] > vk) {
        keys[first + hole] = keys[first + parent]; idx[first + hole] = idx[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    keys[first + hole] = vk; idx[first + hole] = vi;
}

// introselect's fallback when the depth limit is hit (adversarial inputs only): serial, one thread
template <typename KeyT>
__device__ void cvo_heap_select(KeyT* keys, int* idx, int first, int middle, int last) {
    const int len = middle - first;
    if (len >= 2) {
        for (int parent = (len - 2) / 2;; --parent) {
            cvo_adjust_heap(keys, idx, first, parent, len, keys[first + parent], idx[first + parent]);
            if (parent == 0) break;
        }
    }
    for (int i = middle; i < last; ++i)
        if (keys[i] > keys[first]) {
            const KeyT vk = keys[i]; const int vi = idx[i];
            keys[i] = keys[first]; idx[i] = idx[first];
            cvo_adjust_heap(keys, idx, first, 0, len, vk, vi);
        }
}

// retainBest(n_points) over keys[0 .. n) / idx[0 .. n) (permuted together, in place).  Returns the number of survivors m;
// afterwards idx[0 .. m) are the survivors in cv2's order and keys[i] is the key of idx[i].  listL / listR: n + 2 ints each.
// Must be called by all CVO_THREADS threads of the CTA with identical arguments.
template <typename KeyT>
__device__ int cvo_retain_best(KeyT* keys, int* idx, int* listL, int* listR, int n, int n_points, CvoShared& sh) {
    if (n_points < 0 || n <= n_points) return n;
    if (n_points == 0) return 0;
    const int nth = n_points - 1;
    if (threadIdx.x == 0) { sh.first = 0; sh.last = n; sh.depth = 2 * (31 - __clz(n)); sh.heap_done = 0; }
    __syncthreads();
    while (true) {
        const int first = sh.first, last = sh.last, depth = sh.depth;
        __syncthreads();                                                   // everyone has read the range before thread 0 edits it
        if (last - first <= 3) break;
        if (depth == 0) {
            if (threadIdx.x == 0) {
                cvo_heap_select(keys, idx, first, nth + 1, last);
                cvo_swap(keys, idx, first, nth);
                sh.heap_done = 1;
            }
            __syncthreads();
            break;
        }
        if (threadIdx.x == 0) {
            sh.depth -= 1;
            const int a = first + 1, b = first + (last - first) / 2, c = last - 1;
            const KeyT ka = keys[a], kb = keys[b], kc = keys[c];
            int pick;
            if (ka > kb) pick = kb > kc ? b : (ka > kc ? c : a);
            else pick = ka > kc ? a : (kb > kc ? c : b);
            cvo_swap(keys, idx, first, pick);
        }
        __syncthreads();
        const KeyT pv = keys[first];
        cvo_pair_pass(keys, idx, listL, listR, first + 1, last,
                      [pv](KeyT k) { return !(k > pv); }, [pv](KeyT k) { return !(pv > k); }, sh);
        if (threadIdx.x == 0) {
            if (sh.cut <= nth) sh.first = sh.cut; else sh.last = sh.cut;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 && !sh.heap_done) {                               // __insertion_sort on <= 3 elements
        const int first = sh.first, last = sh.last;
        for (int i = first + 1; i < last; ++i) {
            const KeyT vk = keys[i]; const int vi = idx[i];
            int j = i;
            if (vk > keys[first]) {
                for (; j > first; --j) { keys[j] = keys[j - 1]; idx[j] = idx[j - 1]; }
            } else {
                for (; vk > keys[j - 1]; --j) { keys[j] = keys[j - 1]; idx[j] = idx[j - 1]; }
            }
            keys[j] = vk; idx[j] = vi;
        }
    }
    __syncthreads();
    const KeyT amb = keys[nth];
    cvo_pair_pass(keys, idx, listL, listR, n_points, n,
                  [amb](KeyT k) { return !(k >= amb); }, [amb](KeyT k) { return k >= amb; }, sh);
    const int m = n_points + sh.nR;
    __syncthreads();                                                       // sh may be reused by the caller's next selection
    return m;
}
