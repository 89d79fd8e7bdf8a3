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
// L_k <-> R_k, and the cut is min(L_{K+1}, R_K).  One step is: one pass over the keys that leaves the two stopper bit masks of
// every 32-element row in shared memory, a pass over those masks for the ranks and K, and the swaps -- the partner of the k-th left
// stopper is found by a binary search over the per-row suffix counts plus __fns inside the row's mask, so no stopper lists are
// materialised.  Ranges of <= CVO_TAIL elements are processed on shared-memory copies of keys and indices, the last ones (<= CVO_WARP_TAIL) by one warp without block barriers
// (introselect needs ~lg n steps over geometrically shrinking ranges: most steps are small).  The result is the permutation
// libstdc++ produces, element for element (pinned against the real std::nth_element in tests/test_order_gpu.py).
#pragma once
#include "common.cuh"

#define CVO_THREADS 1024
#define CVO_TAIL 2048               // ranges up to this many elements are finished on shared-memory copies of keys and indices
#define CVO_WARP_TAIL 64            // ... and up to this many by warp 0 alone (two rows: no block barriers)

struct CvoShared {
    int first, last, depth, cut, K, nL, nR, heap_done, cutL, cutR;
    int wL[32], wR[32];
};

// per-row scratch in shared memory: stopper masks and the number of right-stoppers at or after the row's first element
struct CvoRows {
    unsigned* bl; unsigned* br; int* sufR; int* preL;
    int* tail_idx; void* tail_keys;                                         // CVO_TAIL entries each: the small steps run on copies
    int cap;                                                                // rows
    __host__ __device__ static size_t bytes(int rows) { return (size_t)rows * 16 + (size_t)CVO_TAIL * 8; }
    __device__ void bind(void* p, int rows) {
        bl = reinterpret_cast<unsigned*>(p); br = bl + rows; sufR = reinterpret_cast<int*>(br + rows); preL = sufR + rows; cap = rows;
        tail_idx = preL + rows; tail_keys = tail_idx + CVO_TAIL;
    }
};

template <typename KeyT>
__device__ __forceinline__ void cvo_swap(KeyT* keys, int* idx, int i, int j) {
    const KeyT k = keys[i]; keys[i] = keys[j]; keys[j] = k;
    const int t = idx[i]; idx[i] = idx[j]; idx[j] = t;
}

// One pairing pass over [lo, hi): left-stoppers = isL(key), right-stoppers = isR(key).  Performs the K swaps; leaves
// sh.K, sh.nL, sh.nR and sh.cut = min(L_{K+1}, R_K) (R_0 = hi, L_{nL+1} = INT_MAX).
// BLOCK = true : called by all CVO_THREADS threads of the CTA.   BLOCK = false : called by the 32 lanes of warp 0 only.
// Needs ceil(len / 32) (rounded up to a multiple of the warp count) <= rows.cap.
// sub_pos / sub_key: position whose key is to be read as sub_key during the pass (the pivot swap of introselect is applied by
// thread 0 through `mid` only after every thread has read the keys); `fin(cut, nL, nR)` runs on thread 0 before the last barrier.
template <bool BLOCK, typename KeyT, class FL, class FR, class Mid, class Fin>
__device__ void cvo_pair_pass(KeyT* keys, int* idx, int lo, int hi, FL isL, FR isR, CvoShared& sh, const CvoRows& rows,
                              int sub_pos, KeyT sub_key, Mid mid, Fin fin) {
    constexpr int NW = BLOCK ? CVO_THREADS / 32 : 1;
    const int warp = BLOCK ? (threadIdx.x >> 5) : 0, lane = threadIdx.x & 31;
    const int len = hi - lo;
    const int rpw = (len + 32 * NW - 1) / (32 * NW);                         // rows per warp
    const int r0 = warp * rpw;
    const unsigned le = 0xffffffffu >> (31 - lane);
    auto sync = [] { if (BLOCK) __syncthreads(); else __syncwarp(); };
    int cL = 0, cR = 0;
    for (int i = 0; i < rpw; ++i) {
        const int p = lo + (r0 + i) * 32 + lane;
        bool l = false, r = false;
        if (p < hi) { const KeyT k = p == sub_pos ? sub_key : keys[p]; l = isL(k); r = isR(k); }
        const unsigned bl = __ballot_sync(0xffffffffu, l), br = __ballot_sync(0xffffffffu, r);
        if (lane == 0) { rows.bl[r0 + i] = bl; rows.br[r0 + i] = br; }
        cL += __popc(bl); cR += __popc(br);
    }
    int preL = 0, sufR = 0, nL = cL, nR = cR;
    if (BLOCK) {
        if (threadIdx.x == 0) { sh.K = 0; sh.cutL = 0x7fffffff; sh.cutR = hi; }
        if (lane == 0) { sh.wL[warp] = cL; sh.wR[warp] = cR; }
        __syncthreads();
        if (threadIdx.x == 0) mid();                                        // every thread has read its keys
        const int vL = sh.wL[lane], vR = sh.wR[lane];
        preL = lane < warp ? vL : 0; sufR = lane > warp ? vR : 0; nL = vL; nR = vR;
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            preL += __shfl_xor_sync(0xffffffffu, preL, d);
            sufR += __shfl_xor_sync(0xffffffffu, sufR, d);
            nL += __shfl_xor_sync(0xffffffffu, nL, d);
            nR += __shfl_xor_sync(0xffffffffu, nR, d);
        }
    } else {
        if (lane == 0) { sh.cutL = 0x7fffffff; sh.cutR = hi; }
        __syncwarp();
        if (lane == 0) mid();
    }
    // ranks from the cached masks: lane i handles row r0 + i of every group of 32 rows
    int best = 0;
    {
        int runL = preL, runRge = sufR + cR;
        for (int i0 = 0; i0 < rpw; i0 += 32) {
            const int i = i0 + lane;
            const unsigned bl = i < rpw ? rows.bl[r0 + i] : 0u, br = i < rpw ? rows.br[r0 + i] : 0u;
            const int nl = __popc(bl), nr = __popc(br);
            int il = nl, ir = nr;                                             // inclusive scans over the 32 rows of the group
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int a = __shfl_up_sync(0xffffffffu, il, d), b = __shfl_up_sync(0xffffffffu, ir, d);
                if (lane >= d) { il += a; ir += b; }
            }
            const int rowPreL = runL + il - nl, rowSufR = runRge - (ir - nr);  // L before the row, R at or after the row's start
            if (i < rpw) {
                rows.preL[r0 + i] = rowPreL; rows.sufR[r0 + i] = rowSufR;
                // max over the row's positions of min(cL(<= p), cR(> p)): f(t) = rowPreL + popc(bl & upto(t)) grows and
                // g(t) = rowSufR - popc(br & upto(t)) shrinks along the row, so the maximum of min(f, g) sits at the crossing: binary
                // search for the last bit t with f(t) <= g(t), then compare t and t + 1
                int t = -1;                                                   // "before the row": f = rowPreL, g = rowSufR
#pragma unroll
                for (int step = 16; step; step >>= 1) {
                    const int c = t + step;                                   // c <= 30 + ... stays <= 31
                    const unsigned upto = 0xffffffffu >> (31 - c);
                    if (rowPreL + __popc(bl & upto) <= rowSufR - __popc(br & upto)) t = c;
                }
                {
                    const unsigned u0 = t < 0 ? 0u : 0xffffffffu >> (31 - t);
                    if (t >= 0) best = max(best, min(rowPreL + __popc(bl & u0), rowSufR - __popc(br & u0)));
                    if (t < 31) { const unsigned u1 = 0xffffffffu >> (30 - t); best = max(best, min(rowPreL + __popc(bl & u1), rowSufR - __popc(br & u1))); }
                }
            }
            runL += __shfl_sync(0xffffffffu, il, 31);
            runRge -= __shfl_sync(0xffffffffu, ir, 31);
        }
    }
    best = __reduce_max_sync(0xffffffffu, best);
    int K;
    if (BLOCK) {
        if (lane == 0 && best > 0) atomicMax(&sh.K, best);
        __syncthreads();
        K = sh.K;
    } else {
        __syncwarp();
        K = best;
    }
    // swaps: the thread that owns the k-th left stopper (k <= K) finds the k-th right stopper from the right and exchanges them.
    // Four rows per batch: all loads of a batch are issued before its stores, so the (global-memory) index array costs one
    // round trip per batch instead of one per row.
    // Rows are dealt round-robin over the warps here (the ranks <= K sit in the LEFT part of the range: contiguous segments would
    // leave the swaps to the first few warps).
    const int nrows = rpw * NW;
    for (int rb = warp; rb < nrows; rb += 4 * NW) {
        if (rows.preL[rb] >= K + 1) break;                                   // every later row has larger ranks
        int pp[4], qq[4];
        bool vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            vv[u] = false; pp[u] = 0; qq[u] = 0;
            const int row = rb + u * NW;
            if (row >= nrows) continue;
            const unsigned bl = rows.bl[row];
            const bool l = (bl >> lane) & 1u;
            const int rank = rows.preL[row] + __popc(bl & le);
            const int p = lo + row * 32 + lane;
            if (l && rank == K + 1) sh.cutL = p;
            if (l && rank <= K) {
                int a = 0, b = nrows - 1;                                     // last row with sufR >= rank (sufR is non-increasing)
                while (a < b) { const int m = (a + b + 1) >> 1; if (rows.sufR[m] >= rank) a = m; else b = m - 1; }
                const unsigned br = rows.br[a];
                const int after = rows.sufR[a] - __popc(br);                  // right stoppers in the rows behind row a
                const int bit = 31 - (int)__fns(__brev(br), 0, rank - after);  // (rank - after)-th set bit of br counted from bit 31
                const int q = lo + a * 32 + bit;
                if (rank == K) sh.cutR = q;
                vv[u] = true; pp[u] = p; qq[u] = q;
            }
        }
        KeyT kp[4], kq[4];
        int ip[4], iq[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (vv[u]) { kp[u] = keys[pp[u]]; kq[u] = keys[qq[u]]; ip[u] = idx[pp[u]]; iq[u] = idx[qq[u]]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) if (vv[u]) { keys[pp[u]] = kq[u]; keys[qq[u]] = kp[u]; idx[pp[u]] = iq[u]; idx[qq[u]] = ip[u]; }
    }
    sync();
    if (threadIdx.x == 0) { sh.nL = nL; sh.nR = nR; fin(min(sh.cutL, sh.cutR)); }
    sync();
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

// one introselect step on [sh.first, sh.last): median of three to the front, Hoare partition, range update.  Every thread derives the
// pivot itself; thread 0 applies the pivot swap once the keys have been read and narrows the range at the end (4 barriers per step).
// Returns false when the loop is over (range <= 3, or the heap-select fallback ran).
template <bool BLOCK, typename KeyT>
__device__ bool cvo_select_step(KeyT* keys, int* idx, int nth, CvoShared& sh, const CvoRows& rows) {
    auto sync = [] { if (BLOCK) __syncthreads(); else __syncwarp(); };
    const int first = sh.first, last = sh.last, depth = sh.depth;          // stable: written before the previous step's last barrier
    if (last - first <= 3) return false;
    if (depth == 0) {
        sync();
        if (threadIdx.x == 0) {
            cvo_heap_select(keys, idx, first, nth + 1, last);
            cvo_swap(keys, idx, first, nth);
            sh.heap_done = 1;
        }
        sync();
        return false;
    }
    const int a = first + 1, b = first + (last - first) / 2, c = last - 1;
    const KeyT ka = keys[a], kb = keys[b], kc = keys[c], kfirst = keys[first];
    int pick;
    if (ka > kb) pick = kb > kc ? b : (ka > kc ? c : a);
    else pick = ka > kc ? a : (kb > kc ? c : b);
    const KeyT pv = pick == a ? ka : (pick == b ? kb : kc);
    cvo_pair_pass<BLOCK>(keys, idx, first + 1, last, [pv](KeyT k) { return !(k > pv); }, [pv](KeyT k) { return !(pv > k); }, sh, rows,
                         pick, kfirst,
                         [&] { cvo_swap(keys, idx, first, pick); sh.depth = depth - 1; },
                         [&](int cut) { if (cut <= nth) sh.first = cut; else sh.last = cut; });
    return true;
}

// retainBest(n_points) over keys[0 .. n) / idx[0 .. n) (permuted together, in place).  Returns the number of survivors m;
// afterwards idx[0 .. m) are the survivors in cv2's order and keys[i] is the key of idx[i].
// rows: shared-memory row scratch with cap >= ceil(n / 1024) * 32 rows.  Must be called by all CVO_THREADS threads of the CTA
// with identical arguments.
template <typename KeyT>
__device__ int cvo_retain_best(KeyT* keys, int* idx, int n, int n_points, CvoShared& sh, const CvoRows& rows) {
    if (n_points < 0 || n <= n_points) return n;
    if (n_points == 0) return 0;
    const int nth = n_points - 1;
    if (threadIdx.x == 0) { sh.first = 0; sh.last = n; sh.depth = 2 * (31 - __clz(n)); sh.heap_done = 0; }
    __syncthreads();
    while (sh.last - sh.first > CVO_TAIL) {                                 // (uniform: written before the last barrier of a step)
        if (!cvo_select_step<true>(keys, idx, nth, sh, rows)) break;
    }
    __syncthreads();
    if (!sh.heap_done) {
        // The remaining ~lg(CVO_TAIL) steps run on shared-memory copies of the range's keys and indices (the index array lives in
        // global memory: a round trip per swapped row would dominate these short steps): by the whole CTA while a step still has a
        // row per warp to offer, by warp 0 alone (no block barriers) below that.
        const int f0 = sh.first, l0 = sh.last;
        KeyT* tk = reinterpret_cast<KeyT*>(rows.tail_keys) - f0;             // indexed with absolute positions
        int* ti = rows.tail_idx - f0;
        for (int i = f0 + threadIdx.x; i < l0; i += CVO_THREADS) { tk[i] = keys[i]; ti[i] = idx[i]; }
        __syncthreads();
        while (sh.last - sh.first > CVO_WARP_TAIL) {
            if (!cvo_select_step<true>(tk, ti, nth, sh, rows)) break;
        }
        __syncthreads();
        if (threadIdx.x < 32 && !sh.heap_done) {
            while (cvo_select_step<false>(tk, ti, nth, sh, rows)) {}
            __syncwarp();
            if (threadIdx.x == 0 && !sh.heap_done) {                       // __insertion_sort on <= 3 elements
                const int first = sh.first, last = sh.last;
                for (int i = first + 1; i < last; ++i) {
                    const KeyT vk = tk[i]; const int vi = ti[i];
                    int j = i;
                    if (vk > tk[first]) {
                        for (; j > first; --j) { tk[j] = tk[j - 1]; ti[j] = ti[j - 1]; }
                    } else {
                        for (; vk > tk[j - 1]; --j) { tk[j] = tk[j - 1]; ti[j] = ti[j - 1]; }
                    }
                    tk[j] = vk; ti[j] = vi;
                }
            }
        }
        __syncthreads();
        for (int i = f0 + threadIdx.x; i < l0; i += CVO_THREADS) { keys[i] = tk[i]; idx[i] = ti[i]; }
    }
    __syncthreads();
    const KeyT amb = keys[nth];
    int m;
    if (n - n_points > CVO_WARP_TAIL) {   // (one pass; its few swaps -- the ties of the boundary response -- go to global memory)
        cvo_pair_pass<true>(keys, idx, n_points, n, [amb](KeyT k) { return !(k >= amb); }, [amb](KeyT k) { return k >= amb; }, sh, rows,
                            -1, amb, [] {}, [](int) {});
        m = n_points + sh.nR;
    } else {
        if (threadIdx.x < 32)
            cvo_pair_pass<false>(keys, idx, n_points, n, [amb](KeyT k) { return !(k >= amb); }, [amb](KeyT k) { return k >= amb; }, sh, rows,
                                 -1, amb, [] {}, [](int) {});
        __syncthreads();
        m = n_points + sh.nR;
    }
    __syncthreads();                                                       // sh may be reused by the caller's next selection
    return m;
}

// rows of shared scratch cvo_retain_best needs for n elements
__host__ __device__ inline int cvo_rows_needed(int n) {
    const int block = ((n + CVO_THREADS - 1) / CVO_THREADS) * 32;          // CTA passes: rows per warp x 32 warps
    const int warp = (CVO_TAIL + 31) / 32 + 32;                            // passes over the shared-memory tail
    return block > warp ? block : warp;
}
