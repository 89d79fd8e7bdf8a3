// oracle/csrc/stl_order.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
//
// cv2's keypoint ORDER is whatever libstdc++'s std::nth_element + std::partition leave behind inside
// cv::KeyPointsFilter::retainBest (OpenCV 4.x features2d/src/keypoint.cpp; called by ORB twice per pyramid level and by
// SIFT once -- the reference reaches it through `self.detector.detectAndCompute`, /root/reference/main.py:112,718).
// This file does not restate that algorithm: it calls the REAL std::nth_element / std::partition / std::sort of the
// libstdc++ in this image, so that oracle/cvorder.py (the restatement) and the CUDA emulation can be checked against the
// genuine article on arbitrary inputs, including ones that drive introselect into its heap-select fallback.
//
//   g++ -O2 -shared -fPIC oracle/csrc/stl_order.cpp -o oracle/_ref/libstlorder.so      (done by oracle/build_ref.py)
#include <algorithm>
#include <cstdint>
#include <vector>

namespace {
struct Item { float response; int index; };
struct ResponseGreater { bool operator()(const Item& a, const Item& b) const { return a.response > b.response; } };
struct ResponseGE {
    float v;
    bool operator()(const Item& a) const { return a.response >= v; }
};
}  // namespace

extern "C" {

// KeyPointsFilter::retainBest(keypoints, n_points) on `n` items whose responses are `resp` (input order = index order).
// Writes the surviving input indices, in the order retainBest leaves them, to out_idx (capacity n); returns how many.
int stl_retain_best(const float* resp, int n, int n_points, int* out_idx) {
    std::vector<Item> v(n);
    for (int i = 0; i < n; ++i) v[i] = Item{resp[i], i};
    if (n_points >= 0 && v.size() > (size_t)n_points) {
        if (n_points == 0) return 0;
        std::nth_element(v.begin(), v.begin() + n_points - 1, v.end(), ResponseGreater());
        const float ambiguous = v[n_points - 1].response;
        auto new_end = std::partition(v.begin() + n_points, v.end(), ResponseGE{ambiguous});
        v.resize(new_end - v.begin());
    }
    for (size_t i = 0; i < v.size(); ++i) out_idx[i] = v[i].index;
    return (int)v.size();
}

// the bare std::nth_element permutation (for the heap-select fallback tests): out_idx = input indices after the call
void stl_nth_element(const float* resp, int n, int nth, int* out_idx) {
    std::vector<Item> v(n);
    for (int i = 0; i < n; ++i) v[i] = Item{resp[i], i};
    std::nth_element(v.begin(), v.begin() + nth, v.end(), ResponseGreater());
    for (int i = 0; i < n; ++i) out_idx[i] = v[i].index;
}

// KeyPointsFilter::removeDuplicatedSorted's comparator (KeyPoint12_LessThan) + std::sort, as SIFT applies it before
// retainBest.  kp = n rows of (x, y, size, angle, response, octave, class_id) float; out_idx = sorted input indices
// with the duplicates (equal x, y, size, angle) dropped; returns the count.
int stl_sift_sort_unique(const float* kp, int n, int* out_idx) {
    std::vector<int> id(n);
    for (int i = 0; i < n; ++i) id[i] = i;
    auto less = [kp](int a, int b) {
        const float *p = kp + 7 * a, *q = kp + 7 * b;
        if (p[0] != q[0]) return p[0] < q[0];
        if (p[1] != q[1]) return p[1] < q[1];
        if (p[2] != q[2]) return p[2] > q[2];
        if (p[3] != q[3]) return p[3] < q[3];
        if (p[4] != q[4]) return p[4] > q[4];
        if (p[5] != q[5]) return p[5] > q[5];
        return p[6] > q[6];
    };
    std::sort(id.begin(), id.end(), less);
    if (n < 2) { for (int i = 0; i < n; ++i) out_idx[i] = id[i]; return n; }
    int i = 0;
    for (int j = 1; j < n; ++j) {
        const float *p = kp + 7 * id[i], *q = kp + 7 * id[j];
        if (p[0] != q[0] || p[1] != q[1] || p[2] != q[2] || p[3] != q[3]) id[++i] = id[j];
    }
    for (int k = 0; k <= i; ++k) out_idx[k] = id[k];
    return i + 1;
}

}  // extern "C"
