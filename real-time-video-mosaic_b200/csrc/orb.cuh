// orb.cuh -- ORB (FAST-9 + Harris + IC angle + rBRIEF-256) detector object.
#pragma once
#include "common.cuh"

#define BM_ORB_LEVELS 8
#define BM_KP_CAP 8192          // final keypoints per frame (700 + ties); exceeding it is reported as an error

struct BmOrbLevel {
    int w, h;           // level size
    int off;            // pixel offset of the level in pyr / blur / score
    int cand_off;       // offset of the level's segment in the candidate arrays
    int cand_cap;
    int quota;          // features to keep on this level (cv2: nfeaturesPerLevel)
    int wpr;            // words per row of the level's NMS bitmap
    int bits_off;       // word offset of the level in nmsbits
    int row_off;        // offset of the level's rows in rowcnt
    float scale;        // layerScale (float32 pow(1.2f, l))
    float inv_scale;    // 1.f / scale
};
struct BmOrbLevels { BmOrbLevel l[BM_ORB_LEVELS]; int total_px; int total_cand; int total_rows; int total_words; };

// final keypoints, SoA on the device; order: cv2's (level-major; inside a level whatever KeyPointsFilter::retainBest leaves, cvorder.cuh)
struct BmKeypoints {
    float2* pt;         // image coordinates (level coords * scale)
    float* size;
    float* angle;       // degrees
    float* response;
    int* octave;
    int2* lxy;          // level coordinates (ORB) / unused (SIFT)
    uint8_t* desc;      // ORB: 32 B per keypoint; SIFT: 128 floats per keypoint (512 B)
    int* count;         // device scalar
    int* flags;         // device, [0] != 0: a candidate / keypoint list overflowed while these features were made (result unusable)
};

struct BmOrbGraph { const uint8_t* gray; const void* out_pt; cudaGraphExec_t exec; int launches; };

struct BmOrb {
    int w, h, nfeatures;
    BmOrbLevels lv;
    uint8_t *pyr, *blur, *score;
    unsigned* corners;  // FAST corners before NMS, x | y << 16; level l owns [off_l / 2, off_l / 2 + w_l * h_l / 2)
    // one allocation, cleared by one memset per frame: ctr | rowcnt | nmsbits
    int* ctr;           // device counters: [0..7] NMS survivors n1, [24..31] kept, [32] overflow flag, [40..47] FAST corners
    int* rowcnt;        // NMS survivors per (level, row)
    unsigned* nmsbits;  // NMS survivors, one bit per pixel (row-major FAST output order = cv2's, before retainBest)
    size_t zero_bytes;
    uint8_t* ckey;      // [total_cand] FAST scores of the survivors in row-major order (permuted in place by the selection)
    unsigned* cxy;      // [total_cand] x | y << 16 of the survivors, row-major
    int *idx, *idx2;    // [total_cand] permutations of the two retainBest stages
    float* resp2;       // [total_cand] Harris responses when they do not fit in shared memory
    uint2* cand2;       // [total_cand] per level: the final keypoints (xy, response bits) in cv2's order
    cudaStream_t stream;
    BmOrbGraph graphs[BM_DET_MAX_GRAPHS];   // captured detect sequences, one per (input buffer, output buffer): 5 frame slots x 5 keypoint slots
    int ngraphs, graphs_disabled;
};

int bm_orb_create(BmOrb** out, int h, int w, int nfeatures, cudaStream_t s);
void bm_orb_destroy(BmOrb* o);
int bm_kp_alloc(BmKeypoints* k, int desc_bytes);
void bm_kp_free(BmKeypoints* k);
// gray (device, tightly packed h*w) -> keypoints + descriptors into `out`
// launch == false: only capture + instantiate the graph of this (input, output) pair if it is not cached yet
cudaError_t bm_orb_detect(BmOrb* o, const uint8_t* d_gray, BmKeypoints* out, bool launch = true);
