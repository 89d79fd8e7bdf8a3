// sift.cu -- placeholder until the SIFT kernels land (build order: ORB path first).
#include "sift.cuh"
struct BmSift { int dummy; };
int bm_sift_create(BmSift** out, int, int, int, cudaStream_t) { *out = nullptr; bm_set_error("SIFT detector not built yet"); return -1; }
void bm_sift_destroy(BmSift*) {}
cudaError_t bm_sift_detect(BmSift*, const uint8_t*, BmKeypoints*) { return cudaErrorNotSupported; }
