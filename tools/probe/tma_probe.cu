// minimal TMA 2-D box load probe: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>

__device__ __forceinline__ unsigned su32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int ROWS>
__global__ void k_probe(const __grid_constant__ CUtensorMap tm, float* out, int x0, int y0, int swz) {
    extern __shared__ unsigned char raw[];
    __shared__ __align__(8) unsigned long long mbar;
    unsigned char* tile = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    if (threadIdx.x == 0) {
        const unsigned mb = su32(&mbar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((unsigned)(ROWS * 128)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(su32(tile)), "l"(&tm), "r"(x0), "r"(y0), "r"(mb) : "memory");
    }
    __syncthreads();
    {
        const unsigned mb = su32(&mbar);
        unsigned done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(mb), "r"(0u) : "memory");
    }
    for (int i = threadIdx.x; i < ROWS * 32; i += blockDim.x) {
        const int r = i / 32, c = i % 32;
        const int q = c >> 2;
        const unsigned off = r * 128 + ((swz ? (q ^ (r & 7)) : q) << 4) + ((c & 3) << 2);
        out[i] = *reinterpret_cast<const float*>(tile + off);
    }
}

template <int ROWS>
int run(const char* name, CUtensorMapSwizzle sw, float* d_img, int w, int h, const std::vector<float>& img) {
    CUtensorMap tm;
    const cuuint64_t gdim[2] = {(cuuint64_t)w, (cuuint64_t)h};
    const cuuint64_t gstr[1] = {(cuuint64_t)w * 4};
    const cuuint32_t box[2] = {32u, (cuuint32_t)ROWS};
    const cuuint32_t es[2] = {1u, 1u};
    CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_img, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("%s: encode -> %d\n", name, (int)r);
    if (r != CUDA_SUCCESS) return 1;
    float* d_out; cudaMalloc(&d_out, ROWS * 32 * 4);
    cudaFuncSetAttribute(k_probe<ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, ROWS * 128 + 1024);
    const int x0 = 40, y0 = 8;
    k_probe<ROWS><<<1, 256, ROWS * 128 + 1024>>>(tm, d_out, x0, y0, sw == CU_TENSOR_MAP_SWIZZLE_128B);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s: kernel -> %s\n", name, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> o(ROWS * 32);
    cudaMemcpy(o.data(), d_out, o.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int rr = 0; rr < ROWS; ++rr) for (int c = 0; c < 32; ++c) if (o[rr * 32 + c] != img[(size_t)(y0 + rr) * w + x0 + c]) ++bad;
    printf("%s: mismatches %d of %d\n", name, bad, ROWS * 32);
    return bad != 0;
}

int main() {
    cudaFree(0);
    const int w = 1280, h = 720;
    std::vector<float> img((size_t)w * h);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (float)(i % 100003);
    float* d; cudaMalloc(&d, img.size() * 4); cudaMemcpy(d, img.data(), img.size() * 4, cudaMemcpyHostToDevice);
    int rc = 0;
    rc |= run<64>("none/64", CU_TENSOR_MAP_SWIZZLE_NONE, d, w, h, img);
    rc |= run<64>("swz128/64", CU_TENSOR_MAP_SWIZZLE_128B, d, w, h, img);
    rc |= run<128>("swz128/128", CU_TENSOR_MAP_SWIZZLE_128B, d, w, h, img);
    return rc;
}
