#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* o, float* of, long long* c, double a, double b, float af, float bf) {
    double x = a; float xf = af;
    long long t0 = clock64();
    #pragma unroll 1
    for (int i = 0; i < 1000; ++i) x = x * b + a;
    long long t1 = clock64();
    #pragma unroll 1
    for (int i = 0; i < 1000; ++i) xf = xf * bf + af;
    long long t2 = clock64();
    double y = a;
    #pragma unroll 1
    for (int i = 0; i < 100; ++i) y = 1.0 / (y + b);
    long long t3 = clock64();
    double z = a;
    #pragma unroll 1
    for (int i = 0; i < 100; ++i) z = __drcp_rn(z + b);
    long long t4 = clock64();
    __shared__ double sm[64];
    sm[threadIdx.x] = a; __syncwarp();
    double w = 0;
    #pragma unroll 1
    for (int i = 0; i < 1000; ++i) { w += sm[(threadIdx.x + i) & 31]; }
    long long t5 = clock64();
    double v = a;
    #pragma unroll 1
    for (int i = 0; i < 1000; ++i) v = __shfl_xor_sync(0xffffffffu, v, 1) + b;
    long long t6 = clock64();
    o[threadIdx.x] = x + y + z + w + v; of[threadIdx.x] = xf;
    if (threadIdx.x == 0) { c[0] = t1 - t0; c[1] = t2 - t1; c[2] = t3 - t2; c[3] = t4 - t3; c[4] = t5 - t4; c[5] = t6 - t5; }
}
int main() {
    double* o; float* of; long long* c; cudaMalloc(&o, 512); cudaMalloc(&of, 512); cudaMalloc(&c, 64);
    for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(o, of, c, 1.0000001, 0.9999999, 1.0001f, 0.9999f);
    long long h[6]; cudaMemcpy(h, c, 48, cudaMemcpyDeviceToHost);
    printf("per-op cycles: dfma %.1f ffma %.1f ddiv %.1f drcp %.1f lds+dadd %.1f shfl_d+dadd %.1f\n", h[0] / 1000.0, h[1] / 1000.0, h[2] / 100.0, h[3] / 100.0, h[4] / 1000.0, h[5] / 1000.0);
    return 0;
}
