// Throughput probe: scalar FFMA (constant-bank / register multiplier) against packed FFMA2 on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>
__constant__ float c_k[32];
__constant__ float2 c_k2[32];
#define ITERS 4096
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { return ((u64)__float_as_uint(b) << 32) | __float_as_uint(a); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float seed) {
    float v = seed + threadIdx.x;
    if (MODE == 0) {          // FFMA acc = c[k] * v + acc  (constant-bank multiplier)
        float a[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = v + i;
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = __fmaf_rn(c_k[i], v, a[i]);
            v += 1.0f;
        }
        float s = 0; for (int i = 0; i < 16; ++i) s += a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (MODE == 1) {   // FFMA with three register operands
        float a[16], kk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { a[i] = v + i; kk[i] = seed * i; }
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = __fmaf_rn(kk[i], v, a[i]);
            v += 1.0f;
        }
        float s = 0; for (int i = 0; i < 16; ++i) s += a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (MODE == 2) {   // FFMA2, three register pairs
        u64 a[16], kk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { a[i] = pk(v + i, v - i); kk[i] = pk(seed * i, seed + i); }
        u64 vv = pk(v, v + 0.5f);
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fma2(kk[i], vv, a[i]);
            vv += 0x0000000100000001ull;
        }
        u64 s = 0; for (int i = 0; i < 16; ++i) s ^= a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)s ^ (unsigned)(s >> 32));
    } else {                  // FFMA2 with the multiplier pair read from the constant bank
        u64 a[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = pk(v + i, v - i);
        u64 vv = pk(v, v + 0.5f);
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fma2(pk(c_k2[i].x, c_k2[i].y), vv, a[i]);
            vv += 0x0000000100000001ull;
        }
        u64 s = 0; for (int i = 0; i < 16; ++i) s ^= a[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)s ^ (unsigned)(s >> 32));
    }
}
template <int MODE> void run(const char* name, float* d, int flops_per) {
    const int blocks = 148 * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, 1.0f); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 10; ++r) k<MODE><<<blocks, 256>>>(d, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = 10.0 * blocks * 256 * (double)ITERS * 16;
    printf("%-28s %8.3f ms  %7.2f G thread-instr/s  %7.2f TFLOP/s  (%s)\n", name, ms, inst / ms / 1e6, inst * flops_per / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    float h[32]; for (int i = 0; i < 32; ++i) h[i] = 1e-3f * i; cudaMemcpyToSymbol(c_k, h, sizeof h);
    float2 h2[32]; for (int i = 0; i < 32; ++i) h2[i] = make_float2(1e-3f * i, 2e-3f * i); cudaMemcpyToSymbol(c_k2, h2, sizeof h2);
    run<0>("FFMA  const multiplier", d, 2);
    run<1>("FFMA  3 registers", d, 2);
    run<2>("FFMA2 3 register pairs", d, 4);
    run<3>("FFMA2 const pair", d, 4);
    return 0;
}
