#include <cstdio>
#include <cuda_runtime.h>
__device__ __noinline__ bool warp_solve(double (*M)[10], double* X, double* P, int N, double tiny, long long* tc) {
    const int lane = threadIdx.x & 31;
    bool ok = true;
    long long ta=0,tb=0,tcx=0,td=0;
    for (int c = 0; c < N; ++c) {
        long long t0=clock64();
        double best = (lane >= c && lane < N) ? fabs(M[lane][c]) : -1.0;
        int bi = lane;
        for (int o = 8; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        best = __shfl_sync(0xffffffffu, best, 0); bi = __shfl_sync(0xffffffffu, bi, 0);
        if (!(best > tiny)) ok = false;
        long long t1=clock64(); ta+=t1-t0;
        if (bi != c && lane <= N) { const double t = M[c][lane]; M[c][lane] = M[bi][lane]; M[bi][lane] = t; }
        __syncwarp();
        const double inv = __drcp_rn(M[c][c]);
        if (lane == 0) P[c] = inv;
        long long t2=clock64(); tb+=t2-t1;
        const int k = c + 1 + lane;
        for (int r = c + 1; r < N; ++r) {
            const double f = M[r][c] * inv;
            if (k <= N) M[r][k] -= f * M[c][k];
        }
        __syncwarp();
        tcx+=clock64()-t2;
    }
    long long t3=clock64();
    for (int r = N - 1; r >= 0; --r) {
        double part = (lane > r && lane < N) ? M[r][lane] * X[lane] : 0.0;
        for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == 0) X[r] = (M[r][N] - part) * P[r];
        __syncwarp();
    }
    td=clock64()-t3;
    if (lane==0){tc[0]=ta;tc[1]=tb;tc[2]=tcx;tc[3]=td;}
    return ok;
}
__global__ void k(double* o, long long* c) {
    __shared__ double M[9][10], X[9], P[9];
    int lane = threadIdx.x;
    for (int rep = 0; rep < 3; ++rep) {
        if (lane < 9) { for (int q = 0; q < 9; ++q) { double v = 0; for (int j = 0; j < 8; ++j) v += sin(1.0 + lane * (j + 1)) * sin(1.0 + q * (j + 1)); M[lane][q] = 300.0 * v + (lane == q ? 1e-11 : 0.0); } M[lane][9] = 1.0 + 0.37 * lane; }
        __syncwarp();
        long long t0 = clock64();
        warp_solve(M, X, P, 9, 0.0, c+2);
        long long t1 = clock64();
        if (lane == 0) c[rep==2] = t1 - t0;
    }
    if (lane < 9) o[lane] = X[lane];
}
int main() {
    double* o; long long* c; cudaMalloc(&o, 512); cudaMalloc(&c, 64);
    k<<<1, 32>>>(o, c);
    long long h[6]; cudaMemcpy(h, c, 48, cudaMemcpyDeviceToHost);
    printf("warp_solve(9): first %lld  third %lld cycles; pivot %lld swap+rcp %lld elim %lld backsub %lld\n", h[0], h[1], h[2],h[3],h[4],h[5]);
    return 0;
}
