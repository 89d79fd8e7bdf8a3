// match_tc.cu -- SIFT brute-force kNN (k = 2) on the 5th-generation tensor cores (reference: BFMatcher().knnMatch at
// /root/reference/main.py:687-688).
//
// SIFT descriptors are integers 0..255, so  |a-b|^2 = |a|^2 + |b|^2 - 2 a.b  with every term an exact integer
// below 2^24: bf16 holds 0..255 exactly, the products are exact and the fp32 accumulation in TMEM is exact in any order.  The
// distances are therefore bit-identical to OpenCV's float loop, and the "fp32 re-rank" only has to implement the tie rule
// (distance, trainIdx) -- which the epilogue does by scanning the train index in increasing order with strict comparisons.
//
// One CTA = 128 queries (the M = 128 rows of a tcgen05.mma.cta_group::1 tile = the 128 TMEM lanes) x every TC_SPLIT-th tile of
// N = 256 train descriptors (blockIdx.y; ~700 descriptors = 3 tiles run on 3 CTAs side by side); K = 128.  A second tiny kernel
// merges the per-split (best, second) pairs lexicographically by (distance, trainIdx).  Operands are converted u8 -> bf16 while they are staged into shared memory in the
// canonical K-major SWIZZLE_128B layout (rows of 64 bf16 = 128 B, 16-byte chunk index XOR (row & 7), 8-row groups 1024 B apart),
// two 64-element K blocks per operand.  One elected thread issues 8 MMAs (K = 16 each) per train tile into a 256-column fp32
// accumulator in TMEM and commits to an mbarrier; the four warps then read their 32 lanes with tcgen05.ld (32x32b.x32) and
// keep the two nearest train descriptors of their query in registers.
#include "match.cuh"
#include <stdint.h>

#define TC_M 128
#define TC_N 256
#define TC_TMEM_COLS 256
#define TC_SPLIT BM_L2_SPLIT                   // train tiles are dealt round-robin to TC_SPLIT CTAs per query tile

struct TcSmem {
    uint8_t A[2][TC_M * 128];        // [k block][row * 128 B]   32 KB
    uint8_t B[2][TC_N * 128];        //                           64 KB
    unsigned nb[TC_N];               // |b|^2 of the train tile
    unsigned long long mbar;
    unsigned tmem_base;
};

__device__ __forceinline__ unsigned tc_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits [0,14),
// leading byte offset (unused for swizzled K-major, 1) in [16,30), stride byte offset (1024 B between 8-row groups) >> 4 in
// [32,46), descriptor version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64)
__device__ __forceinline__ unsigned long long tc_desc(unsigned smem_addr) {
    return (unsigned long long)((smem_addr >> 4) & 0x3fffu) | (1ull << 16) | ((unsigned long long)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// 8 u8 -> 8 bf16 (exact: an integer below 256 has at most 8 significant bits, so its bf16 is the top half of its fp32)
__device__ __forceinline__ uint4 tc_u8x8_to_bf16(unsigned lo, unsigned hi) {
    auto two = [](unsigned a, unsigned b) { return (__float_as_uint((float)a) >> 16) | (__float_as_uint((float)b) & 0xffff0000u); };
    return make_uint4(two(lo & 0xff, (lo >> 8) & 0xff), two((lo >> 16) & 0xff, lo >> 24), two(hi & 0xff, (hi >> 8) & 0xff), two((hi >> 16) & 0xff, hi >> 24));
}

// stage one descriptor (128 u8, global) as row `r` of an operand tile; returns |d|^2.  Rows beyond n are zero.
__device__ __forceinline__ unsigned tc_stage_row(const uint8_t* __restrict__ desc, int idx, int n, uint8_t* k0, uint8_t* k1, int r) {
    unsigned nrm = 0;
    const uint4* p = reinterpret_cast<const uint4*>(desc + (size_t)idx * 128);
#pragma unroll
    for (int i = 0; i < 8; ++i) {                       // 16 source bytes = 2 chunks of 8 elements
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (idx < n) v = __ldg(p + i);
        nrm = __dp4a(v.x, v.x, nrm); nrm = __dp4a(v.y, v.y, nrm); nrm = __dp4a(v.z, v.z, nrm); nrm = __dp4a(v.w, v.w, nrm);
        uint8_t* kb = (i < 4) ? k0 : k1;                // elements 0..63 -> K block 0, 64..127 -> K block 1
        const int c0 = (2 * i) & 7, c1 = (2 * i + 1) & 7;
        *reinterpret_cast<uint4*>(kb + r * 128 + ((c0 ^ (r & 7)) << 4)) = tc_u8x8_to_bf16(v.x, v.y);
        *reinterpret_cast<uint4*>(kb + r * 128 + ((c1 ^ (r & 7)) << 4)) = tc_u8x8_to_bf16(v.z, v.w);
    }
    return nrm;
}

__global__ void __launch_bounds__(128, 1) k_l2_knn2_tc(const uint8_t* __restrict__ A, const int* __restrict__ nAp, const uint8_t* __restrict__ B,
                                                       const int* __restrict__ nBp, int4* __restrict__ part) {
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    // SWIZZLE_128B operands need 1024-byte aligned tiles: align by hand (1 KB of slack is allocated)
    TcSmem& sm = *reinterpret_cast<TcSmem*>(tc_smem_raw + ((1024u - (tc_smem_u32(tc_smem_raw) & 1023u)) & 1023u));
    const int nA = *nAp, nB = *nBp;
    const int q0 = blockIdx.x * TC_M;
    if (q0 >= nA || (int)blockIdx.y * TC_N >= nB) return;
    const int tid = threadIdx.x, warp = tid >> 5;
    const unsigned mbar = tc_smem_u32(&sm.mbar);

    if (warp == 0) {                                     // one warp owns the TMEM allocation
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&sm.tmem_base)), "r"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const unsigned na = tc_stage_row(A, q0 + tid, nA, sm.A[0], sm.A[1], tid);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = sm.tmem_base;

    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = BF16 (1 << 7, 1 << 10), both K-major,
    // N >> 3 in [17,23), M >> 4 in [24,29)
    const unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(TC_N >> 3) << 17) | ((unsigned)(TC_M >> 4) << 24);
    int b1 = 0x7fffffff, i1 = -1, b2 = 0x7fffffff, i2 = -1;
    unsigned phase = 0;
    for (int t0 = blockIdx.y * TC_N; t0 < nB; t0 += TC_SPLIT * TC_N) {
        // stage the train tile (2 rows per thread)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = tid + 128 * h;
            sm.nb[r] = tc_stage_row(B, t0 + r, nB, sm.B[0], sm.B[1], r);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int k = 0; k < 8; ++k) {                // K = 16 per MMA: 32 B steps inside the 128 B swizzle row, 2 K blocks
                const unsigned long long da = tc_desc(tc_smem_u32(sm.A[k >> 2]) + (k & 3) * 32);
                const unsigned long long db = tc_desc(tc_smem_u32(sm.B[k >> 2]) + (k & 3) * 32);
                const unsigned acc = k > 0 ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
            // commit: the mbarrier is arrived on when all MMAs above have completed (implies fence::before_thread_sync)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
        }
        {
            unsigned done;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(mbar), "r"(phase) : "memory");
            } while (!done);
            phase ^= 1u;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // epilogue: this thread's query is TMEM lane (warp * 32 + lane); 8 chunks of 32 accumulator columns
#pragma unroll 1
        for (int c = 0; c < TC_N / 32; ++c) {
            unsigned v[32];
            const unsigned taddr = tmem + ((unsigned)(warp * 32) << 16) + (unsigned)(c * 32);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                         "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int t = t0 + c * 32 + j;
                if (t >= nB) break;
                const int d = (int)na + (int)sm.nb[c * 32 + j] - 2 * __float2int_rn(__uint_as_float(v[j]));
                if (d < b1) { b2 = b1; i2 = i1; b1 = d; i1 = t; }
                else if (d < b2) { b2 = d; i2 = t; }
            }
        }
        // all TMEM reads and shared-memory reads of this tile are done before the next tile overwrites them
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }
    if (q0 + tid < nA) part[(size_t)blockIdx.y * BM_KP_CAP + q0 + tid] = make_int4(b1, i1, b2, i2);
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS));
}

// merge the (best, second) pairs of the splits: candidates ordered by (squared distance, train index)
__global__ void __launch_bounds__(256) k_l2_knn2_merge(const int4* __restrict__ part, const int* __restrict__ nAp, const int* __restrict__ nBp,
                                                       int* __restrict__ nn1, float* __restrict__ d1o, int* __restrict__ nn2, float* __restrict__ d2o) {
    BM_PDL_WAIT();
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int nA = *nAp, nB = *nBp;
    if (q >= nA) return;
    int b1 = 0x7fffffff, i1 = 0x7fffffff, b2 = 0x7fffffff, i2 = 0x7fffffff;
    auto push = [&](int d, int i) {
        if (i < 0) return;
        if (d < b1 || (d == b1 && i < i1)) { b2 = b1; i2 = i1; b1 = d; i1 = i; }
        else if (d < b2 || (d == b2 && i < i2)) { b2 = d; i2 = i; }
    };
    for (int s = 0; s < TC_SPLIT && s * TC_N < nB; ++s) {
        const int4 p = part[(size_t)s * BM_KP_CAP + q];
        push(p.x, p.y); push(p.z, p.w);
    }
    nn1[q] = i1 != 0x7fffffff ? i1 : -1; nn2[q] = i2 != 0x7fffffff ? i2 : -1;
    d1o[q] = __fsqrt_rn((float)b1); d2o[q] = __fsqrt_rn((float)b2);
}

cudaError_t bm_launch_l2_knn2_tc(const uint8_t* A, const int* nA, const uint8_t* B, const int* nB, int4* part, int* nn1, float* d1, int* nn2,
                                 float* d2, cudaStream_t s) {
    cudaError_t attr;
    BM_SMEM_OPTIN(k_l2_knn2_tc, sizeof(TcSmem) + 1024, attr);
    if (attr != cudaSuccess) return attr;
    BM_COUNT_LAUNCHES(1), k_l2_knn2_tc<<<dim3(BM_KP_CAP / TC_M, TC_SPLIT), 128, sizeof(TcSmem) + 1024, s>>>(A, nA, B, nB, part);
    BM_COUNT_LAUNCHES(1);
    return bm_launch_pdl(k_l2_knn2_merge, dim3(BM_KP_CAP / 256), dim3(256), 0, s, (const int4*)part, nA, nB, nn1, d1, nn2, d2);
}
