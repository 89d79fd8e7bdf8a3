// preview.cu -- live preview without full-canvas copies.  The reference hands a copy of the whole canvas to
// its progress callback every frame (/root/reference/main.py:1630-1632) and the GUI turns it into a 400 x 300 thumbnail with
// cv2.cvtColor(BGR2RGB) -> Image.fromarray -> Image.resize((400, 300)) (/root/reference/gui.py:143-158).  Here the thumbnail is made
// on the device and only out_w * out_h * 3 bytes cross PCIe (360 KB instead of 14.9 MB at 1080p, 3.2 GB in config 5).
//
// Image.resize defaults to BICUBIC with an anti-aliasing support scaled by the reduction factor (third-party Pillow, 12.2.0 in this
// image, unpinned in requirements.txt).  Published algorithm (ImagingResample, 8 bits per channel), restated in oracle/preview.py
// and pinned there against live Pillow:
//   scale = in / out, filterscale = max(scale, 1), support = 2 * filterscale, ksize = 2 * ceil(support) + 1
//   per output index: center = (xx + 0.5) * scale, xmin = max(int(center - support + 0.5), 0), xmax = min(int(center + support + 0.5), in)
//   w_j = bicubic((j + xmin - center + 0.5) / filterscale) (a = -0.5), normalised by their double-precision sum,
//   fixed point k_j = int(w_j * 2^22 -+ 0.5); pixel = clamp((2^21 + sum p_j k_j) >> 22, 0, 255)
//   horizontal pass first into an 8-bit intermediate, then the vertical pass.
#include "preview.cuh"
#include <math.h>
#include <vector>

#define PV_PRECISION_BITS 22

static double pv_bicubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

static int pv_axis_taps(int n_in, int n_out) {
    double fs = (double)n_in / n_out;
    if (fs < 1.0) fs = 1.0;
    return (int)ceil(2.0 * fs) * 2 + 1;
}

// bounds[2*n_out] = (first source index, tap count), kk[n_out * ksize] fixed-point coefficients
static void pv_axis_table(int n_in, int n_out, int ksize, int* bounds, int* kk) {
    const double scale = (double)n_in / n_out;
    const double fs = scale < 1.0 ? 1.0 : scale;
    const double support = 2.0 * fs, ss = 1.0 / fs;
    std::vector<double> w(ksize);
    for (int xx = 0; xx < n_out; ++xx) {
        const double center = 0.0 + (xx + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > n_in) xmax = n_in;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) { w[x] = pv_bicubic((x + xmin - center + 0.5) * ss); ww += w[x]; }
        for (int x = 0; x < ksize; ++x) {
            double v = 0.0;
            if (x < xmax) v = ww != 0.0 ? w[x] / ww : w[x];
            kk[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << PV_PRECISION_BITS)) : (int)(0.5 + v * (1 << PV_PRECISION_BITS));
        }
        bounds[2 * xx] = xmin; bounds[2 * xx + 1] = xmax;
    }
}

__device__ __forceinline__ int pv_clip8(int v) { return min(max(v >> PV_PRECISION_BITS, 0), 255); }

// horizontal pass: one thread per (row, output column); lanes along the output columns
__global__ void __launch_bounds__(256) k_preview_rows(const uchar4* __restrict__ canvas, int in_w, int in_h, int out_w, int kx,
                                                      const int* __restrict__ bounds, const int* __restrict__ kk, uchar4* __restrict__ tmp) {
    const int xx = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (xx >= out_w || y >= in_h) return;
    const int x0 = __ldg(bounds + 2 * xx), n = __ldg(bounds + 2 * xx + 1);
    const uchar4* __restrict__ row = canvas + (size_t)y * in_w + x0;
    const int* __restrict__ k = kk + (size_t)xx * kx;
    int s0 = 1 << (PV_PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int j = 0; j < n; ++j) {
        const uchar4 p = __ldg(row + j);
        const int c = __ldg(k + j);
        s0 += p.x * c; s1 += p.y * c; s2 += p.z * c;
    }
    tmp[(size_t)y * out_w + xx] = make_uchar4(pv_clip8(s0), pv_clip8(s1), pv_clip8(s2), 0);
}

// vertical pass + channel order + packing
__global__ void __launch_bounds__(256) k_preview_cols(const uchar4* __restrict__ tmp, int out_w, int out_h, int ky,
                                                      const int* __restrict__ bounds, const int* __restrict__ kk, uint8_t* __restrict__ out, int rgb) {
    const int xx = blockIdx.x * 32 + (threadIdx.x & 31), yy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (xx >= out_w || yy >= out_h) return;
    const int y0 = __ldg(bounds + 2 * yy), n = __ldg(bounds + 2 * yy + 1);
    const uchar4* __restrict__ col = tmp + (size_t)y0 * out_w + xx;
    const int* __restrict__ k = kk + (size_t)yy * ky;
    int s0 = 1 << (PV_PRECISION_BITS - 1), s1 = s0, s2 = s0;
    for (int j = 0; j < n; ++j) {
        const uchar4 p = __ldg(col + (size_t)j * out_w);
        const int c = __ldg(k + j);
        s0 += p.x * c; s1 += p.y * c; s2 += p.z * c;
    }
    uint8_t* q = out + ((size_t)yy * out_w + xx) * 3;
    const uint8_t b = (uint8_t)pv_clip8(s0), g = (uint8_t)pv_clip8(s1), r = (uint8_t)pv_clip8(s2);
    q[0] = rgb ? r : b; q[1] = g; q[2] = rgb ? b : r;
}

cudaError_t bm_preview_prepare(BmPreviewPlan* p, int in_w, int in_h, int out_w, int out_h, cudaStream_t s) {
    if (p->in_w == in_w && p->in_h == in_h && p->out_w == out_w && p->out_h == out_h && p->d_tab) return cudaSuccess;
    const int kx = pv_axis_taps(in_w, out_w), ky = pv_axis_taps(in_h, out_h);
    const size_t n_tab = (size_t)out_w * (2 + kx) + (size_t)out_h * (2 + ky);
    std::vector<int> tab(n_tab);
    int* bx = tab.data();
    int* kxp = bx + 2 * out_w;
    int* byp = kxp + (size_t)out_w * kx;
    int* kyp = byp + 2 * out_h;
    pv_axis_table(in_w, out_w, kx, bx, kxp);
    pv_axis_table(in_h, out_h, ky, byp, kyp);
    cudaError_t e;
    if (p->tab_cap < n_tab) {
        cudaFree(p->d_tab); p->d_tab = nullptr; p->tab_cap = 0;
        if ((e = cudaMalloc(&p->d_tab, n_tab * sizeof(int))) != cudaSuccess) return e;
        p->tab_cap = n_tab;
    }
    const size_t n_tmp = (size_t)in_h * out_w, n_out = (size_t)out_w * out_h * 3;
    if (p->tmp_cap < n_tmp) {
        cudaFree(p->d_tmp); p->d_tmp = nullptr; p->tmp_cap = 0;
        if ((e = cudaMalloc(&p->d_tmp, n_tmp * sizeof(uchar4))) != cudaSuccess) return e;
        p->tmp_cap = n_tmp;
    }
    if (p->out_cap < n_out) {
        cudaFree(p->d_out); p->d_out = nullptr; p->out_cap = 0;
        if ((e = cudaMalloc(&p->d_out, n_out + 16)) != cudaSuccess) return e;
        p->out_cap = n_out;
    }
    if ((e = cudaMemcpyAsync(p->d_tab, tab.data(), n_tab * sizeof(int), cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;          // `tab` is a local
    p->in_w = in_w; p->in_h = in_h; p->out_w = out_w; p->out_h = out_h; p->kx = kx; p->ky = ky;
    return cudaSuccess;
}

cudaError_t bm_launch_preview(const BmPreviewPlan& p, const uchar4* canvas, int rgb, cudaStream_t s) {
    const int* bx = p.d_tab;
    const int* kxp = bx + 2 * p.out_w;
    const int* byp = kxp + (size_t)p.out_w * p.kx;
    const int* kyp = byp + 2 * p.out_h;
    BM_COUNT_LAUNCHES(1), k_preview_rows<<<dim3(bm_div_up(p.out_w, 32), bm_div_up(p.in_h, 8)), 256, 0, s>>>(canvas, p.in_w, p.in_h, p.out_w, p.kx, bx, kxp, p.d_tmp);
    BM_COUNT_LAUNCHES(1), k_preview_cols<<<dim3(bm_div_up(p.out_w, 32), bm_div_up(p.out_h, 8)), 256, 0, s>>>(p.d_tmp, p.out_w, p.out_h, p.ky, byp, kyp, p.d_out, rgb);
    return cudaGetLastError();
}

void bm_preview_free(BmPreviewPlan* p) {
    cudaFree(p->d_tab); cudaFree(p->d_tmp); cudaFree(p->d_out);
    *p = BmPreviewPlan();
}
