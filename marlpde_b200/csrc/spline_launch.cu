// Tensor-product B-spline evaluation on the device: the ground-truth field S(x, t) that
// setGroundTruth / getMseReward (/root/reference/python/_model/Burger.py:322-327, 578-589) obtain from
// scipy.interpolate.interp2d, i.e. FITPACK's bispev on the knots / coefficients of the interpolating spline.
// The spline is still FITTED on the host (SciPy RectBivariateSpline, the library the reference calls); this kernel
// samples it for every (shifted grid, time row) pair of a batch, which is the part that scales with the number of
// environments: out[q, i, j] = S(xq[q, j], tq[i]).
// Restates FITPACK fpbisp (interval search with clamping to [t_k, t_{n-k-1}]) and fpbspl (de Boor recurrence) in fp64.
#include <cuda_runtime.h>
#include <cstdint>
#include <string>

#include "../../include/marlpde_b200.h"

namespace {

// index l with t[l] <= x < t[l+1], clamped to k .. n-k-2 (fpbisp: the last interval takes x == t_end)
__device__ __forceinline__ int find_span(const double* __restrict__ t, int n, int k, double x) {
    int lo = k, hi = n - k - 1;              // t[lo] <= x <= t[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (x >= t[mid]) lo = mid; else hi = mid;
    }
    return lo;
}

// fpbspl: the k+1 non-zero B-splines of degree k at x in the knot interval l
__device__ __forceinline__ void bspl(const double* __restrict__ t, int k, double x, int l, double (&h)[4]) {
    double hh[4];
    h[0] = 1.0;
    for (int j = 1; j <= k; ++j) {
        for (int i = 0; i < j; ++i) hh[i] = h[i];
        h[0] = 0.0;
        for (int i = 1; i <= j; ++i) {
            const int li = l + i, lj = li - j;
            if (t[li] == t[lj]) { h[i] = 0.0; continue; }
            const double f = hh[i - 1] / (t[li] - t[lj]);
            h[i - 1] = h[i - 1] + f * (t[li] - x);
            h[i] = f * (x - t[lj]);
        }
    }
}

template <typename T>
__global__ void spline_table_kernel(const double* __restrict__ tx, int ntx, const double* __restrict__ ty, int nty,
                                    const double* __restrict__ c, int kx, int ky, const double* __restrict__ xq, int64_t nq, int N,
                                    const double* __restrict__ tq, int64_t rows, T* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nq * rows * N) return;
    const int j = (int)(idx % N);
    const int64_t i = (idx / N) % rows, q = idx / ((int64_t)N * rows);
    double x = xq[q * N + j], y = tq[i];
    x = fmin(fmax(x, tx[kx]), tx[ntx - kx - 1]);          // fpbisp clamps the argument to the spline's domain
    y = fmin(fmax(y, ty[ky]), ty[nty - ky - 1]);
    const int lx = find_span(tx, ntx, kx, x), ly = find_span(ty, nty, ky, y);
    double hx[4], hy[4];
    bspl(tx, kx, x, lx, hx);
    bspl(ty, ky, y, ly, hy);
    const int ncy = nty - ky - 1;
    double sp = 0.0;
    for (int a = 0; a <= kx; ++a) {
        const double* row = c + (int64_t)(lx - kx + a) * ncy + (ly - ky);
        for (int b = 0; b <= ky; ++b) sp += row[b] * hx[a] * hy[b];       // fpbisp: sp = sp + c(l2) * h(i1) * wy(j, j1)
    }
    out[idx] = (T)sp;
}

thread_local std::string g_serr;
}  // namespace

extern "C" {

int mpde_eval_spline_table(const double* tx_dev, int32_t ntx, const double* ty_dev, int32_t nty, const double* c_dev, int32_t kx,
                           int32_t ky, const double* xq_dev, int64_t nq, int32_t N, const double* tq_dev, int64_t rows, void* out_dev,
                           int32_t dtype, void* stream) {
    if (!tx_dev || !ty_dev || !c_dev || !xq_dev || !tq_dev || !out_dev) return -1;
    if (kx < 1 || kx > 3 || ky < 1 || ky > 3 || ntx < 2 * (kx + 1) || nty < 2 * (ky + 1) || nq < 1 || rows < 1 || N < 1) return -1;
    const int64_t n = nq * rows * N;
    const unsigned grid = (unsigned)((n + 255) / 256);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == MPDE_F64)
        spline_table_kernel<double><<<grid, 256, 0, st>>>(tx_dev, ntx, ty_dev, nty, c_dev, kx, ky, xq_dev, nq, N, tq_dev, rows,
                                                          static_cast<double*>(out_dev));
    else if (dtype == MPDE_F32)
        spline_table_kernel<float><<<grid, 256, 0, st>>>(tx_dev, ntx, ty_dev, nty, c_dev, kx, ky, xq_dev, nq, N, tq_dev, rows,
                                                         static_cast<float*>(out_dev));
    else
        return -1;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // extern "C"
