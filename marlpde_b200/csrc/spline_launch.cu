// Tensor-product B-spline evaluation on the device: the ground-truth field S(x, t) that
// setGroundTruth / getMseReward (/root/reference/python/_model/Burger.py:322-327, 578-589) obtain from
// scipy.interpolate.interp2d, i.e. FITPACK's bispev on the knots / coefficients of the interpolating spline.
// mpde_eval_spline_table samples the spline for every (shifted grid, time row) pair of a batch -- the part that scales with
// the number of environments: out[q, i, j] = S(xq[q, j], tq[i]); mpde_fit_spline (round 2) computes knots and coefficients
// of the interpolating spline on the device as well (FITPACK regrid with s = 0), so setGroundTruth needs no host fit.
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

// ---- interpolating spline FIT: what scipy.interpolate.interp2d / RectBivariateSpline(s = 0) compute on the host --
// FITPACK regrid / fpregr with s = 0: knots t = [x_0 (k+1 times), x_{k/2+1} .. x_{m-2-k/2}, x_{m-1} (k+1 times)] (odd degree k),
// coefficients = the solution of the collocation system  sum_ab B_a(x_i) c_ab B_b(t_j) = z_ji.  FITPACK solves it by Givens
// rotations of the banded observation matrix; the matrix is totally positive, so plain banded LU without pivoting is stable and
// gives the same coefficients to rounding.  Band layout: ab[i * 7 + (j - i + 3)], |j - i| <= 3.
constexpr int BW = 7;

__global__ void spline_knots_lu_kernel(const double* __restrict__ x, int mx, const double* __restrict__ y, int my, int k,
                                       double* __restrict__ tx, double* __restrict__ ty, double* __restrict__ abx,
                                       double* __restrict__ aby) {
    const double* g = blockIdx.x == 0 ? x : y;
    const int m = blockIdx.x == 0 ? mx : my;
    double* t = blockIdx.x == 0 ? tx : ty;
    double* ab = blockIdx.x == 0 ? abx : aby;
    const int n = m + k + 1, k3 = k / 2;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double v;
        if (i <= k) v = g[0];
        else if (i >= m) v = g[m - 1];
        else v = g[i - k - 1 + k3 + 1];                      // fpregr: tx(kx+2 ..) = x(kx/2+2 ..) (1-based)
        t[i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x) {     // collocation row i: the k+1 non-zero B-splines at g[i]
        for (int d = 0; d < BW; ++d) ab[i * BW + d] = 0.0;
        const double xv = g[i];
        const int l = find_span(t, n, k, xv);
        double h[4];
        bspl(t, k, xv, l, h);
        for (int a = 0; a <= k; ++a) {
            const int j = l - k + a;
            if (j - i + 3 >= 0 && j - i + 3 < BW) ab[i * BW + (j - i + 3)] = h[a];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {                                 // banded LU, no pivoting (multipliers stored below the diagonal)
        for (int i = 0; i < m; ++i) {
            const double piv = ab[i * BW + 3];
            for (int r = i + 1; r <= i + 3 && r < m; ++r) {
                const double mlt = ab[r * BW + (i - r + 3)] / piv;
                ab[r * BW + (i - r + 3)] = mlt;
                if (mlt != 0.0)
                    for (int j = i + 1; j <= i + 3 && j < m; ++j) ab[r * BW + (j - r + 3)] -= mlt * ab[i * BW + (j - i + 3)];
            }
        }
    }
}

// in-place solve of the factored band system for one right-hand side with element stride `st`
__device__ __forceinline__ void band_solve(const double* __restrict__ ab, int m, double* __restrict__ rhs, int64_t st) {
    for (int i = 1; i < m; ++i) {
        double acc = rhs[i * st];
        for (int j = (i - 3 > 0 ? i - 3 : 0); j < i; ++j) acc -= ab[i * BW + (j - i + 3)] * rhs[j * st];
        rhs[i * st] = acc;
    }
    for (int i = m - 1; i >= 0; --i) {
        double acc = rhs[i * st];
        for (int j = i + 1; j <= i + 3 && j < m; ++j) acc -= ab[i * BW + (j - i + 3)] * rhs[j * st];
        rhs[i * st] = acc / ab[i * BW + 3];
    }
}
// along x: one thread per time row; z [my][mx] -> w [my][mx]
__global__ void spline_solve_x_kernel(const double* __restrict__ abx, int mx, int my, const double* __restrict__ z, double* __restrict__ w) {
    const int it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= my) return;
    for (int i = 0; i < mx; ++i) w[(int64_t)it * mx + i] = z[(int64_t)it * mx + i];
    band_solve(abx, mx, w + (int64_t)it * mx, 1);
}
// along t: one thread per x coefficient (coalesced across threads); w [my][mx] in place, then c[ix * my + it]
__global__ void spline_solve_y_kernel(const double* __restrict__ aby, int mx, int my, double* __restrict__ w, double* __restrict__ c) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    if (ix >= mx) return;
    band_solve(aby, my, w + ix, mx);
    for (int it = 0; it < my; ++it) c[(int64_t)ix * my + it] = w[(int64_t)it * mx + ix];
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

int mpde_fit_spline(const double* x_dev, int32_t mx, const double* t_dev, int32_t mt, const double* z_dev, int32_t k, double* tx_dev,
                    double* ty_dev, double* c_dev, double* work_dev, void* stream) {
    if (!x_dev || !t_dev || !z_dev || !tx_dev || !ty_dev || !c_dev || !work_dev) return -1;
    if ((k != 1 && k != 3) || mx < 2 * (k + 1) || mt < 2 * (k + 1)) return -1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* abx = work_dev;
    double* aby = abx + (size_t)BW * mx;
    double* w = aby + (size_t)BW * mt;
    spline_knots_lu_kernel<<<2, 256, 0, st>>>(x_dev, mx, t_dev, mt, k, tx_dev, ty_dev, abx, aby);
    spline_solve_x_kernel<<<(unsigned)((mt + 63) / 64), 64, 0, st>>>(abx, mx, mt, z_dev, w);
    spline_solve_y_kernel<<<(unsigned)((mx + 63) / 64), 64, 0, st>>>(aby, mx, mt, w, c_dev);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // extern "C"
