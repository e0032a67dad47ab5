// Reset / read-back kernels of the warp-resident spectral solvers (N <= 128).
// IC(u0) / IC(v0): Burger.py:289-320, KS.py:191-219.  Re ifft(v): Burger.py:491.
#pragma once
#include "params.h"
#include "warp_fft.cuh"

namespace mpde {


// energy-spectrum row in the reference's float32 chain (Burger.py:562 on complex64 data)
template <typename T>
__device__ __forceinline__ float ek_row_f32(Cx<T> v, int N, float dxf) {
    const float re = (float)v.re, im = (float)v.im;
    const float en = __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));
    return __fmul_rn(en * (0.5f / (float)N), dxf);
}

template <typename T, int N, int MODE>
__global__ void __launch_bounds__(128) aux_warp_kernel(const SpectralParams<T> prm, const void* __restrict__ src_,
                                                       const uint8_t* __restrict__ mask, void* __restrict__ dst_,
                                                       int equation) {
    using F = WarpFFT<T, N>;
    constexpr int TS = F::TS, P = F::P, NH = N / 2 + 1, TPW = 32 / TS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    F f;
    f.init(prm.tw);
    const int team = lane / TS;
    const int64_t pair = ((int64_t)blockIdx.x * wpc + warp) * TPW + team;
    if (2 * (((int64_t)blockIdx.x * wpc + warp) * TPW) >= prm.B) return;
    const int64_t e[2] = {2 * pair, 2 * pair + 1};
    bool has[2], sel[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        has[s] = e[s] < prm.B;
        sel[s] = has[s] && (mask == nullptr || mask[e[s]] != 0);
    }
    const T invN = T(1) / T(N);
    Cx<T> v[2][P];
    T u[2][P];
    Cx<T> z[P];

    if constexpr (MODE == AUX_RESET_U) {
        const T* src = static_cast<const T*>(src_);
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int p = 0; p < P; ++p) u[s][p] = sel[s] ? src[e[s] * N + p * TS + f.tl] : T(0);
#pragma unroll
        for (int p = 0; p < P; ++p) z[p] = cx<T>(u[0][p], u[1][p]);
        f.fwd(z);
        f.untangle(z, v[0], v[1], T(1));
    } else {
        const Cx<T>* src = MODE == AUX_RESET_V ? static_cast<const Cx<T>*>(src_) : nullptr;
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const int k = F::kidx(p, f.tl);
                Cx<T> a = cx<T>(0, 0);
                if (MODE == AUX_RESET_V) {
                    if (sel[s]) {
                        a = ldcx(src + e[s] * N + k);
                        if (k != 0 && k != N / 2) {     // Hermitian part: all that Re ifft ever sees
                            const Cx<T> b = ldcx(src + e[s] * N + (N - k));
                            a = cx<T>(T(0.5) * (a.re + b.re), T(0.5) * (a.im - b.im));
                        }
                    }
                } else if (has[s]) {
                    const int kh = k <= N / 2 ? k : N - k;
                    a = ldcx(prm.v + e[s] * NH + kh);
                    if (k > N / 2) a = conj(a);
                }
                v[s][p] = a;
            }
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int k = F::kidx(p, f.tl);
            z[p] = F::tangle(v[0][p], v[1][p], k == 0 || k == N / 2);
        }
        f.inv(z);
#pragma unroll
        for (int p = 0; p < P; ++p) { u[0][p] = z[p].re * invN; u[1][p] = z[p].im * invN; }
    }

    if constexpr (MODE == AUX_GET_U) {
        T* dst = static_cast<T*>(dst_);
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int p = 0; p < P; ++p)
                if (has[s]) dst[e[s] * N + p * TS + f.tl] = u[s][p];
        return;
    } else {
        // Fn_old = i k fft(u0^2 / 2)  (Burger.py:320)
        Cx<T> X[2][P];
        if (equation == 0) {
#pragma unroll
            for (int p = 0; p < P; ++p) z[p] = cx<T>(u[0][p] * u[0][p], u[1][p] * u[1][p]);
            f.fwd(z);
            f.untangle(z, X[0], X[1], T(0.5));
        }
        const float dxf = (float)prm.dx;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (!sel[s]) continue;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const int k = F::kidx(p, f.tl);
                if (k <= N / 2) {
                    stcx(prm.v + e[s] * NH + k, v[s][p]);
                    if (equation == 0) {
                        const T kw = prm.kwave[k];
                        stcx(prm.fn + e[s] * NH + k, cx<T>(-kw * X[s][p].im, kw * X[s][p].re));
                    }
                    prm.acc[e[s] * NH + k] = ek_row_f32(v[s][p], N, dxf);
                }
                if (prm.uprev) prm.uprev[e[s] * N + p * TS + f.tl] = u[s][p];
                if (prm.hist_rows > 0) {
                    const int64_t hrow = e[s] * prm.hist_rows;
                    if (prm.uu_hist) prm.uu_hist[hrow * N + p * TS + f.tl] = u[s][p];
                    if (prm.vv_hist) {
                        Cx<float> c; c.re = (float)v[s][p].re; c.im = (float)v[s][p].im;
                        prm.vv_hist[hrow * N + k] = c;
                    }
                    if (prm.ektt_hist && k <= N / 2) prm.ektt_hist[hrow * NH + k] = (double)ek_row_f32(v[s][p], N, dxf);
                }
            }
            if (f.tl == 0) {
                prm.iout[e[s]] = 0;
                prm.tnow[e[s]] = T(0);
                prm.kprev[e[s]] = T(0);
                prm.status[e[s]] = 0;
            }
        }
    }
}

}  // namespace mpde
