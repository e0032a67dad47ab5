// Reset / read-back kernels of the warp-resident spectral solvers (N <= 256).
// IC(u0) / IC(v0): Burger.py:289-320, KS.py:191-219.  Re ifft(v): Burger.py:491.
#pragma once
#include "params.h"
#include "warp_fft.cuh"
#include "burgers_warp.cuh"   // ek_row_f32

namespace mpde {

template <typename T, int N, int MODE>
__global__ void __launch_bounds__(128) aux_warp_kernel(const SpectralParams<T> prm, const void* __restrict__ src_,
                                                       const uint8_t* __restrict__ mask, void* __restrict__ dst_,
                                                       int equation) {
    using R = RealFFT<T, N>;
    constexpr int TS = R::TS, P = R::P, H = N / 2, NH = N / 2 + 1, TPW = 32 / TS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    const int64_t first = ((int64_t)blockIdx.x * wpc + warp) * TPW;
    if (first >= prm.B) return;
    R f;
    f.init(prm.tw);
    const int tl = f.c.tl;
    const int64_t e = first + lane / TS;
    const bool has = e < prm.B;
    const int64_t ec = has ? e : 0;
    const bool sel = has && (mask == nullptr || mask[ec] != 0);
    const T invN = T(1) / T(N);
    Cx<T> v[P], u[P];
    Cx<T> vN = cx<T>(0, 0);
    int kk[P];
#pragma unroll
    for (int p = 0; p < P; ++p) kk[p] = f.k(p);

    if constexpr (MODE == AUX_RESET_U) {
        const Cx<T>* src = static_cast<const Cx<T>*>(src_) + ec * H;     // (u_{2j}, u_{2j+1}) pairs
        Cx<T> z[P];
#pragma unroll
        for (int p = 0; p < P; ++p) { u[p] = ldcx(src + p * TS + tl); z[p] = u[p]; }
        T nyq;
        Cx<T> ws1[P];
        f.scaled_twiddles(T(1), ws1);
        f.fwd(z, v, nyq, T(1), ws1);
        vN = cx<T>(nyq, T(0));
    } else {
        if constexpr (MODE == AUX_RESET_V) {
            const Cx<T>* src = static_cast<const Cx<T>*>(src_) + ec * N;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                Cx<T> a = ldcx(src + kk[p]);
                if (kk[p] != 0) {       // Hermitian part: all that Re ifft(v0) ever sees of 0 < k < N/2
                    const Cx<T> b = ldcx(src + (N - kk[p]));
                    a = cx<T>(T(0.5) * (a.re + b.re), T(0.5) * (a.im - b.im));
                }
                v[p] = a;
            }
            vN = ldcx(src + H);         // kept complex (quirk Q5)
        } else {
#pragma unroll
            for (int p = 0; p < P; ++p) v[p] = ldcx(prm.v + ec * NH + kk[p]);
            vN = ldcx(prm.v + ec * NH + H);
        }
        f.inv(v, vN.re, u);
#pragma unroll
        for (int p = 0; p < P; ++p) u[p] = cx<T>(u[p].re * invN, u[p].im * invN);
    }

    if constexpr (MODE == AUX_GET_U) {
        Cx<T>* dst = static_cast<Cx<T>*>(dst_) + ec * H;
        if (has) {
#pragma unroll
            for (int p = 0; p < P; ++p) stcx(dst + p * TS + tl, u[p]);
        }
        return;
    } else {
        // Fn_old = i k fft(u0^2 / 2)  (Burger.py:320)
        Cx<T> X[P];
        T XN = T(0);
        if (equation == 0) {
            Cx<T> z[P];
#pragma unroll
            for (int p = 0; p < P; ++p) z[p] = cx<T>(u[p].re * u[p].re, u[p].im * u[p].im);
            Cx<T> wsh[P];
            f.scaled_twiddles(T(0.5), wsh);
            f.fwd(z, X, XN, T(0.5), wsh);
        }
        if (!sel) return;
        const float dxf = (float)prm.dx;
        const int64_t hrow = e * prm.hist_rows;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int k = kk[p];
            stcx(prm.v + e * NH + k, v[p]);
            if (equation == 0) {
                const T kw = prm.kwave[k];
                stcx(prm.fn + e * NH + k, cx<T>(-kw * X[p].im, kw * X[p].re));
            }
            const float ek = ek_row_f32((float)v[p].re, (float)v[p].im, N, dxf);
            prm.acc[e * NH + k] = ek;
            stcx(reinterpret_cast<Cx<T>*>(prm.uprev + e * N) + p * TS + tl, u[p]);
            if (prm.hist_rows > 0) {
                if (prm.uu_hist) stcx(reinterpret_cast<Cx<T>*>(prm.uu_hist + hrow * N) + p * TS + tl, u[p]);
                if (prm.vv_hist) {
                    Cx<float> c; c.re = (float)v[p].re; c.im = (float)v[p].im;
                    prm.vv_hist[hrow * N + k] = c;
                    if (k != 0) { c.im = -c.im; prm.vv_hist[hrow * N + N - k] = c; }
                }
                if (prm.ektt_hist) prm.ektt_hist[hrow * NH + k] = (double)ek;
            }
        }
        if (f.dc) {
            stcx(prm.v + e * NH + H, vN);
            if (equation == 0) stcx(prm.fn + e * NH + H, cx<T>(T(0), prm.kwave[H] * XN));
            const float ekN = ek_row_f32((float)vN.re, (float)vN.im, N, dxf);
            prm.acc[e * NH + H] = ekN;
            if (prm.hist_rows > 0) {
                if (prm.vv_hist) { Cx<float> c; c.re = (float)vN.re; c.im = (float)vN.im; prm.vv_hist[hrow * N + H] = c; }
                if (prm.ektt_hist) prm.ektt_hist[hrow * NH + H] = (double)ekN;
            }
            prm.iout[e] = 0;
            prm.tnow[e] = T(0);
            prm.kprev[e] = T(0);
            prm.status[e] = 0;
        }
    }
}

}  // namespace mpde
