// Kuramoto-Sivashinsky environment step (ETDRK4), warp-resident variant (N = 8..256), sm_100a.
//
// Restates /root/reference/python/_model/KS.py: step :230-274 (Kassam-Trefethen ETDRK4 with the
// host-computed tables E, E2, Q, f1, f2, f3 of :127-137), fou2real :316-320, compute_Ek
// :322-343, getState :369-383, and the spectral reward of ks_environment.py:98-100.
// Same team layout as the Burgers kernel (one environment per team of TS lanes, half
// spectrum in registers, shuffle real-FFT); one step = 4 x (inverse, square, forward).
#pragma once
#include "params.h"
#include "warp_fft.cuh"
#include "burgers_warp.cuh"

namespace mpde {

constexpr int F_KS_UUROW = 1 << 8;   // the float32 row uu[ioutnum] is current (fou2real was called)

// WW ("whole warp"): compile-time promise that no history is recorded -- the only place where a team's control flow around
// a collective depends on its own environment (a blown-up environment stops writing rows) -- so every shuffle and
// __syncwarp may name the full warp mask (RealFFT::whole_warp).
template <typename T, int N, int TS_, bool WW = false>
struct KSWarp {
    using R = RealFFT<T, N, TS_>;
    using BW = BurgersWarp<T, N, TS_, -1, false>;
    static constexpr int H = N / 2, TS = R::TS, P = R::P, NH = N / 2 + 1, TPW = 32 / TS;

    // float32 periodic stencils of getState / the dforce=False forcing (KS.py:374-379, 241-244):
    // uu is float32 there, so differences and divisions round to float32
    __device__ __forceinline__ static float d2_f32(float um, float u, float up, float dx2f) {
        return __fdiv_rn(__fadd_rn(__fsub_rn(up, __fmul_rn(2.0f, u)), um), dx2f);
    }

    __device__ static void run(const SpectralParams<T>& prm, T* smem) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
        // ETDRK4 tables of the half spectrum, shared by every environment of the CTA: [7][H+1] = E, E2, Q, f1, f2, f3, -k/2
        // (read from shared memory at each use instead of living in 7 P registers per lane: fewer registers, more
        // resident warps, fewer waves)
        constexpr int TW = H + 1;
        // per team: [FFT exchange area (transposed transforms only) | scratch]; then the shared tables
        const int scr_team = 2 * R::SMEM_CX + max(prm.M, 2 * N + N / 2);
        T* const tab = smem + (size_t)wpc * TPW * scr_team;
        for (int i = threadIdx.x; i < 7 * TW; i += blockDim.x) {
            const int t = i / TW, k = i - t * TW;
            tab[i] = t < 6 ? prm.etd[t * N + k] : T(-0.5) * prm.kwave[k];
        }
        __syncthreads();
        const int64_t first = ((int64_t)blockIdx.x * wpc + warp) * TPW;
        if (first >= prm.B) return;
        const int team = lane / TS;
        T* const team_smem = smem + (size_t)(warp * TPW + team) * scr_team;
        R f;
        f.init(prm.tw, reinterpret_cast<Cx<T>*>(team_smem));
        if constexpr (WW) f.whole_warp();
        const int tl = f.c.tl;
        const int64_t e = first + team;
        const bool has = e < prm.B;
        const int64_t ec = has ? e : 0;
        const int flags = prm.flags;
        T* scratch = team_smem + 2 * R::SMEM_CX;
        const T dt = prm.dt, invN = T(1) / T(N);

        int kk[P];
#pragma unroll
        for (int p = 0; p < P; ++p) kk[p] = f.k(p);
        // table t at the wavenumber of register p / at the Nyquist mode; gk = -k/2: N(w) = g fft(u^2), g = -i k / 2
        auto tb = [&](int t, int p) { return tab[t * TW + kk[p]]; };
        auto tbN = [&](int t) { return tab[t * TW + H]; };
        enum { TE = 0, TE2 = 1, TQ = 2, TF1 = 3, TF2 = 4, TF3 = 5, TG = 6 };
        Cx<T> ws1[P];
        f.scaled_twiddles(T(1), ws1);
        // the nonlinear term uses the UNSCALED split step (X2 = 2 fft(U^2), U = N u): 1/(2 N^2) is folded into g = -k/2
        const T nl_scale = T(0.5) * invN * invN;

        pdl_wait();                  // everything above reads constant tables only
        pdl_launch_dependents();
        const bool was_live = has && prm.status[ec] == 0;
        bool live = was_live;
        int iout = prm.iout[ec];
        T tnow = prm.tnow[ec];
        Cx<T> v[P];
#pragma unroll
        for (int p = 0; p < P; ++p) v[p] = ldcx(prm.v + ec * NH + kk[p]);
        Cx<T> vN = ldcx(prm.v + ec * NH + H);
        const T v0im = v[0].im;

        // float32 row of uu (fou2real: Re ifft of the complex64 history row, KS.py:316-320)
        auto real_row_f32 = [&](Cx<T> (&u32)[P]) {
            Cx<T> w[P];
#pragma unroll
            for (int p = 0; p < P; ++p) w[p] = cx<T>((T)(float)v[p].re, (T)(float)v[p].im);
            f.inv(w, (T)(float)vN.re, u32);
#pragma unroll
            for (int p = 0; p < P; ++p) u32[p] = cx<T>((T)(float)(u32[p].re * invN), (T)(float)(u32[p].im * invN));
        };

        // ---- action forcing (KS.py:233-249) ------------------------------------------------------
        Cx<T> fa[P], F[P];
        T FN = T(0);
#pragma unroll
        for (int p = 0; p < P; ++p) { fa[p] = cx<T>(0, 0); F[p] = cx<T>(0, 0); }
        if (flags & F_ACTIONS) {
            for (int i = tl; i < prm.M; i += TS) scratch[i] = prm.actions[ec * prm.M + i];
            __syncwarp(f.c.smask);
#pragma unroll
            for (int p = 0; p < P; ++p) {
                T val[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int n = 2 * (p * TS + tl) + h;
                    T acc = T(0);
                    if (flags & F_BASIS_DENSE) {
                        for (int i = 0; i < prm.M; ++i) acc = fma(scratch[i], prm.basis[(size_t)i * N + n], acc);
                    } else {
                        acc = prm.tap_w[2 * n] * scratch[prm.tap_idx[2 * n]] +
                              prm.tap_w[2 * n + 1] * scratch[prm.tap_idx[2 * n + 1]];
                    }
                    val[h] = acc;
                }
                fa[p] = cx<T>(val[0], val[1]);
            }
            __syncwarp(f.c.smask);
            if (flags & F_DFORCE) {
                Cx<T> z[P];
#pragma unroll
                for (int p = 0; p < P; ++p) z[p] = fa[p];
                f.fwd(z, F, FN, T(1), ws1);
            }
        }

        float acc32[P];
#pragma unroll
        for (int p = 0; p < P; ++p) acc32[p] = prm.acc[ec * NH + kk[p]];
        float accN = prm.acc[ec * NH + H];
        const float dxf = (float)prm.dx;
        const float dx2f = __fmul_rn(dxf, dxf);         // self.dx**2 is a python float -> weak scalar -> float32(dx^2)
        const float dx2w = (float)((double)prm.dx * (double)prm.dx);
        bool bad = false;

        // N(w) = g * fft(Re(ifft(w))^2), g = -i k / 2   (KS.py:256-262)
        auto nonlinear = [&](const Cx<T> (&w)[P], T wNre, Cx<T> (&out)[P], T& outNim) {
            Cx<T> z[P], X[P];
            T XN;
            f.inv(w, wNre, z);
#pragma unroll
            for (int p = 0; p < P; ++p) z[p] = cx<T>(z[p].re * z[p].re, z[p].im * z[p].im);
            f.fwd_raw(z, X, XN);
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const T g = tb(TG, p) * nl_scale;
                out[p] = cx<T>(-g * X[p].im, g * X[p].re);     // (i gk) X, gk = -k/2
            }
            outNim = (tbN(TG) * nl_scale) * XN;
        };

        const int nsub = (flags & F_NO_ADVANCE) ? 0 : prm.nsub;
        const bool eddy = (flags & F_ACTIONS) && !(flags & F_DFORCE);
        for (int it = 0; it < nsub; ++it) {
            if (eddy) {
                // dforce == False multiplies by the stencil of the FLOAT32 row uu[ioutnum], which is only
                // current right after fou2real/getState; later rows of the reference's uu are zero, so
                // every further sub-step of the same call gets zero forcing (KS.py:241-245).
                Cx<T> z[P];
                if ((flags & F_KS_UUROW) && it == 0) {
                    Cx<T> u32[P];
                    real_row_f32(u32);
                    T left[P], right[P];
                    BW::halo(f, u32, left, right);
#pragma unroll
                    for (int p = 0; p < P; ++p)
                        z[p] = cx<T>(fa[p].re * (T)d2_f32((float)right_even(u32[p]), (float)u32[p].re, (float)left[p], dx2w),
                                     fa[p].im * (T)d2_f32((float)right[p], (float)u32[p].im, (float)u32[p].re, dx2w));
                } else {
#pragma unroll
                    for (int p = 0; p < P; ++p) z[p] = cx<T>(0, 0);
                }
                f.fwd(z, F, FN, T(1), ws1);
            }
            Cx<T> Nv[P], Na[P], Nb[P], Nc[P], a[P], b[P];
            T NvN, NaN, NbN, NcN;
            nonlinear(v, vN.re, Nv, NvN);
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const T e2 = tb(TE2, p), q = tb(TQ, p);
                a[p] = cx<T>(fma(e2, v[p].re, q * Nv[p].re), fma(e2, v[p].im, q * Nv[p].im));
            }
            const T E2N = tbN(TE2);
            const T aNre = E2N * vN.re;                  // the Nyquist entry of N(.) is purely imaginary
            nonlinear(a, aNre, Na, NaN);
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const T e2 = tb(TE2, p), q = tb(TQ, p);
                b[p] = cx<T>(fma(e2, v[p].re, q * Na[p].re), fma(e2, v[p].im, q * Na[p].im));
            }
            nonlinear(b, aNre, Nb, NbN);
#pragma unroll
            for (int p = 0; p < P; ++p) {     // c = E2 a + Q (2 Nb - Nv), reusing b
                const T e2 = tb(TE2, p), q = tb(TQ, p);
                b[p] = cx<T>(fma(e2, a[p].re, q * (T(2) * Nb[p].re - Nv[p].re)),
                             fma(e2, a[p].im, q * (T(2) * Nb[p].im - Nv[p].im)));
            }
            nonlinear(b, E2N * aNre, Nc, NcN);
            // v <- E v + (Nv + F) f1 + 2 (Na + Nb + 2 F) f2 + (Nc + F) f3   (KS.py:265)
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const T Ep = tb(TE, p), f1p = tb(TF1, p), f2p = tb(TF2, p), f3p = tb(TF3, p);
                v[p] = cx<T>(Ep * v[p].re + (Nv[p].re + F[p].re) * f1p + T(2) * (Na[p].re + Nb[p].re + T(2) * F[p].re) * f2p +
                                 (Nc[p].re + F[p].re) * f3p,
                             Ep * v[p].im + (Nv[p].im + F[p].im) * f1p + T(2) * (Na[p].im + Nb[p].im + T(2) * F[p].im) * f2p +
                                 (Nc[p].im + F[p].im) * f3p);
            }
            const T EN = tbN(TE), f1N = tbN(TF1), f2N = tbN(TF2), f3N = tbN(TF3);
            vN = cx<T>(EN * vN.re + FN * f1N + T(2) * (T(2) * FN) * f2N + FN * f3N,
                       EN * vN.im + NvN * f1N + T(2) * (NaN + NbN) * f2N + NcN * f3N);
            if (f.dc) v[0].im = tb(TE, 0) * v0im;
            iout += 1;
            tnow += dt;
            // float32 spectrum chain; the complex64 cast is also where the reference detects a blow-up (KS.py:7,273)
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const float fre = (float)v[p].re, fim = (float)v[p].im;
                bad |= !(fabsf(fre) <= FLT_MAX && fabsf(fim) <= FLT_MAX);
                acc32[p] = __fadd_rn(acc32[p], ek_row_f32(fre, fim, N, dxf));
            }
            {
                const float fre = (float)vN.re, fim = (float)vN.im;
                if (f.dc) bad |= !(fabsf(fre) <= FLT_MAX && fabsf(fim) <= FLT_MAX);
                accN = __fadd_rn(accN, ek_row_f32(fre, fim, N, dxf));
            }

            if (!WW && prm.hist_rows > 0) {
                live = live && !BW::team_any(f, bad);
                if (live && iout < prm.hist_rows) {
                    const int64_t hrow = e * prm.hist_rows + iout;
                    Cx<T> u32[P];
                    if (prm.uu_hist) real_row_f32(u32);
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        if (prm.uu_hist) stcx(reinterpret_cast<Cx<T>*>(prm.uu_hist + hrow * N) + p * TS + tl, u32[p]);
                        if (prm.vv_hist) {
                            Cx<float> c; c.re = (float)v[p].re; c.im = (float)v[p].im;
                            prm.vv_hist[hrow * N + kk[p]] = c;
                            if (kk[p] != 0) { c.im = -c.im; prm.vv_hist[hrow * N + N - kk[p]] = c; }
                        }
                        if (prm.ektt_hist) prm.ektt_hist[hrow * NH + kk[p]] = (double)acc32[p] / (double)(iout + 1);
                    }
                    if (f.dc) {
                        if (prm.vv_hist) { Cx<float> c; c.re = (float)vN.re; c.im = (float)vN.im; prm.vv_hist[hrow * N + H] = c; }
                        if (prm.ektt_hist) prm.ektt_hist[hrow * NH + H] = (double)accN / (double)(iout + 1);
                    }
                }
            }
        }

        if (nsub > 0) {
            const bool blew = BW::team_any(f, bad);
            if (was_live && blew && f.dc) prm.status[e] = 1;
            live = live && !blew;
        }
        if (nsub > 0 && live) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                stcx(prm.v + e * NH + kk[p], v[p]);
                prm.acc[e * NH + kk[p]] = acc32[p];
            }
            if (f.dc) {
                stcx(prm.v + e * NH + H, vN);
                prm.acc[e * NH + H] = accN;
                prm.iout[e] = iout;
                prm.tnow[e] = tnow;
            }
        }

        const T inf = T(1) / T(0);
        if (prm.state_out) {
            // getState (KS.py:369-383) on the float32 row: [ (u_{j+1}-u_{j-1})/(2 dx) ; (u_{j+1}-2u_j+u_{j-1})/dx^2 ]
            Cx<T> u32[P];
            real_row_f32(u32);
            T left[P], right[P];
            BW::halo(f, u32, left, right);
            const float two_dx = __fmul_rn(2.0f, dxf);
            if (has) {
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const int j = p * TS + tl;
                    const float ul = (float)left[p], ue = (float)u32[p].re, uo = (float)u32[p].im, ur = (float)right[p];
                    const Cx<T> dudx = cx<T>((T)__fdiv_rn(__fsub_rn(uo, ul), two_dx), (T)__fdiv_rn(__fsub_rn(ur, ue), two_dx));
                    const Cx<T> d2 = cx<T>((T)d2_f32(ul, ue, uo, dx2w), (T)d2_f32(ue, uo, ur, dx2w));
                    const Cx<T> ii = cx<T>(inf, inf);
                    stcx(reinterpret_cast<Cx<T>*>(prm.state_out + e * 2 * N) + j, live ? dudx : ii);
                    stcx(reinterpret_cast<Cx<T>*>(prm.state_out + e * 2 * N + N) + j, live ? d2 : ii);
                }
            }
            (void)dx2f;
        }

        if (prm.reward_out && prm.reward_mode == REWARD_SPECTRAL && nsub > 0) {
            // ks_environment.py:98-100 on the running float32 sums
            const int A = prm.A;
            const int64_t ref = prm.ek_map ? prm.ek_map[ec] : 0;
            const int64_t row = iout < prm.ek_rows ? iout : prm.ek_rows - 1;
            T part = T(0);
#pragma unroll
            for (int p = 0; p < P; ++p)
                if (kk[p] >= 1) {
                    const T ed = (T)prm.ek_ref[(ref * prm.ek_rows + row) * H + kk[p]];
                    const T es = (T)((double)acc32[p] / (double)(iout + 1));
                    const T q = fabs(ed - es) / ed;
                    part += q * q;
                }
            part = BW::team_sum(f, part) / T(H - 1);
            const T prev = prm.kprev[ec];
            const T r = live ? prev - part : -inf;
            if (has) {
                for (int a = tl; a < A; a += TS) prm.reward_out[e * A + a] = r;
                if (f.dc && live) prm.kprev[e] = part;
            }
        }
    }

    __device__ __forceinline__ static T right_even(const Cx<T>& x) { return x.im; }   // u_{2j+1}: right neighbour of u_{2j}
};

}  // namespace mpde
