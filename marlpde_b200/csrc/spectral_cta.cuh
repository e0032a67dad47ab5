// CTA-per-environment variants of the Burgers (ABCN) and KS (ETDRK4) steps for N = 256..4096
// (DNS / ground-truth generation, Burger.py:333-499 and KS.py:230-274), sm_100a.
// The environment state (half spectra, real field) stays in shared memory across the `nsub`
// fused sub-steps; per-step history rows (uu, vv, Ek_ktt) stream out coalesced.
// The Smagorinsky closures are only available in the warp-resident kernels (N <= 256).
#pragma once
#include "params.h"
#include "cta_fft.cuh"
#include "burgers_warp.cuh"   // ek_row_f32
#include "ks_warp.cuh"        // F_KS_UUROW

namespace mpde {

template <typename T, int N, int NT, int EQ>
struct SpectralCta {
    using F = CtaFFT<T, N, NT>;
    static constexpr int H = N / 2, NH = H + 1;
    // shared-memory budget in units of Cx<T>
    static constexpr int n_cx = (EQ == 0) ? (H /*buf*/ + NH /*X*/ + NH /*v*/ + NH /*fn*/ + H /*U*/ + NH /*Fh*/ + H /*fa*/ + NH /*Fa*/)
                                           : (H /*buf*/ + NH /*X*/ + 7 * NH /*v a b Nv Na Nb Nc*/ + NH /*Fh*/ + H /*fa*/);
    static size_t smem_bytes(int M) { return sizeof(Cx<T>) * n_cx + sizeof(T) * (H + (M > 0 ? M : 1)) + 16; }

    __device__ __forceinline__ static T r32(T a) { return (T)(float)a; }

    __device__ static void run(const SpectralParams<T>& prm, unsigned char* smem_raw) {
        const int64_t e = blockIdx.x;
        const int t = threadIdx.x;
        const int flags = prm.flags;
        const bool q1 = EQ == 0 && !(flags & F_FORCING);
        Cx<T>* buf = reinterpret_cast<Cx<T>*>(smem_raw);
        Cx<T>* X = buf + H;
        Cx<T>* v = X + NH;
        Cx<T>* w1 = v + NH;                    // Burgers: fn            | KS: a
        Cx<T>* w2 = w1 + NH;                   // Burgers: U (H entries) | KS: b / c
        Cx<T>* rest = w2 + (EQ == 0 ? H : NH);
        Cx<T>* Fh = rest;                      // forcing spectrum
        Cx<T>* fa = Fh + NH;                   // action field (H pairs)
        Cx<T>* Fa = fa + H;                    // Burgers: spectrum of a direct forcing
        Cx<T>* Nv = fa + H;                    // KS: nonlinear stage values (alias of Fa's region onwards)
        Cx<T>* Na = Nv + NH;
        Cx<T>* Nb = Na + NH;
        Cx<T>* Nc = Nb + NH;
        T* ekbuf = reinterpret_cast<T*>((EQ == 0 ? Fa + NH : Nc + NH));
        T* act = ekbuf + H;
        const Cx<T>* tw = prm.tw;
        const T dt = prm.dt, invN = T(1) / T(N);
        const T nu = prm.nu[e];
        const float dxf = (float)prm.dx, dtf = (float)dt;

        const bool was_live = prm.status[e] == 0;
        int iout = prm.iout[e];
        T tnow = prm.tnow[e];
        for (int k = t; k < NH; k += NT) {
            v[k] = ldcx(prm.v + e * NH + k);
            if (EQ == 0) w1[k] = ldcx(prm.fn + e * NH + k);
            Fh[k] = cx<T>(0, 0);
        }
        __syncthreads();

        // ---- action field ---------------------------------------------------------------------------
        const bool has_act = flags & F_ACTIONS;
        const bool eddy = has_act && !(flags & F_DFORCE);
        if (has_act) {
            for (int i = t; i < prm.M; i += NT) act[i] = prm.actions[e * prm.M + i];
            __syncthreads();
            T* far = reinterpret_cast<T*>(fa);
            for (int n = t; n < N; n += NT) {
                T acc = T(0);
                if (flags & F_BASIS_DENSE) {
                    for (int i = 0; i < prm.M; ++i) acc = fma(act[i], prm.basis[(size_t)i * N + n], acc);
                } else {
                    acc = prm.tap_w[2 * n] * act[prm.tap_idx[2 * n]] + prm.tap_w[2 * n + 1] * act[prm.tap_idx[2 * n + 1]];
                }
                far[n] = acc;
            }
            __syncthreads();
            if (!eddy) {      // direct forcing: constant spectrum for the whole call
                for (int j = t; j < H; j += NT) buf[j] = fa[j];
                __syncthreads();
                F::rfwd(buf, EQ == 0 ? Fa : Fh, T(1), tw);
            }
        }

        bool bad = false;
        bool live = was_live;
        T* Ur = reinterpret_cast<T*>(w2);     // Burgers: real field u (natural order, N reals)

        // Burgers: u = Re ifft(v)
        if (EQ == 0) {
            F::rinv(v, buf, tw);
            for (int j = t; j < H; j += NT) w2[j] = cx<T>(buf[j].re * invN, buf[j].im * invN);
            __syncthreads();
        }

        const int nsub = (flags & F_NO_ADVANCE) ? 0 : prm.nsub;

        for (int it = 0; it < nsub; ++it) {
            if (EQ == 0 && it == nsub - 1) {       // u before the last sub-step (dudt of state version 1)
                for (int j = t; j < H; j += NT) stcx(reinterpret_cast<Cx<T>*>(prm.uprev + e * N) + j, w2[j]);
            }
            if (EQ == 0) {
                // ---------------- Burgers ABCN (Burger.py:486-491) ----------------
                for (int j = t; j < H; j += NT) buf[j] = cx<T>(w2[j].re * w2[j].re, w2[j].im * w2[j].im);
                __syncthreads();
                F::rfwd(buf, X, T(0.5), tw);                              // X = fft(u^2/2)
                if (flags & F_FORCING) {
                    const int64_t row = (flags & F_FORCING_PER_ENV) ? e : 0;
                    const int col = iout % prm.stepper;
                    for (int k = t; k < NH; k += NT)
                        Fh[k] = (k >= 1 && k <= 3) ? ldcx(prm.fcoef + (row * prm.stepper + col) * 3 + (k - 1)) : cx<T>(0, 0);
                } else {
                    for (int k = t; k < NH; k += NT) Fh[k] = cx<T>(0, 0);
                }
                if (eddy) {
                    const T s = T(1) / (prm.dx * prm.dx);
                    for (int j = t; j < H; j += NT) {
                        const T ul = Ur[(2 * j + N - 1) & (N - 1)], ue = Ur[2 * j], uo = Ur[2 * j + 1], ur = Ur[(2 * j + 2) & (N - 1)];
                        buf[j] = cx<T>(fa[j].re * ((ul - T(2) * ue + uo) * s), fa[j].im * ((ue - T(2) * uo + ur) * s));
                    }
                    __syncthreads();
                    F::rfwd(buf, Fa, T(1), tw);
                }
                __syncthreads();
                for (int k = t; k < NH; k += NT) {
                    Cx<T> Fk = Fh[k];
                    if (has_act) {
                        Fk = Fk + Fa[k];
                        if (q1) Fk = cx<T>(r32(Fk.re), r32(Fk.im));
                    }
                    const T kw = prm.kwave[k];
                    const T C = T(0.5) * (kw * kw) * nu * dt;
                    const T r = T(1) / (T(1) + C);
                    const Cx<T> fnn = cx<T>(-kw * X[k].im, kw * X[k].re);
                    const Cx<T> dtF = q1 ? cx<T>((T)__fmul_rn(dtf, (float)Fk.re), (T)__fmul_rn(dtf, (float)Fk.im))
                                         : cx<T>(dt * Fk.re, dt * Fk.im);
                    const Cx<T> vo = v[k], fo = w1[k];
                    Cx<T> vn = cx<T>(((T(1) - C) * vo.re - T(0.5) * dt * (T(3) * fnn.re - fo.re) + dtF.re) * r,
                                     ((T(1) - C) * vo.im - T(0.5) * dt * (T(3) * fnn.im - fo.im) + dtF.im) * r);
                    v[k] = vn;
                    w1[k] = fnn;
                    bad |= blown(vn);
                }
                __syncthreads();
                F::rinv(v, buf, tw);
                for (int j = t; j < H; j += NT) w2[j] = cx<T>(buf[j].re * invN, buf[j].im * invN);
                __syncthreads();
            } else {
                // ---------------- KS ETDRK4 (KS.py:255-267) ----------------
                auto nonlinear = [&](const Cx<T>* w, Cx<T>* out) {
                    F::rinv(w, buf, tw);
                    for (int j = t; j < H; j += NT) buf[j] = cx<T>(buf[j].re * buf[j].re, buf[j].im * buf[j].im);
                    __syncthreads();
                    F::rfwd(buf, X, invN * invN, tw);
                    for (int k = t; k < NH; k += NT) {
                        const T gk = T(-0.5) * prm.kwave[k];
                        out[k] = cx<T>(-gk * X[k].im, gk * X[k].re);
                    }
                    __syncthreads();
                };
                if (eddy) {       // KS.py:241-245 quirk: float32 row, only current right after fou2real
                    if ((flags & F_KS_UUROW) && it == 0) {
                        for (int k = t; k < NH; k += NT) X[k] = cx<T>((T)(float)v[k].re, (T)(float)v[k].im);
                        __syncthreads();
                        F::rinv(X, buf, tw);
                        T* ur = reinterpret_cast<T*>(buf);
                        for (int n = t; n < N; n += NT) ur[n] = (T)(float)(ur[n] * invN);
                        __syncthreads();
                        const float dx2w = (float)((double)prm.dx * (double)prm.dx);
                        T* far = reinterpret_cast<T*>(fa);
                        T* out = reinterpret_cast<T*>(w2);      // b is free here
                        for (int n = t; n < N; n += NT) {
                            const float um = (float)ur[(n + N - 1) & (N - 1)], u0 = (float)ur[n], up = (float)ur[(n + 1) & (N - 1)];
                            out[n] = far[n] * (T)__fdiv_rn(__fadd_rn(__fsub_rn(um, __fmul_rn(2.0f, u0)), up), dx2w);
                        }
                        __syncthreads();
                        for (int j = t; j < H; j += NT) buf[j] = w2[j];
                        __syncthreads();
                        F::rfwd(buf, Fh, T(1), tw);
                    } else {
                        for (int k = t; k < NH; k += NT) Fh[k] = cx<T>(0, 0);
                        __syncthreads();
                    }
                }
                const T* etd = prm.etd;
                nonlinear(v, Nv);
                for (int k = t; k < NH; k += NT) {
                    const T e2 = etd[1 * N + k], q = etd[2 * N + k];
                    w1[k] = cx<T>(fma(e2, v[k].re, q * Nv[k].re), fma(e2, v[k].im, q * Nv[k].im));
                }
                __syncthreads();
                nonlinear(w1, Na);
                for (int k = t; k < NH; k += NT) {
                    const T e2 = etd[1 * N + k], q = etd[2 * N + k];
                    w2[k] = cx<T>(fma(e2, v[k].re, q * Na[k].re), fma(e2, v[k].im, q * Na[k].im));
                }
                __syncthreads();
                nonlinear(w2, Nb);
                for (int k = t; k < NH; k += NT) {
                    const T e2 = etd[1 * N + k], q = etd[2 * N + k];
                    w2[k] = cx<T>(fma(e2, w1[k].re, q * (T(2) * Nb[k].re - Nv[k].re)),
                                  fma(e2, w1[k].im, q * (T(2) * Nb[k].im - Nv[k].im)));
                }
                __syncthreads();
                nonlinear(w2, Nc);
                for (int k = t; k < NH; k += NT) {
                    const T E = etd[0 * N + k], f1 = etd[3 * N + k], f2 = etd[4 * N + k], f3 = etd[5 * N + k];
                    const Cx<T> Fk = Fh[k];
                    const Cx<T> vn = cx<T>(
                        E * v[k].re + (Nv[k].re + Fk.re) * f1 + T(2) * (Na[k].re + Nb[k].re + T(2) * Fk.re) * f2 + (Nc[k].re + Fk.re) * f3,
                        E * v[k].im + (Nv[k].im + Fk.im) * f1 + T(2) * (Na[k].im + Nb[k].im + T(2) * Fk.im) * f2 + (Nc[k].im + Fk.im) * f3);
                    v[k] = vn;
                    bad |= blown(vn);
                }
                __syncthreads();
            }
            iout += 1;
            tnow += dt;

            // float32 spectrum chain + history rows
            const bool write_hist = prm.hist_rows > 0 && iout < prm.hist_rows;
            if (write_hist) live = live && !__syncthreads_or(bad);
            const int64_t hrow = e * prm.hist_rows + iout;
            for (int k = t; k < NH; k += NT) {
                float a = prm.acc[e * NH + k];
                a = __fadd_rn(a, ek_row_f32((float)v[k].re, (float)v[k].im, N, dxf));
                prm.acc[e * NH + k] = a;      // (an env that later blows up within this call keeps these sums; it is dead anyway)
                if (write_hist && live) {
                    if (prm.vv_hist) {
                        Cx<float> c; c.re = (float)v[k].re; c.im = (float)v[k].im;
                        prm.vv_hist[hrow * N + k] = c;
                        if (k != 0 && k != H) { c.im = -c.im; prm.vv_hist[hrow * N + N - k] = c; }
                    }
                    if (prm.ektt_hist) prm.ektt_hist[hrow * NH + k] = (double)a / (double)(iout + 1);
                }
            }
            if (write_hist && live && prm.uu_hist) {
                if (EQ == 0) {
                    for (int j = t; j < H; j += NT) stcx(reinterpret_cast<Cx<T>*>(prm.uu_hist + hrow * N) + j, w2[j]);
                } else {           // KS: uu = Re ifft(complex64(vv)) in float32 (fou2real)
                    for (int k = t; k < NH; k += NT) X[k] = cx<T>((T)(float)v[k].re, (T)(float)v[k].im);
                    __syncthreads();
                    F::rinv(X, buf, tw);
                    for (int j = t; j < H; j += NT)
                        stcx(reinterpret_cast<Cx<T>*>(prm.uu_hist + hrow * N) + j,
                             cx<T>((T)(float)(buf[j].re * invN), (T)(float)(buf[j].im * invN)));
                    __syncthreads();
                }
            }
            __syncthreads();
        }

        // ---- epilogue -----------------------------------------------------------------------------------
        if (nsub > 0) {
            const bool blew = __syncthreads_or(bad);
            if (was_live && blew && t == 0) prm.status[e] = 1;
            live = live && !blew;
            if (live) {
                for (int k = t; k < NH; k += NT) {
                    stcx(prm.v + e * NH + k, v[k]);
                    if (EQ == 0) stcx(prm.fn + e * NH + k, w1[k]);
                }
                if (t == 0) { prm.iout[e] = iout; prm.tnow[e] = tnow; }
            }
        }
        const T inf = T(1) / T(0);
        if (prm.state_out) {
            if (EQ == 0) {
                // getState (Burger.py:604-675)
                const int ver = prm.version, A = prm.A;
                T* f0 = reinterpret_cast<T*>(buf);
                T* f1 = reinterpret_cast<T*>(X);
                const T s = T(1) / (prm.dx * prm.dx);
                for (int n = t; n < N; n += NT) {
                    const T ul = Ur[(n + N - 1) & (N - 1)], u0 = Ur[n], ur = Ur[(n + 1) & (N - 1)];
                    const T d2 = (ul - T(2) * u0 + ur) * s;
                    const T up = prm.uprev[e * N + n];
                    const T dudt = (nsub > 0 || iout > 0) ? (u0 - up) / dt : T(0);
                    T a = d2, b = d2;
                    if (ver == 1) a = dudt;
                    else if (ver == 2) { a = u0; b = u0 * u0; }
                    else if (ver == 4) a = u0;
                    f0[n] = a;
                    f1[n] = b;
                }
                for (int k = t; k < H; k += NT)
                    ekbuf[k] = T(0.5) * ((v[k].re * v[k].re + v[k].im * v[k].im) / T(N)) * prm.dx;
                __syncthreads();
                const int nf = (ver == 1 || ver == 2) ? 2 : 1;
                const int seg = A == 1 ? N : N / A + 2;
                const int tail = (ver == 3 || ver == 4) ? N / 2 : 0;
                const int RL = nf * seg + tail, S = A * RL;
                for (int o = t; o < S; o += NT) {
                    const int a = o / RL, r = o - a * RL;
                    T val;
                    if (r < nf * seg) {
                        const int fld = r / seg, w = r - fld * seg;
                        const int start = A == 1 ? 0 : a * (N / A) - 1;
                        const int j = (start + w + N) & (N - 1);
                        val = fld == 0 ? f0[j] : f1[j];
                    } else {
                        val = ekbuf[r - nf * seg];
                    }
                    prm.state_out[e * S + o] = live ? val : inf;
                }
            } else {
                // KS getState (KS.py:369-383) on the float32 row
                for (int k = t; k < NH; k += NT) X[k] = cx<T>((T)(float)v[k].re, (T)(float)v[k].im);
                __syncthreads();
                F::rinv(X, buf, tw);
                T* ur = reinterpret_cast<T*>(buf);
                for (int n = t; n < N; n += NT) ur[n] = (T)(float)(ur[n] * invN);
                __syncthreads();
                const float two_dx = __fmul_rn(2.0f, dxf);
                const float dx2w = (float)((double)prm.dx * (double)prm.dx);
                for (int n = t; n < N; n += NT) {
                    const float um = (float)ur[(n + N - 1) & (N - 1)], u0 = (float)ur[n], up = (float)ur[(n + 1) & (N - 1)];
                    const T dudx = (T)__fdiv_rn(__fsub_rn(up, um), two_dx);
                    const T d2 = (T)__fdiv_rn(__fadd_rn(__fsub_rn(up, __fmul_rn(2.0f, u0)), um), dx2w);
                    prm.state_out[e * 2 * N + n] = live ? dudx : inf;
                    prm.state_out[e * 2 * N + N + n] = live ? d2 : inf;
                }
            }
            __syncthreads();
        }
        if (prm.reward_out && prm.reward_mode == REWARD_SPECTRAL && nsub > 0) {
            const int64_t ref = prm.ek_map ? prm.ek_map[e] : 0;
            const int64_t row = iout < prm.ek_rows ? iout : prm.ek_rows - 1;
            __syncthreads();
            for (int k = t; k < H; k += NT) {
                T q = T(0);
                if (k >= 1) {
                    const T ed = (T)prm.ek_ref[(ref * prm.ek_rows + row) * H + k];
                    const T es = (T)((double)prm.acc[e * NH + k] / (double)(iout + 1));
                    q = fabs(ed - es) / ed;
                }
                ekbuf[k] = q * q;
            }
            __syncthreads();
            if (t == 0) {
                T part = T(0);
                for (int k = 1; k < H; ++k) part += ekbuf[k];
                part /= T(H - 1);
                const T prev = prm.kprev[e];
                const T r = live ? prev - part : -inf;
                for (int a = 0; a < prm.A; ++a) prm.reward_out[e * prm.A + a] = r;
                if (live) prm.kprev[e] = part;
            }
        }
    }
};

// reset / read-back for the CTA-resident solvers (IC(u0), IC(v0), Re ifft(v))
template <typename T, int N, int NT, int MODE>
__global__ void __launch_bounds__(NT) aux_cta_kernel(const SpectralParams<T> prm, const void* __restrict__ src_,
                                                     const uint8_t* __restrict__ mask, void* __restrict__ dst_, int equation) {
    using F = CtaFFT<T, N, NT>;
    constexpr int H = N / 2, NH = H + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cx<T>* buf = reinterpret_cast<Cx<T>*>(smem_raw);
    Cx<T>* v = buf + H;
    Cx<T>* X = v + NH;
    Cx<T>* U = X + NH;
    const int64_t e = blockIdx.x;
    const int t = threadIdx.x;
    if (MODE != AUX_GET_U && mask != nullptr && mask[e] == 0) return;
    const T invN = T(1) / T(N);
    if (MODE == AUX_RESET_U) {
        const Cx<T>* src = static_cast<const Cx<T>*>(src_) + e * H;
        for (int j = t; j < H; j += NT) { U[j] = ldcx(src + j); buf[j] = U[j]; }
        __syncthreads();
        F::rfwd(buf, v, T(1), prm.tw);
    } else {
        if (MODE == AUX_RESET_V) {
            const Cx<T>* src = static_cast<const Cx<T>*>(src_) + e * N;
            for (int k = t; k < NH; k += NT) {
                Cx<T> a = ldcx(src + k);
                if (k != 0 && k != H) {
                    const Cx<T> b = ldcx(src + (N - k));
                    a = cx<T>(T(0.5) * (a.re + b.re), T(0.5) * (a.im - b.im));
                }
                v[k] = a;
            }
        } else {
            for (int k = t; k < NH; k += NT) v[k] = ldcx(prm.v + e * NH + k);
        }
        __syncthreads();
        F::rinv(v, buf, prm.tw);
        for (int j = t; j < H; j += NT) U[j] = cx<T>(buf[j].re * invN, buf[j].im * invN);
        __syncthreads();
    }
    if (MODE == AUX_GET_U) {
        Cx<T>* dst = static_cast<Cx<T>*>(dst_) + e * H;
        for (int j = t; j < H; j += NT) stcx(dst + j, U[j]);
        return;
    }
    if (equation == 0) {
        for (int j = t; j < H; j += NT) buf[j] = cx<T>(U[j].re * U[j].re, U[j].im * U[j].im);
        __syncthreads();
        F::rfwd(buf, X, T(0.5), prm.tw);
    }
    const float dxf = (float)prm.dx;
    const int64_t hrow = e * prm.hist_rows;
    for (int k = t; k < NH; k += NT) {
        stcx(prm.v + e * NH + k, v[k]);
        if (equation == 0) {
            const T kw = prm.kwave[k];
            stcx(prm.fn + e * NH + k, cx<T>(-kw * X[k].im, kw * X[k].re));
        }
        const float ek = ek_row_f32((float)v[k].re, (float)v[k].im, N, dxf);
        prm.acc[e * NH + k] = ek;
        if (prm.hist_rows > 0) {
            if (prm.vv_hist) {
                Cx<float> c; c.re = (float)v[k].re; c.im = (float)v[k].im;
                prm.vv_hist[hrow * N + k] = c;
                if (k != 0 && k != H) { c.im = -c.im; prm.vv_hist[hrow * N + N - k] = c; }
            }
            if (prm.ektt_hist) prm.ektt_hist[hrow * NH + k] = (double)ek;
        }
    }
    for (int j = t; j < H; j += NT) {
        stcx(reinterpret_cast<Cx<T>*>(prm.uprev + e * N) + j, U[j]);
        if (prm.hist_rows > 0 && prm.uu_hist) stcx(reinterpret_cast<Cx<T>*>(prm.uu_hist + hrow * N) + j, U[j]);
    }
    if (t == 0) {
        prm.iout[e] = 0;
        prm.tnow[e] = T(0);
        prm.kprev[e] = T(0);
        prm.status[e] = 0;
    }
}

}  // namespace mpde
