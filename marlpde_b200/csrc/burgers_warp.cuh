// Burgers environment step, warp-resident variant (N = 8..256), sm_100a.
//
// Restates the arithmetic of the reference Burger.step() + getState() + rewards
// (/root/reference/python/_model/Burger.py:333-499, 541-576, 578-675 and
// burger_environment.py:148-176) for a batch of independent environments:
//   * one team of min(N/2, 32) lanes owns one environment (N = 32: two environments per
//     warp); lane registers hold the half spectrum of v and Fn_old, one or a few
//     wavenumbers each, plus two adjacent grid points of u;
//   * the state is read once, `nsub` ABCN sub-steps run out of registers on a shuffle
//     real-FFT, then state / reward / spectrum sums are written back coalesced;
//   * action forcing, 3-mode stochastic forcing, Smagorinsky closures, the float32
//     spectrum chain and the spectral / MSE rewards are fused in.
// No arithmetic of one environment depends on another one: results are bitwise
// independent of batch size and packing.
#pragma once
#include "params.h"
#include "warp_fft.cuh"

namespace mpde {

// energy-spectrum row in the reference's float32 chain (Burger.py:562 on complex64 data)
__device__ __forceinline__ float ek_row_f32(float re, float im, int N, float dxf) {
    const float en = __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));
    return __fmul_rn(en * (0.5f / (float)N), dxf);
}

template <typename T, int N>
struct BurgersWarp {
    using R = RealFFT<T, N>;
    static constexpr int H = N / 2, TS = R::TS, P = R::P, NH = N / 2 + 1;
    static constexpr unsigned TEAM_MASK = TS == 32 ? 0xffffffffu : ((1u << TS) - 1u);

    __device__ __forceinline__ static bool team_any(const R& f, bool pred) {
        const unsigned b = __ballot_sync(0xffffffffu, pred);
        return ((b >> f.c.base) & TEAM_MASK) != 0u;
    }
    __device__ __forceinline__ static T team_sum(T x) {
#pragma unroll
        for (int h = TS / 2; h >= 1; h >>= 1) x += shfl_xor(x, h);
        return x;
    }
    // Real field stored as x[p] = (x_{2j}, x_{2j+1}), j = p*TS + tl.  Returns the left
    // neighbour of the even point (x_{2j-1}) and the right neighbour of the odd point
    // (x_{2j+2}), periodic.
    __device__ __forceinline__ static void halo(const R& f, const Cx<T> (&x)[P], T (&left)[P], T (&right)[P]) {
        const int tl = f.c.tl;
        const int lr = f.c.base + ((tl + 1) & (TS - 1));
        const int ll = f.c.base + ((tl - 1) & (TS - 1));
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const T a = shfl(x[p].re, lr), b = shfl(x[(p + 1) % P].re, lr);
            right[p] = (tl == TS - 1) ? b : a;
            const T c = shfl(x[p].im, ll), d = shfl(x[(p + P - 1) % P].im, ll);
            left[p] = (tl == 0) ? d : c;
        }
    }
    // float64 -> float32 -> float64 (quirk Q1: the forcing accumulator of the reference is
    // complex64 unless the stochastic forcing replaced it, Burger.py:335,466)
    __device__ __forceinline__ static T r32(T a) { return (T)(float)a; }

    __device__ static void run(const SpectralParams<T>& prm, T* smem) {
        const int lane = threadIdx.x & 31;
        const int warp = threadIdx.x >> 5;
        const int wpc = blockDim.x >> 5;
        constexpr int TPW = 32 / TS;
        const int64_t first = ((int64_t)blockIdx.x * wpc + warp) * TPW;
        if (first >= prm.B) return;                                  // whole warp idle
        R f;
        f.init(prm.tw);
        const int tl = f.c.tl;
        const int team = lane / TS;
        const int64_t e = first + team;
        const bool has = e < prm.B;
        const int64_t ec = has ? e : 0;
        const int flags = prm.flags;
        const bool q1 = !(flags & F_FORCING);
        const int scr = max(prm.M, 2 * N + N / 2);
        T* scratch = smem + (size_t)(warp * TPW + team) * scr;

        // ---- per-register constants -------------------------------------------------------
        const T dt = prm.dt;
        const T nu = prm.nu[ec];
        int kk[P];
        T kw[P], g1[P], g2[P], g3[P];       // (1-C)/(1+C), dt/(1+C), 1/(1+C); C = nu k^2 dt/2 (Burger.py:486-488)
#pragma unroll
        for (int p = 0; p < P; ++p) {
            kk[p] = f.k(p);
            kw[p] = prm.kwave[kk[p]];
            const T C = T(0.5) * (kw[p] * kw[p]) * nu * dt;
            g1[p] = (T(1) - C) / (T(1) + C);
            g2[p] = dt / (T(1) + C);
            g3[p] = T(1) / (T(1) + C);
        }
        const T kwN = prm.kwave[H];
        const T CN = T(0.5) * (kwN * kwN) * nu * dt;
        const T g1N = (T(1) - CN) / (T(1) + CN), g2N = dt / (T(1) + CN), g3N = T(1) / (T(1) + CN);

        // ---- load state ---------------------------------------------------------------------
        bool live = has && prm.status[ec] == 0;
        int iout = prm.iout[ec];
        T tnow = prm.tnow[ec];
        Cx<T> v[P], fn[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            v[p] = ldcx(prm.v + ec * NH + kk[p]);
            fn[p] = ldcx(prm.fn + ec * NH + kk[p]);
        }
        // Nyquist mode: complex in the reference when the IC came from a truncated DNS spectrum
        // (quirk Q5); its imaginary part and Im v[0] never reach u.  Carried by the dc lane.
        Cx<T> vN = ldcx(prm.v + ec * NH + H);
        T fnN = ldcx(prm.fn + ec * NH + H).im;
        const T v0im = v[0].im;                 // meaningful on the dc lane only

        // ---- u = Re ifft(v) -------------------------------------------------------------------
        const T invN = T(1) / T(N);
        Cx<T> u[P], uprev[P];
        f.inv(v, vN.re, u, invN);
#pragma unroll
        for (int p = 0; p < P; ++p) uprev[p] = u[p];
        if ((flags & F_NO_ADVANCE) && prm.version == 1 && iout > 0) {
#pragma unroll
            for (int p = 0; p < P; ++p)
                uprev[p] = ldcx(reinterpret_cast<const Cx<T>*>(prm.uprev + ec * N) + p * TS + tl);
        }

        // ---- action field a @ basis (Burger.py:442) --------------------------------------------
        Cx<T> fa[P], Fa[P];
        T FaN = T(0);
#pragma unroll
        for (int p = 0; p < P; ++p) { fa[p] = cx<T>(0, 0); Fa[p] = cx<T>(0, 0); }
        if (flags & F_ACTIONS) {
            for (int i = tl; i < prm.M; i += TS) scratch[i] = prm.actions[ec * prm.M + i];
            __syncwarp();
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
            __syncwarp();
            if (flags & F_DFORCE) {           // spectrum of a direct forcing is constant over the sub-steps
                Cx<T> z[P];
#pragma unroll
                for (int p = 0; p < P; ++p) z[p] = fa[p];
                f.fwd(z, Fa, FaN, T(1));
            }
        }

        // ---- reward bookkeeping ----------------------------------------------------------------
        float acc32[P];
#pragma unroll
        for (int p = 0; p < P; ++p) acc32[p] = prm.acc[ec * NH + kk[p]];
        float accN = prm.acc[ec * NH + H];
        const float dxf = (float)prm.dx, dtf = (float)dt;
        Cx<T> mse[P];
#pragma unroll
        for (int p = 0; p < P; ++p) mse[p] = cx<T>(0, 0);
        const T inv_dx = T(1) / prm.dx, inv_dx2 = T(1) / (prm.dx * prm.dx);
        const int64_t truth_base =
            prm.truth ? ((prm.truth_map ? prm.truth_map[ec] : 0) * prm.truth_rows) : 0;

        // =============================== sub-steps ==============================================
        const int nsub = (flags & F_NO_ADVANCE) ? 0 : prm.nsub;
        for (int it = 0; it < nsub; ++it) {
            // nonlinear term: X = fft(u^2 / 2) (Burger.py:487)
            Cx<T> z[P], X[P];
            T XN;
#pragma unroll
            for (int p = 0; p < P; ++p) z[p] = cx<T>(u[p].re * u[p].re, u[p].im * u[p].im);
            f.fwd(z, X, XN, T(0.5));

            Cx<T> Fh[P];
            T FhN = T(0);
#pragma unroll
            for (int p = 0; p < P; ++p) Fh[p] = cx<T>(0, 0);

            T left[P], right[P];
            const bool need_nb = (flags & (F_SSM | F_DSM)) || ((flags & F_ACTIONS) && !(flags & F_DFORCE));
            if (need_nb) halo(f, u, left, right);

            if (flags & (F_SSM | F_DSM)) {
                Cx<T> sgs[P];
                Cx<T> dudx[P], d2[P];
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    // upwind first difference, centred second difference (Burger.py:342-346)
                    dudx[p] = cx<T>((u[p].re - left[p]) * inv_dx, (u[p].im - u[p].re) * inv_dx);
                    d2[p] = cx<T>((u[p].im - T(2) * u[p].re + left[p]) * inv_dx2,
                                  (right[p] - T(2) * u[p].im + u[p].re) * inv_dx2);
                }
                if (flags & F_SSM) {
                    // Burger.py:339-349, delta = 2 pi / N whatever L is
                    const T cd = T(0.1) * T(2.0 * 3.14159265358979323846 / N);
                    const T cd2 = cd * cd;
#pragma unroll
                    for (int p = 0; p < P; ++p)
                        sgs[p] = cx<T>(cd2 * fabs(dudx[p].re) * d2[p].re, cd2 * fabs(dudx[p].im) * d2[p].im);
                } else {
                    // dynamic Smagorinsky, Burger.py:357-399 (Germano identity with a sharp spectral
                    // test filter |k| > N//4 applied IN PLACE to the state v, :369-370)
                    const T delta = T(2.0 * 3.14159265358979323846 / N), deltah = T(4.0 * 3.14159265358979323846 / N);
                    bool cut[P];
#pragma unroll
                    for (int p = 0; p < P; ++p) cut[p] = fabs(kw[p]) > T(N / 4);
                    const bool cutN = fabs(kwN) > T(N / 4);
                    Cx<T> w[P], L1[P], uh[P];
#pragma unroll
                    for (int p = 0; p < P; ++p)      // filtered fft(u^2) = 2 X
                        w[p] = cut[p] ? cx<T>(0, 0) : cx<T>(T(2) * X[p].re, T(2) * X[p].im);
                    f.inv(w, cutN ? T(0) : T(2) * XN, L1, T(0.5) * invN);
#pragma unroll
                    for (int p = 0; p < P; ++p)
                        if (cut[p]) v[p] = cx<T>(0, 0);
                    if (cutN) vN = cx<T>(0, 0);
                    f.inv(v, vN.re, uh, invN);
#pragma unroll
                    for (int p = 0; p < P; ++p)
                        z[p] = cx<T>(fabs(dudx[p].re) * dudx[p].re, fabs(dudx[p].im) * dudx[p].im);
                    Cx<T> W2[P], M1[P];
                    T W2N;
                    f.fwd(z, W2, W2N, T(1));
#pragma unroll
                    for (int p = 0; p < P; ++p)
                        if (cut[p]) W2[p] = cx<T>(0, 0);
                    f.inv(W2, cutN ? T(0) : W2N, M1, delta * delta * invN);
                    T uhl[P], uhr[P];
                    halo(f, uh, uhl, uhr);
                    Cx<T> malt[P];
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const T da = (uh[p].re - uhl[p]) * inv_dx, db = (uh[p].im - uh[p].re) * inv_dx;
                        const T M2a = deltah * deltah * fabs(da) * da, M2b = deltah * deltah * fabs(db) * db;
                        malt[p] = cx<T>(T(4) / (deltah * deltah) * M2a - T(1) / (delta * delta) * M1[p].re,
                                        T(4) / (deltah * deltah) * M2b - T(1) / (delta * delta) * M1[p].im);
                    }
                    T ml[P], mr[P];
                    halo(f, malt, ml, mr);
                    T num = T(0), den = T(0);
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const T Lga = L1[p].re - T(0.5) * uh[p].re * uh[p].re;
                        const T Lgb = L1[p].im - T(0.5) * uh[p].im * uh[p].im;
                        const T Mga = (malt[p].re - ml[p]) * inv_dx, Mgb = (malt[p].im - malt[p].re) * inv_dx;
                        num += -Lga * Mga - Lgb * Mgb;
                        den += Mga * Mga + Mgb * Mgb;
                    }
                    num = team_sum(num);
                    den = team_sum(den);
                    const T c = num / den;          // mean/mean: the 1/N cancels (Burger.py:397)
#pragma unroll
                    for (int p = 0; p < P; ++p)
                        sgs[p] = cx<T>(c * fabs(dudx[p].re) * d2[p].re, c * fabs(dudx[p].im) * d2[p].im);
                }
                Cx<T> S[P];
                T SN;
                f.fwd(sgs, S, SN, T(1));
#pragma unroll
                for (int p = 0; p < P; ++p) Fh[p] = q1 ? cx<T>(r32(S[p].re), r32(S[p].im)) : S[p];
                FhN = q1 ? r32(SN) : SN;
            }

            if (flags & F_FORCING) {
                // Burger.py:410-421: the spectrum of sum_k c_k cos(...) is non-zero at k = +-1,2,3 only;
                // it REPLACES whatever the closures accumulated (Q2).  Column ioutnum % stepper (Q3).
                const int64_t row = (flags & F_FORCING_PER_ENV) ? ec : 0;
                const int col = iout % prm.stepper;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    Cx<T> c = cx<T>(0, 0);
                    if (kk[p] >= 1 && kk[p] <= 3) c = ldcx(prm.fcoef + (row * prm.stepper + col) * 3 + (kk[p] - 1));
                    Fh[p] = c;
                }
                FhN = T(0);
            }

            if (flags & F_ACTIONS) {
                Cx<T> S[P];
                T SN;
                if (flags & F_DFORCE) {
#pragma unroll
                    for (int p = 0; p < P; ++p) S[p] = Fa[p];
                    SN = FaN;
                } else {
                    // eddy-viscosity action: forcing = (a @ basis) * d2u/dx2 (Burger.py:445-450)
#pragma unroll
                    for (int p = 0; p < P; ++p)
                        z[p] = cx<T>(fa[p].re * ((left[p] - T(2) * u[p].re + u[p].im) * inv_dx2),
                                     fa[p].im * ((u[p].re - T(2) * u[p].im + right[p]) * inv_dx2));
                    f.fwd(z, S, SN, T(1));
                }
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const Cx<T> t = Fh[p] + S[p];
                    Fh[p] = q1 ? cx<T>(r32(t.re), r32(t.im)) : t;
                }
                FhN = q1 ? r32(FhN + SN) : FhN + SN;
            }

            // ABCN update (Burger.py:486-489): v <- ((1-C) v - dt/2 (3 Fn - Fn_old) + dt F) / (1+C).
            // Q1 (cont.): while the forcing accumulator is complex64, `self.dt*Fforcing` is a
            // complex64 product: float32(dt) * float32(F), rounded to float32.
            Cx<T> vn[P], fnn[P];
            bool bad = false;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                fnn[p] = cx<T>(-kw[p] * X[p].im, kw[p] * X[p].re);                       // i k X
                const Cx<T> dtF = q1 ? cx<T>((T)__fmul_rn(dtf, (float)Fh[p].re), (T)__fmul_rn(dtf, (float)Fh[p].im))
                                     : cx<T>(dt * Fh[p].re, dt * Fh[p].im);
                const T tr = T(0.5) * fn[p].re - T(1.5) * fnn[p].re;
                const T ti = T(0.5) * fn[p].im - T(1.5) * fnn[p].im;
                vn[p] = cx<T>(fma(g3[p], dtF.re, fma(g2[p], tr, g1[p] * v[p].re)),
                              fma(g3[p], dtF.im, fma(g2[p], ti, g1[p] * v[p].im)));
                bad |= blown(vn[p]);
            }
            // k = 0: Fn = 0, F real -> Im v[0] is a constant of the motion; Nyquist: F real, Fn imaginary
            const T fnnN = kwN * XN;
            const T dtFN = q1 ? (T)__fmul_rn(dtf, (float)FhN) : dt * FhN;
            const Cx<T> vnN = cx<T>(fma(g3N, dtFN, g1N * vN.re), fma(g2N, T(0.5) * fnN - T(1.5) * fnnN, g1N * vN.im));
            if (f.dc) {
                vn[0].im = v0im;
                bad |= blown(vnN);
            }
            bad = team_any(f, bad);
            if (live && bad) { prm.status[e] = 1; live = false; }
            if (live) {
#pragma unroll
                for (int p = 0; p < P; ++p) { v[p] = vn[p]; fn[p] = fnn[p]; uprev[p] = u[p]; }
                vN = vnN;
                fnN = fnnN;
                iout += 1;
                tnow += dt;
            }
            // u = Re ifft(v) (Burger.py:491); a blown-up env keeps its last good field
            {
                Cx<T> un[P];
                f.inv(v, vN.re, un, invN);
                if (live) {
#pragma unroll
                    for (int p = 0; p < P; ++p) u[p] = un[p];
                }
            }

            if (live) {
                // float32 spectrum chain (Q6): Ek row from complex64(v), sequential float32 sum
#pragma unroll
                for (int p = 0; p < P; ++p)
                    acc32[p] = __fadd_rn(acc32[p], ek_row_f32((float)v[p].re, (float)v[p].im, N, dxf));
                accN = __fadd_rn(accN, ek_row_f32((float)vN.re, (float)vN.im, N, dxf));

                if (prm.hist_rows > 0 && iout < prm.hist_rows) {
                    const int64_t hrow = e * prm.hist_rows + iout;
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const int j = p * TS + tl;
                        if (prm.uu_hist) stcx(reinterpret_cast<Cx<T>*>(prm.uu_hist + hrow * N) + j, u[p]);
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

                if (prm.reward_mode == REWARD_MSE && prm.truth) {
                    // Burger.py:589-599 after every sub-step, averaged over them (burger_environment.py:153)
                    const int64_t row = iout < prm.truth_rows ? iout : prm.truth_rows - 1;
                    const Cx<T>* tr = reinterpret_cast<const Cx<T>*>(prm.truth + (truth_base + row) * N);
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const Cx<T> t = ldcx(tr + p * TS + tl);
                        const T da = t.re - u[p].re, db = t.im - u[p].im;
                        mse[p].re += da * da;
                        mse[p].im += db * db;
                    }
                }
            }
        }

        // =============================== epilogue ===============================================
        if (nsub > 0 && live) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                stcx(prm.v + e * NH + kk[p], v[p]);
                stcx(prm.fn + e * NH + kk[p], fn[p]);
                prm.acc[e * NH + kk[p]] = acc32[p];
                stcx(reinterpret_cast<Cx<T>*>(prm.uprev + e * N) + p * TS + tl, uprev[p]);
            }
            if (f.dc) {
                stcx(prm.v + e * NH + H, vN);
                stcx(prm.fn + e * NH + H, cx<T>(T(0), fnN));
                prm.acc[e * NH + H] = accN;
                prm.iout[e] = iout;
                prm.tnow[e] = tnow;
            }
        }

        const T inf = T(1) / T(0);
        if (prm.state_out) {
            // getState (Burger.py:604-675) through a shared-memory gather so that every
            // version / agent-window layout becomes one coalesced row store
            const int ver = prm.version, A = prm.A;
            T left[P], right[P];
            halo(f, u, left, right);
            __syncwarp();
            T* f0 = scratch;
            T* f1 = f0 + N;
            T* ek = f1 + N;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const int j = p * TS + tl;
                const Cx<T> d2 = cx<T>((left[p] - T(2) * u[p].re + u[p].im) * inv_dx2,
                                       (u[p].re - T(2) * u[p].im + right[p]) * inv_dx2);
                const Cx<T> dudt = cx<T>((u[p].re - uprev[p].re) / dt, (u[p].im - uprev[p].im) / dt);
                Cx<T> a = d2, b = d2;
                if (ver == 1) { a = dudt; }
                else if (ver == 2) { a = u[p]; b = cx<T>(u[p].re * u[p].re, u[p].im * u[p].im); }
                else if (ver == 4) { a = u[p]; }
                stcx(reinterpret_cast<Cx<T>*>(f0) + j, a);
                stcx(reinterpret_cast<Cx<T>*>(f1) + j, b);
                // Burger.py:653, from the live float64 v
                ek[kk[p]] = T(0.5) * ((v[p].re * v[p].re + v[p].im * v[p].im) / T(N)) * prm.dx;
            }
            __syncwarp();
            const int nf = (ver == 1 || ver == 2) ? 2 : 1;
            const int seg = A == 1 ? N : N / A + 2;
            const int tail = (ver == 3 || ver == 4) ? N / 2 : 0;
            const int RL = nf * seg + tail;
            const int S = A * RL;
            if (has) {
                for (int o = tl; o < S; o += TS) {
                    const int a = o / RL, r = o - a * RL;
                    T val;
                    if (r < nf * seg) {
                        const int fld = r / seg, w = r - fld * seg;
                        const int start = A == 1 ? 0 : a * (N / A) - 1;
                        const int j = (start + w + N) & (N - 1);
                        val = f0[fld * N + j];
                    } else {
                        val = f0[2 * N + (r - nf * seg)];
                    }
                    prm.state_out[e * S + o] = live ? val : inf;          // Burger.py:633-643
                }
            }
            __syncwarp();
        }

        if (prm.reward_out && prm.reward_mode == REWARD_SPECTRAL && nsub > 0) {
            // burger_environment.py:172-176 on the running float32 sums
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
            part = team_sum(part) / T(H - 1);
            const T prev = prm.kprev[ec];
            const T r = live ? prev - part : -inf;
            if (has) {
                for (int a = tl; a < A; a += TS) prm.reward_out[e * A + a] = r;
                if (f.dc && live) prm.kprev[e] = part;
            }
        }
        if (prm.reward_out && prm.reward_mode == REWARD_MSE && prm.truth) {
            const int A = prm.A, W = N / A;
            if (nsub == 0 && live) {     // getMseReward() of the current state (Burger.py:578-601)
                const int64_t row = iout < prm.truth_rows ? iout : prm.truth_rows - 1;
                const Cx<T>* tr = reinterpret_cast<const Cx<T>*>(prm.truth + (truth_base + row) * N);
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const Cx<T> t = ldcx(tr + p * TS + tl);
                    const T da = t.re - u[p].re, db = t.im - u[p].im;
                    mse[p] = cx<T>(da * da, db * db);
                }
            }
            __syncwarp();
#pragma unroll
            for (int p = 0; p < P; ++p) stcx(reinterpret_cast<Cx<T>*>(scratch) + p * TS + tl, mse[p]);
            __syncwarp();
            if (has) {
                for (int a = tl; a < A; a += TS) {       // agent a owns points [a N/A, (a+1) N/A)
                    T sum = T(0);
                    for (int j = 0; j < W; ++j) sum += scratch[a * W + j];
                    prm.reward_out[e * A + a] = live ? -(sum / T(W)) / T(nsub > 0 ? nsub : 1) : -inf;
                }
            }
        }
    }
};

}  // namespace mpde
