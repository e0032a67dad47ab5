// Burgers environment step, warp-resident variant (N = 4..128), sm_100a.
//
// Restates the arithmetic of the reference Burger.step() + getState() + rewards
// (/root/reference/python/_model/Burger.py:333-499, 541-576, 578-675 and
// burger_environment.py:148-176) for a batch of independent environments:
//   * one team of min(N,32) lanes carries a PAIR of environments (A in the real part of
//     every complex transform, B in the imaginary part),
//   * the half-spectrum state (v, Fn_old) is read once, `nsub` ABCN sub-steps run out of
//     registers with a shuffle FFT, then state / reward / spectrum sums are written back,
//   * the action forcing, the 3-mode stochastic forcing, the Smagorinsky closures, the
//     float32 spectrum chain and the spectral / MSE rewards are fused in.
#pragma once
#include "params.h"
#include "warp_fft.cuh"

namespace mpde {

template <typename T, int N>
struct BurgersWarp {
    using F = WarpFFT<T, N>;
    static constexpr int TS = F::TS, P = F::P, NH = N / 2 + 1;
    static constexpr unsigned TEAM_MASK = TS == 32 ? 0xffffffffu : ((1u << TS) - 1u);

    __device__ __forceinline__ static bool team_any(const F& f, bool pred) {
        const unsigned b = __ballot_sync(0xffffffffu, pred);
        return ((b >> f.base) & TEAM_MASK) != 0u;
    }
    __device__ __forceinline__ static T team_sum(T x) {
#pragma unroll
        for (int h = TS / 2; h >= 1; h >>= 1) x += shfl_xor(x, h);
        return x;
    }
    // periodic neighbours of a real field held as n = p*TS + tl
    __device__ __forceinline__ static void neighbours(const F& f, const T (&u)[P], T (&um)[P], T (&up)[P]) {
        const int lr = f.base + ((f.tl + 1) & (TS - 1));
        const int ll = f.base + ((f.tl - 1) & (TS - 1));
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const T a = shfl(u[p], lr), b = shfl(u[(p + 1) % P], lr);
            up[p] = (f.tl == TS - 1) ? b : a;
            const T c = shfl(u[p], ll), d = shfl(u[(p + P - 1) % P], ll);
            um[p] = (f.tl == 0) ? d : c;
        }
    }
    // complex128 -> complex64 -> complex128 (quirk Q1: the forcing accumulator of the
    // reference is complex64 unless the stochastic forcing replaced it, Burger.py:335,466)
    __device__ __forceinline__ static Cx<T> round_c64(Cx<T> a) { return cx<T>((T)(float)a.re, (T)(float)a.im); }

    __device__ static void run(const SpectralParams<T>& prm, T* smem) {
        const int lane = threadIdx.x & 31;
        const int warp = threadIdx.x >> 5;
        const int wpc = blockDim.x >> 5;
        constexpr int TPW = 32 / TS;
        F f;
        f.init(prm.tw);
        const int team = lane / TS;
        const int64_t pair = ((int64_t)blockIdx.x * wpc + warp) * TPW + team;
        const int64_t e[2] = {2 * pair, 2 * pair + 1};
        if (2 * (((int64_t)blockIdx.x * wpc + warp) * TPW) >= prm.B) return;   // whole warp idle
        const int flags = prm.flags;
        const bool q1 = !(flags & F_FORCING);
        const int scr = max(prm.M, 2 * N + N / 2);
        T* scratch = smem + ((size_t)(warp * TPW + team) * 2) * scr;   // [2][scr]

        // ---- per-position constants ---------------------------------------------------
        int kh[P];            // index into the stored half spectrum
        bool cj[P], selfc[P], owner[P];
        T kw[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int k = F::kidx(p, f.tl);
            kh[p] = k <= N / 2 ? k : N - k;
            cj[p] = k > N / 2;
            selfc[p] = (k == 0) || (k == N / 2);
            owner[p] = k <= N / 2;
            kw[p] = prm.kwave[k];
        }

        // ---- load state ---------------------------------------------------------------
        bool live[2];
        Cx<T> v[2][P], fn[2][P];
        T g1[2][P], g2[2][P];         // (1-C)/(1+C) and dt/(1+C), C = nu k^2 dt / 2 (Burger.py:486-488)
        int iout[2];
        T tnow[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const bool has = e[s] < prm.B;
            live[s] = has && prm.status[has ? e[s] : 0] == 0;
            iout[s] = has ? prm.iout[e[s]] : 0;
            tnow[s] = has ? prm.tnow[e[s]] : T(0);
            const T nu = has ? prm.nu[e[s]] : T(0);
#pragma unroll
            for (int p = 0; p < P; ++p) {
                Cx<T> a = cx<T>(T(0), T(0)), b = a;
                if (has) {
                    a = ldcx(prm.v + e[s] * NH + kh[p]);
                    b = ldcx(prm.fn + e[s] * NH + kh[p]);
                    if (cj[p]) { a = conj(a); b = conj(b); }
                }
                v[s][p] = a;
                fn[s][p] = b;
                const T C = T(0.5) * (kw[p] * kw[p]) * nu * prm.dt;
                g1[s][p] = (T(1) - C) / (T(1) + C);
                g2[s][p] = prm.dt / (T(1) + C);
            }
        }

        // ---- u = Re ifft(v) -------------------------------------------------------------
        const T invN = T(1) / T(N);
        T u[2][P], uprev[2][P];
        {
            Cx<T> z[P];
#pragma unroll
            for (int p = 0; p < P; ++p)
                z[p] = F::tangle(live[0] ? v[0][p] : cx<T>(0, 0), live[1] ? v[1][p] : cx<T>(0, 0), selfc[p]);
            f.inv(z);
#pragma unroll
            for (int p = 0; p < P; ++p) {
                u[0][p] = z[p].re * invN;
                u[1][p] = z[p].im * invN;
                uprev[0][p] = u[0][p];
                uprev[1][p] = u[1][p];
            }
        }
        if ((flags & F_NO_ADVANCE) && prm.version == 1 && prm.uprev) {
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int p = 0; p < P; ++p)
                    if (e[s] < prm.B && iout[s] > 0) uprev[s][p] = prm.uprev[e[s] * N + p * TS + f.tl];
        }

        // ---- action field a @ basis (Burger.py:442) ---------------------------------------
        T fa[2][P];
        Cx<T> Fa[2][P];          // its spectrum when the forcing is direct (constant over the sub-steps)
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int p = 0; p < P; ++p) { fa[s][p] = T(0); Fa[s][p] = cx<T>(0, 0); }
        if (flags & F_ACTIONS) {
#pragma unroll
            for (int s = 0; s < 2; ++s)
                for (int i = f.tl; i < prm.M; i += TS)
                    scratch[s * scr + i] = e[s] < prm.B ? prm.actions[e[s] * prm.M + i] : T(0);
            __syncwarp();
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const int n = p * TS + f.tl;
                    T acc = T(0);
                    if (flags & F_BASIS_DENSE) {
                        for (int i = 0; i < prm.M; ++i) acc = fma(scratch[s * scr + i], prm.basis[(size_t)i * N + n], acc);
                    } else {
                        acc = prm.tap_w[2 * n] * scratch[s * scr + prm.tap_idx[2 * n]] +
                              prm.tap_w[2 * n + 1] * scratch[s * scr + prm.tap_idx[2 * n + 1]];
                    }
                    fa[s][p] = acc;
                }
            __syncwarp();
            // a non-finite action would poison the partner env through the packed FFT
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                bool bad = false;
#pragma unroll
                for (int p = 0; p < P; ++p) bad |= !(fabs((double)fa[s][p]) <= 1e150);
                if (team_any(f, bad)) {
                    if (live[s]) prm.status[e[s]] = 1;
                    live[s] = false;
#pragma unroll
                    for (int p = 0; p < P; ++p) fa[s][p] = T(0);
                }
            }
            if (flags & F_DFORCE) {
                Cx<T> z[P];
#pragma unroll
                for (int p = 0; p < P; ++p) z[p] = cx<T>(fa[0][p], fa[1][p]);
                f.fwd(z);
                f.untangle(z, Fa[0], Fa[1], T(1));
            }
        }

        // ---- rewards bookkeeping ------------------------------------------------------------
        float acc32[2][P];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int p = 0; p < P; ++p)
                acc32[s][p] = (e[s] < prm.B && owner[p]) ? prm.acc[e[s] * NH + kh[p]] : 0.f;
        const float dxf = (float)prm.dx;
        T mse[2][P];                       // per-lane partial of the segment means
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int p = 0; p < P; ++p) mse[s][p] = T(0);
        const T inv_dx = T(1) / prm.dx, inv_dx2 = T(1) / (prm.dx * prm.dx);

        // =========================== sub-steps ============================================
        const int nsub = (flags & F_NO_ADVANCE) ? 0 : prm.nsub;
        for (int it = 0; it < nsub; ++it) {
            // nonlinear term: X = fft(u^2 / 2)
            Cx<T> z[P], X[2][P];
#pragma unroll
            for (int p = 0; p < P; ++p)
                z[p] = cx<T>(live[0] ? u[0][p] * u[0][p] : T(0), live[1] ? u[1][p] * u[1][p] : T(0));
            f.fwd(z);
            f.untangle(z, X[0], X[1], T(0.5));

            Cx<T> Fh[2][P];
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int p = 0; p < P; ++p) Fh[s][p] = cx<T>(0, 0);

            T um[2][P], up[2][P];
            const bool need_nb = flags & (F_SSM | F_DSM) || ((flags & F_ACTIONS) && !(flags & F_DFORCE));
            if (need_nb) {
                neighbours(f, u[0], um[0], up[0]);
                neighbours(f, u[1], um[1], up[1]);
            }

            if (flags & (F_SSM | F_DSM)) {
                T sgs[2][P];
                if (flags & F_SSM) {
                    // Burger.py:339-349, delta = 2 pi / N whatever L is
                    const T cd = T(0.1) * T(2.0 * 3.14159265358979323846 / N);
                    const T cd2 = cd * cd;
#pragma unroll
                    for (int s = 0; s < 2; ++s)
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            const T dudx = (u[s][p] - um[s][p]) * inv_dx;
                            const T d2 = (up[s][p] - T(2) * u[s][p] + um[s][p]) * inv_dx2;
                            sgs[s][p] = live[s] ? cd2 * fabs(dudx) * d2 : T(0);
                        }
                } else {
                    // dynamic Smagorinsky, Burger.py:357-399 (Germano identity with a sharp
                    // spectral test filter |k| > N//4 applied IN PLACE to the state v, :369-370)
                    const T delta = T(2.0 * 3.14159265358979323846 / N), deltah = T(4.0 * 3.14159265358979323846 / N);
                    bool cut[P];
#pragma unroll
                    for (int p = 0; p < P; ++p) cut[p] = fabs(kw[p]) > T(N / 4);
                    Cx<T> w[P];
                    T L1[2][P], uh[2][P];
#pragma unroll
                    for (int p = 0; p < P; ++p) {      // v2h: filtered fft(u^2) = 2 X
                        const Cx<T> a = cut[p] ? cx<T>(0, 0) : cx<T>(T(2) * X[0][p].re, T(2) * X[0][p].im);
                        const Cx<T> b = cut[p] ? cx<T>(0, 0) : cx<T>(T(2) * X[1][p].re, T(2) * X[1][p].im);
                        w[p] = F::tangle(a, b, selfc[p]);
                    }
                    f.inv(w);
#pragma unroll
                    for (int p = 0; p < P; ++p) { L1[0][p] = T(0.5) * w[p].re * invN; L1[1][p] = T(0.5) * w[p].im * invN; }
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        if (cut[p]) { v[0][p] = cx<T>(0, 0); v[1][p] = cx<T>(0, 0); }
                        w[p] = F::tangle(live[0] ? v[0][p] : cx<T>(0, 0), live[1] ? v[1][p] : cx<T>(0, 0), selfc[p]);
                    }
                    f.inv(w);
#pragma unroll
                    for (int p = 0; p < P; ++p) { uh[0][p] = w[p].re * invN; uh[1][p] = w[p].im * invN; }
                    T dudx[2][P], d2[2][P];
#pragma unroll
                    for (int s = 0; s < 2; ++s)
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            dudx[s][p] = (u[s][p] - um[s][p]) * inv_dx;
                            d2[s][p] = (up[s][p] - T(2) * u[s][p] + um[s][p]) * inv_dx2;
                        }
#pragma unroll
                    for (int p = 0; p < P; ++p) w[p] = cx<T>(fabs(dudx[0][p]) * dudx[0][p], fabs(dudx[1][p]) * dudx[1][p]);
                    f.fwd(w);
                    Cx<T> W2[2][P];
                    f.untangle(w, W2[0], W2[1], T(1));
#pragma unroll
                    for (int p = 0; p < P; ++p)
                        w[p] = cut[p] ? cx<T>(0, 0) : F::tangle(W2[0][p], W2[1][p], selfc[p]);
                    f.inv(w);
                    T uhm[2][P], uhp[2][P], malt[2][P], maltm[2][P], maltp[2][P];
                    neighbours(f, uh[0], uhm[0], uhp[0]);
                    neighbours(f, uh[1], uhm[1], uhp[1]);
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const T M1a = delta * delta * w[p].re * invN, M1b = delta * delta * w[p].im * invN;
                        const T da = (uh[0][p] - uhm[0][p]) * inv_dx, db = (uh[1][p] - uhm[1][p]) * inv_dx;
                        const T M2a = deltah * deltah * fabs(da) * da, M2b = deltah * deltah * fabs(db) * db;
                        malt[0][p] = T(4) / (deltah * deltah) * M2a - T(1) / (delta * delta) * M1a;
                        malt[1][p] = T(4) / (deltah * deltah) * M2b - T(1) / (delta * delta) * M1b;
                    }
                    neighbours(f, malt[0], maltm[0], maltp[0]);
                    neighbours(f, malt[1], maltm[1], maltp[1]);
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        T num = T(0), den = T(0);
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            const T Lg = L1[s][p] - T(0.5) * uh[s][p] * uh[s][p];
                            const T Mg = (malt[s][p] - maltm[s][p]) * inv_dx;
                            num += -Lg * Mg;
                            den += Mg * Mg;
                        }
                        num = team_sum(num);
                        den = team_sum(den);
                        const T c = num / den;          // mean/mean: the 1/N cancels (Burger.py:397)
                        if (live[s] && !(fabs((double)c) <= 1e150)) { prm.status[e[s]] = 1; live[s] = false; }
#pragma unroll
                        for (int p = 0; p < P; ++p) sgs[s][p] = live[s] ? c * fabs(dudx[s][p]) * d2[s][p] : T(0);
                    }
                }
#pragma unroll
                for (int p = 0; p < P; ++p) z[p] = cx<T>(sgs[0][p], sgs[1][p]);
                f.fwd(z);
                Cx<T> S[2][P];
                f.untangle(z, S[0], S[1], T(1));
#pragma unroll
                for (int s = 0; s < 2; ++s)
#pragma unroll
                    for (int p = 0; p < P; ++p) Fh[s][p] = q1 ? round_c64(S[s][p]) : S[s][p];
            }

            if (flags & F_FORCING) {
                // Burger.py:410-421: the spectrum of sum_k c_k cos(...) is non-zero at k = +-1,2,3 only;
                // it REPLACES whatever the closures accumulated (Q2).  Column index ioutnum % stepper (Q3).
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const int64_t row = (flags & F_FORCING_PER_ENV) ? (e[s] < prm.B ? e[s] : 0) : 0;
                    const int col = iout[s] % prm.stepper;
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const int k = F::kidx(p, f.tl);
                        const int ka = cj[p] ? N - k : k;
                        Cx<T> c = cx<T>(0, 0);
                        if (ka >= 1 && ka <= 3 && N >= 8) {
                            c = ldcx(prm.fcoef + (row * prm.stepper + col) * 3 + (ka - 1));
                            if (cj[p]) c = conj(c);
                        }
                        Fh[s][p] = c;
                    }
                }
            }

            if (flags & F_ACTIONS) {
                if (flags & F_DFORCE) {
#pragma unroll
                    for (int s = 0; s < 2; ++s)
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            const Cx<T> t = Fh[s][p] + Fa[s][p];
                            Fh[s][p] = q1 ? round_c64(t) : t;
                        }
                } else {
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const T d2a = (um[0][p] - T(2) * u[0][p] + up[0][p]) * inv_dx2;
                        const T d2b = (um[1][p] - T(2) * u[1][p] + up[1][p]) * inv_dx2;
                        z[p] = cx<T>(live[0] ? fa[0][p] * d2a : T(0), live[1] ? fa[1][p] * d2b : T(0));
                    }
                    f.fwd(z);
                    Cx<T> S[2][P];
                    f.untangle(z, S[0], S[1], T(1));
#pragma unroll
                    for (int s = 0; s < 2; ++s)
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            const Cx<T> t = Fh[s][p] + S[s][p];
                            Fh[s][p] = q1 ? round_c64(t) : t;
                        }
                }
            }

            // ABCN update (Burger.py:486-489): v <- ((1-C) v - dt/2 (3 Fn - Fn_old) + dt F) / (1+C)
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                Cx<T> vn[P], fnn[P];
                bool bad = false;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    fnn[p] = cx<T>(-kw[p] * X[s][p].im, kw[p] * X[s][p].re);             // i k X
                    const T tr = Fh[s][p].re - T(1.5) * fnn[p].re + T(0.5) * fn[s][p].re;
                    const T ti = Fh[s][p].im - T(1.5) * fnn[p].im + T(0.5) * fn[s][p].im;
                    vn[p] = cx<T>(fma(g2[s][p], tr, g1[s][p] * v[s][p].re), fma(g2[s][p], ti, g1[s][p] * v[s][p].im));
                    bad |= blown(vn[p]);
                }
                bad = team_any(f, bad);
                if (live[s] && bad) { prm.status[e[s]] = 1; live[s] = false; }
                if (live[s]) {
#pragma unroll
                    for (int p = 0; p < P; ++p) { v[s][p] = vn[p]; fn[s][p] = fnn[p]; }
                    iout[s] += 1;
                    tnow[s] += prm.dt;
                }
            }

            // u = Re ifft(v) (Burger.py:491)
#pragma unroll
            for (int p = 0; p < P; ++p)
                z[p] = F::tangle(live[0] ? v[0][p] : cx<T>(0, 0), live[1] ? v[1][p] : cx<T>(0, 0), selfc[p]);
            f.inv(z);
#pragma unroll
            for (int p = 0; p < P; ++p) {
                if (live[0]) { uprev[0][p] = u[0][p]; u[0][p] = z[p].re * invN; }
                if (live[1]) { uprev[1][p] = u[1][p]; u[1][p] = z[p].im * invN; }
            }

            // float32 spectrum chain (Q6): Ek row from complex64(v), sequential float32 sum
#pragma unroll
            for (int s = 0; s < 2; ++s)
                if (live[s]) {
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const float re = (float)v[s][p].re, im = (float)v[s][p].im;
                        const float en = __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));
                        acc32[s][p] = __fadd_rn(acc32[s][p], __fmul_rn(en * (0.5f / (float)N), dxf));
                    }
                }

            if (prm.hist_rows > 0) {
#pragma unroll
                for (int s = 0; s < 2; ++s)
                    if (live[s] && iout[s] < prm.hist_rows) {
                        const int64_t hrow = e[s] * prm.hist_rows + iout[s];
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            if (prm.uu_hist) prm.uu_hist[hrow * N + p * TS + f.tl] = u[s][p];
                            if (prm.vv_hist) {
                                Cx<float> c; c.re = (float)v[s][p].re; c.im = (float)v[s][p].im;
                                prm.vv_hist[hrow * N + F::kidx(p, f.tl)] = c;
                            }
                            if (prm.ektt_hist && owner[p])
                                prm.ektt_hist[hrow * NH + kh[p]] = (double)acc32[s][p] / (double)(iout[s] + 1);
                        }
                    }
            }

            if (prm.reward_mode == REWARD_MSE) {
                // Burger.py:589-599 evaluated after every sub-step, averaged over the sub-steps
                // (burger_environment.py:153)
#pragma unroll
                for (int s = 0; s < 2; ++s)
                    if (live[s]) {
                        const int64_t tr = prm.truth_map ? prm.truth_map[e[s]] : 0;
                        const int64_t row = iout[s] < prm.truth_rows ? iout[s] : prm.truth_rows - 1;
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            const T d = prm.truth[(tr * prm.truth_rows + row) * N + p * TS + f.tl] - u[s][p];
                            mse[s][p] += d * d;
                        }
                    }
            }
        }

        // =========================== epilogue ===============================================
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (e[s] >= prm.B) continue;
            const bool ok = live[s];
            if (nsub > 0 && ok) {
#pragma unroll
                for (int p = 0; p < P; ++p)
                    if (owner[p]) {
                        stcx(prm.v + e[s] * NH + kh[p], v[s][p]);
                        stcx(prm.fn + e[s] * NH + kh[p], fn[s][p]);
                        prm.acc[e[s] * NH + kh[p]] = acc32[s][p];
                    }
                if (prm.uprev) {
#pragma unroll
                    for (int p = 0; p < P; ++p) prm.uprev[e[s] * N + p * TS + f.tl] = uprev[s][p];
                }
                if (f.tl == 0) { prm.iout[e[s]] = iout[s]; prm.tnow[e[s]] = tnow[s]; }
            }
        }

        const T inf = T(1) / T(0);
        if (prm.state_out) {
            // getState (Burger.py:604-675) through a shared-memory gather so that every
            // version / agent-window layout becomes one coalesced row store
            const int ver = prm.version, A = prm.A;
            T um[2][P], up[2][P];
            neighbours(f, u[0], um[0], up[0]);
            neighbours(f, u[1], um[1], up[1]);
            __syncwarp();
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                T* f0 = scratch + s * scr;
                T* f1 = f0 + N;
                T* ek = f1 + N;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const int n = p * TS + f.tl;
                    const T d2 = (um[s][p] - T(2) * u[s][p] + up[s][p]) * inv_dx2;
                    const T dudt = (u[s][p] - uprev[s][p]) / prm.dt;
                    T a = d2, b = d2;
                    if (ver == 1) { a = dudt; b = d2; }
                    else if (ver == 2) { a = u[s][p]; b = u[s][p] * u[s][p]; }
                    else if (ver == 4) { a = u[s][p]; }
                    f0[n] = a;
                    f1[n] = b;
                    const int k = F::kidx(p, f.tl);
                    if (k < N / 2)    // Burger.py:653, from the live float64 v
                        ek[k] = T(0.5) * ((v[s][p].re * v[s][p].re + v[s][p].im * v[s][p].im) / T(N)) * prm.dx;
                }
            }
            __syncwarp();
            const int nf = (ver == 1 || ver == 2) ? 2 : 1;
            const int seg = A == 1 ? N : N / A + 2;
            const int tail = (ver == 3 || ver == 4) ? N / 2 : 0;
            const int RL = nf * seg + tail;
            const int S = A * RL;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (e[s] >= prm.B) continue;
                const T* f0 = scratch + s * scr;
                for (int o = f.tl; o < S; o += TS) {
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
                    prm.state_out[e[s] * S + o] = live[s] ? val : inf;     // Burger.py:633-643
                }
            }
        }

        if (prm.reward_out && nsub == 0 && prm.reward_mode == REWARD_MSE) {
            // getMseReward() of the current state (Burger.py:578-601) without stepping
#pragma unroll
            for (int s = 0; s < 2; ++s)
                if (live[s]) {
                    const int64_t tr = prm.truth_map ? prm.truth_map[e[s]] : 0;
                    const int64_t row = iout[s] < prm.truth_rows ? iout[s] : prm.truth_rows - 1;
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const T d = prm.truth[(tr * prm.truth_rows + row) * N + p * TS + f.tl] - u[s][p];
                        mse[s][p] = d * d;
                    }
                }
        }
        if (prm.reward_out && (nsub > 0 || prm.reward_mode == REWARD_MSE)) {
            const int A = prm.A;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const bool has = e[s] < prm.B;       // no early-out: the team shuffles below need every lane
                if (prm.reward_mode == REWARD_SPECTRAL && nsub > 0) {
                    // burger_environment.py:172-176 on the running float32 sums
                    const int64_t ref = (prm.ek_map && has) ? prm.ek_map[e[s]] : 0;
                    const int64_t row = iout[s] < prm.ek_rows ? iout[s] : prm.ek_rows - 1;
                    T part = T(0);
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const int k = F::kidx(p, f.tl);
                        if (k >= 1 && k < N / 2) {
                            const T ed = (T)prm.ek_ref[(ref * prm.ek_rows + row) * (N / 2) + k];
                            const T es = (T)((double)acc32[s][p] / (double)(iout[s] + 1));
                            const T q = fabs(ed - es) / ed;
                            part += q * q;
                        }
                    }
                    part = team_sum(part) / T(N / 2 - 1);
                    const T prev = has ? prm.kprev[e[s]] : T(0);
                    const T r = live[s] ? prev - part : -inf;
                    __syncwarp();
                    if (f.tl == 0 && live[s]) prm.kprev[e[s]] = part;
                    if (has)
                        for (int a = f.tl; a < A; a += TS) prm.reward_out[e[s] * A + a] = r;
                } else if (prm.reward_mode == REWARD_MSE) {
                    // segment means: agent a owns points [a N/A, (a+1) N/A)
                    const int W = N / A;
                    T* buf = scratch + s * scr;
                    __syncwarp();
#pragma unroll
                    for (int p = 0; p < P; ++p) buf[p * TS + f.tl] = mse[s][p];
                    __syncwarp();
                    for (int a = f.tl; a < A && has; a += TS) {
                        T sum = T(0);
                        for (int j = 0; j < W; ++j) sum += buf[a * W + j];
                        prm.reward_out[e[s] * A + a] = live[s] ? -(sum / T(W)) / T(nsub > 0 ? nsub : 1) : -inf;
                    }
                }
            }
        }
    }
};

}  // namespace mpde
