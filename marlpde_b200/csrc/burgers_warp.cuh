// Burgers environment step, warp-resident variant (N = 8..256), sm_100a.
//
// Restates the arithmetic of the reference Burger.step() + getState() + rewards
// (/root/reference/python/_model/Burger.py:333-499, 541-576, 578-675 and
// burger_environment.py:148-176) for a batch of independent environments:
//   * a team of TS lanes owns one environment (32/TS environments per warp); lane
//     registers hold P = N/(2 TS) wavenumbers of the half spectra of v and Fn_old plus 2P
//     adjacent grid points of u.  TS is a template parameter: wide teams minimise latency
//     for small batches, narrow teams minimise instructions per environment for large ones;
//   * the state is read once, `nsub` ABCN sub-steps run out of registers on a shuffle
//     real-FFT, then state / reward / spectrum sums are written back coalesced;
//   * action forcing, 3-mode stochastic forcing, Smagorinsky closures, the float32
//     spectrum chain and the spectral / MSE rewards are fused in;
//   * SF >= 0 fixes the structural mode flags at compile time (hot configurations), SF < 0 (-1 generic, -2 Burger_fd)
//     reads them at run time (generic kernel).
// No arithmetic of one environment depends on another one: results are bitwise
// independent of batch size and packing.
#pragma once
#include "params.h"
#include "warp_fft.cuh"

namespace mpde {

constexpr int STRUCT_FLAGS = F_DFORCE | F_FORCING | F_SSM | F_DSM | F_ACTIONS;

// energy-spectrum row in the reference's float32 chain (Burger.py:562 on complex64 data)
__device__ __forceinline__ float ek_row_f32(float re, float im, int N, float dxf) {
    const float en = __fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im));
    return __fmul_rn(en * (0.5f / (float)N), dxf);
}

// LEAN: compile-time promise of the common training configuration -- no history buffers, no MSE
// truth table, forcing column period 1, state version != 1 (no dudt, so u_prev is not tracked) --
// which removes those branches and registers from the sub-step loop.
// LEAN = 2 ("hot") additionally promises: every call advances (nsub > 0) and writes the state, one agent, state version 0
// or 2 (per-point rows stored straight from registers), sparse (<= 2-tap) action basis, no MSE reward -- the cold
// epilogue paths (shared-memory gather with integer divisions, dense basis product, MSE segments) are not even
// compiled in, which shortens the once-per-launch code the instruction cache has to stream.
// LEAN = 3 ("multi-agent"): the LEAN = 1 promises except that the MSE reward against a truth table stays (per-agent segment
// means, any agent count, every state layout), plus "every call advances" and "sparse action basis" as for LEAN = 2: the MARL
// training configuration.
template <typename T, int N, int TS_, int SF, int LEAN = 0>
struct BurgersWarp {
    static constexpr bool HOT = LEAN == 2;
    static constexpr bool SLIM = LEAN != 0;                   // no history, forcing column period 1, no u_prev tracking
    static constexpr bool NO_MSE = LEAN == 1 || LEAN == 2;    // LEAN = 3: SLIM, but the MSE reward (and any agent count) stays
    static constexpr bool TRAIN = LEAN >= 2;                  // every call advances (nsub > 0), sparse (<= 2-tap) action basis
    using R = RealFFT<T, N, TS_>;
    static constexpr int H = N / 2, TS = R::TS, P = R::P, NH = N / 2 + 1, TPW = 32 / TS;
    // shared-memory stash per team: [0..3] Nyquist-mode constants, [4] kPrevRelErr, [5..] reference spectrum row, [RCP..] its
    // reciprocals (the reward's divisions become correction steps: div_by_rcp)
    static constexpr int STASH = 6 + 2 * H, RCP = 6 + H;          // [RCP + k]: reciprocal of the reference spectrum row
    // doubles of shared memory per team: [FFT exchange area | work (dense actions / state gather / MSE) | stash]
    __host__ __device__ static constexpr int work_doubles(int M) { return M > 2 * N + N / 2 ? M : 2 * N + N / 2; }
    __host__ __device__ static constexpr int scratch_doubles(int M) {
        int scr = (2 * R::SMEM_CX + work_doubles(M) + STASH + 1) & ~1;      // 16-byte granularity
        if (R::SMEM_CX)                                                      // team stride = 64 (mod 128) bytes: see WarpFFT<T,16,4>
            while ((scr & 15) != 8) scr += 2;
        return scr;
    }

    // all cross-lane traffic is scoped to the team: teams share a warp but never each other's data, and -- unless the
    // kernel variant promised warp-uniform control flow (R::whole_warp) -- never each other's control flow
    __device__ __forceinline__ static bool team_any(const R& f, bool pred) {
        return (__ballot_sync(f.c.smask, pred) & f.c.tmask) != 0u;
    }
    __device__ __forceinline__ static T team_sum(const R& f, T x) {
#pragma unroll
        for (int h = TS / 2; h >= 1; h >>= 1) x += shfl_xor(x, h, f.c.smask);
        return x;
    }
    // Real field stored as x[p] = (x_{2j}, x_{2j+1}), j = p*TS + tl.  Returns the left
    // neighbour of the even point (x_{2j-1}) and the right neighbour of the odd point
    // (x_{2j+2}), periodic.
    __device__ __forceinline__ static void halo(const R& f, const Cx<T> (&x)[P], T (&left)[P], T (&right)[P]) {
        const int tl = f.c.tl;
        if constexpr (TS == 1) {
#pragma unroll
            for (int p = 0; p < P; ++p) { right[p] = x[(p + 1) % P].re; left[p] = x[(p + P - 1) % P].im; }
        } else {
            const int lr = f.c.base + ((tl + 1) & (TS - 1));
            const int ll = f.c.base + ((tl - 1) & (TS - 1));
#pragma unroll
            for (int p = 0; p < P; ++p) {
                if constexpr (P == 1) {
                    right[p] = shfl(x[p].re, lr, f.c.smask);
                    left[p] = shfl(x[p].im, ll, f.c.smask);
                } else {
                    // the SENDER picks the register: lane 0 holds the point right of lane TS-1's register p-1 ...
                    const T sr = (tl == 0) ? x[(p + 1) % P].re : x[p].re;          // wanted by my left neighbour
                    const T sl = (tl == TS - 1) ? x[(p + P - 1) % P].im : x[p].im;  // wanted by my right neighbour
                    right[p] = shfl(sr, lr, f.c.smask);
                    left[p] = shfl(sl, ll, f.c.smask);
                }
            }
        }
    }
    // float64 -> float32 -> float64 (quirk Q1: the forcing accumulator of the reference is
    // complex64 unless the stochastic forcing replaced it, Burger.py:335,466)
    __device__ __forceinline__ static T r32(T a) { return (T)(float)a; }

    // Dynamic Smagorinsky closure (Burger.py:357-399 = Burger_fd.py:358-411): Germano identity with a sharp spectral test
    // filter |k| > N//4 that the reference applies IN PLACE to the state v (quirk Q4; in Burger_fd v is refreshed from u at
    // the end of the step, so the filter has no lasting effect there).  xscale X = fft(u^2) (X comes from U = N u).
    __device__ __forceinline__ static void dsm_sgs(const R& f, const Cx<T> (&X)[P], T XN, Cx<T> (&v)[P], Cx<T>& vN, const T (&kw)[P],
                                                   T kwN, const Cx<T> (&ws1)[P], const Cx<T> (&dudx)[P], const Cx<T> (&d2)[P],
                                                   T xscale, T invN, T inv_dx, Cx<T> (&sgs)[P]) {
        Cx<T> z[P];
        // dynamic Smagorinsky, Burger.py:357-399 (Germano identity with a sharp spectral
        // test filter |k| > N//4 applied IN PLACE to the state v, :369-370)
        const T delta = T(2.0 * 3.14159265358979323846 / N), deltah = T(4.0 * 3.14159265358979323846 / N);
        bool cut[P];
#pragma unroll
        for (int p = 0; p < P; ++p) cut[p] = fabs(kw[p]) > T(N / 4);
        const bool cutN = fabs(kwN) > T(N / 4);
        Cx<T> w[P], L1[P], uh[P];
#pragma unroll
        for (int p = 0; p < P; ++p)      // filtered fft(u^2) = xscale X
            w[p] = cut[p] ? cx<T>(0, 0) : cx<T>(xscale * X[p].re, xscale * X[p].im);
        f.inv(w, cutN ? T(0) : xscale * XN, L1);
#pragma unroll
        for (int p = 0; p < P; ++p) {
            L1[p] = cx<T>(L1[p].re * (T(0.5) * invN), L1[p].im * (T(0.5) * invN));
            if (cut[p]) v[p] = cx<T>(0, 0);
        }
        if (cutN) vN = cx<T>(0, 0);
        f.inv(v, vN.re, uh);
#pragma unroll
        for (int p = 0; p < P; ++p) {
            uh[p] = cx<T>(uh[p].re * invN, uh[p].im * invN);
            z[p] = cx<T>(fabs(dudx[p].re) * dudx[p].re, fabs(dudx[p].im) * dudx[p].im);
        }
        Cx<T> W2[P], M1[P];
        T W2N;
        f.fwd(z, W2, W2N, T(1), ws1);
#pragma unroll
        for (int p = 0; p < P; ++p)
            if (cut[p]) W2[p] = cx<T>(0, 0);
        f.inv(W2, cutN ? T(0) : W2N, M1);
        T uhl[P], uhr[P];
        halo(f, uh, uhl, uhr);
        Cx<T> malt[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const T m1a = delta * delta * invN * M1[p].re, m1b = delta * delta * invN * M1[p].im;
            const T da = (uh[p].re - uhl[p]) * inv_dx, db = (uh[p].im - uh[p].re) * inv_dx;
            const T M2a = deltah * deltah * fabs(da) * da, M2b = deltah * deltah * fabs(db) * db;
            malt[p] = cx<T>(T(4) / (deltah * deltah) * M2a - T(1) / (delta * delta) * m1a,
                            T(4) / (deltah * deltah) * M2b - T(1) / (delta * delta) * m1b);
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
        num = team_sum(f, num);
        den = team_sum(f, den);
        const T c = num / den;          // mean/mean: the 1/N cancels (Burger.py:397)
#pragma unroll
        for (int p = 0; p < P; ++p)
            sgs[p] = cx<T>(c * fabs(dudx[p].re) * d2[p].re, c * fabs(dudx[p].im) * d2[p].im);
    }

    __device__ static void run(const SpectralParams<T>& prm, T* smem) {
        const int lane = threadIdx.x & 31;
        const int warp = threadIdx.x >> 5;
        const int wpc = blockDim.x >> 5;
        const int64_t first = ((int64_t)blockIdx.x * wpc + warp) * TPW;
        if (first >= prm.B) return;                                  // whole warp idle 
        const int team = lane / TS;
        const int scr = scratch_doubles(prm.M);
        T* const team_smem = smem + (size_t)(warp * TPW + team) * scr;
        R f;
        f.init(prm.tw, reinterpret_cast<Cx<T>*>(team_smem));
        // the training variants have warp-uniform control flow around every collective (mode flags and trip counts are
        // launch-wide; per-environment conditions only guard loads, stores and selects)
        if constexpr (TRAIN) f.whole_warp();
        const int tl = f.c.tl;
        const int64_t e = first + team;
        const bool has = e < prm.B;
        const int64_t ec = has ? e : 0;
        const int flags = (SF < 0 ? prm.flags : ((prm.flags & ~STRUCT_FLAGS) | SF)) & (TRAIN ? ~(F_NO_ADVANCE | F_BASIS_DENSE) : ~0);
        const bool q1 = !(flags & F_FORCING);
        T* scratch = team_smem + 2 * R::SMEM_CX;
        T* stash = scratch + work_doubles(prm.M);      // per-team constants parked in shared memory (register relief)

        // ---- constant tables first: they may be read while the previous kernel of the stream still runs ----
        int kk[P];
        T kw[P];
#pragma unroll
        for (int p = 0; p < P; ++p) { kk[p] = f.k(p); kw[p] = prm.kwave[kk[p]]; }
        const T kwN = prm.kwave[H];
        const bool sparse_actions = (flags & F_ACTIONS) && !(flags & F_BASIS_DENSE);
        int i_tap[P][2][2];
        T w_tap[P][2][2];
        if (sparse_actions) {
            // the two points of a lane's register (n = 2 j, 2 j + 1) have their taps next to each other in the tables:
            // one 16-byte load of indices and two of weights per register instead of eight scalar loads
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const int j = p * TS + tl;
                const int4 ti = *reinterpret_cast<const int4*>(prm.tap_idx + 4 * j);
                i_tap[p][0][0] = ti.x; i_tap[p][0][1] = ti.y; i_tap[p][1][0] = ti.z; i_tap[p][1][1] = ti.w;
                const Cx<T> wa = ldcx(reinterpret_cast<const Cx<T>*>(prm.tap_w) + 2 * j);
                const Cx<T> wb = ldcx(reinterpret_cast<const Cx<T>*>(prm.tap_w) + 2 * j + 1);
                w_tap[p][0][0] = wa.re; w_tap[p][0][1] = wa.im; w_tap[p][1][0] = wb.re; w_tap[p][1][1] = wb.im;
            }
        }
        pdl_wait();
        pdl_launch_dependents();
        // ---- issue every global load of the mutable state up front (one memory round trip) -----------
        const bool was_live = has && prm.status[ec] == 0;
        bool live = was_live;
        int iout = prm.iout[ec];
        T tnow = prm.tnow[ec];
        const T nu = prm.nu[ec];
        Cx<T> v[P], fn[P];
        float acc32[P];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            v[p] = ldcx(prm.v + ec * NH + kk[p]);
            fn[p] = ldcx(prm.fn + ec * NH + kk[p]);
            acc32[p] = prm.acc[ec * NH + kk[p]];
        }
        // Nyquist mode: complex in the reference when the IC came from a truncated DNS spectrum
        // (quirk Q5); its imaginary part and Im v[0] never reach u.  Carried by the dc lane.
        Cx<T> vN = ldcx(prm.v + ec * NH + H);
        T fnN = ldcx(prm.fn + ec * NH + H).im;
        float accN = prm.acc[ec * NH + H];
        // 2-tap action basis: f_n = w0 a[i0] + w1 a[i1]; the actions are gathered straight from global memory
        T a_tap[P][2][2];
        if (sparse_actions) {
#pragma unroll
            for (int p = 0; p < P; ++p)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    a_tap[p][h][0] = prm.actions[ec * prm.M + i_tap[p][h][0]];
                    a_tap[p][h][1] = prm.actions[ec * prm.M + i_tap[p][h][1]];
                }
        }
        const T v0im = v[0].im;                 // meaningful on the dc lane only
        // reference spectrum row of the step this call ends at + kPrevRelErr: their addresses depend on iout, so the
        // loads go out as early as possible; the values are parked in shared memory at the end of the prologue
        const int nsub = (flags & F_NO_ADVANCE) ? 0 : prm.nsub;
        const bool spec_reward = prm.reward_out && prm.reward_mode == REWARD_SPECTRAL && nsub > 0;
        T ek_pre[P], rk_pre[P], kprev_pre = T(0);
        // the reward's running-mean divisor (iout + 1 at the END of this call) and its reciprocal
        const double cnt_end = (double)(iout + nsub + 1), rcnt_end = 1.0 / cnt_end;
        if (spec_reward) {
            const int64_t ref = prm.ek_map ? prm.ek_map[ec] : 0;
            const int64_t row = iout + nsub < prm.ek_rows ? iout + nsub : prm.ek_rows - 1;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const Cx<T> er = ldcx(prm.ek_pair + (ref * prm.ek_rows + row) * H + kk[p]);      // (E_ref, 1 / E_ref)
                ek_pre[p] = er.re;
                rk_pre[p] = er.im;
            }
            kprev_pre = prm.kprev[ec];
        }

        // ---- per-register constants -------------------------------------------------------
        const T dt = prm.dt;
        const T invN = T(1) / T(N);
        T cv[P], cfo[P], cF[P];
        // v' = [(1-C) v - dt/2 (3 Fn - Fn_old) + dt F] / (1+C),  C = nu k^2 dt / 2  (Burger.py:486-488)
        //    = cv v + cfo (Fn_old - 3 Fn) + cF (dt F)
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const T C = T(0.5) * (kw[p] * kw[p]) * nu * dt;
            const T r = T(1) / (T(1) + C);
            cv[p] = (T(1) - C) * r;
            cfo[p] = T(0.5) * dt * r;
            cF[p] = q1 ? r : dt * r;
        }
        // u is kept as U = N u in registers and the transforms are unscaled: the factor 1/(2 N^2) of
        // fft(u^2 / 2) is folded into the wavenumber that multiplies it (Fn = i k X)
        const T scale_nl = T(0.5) * invN * invN;
        // the nonlinear term comes out of the UNSCALED split step (X2 = 2 fft(U^2)): its 1/2 is folded in here as well
        T kws[P];
#pragma unroll
        for (int p = 0; p < P; ++p) kws[p] = kw[p] * (T(0.5) * scale_nl);
        {
            const T CN = T(0.5) * (kwN * kwN) * nu * dt;
            const T rN = T(1) / (T(1) + CN);
            if (f.dc) {
                stash[0] = (T(1) - CN) * rN;          // cvN
                stash[1] = T(0.5) * dt * rN;          // cfoN
                stash[2] = q1 ? rN : dt * rN;         // cFN
                stash[3] = kwN * (T(0.5) * scale_nl); // kwN (scaled)
            }
        }
        Cx<T> ws1[P];
        f.scaled_twiddles(T(1), ws1);

        // ---- U = N * Re ifft(v) ---------------------------------------------------------------
        Cx<T> U[P], Uprev[P];
        // Burger_fd (generic kernel only): u is the primary variable and lives in the uprev slot; v = fft(u) is derived
        constexpr bool fd = SF == -2;          // Burger_fd has its own instantiation: the generic kernel (SF = -1) carries none of it
        if (fd) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const Cx<T> t = ldcx(reinterpret_cast<const Cx<T>*>(prm.uprev + ec * N) + p * TS + tl);
                U[p] = cx<T>(t.re * T(N), t.im * T(N));          // exact: N is a power of two
            }
        } else {
            f.inv(v, vN.re, U);
        }
#pragma unroll
        for (int p = 0; p < P; ++p) Uprev[p] = U[p];
        const bool v1 = !SLIM && prm.version == 1;             // state version 1 needs u of the previous step (dudt)
        // Burger_fd keeps u itself in the uprev slot; its previous row lives in the (otherwise unused) Fn_old slot
        const Cx<T>* const uprev_row = fd ? reinterpret_cast<const Cx<T>*>(prm.fn + ec * NH) : reinterpret_cast<const Cx<T>*>(prm.uprev + ec * N);
        if ((flags & F_NO_ADVANCE) && v1 && iout > 0) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const Cx<T> t = ldcx(uprev_row + p * TS + tl);
                Uprev[p] = cx<T>(t.re * T(N), t.im * T(N));
            }
        }

        // ---- action field a @ basis (Burger.py:442) --------------------------------------------
        Cx<T> fa[P], Fa[P];
        T FaN = T(0);
#pragma unroll
        for (int p = 0; p < P; ++p) { fa[p] = cx<T>(0, 0); Fa[p] = cx<T>(0, 0); }
        if (flags & F_ACTIONS) {
            if (flags & F_BASIS_DENSE) {
                for (int i = tl; i < prm.M; i += TS) scratch[i] = prm.actions[ec * prm.M + i];
                __syncwarp(f.c.smask);
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    T val[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int n = 2 * (p * TS + tl) + h;
                        T acc = T(0);
                        for (int i = 0; i < prm.M; ++i) acc = fma(scratch[i], prm.basis[(size_t)i * N + n], acc);
                        val[h] = acc;
                    }
                    fa[p] = cx<T>(val[0], val[1]);
                }
                __syncwarp(f.c.smask);
            } else {
#pragma unroll
                for (int p = 0; p < P; ++p)
                    fa[p] = cx<T>(w_tap[p][0][0] * a_tap[p][0][0] + w_tap[p][0][1] * a_tap[p][0][1],
                                  w_tap[p][1][0] * a_tap[p][1][0] + w_tap[p][1][1] * a_tap[p][1][1]);
            }
            if (flags & F_DFORCE) {           // spectrum of a direct forcing is constant over the sub-steps
                Cx<T> z[P];
#pragma unroll
                for (int p = 0; p < P; ++p) z[p] = fa[p];
                f.fwd(z, Fa, FaN, T(1), ws1);
            } else {                          // eddy viscosity: fold 1/dx^2, the 1/N of U and the 1/2 of the raw split step once
                // (Burger_fd applies the field in real space: no split step, no 1/2)
                const T s = (fd ? T(1) : T(0.5)) * invN / (prm.dx * prm.dx);
#pragma unroll
                for (int p = 0; p < P; ++p) fa[p] = cx<T>(fa[p].re * s, fa[p].im * s);
            }
        }

        // ---- stochastic forcing spectrum (Burger.py:410-421): non-zero at k = 1,2,3 only ----------
        // natural-order layouts (k = tl + TS p, TS >= 4) hold k = 1, 2, 3 in register 0 only: one complex register instead of P
        constexpr int FP = (R::C::NATURAL && TS >= 4) ? 1 : P;
        Cx<T> Fc[FP];
#pragma unroll
        for (int p = 0; p < FP; ++p) Fc[p] = cx<T>(0, 0);
        const Cx<T>* fc_row = prm.fcoef + ((flags & F_FORCING_PER_ENV) ? ec : 0) * prm.stepper * 3;
        int col = (flags & F_FORCING) ? iout % prm.stepper : 0;      // column ioutnum % stepper (Q3)
        if ((flags & F_FORCING) && prm.stepper == 1) {
#pragma unroll
            for (int p = 0; p < FP; ++p)
                if (kk[p] >= 1 && kk[p] <= 3) Fc[p] = ldcx(fc_row + (kk[p] - 1));
        }

        // ---- reward bookkeeping ----------------------------------------------------------------
        const float dxf = (float)prm.dx, dtf = (float)dt;
        Cx<T> mse[P];
#pragma unroll
        for (int p = 0; p < P; ++p) mse[p] = cx<T>(0, 0);
        const T inv_dx = T(1) / prm.dx, inv_dx2 = T(1) / (prm.dx * prm.dx);
        const int64_t truth_base =
            prm.truth ? ((prm.truth_map ? prm.truth_map[ec] : 0) * prm.truth_rows) : 0;
        bool bad = false;

        // =============================== sub-steps ==============================================
        if (spec_reward) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                stash[5 + kk[p]] = ek_pre[p];
                stash[RCP + kk[p]] = rk_pre[p];
            }
            if (f.dc) stash[4] = kprev_pre;
        }
        __syncwarp(f.c.smask);
        const bool hist = !SLIM && prm.hist_rows > 0;
        const bool do_mse = !NO_MSE && prm.reward_mode == REWARD_MSE && prm.truth != nullptr;
        const bool multi_col = !SLIM && prm.stepper > 1;
        const bool eddy = (flags & F_ACTIONS) && !(flags & F_DFORCE);
        for (int it = 0; it < nsub; ++it) {
          if (fd) {
            // ---- Burger_fd.step (Burger_fd.py:335-476): u += dt (nu u_xx - u u_x + forcing), v = fft(u) ----------------
            T left[P], right[P];
            halo(f, U, left, right);
            const T delta = T(2.0 * 3.14159265358979323846 / N);
            Cx<T> c3[3];                       // spectrum of the stochastic forcing at k = 1, 2, 3 (this step's column)
            if (flags & F_FORCING) {
#pragma unroll
                for (int k = 0; k < 3; ++k) c3[k] = ldcx(fc_row + col * 3 + k);
                col = (col + 1 == prm.stepper) ? 0 : col + 1;
            }
            Cx<T> zz[P], sgs_d[P];
            if (flags & F_DSM) {               // dynamic Smagorinsky (Burger_fd.py:358-411): same closure as the spectral solver
                Cx<T> dudx_c[P], d2_c[P], X2[P];
                T X2N;
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const T ue = U[p].re * invN, uo = U[p].im * invN, ul = left[p] * invN, ur = right[p] * invN;
                    dudx_c[p] = cx<T>((ue - ul) * inv_dx, (uo - ue) * inv_dx);
                    d2_c[p] = cx<T>((uo - T(2) * ue + ul) * inv_dx2, (ur - T(2) * uo + ue) * inv_dx2);
                    zz[p] = cx<T>(U[p].re * U[p].re, U[p].im * U[p].im);
                }
                f.fwd(zz, X2, X2N, T(1), ws1);
                dsm_sgs(f, X2, X2N, v, vN, kw, kwN, ws1, dudx_c, d2_c, T(2) * scale_nl, invN, inv_dx, sgs_d);
            }
#pragma unroll
            for (int p = 0; p < P; ++p) {
                if (it == nsub - 1) Uprev[p] = U[p];
                const T ue = U[p].re * invN, uo = U[p].im * invN, ul = left[p] * invN, ur = right[p] * invN;
                const T dudx[2] = {(ue - ul) * inv_dx, (uo - ue) * inv_dx};                                   // :465
                const T d2[2] = {(uo - T(2) * ue + ul) * inv_dx2, (ur - T(2) * uo + ue) * inv_dx2};            // :466
                const T uu2[2] = {ue, uo};
                const T act[2] = {(flags & F_DFORCE) ? fa[p].re : fa[p].re * (left[p] - T(2) * U[p].re + U[p].im),
                                  (flags & F_DFORCE) ? fa[p].im : fa[p].im * (U[p].re - T(2) * U[p].im + right[p])};
                T un[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    T forc = T(0);
                    if (flags & F_SSM) forc = (T(0.1) * delta) * (T(0.1) * delta) * fabs(dudx[h]) * d2[h];             // :343-355
                    if (flags & F_DSM) forc = h == 0 ? sgs_d[p].re : sgs_d[p].im;
                    if (flags & F_FORCING) {           // f_j = (2/N) Re sum_k c_k exp(2 pi i k j / N)  (:406-417, replaces)
                        const int j = 2 * (p * TS + tl) + h;
                        forc = T(0);
#pragma unroll
                        for (int k = 1; k <= 3; ++k) {
                            const int m = (k * j) & (N - 1);
                            const Cx<T> w = ldcx(prm.tw + (m & (N / 2 - 1)));          // exp(-2 pi i m / N), m < N/2
                            const T sg = m >= N / 2 ? T(-1) : T(1);
                            forc += (T(2) * invN) * sg * (c3[k - 1].re * w.re + c3[k - 1].im * w.im);
                        }
                    }
                    if (flags & F_ACTIONS) {
                        T af = act[h];
                        if (flags & F_SSMFORCE) af = (af * delta) * (af * delta) * fabs(dudx[h]) * d2[h];               // :447-455
                        forc += af;
                    }
                    un[h] = uu2[h] + dt * (nu * d2[h] - uu2[h] * dudx[h] + forc);                                       // :468
                }
                U[p] = cx<T>(un[0] * T(N), un[1] * T(N));
                zz[p] = U[p];
            }
            Cx<T> X[P];
            T XN;
            f.fwd(zz, X, XN, T(1), ws1);                                                                        // :469
#pragma unroll
            for (int p = 0; p < P; ++p) {
                v[p] = cx<T>(X[p].re * invN, X[p].im * invN);
            }
            vN = cx<T>(XN * invN, T(0));
            iout += 1;
          } else {
            // nonlinear term X = fft(u^2 / 2) (Burger.py:487); with eddy-viscosity actions the forcing
            // (a @ basis) * d2u/dx2 (Burger.py:445-450) is transformed in the same pass
            Cx<T> z[P], X[P], S[P];
            T XN, SN = T(0);
            T left[P], right[P];
            const bool need_nb = (flags & (F_SSM | F_DSM)) || eddy;
            if (need_nb) halo(f, U, left, right);
#pragma unroll
            for (int p = 0; p < P; ++p) z[p] = cx<T>(U[p].re * U[p].re, U[p].im * U[p].im);
            if (eddy) {
                Cx<T> zf[P];
#pragma unroll
                for (int p = 0; p < P; ++p)
                    zf[p] = cx<T>(fa[p].re * (left[p] - T(2) * U[p].re + U[p].im),
                                  fa[p].im * (U[p].re - T(2) * U[p].im + right[p]));
                f.fwd2_raw(z, X, XN, zf, S, SN);          // X = 2 fft(U^2); S = fft of the forcing (its input carries the 1/2)
            } else {
                f.fwd_raw(z, X, XN);
            }

            Cx<T> Fh[P];
            T FhN = T(0);
#pragma unroll
            for (int p = 0; p < P; ++p) Fh[p] = cx<T>(0, 0);

            if (flags & (F_SSM | F_DSM)) {
                Cx<T> sgs[P], dudx[P], d2[P];
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    // upwind first difference, centred second difference (Burger.py:342-346)
                    const T ue = U[p].re * invN, uo = U[p].im * invN, ul = left[p] * invN, ur = right[p] * invN;
                    dudx[p] = cx<T>((ue - ul) * inv_dx, (uo - ue) * inv_dx);
                    d2[p] = cx<T>((uo - T(2) * ue + ul) * inv_dx2, (ur - T(2) * uo + ue) * inv_dx2);
                }
                if (flags & F_SSM) {
                    // Burger.py:339-349, delta = 2 pi / N whatever L is
                    const T cd = T(0.1) * T(2.0 * 3.14159265358979323846 / N);
                    const T cd2 = cd * cd;
#pragma unroll
                    for (int p = 0; p < P; ++p)
                        sgs[p] = cx<T>(cd2 * fabs(dudx[p].re) * d2[p].re, cd2 * fabs(dudx[p].im) * d2[p].im);
                } else {
                    dsm_sgs(f, X, XN, v, vN, kw, kwN, ws1, dudx, d2, scale_nl, invN, inv_dx, sgs);      // X = 2 fft(U^2)
                }
                Cx<T> G[P];
                T GN;
                f.fwd(sgs, G, GN, T(1), ws1);
#pragma unroll
                for (int p = 0; p < P; ++p) Fh[p] = q1 ? cx<T>(r32(G[p].re), r32(G[p].im)) : G[p];
                FhN = q1 ? r32(GN) : GN;
            }

            if (flags & F_FORCING) {
                // the forcing spectrum REPLACES whatever the closures accumulated (Q2)
                if (multi_col) {
#pragma unroll
                    for (int p = 0; p < FP; ++p)
                        if (kk[p] >= 1 && kk[p] <= 3) Fc[p] = ldcx(fc_row + col * 3 + (kk[p] - 1));
                    col = (col + 1 == prm.stepper) ? 0 : col + 1;
                }
#pragma unroll
                for (int p = 0; p < P; ++p) Fh[p] = p < FP ? Fc[p] : cx<T>(0, 0);
                FhN = T(0);
            }

            if (flags & F_ACTIONS) {
                if (flags & F_DFORCE) {
#pragma unroll
                    for (int p = 0; p < P; ++p) S[p] = Fa[p];
                    SN = FaN;
                }
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const Cx<T> t = Fh[p] + S[p];
                    Fh[p] = q1 ? cx<T>(r32(t.re), r32(t.im)) : t;
                }
                FhN = q1 ? r32(FhN + SN) : FhN + SN;
            }

            // ABCN update.  Q1 (cont.): while the forcing accumulator is complex64, `self.dt*Fforcing`
            // (Burger.py:488) is a complex64 product: float32(dt) * float32(F), rounded to float32.
#pragma unroll
            for (int p = 0; p < P; ++p) {
                if (!SLIM && it == nsub - 1) Uprev[p] = U[p];       // u before the last sub-step (dudt of state v1)
                const Cx<T> fnn = cx<T>(-kws[p] * X[p].im, kws[p] * X[p].re);              // i k X
                const Cx<T> F = q1 ? cx<T>((T)__fmul_rn(dtf, (float)Fh[p].re), (T)__fmul_rn(dtf, (float)Fh[p].im)) : Fh[p];
                v[p] = cx<T>(fma(cF[p], F.re, fma(cfo[p], fma(T(-3), fnn.re, fn[p].re), cv[p] * v[p].re)),
                             fma(cF[p], F.im, fma(cfo[p], fma(T(-3), fnn.im, fn[p].im), cv[p] * v[p].im)));
                fn[p] = fnn;
            }
            {   // k = 0: Fn = 0, F real -> Im v[0] is a constant of the motion; Nyquist: F real, Fn imaginary
                const T cvN = stash[0], cfoN = stash[1], cFN = stash[2];
                const T fnnN = stash[3] * XN;
                const T FN = q1 ? (T)__fmul_rn(dtf, (float)FhN) : FhN;
                vN = cx<T>(fma(cFN, FN, cvN * vN.re), fma(cfoN, fma(T(-3), fnnN, fnN), cvN * vN.im));
                fnN = fnnN;
                if (f.dc) v[0].im = v0im;
            }
            iout += 1;

            // U = N Re ifft(v) (Burger.py:491)
            f.inv(v, vN.re, U);
          }

            // float32 spectrum chain (Q6): Ek row from complex64(v), sequential float32 sum.  The same cast is where the
            // reference detects a blow-up (`vv[i] = v` overflows complex64 with np.seterr(over='raise'), Burger.py:8,498):
            // a component that is inf / nan AFTER the cast marks the environment -- checked on the FP32 pipe
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

            if (hist) {
                // blow-up semantics: rows are only written while the env is healthy
                live = live && !team_any(f, bad);
                if (live && iout < prm.hist_rows) {
                    const int64_t hrow = e * prm.hist_rows + iout;
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const int j = p * TS + tl;
                        if (prm.uu_hist)
                            stcx(reinterpret_cast<Cx<T>*>(prm.uu_hist + hrow * N) + j, cx<T>(U[p].re * invN, U[p].im * invN));
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

            if (do_mse) {
                // Burger.py:589-599 after every sub-step, averaged over them (burger_environment.py:153)
                const int64_t row = iout < prm.truth_rows ? iout : prm.truth_rows - 1;
                const Cx<T>* tr = reinterpret_cast<const Cx<T>*>(prm.truth + (truth_base + row) * N);
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const Cx<T> t = ldcx(tr + p * TS + tl);
                    const T da = t.re - U[p].re * invN, db = t.im - U[p].im * invN;
                    mse[p].re += da * da;
                    mse[p].im += db * db;
                }
            }
        }

        // =============================== epilogue ===============================================
        // A blown-up environment (non-finite / > FLT_MAX spectrum, the reference's FloatingPointError)
        // keeps the state it had before this call and is marked TRUNCATED.
        for (int it = 0; it < nsub; ++it) tnow += dt;            // t += dt per step (Burger.py:494)
        if (nsub > 0) {
            const bool blew = team_any(f, bad);
            if (was_live && blew && f.dc) prm.status[e] = 1;
            live = live && !blew;
        }
        if (nsub > 0 && live) {
#pragma unroll
            for (int p = 0; p < P; ++p) {
                stcx(prm.v + e * NH + kk[p], v[p]);
                if (!fd) stcx(prm.fn + e * NH + kk[p], fn[p]);
                prm.acc[e * NH + kk[p]] = acc32[p];
                if (fd) {                  // Burger_fd: the field itself is the primary variable
                    stcx(reinterpret_cast<Cx<T>*>(prm.uprev + e * N) + p * TS + tl, cx<T>(U[p].re * invN, U[p].im * invN));
                    if (v1) stcx(reinterpret_cast<Cx<T>*>(prm.fn + e * NH) + p * TS + tl, cx<T>(Uprev[p].re * invN, Uprev[p].im * invN));
                } else if (v1) {           // u before the last sub-step: only state version 1 (dudt) reads it
                    stcx(reinterpret_cast<Cx<T>*>(prm.uprev + e * N) + p * TS + tl, cx<T>(Uprev[p].re * invN, Uprev[p].im * invN));
                }
            }
            if (f.dc) {
                stcx(prm.v + e * NH + H, vN);
                if (!fd) stcx(prm.fn + e * NH + H, cx<T>(T(0), fnN));
                prm.acc[e * NH + H] = accN;
                prm.iout[e] = iout;
                prm.tnow[e] = tnow;
            }
        }

        const T inf = T(1) / T(0);
        // double-buffered gather (PeerSink): this step's copy of the output buffers
        const int64_t poff = prm.peer.parity * prm.peer.parity_stride;
        T* const state_out = prm.state_out ? prm.state_out + poff : nullptr;
        T* const reward_out = prm.reward_out ? prm.reward_out + poff : nullptr;
        if (HOT || (prm.state_out && prm.A == 1 && prm.version <= 2)) {
            // getState, single agent, versions 0/1/2 (Burger.py:617-622): rows are per-point fields, so the
            // lane's two adjacent points go out as one 16-byte store each -- no shared-memory gather
            const int ver = (SLIM && prm.version == 1) ? 0 : prm.version;
            T left[P], right[P];
            halo(f, U, left, right);
            const T sd2 = inv_dx2 * invN, sdt = invN / dt;
            const int rl_shift = ilog2(H) + (ver == 0 ? 0 : 1);        // complex words per state row = 1 << rl_shift
            const Cx<T> ii = cx<T>(inf, inf);
            Cx<T> ra[P], rb[P];                                       // first / second field of the row at this lane's points
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const Cx<T> u = cx<T>(U[p].re * invN, U[p].im * invN);
                const Cx<T> d2 = cx<T>((left[p] - T(2) * U[p].re + U[p].im) * sd2, (U[p].re - T(2) * U[p].im + right[p]) * sd2);
                if (ver == 0) {
                    ra[p] = d2;
                } else if (ver == 1) {
                    ra[p] = cx<T>((U[p].re - Uprev[p].re) * sdt, (U[p].im - Uprev[p].im) * sdt);
                    rb[p] = d2;
                } else {
                    ra[p] = u;
                    rb[p] = cx<T>(u.re * u.re, u.im * u.im);
                }
                if (!live) ra[p] = rb[p] = ii;
            }
            const bool gather = prm.peer.mc_state != nullptr || prm.peer.n_data > 0;
            if (!gather || !prm.peer.row_stores) {
                // every lane stores its own points (a team covers 64 contiguous bytes per store instruction): one GPU, or a
                // gather whose links are not the bound
                if (has) {
                    const int64_t off = e * ((int64_t)2 << rl_shift);                       // row start, in T
                    auto put = [&](T* base) {
                        Cx<T>* const row = reinterpret_cast<Cx<T>*>(base + off);
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            stcx(row + p * TS + tl, ra[p]);
                            if (ver != 0) stcx(row + H + p * TS + tl, rb[p]);
                        }
                    };
                    if (prm.peer.mc_state) {
                        Cx<T>* const row = reinterpret_cast<Cx<T>*>(static_cast<T*>(prm.peer.mc_state) + poff + off);
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            st_multicast(row + p * TS + tl, ra[p]);
                            if (ver != 0) st_multicast(row + H + p * TS + tl, rb[p]);
                        }
                    } else {
                        put(state_out);
#pragma unroll 1
                        for (int q = 0; q < prm.peer.n_data; ++q) put(static_cast<T*>(prm.peer.state[q]) + poff);
                    }
                }
            } else {
                // Fused gather: stage this team's row in its scratch area; then the WARP writes whole rows: 16 consecutive
                // lanes cover 256 contiguous bytes of one environment's row.  (Those stores travel over NVLink, whose write
                // efficiency follows the size of the contiguous segment: 8 GPUs reached 0.55 TB/s of ingress with the 64-byte
                // segments a 4-lane team covers by itself, 0.60 - 0.66 TB/s with whole rows.)
                Cx<T>* const stage = reinterpret_cast<Cx<T>*>(scratch);
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    stcx(stage + p * TS + tl, ra[p]);
                    if (ver != 0) stcx(stage + H + p * TS + tl, rb[p]);
                }
                __syncwarp();               // the whole warp: every lane is here (mode flags and trip counts are warp-uniform)
                for (int idx = lane; idx < (TPW << rl_shift); idx += 32) {
                    const int t = idx >> rl_shift, c = idx & ((1 << rl_shift) - 1);
                    if (first + t >= prm.B) continue;
                    const Cx<T> val = ldcx(reinterpret_cast<const Cx<T>*>(smem + (size_t)(warp * TPW + t) * scr + 2 * R::SMEM_CX) + c);
                    const int64_t off = (first + t) * ((int64_t)2 << rl_shift);         // row start, in T
                    // local row + the same row in every peer's gather buffer (multi-GPU, PeerSink)
                    if (prm.peer.mc_state) {      // one multicast store reaches every rank's buffer (this one included)
                        st_multicast(reinterpret_cast<Cx<T>*>(static_cast<T*>(prm.peer.mc_state) + poff + off) + c, val);
                    } else {
                        stcx(reinterpret_cast<Cx<T>*>(state_out + off) + c, val);
#pragma unroll 1
                        for (int q = 0; q < prm.peer.n_data; ++q)
                            stcx(reinterpret_cast<Cx<T>*>(static_cast<T*>(prm.peer.state[q]) + poff + off) + c, val);
                    }
                }
                __syncwarp();
            }
        } else if (!HOT && prm.state_out) {
            // getState (Burger.py:604-675) through a shared-memory gather so that every
            // version / agent-window layout becomes one coalesced row store
            const int ver = (SLIM && prm.version == 1) ? 0 : prm.version, A = prm.A;
            T left[P], right[P];
            halo(f, U, left, right);
            __syncwarp(f.c.smask);
            T* f0 = scratch;
            T* f1 = f0 + N;
            T* ek = f1 + N;
            const T sd2 = inv_dx2 * invN, sdt = invN / dt;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const int j = p * TS + tl;
                const Cx<T> u = cx<T>(U[p].re * invN, U[p].im * invN);
                const Cx<T> d2 = cx<T>((left[p] - T(2) * U[p].re + U[p].im) * sd2,
                                       (U[p].re - T(2) * U[p].im + right[p]) * sd2);
                const Cx<T> dudt = cx<T>((U[p].re - Uprev[p].re) * sdt, (U[p].im - Uprev[p].im) * sdt);
                Cx<T> a = d2, b = d2;
                if (ver == 1) { a = dudt; }
                else if (ver == 2) { a = u; b = cx<T>(u.re * u.re, u.im * u.im); }
                else if (ver == 4) { a = u; }
                stcx(reinterpret_cast<Cx<T>*>(f0) + j, a);
                stcx(reinterpret_cast<Cx<T>*>(f1) + j, b);
                // Burger.py:653, from the live float64 v (the spectrum tail of state versions 3 / 4)
                if (ver >= 3) ek[kk[p]] = T(0.5) * ((v[p].re * v[p].re + v[p].im * v[p].im) / T(N)) * prm.dx;
            }
            __syncwarp(f.c.smask);
            const int nf = (ver == 1 || ver == 2) ? 2 : 1;
            const int seg = A == 1 ? N : N / A + 2;
            const int tail = (ver == 3 || ver == 4) ? N / 2 : 0;
            const int RL = nf * seg + tail;
            const int S = A * RL;
            // (a, r) = (agent, position inside the agent's row) of output number o = a RL + r, by a multiply-high division:
            // floor(o m / 2^32) with m = floor(2^32 / RL) + 1 is exact while o RL < 2^32 (here o < S = A RL <= 33536 and
            // RL <= 640 for N <= 256; RL >= 3, so m fits 32 bits) -- an integer division per element would cost more than
            // the ten solver steps' worth of state arithmetic at A = N, and an incremental (a, r) needs divergent loops
            static_assert(N <= 256, "multiply-high (a, r) split assumes S * RL < 2^32");
            const int npa = A == 1 ? N : N / A;
            const unsigned magic = (unsigned)(0x100000000ull / (unsigned long long)RL) + 1u;
            const int nfseg = nf * seg, start0 = A == 1 ? N : N - 1;
            auto element = [&](int o) {
                const int a = (int)__umulhi((unsigned)o, magic), r = o - a * RL;
                int idx;
                if (r < nfseg) {
                    const int fld = r >= seg ? 1 : 0;
                    idx = fld * N + ((start0 + a * npa + r - fld * seg) & (N - 1));
                } else {
                    idx = 2 * N + (r - nfseg);
                }
                const T val = f0[idx];
                return live ? val : inf;                                     // Burger.py:633-643
            };
            if (has && (S & 1) == 0) {
                // rows of even length: two numbers per 16-byte store, the lanes of a team cover contiguous 16 TS bytes
                for (int o = 2 * tl; o < S; o += 2 * TS) {
                    const Cx<T> out = cx<T>(element(o), element(o + 1));
                    if (prm.peer.mc_state) {
                        st_multicast(reinterpret_cast<Cx<T>*>(static_cast<T*>(prm.peer.mc_state) + poff + e * S + o), out);
                    } else {
                        stcx(reinterpret_cast<Cx<T>*>(state_out + e * S + o), out);
#pragma unroll 1
                        for (int q = 0; q < prm.peer.n_data; ++q)
                            stcx(reinterpret_cast<Cx<T>*>(static_cast<T*>(prm.peer.state[q]) + poff + e * S + o), out);
                    }
                }
            } else if (has) {
                for (int o = tl; o < S; o += TS) {
                    const T out = element(o);
                    if (prm.peer.mc_state) {
                        st_multicast(static_cast<T*>(prm.peer.mc_state) + poff + e * S + o, out);
                    } else {
                        state_out[e * S + o] = out;
#pragma unroll 1
                        for (int q = 0; q < prm.peer.n_data; ++q) static_cast<T*>(prm.peer.state[q])[poff + e * S + o] = out;
                    }
                }
            }
            __syncwarp(f.c.smask);
        }

        if (spec_reward) {
            // burger_environment.py:172-176 on the running float32 sums
            const int A = prm.A;
            // squared relative errors go back into the stash (over the reference row) and are summed in WAVENUMBER order
            // by every lane: the same sequence of additions whatever the team size (variants must agree bitwise)
            // every division here is a correction step on a reciprocal computed in the prologue (rcnt_end) or when the
            // reference was set (stash[RCP + k]); one range check for the whole lane, the divisions themselves otherwise
            T qq[P];
            bool safe = true;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const double a = (double)acc32[p];
                const T es = (T)div_rcp_fast(a, cnt_end, rcnt_end);
                const T er = stash[5 + kk[p]];
                const T x = fabs(er - es);
                if (kk[p] >= 1) safe = safe && rcp_ok(a, cnt_end) && rcp_ok(x, er);      // k = 0 is not part of the reward
                const T q = div_rcp_fast(x, er, stash[RCP + kk[p]]);
                qq[p] = q * q;
            }
            if (!safe) {
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const T es = (T)((double)acc32[p] / cnt_end);
                    const T er = stash[5 + kk[p]];
                    const T q = fabs(er - es) / er;
                    qq[p] = q * q;
                }
            }
#pragma unroll
            for (int p = 0; p < P; ++p)
                if (kk[p] >= 1) stash[5 + kk[p]] = qq[p];
            __syncwarp(f.c.smask);
            T part = T(0);
            for (int k = 1; k < H; ++k) part += stash[5 + k];
            part = div_by_rcp(part, T(H - 1), T(1) / T(H - 1));
            const T r = live ? stash[4] - part : -inf;
            if (has) {
                for (int a = tl; a < A; a += TS) {
                    if (prm.peer.mc_reward) {
                        st_multicast(static_cast<T*>(prm.peer.mc_reward) + poff + e * A + a, r);
                    } else {
                        reward_out[e * A + a] = r;
#pragma unroll 1
                        for (int q = 0; q < prm.peer.n_data; ++q) static_cast<T*>(prm.peer.reward[q])[poff + e * A + a] = r;
                    }
                }
                if (f.dc && live) prm.kprev[e] = part;
            }
        }
        if (!NO_MSE && prm.reward_out && prm.reward_mode == REWARD_MSE && prm.truth) {
            const int A = prm.A, W = N / A;
            if (!TRAIN && nsub == 0) {   // getMseReward() of the current state (Burger.py:578-601)
                const int64_t row = iout < prm.truth_rows ? iout : prm.truth_rows - 1;
                const Cx<T>* tr = reinterpret_cast<const Cx<T>*>(prm.truth + (truth_base + row) * N);
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const Cx<T> t = ldcx(tr + p * TS + tl);
                    const T da = t.re - U[p].re * invN, db = t.im - U[p].im * invN;
                    mse[p] = cx<T>(da * da, db * db);
                }
            }
            __syncwarp(f.c.smask);
#pragma unroll
            for (int p = 0; p < P; ++p) stcx(reinterpret_cast<Cx<T>*>(scratch) + p * TS + tl, mse[p]);
            __syncwarp(f.c.smask);
            // x / nsub for every agent of the row: one reciprocal, then the quotient corrected with the exact remainder
            // (q0 = x r, q = q0 + (x - q0 n) r is the correctly rounded x / n when r is the correctly rounded 1 / n and nothing
            // under- or overflows); operands outside the safe range take the division itself
            const T dn = T(nsub > 0 ? nsub : 1), rn = T(1) / dn;
            auto div_n = [&](T x) {
                if (nsub <= 1) return x;
                if (sizeof(T) == 8 && fabs(x) > T(1e-30) && fabs(x) < T(1e30)) {
                    const T q0 = x * rn;
                    return fma(fma(-q0, dn, x), rn, q0);
                }
                return x / dn;
            };
            auto agent_reward = [&](int a) {             // agent a owns points [a N/A, (a+1) N/A)
                T sum = T(0);
                for (int j = 0; j < W; ++j) sum += scratch[a * W + j];
                const T mean = W == 1 ? sum : sum / T(W);                     // x / 1 == x: skip the division for per-point agents
                return live ? div_n(-mean) : -inf;
            };
            if (has && (A & 1) == 0) {                   // two agents per 16-byte store
                for (int a = 2 * tl; a < A; a += 2 * TS) {
                    const Cx<T> r = cx<T>(agent_reward(a), agent_reward(a + 1));
                    if (prm.peer.mc_reward) {
                        st_multicast(reinterpret_cast<Cx<T>*>(static_cast<T*>(prm.peer.mc_reward) + poff + e * A + a), r);
                    } else {
                        stcx(reinterpret_cast<Cx<T>*>(reward_out + e * A + a), r);
#pragma unroll 1
                        for (int q = 0; q < prm.peer.n_data; ++q)
                            stcx(reinterpret_cast<Cx<T>*>(static_cast<T*>(prm.peer.reward[q]) + poff + e * A + a), r);
                    }
                }
            } else if (has) {
                for (int a = tl; a < A; a += TS) {
                    const T r = agent_reward(a);
                    if (prm.peer.mc_reward) {
                        st_multicast(static_cast<T*>(prm.peer.mc_reward) + poff + e * A + a, r);
                    } else {
                        reward_out[e * A + a] = r;
#pragma unroll 1
                        for (int q = 0; q < prm.peer.n_data; ++q) static_cast<T*>(prm.peer.reward[q])[poff + e * A + a] = r;
                    }
                }
            }
        }
    }
};

}  // namespace mpde
