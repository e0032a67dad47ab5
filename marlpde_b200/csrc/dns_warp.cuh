// Burgers DNS step (ground-truth generation, /root/reference/python/_model/Burger.py:486-499 inside simulate() :501-539),
// N = 1024, ONE WARP PER ENVIRONMENT, sm_100a.
//
// The CTA-resident kernel of round 1 (spectral_cta.cuh) kept the field in shared memory and ran radix-2 passes with a
// __syncthreads and a global twiddle load per butterfly: 4 % of the FP64 peak.  Here the whole environment lives in
// the REGISTERS of one warp for all `nsub` fused steps:
//   * the real field is 512 complex points z_j = (x_2j, x_2j+1), 16 per lane ("T layout": j = 32 r + lane); the half
//     spectra v, Fn_old are 16 wavenumbers per lane ("F layout": k = 32 r + lane, the Nyquist value on lane 0);
//   * 512 = 16 x 2 x 16: a 16-point transform in registers (radix 4 x 4, compile-time twiddles), one twiddle from a
//     shared-memory table, ONE transpose through shared memory (bank-conflict-free, row stride 33), a radix-2 stage
//     between lanes l and l ^ 16 (shuffles), a second 16-point transform in registers -- natural order in and out, so
//     history rows leave as fully coalesced 512-byte stores;
//   * the Hermitian partner X[512 - k] needed by the real-transform split comes by shuffle from lane 32 - l
//     (register 15 - r; lane 0 pairs with itself);
//   * per step and warp: ~1500 FP64 and ~700 LSU instructions, no block-wide barrier (__syncwarp only).
// Scope: the DNS configuration -- no actions, no closures, no state / reward outputs (those calls take the generic
// CTA kernel); stochastic forcing, uu / vv / Ek_ktt history rows and the float32 spectrum chain are fused in.
#pragma once
#include "params.h"
#include "burgers_warp.cuh"   // ek_row_f32

namespace mpde {

// VS: keep the spectrum v in shared memory instead of 64 registers (the Hermitian partner is then a second shared-memory
// read instead of a shuffle): no spills and more scheduling freedom for the transforms, +128 LSU cycles per step
template <typename T, bool VS = true>
struct Dns1024 {
    static constexpr int N = 1024, H = 512, NH = 513;
    static constexpr int ROW = 33;                       // exchange-buffer row stride (complex words): conflict-free transposes
    // shared memory per warp, in complex words: exchange 16 x 33 | T1 16 x 32 | T2 16 x 2 | (cv, cfo) 513 (+1) | Fn_old 512 |
    // acc (float) 513; entry 512 of the last three tables and FN[513] (= v[N/2]) belong to the Nyquist mode, which lane 0
    // updates out of shared memory.  Fn_old is touched once per step, so it lives here instead of in 64 more registers (v, the work
    // array and their temporaries already fill the 255-register budget).
    static constexpr int CX_WORDS = 16 * ROW + 512 + 32 + 514 + 514 + (VS ? 512 : 0);
    static size_t smem_bytes() { return sizeof(Cx<T>) * CX_WORDS + sizeof(float) * 516; }

    // ---- 16-point transform in registers, natural order in and out ---------------------------------------------
    template <bool INV>
    __device__ __forceinline__ static void dft4(Cx<T>& x0, Cx<T>& x1, Cx<T>& x2, Cx<T>& x3) {
        const Cx<T> a0 = x0 + x2, a1 = x0 - x2, a2 = x1 + x3, a3 = x1 - x3;
        x0 = a0 + a2;
        x2 = a0 - a2;
        const Cx<T> m = cx<T>(a1.re + a3.im, a1.im - a3.re), q = cx<T>(a1.re - a3.im, a1.im + a3.re);   // a1 -/+ i a3
        x1 = INV ? q : m;
        x3 = INV ? m : q;
    }
    // x * W16^E (forward) or x * conj(W16^E) (inverse), E in {1, 2, 3, 4, 6, 9}
    template <int E, bool INV>
    __device__ __forceinline__ static Cx<T> tw16(Cx<T> a) {
        constexpr double C1 = 0.92387953251128673848, S1 = 0.38268343236508978178, R = 0.70710678118654752440;
        if constexpr (E == 4) return INV ? cx<T>(-a.im, a.re) : cx<T>(a.im, -a.re);
        else if constexpr (E == 2) return INV ? cx<T>((a.re - a.im) * T(R), (a.im + a.re) * T(R)) : cx<T>((a.re + a.im) * T(R), (a.im - a.re) * T(R));
        else if constexpr (E == 6) return INV ? cx<T>(-(a.re + a.im) * T(R), (a.re - a.im) * T(R)) : cx<T>((a.im - a.re) * T(R), -(a.re + a.im) * T(R));
        else {
            constexpr double wr = E == 1 ? C1 : (E == 3 ? S1 : -C1);
            constexpr double wi = (E == 1 ? -S1 : (E == 3 ? -C1 : S1)) * (INV ? -1.0 : 1.0);
            return cx<T>(fma(a.re, T(wr), -(a.im * T(wi))), fma(a.re, T(wi), a.im * T(wr)));
        }
    }
    template <bool INV>
    __device__ __forceinline__ static void fft16(Cx<T> (&x)[16]) {
        // n = 4 n1 + n2, k = k1 + 4 k2:  W16^(nk) = W4^(n1 k1) W16^(n2 k1) W4^(n2 k2)
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) dft4<INV>(x[n2], x[4 + n2], x[8 + n2], x[12 + n2]);      // slot 4 k1 + n2
        x[5] = tw16<1, INV>(x[5]);  x[6] = tw16<2, INV>(x[6]);   x[7] = tw16<3, INV>(x[7]);
        x[9] = tw16<2, INV>(x[9]);  x[10] = tw16<4, INV>(x[10]); x[11] = tw16<6, INV>(x[11]);
        x[13] = tw16<3, INV>(x[13]); x[14] = tw16<6, INV>(x[14]); x[15] = tw16<9, INV>(x[15]);
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) dft4<INV>(x[4 * k1], x[4 * k1 + 1], x[4 * k1 + 2], x[4 * k1 + 3]);   // slot 4 k1 + k2
        // natural order: out[k1 + 4 k2] = slot[4 k1 + k2] (a transpose of the 4 x 4 register tile: free renaming)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = a + 1; b < 4; ++b) { const Cx<T> t = x[4 * a + b]; x[4 * a + b] = x[4 * b + a]; x[4 * b + a] = t; }
    }

    // ---- 512-point complex transform over the warp ---------------------------------------------------------------
    // forward: T layout (j = 32 r + lane) -> F layout (k = 32 r + lane); inverse: F -> T, unnormalised.
    template <bool INV>
    __device__ __forceinline__ static void fft512(Cx<T> (&z)[16], Cx<T>* E, const Cx<T>* T1, const Cx<T>* T2, int lane) {
        const int k1 = lane & 15, hbit = lane >> 4;
        const T sg = hbit ? T(-1) : T(1);
        if constexpr (!INV) {
            fft16<false>(z);                                                      // over r (n1) -> k1 in registers
#pragma unroll
            for (int r = 1; r < 16; ++r) z[r] = cmul(z[r], ldcx(T1 + r * 32 + lane));      // W512^(lane k1)
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 16; ++r) stcx(E + r * ROW + lane, z[r]);
            __syncwarp();
#pragma unroll
            for (int m = 0; m < 16; ++m) z[m] = ldcx(E + k1 * ROW + 16 * hbit + m);        // lane (k1, h): n2 = 16 h + m
#pragma unroll
            for (int m = 0; m < 16; ++m) {                                         // radix 2 between lanes (k1, 0) and (k1, 1)
                const Cx<T> o = shfl_xor(z[m], 16, 0xffffffffu);
                z[m] = cx<T>(fma(sg, z[m].re, o.re), fma(sg, z[m].im, o.im));
            }
#pragma unroll
            for (int m = 1; m < 16; ++m) z[m] = cmul(z[m], ldcx(T2 + 2 * m + hbit));       // W32^(m h)
            fft16<false>(z);                                                      // over m -> q: k = k1 + 16 h + 32 q
        } else {
            fft16<true>(z);
#pragma unroll
            for (int m = 1; m < 16; ++m) z[m] = cmulc(z[m], ldcx(T2 + 2 * m + hbit));
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const Cx<T> o = shfl_xor(z[m], 16, 0xffffffffu);
                z[m] = cx<T>(fma(sg, z[m].re, o.re), fma(sg, z[m].im, o.im));
            }
            __syncwarp();
#pragma unroll
            for (int m = 0; m < 16; ++m) stcx(E + k1 * ROW + 16 * hbit + m, z[m]);
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 16; ++r) z[r] = ldcx(E + r * ROW + lane);
#pragma unroll
            for (int r = 1; r < 16; ++r) z[r] = cmulc(z[r], ldcx(T1 + r * 32 + lane));
            fft16<true>(z);
        }
    }

    // value at the Hermitian partner index (512 - k) mod 512 of k = 32 r + lane: lane 32 - l holds it in register 15 - r;
    // lane 0 pairs with itself (register (16 - r) & 15).  `r` is a compile-time constant after unrolling.
    __device__ __forceinline__ static Cx<T> mirror(const Cx<T> (&z)[16], int r, int lane) {
        const Cx<T> far = shfl(z[15 - r], (32 - lane) & 31, 0xffffffffu);
        const Cx<T> own = z[(16 - r) & 15];
        return lane == 0 ? own : far;
    }

    __device__ static void run(const SpectralParams<T>& prm, unsigned char* smem_raw) {
        const int lane = threadIdx.x & 31;
        const int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        if (e >= prm.B) return;
        Cx<T>* E = reinterpret_cast<Cx<T>*>(smem_raw);                             // one warp per CTA
        Cx<T>* T1 = E + 16 * ROW;
        Cx<T>* T2 = T1 + 512;
        Cx<T>* CC = T2 + 32;
        Cx<T>* FN = CC + 514;
        Cx<T>* V = FN + 514;                              // VS only: v[k], natural order
        float* acc = reinterpret_cast<float*>(V + (VS ? 512 : 0));
        const int flags = prm.flags;
        const T dt = prm.dt, invN = T(1) / T(N);
        const T nu = prm.nu[e];
        const float dxf = (float)prm.dx;
        auto tw1024 = [&](int j) {                       // exp(-2 pi i j / 1024), j in [0, 1024)
            const Cx<T> w = ldcx(prm.tw + (j & 511));
            return (j & 512) ? cx<T>(-w.re, -w.im) : w;
        };
        // ---- tables ------------------------------------------------------------------------------------------
#pragma unroll 4
        for (int r = 0; r < 16; ++r) T1[r * 32 + lane] = tw1024((2 * r * lane) & 1023);         // W512^(lane r)
        T2[lane] = tw1024((32 * (lane >> 1) * (lane & 1)) & 1023);                               // [m][h]: W32^(m h)
        // wavenumber of register r: kw0 + r dk (the table value to <= 1 ulp; keeps 16 doubles out of the register file)
        const T kw0 = prm.kwave[lane], dk = prm.kwave[32];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int k = 32 * r + lane;
            const T kwr = prm.kwave[k];
            const T C = T(0.5) * (kwr * kwr) * nu * dt;                           // Burger.py:486
            const T rr = T(1) / (T(1) + C);
            CC[k] = cx<T>((T(1) - C) * rr, T(0.5) * dt * rr);
            acc[k] = prm.acc[e * NH + k];
        }
        if (lane == 0) {                                 // Nyquist mode (k = N/2): constants, Fn_old (imaginary) and v in shared memory
            const T kwN = prm.kwave[H];
            const T C = T(0.5) * (kwN * kwN) * nu * dt;
            const T rr = T(1) / (T(1) + C);
            CC[H] = cx<T>((T(1) - C) * rr, T(0.5) * dt * rr);
            acc[H] = prm.acc[e * NH + H];
            FN[H] = ldcx(prm.fn + e * NH + H);
            FN[H + 1] = ldcx(prm.v + e * NH + H);
        }
        const Cx<T> wbase = tw1024(lane);                                          // W1024^lane; W1024^k = wbase * W32^r
        // fft(u^2 / 2) from U = N u: scale 0.5 / N^2, times the 1/2 of the split step
        const T hs = T(0.25) * invN * invN;
        // ---- state -------------------------------------------------------------------------------------------
        const bool was_live = prm.status[e] == 0;
        bool live = was_live, bad = false;
        int iout = prm.iout[e];
        T tnow = prm.tnow[e];
        Cx<T> v[VS ? 1 : 16], z[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const Cx<T> t = ldcx(prm.v + e * NH + 32 * r + lane);
            if constexpr (VS) V[32 * r + lane] = t; else v[r] = t;
            FN[32 * r + lane] = ldcx(prm.fn + e * NH + 32 * r + lane);
        }
        __syncwarp();
        auto vget = [&](int r) { if constexpr (VS) return ldcx(V + 32 * r + lane); else return v[r]; };
        auto vmirror = [&](int r) {
            if constexpr (VS) return ldcx(V + ((512 - 32 * r - lane) & 511));
            else return mirror(v, r, lane);
        };

        // pre-processing of the inverse real transform + inverse: z = N Re ifft(v) as pairs (x_2j, x_2j+1)
        auto to_real = [&]() {
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const Cx<T> wk = cmul(wbase, w32(r));
                const Cx<T> vr = vget(r), vm = vmirror(r);
                const Cx<T> Ee = cx<T>(vr.re + vm.re, vr.im - vm.im);
                const Cx<T> Dd = cx<T>(vr.re - vm.re, vr.im + vm.im);
                z[r] = cx<T>(fma(-Dd.im, wk.re, fma(Dd.re, wk.im, Ee.re)), fma(Dd.re, wk.re, fma(Dd.im, wk.im, Ee.im)));
            }
            {   // k = 0 (lane 0, register 0): only Re v[0], Re v[N/2] enter (selects, no branch: the warp stays converged)
                const T vNre = FN[H + 1].re, v0re = vget(0).re;
                const Cx<T> dc = cx<T>(v0re + vNre, v0re - vNre);
                z[0] = lane == 0 ? dc : z[0];
            }
            fft512<true>(z, E, T1, T2, lane);
        };
        to_real();

        const int nsub = (flags & F_NO_ADVANCE) ? 0 : prm.nsub;
        const int64_t fc_row = (flags & F_FORCING_PER_ENV) ? e : 0;
        const bool forcing = (flags & F_FORCING) != 0;
        for (int it = 0; it < nsub; ++it) {
            if (it == nsub - 1) {                          // u before the last sub-step (dudt of state version 1)
#pragma unroll
                for (int r = 0; r < 16; ++r)
                    stcx(reinterpret_cast<Cx<T>*>(prm.uprev + e * N) + 32 * r + lane, cx<T>(z[r].re * invN, z[r].im * invN));
            }
            // ---- Fn = i k fft(u^2 / 2)  (Burger.py:487) ------------------------------------------------------------
#pragma unroll
            for (int r = 0; r < 16; ++r) z[r] = cx<T>(z[r].re * z[r].re, z[r].im * z[r].im);
            fft512<false>(z, E, T1, T2, lane);
            const T XN = ((z[0].re - z[0].im)) * (T(2) * hs);      // lane 0: fft(u^2/2)[N/2] (real)
            // ---- ABCN update (Burger.py:486-489) ------------------------------------------------------------------
            float fre[16], fim[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const Cx<T> wk = cmul(wbase, w32(r));
                const Cx<T> zm = mirror(z, r, lane);
                const Cx<T> Ee = cx<T>(z[r].re + zm.re, z[r].im - zm.im);                  // 2 fft(x_even)[k]
                const Cx<T> Oo = cx<T>(z[r].im + zm.im, zm.re - z[r].re);                  // 2 fft(x_odd)[k]
                const Cx<T> X = cx<T>(hs * (Ee.re + fma(wk.re, Oo.re, -(wk.im * Oo.im))), hs * (Ee.im + fma(wk.re, Oo.im, wk.im * Oo.re)));
                const T kwr = fma(T(r), dk, kw0);
                const Cx<T> fnn = cx<T>(-kwr * X.im, kwr * X.re);
                const Cx<T> c = ldcx(CC + 32 * r + lane);                                  // (cv, cfo)
                const Cx<T> fo = ldcx(FN + 32 * r + lane);
                const Cx<T> vo = vget(r);
                Cx<T> vn = cx<T>(fma(c.im, fma(T(-3), fnn.re, fo.re), c.re * vo.re),
                                 fma(c.im, fma(T(-3), fnn.im, fo.im), c.re * vo.im));
                if (r == 0 && forcing) {
                    // 3-mode forcing (Burger.py:410-421) on k = 1, 2, 3 = lanes 1..3: dt F / (1 + C) = 2 cfo F (other lanes: weight 0)
                    const int m = lane - 1 < 0 ? 0 : (lane - 1 > 2 ? 2 : lane - 1);
                    const Cx<T> F = ldcx(prm.fcoef + (fc_row * prm.stepper + iout % prm.stepper) * 3 + m);
                    const T wgt = (lane >= 1 && lane <= 3) ? T(2) * c.im : T(0);
                    vn = cx<T>(fma(wgt, F.re, vn.re), fma(wgt, F.im, vn.im));
                }
                if constexpr (VS) stcx(V + 32 * r + lane, vn); else v[r] = vn;
                fre[r] = (float)vn.re;
                fim[r] = (float)vn.im;
                stcx(FN + 32 * r + lane, fnn);
            }
            iout += 1;
            tnow += dt;
            if (lane == 0) {
                // Nyquist mode: Fn imaginary, no forcing there; Im v[0] is a constant of the motion (k = 0: Fn = 0).
                // Touches shared memory only, so the divergence ends with the block.
                const Cx<T> c = CC[H], vN = FN[H + 1];
                const T fnnN = prm.kwave[H] * XN;
                const Cx<T> vn = cx<T>(c.re * vN.re, fma(c.im, fma(T(-3), fnnN, FN[H].im), c.re * vN.im));
                FN[H + 1] = vn;
                FN[H] = cx<T>(T(0), fnnN);
                const float fNre = (float)vn.re, fNim = (float)vn.im;
                bad |= !(fabsf(fNre) <= FLT_MAX && fabsf(fNim) <= FLT_MAX);
                acc[H] = __fadd_rn(acc[H], ek_row_f32(fNre, fNim, N, dxf));
            }
            __syncwarp();
            // ---- float32 spectrum chain (Q6) + blow-up detection on the complex64 cast (Burger.py:498) -------------
            const bool write_hist = prm.hist_rows > 0 && iout < prm.hist_rows;
            const int64_t hrow = e * prm.hist_rows + iout;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                bad |= !(fabsf(fre[r]) <= FLT_MAX && fabsf(fim[r]) <= FLT_MAX);
                const float a = __fadd_rn(acc[32 * r + lane], ek_row_f32(fre[r], fim[r], N, dxf));
                acc[32 * r + lane] = a;
            }
            if (write_hist) live = live && !__any_sync(0xffffffffu, bad);
            if (write_hist && live) {
                // Ek_ktt = cumsum / (i + 1) (Burger.py:555): one reciprocal per row, then the quotient is corrected with the
                // exact remainder (q0 = a r, q = q0 + (a - q0 b) r: the correctly rounded a / b for these operands) -- 3 FMAs
                // per entry instead of a division
                const double bdiv = (double)(iout + 1), rdiv = 1.0 / bdiv;
                auto quot = [&](float a_) {
                    const double a = (double)a_, q0 = a * rdiv;
                    return fma(fma(-q0, bdiv, a), rdiv, q0);
                };
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const int k = 32 * r + lane;
                    if (prm.vv_hist) {
                        Cx<float> c; c.re = fre[r]; c.im = fim[r];
                        prm.vv_hist[hrow * N + k] = c;
                        if (k != 0) { c.im = -c.im; prm.vv_hist[hrow * N + N - k] = c; }
                    }
                    if (prm.ektt_hist) prm.ektt_hist[hrow * NH + k] = quot(acc[k]);
                }
                if (lane == 0) {
                    const Cx<T> vN = FN[H + 1];
                    if (prm.vv_hist) { Cx<float> c; c.re = (float)vN.re; c.im = (float)vN.im; prm.vv_hist[hrow * N + H] = c; }
                    if (prm.ektt_hist) prm.ektt_hist[hrow * NH + H] = quot(acc[H]);
                }
            }
            // ---- u = Re ifft(v)  (Burger.py:491) ------------------------------------------------------------------
            to_real();
            if (write_hist && live && prm.uu_hist) {
#pragma unroll
                for (int r = 0; r < 16; ++r)
                    stcx(reinterpret_cast<Cx<T>*>(prm.uu_hist + hrow * N) + 32 * r + lane, cx<T>(z[r].re * invN, z[r].im * invN));
            }
        }
        // ---- epilogue --------------------------------------------------------------------------------------------
        if (nsub > 0) {
            const bool blew = __any_sync(0xffffffffu, bad);
            if (was_live && blew && lane == 0) prm.status[e] = 1;
            live = live && !blew;
            if (live) {
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    const int k = 32 * r + lane;
                    stcx(prm.v + e * NH + k, vget(r));
                    stcx(prm.fn + e * NH + k, FN[k]);
                }
                if (lane == 0) {
                    stcx(prm.v + e * NH + H, FN[H + 1]);
                    stcx(prm.fn + e * NH + H, FN[H]);
                    prm.iout[e] = iout;
                    prm.tnow[e] = tnow;
                }
            }
            // the running spectrum sums are kept even for an environment that blew up within this call (it is dead anyway)
#pragma unroll
            for (int r = 0; r < 16; ++r) prm.acc[e * NH + 32 * r + lane] = acc[32 * r + lane];
            if (lane == 0) prm.acc[e * NH + H] = acc[H];
        }
    }

    // W32^r = exp(-2 pi i r / 32), r = 0..15 (compile-time constants after unrolling)
    __device__ __forceinline__ static Cx<T> w32(int r) {
        constexpr double c[16] = {1.0, 0.98078528040323044913, 0.92387953251128675613, 0.83146961230254523708, 0.70710678118654752440,
                                  0.55557023301960222474, 0.38268343236508977173, 0.19509032201612826785, 0.0, -0.19509032201612826785,
                                  -0.38268343236508977173, -0.55557023301960222474, -0.70710678118654752440, -0.83146961230254523708,
                                  -0.92387953251128675613, -0.98078528040323044913};
        constexpr double s[16] = {0.0, 0.19509032201612826785, 0.38268343236508977173, 0.55557023301960222474, 0.70710678118654752440,
                                  0.83146961230254523708, 0.92387953251128675613, 0.98078528040323044913, 1.0, 0.98078528040323044913,
                                  0.92387953251128675613, 0.83146961230254523708, 0.70710678118654752440, 0.55557023301960222474,
                                  0.38268343236508977173, 0.19509032201612826785};
        return cx<T>(T(c[r]), T(-s[r]));
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// Two warps per environment (CTA of 64 threads): 512 = 8 x 8 x 8.  Each thread holds 8 complex points (T layout
// j = 64 r + t, F layout k = 64 r + t); three 8-point transforms in registers, two twiddles from shared-memory tables, two
// conflict-free transposes through shared memory, no shuffles.  Twice the warps per SM sub-partition of the one-warp kernel
// (1.7 instead of 0.9) and ~100 registers per thread instead of 255: the one-warp kernel is bound by its own latencies.
// Same arithmetic per wavenumber as Dns1024 (split step, ABCN update, spectrum chain, history rows).
template <typename T>
struct Dns1024x2 {
    using D1 = Dns1024<T, true>;
    static constexpr int N = 1024, H = 512, NH = 513, NT = 64;
    static constexpr int ROW = 65;                       // transpose 1: element (k1, n2) at k1 * 65 + n2
    // shared memory per CTA in complex words: exchange 8 x 65 | T1 8 x 64 | T2 8 x 8 | (cv, cfo) 514 | Fn_old 514 | v 512 | acc (float)
    static constexpr int CX_WORDS = 8 * ROW + 512 + 64 + 514 + 514 + 512;
    static size_t smem_bytes() { return sizeof(Cx<T>) * CX_WORDS + sizeof(float) * 516; }

    // 8-point transform in registers, natural order in and out: n = 4 n1 + n2, k = k1 + 2 k2
    template <bool INV>
    __device__ __forceinline__ static void fft8(Cx<T> (&x)[8]) {
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) {
            const Cx<T> a = x[n2] + x[4 + n2], b = x[n2] - x[4 + n2];
            x[n2] = a;
            x[4 + n2] = b;                                                   // slot 4 k1 + n2
        }
        x[5] = D1::template tw16<2, INV>(x[5]);                              // W8^(n2 k1): W8 = W16^2
        x[6] = D1::template tw16<4, INV>(x[6]);
        x[7] = D1::template tw16<6, INV>(x[7]);
        D1::template dft4<INV>(x[0], x[1], x[2], x[3]);                      // k1 = 0 -> X[0], X[2], X[4], X[6]
        D1::template dft4<INV>(x[4], x[5], x[6], x[7]);                      // k1 = 1 -> X[1], X[3], X[5], X[7]
        const Cx<T> y1 = x[4], y2 = x[1], y3 = x[5], y4 = x[2], y5 = x[6], y6 = x[3];
        x[1] = y1; x[2] = y2; x[3] = y3; x[4] = y4; x[5] = y5; x[6] = y6;   // out[k1 + 2 k2] = slot[4 k1 + k2]
    }

    // W16^r (forward), r = 0..7
    __device__ __forceinline__ static Cx<T> w16(int r) {
        constexpr double c[8] = {1.0, 0.92387953251128673848, 0.70710678118654752440, 0.38268343236508978178, 0.0,
                                 -0.38268343236508978178, -0.70710678118654752440, -0.92387953251128673848};
        constexpr double s[8] = {0.0, 0.38268343236508978178, 0.70710678118654752440, 0.92387953251128673848, 1.0,
                                 0.92387953251128673848, 0.70710678118654752440, 0.38268343236508978178};
        return cx<T>(T(c[r]), T(-s[r]));
    }

    // forward: T layout -> F layout; inverse: F -> T (unnormalised).  All 64 threads of the CTA call it.
    template <bool INV>
    __device__ __forceinline__ static void fft512(Cx<T> (&z)[8], Cx<T>* E, const Cx<T> (&w1)[7], const Cx<T> (&w2)[7], int t) {
        // pass-B thread roles: b + 8 k1 for the first in-register pass over a, k1 + 8 c for the second over b
        const int bB = t & 7, kB = t >> 3;           // t = b + 8 k1
        const int kC = t & 7, cC = t >> 3;           // t = k1 + 8 c
        if constexpr (!INV) {
            fft8<false>(z);                                                          // over r (n1) -> k1
#pragma unroll
            for (int r = 1; r < 8; ++r) z[r] = cmul(z[r], w1[r - 1]);               // W512^(t k1)
            __syncthreads();
#pragma unroll
            for (int r = 0; r < 8; ++r) stcx(E + r * ROW + t, z[r]);                // (k1 = r, n2 = t)
            __syncthreads();
#pragma unroll
            for (int a = 0; a < 8; ++a) z[a] = ldcx(E + kB * ROW + 8 * a + bB);     // thread (b, k1): n2 = 8 a + b
            fft8<false>(z);                                                          // over a -> c
#pragma unroll
            for (int c = 1; c < 8; ++c) z[c] = cmul(z[c], w2[c - 1]);               // W64^(b c)
            __syncthreads();
#pragma unroll
            for (int c = 0; c < 8; ++c) stcx(E + 64 * c + 8 * kB + ((bB + kB) & 7), z[c]);      // (c, k1, b), b swizzled by k1
            __syncthreads();
#pragma unroll
            for (int b = 0; b < 8; ++b) z[b] = ldcx(E + 64 * cC + 8 * kC + ((b + kC) & 7));     // thread (k1, c)
            fft8<false>(z);                                                          // over b -> d: k = k1 + 8 c + 64 d = t + 64 d
        } else {
            fft8<true>(z);                                                           // over d -> b
            __syncthreads();
#pragma unroll
            for (int b = 0; b < 8; ++b) stcx(E + 64 * cC + 8 * kC + ((b + kC) & 7), z[b]);
            __syncthreads();
#pragma unroll
            for (int c = 0; c < 8; ++c) z[c] = ldcx(E + 64 * c + 8 * kB + ((bB + kB) & 7));
#pragma unroll
            for (int c = 1; c < 8; ++c) z[c] = cmulc(z[c], w2[c - 1]);
            fft8<true>(z);                                                           // over c -> a
            __syncthreads();
#pragma unroll
            for (int a = 0; a < 8; ++a) stcx(E + kB * ROW + 8 * a + bB, z[a]);
            __syncthreads();
#pragma unroll
            for (int r = 0; r < 8; ++r) z[r] = ldcx(E + r * ROW + t);
#pragma unroll
            for (int r = 1; r < 8; ++r) z[r] = cmulc(z[r], w1[r - 1]);
            fft8<true>(z);                                                           // over k1 -> r
        }
    }

    __device__ static void run(const SpectralParams<T>& prm, unsigned char* smem_raw) {
        const int t = threadIdx.x;
        const int64_t e = blockIdx.x;
        Cx<T>* E = reinterpret_cast<Cx<T>*>(smem_raw);
        Cx<T>* T1 = E + 8 * ROW;
        Cx<T>* T2 = T1 + 512;
        Cx<T>* CC = T2 + 64;
        Cx<T>* FN = CC + 514;
        Cx<T>* V = FN + 514;
        float* acc = reinterpret_cast<float*>(V + 512);
        const int flags = prm.flags;
        const T dt = prm.dt, invN = T(1) / T(N);
        const T nu = prm.nu[e];
        const float dxf = (float)prm.dx;
        auto tw1024 = [&](int j) {
            const Cx<T> w = ldcx(prm.tw + (j & 511));
            return (j & 512) ? cx<T>(-w.re, -w.im) : w;
        };
        // both twiddle sets and the Crank-Nicolson factors live in registers (148 -> ~230 of the 255 a 64-thread CTA may use
        // with 4 CTAs per SM): the kernel is bound by shared-memory wavefronts + FP64 issue, and these were 30 % of the former
        Cx<T> w1[7], w2[7], cc[8];
#pragma unroll
        for (int r = 1; r < 8; ++r) {
            w1[r - 1] = tw1024((2 * r * t) & 1023);                       // W512^(t r)
            w2[r - 1] = tw1024((16 * (t & 7) * r) & 1023);                // W64^(b c), b = t & 7 (pass-B role), c = r
        }
        const T kw0 = prm.kwave[t], dk = prm.kwave[64];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int k = 64 * r + t;
            const T kwr = prm.kwave[k];
            const T C = T(0.5) * (kwr * kwr) * nu * dt;
            const T rr = T(1) / (T(1) + C);
            cc[r] = cx<T>((T(1) - C) * rr, T(0.5) * dt * rr);
            acc[k] = prm.acc[e * NH + k];
            V[k] = ldcx(prm.v + e * NH + k);
            FN[k] = ldcx(prm.fn + e * NH + k);
        }
        if (t == 0) {
            const T kwN = prm.kwave[H];
            const T C = T(0.5) * (kwN * kwN) * nu * dt;
            const T rr = T(1) / (T(1) + C);
            CC[H] = cx<T>((T(1) - C) * rr, T(0.5) * dt * rr);
            acc[H] = prm.acc[e * NH + H];
            FN[H] = ldcx(prm.fn + e * NH + H);
            FN[H + 1] = ldcx(prm.v + e * NH + H);
        }
        const Cx<T> wbase = tw1024(t);                                                        // W1024^t; W1024^k = wbase W16^r
        const T hs = T(0.25) * invN * invN;
        const bool was_live = prm.status[e] == 0;
        bool live = was_live, bad = false;
        int iout = prm.iout[e];
        T tnow = prm.tnow[e];
        Cx<T> z[8];
        __syncthreads();

        auto to_real = [&]() {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int k = 64 * r + t;
                const Cx<T> wk = cmul(wbase, w16(r));
                const Cx<T> vr = ldcx(V + k), vm = ldcx(V + ((512 - k) & 511));
                const Cx<T> Ee = cx<T>(vr.re + vm.re, vr.im - vm.im);
                const Cx<T> Dd = cx<T>(vr.re - vm.re, vr.im + vm.im);
                z[r] = cx<T>(fma(-Dd.im, wk.re, fma(Dd.re, wk.im, Ee.re)), fma(Dd.re, wk.re, fma(Dd.im, wk.im, Ee.im)));
            }
            {
                const T vNre = FN[H + 1].re, v0re = V[0].re;
                const Cx<T> dc = cx<T>(v0re + vNre, v0re - vNre);
                z[0] = t == 0 ? dc : z[0];
            }
            fft512<true>(z, E, w1, w2, t);
        };
        to_real();

        const int nsub = (flags & F_NO_ADVANCE) ? 0 : prm.nsub;
        const int64_t fc_row = (flags & F_FORCING_PER_ENV) ? e : 0;
        const bool forcing = (flags & F_FORCING) != 0;
        for (int it = 0; it < nsub; ++it) {
            if (it == nsub - 1) {
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    stcx(reinterpret_cast<Cx<T>*>(prm.uprev + e * N) + 64 * r + t, cx<T>(z[r].re * invN, z[r].im * invN));
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) z[r] = cx<T>(z[r].re * z[r].re, z[r].im * z[r].im);
            fft512<false>(z, E, w1, w2, t);
            // Hermitian partners through the exchange buffer (natural order)
            __syncthreads();
#pragma unroll
            for (int r = 0; r < 8; ++r) stcx(E + 64 * r + t, z[r]);
            __syncthreads();
            const T XN = (z[0].re - z[0].im) * (T(2) * hs);                    // thread 0: fft(u^2/2)[N/2]
            float fre[8], fim[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int k = 64 * r + t;
                const Cx<T> wk = cmul(wbase, w16(r));
                const Cx<T> zm = ldcx(E + ((512 - k) & 511));
                const Cx<T> Ee = cx<T>(z[r].re + zm.re, z[r].im - zm.im);
                const Cx<T> Oo = cx<T>(z[r].im + zm.im, zm.re - z[r].re);
                const Cx<T> X = cx<T>(hs * (Ee.re + fma(wk.re, Oo.re, -(wk.im * Oo.im))), hs * (Ee.im + fma(wk.re, Oo.im, wk.im * Oo.re)));
                const T kwr = fma(T(r), dk, kw0);
                const Cx<T> fnn = cx<T>(-kwr * X.im, kwr * X.re);
                const Cx<T> c = cc[r];
                const Cx<T> fo = ldcx(FN + k);
                const Cx<T> vo = ldcx(V + k);
                Cx<T> vn = cx<T>(fma(c.im, fma(T(-3), fnn.re, fo.re), c.re * vo.re),
                                 fma(c.im, fma(T(-3), fnn.im, fo.im), c.re * vo.im));
                if (r == 0 && forcing) {
                    const int m = t - 1 < 0 ? 0 : (t - 1 > 2 ? 2 : t - 1);
                    const Cx<T> F = ldcx(prm.fcoef + (fc_row * prm.stepper + iout % prm.stepper) * 3 + m);
                    const T wgt = (t >= 1 && t <= 3) ? T(2) * c.im : T(0);
                    vn = cx<T>(fma(wgt, F.re, vn.re), fma(wgt, F.im, vn.im));
                }
                fre[r] = (float)vn.re;
                fim[r] = (float)vn.im;
                // V[k] is read by the partner thread too (as v[512 - k] in to_real only, after the barrier below): safe to overwrite
                stcx(V + k, vn);
                stcx(FN + k, fnn);
            }
            iout += 1;
            tnow += dt;
            if (t == 0) {
                const Cx<T> c = CC[H], vN = FN[H + 1];
                const T fnnN = prm.kwave[H] * XN;
                const Cx<T> vn = cx<T>(c.re * vN.re, fma(c.im, fma(T(-3), fnnN, FN[H].im), c.re * vN.im));
                FN[H + 1] = vn;
                FN[H] = cx<T>(T(0), fnnN);
                const float fNre = (float)vn.re, fNim = (float)vn.im;
                bad |= !(fabsf(fNre) <= FLT_MAX && fabsf(fNim) <= FLT_MAX);
                acc[H] = __fadd_rn(acc[H], ek_row_f32(fNre, fNim, N, dxf));
            }
            const bool write_hist = prm.hist_rows > 0 && iout < prm.hist_rows;
            const int64_t hrow = e * prm.hist_rows + iout;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                bad |= !(fabsf(fre[r]) <= FLT_MAX && fabsf(fim[r]) <= FLT_MAX);
                const float a = __fadd_rn(acc[64 * r + t], ek_row_f32(fre[r], fim[r], N, dxf));
                acc[64 * r + t] = a;
            }
            // one barrier: the new v / Fn_old / Nyquist values are visible, and (with history) the blow-up vote
            if (write_hist) live = live && !__syncthreads_or(bad);
            else __syncthreads();
            if (write_hist && live) {
                const double bdiv = (double)(iout + 1), rdiv = 1.0 / bdiv;
                auto quot = [&](float a_) {
                    const double a = (double)a_, q0 = a * rdiv;
                    return fma(fma(-q0, bdiv, a), rdiv, q0);
                };
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int k = 64 * r + t;
                    if (prm.vv_hist) {
                        Cx<float> c; c.re = fre[r]; c.im = fim[r];
                        prm.vv_hist[hrow * N + k] = c;
                        if (k != 0) { c.im = -c.im; prm.vv_hist[hrow * N + N - k] = c; }
                    }
                    if (prm.ektt_hist) prm.ektt_hist[hrow * NH + k] = quot(acc[k]);
                }
                if (t == 0) {
                    const Cx<T> vN = FN[H + 1];
                    if (prm.vv_hist) { Cx<float> c; c.re = (float)vN.re; c.im = (float)vN.im; prm.vv_hist[hrow * N + H] = c; }
                    if (prm.ektt_hist) prm.ektt_hist[hrow * NH + H] = quot(acc[H]);
                }
            }
            to_real();
            if (write_hist && live && prm.uu_hist) {
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    stcx(reinterpret_cast<Cx<T>*>(prm.uu_hist + hrow * N) + 64 * r + t, cx<T>(z[r].re * invN, z[r].im * invN));
            }
        }
        if (nsub > 0) {
            const bool blew = __syncthreads_or(bad);
            if (was_live && blew && t == 0) prm.status[e] = 1;
            live = live && !blew;
            if (live) {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const int k = 64 * r + t;
                    stcx(prm.v + e * NH + k, V[k]);
                    stcx(prm.fn + e * NH + k, FN[k]);
                }
                if (t == 0) {
                    stcx(prm.v + e * NH + H, FN[H + 1]);
                    stcx(prm.fn + e * NH + H, FN[H]);
                    prm.iout[e] = iout;
                    prm.tnow[e] = tnow;
                }
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) prm.acc[e * NH + 64 * r + t] = acc[64 * r + t];
            if (t == 0) prm.acc[e * NH + H] = acc[H];
        }
    }
};

}  // namespace mpde
