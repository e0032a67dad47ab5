// Warp-level FFTs on register-resident data (sm_100a).
//
// WarpFFT<T, H>: radix-2 complex FFT of H <= 128 points.  A "team" of TS = min(H, 32)
// lanes owns one sequence; lane tl holds the P = H/TS points j = p*TS + tl.  Forward =
// decimation-in-frequency (natural order in, BIT-REVERSED order out): the first log2(P)
// stages pair registers of one lane, the last log2(TS) stages pair lanes through
// __shfl_xor.  Inverse = the exact transpose (decimation-in-time, bit-reversed in, natural
// out, unnormalised).  Spectral arithmetic is done in the bit-reversed layout, so no
// permutation pass is ever needed.
//
// RealFFT<T, N>: transform of ONE real length-N field per team through a complex FFT of
// H = N/2 points on z_j = x_{2j} + i x_{2j+1} plus the usual split/merge step.  The team
// then holds exactly the half spectrum X_0..X_{H-1} (one wavenumber per register, bit-
// reversed order) and the real Nyquist value X_H on its first lane.  Every environment's
// arithmetic is independent of its neighbours in the warp, which is what makes results
// bitwise invariant to batch size and packing.
#pragma once
#include "common.cuh"

namespace mpde {

__host__ __device__ constexpr int brev_bits(int x, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}

// TS_: lanes per sequence (team size); TWS: stride into the twiddle table (table holds
// exp(-2 pi i j / (H*TWS)))
template <typename T, int H, int TS_ = (H < 32 ? H : 32), int TWS = 1>
struct WarpFFT {
    static constexpr int TS = TS_;
    static constexpr int P = H / TS;
    static constexpr int SMEM_CX = 0;   // complex words of shared memory per team (none: shuffles only)
    static constexpr bool NATURAL = false;   // wavenumber of (register p, lane t) is NOT t + TS p (bit-reversed layout)
    static_assert(TS >= 1 && TS <= 32 && (TS & (TS - 1)) == 0 && TS <= H, "team size: power of two, <= 32, <= H");
    static constexpr int LOGH = ilog2(H);
    static constexpr int LOGTS = ilog2(TS);
    static constexpr int LOGP = ilog2(P);
    static_assert((1 << LOGH) == H && H >= 2 && H <= 128, "warp FFT handles 2..128 points, power of two");

    int tl;                         // lane within the team
    int base;                       // first lane of the team within the warp
    unsigned tmask;                 // lanes of this team
    unsigned smask;                 // mask named by every shuffle / __syncwarp: the team (default: teams never depend on each
                                    // other's control flow, a team may be idle or dead) or the whole warp (whole_warp())
    Cx<T> wx[LOGTS > 0 ? LOGTS : 1];  // cross-lane twiddles, (1,0) on the lower lane of a pair
    Cx<T> wl[P > 1 ? P - 1 : 1];    // in-register twiddles
    int part[P];                    // lane holding wavenumber -k (mod H) for register pp(p)

    // register index that holds -k for the k held in register p (depends on p only)
    __host__ __device__ static constexpr int pp(int p) {
        return brev_bits((P - brev_bits(p, LOGP)) % P, LOGP);
    }
    // wavenumber index (0..H-1) held at register p of team-lane t after fwd()
    __device__ __forceinline__ static int kidx(int p, int t) {
        return (int)(__brev((unsigned)(p * TS + t)) >> (32 - LOGH));
    }

    __device__ __forceinline__ void init(const Cx<T>* __restrict__ tw, Cx<T>* /*team_smem*/ = nullptr) {
        const int lane = threadIdx.x & 31;
        tl = lane & (TS - 1);
        base = lane & ~(TS - 1);
        tmask = TS == 32 ? 0xffffffffu : (((1u << TS) - 1u) << base);
        smask = tmask;
#pragma unroll
        for (int s = 0; s < LOGTS; ++s) {
            const int h = TS >> (s + 1);
            wx[s] = (tl & h) ? ldcx(tw + (tl & (h - 1)) * (H / (2 * h)) * TWS) : cx<T>(T(1), T(0));
        }
        if constexpr (P > 1) {
#pragma unroll
            for (int hp = P / 2; hp >= 1; hp >>= 1)
#pragma unroll
                for (int q = 0; q < hp; ++q)
                    wl[P - 2 * hp + q] = ldcx(tw + (q * TS + tl) * (P / (2 * hp)) * TWS);
        }
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int k = kidx(p, tl);
            const int kneg = (H - k) & (H - 1);
            const int npos = (int)(__brev((unsigned)kneg) >> (32 - LOGH));   // position holding -k
            part[p] = base + (npos & (TS - 1));
        }
    }

    // natural order -> bit-reversed spectrum, unnormalised forward DFT
    __device__ __forceinline__ void fwd(Cx<T> (&z)[P]) const {
        if constexpr (P > 1) {
#pragma unroll
            for (int hp = P / 2; hp >= 1; hp >>= 1)
#pragma unroll
                for (int p = 0; p < P; ++p)
                    if ((p & hp) == 0) {
                        const Cx<T> a = z[p], b = z[p + hp];
                        z[p] = a + b;
                        z[p + hp] = cmul(a - b, wl[P - 2 * hp + (p & (hp - 1))]);
                    }
        }
#pragma unroll
        for (int s = 0; s < LOGTS; ++s) {
            const int h = TS >> (s + 1);
            const T sg = (tl & h) ? T(-1) : T(1);
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const Cx<T> o = shfl_xor(z[p], h, smask);
                const Cx<T> t = cx<T>(fma(sg, z[p].re, o.re), fma(sg, z[p].im, o.im));
                z[p] = (h == 1) ? t : cmul(t, wx[s]);          // last stage: twiddle is 1
            }
        }
    }

    // two independent sequences at once, interleaved stage by stage (instruction-level parallelism
    // hides the ~37-cycle shuffle and ~13-cycle FP64 latencies)
    __device__ __forceinline__ void fwd2(Cx<T> (&za)[P], Cx<T> (&zb)[P]) const {
        if constexpr (P > 1) {
#pragma unroll
            for (int hp = P / 2; hp >= 1; hp >>= 1)
#pragma unroll
                for (int p = 0; p < P; ++p)
                    if ((p & hp) == 0) {
                        const Cx<T> w = wl[P - 2 * hp + (p & (hp - 1))];
                        const Cx<T> a = za[p], b = za[p + hp], c = zb[p], d = zb[p + hp];
                        za[p] = a + b;
                        zb[p] = c + d;
                        za[p + hp] = cmul(a - b, w);
                        zb[p + hp] = cmul(c - d, w);
                    }
        }
#pragma unroll
        for (int s = 0; s < LOGTS; ++s) {
            const int h = TS >> (s + 1);
            const T sg = (tl & h) ? T(-1) : T(1);
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const Cx<T> oa = shfl_xor(za[p], h, smask);
                const Cx<T> ob = shfl_xor(zb[p], h, smask);
                const Cx<T> ta = cx<T>(fma(sg, za[p].re, oa.re), fma(sg, za[p].im, oa.im));
                const Cx<T> tb = cx<T>(fma(sg, zb[p].re, ob.re), fma(sg, zb[p].im, ob.im));
                za[p] = (h == 1) ? ta : cmul(ta, wx[s]);
                zb[p] = (h == 1) ? tb : cmul(tb, wx[s]);
            }
        }
    }

    // bit-reversed spectrum -> natural order, unnormalised inverse DFT (H * ifft)
    __device__ __forceinline__ void inv(Cx<T> (&z)[P]) const {
#pragma unroll
        for (int s = LOGTS - 1; s >= 0; --s) {
            const int h = TS >> (s + 1);
            const T sg = (tl & h) ? T(-1) : T(1);
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const Cx<T> m = (h == 1) ? z[p] : cmulc(z[p], wx[s]);
                const Cx<T> o = shfl_xor(m, h, smask);
                z[p] = cx<T>(fma(sg, m.re, o.re), fma(sg, m.im, o.im));
            }
        }
        if constexpr (P > 1) {
#pragma unroll
            for (int hp = 1; hp <= P / 2; hp <<= 1)
#pragma unroll
                for (int p = 0; p < P; ++p)
                    if ((p & hp) == 0) {
                        const Cx<T> a = z[p];
                        const Cx<T> b = cmulc(z[p + hp], wl[P - 2 * hp + (p & (hp - 1))]);
                        z[p] = a + b;
                        z[p + hp] = a - b;
                    }
        }
    }

    // value held for wavenumber -k (mod H) by the team, for the k of register p
    __device__ __forceinline__ Cx<T> mirrored(const Cx<T> (&z)[P], int p) const {
        if constexpr (TS == 1) return z[pp(p)];
        else return shfl(z[pp(p)], part[p], smask);
    }
};

// 16 points on 4 lanes x 4 registers as a 4 x 4 Cooley-Tukey transform: an in-register radix-4 pass, twiddles, a
// TRANSPOSE through shared memory (4 STS.128 + 4 LDS.128 per lane, bank-conflict free) and a second in-register
// radix-4 pass.  Against the shuffle butterflies above this moves each value across lanes once instead of
// log2(TS) times and needs no multiplies in the passes: per lane 44 FP64 + 8 LSU instructions instead of
// ~90 FP64 + 32 SHFL.  Point j = 4 p + tl lives in register p of lane tl; wavenumber k = tl + 4 d in register d
// (NATURAL order in and out).  The team's exchange area is SMEM_CX complex words; a team stride of
// 64 (mod 128) bytes keeps the two teams of a quarter-warp on disjoint banks.
template <typename T, int TWS>
struct WarpFFT<T, 16, 4, TWS> {
    static constexpr int H = 16, TS = 4, P = 4;
    static constexpr int SMEM_CX = 32;  // two 16-entry slots (fwd2 transforms two sequences at once)
    static constexpr bool NATURAL = true;    // k = t + 4 p
    int tl, base;
    unsigned tmask, smask;
    Cx<T> wl[3];                        // W16^(tl * r), r = 1..3
    Cx<T>* sm;

    __device__ __forceinline__ static int kidx(int p, int t) { return t + 4 * p; }

    __device__ __forceinline__ void init(const Cx<T>* __restrict__ tw, Cx<T>* team_smem) {
        const int lane = threadIdx.x & 31;
        tl = lane & 3;
        base = lane & ~3;
        tmask = 0xfu << base;
        smask = tmask;
        sm = team_smem;
#pragma unroll
        for (int r = 1; r < 4; ++r) {           // the table holds W16^m for m < 8; W16^(m+8) = -W16^m  (tl * r <= 9)
            const int m = tl * r;
            const Cx<T> w = ldcx(tw + (m & 7) * TWS);
            wl[r - 1] = m >= 8 ? cx<T>(-w.re, -w.im) : w;
        }
    }

    // 4-point DFT of the registers; INV = conjugate twiddles (W4 = +i)
    template <bool INV>
    __device__ __forceinline__ static void dft4(Cx<T> (&x)[4]) {
        const Cx<T> a0 = x[0] + x[2], a1 = x[0] - x[2], a2 = x[1] + x[3], a3 = x[1] - x[3];
        x[0] = a0 + a2;
        x[2] = a0 - a2;
        // forward: y1 = a1 - i a3, y3 = a1 + i a3
        const Cx<T> m = cx<T>(a1.re + a3.im, a1.im - a3.re), q = cx<T>(a1.re - a3.im, a1.im + a3.re);
        x[1] = INV ? q : m;
        x[3] = INV ? m : q;
    }
    // element (lane a, register b) -> (lane b, register a); entry of (destination lane b, register a) is
    // 4 b + ((a + b) & 3): writes of one register fill one 64-byte block, reads of one register hit 4 distinct
    // bank groups
    __device__ __forceinline__ void put(const Cx<T> (&z)[4], int slot) const {
#pragma unroll
        for (int b = 0; b < 4; ++b) stcx(sm + 16 * slot + 4 * b + ((tl + b) & 3), z[b]);
    }
    __device__ __forceinline__ void get(Cx<T> (&z)[4], int slot) const {
#pragma unroll
        for (int a = 0; a < 4; ++a) z[a] = ldcx(sm + 16 * slot + 4 * tl + ((a + tl) & 3));
    }
    template <bool INV>
    __device__ __forceinline__ void twiddle(Cx<T> (&z)[4]) const {
#pragma unroll
        for (int r = 1; r < 4; ++r) z[r] = INV ? cmulc(z[r], wl[r - 1]) : cmul(z[r], wl[r - 1]);
    }
    template <bool INV>
    __device__ __forceinline__ void run(Cx<T> (&z)[4]) const {
        dft4<INV>(z);
        twiddle<INV>(z);
        __syncwarp(smask);
        put(z, 0);
        __syncwarp(smask);
        get(z, 0);
        dft4<INV>(z);
    }
    __device__ __forceinline__ void fwd(Cx<T> (&z)[4]) const { run<false>(z); }
    __device__ __forceinline__ void inv(Cx<T> (&z)[4]) const { run<true>(z); }
    __device__ __forceinline__ void fwd2(Cx<T> (&za)[4], Cx<T> (&zb)[4]) const {
        dft4<false>(za);
        dft4<false>(zb);
        twiddle<false>(za);
        twiddle<false>(zb);
        __syncwarp(smask);
        put(za, 0);
        put(zb, 1);
        __syncwarp(smask);
        get(za, 0);
        get(zb, 1);
        dft4<false>(za);
        dft4<false>(zb);
    }
    // value held for wavenumber -k (mod 16): k = tl + 4 p -> lane (4 - tl) & 3, register 3 - p (tl > 0) or (4 - p) & 3
    __device__ __forceinline__ Cx<T> mirrored(const Cx<T> (&z)[4], int p) const {
        const Cx<T> far = shfl(z[3 - p], base + ((4 - tl) & 3), smask);
        return tl == 0 ? z[(4 - p) & 3] : far;
    }
};

// 16 points on 8 lanes x 2 registers with EXACTLY the arithmetic of the 4 x 4 transform above (a radix-2^2
// factorisation: the same two radix-4 passes, their butterflies spread over one in-register and three shuffle
// stages, the same single twiddle W16^(t c) between the passes).  Every add, subtract and multiply has the same
// operands in the same order, so both variants produce bit-identical spectra and fields -- which is what lets the
// host choose the team size from the batch size without changing any result.
// Layout: point j = 8 p + tl in register p (as the generic transform); wavenumber k = c + 4 d with
// c = 2 (tl >> 2) + q in register q, d = bitrev2(tl & 3).
// Selected with the team-size tag -8 (the plain 8 keeps the faster radix-2 shuffle network above, whose rounding differs).
template <typename T, int TWS>
struct WarpFFT<T, 16, -8, TWS> {
    static constexpr int H = 16, TS = 8, P = 2;
    static constexpr int SMEM_CX = 0;
    static constexpr bool NATURAL = false;
    int tl, base;
    unsigned tmask, smask;
    bool lo4, b1, b0;                   // tl < 4, bit 1 and bit 0 of tl
    Cx<T> wa, wb;                       // W16^(t' c0), W16^(t' (c0 + 1)): t' = tl & 3, c0 = 2 (tl >> 2)
    int part[2];

    __device__ __forceinline__ static int brev2(int x) { return ((x & 1) << 1) | ((x >> 1) & 1); }
    __device__ __forceinline__ static int kidx(int q, int t) { return 2 * (t >> 2) + q + 4 * brev2(t & 3); }

    __device__ __forceinline__ static Cx<T> tw16(const Cx<T>* __restrict__ tw, int m) {      // W16^m, m <= 9
        const Cx<T> w = ldcx(tw + (m & 7) * TWS);
        return m >= 8 ? cx<T>(-w.re, -w.im) : w;
    }
    __device__ __forceinline__ void init(const Cx<T>* __restrict__ tw, Cx<T>* /*team_smem*/ = nullptr) {
        const int lane = threadIdx.x & 31;
        tl = lane & 7;
        base = lane & ~7;
        tmask = 0xffu << base;
        smask = tmask;
        lo4 = tl < 4;
        b1 = (tl & 2) != 0;
        b0 = (tl & 1) != 0;
        const int tp = tl & 3, c0 = 2 * (tl >> 2);
        wa = tw16(tw, tp * c0);
        wb = tw16(tw, tp * (c0 + 1));
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int kn = (16 - kidx(q, tl)) & 15, cn = kn & 3, dn = kn >> 2;     // (cn & 1) == q
            part[q] = base + 4 * (cn >> 1) + brev2(dn);
        }
    }
    // r <- (lower lane of the pair ? r + o : o - r), the pair being lanes that differ in the bit whose sign is sg
    __device__ __forceinline__ static Cx<T> pm(T sg, Cx<T> r, Cx<T> o) { return cx<T>(fma(sg, r.re, o.re), fma(sg, r.im, o.im)); }
    // a1 -/+ i a3: sg = +1 -> (a1.re + a3.im, a1.im - a3.re) = a1 - i a3;  sg = -1 -> a1 + i a3
    __device__ __forceinline__ static Cx<T> rot(T sg, Cx<T> a1, Cx<T> a3) { return cx<T>(fma(sg, a3.im, a1.re), fma(-sg, a3.re, a1.im)); }

    template <bool INV>
    __device__ __forceinline__ void first_pass(Cx<T> (&z)[2]) const {
        if constexpr (!INV) {
            // pass 1 (points j, j+4, j+8, j+12 of column t'): in-register stage, then lanes t' <-> t' + 4
            const Cx<T> s = z[0] + z[1], d = z[0] - z[1];
            const Cx<T> os = shfl_xor(s, 4, smask), od = shfl_xor(d, 4, smask);
            const T sg = lo4 ? T(1) : T(-1);
            z[0] = pm(sg, s, os);                                       // y0 = a0 + a2 | y2 = a0 - a2
            z[1] = rot(sg, lo4 ? d : od, lo4 ? od : d);                 // y1 = a1 - i a3 | y3 = a1 + i a3
        } else {
            // pass 1 of the inverse runs over d = bitrev2(t'): lanes t' <-> t' ^ 1, then t' <-> t' ^ 2
            const T sg0 = b0 ? T(-1) : T(1), sg1 = b1 ? T(-1) : T(1);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const Cx<T> o = shfl_xor(z[q], 1, smask);
                z[q] = pm(sg0, z[q], o);                                 // a0, a1 (d = 0, 2) | a2, a3 (d = 1, 3)
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const Cx<T> o = shfl_xor(z[q], 2, smask);
                const Cx<T> plain = pm(sg1, z[q], o);                    // B0 = a0 + a2 | B2 = a0 - a2
                const Cx<T> turned = rot(-sg1, b1 ? o : z[q], b1 ? z[q] : o);   // B1 = a1 + i a3 | B3 = a1 - i a3
                z[q] = b0 ? turned : plain;
            }
        }
    }
    template <bool INV>
    __device__ __forceinline__ void twiddle(Cx<T> (&z)[2]) const {
        const Cx<T> t0 = INV ? cmulc(z[0], wa) : cmul(z[0], wa);
        z[0] = lo4 ? z[0] : t0;                                          // c = 0 is never multiplied
        z[1] = INV ? cmulc(z[1], wb) : cmul(z[1], wb);
    }
    template <bool INV>
    __device__ __forceinline__ void second_pass(Cx<T> (&z)[2]) const {
        if constexpr (!INV) {
            // pass 2 over t' (lanes): t' <-> t' ^ 2, then t' <-> t' ^ 1
            const T sg1 = b1 ? T(-1) : T(1), sg0 = b0 ? T(-1) : T(1);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const Cx<T> o = shfl_xor(z[q], 2, smask);
                z[q] = pm(sg1, z[q], o);                                 // a0, a2 | a1, a3
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const Cx<T> o = shfl_xor(z[q], 1, smask);
                const Cx<T> plain = pm(sg0, z[q], o);                    // y0 = a0 + a2 | y2 = a0 - a2
                const Cx<T> turned = rot(sg0, b0 ? o : z[q], b0 ? z[q] : o);    // y1 = a1 - i a3 | y3 = a1 + i a3
                z[q] = b1 ? turned : plain;
            }
        } else {
            // pass 2 of the inverse over c = 2 (tl >> 2) + q: lanes t' <-> t' + 4, then the in-register stage
            const T sg = lo4 ? T(1) : T(-1);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const Cx<T> o = shfl_xor(z[q], 4, smask);
                z[q] = pm(sg, z[q], o);                                  // a0, a2 | a1, a3
            }
            // lo: x[p=0] = a0 + a2, x[p=2] = a0 - a2;  hi: x[p=1] = a1 + i a3, x[p=3] = a1 - i a3
            const Cx<T> u = lo4 ? z[1] : cx<T>(-z[1].im, z[1].re);
            const Cx<T> r0 = z[0];
            z[0] = r0 + u;
            z[1] = r0 - u;
        }
    }
    __device__ __forceinline__ void fwd(Cx<T> (&z)[2]) const {
        first_pass<false>(z);
        twiddle<false>(z);
        second_pass<false>(z);
    }
    __device__ __forceinline__ void fwd2(Cx<T> (&za)[2], Cx<T> (&zb)[2]) const {
        first_pass<false>(za);
        first_pass<false>(zb);
        twiddle<false>(za);
        twiddle<false>(zb);
        second_pass<false>(za);
        second_pass<false>(zb);
    }
    __device__ __forceinline__ void inv(Cx<T> (&z)[2]) const {
        first_pass<true>(z);
        twiddle<true>(z);
        second_pass<true>(z);
    }
    __device__ __forceinline__ Cx<T> mirrored(const Cx<T> (&z)[2], int q) const { return shfl(z[q], part[q], smask); }
};

// 32 points on 8 lanes x 4 registers as 4 x (2 x 4): a 4-point transform in registers, the twiddle W32^(lane r), ONE
// transpose through shared memory (row stride 9 complex words: conflict-free), a radix-2 stage between lanes l and l ^ 4,
// the twiddle W8^(m h) and a second 4-point transform in registers.  Natural order in and out: point j = 8 p + tl in
// register p, wavenumber k = tl + 8 q in register q.  Per environment 8 x 64 = 512 FP64 and 128 SHFL.32 instructions
// instead of 16 x 56 = 896 and 512 for the 16-lane radix-2 shuffle network: the KS step is nine of these transforms.
// Selected with the team-size tag -8 for H = 32 (N = 64).
template <typename T, int TWS>
struct WarpFFT<T, 32, -8, TWS> {
    static constexpr int H = 32, TS = 8, P = 4, ROW = 9;
    static constexpr int SMEM_CX = 4 * ROW;
    static constexpr bool NATURAL = true;    // k = t + 8 p
    int tl, base, k1, hbit;
    unsigned tmask, smask;
    T sg;
    Cx<T> w1[3];                        // W32^(tl r), r = 1..3
    Cx<T> w2[3];                        // W8^(m h),  m = 1..3 (1 on the lanes with h = 0)
    Cx<T>* sm;

    __device__ __forceinline__ static int kidx(int p, int t) { return t + 8 * p; }
    // table entry exp(-2 pi i m / 32) from tw[j] = exp(-2 pi i j / (32 TWS)), j < 32 (TWS = 2): W32^m = tw[2 m], m < 32
    __device__ __forceinline__ static Cx<T> w32(const Cx<T>* __restrict__ tw, int m) {
        const int j = m * TWS;                                   // j < 32 TWS
        const Cx<T> w = ldcx(tw + (j & 31));
        return (j & 32) ? cx<T>(-w.re, -w.im) : w;               // exp(-2 pi i (j + 32) / 64) = -exp(-2 pi i j / 64)
    }
    __device__ __forceinline__ void init(const Cx<T>* __restrict__ tw, Cx<T>* team_smem) {
        static_assert(TWS == 2, "the 32-point transposed transform reads the N = 64 twiddle table");
        const int lane = threadIdx.x & 31;
        tl = lane & 7;
        base = lane & ~7;
        tmask = 0xffu << base;
        smask = tmask;
        k1 = tl & 3;
        hbit = tl >> 2;
        sg = hbit ? T(-1) : T(1);
        sm = team_smem;
#pragma unroll
        for (int r = 1; r < 4; ++r) {
            w1[r - 1] = w32(tw, tl * r);                         // tl r <= 21
            w2[r - 1] = hbit ? w32(tw, 4 * r) : cx<T>(T(1), T(0));   // W8^r = W32^(4 r)
        }
    }
    using Q = WarpFFT<T, 16, 4, TWS>;                            // its dft4 is the natural-order 4-point transform
    __device__ __forceinline__ void fwd(Cx<T> (&z)[4]) const {
        Q::template dft4<false>(z);                              // over p -> k1 in registers
#pragma unroll
        for (int r = 1; r < 4; ++r) z[r] = cmul(z[r], w1[r - 1]);
        __syncwarp(smask);
#pragma unroll
        for (int r = 0; r < 4; ++r) stcx(sm + r * ROW + tl, z[r]);
        __syncwarp(smask);
#pragma unroll
        for (int m = 0; m < 4; ++m) z[m] = ldcx(sm + k1 * ROW + 4 * hbit + m);      // lane (k1, h): points n2 = 4 h + m
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const Cx<T> o = shfl_xor(z[m], 4, smask);
            z[m] = cx<T>(fma(sg, z[m].re, o.re), fma(sg, z[m].im, o.im));
        }
#pragma unroll
        for (int m = 1; m < 4; ++m) z[m] = cmul(z[m], w2[m - 1]);
        Q::template dft4<false>(z);                              // over m -> q: k = k1 + 4 h + 8 q
    }
    __device__ __forceinline__ void inv(Cx<T> (&z)[4]) const {
        Q::template dft4<true>(z);
#pragma unroll
        for (int m = 1; m < 4; ++m) z[m] = cmulc(z[m], w2[m - 1]);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const Cx<T> o = shfl_xor(z[m], 4, smask);
            z[m] = cx<T>(fma(sg, z[m].re, o.re), fma(sg, z[m].im, o.im));
        }
        __syncwarp(smask);
#pragma unroll
        for (int m = 0; m < 4; ++m) stcx(sm + k1 * ROW + 4 * hbit + m, z[m]);
        __syncwarp(smask);
#pragma unroll
        for (int r = 0; r < 4; ++r) z[r] = ldcx(sm + r * ROW + tl);
#pragma unroll
        for (int r = 1; r < 4; ++r) z[r] = cmulc(z[r], w1[r - 1]);
        Q::template dft4<true>(z);
    }
    __device__ __forceinline__ void fwd2(Cx<T> (&za)[4], Cx<T> (&zb)[4]) const { fwd(za); fwd(zb); }
    // value held for wavenumber -k (mod 32): k = tl + 8 q -> lane (8 - tl) & 7, register 3 - q (tl > 0) or (4 - q) & 3
    __device__ __forceinline__ Cx<T> mirrored(const Cx<T> (&z)[4], int q) const {
        const Cx<T> far = shfl(z[3 - q], base + ((8 - tl) & 7), smask);
        return tl == 0 ? z[(4 - q) & 3] : far;
    }
};

template <typename T, int N, int TS_ = (N / 2 < 32 ? N / 2 : 32)>      // TS_ < 0: tag of an alternative transform with |TS_| lanes
struct RealFFT {
    static constexpr int H = N / 2;
    using C = WarpFFT<T, H, TS_, 2>;
    static constexpr int TS = C::TS, P = C::P;
    static_assert(N >= 8 && N <= 256, "warp-resident real FFT handles N = 8..256");

    C c;
    Cx<T> wk[P];     // exp(-2 pi i k / N) for the wavenumber k of register p
    bool dc;         // this lane's register 0 holds k = 0 (and owns the Nyquist value)

    static constexpr int SMEM_CX = C::SMEM_CX;       // complex words of shared memory this transform needs per team
    __device__ __forceinline__ void init(const Cx<T>* __restrict__ tw /* [N/2]: exp(-2 pi i j / N) */, Cx<T>* team_smem = nullptr) {
        c.init(tw, team_smem);
#pragma unroll
        for (int p = 0; p < P; ++p) wk[p] = ldcx(tw + C::kidx(p, c.tl));
        dc = c.tl == 0;
    }
    // wavenumber index 0..H-1 of register p
    __device__ __forceinline__ int k(int p) const { return C::kidx(p, c.tl); }
    // Promise that all 32 lanes of the warp execute every shuffle / __syncwarp of this object together (warp-uniform control
    // flow): the collectives then name the compile-time full mask, and the compiler drops the run-time mask validation
    // (MATCH.ANY + a divergence branch ahead of each group of collectives -- 4 % of the hot kernel's time).
    __device__ __forceinline__ void whole_warp() { c.smask = 0xffffffffu; }

    // z[p] = (x_{2j}, x_{2j+1}), j = p*TS + tl   ->   X[p] = scale * fft(x)[k(p)],
    // nyq = scale * fft(x)[N/2] (real; meaningful on the dc lane only).  z is clobbered.
    // ws[p] must hold (scale/2) * wk[p] (pre-scaled by the caller, once per kernel).
    __device__ __forceinline__ void fwd(Cx<T> (&z)[P], Cx<T> (&X)[P], T& nyq, T scale, const Cx<T> (&ws)[P]) const {
        c.fwd(z);
        const T hs = T(0.5) * scale;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const Cx<T> zp = c.mirrored(z, p);                                   // Z[H-k]
            const Cx<T> E = cx<T>(z[p].re + zp.re, z[p].im - zp.im);            // 2 * fft(x_even)[k]
            const Cx<T> O = cx<T>(z[p].im + zp.im, zp.re - z[p].re);            // 2 * fft(x_odd)[k]
            X[p] = cx<T>(fma(hs, E.re, fma(ws[p].re, O.re, -ws[p].im * O.im)),
                         fma(hs, E.im, fma(ws[p].re, O.im, ws[p].im * O.re)));
            if (p == 0) nyq = (E.re - O.re) * hs;                                // k = 0: E, O real
        }
    }
    // two real fields at once (see WarpFFT::fwd2)
    __device__ __forceinline__ void fwd2(Cx<T> (&za)[P], Cx<T> (&Xa)[P], T& nyqa, T sa, const Cx<T> (&wsa)[P],
                                         Cx<T> (&zb)[P], Cx<T> (&Xb)[P], T& nyqb, T sb, const Cx<T> (&wsb)[P]) const {
        c.fwd2(za, zb);
        const T ha = T(0.5) * sa, hb = T(0.5) * sb;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const Cx<T> zpa = c.mirrored(za, p), zpb = c.mirrored(zb, p);
            const Cx<T> Ea = cx<T>(za[p].re + zpa.re, za[p].im - zpa.im), Oa = cx<T>(za[p].im + zpa.im, zpa.re - za[p].re);
            const Cx<T> Eb = cx<T>(zb[p].re + zpb.re, zb[p].im - zpb.im), Ob = cx<T>(zb[p].im + zpb.im, zpb.re - zb[p].re);
            Xa[p] = cx<T>(fma(ha, Ea.re, fma(wsa[p].re, Oa.re, -wsa[p].im * Oa.im)),
                          fma(ha, Ea.im, fma(wsa[p].re, Oa.im, wsa[p].im * Oa.re)));
            Xb[p] = cx<T>(fma(hb, Eb.re, fma(wsb[p].re, Ob.re, -wsb[p].im * Ob.im)),
                          fma(hb, Eb.im, fma(wsb[p].re, Ob.im, wsb[p].im * Ob.re)));
            if (p == 0) { nyqa = (Ea.re - Oa.re) * ha; nyqb = (Eb.re - Ob.re) * hb; }
        }
    }
    // UNSCALED split step: X2[p] = E + wk O = 2 fft(x)[k(p)], nyq2 = 2 fft(x)[N/2] -- two FMAs per component instead of
    // a multiply and two FMAs; the caller folds the 1/2 into a constant it multiplies with anyway.
    __device__ __forceinline__ void split_raw(const Cx<T> (&z)[P], Cx<T> (&X)[P], T& nyq) const {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const Cx<T> zp = c.mirrored(z, p);
            const Cx<T> E = cx<T>(z[p].re + zp.re, z[p].im - zp.im);
            const Cx<T> O = cx<T>(z[p].im + zp.im, zp.re - z[p].re);
            X[p] = cx<T>(fma(wk[p].re, O.re, fma(-wk[p].im, O.im, E.re)), fma(wk[p].re, O.im, fma(wk[p].im, O.re, E.im)));
            if (p == 0) nyq = E.re - O.re;
        }
    }
    __device__ __forceinline__ void fwd_raw(Cx<T> (&z)[P], Cx<T> (&X)[P], T& nyq) const {
        c.fwd(z);
        split_raw(z, X, nyq);
    }
    __device__ __forceinline__ void fwd2_raw(Cx<T> (&za)[P], Cx<T> (&Xa)[P], T& nyqa, Cx<T> (&zb)[P], Cx<T> (&Xb)[P], T& nyqb) const {
        c.fwd2(za, zb);
        split_raw(za, Xa, nyqa);
        split_raw(zb, Xb, nyqb);
    }
    __device__ __forceinline__ void scaled_twiddles(T scale, Cx<T> (&ws)[P]) const {
#pragma unroll
        for (int p = 0; p < P; ++p) ws[p] = cx<T>(wk[p].re * (T(0.5) * scale), wk[p].im * (T(0.5) * scale));
    }
    // X[p] = spectrum at k(p), nyq = real Nyquist value (dc lane)  ->  z[p] = N * ifft(X) at points
    // (2j, 2j+1), UNNORMALISED.  Only the real parts of X_0 and X_{N/2} enter (np.real(ifft(.))).
    __device__ __forceinline__ void inv(const Cx<T> (&X)[P], T nyq, Cx<T> (&z)[P]) const {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const Cx<T> xp = c.mirrored(X, p);                                   // X[H-k]
            const Cx<T> E = cx<T>(X[p].re + xp.re, X[p].im - xp.im);
            const Cx<T> D = cx<T>(X[p].re - xp.re, X[p].im + xp.im);
            // z = E + i * (D * conj(wk))
            z[p] = cx<T>(fma(-D.im, wk[p].re, fma(D.re, wk[p].im, E.re)),
                         fma(D.re, wk[p].re, fma(D.im, wk[p].im, E.im)));
            if (p == 0 && dc) z[p] = cx<T>(X[p].re + nyq, X[p].re - nyq);
        }
        c.inv(z);
    }
};

}  // namespace mpde
