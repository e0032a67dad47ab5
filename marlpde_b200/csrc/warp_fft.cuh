// Warp-level radix-2 FFT on register-resident data, for N <= 128 (sm_100a).
//
// A "team" of TS = min(N, 32) lanes owns one length-N complex sequence; lane tl holds
// the P = N/TS points n = p*TS + tl.  Forward = decimation-in-frequency (natural order
// in, BIT-REVERSED order out): the first log2(P) stages pair registers of one lane, the
// last log2(TS) stages pair lanes through __shfl_xor.  Inverse = the exact transpose
// (decimation-in-time, bit-reversed in, natural out, unnormalised).  Spectral
// arithmetic is done in the bit-reversed layout, so no permutation pass is ever needed.
//
// One complex transform carries TWO real fields (environment A in the real part,
// environment B in the imaginary part): untangle() splits the two Hermitian spectra,
// tangle() packs two Hermitian spectra for a single inverse transform.
#pragma once
#include "common.cuh"

namespace mpde {

__host__ __device__ constexpr int brev_bits(int x, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((x >> i) & 1) << (bits - 1 - i);
    return r;
}

template <typename T, int N>
struct WarpFFT {
    static constexpr int TS = N < 32 ? N : 32;
    static constexpr int P = N / TS;
    static constexpr int LOGN = ilog2(N);
    static constexpr int LOGTS = ilog2(TS);
    static constexpr int LOGP = ilog2(P);
    static_assert((1 << LOGN) == N && N >= 4 && N <= 128, "warp FFT handles N = 4..128, power of two");

    int tl;                         // lane within the team
    int base;                       // first lane of the team within the warp
    Cx<T> wx[LOGTS];                // cross-lane twiddles, (1,0) on the lower lane of a pair
    Cx<T> wl[P > 1 ? P - 1 : 1];    // in-register twiddles
    int part[P];                    // lane holding wavenumber -k for register pp(p)

    // register index that holds -k for the k held in register p (depends on p only)
    __host__ __device__ static constexpr int pp(int p) {
        return brev_bits((P - brev_bits(p, LOGP)) % P, LOGP);
    }
    // wavenumber index (0..N-1, FFT order) held at register p of team-lane t after fwd()
    __device__ __forceinline__ static int kidx(int p, int t) {
        return (int)(__brev((unsigned)(p * TS + t)) >> (32 - LOGN));
    }

    __device__ __forceinline__ void init(const Cx<T>* __restrict__ tw /* [N/2]: exp(-2 pi i j/N) */) {
        const int lane = threadIdx.x & 31;
        tl = lane & (TS - 1);
        base = lane & ~(TS - 1);
#pragma unroll
        for (int s = 0; s < LOGTS; ++s) {
            const int h = TS >> (s + 1);
            wx[s] = (tl & h) ? ldcx(tw + (tl & (h - 1)) * (N / (2 * h))) : cx<T>(T(1), T(0));
        }
        if constexpr (P > 1) {
#pragma unroll
            for (int hp = P / 2; hp >= 1; hp >>= 1)
#pragma unroll
                for (int q = 0; q < hp; ++q)
                    wl[P - 2 * hp + q] = ldcx(tw + (q * TS + tl) * (P / (2 * hp)));
        }
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int k = kidx(p, tl);
            const int kneg = (N - k) & (N - 1);
            const int npos = (int)(__brev((unsigned)kneg) >> (32 - LOGN));   // position holding -k
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
                const Cx<T> o = shfl_xor(z[p], h);
                const Cx<T> t = cx<T>(fma(sg, z[p].re, o.re), fma(sg, z[p].im, o.im));
                z[p] = (h == 1) ? t : cmul(t, wx[s]);          // last stage: twiddle is 1
            }
        }
    }

    // bit-reversed spectrum -> natural order, unnormalised inverse DFT (N * ifft)
    __device__ __forceinline__ void inv(Cx<T> (&z)[P]) const {
#pragma unroll
        for (int s = LOGTS - 1; s >= 0; --s) {
            const int h = TS >> (s + 1);
            const bool up = (tl & h) != 0;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const Cx<T> m = (h == 1) ? z[p] : cmulc(z[p], wx[s]);
                const Cx<T> o = shfl_xor(m, h);
                z[p] = up ? (o - m) : (m + o);
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

    // Z = fwd(xA + i xB)  ->  XA = fft(xA), XB = fft(xB) at this lane's wavenumbers,
    // both multiplied by `scale`
    __device__ __forceinline__ void untangle(const Cx<T> (&z)[P], Cx<T> (&xa)[P], Cx<T> (&xb)[P], T scale) const {
        const T hs = T(0.5) * scale;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const Cx<T> zp = shfl(z[pp(p)], part[p]);                  // Z[-k]
            xa[p] = cx<T>((z[p].re + zp.re) * hs, (z[p].im - zp.im) * hs);
            xb[p] = cx<T>((z[p].im + zp.im) * hs, (zp.re - z[p].re) * hs);
        }
    }

    // Z = SA + i SB for two Hermitian spectra.  `selfc[p]` marks k = 0 and k = N/2, where
    // only the real part of a spectrum reaches Re(ifft) (Burger.py:491 takes np.real).
    __device__ __forceinline__ static Cx<T> tangle(Cx<T> sa, Cx<T> sb, bool selfc) {
        return selfc ? cx<T>(sa.re, sb.re) : cx<T>(sa.re - sb.im, sa.im + sb.re);
    }
};

}  // namespace mpde
