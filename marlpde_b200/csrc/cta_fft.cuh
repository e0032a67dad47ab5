// CTA-level real FFT in shared memory for N = 256..4096 (sm_100a).
//
// One CTA owns one environment.  A real length-N field lives in shared memory as H = N/2
// complex numbers z_j = (x_{2j}, x_{2j+1}); the transform is an in-place radix-2 FFT of those H
// points (each thread keeps PT = H/NT points in registers per pass and does log2(PT) butterfly
// stages between two __syncthreads) followed by the split step that yields the half spectrum
// X_0..X_H in NATURAL order in a second shared array.  The inverse runs the same passes
// backwards.  Twiddles come from the global table exp(-2 pi i j / N) (L1-resident).
#pragma once
#include "common.cuh"

namespace mpde {

template <typename T, int N, int NT>
struct CtaFFT {
    static constexpr int H = N / 2;
    static constexpr int PT = H / NT;                 // points per thread per pass
    static constexpr int LOGH = ilog2(H), R = ilog2(PT);
    static_assert(PT >= 2 && (1 << R) == PT && (1 << LOGH) == H, "need a power-of-two number (>= 2) of points per thread");

    // index of register p of thread t in the pass that owns bits [b0, b0 + r)
    __device__ __forceinline__ static int index(int t, int p, int b0, int r) {
        return ((t >> b0) << (b0 + r)) | (p << b0) | (t & ((1 << b0) - 1));
    }

    // forward: buf holds the sequence in natural order -> bit-reversed spectrum in place
    __device__ static void fwd_inplace(Cx<T>* buf, const Cx<T>* __restrict__ tw) {
        const int t = threadIdx.x;
        for (int hi = LOGH; hi > 0;) {
            const int r = hi >= R ? R : hi;
            const int b0 = hi - r;
            // a short last pass (r < R) keeps PT >> r independent groups per thread
            const int groups = PT >> r;
            for (int gidx = 0; gidx < groups; ++gidx) {
                const int tt = t * groups + gidx;
                Cx<T> x[PT];
#pragma unroll
                for (int p = 0; p < PT; ++p)
                    if (p < (1 << r)) x[p] = buf[index(tt, p, b0, r)];
#pragma unroll
                for (int q = R - 1; q >= 0; --q) {
                    if (q >= r) continue;
                    const int half = 1 << (b0 + q);
#pragma unroll
                    for (int p = 0; p < PT; ++p) {
                        if (p >= (1 << r) || (p & (1 << q))) continue;
                        const int n = index(tt, p, b0, r);
                        const Cx<T> w = ldcx(tw + (size_t)(n & (half - 1)) * (H / (2 * half)) * 2);
                        const Cx<T> a = x[p], b = x[p | (1 << q)];
                        x[p] = a + b;
                        x[p | (1 << q)] = cmul(a - b, w);
                    }
                }
#pragma unroll
                for (int p = 0; p < PT; ++p)
                    if (p < (1 << r)) buf[index(tt, p, b0, r)] = x[p];
            }
            __syncthreads();
            hi = b0;
        }
    }

    // inverse: bit-reversed spectrum in buf -> natural-order sequence in place (unnormalised)
    __device__ static void inv_inplace(Cx<T>* buf, const Cx<T>* __restrict__ tw) {
        const int t = threadIdx.x;
        for (int b0 = 0; b0 < LOGH;) {
            const int r = (LOGH - b0) >= R ? R : (LOGH - b0);
            const int groups = PT >> r;
            for (int gidx = 0; gidx < groups; ++gidx) {
                const int tt = t * groups + gidx;
                Cx<T> x[PT];
#pragma unroll
                for (int p = 0; p < PT; ++p)
                    if (p < (1 << r)) x[p] = buf[index(tt, p, b0, r)];
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    if (q >= r) continue;
                    const int half = 1 << (b0 + q);
#pragma unroll
                    for (int p = 0; p < PT; ++p) {
                        if (p >= (1 << r) || (p & (1 << q))) continue;
                        const int n = index(tt, p, b0, r);
                        const Cx<T> w = ldcx(tw + (size_t)(n & (half - 1)) * (H / (2 * half)) * 2);
                        const Cx<T> a = x[p], b = cmulc(x[p | (1 << q)], w);
                        x[p] = a + b;
                        x[p | (1 << q)] = a - b;
                    }
                }
#pragma unroll
                for (int p = 0; p < PT; ++p)
                    if (p < (1 << r)) buf[index(tt, p, b0, r)] = x[p];
            }
            __syncthreads();
            b0 += r;
        }
    }

    __device__ __forceinline__ static int brev(int k) { return (int)(__brev((unsigned)k) >> (32 - LOGH)); }

    // real forward transform: buf (H complex = N reals, natural order, clobbered) -> X[0..H] natural
    // order, multiplied by `scale`.  X[H] is stored as (value, 0).  Ends with a __syncthreads().
    __device__ static void rfwd(Cx<T>* buf, Cx<T>* X, T scale, const Cx<T>* __restrict__ tw) {
        fwd_inplace(buf, tw);
        const T hs = T(0.5) * scale;
        for (int k = threadIdx.x; k <= H / 2; k += NT) {
            const int km = (H - k) & (H - 1);
            const Cx<T> zk = buf[brev(k)], zm = buf[brev(km)];
            // k
            {
                const Cx<T> E = cx<T>(zk.re + zm.re, zk.im - zm.im), O = cx<T>(zk.im + zm.im, zm.re - zk.re);
                const Cx<T> w = ldcx(tw + k);
                const Cx<T> wO = cmul(w, O);
                X[k] = cx<T>((E.re + wO.re) * hs, (E.im + wO.im) * hs);
                if (k == 0) X[H] = cx<T>((E.re - O.re) * hs, T(0));
            }
            if (km != k && k != 0) {
                const Cx<T> E = cx<T>(zm.re + zk.re, zm.im - zk.im), O = cx<T>(zm.im + zk.im, zk.re - zm.re);
                const Cx<T> w = ldcx(tw + km);
                const Cx<T> wO = cmul(w, O);
                X[km] = cx<T>((E.re + wO.re) * hs, (E.im + wO.im) * hs);
            }
        }
        __syncthreads();
    }

    // real inverse transform: X[0..H] natural order (only Re X[0], Re X[H] are used) -> buf = N * ifft
    // as H complex pairs in natural order.  Ends with a __syncthreads().
    __device__ static void rinv(const Cx<T>* X, Cx<T>* buf, const Cx<T>* __restrict__ tw) {
        for (int k = threadIdx.x; k <= H / 2; k += NT) {
            const int km = (H - k) & (H - 1);
            if (k == 0) {
                buf[0] = cx<T>(X[0].re + X[H].re, X[0].re - X[H].re);
                continue;
            }
            const Cx<T> xk = X[k], xm = X[km];
            {
                const Cx<T> E = cx<T>(xk.re + xm.re, xk.im - xm.im), D = cx<T>(xk.re - xm.re, xk.im + xm.im);
                const Cx<T> O = cmulc(D, ldcx(tw + k));
                buf[brev(k)] = cx<T>(E.re - O.im, E.im + O.re);
            }
            if (km != k) {
                const Cx<T> E = cx<T>(xm.re + xk.re, xm.im - xk.im), D = cx<T>(xm.re - xk.re, xm.im + xk.im);
                const Cx<T> O = cmulc(D, ldcx(tw + km));
                buf[brev(km)] = cx<T>(E.re - O.im, E.im + O.re);
            }
        }
        __syncthreads();
        inv_inplace(buf, tw);
    }
};

}  // namespace mpde
