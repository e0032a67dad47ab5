// A-priori sub-grid-scale term of a DNS history, Burger.compute_Sgs(nURG) (/root/reference/python/_model/Burger.py:677-736)
// and KS.compute_Sgs(nURG) (KS.py:385-409): testing-mode diagnostic of burger_testing_environment.py.  One CTA per history
// row: sharp spectral filter |k| > nURG // 2 of u, u^2 (and of the next row for d/dt) through the shared-memory real FFT,
// first-order upwind / centred differences of the filtered fields, plus the "Alt2" variant on the coarse nURG-point grid
// (truncated spectrum -> nURG-point inverse transform, evaluated as a direct sum: nURG <= 64).
#include "dispatch.h"
#include "cta_fft.cuh"

namespace mpde {

template <typename T, int N, int NT>
__global__ void __launch_bounds__(NT) sgs_rows_kernel(const SpectralParams<T> prm, const T* __restrict__ uu, int64_t rows, int nURG,
                                                       int ks, T* __restrict__ sgs, T* __restrict__ alt, T* __restrict__ alt2) {
    using F = CtaFFT<T, N, NT>;
    constexpr int H = N / 2, NH = H + 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cx<T>* buf = reinterpret_cast<Cx<T>*>(smem_raw);
    Cx<T>* X = buf + H;
    T* uh = reinterpret_cast<T*>(X + NH);
    T* u2h = uh + N;
    T* uhpt = u2h + N;
    Cx<T>* c0 = reinterpret_cast<Cx<T>*>(uhpt + N);      // truncated spectrum of the row, nURG entries
    Cx<T>* c1 = c0 + 64;                                  // ... of the neighbouring row
    T* a0 = reinterpret_cast<T*>(c1 + 64);               // coarse fields
    T* a1 = a0 + 64;
    const int t = threadIdx.x;
    const int64_t e = blockIdx.x / rows, idx = blockIdx.x % rows;
    const int64_t dtidx = idx < rows - 1 ? idx + 1 : idx - 1;                     // Burger.py:686
    const T invN = T(1) / T(N), dx = prm.dx, dt = prm.dt, nu = prm.nu[e];
    const T cutoff = T(nURG / 2);
    const int n_pos = (nURG + 1) / 2;
    const T* row = uu + (e * rows + idx) * N;
    const T* rowp = uu + (e * rows + dtidx) * N;

    // filtered real field of `src` (optionally squared) -> dst; optionally keep the truncated spectrum in cc
    auto filtered = [&](const T* src, bool square, T* dst, Cx<T>* cc) {
        for (int j = t; j < H; j += NT) {
            const Cx<T> z = ldcx(reinterpret_cast<const Cx<T>*>(src) + j);
            buf[j] = square ? cx<T>(z.re * z.re, z.im * z.im) : z;
        }
        __syncthreads();
        F::rfwd(buf, X, T(1), prm.tw);
        for (int k = t; k < NH; k += NT)
            if (fabs(prm.kwave[k]) > cutoff) X[k] = cx<T>(0, 0);                    // hidx = |k| > nURG // 2 (dimensional k)
        __syncthreads();
        if (cc) {
            // concat(v[:(nURG+1)//2], v[-(nURG-1)//2:]) of the FILTERED spectrum (vh aliases v, Burger.py:691-693,701-707)
            for (int m = t; m < nURG; m += NT) {
                Cx<T> val;
                if (m < n_pos) val = X[m];
                else { const int kk = nURG - m; val = cx<T>(X[kk].re, -X[kk].im); }   // v[N - kk] = conj(v[kk])
                cc[m] = val;
            }
        }
        F::rinv(X, buf, prm.tw);
        for (int j = t; j < H; j += NT) {
            dst[2 * j] = buf[j].re * invN;
            dst[2 * j + 1] = buf[j].im * invN;
        }
        __syncthreads();
    };
    filtered(row, false, uh, ks ? nullptr : c0);
    filtered(row, true, u2h, nullptr);
    if (!ks) filtered(rowp, false, uhpt, c1);

    const T sgn = idx == rows - 1 ? T(-1) : T(1);                                    // Burger.py:712-714
    T* out_s = sgs + (e * rows + idx) * N;
    T* out_a = alt ? alt + (e * rows + idx) * N : nullptr;
    for (int n = t; n < N; n += NT) {
        const int nm = (n + N - 1) & (N - 1), np_ = (n + 1) & (N - 1);
        const T duhdx = (uh[n] - uh[nm]) / dx;
        const T du2hdx = (u2h[n] - u2h[nm]) / dx;
        out_s[n] = -uh[n] * duhdx + T(0.5) * du2hdx;                                 // Burger.py:734, KS.py:409
        if (!ks && out_a) {
            const T d2 = (uh[np_] - T(2) * uh[n] + uh[nm]) / (dx * dx);
            const T duhdt = sgn * ((uhpt[n] - uh[n]) / dt);
            out_a[n] = duhdt + uh[n] * duhdx - nu * d2;                               // Burger.py:735
        }
    }
    if (!ks && alt2) {
        const T r = T(nURG) / T(N);
        // nURG-point inverse transforms of the truncated spectra as direct sums
        for (int j = t; j < 2 * nURG; j += NT) {
            const int jj = j % nURG;
            const Cx<T>* cc = j < nURG ? c0 : c1;
            T acc = T(0);
            for (int m = 0; m < nURG; ++m) {
                T s, c;
                sincospi(T(2) * T((jj * m) % nURG) / T(nURG), &s, &c);
                acc += cc[m].re * c - cc[m].im * s;
            }
            (j < nURG ? a0 : a1)[jj] = acc / T(nURG) * r;
        }
        __syncthreads();
        T* out2 = alt2 + (e * rows + idx) * nURG;
        for (int j = t; j < nURG; j += NT) {
            const int jm = (j + nURG - 1) % nURG, jp = (j + 1) % nURG;
            const T duhdt = sgn * ((a1[j] - a0[j]) / dt);
            const T duhdx = (a0[j] - a0[jm]) / dx * r;
            const T d2 = (a0[jp] - T(2) * a0[j] + a0[jm]) / (dx * dx) * r * r;
            out2[j] = duhdt + a0[j] * duhdx - nu * d2;                                // Burger.py:736
        }
    }
}

template <typename T, int N, int NT>
static int launch_sgs_n(const SpectralParams<T>& p, const void* uu, int64_t rows, int nURG, int ks, void* sgs, void* alt, void* alt2,
                        cudaStream_t st) {
    constexpr int H = N / 2, NH = H + 1;
    const size_t smem = sizeof(Cx<T>) * (H + NH + 128) + sizeof(T) * (3 * N + 128) + 16;
    if (smem > 48 * 1024) {
        if (cudaFuncSetAttribute(sgs_rows_kernel<T, N, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
    }
    sgs_rows_kernel<T, N, NT><<<(unsigned)(p.B * rows), NT, smem, st>>>(p, static_cast<const T*>(uu), rows, nURG, ks, static_cast<T*>(sgs),
                                                                         static_cast<T*>(alt), static_cast<T*>(alt2));
    return 1;
}

template <typename T>
int launch_sgs(const SpectralParams<T>& p, const void* uu, int64_t rows, int nURG, int ks, void* sgs, void* alt, void* alt2, cudaStream_t st) {
    switch (p.N) {
        case 256: return launch_sgs_n<T, 256, 64>(p, uu, rows, nURG, ks, sgs, alt, alt2, st);
        case 512: return launch_sgs_n<T, 512, 64>(p, uu, rows, nURG, ks, sgs, alt, alt2, st);
        case 1024: return launch_sgs_n<T, 1024, 128>(p, uu, rows, nURG, ks, sgs, alt, alt2, st);
        case 2048: return launch_sgs_n<T, 2048, 256>(p, uu, rows, nURG, ks, sgs, alt, alt2, st);
        default: return -1;
    }
}
template int launch_sgs<double>(const SpectralParams<double>&, const void*, int64_t, int, int, void*, void*, void*, cudaStream_t);
template int launch_sgs<float>(const SpectralParams<float>&, const void*, int64_t, int, int, void*, void*, void*, cudaStream_t);

}  // namespace mpde
