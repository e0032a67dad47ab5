// Device-side initial conditions for a batch of environments (episode reset without a host loop).
//   handoff_kernel     : DNS -> LES spectral hand-off with a per-environment phase shift
//                        (/root/reference/python/_model/burger_environment.py:109-112, ks_environment.py:52-54)
//   turbulence_kernel  : 'turbulence' initial field (Burger.py:227-260): k^-5/3 spectrum, phases from the 13-bit LCG
//                        seeded with 123456789 + seed, rescaled until rms(u - 1) is in [0.65, 0.75]
// Both write a [B,N] staging buffer that the ordinary reset kernels (IC(v0=...) / IC(u0=...)) consume, so everything
// after the field itself (v0, u0, Fn_old, spectrum sums, counters) is the code path the parity tests already pin.
// Arithmetic is fp64 in the reference's order of operations; sin / cos are CUDA's (<= 2 ulp from libm).
#include "dispatch.h"

namespace mpde {

template <typename T>
__global__ void handoff_kernel(const Cx<double>* __restrict__ vsrc, int nsrc_points, const double* __restrict__ ksrc,
                               const int* __restrict__ src_map, const double* __restrict__ offset,
                               const uint8_t* __restrict__ mask, Cx<T>* __restrict__ out, int64_t B, int N) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * N) return;
    const int64_t e = idx / N;
    const int j = (int)(idx - e * N);
    if (mask && !mask[e]) return;
    // np.concatenate((v[:(g+1)//2], v[-(g-1)//2:])): modes 0..g/2-1 and -g/2..-1 of the source
    const int jj = j < (N + 1) / 2 ? j : nsrc_points - N + j;
    const int64_t s = src_map ? src_map[e] : 0;
    const Cx<double> v = vsrc[s * nsrc_points + jj];
    const double off = offset ? offset[e] : 0.0;
    // np.exp(1j * 2 * np.pi * offset * k): argument ((2 pi) * offset) * k, literal transcription
    const double th = ((2.0 * 3.141592653589793) * off) * ksrc[jj];
    double sn, cs;
    sincos(th, &sn, &cs);
    const double re = __dsub_rn(__dmul_rn(v.re, cs), __dmul_rn(v.im, sn));       // numpy complex product: no fma
    const double im = __dadd_rn(__dmul_rn(v.re, sn), __dmul_rn(v.im, cs));
    out[idx] = cx<T>((T)(re * (double)N / (double)nsrc_points), (T)(im * (double)N / (double)nsrc_points));
}

// one CTA per environment, blockDim.x threads stride over the N grid points
template <typename T>
__global__ void __launch_bounds__(256) turbulence_kernel(const long long* __restrict__ seed, const double* __restrict__ offset,
                                                         const double* __restrict__ x, const double* __restrict__ amp,
                                                         const uint8_t* __restrict__ mask, T* __restrict__ out, int N, double L) {
    extern __shared__ double sm[];          // [N] phases, [32] reduction
    double* phase = sm;
    double* red = sm + N;
    const int64_t e = blockIdx.x;
    if (mask && !mask[e]) return;
    if (threadIdx.x == 0) {
        long long rng = 123456789LL + seed[e];
        for (int k = 1; k < N; ++k) {
            rng = (1103515245LL * rng + 12345LL) % 8192LL;                 // Burger.py:238 (rng >= 0 throughout)
            phase[k] = (double)rng / 8192.0 * 2.0 * 3.141592653589793;       // Burger.py:239
        }
    }
    __syncthreads();
    const double off = offset ? offset[e] : 0.0;
    constexpr int MAXP = 8;                 // N <= 2048 with 256 threads
    double u[MAXP];
    int np = 0;
    for (int j = threadIdx.x; j < N; j += blockDim.x, ++np) {
        double acc = 1.0;
        const double xo = x[j] + off;
        for (int k = 1; k < N; ++k)         // u0 += sqrt(2 Ek) * sin(k * 2 * pi * (x + off) / L + phase)
            acc = __dadd_rn(acc, __dmul_rn(amp[k], sin(__dadd_rn(((double)(2 * k) * 3.141592653589793) * xo / L, phase[k]))));
        u[np] = acc;
    }
    auto rms = [&]() {
        double s = 0.0;
        for (int q = 0; q < np; ++q) s += (u[q] - 1.0) * (u[q] - 1.0);
        for (int h = 16; h >= 1; h >>= 1) s += __shfl_xor_sync(0xffffffffu, s, h);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x + 31) / 32; ++w) t += red[w];
        return sqrt(t / (double)N);
    };
    double crit = rms();
    for (int it = 0; (crit < 0.65 || crit > 0.75) && it <= 100; ++it) {       // Burger.py:247-257
        const double fac = 0.7 / crit;
        for (int q = 0; q < np; ++q) u[q] *= fac;
        crit = rms();
    }
    np = 0;
    for (int j = threadIdx.x; j < N; j += blockDim.x, ++np) out[e * N + j] = (T)u[np];
}

template <typename T>
int launch_handoff(const void* vsrc, int nsrc_points, const double* ksrc, const int* src_map, const double* offset,
                   const uint8_t* mask, void* out, int64_t B, int N, cudaStream_t st) {
    const int64_t n = B * N;
    handoff_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(static_cast<const Cx<double>*>(vsrc), nsrc_points, ksrc, src_map,
                                                                  offset, mask, static_cast<Cx<T>*>(out), B, N);
    return 1;
}
template <typename T>
int launch_turbulence(const long long* seed, const double* offset, const double* x, const double* amp, const uint8_t* mask,
                      void* out, int64_t B, int N, double L, cudaStream_t st) {
    if (N > 2048) return -1;
    const int block = N >= 256 ? 256 : (N < 32 ? 32 : N);
    turbulence_kernel<T><<<(unsigned)B, block, (size_t)(N + 32) * sizeof(double), st>>>(seed, offset, x, amp, mask, static_cast<T*>(out),
                                                                                        N, L);
    return 1;
}

template int launch_handoff<double>(const void*, int, const double*, const int*, const double*, const uint8_t*, void*, int64_t, int, cudaStream_t);
template int launch_handoff<float>(const void*, int, const double*, const int*, const double*, const uint8_t*, void*, int64_t, int, cudaStream_t);
template int launch_turbulence<double>(const long long*, const double*, const double*, const double*, const uint8_t*, void*, int64_t, int, double, cudaStream_t);
template int launch_turbulence<float>(const long long*, const double*, const double*, const double*, const uint8_t*, void*, int64_t, int, double, cudaStream_t);

}  // namespace mpde
