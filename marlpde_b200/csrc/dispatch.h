// Kernel launchers (one translation unit per solver family keeps nvcc times short).
// Each returns the number of kernels it launched, or -1 when N is not supported.
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>
#include "params.h"
#include "pack.cuh"



namespace mpde {

template <typename T> int launch_burgers(const SpectralParams<T>& p, cudaStream_t st);
template <typename T> int launch_burgers_32_16(const SpectralParams<T>& p, cudaStream_t st);
template <typename T> int launch_burgers_32_8(const SpectralParams<T>& p, cudaStream_t st);
template <typename T> int launch_burgers_32_8x(const SpectralParams<T>& p, cudaStream_t st);
template <typename T> int launch_burgers_32_4(const SpectralParams<T>& p, cudaStream_t st);
template <typename T> int launch_burgers_64_32(const SpectralParams<T>& p, cudaStream_t st);
template <typename T> int launch_burgers_64_16(const SpectralParams<T>& p, cudaStream_t st);
template <typename T> int launch_burgers_64_8(const SpectralParams<T>& p, cudaStream_t st);
template <typename T> int launch_burgers_cta(const SpectralParams<T>& p, cudaStream_t st);
template <typename T>
int launch_spectral_aux_cta(const SpectralParams<T>& p, int equation, int mode, const void* src, const uint8_t* mask,
                            void* dst, cudaStream_t st);
template <typename T> int launch_ks(const SpectralParams<T>& p, cudaStream_t st);
template <typename T> int launch_rcp_table(const double* in, Cx<T>* out, int64_t n, cudaStream_t st);
template <typename T>
int launch_handoff(const void* vsrc, int nsrc_points, const double* ksrc, const int* src_map, const double* offset,
                   const uint8_t* mask, void* out, int64_t B, int N, cudaStream_t st);
template <typename T>
int launch_turbulence(const long long* seed, const double* offset, const double* x, const double* amp, const uint8_t* mask,
                      void* out, int64_t B, int N, double L, cudaStream_t st);
template <typename T> int launch_ks_cta(const SpectralParams<T>& p, cudaStream_t st);
template <typename T>
int launch_sgs(const SpectralParams<T>& p, const void* uu, int64_t rows, int nURG, int ks, void* sgs, void* alt, void* alt2, cudaStream_t st);
template <typename T> int launch_fd(const SpectralParams<T>& p, int equation, bool implicit, cudaStream_t st);
template <typename T> int launch_fd_reset(const SpectralParams<T>& p, const void* src, const uint8_t* mask, cudaStream_t st);
template <typename T>
int launch_spectral_aux(const SpectralParams<T>& p, int equation, int mode, const void* src, const uint8_t* mask,
                        void* dst, cudaStream_t st);

// Grid geometry of the warp-resident kernels: a team of min(N/2, 32) lanes per environment.
// Small batches get one warp per CTA so the 148 SMs fill evenly; large batches use 4-warp
// CTAs.  MPDE_WPC overrides (tuning).
inline void warp_geometry(int64_t B, int N, int& grid, int& block) {
    const int H = N / 2, TS = H < 32 ? H : 32, TPW = 32 / TS;
    const int64_t warps = (B + TPW - 1) / TPW;
    int wpc = warps >= 148 * 32 ? 4 : (warps >= 148 * 8 ? 2 : 1);
    if (const char* s = std::getenv("MPDE_WPC")) {
        const int v = std::atoi(s);
        if (v == 1 || v == 2 || v == 4) wpc = v;
    }
    grid = (int)((warps + wpc - 1) / wpc);
    block = 32 * wpc;
}
inline size_t warp_scratch_bytes(int N, int M, int block, size_t elem) {
    const int H = N / 2, TS = H < 32 ? H : 32, TPW = 32 / TS;
    const int scr = M > 2 * N + N / 2 ? M : 2 * N + N / 2;
    return (size_t)(block / 32) * TPW * scr * elem;
}

}  // namespace mpde
