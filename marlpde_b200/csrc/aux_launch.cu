// Launchers of the reset / read-back kernels of the spectral solvers.
#include "dispatch.h"
#include "spectral_aux.cuh"

namespace mpde {

template <typename T, int N>
static int launch_aux_warp(const SpectralParams<T>& p, int equation, int mode, const void* src, const uint8_t* mask,
                           void* dst, cudaStream_t st) {
    int grid, block;
    warp_geometry(p.B, N, grid, block);
    switch (mode) {
        case AUX_RESET_U: aux_warp_kernel<T, N, AUX_RESET_U><<<grid, block, 0, st>>>(p, src, mask, dst, equation); break;
        case AUX_RESET_V: aux_warp_kernel<T, N, AUX_RESET_V><<<grid, block, 0, st>>>(p, src, mask, dst, equation); break;
        default: aux_warp_kernel<T, N, AUX_GET_U><<<grid, block, 0, st>>>(p, src, mask, dst, equation); break;
    }
    return 1;
}

template <typename T>
int launch_spectral_aux(const SpectralParams<T>& p, int equation, int mode, const void* src, const uint8_t* mask,
                        void* dst, cudaStream_t st) {
    switch (p.N) {
        case 8: return launch_aux_warp<T, 8>(p, equation, mode, src, mask, dst, st);
        case 16: return launch_aux_warp<T, 16>(p, equation, mode, src, mask, dst, st);
        case 32: return launch_aux_warp<T, 32>(p, equation, mode, src, mask, dst, st);
        case 64: return launch_aux_warp<T, 64>(p, equation, mode, src, mask, dst, st);
        case 128: return launch_aux_warp<T, 128>(p, equation, mode, src, mask, dst, st);
        case 256: return launch_aux_warp<T, 256>(p, equation, mode, src, mask, dst, st);
        default: return launch_spectral_aux_cta<T>(p, equation, mode, src, mask, dst, st);
    }
}

template int launch_spectral_aux<double>(const SpectralParams<double>&, int, int, const void*, const uint8_t*, void*, cudaStream_t);
template int launch_spectral_aux<float>(const SpectralParams<float>&, int, int, const void*, const uint8_t*, void*, cudaStream_t);

}  // namespace mpde
