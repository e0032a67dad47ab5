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

// out[i] = (T(in[i]), T(1) / T(in[i])): the reward's reference spectrum with its reciprocals (set-up, once per
// mpde_set_spectrum_ref)
template <typename T>
__global__ void rcp_table_kernel(const double* __restrict__ in, Cx<T>* __restrict__ out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) stcx(out + i, cx<T>((T)in[i], T(1) / (T)in[i]));
}
template <typename T>
int launch_rcp_table(const double* in, Cx<T>* out, int64_t n, cudaStream_t st) {
    if (n > 0) rcp_table_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, out, n);
    return 1;
}
template int launch_rcp_table<double>(const double*, Cx<double>*, int64_t, cudaStream_t);
template int launch_rcp_table<float>(const double*, Cx<float>*, int64_t, cudaStream_t);

template int launch_spectral_aux<double>(const SpectralParams<double>&, int, int, const void*, const uint8_t*, void*, cudaStream_t);
template int launch_spectral_aux<float>(const SpectralParams<float>&, int, int, const void*, const uint8_t*, void*, cudaStream_t);

}  // namespace mpde
