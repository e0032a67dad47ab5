// Launchers of the Burgers environment step (Burger.py:333-499).
#include "dispatch.h"
#include "burgers_warp.cuh"

namespace mpde {

template <typename T, int N>
__global__ void __launch_bounds__(128) burgers_warp_kernel(const SpectralParams<T> prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BurgersWarp<T, N>::run(prm, reinterpret_cast<T*>(smem_raw));
}

template <typename T, int N>
static int launch_warp(const SpectralParams<T>& p, cudaStream_t st) {
    int grid, block;
    warp_geometry(p.B, N, grid, block);
    const size_t smem = warp_scratch_bytes(N, p.M, block, sizeof(T));
    burgers_warp_kernel<T, N><<<grid, block, smem, st>>>(p);
    return 1;
}

template <typename T>
int launch_burgers(const SpectralParams<T>& p, cudaStream_t st) {
    switch (p.N) {
        case 8: return launch_warp<T, 8>(p, st);
        case 16: return launch_warp<T, 16>(p, st);
        case 32: return launch_warp<T, 32>(p, st);
        case 64: return launch_warp<T, 64>(p, st);
        case 128: return launch_warp<T, 128>(p, st);
        case 256: return launch_warp<T, 256>(p, st);
        default: return launch_burgers_cta<T>(p, st);
    }
}

template int launch_burgers<double>(const SpectralParams<double>&, cudaStream_t);
template int launch_burgers<float>(const SpectralParams<float>&, cudaStream_t);

}  // namespace mpde
