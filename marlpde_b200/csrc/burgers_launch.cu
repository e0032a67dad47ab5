// Launchers of the Burgers environment step (Burger.py:333-499): grid sizes other than 32 / 64.
#include "burgers_dispatch.cuh"

namespace mpde {

template <typename T>
int launch_burgers(const SpectralParams<T>& p, cudaStream_t st) {
    switch (p.N) {
        case 8: return launch_warp<T, 8, 4, -1>(p, st);
        case 16: return launch_warp<T, 16, 8, -1>(p, st);
        case 32: return launch_burgers_32<T>(p, st);
        case 64: return launch_burgers_64<T>(p, st);
        case 128: return launch_warp<T, 128, 32, -1>(p, st);
        case 256: return launch_warp<T, 256, 32, -1>(p, st);
        default: return launch_burgers_cta<T>(p, st);
    }
}

template int launch_burgers<double>(const SpectralParams<double>&, cudaStream_t);
template int launch_burgers<float>(const SpectralParams<float>&, cudaStream_t);

}  // namespace mpde
