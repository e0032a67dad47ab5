// Launchers of the Burgers environment step for N = 64 (the LES grid of the reference drivers):
// three team sizes x compile-time specialised mode flags.
#include "burgers_dispatch.cuh"

namespace mpde {

template <typename T>
int launch_burgers_64(const SpectralParams<T>& p, cudaStream_t st) {
    switch (pick_team(p.B, 64, 32, 8)) {
        case 32: return launch_warp_sf<T, 64, 32>(p, st);
        case 16: return launch_warp_sf<T, 64, 16>(p, st);
        default: return launch_warp_sf<T, 64, 8>(p, st);
    }
}

template int launch_burgers_64<double>(const SpectralParams<double>&, cudaStream_t);
template int launch_burgers_64<float>(const SpectralParams<float>&, cudaStream_t);

}  // namespace mpde
