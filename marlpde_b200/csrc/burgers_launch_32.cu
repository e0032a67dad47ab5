// Launchers of the Burgers environment step for N = 32 (the LES grid of the reference drivers):
// three team sizes x compile-time specialised mode flags.
#include "burgers_dispatch.cuh"

namespace mpde {

template <typename T>
int launch_burgers_32(const SpectralParams<T>& p, cudaStream_t st) {
    switch (pick_team(p.B, 32, 16, 4)) {
        case 16: return launch_warp_sf<T, 32, 16>(p, st);
        case 8: return launch_warp_sf<T, 32, 8>(p, st);
        default: return launch_warp_sf<T, 32, 4>(p, st);
    }
}

template int launch_burgers_32<double>(const SpectralParams<double>&, cudaStream_t);
template int launch_burgers_32<float>(const SpectralParams<float>&, cudaStream_t);

}  // namespace mpde
