// Launchers of the Burgers environment step (Burger.py:333-499).
#include "burgers_dispatch.cuh"

namespace mpde {

template <typename T>
static int launch_burgers_32(const SpectralParams<T>& p, cudaStream_t st) {
    switch (pick_team(p.team_lanes, p.B, 32, p.flags, 16, 4)) {
        case 16: return launch_burgers_32_16<T>(p, st);
        case 8: return launch_burgers_32_8<T>(p, st);
        case -8: return launch_burgers_32_8x<T>(p, st);
        default: return launch_burgers_32_4<T>(p, st);
    }
}
template <typename T>
static int launch_burgers_64(const SpectralParams<T>& p, cudaStream_t st) {
    switch (pick_team(p.team_lanes, p.B, 64, p.flags, 32, 8)) {
        case 32: return launch_burgers_64_32<T>(p, st);
        case 16: return launch_burgers_64_16<T>(p, st);
        default: return launch_burgers_64_8<T>(p, st);
    }
}

template <typename T>
int launch_burgers(const SpectralParams<T>& p, cudaStream_t st) {
    switch (p.N) {
        case 8: return (p.flags & F_FD) ? launch_warp<T, 8, 4, -2>(p, st) : launch_warp<T, 8, 4, -1>(p, st);
        case 16: return (p.flags & F_FD) ? launch_warp<T, 16, 8, -2>(p, st) : launch_warp<T, 16, 8, -1>(p, st);
        case 32: return launch_burgers_32<T>(p, st);
        case 64: return launch_burgers_64<T>(p, st);
        case 128: return (p.flags & F_FD) ? launch_warp<T, 128, 32, -2>(p, st) : launch_warp<T, 128, 32, -1>(p, st);
        case 256: return (p.flags & F_FD) ? launch_warp<T, 256, 32, -2>(p, st) : launch_warp<T, 256, 32, -1>(p, st);
        default: return launch_burgers_cta<T>(p, st);
    }
}

template int launch_burgers<double>(const SpectralParams<double>&, cudaStream_t);
template int launch_burgers<float>(const SpectralParams<float>&, cudaStream_t);

}  // namespace mpde
