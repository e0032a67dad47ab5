// Shared launcher template of the warp-resident Burgers kernels.
#pragma once
#include "dispatch.h"
#include "burgers_warp.cuh"

namespace mpde {

// register budget: one wave of warps must fit for the small-batch (latency) regime
template <typename T, int N, int TS>
constexpr int min_blocks() {
    constexpr int P = N / 2 / TS;
    return (sizeof(T) == 8 ? (P <= 1 ? 8 : (P <= 2 ? 6 : 4)) : (P <= 2 ? 8 : 4));
}

// 256 threads x (min_blocks / 4) CTAs give the same register budget as 64 x min_blocks
template <typename T, int N, int TS, int SF, bool LEAN>
__global__ void __launch_bounds__(256, (min_blocks<T, N, TS>() / 4 > 0 ? min_blocks<T, N, TS>() / 4 : 1)) burgers_warp_kernel(const SpectralParams<T> prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BurgersWarp<T, N, TS, SF, LEAN>::run(prm, reinterpret_cast<T*>(smem_raw));
}

template <typename T, int N, int TS, int SF, bool LEAN = false>
int launch_warp(const SpectralParams<T>& p, cudaStream_t st) {
    constexpr int TPW = 32 / TS;
    const int64_t warps = (p.B + TPW - 1) / TPW;
    int wpc = 2;                                   // warps per CTA (MPDE_WPC = 1, 2, 4, 8 overrides: tuning)
    if (const char* s = std::getenv("MPDE_WPC")) {
        const int v = std::atoi(s);
        if (v == 1 || v == 2 || v == 4 || v == 8) wpc = v;
    }
    const int block = 32 * wpc;
    const int grid = (int)((warps + wpc - 1) / wpc);
    const int scr = p.M > 2 * N + N / 2 ? p.M : 2 * N + N / 2;
    const size_t smem = (size_t)wpc * TPW * scr * sizeof(T);
    burgers_warp_kernel<T, N, TS, SF, LEAN><<<grid, block, smem, st>>>(p);
    return 1;
}

// structural-flag specialisations compiled for the hot grid sizes (fp64); LEAN drops the
// history / MSE / multi-column-forcing branches when the call does not need them
template <typename T, int N, int TS, int SF>
int launch_warp_lean(const SpectralParams<T>& p, cudaStream_t st) {
    const bool lean = p.hist_rows == 0 && !(p.reward_mode == REWARD_MSE && p.truth) && p.stepper == 1;
    return lean ? launch_warp<T, N, TS, SF, true>(p, st) : launch_warp<T, N, TS, SF, false>(p, st);
}
template <typename T, int N, int TS>
int launch_warp_sf(const SpectralParams<T>& p, cudaStream_t st) {
    if constexpr (sizeof(T) == 8) {
        switch (p.flags & STRUCT_FLAGS) {
            case F_ACTIONS: return launch_warp_lean<T, N, TS, F_ACTIONS>(p, st);
            case F_ACTIONS | F_FORCING: return launch_warp_lean<T, N, TS, F_ACTIONS | F_FORCING>(p, st);
            case F_ACTIONS | F_DFORCE: return launch_warp_lean<T, N, TS, F_ACTIONS | F_DFORCE>(p, st);
            case F_ACTIONS | F_DFORCE | F_FORCING: return launch_warp_lean<T, N, TS, F_ACTIONS | F_DFORCE | F_FORCING>(p, st);
            case 0:
            case F_DFORCE: return launch_warp_lean<T, N, TS, 0>(p, st);
            case F_FORCING:
            case F_FORCING | F_DFORCE: return launch_warp_lean<T, N, TS, F_FORCING>(p, st);
            default: break;
        }
    }
    return launch_warp<T, N, TS, -1>(p, st);
}

// one translation unit per (N, team size): defines launch_burgers_<N>_<TS><T>
#define MPDE_INSTANTIATE_TEAM(N_, TS_)                                                                         \
    template <typename T> int launch_burgers_##N_##_##TS_(const SpectralParams<T>& p, cudaStream_t st) {      \
        return launch_warp_sf<T, N_, TS_>(p, st);                                                              \
    }                                                                                                          \
    template int launch_burgers_##N_##_##TS_<double>(const SpectralParams<double>&, cudaStream_t);            \
    template int launch_burgers_##N_##_##TS_<float>(const SpectralParams<float>&, cudaStream_t);

// Team size: the widest team (lowest latency) unless the batch is large enough to keep every
// SM sub-partition busy with narrower teams (fewer instructions per environment).
// MPDE_TS overrides (tuning).
inline int pick_team(int64_t B, int N, int ts_max, int ts_min) {
    if (const char* s = std::getenv("MPDE_TS")) {
        const int v = std::atoi(s);
        if (v >= ts_min && v <= ts_max && (v & (v - 1)) == 0) return v;
    }
    int ts = ts_max;
    while (ts > ts_min && (B * (ts / 2) / 32) >= (int64_t)148 * 4 * MPDE_WARPS_PER_SMSP_TARGET) ts /= 2;
    return ts;
}

}  // namespace mpde
