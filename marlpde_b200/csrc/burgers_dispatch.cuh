// Shared launcher template of the warp-resident Burgers kernels.
#pragma once
#include "dispatch.h"
#include "burgers_warp.cuh"

namespace mpde {

// Register budget.  CTAs are 64 threads (2 warps).  The hot case -- fp64, N = 32, 16 lanes per environment,
// B = 4096 -> 2048 warps = 13.8 per SM -- must be resident in ONE wave: 7 CTAs per SM = 144 registers per thread
// (at 128 ptxas shuffles 64-bit values through XOR swaps and moves: +50 instructions per sub-step).
template <typename T, int N, int TS>
constexpr int min_blocks() {
    constexpr int P = N / 2 / TS;
    if (sizeof(T) == 8 && N == 32 && P == 1) return 7;
    return 4 * (sizeof(T) == 8 ? (P <= 1 ? 2 : 1) : (P <= 2 ? 2 : 1));
}
#if defined(MPDE_LB_T) && defined(MPDE_LB_B)      // tuning override: -DMPDE_LB_T=64 -DMPDE_LB_B=7
#define MPDE_LB MPDE_LB_T, MPDE_LB_B
#else
#define MPDE_LB 64, min_blocks<T, N, TS>()
#endif
constexpr int WARPS_PER_CTA = 2;

template <typename T, int N, int TS, int SF, int LEAN>
__global__ void __launch_bounds__(MPDE_LB) burgers_warp_kernel(const SpectralParams<T> prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BurgersWarp<T, N, TS, SF, LEAN>::run(prm, reinterpret_cast<T*>(smem_raw));
}

template <typename T, int N, int TS, int SF, int LEAN = 0>
int launch_warp(const SpectralParams<T>& p, cudaStream_t st) {
    constexpr int TPW = 32 / TS;
    const int64_t warps = (p.B + TPW - 1) / TPW;
    constexpr int wpc = WARPS_PER_CTA;
    const int block = 32 * wpc;
    const int grid = (int)((warps + wpc - 1) / wpc);
    const size_t smem = (size_t)wpc * TPW * BurgersWarp<T, N, TS, SF, LEAN>::scratch_doubles(p.M) * sizeof(T);
    // programmatic dependent launch: this kernel may become resident (and read its constant tables) while the
    // previous kernel of the stream drains; it reads mutable state only after pdl_wait().  MPDE_PDL=0 disables.
    static const bool pdl = [] { const char* s = std::getenv("MPDE_PDL"); return !(s && s[0] == '0'); }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, burgers_warp_kernel<T, N, TS, SF, LEAN>, p);
    return 1;
}

// structural-flag specialisations compiled for the hot grid sizes (fp64); LEAN drops the
// history / MSE / multi-column-forcing branches when the call does not need them
template <typename T, int N, int TS, int SF>
int launch_warp_lean(const SpectralParams<T>& p, cudaStream_t st) {
    const bool lean = p.hist_rows == 0 && !(p.reward_mode == REWARD_MSE && p.truth) && p.stepper == 1 && p.version != 1;
    const bool hot = lean && p.nsub > 0 && p.state_out && p.A == 1 && (p.version == 0 || p.version == 2) &&
                     !(p.flags & F_BASIS_DENSE) && p.reward_mode != REWARD_MSE;
    if (hot) return launch_warp<T, N, TS, SF, 2>(p, st);
    return lean ? launch_warp<T, N, TS, SF, 1>(p, st) : launch_warp<T, N, TS, SF, 0>(p, st);
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

// Team size = lanes per environment.  It is a function of N ONLY (two complex points per lane: N = 32 -> 8 lanes,
// N = 64 -> 16 lanes), never of the batch size: the variants factor the FFT differently, so their results differ in
// the last bits, and environment e must give the same bits whether it runs alone or inside a batch of 65536.
// Measured on B200, fp64 N = 32, B = 4096, 10 sub-steps: 16 lanes 15.2 us, 8 lanes 13.9 us, 4 lanes 14.7 us per launch
// (the 4-lane shared-memory-transpose variant has the lowest per-step slope and wins from B ~ 8192 per GPU).
// mpde_config.team_lanes (Burger(team_lanes=...)) or MPDE_TS (whole process: tuning / tests) select another variant;
// not bitwise compatible with the default.
inline int pick_team(int requested, int N, int ts_max, int ts_min) {
    if (requested >= ts_min && requested <= ts_max && (requested & (requested - 1)) == 0) return requested;   // mpde_config.team_lanes
    if (const char* s = std::getenv("MPDE_TS")) {
        const int v = std::atoi(s);
        if (v >= ts_min && v <= ts_max && (v & (v - 1)) == 0) return v;
    }
    const int ts = N / 4;           // P = (N/2) / ts = 2 complex points per lane
    return ts > ts_max ? ts_max : (ts < ts_min ? ts_min : ts);
}

}  // namespace mpde
