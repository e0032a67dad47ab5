// Shared launcher template of the warp-resident Burgers kernels.
#pragma once
#include "dispatch.h"
#include "burgers_warp.cuh"

namespace mpde {

// Register budget.  CTAs are 64 threads (2 warps).  The hot case -- fp64, N = 32, 16 lanes per environment,
// B = 4096 -> 2048 warps = 13.8 per SM -- must be resident in ONE wave: 7 CTAs per SM = 144 registers per thread
// (at 128 ptxas shuffles 64-bit values through XOR swaps and moves: +50 instructions per sub-step).
template <typename T, int N, int TSTAG>
constexpr int min_blocks() {
    constexpr int TS = TSTAG < 0 ? -TSTAG : TSTAG;
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
    constexpr int TPW = BurgersWarp<T, N, TS, SF, LEAN>::TPW;
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
// history / MSE / multi-column-forcing branches when the call does not need them (burgers_warp.cuh: LEAN)
template <typename T, int N, int TS, int SF>
int launch_warp_lean(const SpectralParams<T>& p, cudaStream_t st) {
    const bool slim = p.hist_rows == 0 && p.stepper == 1 && p.version != 1;
    const bool mse = p.reward_mode == REWARD_MSE && p.truth;
    const bool lean = slim && !mse;
    const bool hot = lean && p.nsub > 0 && p.state_out && p.A == 1 && (p.version == 0 || p.version == 2) &&
                     !(p.flags & F_BASIS_DENSE) && p.reward_mode != REWARD_MSE;
    if (hot) return launch_warp<T, N, TS, SF, 2>(p, st);
    // LEAN = 3: the multi-agent training configuration (MSE reward against a truth table, any agent count) without the
    // history / multi-column / u_prev branches in the sub-step loop; compiled for the action-driven modes only
    if constexpr ((SF & F_ACTIONS) != 0)
        if (slim && mse && p.nsub > 0 && !(p.flags & F_BASIS_DENSE)) return launch_warp<T, N, TS, SF, 3>(p, st);
    return lean ? launch_warp<T, N, TS, SF, 1>(p, st) : launch_warp<T, N, TS, SF, 0>(p, st);
}
template <typename T, int N, int TS>
int launch_warp_sf(const SpectralParams<T>& p, cudaStream_t st) {
    if (p.flags & F_FD) return launch_warp<T, N, TS, -2>(p, st);      // Burger_fd: its own instantiation of the generic kernel
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
#define MPDE_INSTANTIATE_TEAM_AS(N_, NAME_, TSTAG_)                                                            \
    template <typename T> int launch_burgers_##N_##_##NAME_(const SpectralParams<T>& p, cudaStream_t st) {    \
        return launch_warp_sf<T, N_, TSTAG_>(p, st);                                                           \
    }                                                                                                          \
    template int launch_burgers_##N_##_##NAME_<double>(const SpectralParams<double>&, cudaStream_t);          \
    template int launch_burgers_##N_##_##NAME_<float>(const SpectralParams<float>&, cudaStream_t);
#define MPDE_INSTANTIATE_TEAM(N_, TS_) MPDE_INSTANTIATE_TEAM_AS(N_, TS_, TS_)

// Team size = lanes per environment: a function of N ONLY by default (N = 32 -> 4 lanes, N = 64 -> 16 lanes), never of the
// batch size, because the variants round differently and environment e must give the same bits alone or inside a batch
// of 65536.  B200, fp64 N = 32, 10 fused sub-steps, us per launch, one batch at a time:
//   B = 4096: 16 lanes 14.8, 8 lanes 13.4, 4 lanes 13.9;  B = 8192: 24.2 / 25.7 / 19.5;  B = 32768: 86.6 / 89.9 / 64.0;
// B = 4096 with 4 independent batches in flight (round 2): 16 lanes 10.1, 8 lanes 10.4, 4 lanes 7.2.
// mpde_config.team_lanes (Burger(team_lanes=...)) or MPDE_TS (whole process) select a variant:
//   4 / 8 / 16 : that many lanes;
//   -8         : 8 lanes with the radix-2^2 shuffle network whose arithmetic is bit-identical to the 4-lane kernel
//                (5 % slower than the plain 8-lane kernel: more selects);
//   -1         : "consistent auto": -8 below ~6000 environments, 4 lanes above -- results do not depend on the batch size
//                although the kernel does (N = 32, not with the dynamic Smagorinsky closure, whose means are summed in
//                team order).
inline int pick_team(int requested, int64_t B, int N, int flags, int ts_max, int ts_min) {
    if (const char* s = std::getenv("MPDE_TS"))
        if (requested == 0) requested = std::atoi(s);
    if (N == 32 && requested == -1) return (B >= 6144 && !(flags & F_DSM)) ? 4 : -8;
    if (N == 32 && requested == -8) return -8;
    if (requested >= ts_min && requested <= ts_max && (requested & (requested - 1)) == 0) return requested;
    // default: N = 32 -> 4 lanes per environment (the 4 x 4 shared-memory-transposed transform: fewest instructions per
    // environment; with several independent batches in flight it is also the fastest at B = 4096: 7.2 us per launch
    // against 10.4 us with 8 lanes, profiles/r2_lanes_chains.md); other N -> two complex points per lane
    const int ts = N == 32 ? 4 : N / 4;
    return ts > ts_max ? ts_max : (ts < ts_min ? ts_min : ts);
}

}  // namespace mpde
