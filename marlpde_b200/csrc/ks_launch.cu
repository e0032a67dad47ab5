// Launchers of the Kuramoto-Sivashinsky ETDRK4 step (KS.py:230-274).
#include "dispatch.h"
#include "ks_warp.cuh"

namespace mpde {

// MINB = CTAs (of 64 threads) per SM the register allocation must allow: a batch of thousands of environments runs in
// several waves, so more resident warps (fewer waves) beats more registers per thread
template <typename T, int N, int TS, int MINB = 1, bool WW = false>
__global__ void __launch_bounds__(64, MINB) ks_warp_kernel(const SpectralParams<T> prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    KSWarp<T, N, TS, WW>::run(prm, reinterpret_cast<T*>(smem_raw));
}

template <typename T, int N, int TS, int MINB = 1, bool WW = false>
static int launch_ks_warp(const SpectralParams<T>& p, cudaStream_t st) {
    using K = KSWarp<T, N, TS, WW>;
    constexpr int TPW = K::TPW;
    const int64_t warps = (p.B + TPW - 1) / TPW;
    const int scr = 2 * K::R::SMEM_CX + (p.M > 2 * N + N / 2 ? p.M : 2 * N + N / 2);
    const size_t smem = ((size_t)2 * TPW * scr + 7 * (N / 2 + 1)) * sizeof(T);      // team areas + shared ETDRK4 tables
    // programmatic dependent launch, as for the Burgers kernels (the kernel reads mutable state after pdl_wait())
    static const bool pdl = [] { const char* s = std::getenv("MPDE_PDL"); return !(s && s[0] == '0'); }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((warps + 1) / 2));
    cfg.blockDim = dim3(64);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, ks_warp_kernel<T, N, TS, MINB, WW>, p);
    return 1;
}

// lanes per environment for N = 64: mpde_config.team_lanes, else MPDE_KS_TS (tuning), else the default
static int ks_team(int requested, int dflt) {
    if (requested == 8 || requested == 16 || requested == 32 || requested == -8) return requested;
    if (const char* s = std::getenv("MPDE_KS_TS")) {
        const int v = std::atoi(s);
        if (v == 8 || v == 16 || v == 32 || v == -8) return v;
    }
    return dflt;
}

template <typename T>
int launch_ks(const SpectralParams<T>& p, cudaStream_t st) {
    if (p.N == 64) {
        // B200, 8192 envs, 10 steps per launch: 32 lanes 122 us, 16 lanes 102.6 us, 8 lanes (radix-2 shuffles) 109 us,
        // 8 lanes with the transposed 4 x (2 x 4) transform (-8, default from round 2) 88.9 us
        switch (ks_team(p.team_lanes, -8)) {
            case 8: return launch_ks_warp<T, 64, 8>(p, st);
            case -8: {      // 8 lanes, transposed 4 x (2 x 4) transform (WarpFFT<T,32,-8>)
                int minb = 4;
                if (const char* s = std::getenv("MPDE_KS_MINB")) minb = std::atoi(s);
                if (minb == 6) return launch_ks_warp<T, 64, -8, 6>(p, st);
                if (minb == 5) return launch_ks_warp<T, 64, -8, 5>(p, st);
                if (p.hist_rows == 0) return launch_ks_warp<T, 64, -8, 4, true>(p, st);     // training: whole-warp collectives
                return launch_ks_warp<T, 64, -8, 4>(p, st);
            }
            case 32: return launch_ks_warp<T, 64, 32>(p, st);
            default: {      // B200, 8192 envs, 10 steps: 236 regs 108.8 us, 168 regs 110.9 us, 128 regs (some spills) 99.5 us
                int minb = 8;
                if (const char* s = std::getenv("MPDE_KS_MINB")) minb = std::atoi(s);
                if (minb == 6) return launch_ks_warp<T, 64, 16, 6>(p, st);
                if (minb == 1) return launch_ks_warp<T, 64, 16, 1>(p, st);
                return launch_ks_warp<T, 64, 16, 8>(p, st);
            }
        }
    }
    switch (p.N) {
        case 8: return launch_ks_warp<T, 8, 4>(p, st);
        case 16: return launch_ks_warp<T, 16, 8>(p, st);
        case 32: return launch_ks_warp<T, 32, 16>(p, st);
        case 64: return launch_ks_warp<T, 64, 32>(p, st);
        case 128: return launch_ks_warp<T, 128, 32>(p, st);
        default: return launch_ks_cta<T>(p, st);
    }
}

template int launch_ks<double>(const SpectralParams<double>&, cudaStream_t);
template int launch_ks<float>(const SpectralParams<float>&, cudaStream_t);

}  // namespace mpde
