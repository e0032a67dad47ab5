// Launchers of the CTA-per-environment spectral kernels (N = 256..2048): Burgers / KS DNS.
#include "dispatch.h"
#include "spectral_cta.cuh"

namespace mpde {

template <typename T, int N, int NT, int EQ>
__global__ void __launch_bounds__(NT) spectral_cta_kernel(const SpectralParams<T> prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SpectralCta<T, N, NT, EQ>::run(prm, smem_raw);
}

template <typename T, int N, int NT, int EQ>
static int launch_cta(const SpectralParams<T>& p, cudaStream_t st) {
    if (p.flags & (F_SSM | F_DSM)) return -2;                          // closures: warp kernels only
    if (p.reward_mode == REWARD_MSE && p.reward_out) return -3;        // MSE reward: warp kernels only
    const size_t smem = SpectralCta<T, N, NT, EQ>::smem_bytes(p.M);
    static size_t configured = 48 * 1024;       // dynamic shared memory above 48 KB is opt-in (227 KB max on sm_100)
    if (smem > 227 * 1024) return -5;
    if (smem > configured) {
        if (cudaFuncSetAttribute(spectral_cta_kernel<T, N, NT, EQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return -4;
        configured = smem;
    }
    spectral_cta_kernel<T, N, NT, EQ><<<(unsigned)p.B, NT, smem, st>>>(p);
    return 1;
}

template <typename T, int EQ>
static int launch_cta_n(const SpectralParams<T>& p, cudaStream_t st) {
    switch (p.N) {
        case 256: return launch_cta<T, 256, 32, EQ>(p, st);
        case 512: return launch_cta<T, 512, 64, EQ>(p, st);
        case 1024: return launch_cta<T, 1024, 128, EQ>(p, st);
        case 2048: return launch_cta<T, 2048, 256, EQ>(p, st);
        default: return -1;
    }
}

template <typename T> int launch_burgers_cta(const SpectralParams<T>& p, cudaStream_t st) { return launch_cta_n<T, 0>(p, st); }
template <typename T> int launch_ks_cta(const SpectralParams<T>& p, cudaStream_t st) { return launch_cta_n<T, 1>(p, st); }

template <typename T, int N, int NT>
static int launch_aux_cta(const SpectralParams<T>& p, int equation, int mode, const void* src, const uint8_t* mask, void* dst,
                          cudaStream_t st) {
    constexpr int H = N / 2, NH = H + 1;
    const size_t smem = sizeof(Cx<T>) * (2 * H + 2 * NH) + 16;
    static bool configured = false;
    if (!configured && smem > 48 * 1024) {
        cudaFuncSetAttribute(aux_cta_kernel<T, N, NT, AUX_RESET_U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(aux_cta_kernel<T, N, NT, AUX_RESET_V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(aux_cta_kernel<T, N, NT, AUX_GET_U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    switch (mode) {
        case AUX_RESET_U: aux_cta_kernel<T, N, NT, AUX_RESET_U><<<(unsigned)p.B, NT, smem, st>>>(p, src, mask, dst, equation); break;
        case AUX_RESET_V: aux_cta_kernel<T, N, NT, AUX_RESET_V><<<(unsigned)p.B, NT, smem, st>>>(p, src, mask, dst, equation); break;
        default: aux_cta_kernel<T, N, NT, AUX_GET_U><<<(unsigned)p.B, NT, smem, st>>>(p, src, mask, dst, equation); break;
    }
    return 1;
}

template <typename T>
int launch_spectral_aux_cta(const SpectralParams<T>& p, int equation, int mode, const void* src, const uint8_t* mask, void* dst,
                            cudaStream_t st) {
    switch (p.N) {
        case 512: return launch_aux_cta<T, 512, 64>(p, equation, mode, src, mask, dst, st);
        case 1024: return launch_aux_cta<T, 1024, 128>(p, equation, mode, src, mask, dst, st);
        case 2048: return launch_aux_cta<T, 2048, 256>(p, equation, mode, src, mask, dst, st);
        default: return -1;
    }
}

#define INST(T)                                                                                                   \
    template int launch_burgers_cta<T>(const SpectralParams<T>&, cudaStream_t);                                   \
    template int launch_ks_cta<T>(const SpectralParams<T>&, cudaStream_t);                                        \
    template int launch_spectral_aux_cta<T>(const SpectralParams<T>&, int, int, const void*, const uint8_t*, void*, cudaStream_t);
INST(double)
INST(float)

}  // namespace mpde
