// Launchers of the CTA-per-environment spectral kernels (N = 256..2048): Burgers / KS DNS.
#include "dispatch.h"
#include "spectral_cta.cuh"
#include "dns_warp.cuh"
#include <map>

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
    // dynamic shared memory above 48 KB is opt-in (227 KB max on sm_100); the attribute is per DEVICE, so the bookkeeping is too
    static std::map<int, size_t> configured;
    if (smem > 227 * 1024) return -5;
    int dev = 0;
    cudaGetDevice(&dev);
    size_t& have = configured.emplace(dev, 48 * 1024).first->second;
    if (smem > have) {
        if (cudaFuncSetAttribute(spectral_cta_kernel<T, N, NT, EQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return -4;
        have = smem;
    }
    spectral_cta_kernel<T, N, NT, EQ><<<(unsigned)p.B, NT, smem, st>>>(p);
    return 1;
}

// N = 1024 Burgers DNS (no actions / closures / state / reward): one warp per environment, state in registers (dns_warp.cuh)
template <typename T, bool VS>
__global__ void __launch_bounds__(32) burgers_dns1024_kernel(const SpectralParams<T> prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Dns1024<T, VS>::run(prm, smem_raw);
}
template <typename T>
__global__ void __launch_bounds__(64) burgers_dns1024x2_kernel(const SpectralParams<T> prm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Dns1024x2<T>::run(prm, smem_raw);
}
template <typename T>
static bool dns_warp_eligible(const SpectralParams<T>& p) {
    static const bool on = [] { const char* s = std::getenv("MPDE_DNS_WARP"); return !(s && s[0] == '0'); }();
    return on && p.N == 1024 && !(p.flags & (F_ACTIONS | F_SSM | F_DSM | F_FD | F_NO_ADVANCE)) && p.nsub > 0 && !p.state_out && !p.reward_out;
}
template <typename T>
static int launch_dns1024(const SpectralParams<T>& p, cudaStream_t st) {
    // MPDE_DNS_WARPS=1: one warp per environment (round-2 first design); MPDE_DNS_VREG=1: that kernel with the spectrum in
    // registers; default: two warps per environment (Dns1024x2)
    static const bool vreg = [] { const char* s = std::getenv("MPDE_DNS_VREG"); return s && s[0] == '1'; }();
    static const bool one_warp = [] { const char* s = std::getenv("MPDE_DNS_WARPS"); return s && s[0] == '1'; }();
    if (vreg) {
        burgers_dns1024_kernel<T, false><<<(unsigned)p.B, 32, Dns1024<T, false>::smem_bytes(), st>>>(p);
    } else if (one_warp) {
        burgers_dns1024_kernel<T, true><<<(unsigned)p.B, 32, Dns1024<T, true>::smem_bytes(), st>>>(p);
    } else {
        const size_t smem = Dns1024x2<T>::smem_bytes();
        static std::map<int, bool> configured;
        int dev = 0;
        cudaGetDevice(&dev);
        if (!configured[dev]) {
            if (cudaFuncSetAttribute(burgers_dns1024x2_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -4;
            configured[dev] = true;
        }
        burgers_dns1024x2_kernel<T><<<(unsigned)p.B, 64, smem, st>>>(p);
    }
    return 1;
}

template <typename T, int EQ>
static int launch_cta_n(const SpectralParams<T>& p, cudaStream_t st) {
    if (EQ == 0 && dns_warp_eligible(p)) return launch_dns1024<T>(p, st);
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
    static std::map<int, bool> configured_dev;
    int dev = 0;
    cudaGetDevice(&dev);
    bool& configured = configured_dev[dev];
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
