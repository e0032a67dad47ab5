// Temporary stubs until the CTA / KS / FD kernels land.
#include "dispatch.h"
namespace mpde {
template <typename T> int launch_burgers_cta(const SpectralParams<T>&, cudaStream_t) { return -1; }
template <typename T> int launch_spectral_aux_cta(const SpectralParams<T>&, int, int, const void*, const uint8_t*, void*, cudaStream_t) { return -1; }
template <typename T> int launch_ks_cta(const SpectralParams<T>&, cudaStream_t) { return -1; }
template <typename T> int launch_fd(const SpectralParams<T>&, int, bool, cudaStream_t) { return -1; }
template <typename T> int launch_fd_reset(const SpectralParams<T>&, const void*, const uint8_t*, cudaStream_t) { return -1; }
#define INST(T) \
  template int launch_burgers_cta<T>(const SpectralParams<T>&, cudaStream_t); \
  template int launch_spectral_aux_cta<T>(const SpectralParams<T>&, int, int, const void*, const uint8_t*, void*, cudaStream_t); \
  template int launch_ks_cta<T>(const SpectralParams<T>&, cudaStream_t); \
  template int launch_fd<T>(const SpectralParams<T>&, int, bool, cudaStream_t); \
  template int launch_fd_reset<T>(const SpectralParams<T>&, const void*, const uint8_t*, cudaStream_t);
INST(double) INST(float)
}
