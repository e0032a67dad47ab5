// Temporary stubs until the CTA / KS / FD kernels land.
#include "dispatch.h"
namespace mpde {
template <typename T> int launch_fd(const SpectralParams<T>&, int, bool, cudaStream_t) { return -1; }
template <typename T> int launch_fd_reset(const SpectralParams<T>&, const void*, const uint8_t*, cudaStream_t) { return -1; }
#define INST(T) \
  template int launch_fd<T>(const SpectralParams<T>&, int, bool, cudaStream_t); \
  template int launch_fd_reset<T>(const SpectralParams<T>&, const void*, const uint8_t*, cudaStream_t);
INST(double) INST(float)
}
