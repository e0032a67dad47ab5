// Burgers step kernels for N = 32, 8 lanes per environment, radix-2^2 shuffle transform (bit-identical to the 4-lane kernels).
#include "burgers_dispatch.cuh"
namespace mpde {
MPDE_INSTANTIATE_TEAM_AS(32, 8x, -8)
}
