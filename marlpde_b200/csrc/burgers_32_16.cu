// Burgers step kernels for N = 32, 16 lanes per environment (all mode specialisations).
#include "burgers_dispatch.cuh"
namespace mpde {
MPDE_INSTANTIATE_TEAM(32, 16)
}
