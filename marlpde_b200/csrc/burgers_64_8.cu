// Burgers step kernels for N = 64, 8 lanes per environment (all mode specialisations).
#include "burgers_dispatch.cuh"
namespace mpde {
MPDE_INSTANTIATE_TEAM(64, 8)
}
