// Half spectrum <-> full FFT-order spectrum copies for mpde_get / mpde_set.
#pragma once
#include "common.cuh"

namespace mpde {

// half spectrum <-> full FFT-order spectrum, for mpde_get / mpde_set
template <typename T>
__global__ void unpack_half_kernel(const Cx<T>* __restrict__ half, Cx<T>* __restrict__ full, int64_t B, int N) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * N) return;
    const int64_t e = i / N;
    const int k = (int)(i - e * N);
    const int NH = N / 2 + 1;
    Cx<T> a = ldcx(half + e * NH + (k <= N / 2 ? k : N - k));
    if (k > N / 2) a = conj(a);
    stcx(full + i, a);
}
template <typename T>
__global__ void pack_half_kernel(const Cx<T>* __restrict__ full, Cx<T>* __restrict__ half, int64_t B, int N) {
    const int NH = N / 2 + 1;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * NH) return;
    const int64_t e = i / NH;
    const int k = (int)(i - e * NH);
    stcx(half + i, ldcx(full + e * N + k));
}

}  // namespace mpde
