// B200 FP64 / shuffle micro-benchmarks (latency of dependent chains, throughput with ILP).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_ffma(float* out, int iters, float a, float b) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP>
__global__ void k_shfl(double* out, int iters) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = __shfl_xor_sync(0xffffffffu, x[i], 1 + (it & 7));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// one radix-2 butterfly stage chain: shuffle + fma + complex multiply (what the FFT does)
template <int ILP>
__global__ void k_stage(double* out, int iters, double wr, double wi) {
    double re[ILP], im[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { re[i] = threadIdx.x + i; im[i] = 1.0 / (threadIdx.x + 1 + i); }
    const double sg = (threadIdx.x & 1) ? -1.0 : 1.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            const double ore = __shfl_xor_sync(0xffffffffu, re[i], 1 + (it & 7));
            const double oim = __shfl_xor_sync(0xffffffffu, im[i], 1 + (it & 7));
            const double tr = fma(sg, re[i], ore), ti = fma(sg, im[i], oim);
            re[i] = tr * wr - ti * wi;
            im[i] = tr * wi + ti * wr;
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += re[i] + im[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("%s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, clk_khz);
    double* out; cudaMalloc(&out, 148 * 64 * 1024 * sizeof(double));
    const int iters = 4096;
    const double ghz = clk_khz * 1e-6;
#define RUN(name, kern, ilp, grid, block, ops_per_iter, ...)                                                      \
    {                                                                                                             \
        float ms = timeit([&] { kern<ilp><<<grid, block>>>(__VA_ARGS__); });                                     \
        double cyc = ms * 1e-3 * ghz * 1e9;                                                                      \
        double warps = (double)grid * block / 32;                                                                 \
        printf("%-28s ilp=%d grid=%5d block=%4d  cycles/iter=%8.2f  per-op latency/throughput: %6.2f cyc/op/warp, " \
               "chip %.3g op-lanes/s\n", name, ilp, grid, block, cyc / iters, cyc / iters / (ilp * ops_per_iter),   \
               warps * 32 * ilp * ops_per_iter * iters / (ms * 1e-3));                                           \
    }
    RUN("dfma dependent 1 warp", k_dfma, 1, 1, 32, 1, out, iters, 1.0000001, 1e-9)
    RUN("dfma ilp2 1 warp", k_dfma, 2, 1, 32, 1, out, iters, 1.0000001, 1e-9)
    RUN("dfma ilp4 1 warp", k_dfma, 4, 1, 32, 1, out, iters, 1.0000001, 1e-9)
    RUN("dfma ilp8 1 warp", k_dfma, 8, 1, 32, 1, out, iters, 1.0000001, 1e-9)
    RUN("dfma ilp8 4 warps/SM", k_dfma, 8, 148, 128, 1, out, iters, 1.0000001, 1e-9)
    RUN("dfma ilp8 16 warps/SM", k_dfma, 8, 148, 512, 1, out, iters, 1.0000001, 1e-9)
    RUN("dfma ilp8 32 warps/SM", k_dfma, 8, 148, 1024, 1, out, iters, 1.0000001, 1e-9)
    RUN("dfma ilp1 14 warps/SM", k_dfma, 1, 148 * 7, 64, 1, out, iters, 1.0000001, 1e-9)
    RUN("dfma ilp2 14 warps/SM", k_dfma, 2, 148 * 7, 64, 1, out, iters, 1.0000001, 1e-9)
    RUN("ffma ilp8 32 warps/SM", k_ffma, 8, 148, 1024, 1, (float*)out, iters, 1.0000001f, 1e-9f)
    RUN("ffma dependent 1 warp", k_ffma, 1, 1, 32, 1, (float*)out, iters, 1.0000001f, 1e-9f)
    RUN("shfl64 dependent 1 warp", k_shfl, 1, 1, 32, 1, out, iters)
    RUN("shfl64 ilp4 1 warp", k_shfl, 4, 1, 32, 1, out, iters)
    RUN("shfl64 ilp8 16 warps/SM", k_shfl, 8, 148, 512, 1, out, iters)
    RUN("stage dependent 1 warp", k_stage, 1, 1, 32, 1, out, iters, 0.9238795, -0.3826834)
    RUN("stage ilp2 1 warp", k_stage, 2, 1, 32, 1, out, iters, 0.9238795, -0.3826834)
    RUN("stage ilp4 1 warp", k_stage, 4, 1, 32, 1, out, iters, 0.9238795, -0.3826834)
    RUN("stage ilp1 14 warps/SM", k_stage, 1, 148 * 7, 64, 1, out, iters, 0.9238795, -0.3826834)
    RUN("stage ilp2 14 warps/SM", k_stage, 2, 148 * 7, 64, 1, out, iters, 0.9238795, -0.3826834)
    RUN("stage ilp2 7 warps/SM", k_stage, 2, 148 * 7, 32, 1, out, iters, 0.9238795, -0.3826834)
    RUN("stage ilp4 7 warps/SM", k_stage, 4, 148 * 7, 32, 1, out, iters, 0.9238795, -0.3826834)
    RUN("stage ilp4 4 warps/SM", k_stage, 4, 148 * 4, 32, 1, out, iters, 0.9238795, -0.3826834)
    RUN("stage ilp8 4 warps/SM", k_stage, 8, 148 * 4, 32, 1, out, iters, 0.9238795, -0.3826834)
    RUN("stage ilp8 2 warps/SM", k_stage, 8, 148 * 2, 32, 1, out, iters, 0.9238795, -0.3826834)
    return 0;
}
