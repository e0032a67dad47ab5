// Peer-memory gather of per-environment summaries over NVLink (no NCCL call, no host round trip):
// every rank's step kernel writes state + reward into a flat local slab; peer_put_kernel stores that
// slab into the gather buffer of EVERY rank (plain 16-byte stores to peer-mapped pointers travel over
// NVLink / NVSwitch) and then publishes a per-source step counter; peer_wait_kernel makes the consumer
// stream wait until all sources have published the current step.  Buffers come from cudaMalloc and are
// shared between the one-process-per-GPU ranks of a node through CUDA IPC handles.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/marlpde_b200.h"

namespace {
thread_local std::string g_perr;
int pfail(const std::string& m) { g_perr = m; return -1; }
#define PCU(call)                                                                      \
    do {                                                                               \
        cudaError_t _e = (call);                                                       \
        if (_e != cudaSuccess) return pfail(std::string(#call) + ": " + cudaGetErrorString(_e)); \
    } while (0)

constexpr int MAX_RANKS = 16;
struct PeerTable {
    void* dst[MAX_RANKS];          // gather buffer of rank p (peer-mapped)
    long long* flags[MAX_RANKS];   // flag array [nranks] of rank p (peer-mapped)
};

__global__ void peer_put_kernel(const int4* __restrict__ src, size_t n16, PeerTable tab, size_t dst_off16, int my_rank,
                                int nranks, long long step, unsigned int* counter) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const int4 v = src[i];
        for (int p = 0; p < nranks; ++p) reinterpret_cast<int4*>(tab.dst[p])[dst_off16 + i] = v;
    }
    __threadfence_system();                       // this block's stores are visible system-wide ...
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(counter, 1u);
        if (ticket == gridDim.x - 1) {            // ... and the last block to finish publishes the step
            *counter = 0;
            __threadfence_system();
            for (int p = 0; p < nranks; ++p) *reinterpret_cast<volatile long long*>(tab.flags[p] + my_rank) = step;
            __threadfence_system();
        }
    }
}

// Waits are bounded in TIME (%globaltimer), not in polls: a peer that does not publish within `timeout_ns` sets
// *err = 1 + its rank.  A timed-out wait does NOT advance the expected step, and once *err is set every later wait
// returns at once, so a dead peer costs one timeout, not one per step; the host sees *err (mapped pinned memory,
// mpde_host_flag_alloc) on its next call and raises.
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// returns true when slot r reached `want`
__device__ __forceinline__ bool spin_until(const long long* flags, int r, long long want, volatile int* err, long long timeout_ns) {
    const volatile long long* f = flags + r;
    if (*f >= want) return true;
    const unsigned long long t0 = global_ns();
    long long next_err_poll = 1000000;            // *err may live in mapped host memory: poll it once per millisecond only
    while (*f < want) {
        const long long waited = (long long)(global_ns() - t0);
        if (waited > next_err_poll) {
            if (*err) return false;               // an earlier wait already timed out: do not wait again
            next_err_poll += 1000000;
        }
        if (waited > timeout_ns) {
            if (*err == 0) *err = 1 + r;          // plain store: *err may be mapped host memory (any timed-out rank will do)
            return false;
        }
        __nanosleep(64);
    }
    return true;
}

__global__ void peer_wait_kernel(const long long* flags, int nranks, long long step, int* err, long long timeout_ns) {
    const int r = threadIdx.x;
    if (r < nranks) spin_until(flags, r, step, err, timeout_ns);
    __threadfence_system();
}
// producer side of the fused gather: runs BEHIND the step kernel in stream order (the kernel boundary makes the step
// kernel's peer stores visible), bumps the device-side step counter and publishes it in slot [my rank] of every
// rank's flag array.  One thread: its system-scope fence has nothing outstanding to wait for.
struct FlagTable { long long* f[MAX_RANKS]; };
__global__ void peer_signal_next_kernel(FlagTable tab, int n, long long* step) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const long long s = *step + 1;
        *step = s;
        __threadfence_system();
        for (int q = 0; q < n; ++q) *reinterpret_cast<volatile long long*>(tab.f[q]) = s;
    }
}
// consumer side of the fused gather: the expected step lives on the device (CUDA-graph replayable)
__global__ void peer_wait_next_kernel(const long long* flags, int nranks, long long* expect, int* err, long long timeout_ns) {
    __shared__ long long want;
    __shared__ int ok;
    if (threadIdx.x == 0) { want = *expect + 1; ok = 1; }
    __syncthreads();
    const int r = threadIdx.x;
    if (r < nranks && !spin_until(flags, r, want, err, timeout_ns)) ok = 0;
    __syncthreads();
    if (threadIdx.x == 0 && ok) *expect = want;
    __threadfence_system();
}
// signal + wait in one launch (what a symmetric learner loop does after every step)
__global__ void peer_exchange_next_kernel(FlagTable tab, int n, long long* step, const long long* flags, int nranks, long long* expect,
                                          int* err, long long timeout_ns) {
    __shared__ long long want;
    __shared__ int ok;
    if (threadIdx.x == 0) {
        const long long s = *step + 1;
        *step = s;
        __threadfence_system();
        for (int q = 0; q < n; ++q) *reinterpret_cast<volatile long long*>(tab.f[q]) = s;
        want = *expect + 1;
        ok = 1;
    }
    __syncthreads();
    const int r = threadIdx.x;
    if (r < nranks && !spin_until(flags, r, want, err, timeout_ns)) ok = 0;
    __syncthreads();
    if (threadIdx.x == 0 && ok) *expect = want;
    __threadfence_system();
}
long long timeout_ns_of(int64_t timeout_us) {
    if (timeout_us <= 0) {                        // default: MPDE_PEER_TIMEOUT_S seconds (30)
        static const long long dflt = [] {
            const char* s = std::getenv("MPDE_PEER_TIMEOUT_S");
            const double v = s ? std::atof(s) : 30.0;
            return (long long)((v > 0 ? v : 30.0) * 1e9);
        }();
        return dflt;
    }
    return (long long)timeout_us * 1000;
}
}  // namespace

extern "C" {

const char* mpde_peer_last_error(void) { return g_perr.c_str(); }

int mpde_peer_alloc(size_t bytes, void** out) {
    if (!out) return pfail("null argument");
    PCU(cudaMalloc(out, bytes));
    PCU(cudaMemset(*out, 0, bytes));
    return 0;
}
int mpde_peer_free(void* p) {
    PCU(cudaFree(p));
    return 0;
}
int mpde_host_flag_alloc(int32_t n_int32, void** out) {
    if (!out || n_int32 < 1) return pfail("host_flag_alloc: bad argument");
    PCU(cudaHostAlloc(out, (size_t)n_int32 * sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable));
    memset(*out, 0, (size_t)n_int32 * sizeof(int));
    return 0;
}
int mpde_host_flag_free(void* p) {
    PCU(cudaFreeHost(p));
    return 0;
}
int mpde_peer_export(void* dev_ptr, void* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    PCU(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle64), dev_ptr));
    return 0;
}
int mpde_peer_open(const void* handle64, void** out) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    PCU(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
int mpde_peer_close(void* p) {
    PCU(cudaIpcCloseMemHandle(p));
    return 0;
}

int mpde_peer_put(const void* src, size_t nbytes, void* const* dst_ptrs, size_t dst_offset_bytes, void* const* flag_ptrs,
                  int32_t my_rank, int32_t nranks, int64_t step, void* counter_dev, void* stream) {
    if (nranks < 1 || nranks > MAX_RANKS) return pfail("peer_put: 1..16 ranks");
    if ((nbytes & 15) || (dst_offset_bytes & 15)) return pfail("peer_put: sizes/offsets must be multiples of 16 bytes");
    PeerTable tab;
    for (int p = 0; p < nranks; ++p) {
        tab.dst[p] = dst_ptrs[p];
        tab.flags[p] = static_cast<long long*>(flag_ptrs[p]);
    }
    const size_t n16 = nbytes / 16;
    int grid = (int)((n16 + 255) / 256);
    if (grid > 64) grid = 64;
    if (grid < 1) grid = 1;
    peer_put_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const int4*>(src), n16, tab, dst_offset_bytes / 16,
                                                                         my_rank, nranks, (long long)step,
                                                                         static_cast<unsigned int*>(counter_dev));
    PCU(cudaGetLastError());
    return 0;
}

int mpde_peer_wait(const void* my_flags_dev, int32_t nranks, int64_t step, void* err_dev, int64_t timeout_us, void* stream) {
    if (nranks < 1 || nranks > MAX_RANKS) return pfail("peer_wait: 1..16 ranks");
    peer_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const long long*>(my_flags_dev), nranks, (long long)step,
                                                                      static_cast<int*>(err_dev), timeout_ns_of(timeout_us));
    PCU(cudaGetLastError());
    return 0;
}

int mpde_peer_signal_next(void* const* flag_ptrs, int32_t n, void* step_dev, void* stream) {
    if (n < 1 || n > MAX_RANKS || !flag_ptrs || !step_dev) return pfail("peer_signal_next: 1..16 flag slots and a step counter");
    FlagTable tab;
    for (int q = 0; q < n; ++q) tab.f[q] = static_cast<long long*>(flag_ptrs[q]);
    peer_signal_next_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(tab, n, static_cast<long long*>(step_dev));
    PCU(cudaGetLastError());
    return 0;
}

int mpde_peer_exchange_next(void* const* flag_ptrs, int32_t n, void* step_dev, const void* my_flags_dev, int32_t nranks,
                            void* expect_dev, void* err_dev, int64_t timeout_us, void* stream) {
    if (n < 1 || n > MAX_RANKS || nranks < 1 || nranks > MAX_RANKS || !flag_ptrs || !step_dev || !expect_dev)
        return pfail("peer_exchange_next: 1..16 ranks, flag slots and both counters");
    FlagTable tab;
    for (int q = 0; q < n; ++q) tab.f[q] = static_cast<long long*>(flag_ptrs[q]);
    peer_exchange_next_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(
        tab, n, static_cast<long long*>(step_dev), static_cast<const long long*>(my_flags_dev), nranks,
        static_cast<long long*>(expect_dev), static_cast<int*>(err_dev), timeout_ns_of(timeout_us));
    PCU(cudaGetLastError());
    return 0;
}

int mpde_peer_wait_next(const void* my_flags_dev, int32_t nranks, void* expect_dev, void* err_dev, int64_t timeout_us, void* stream) {
    if (nranks < 1 || nranks > MAX_RANKS) return pfail("peer_wait_next: 1..16 ranks");
    peer_wait_next_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const long long*>(my_flags_dev), nranks,
                                                                           static_cast<long long*>(expect_dev),
                                                                           static_cast<int*>(err_dev), timeout_ns_of(timeout_us));
    PCU(cudaGetLastError());
    return 0;
}

}  // extern "C"
