// Device generator of the stochastic-forcing tables.
//
// The reference seeds NumPy's legacy generator per environment and draws two (32, nsteps) tables of standard normals
// (/root/reference/python/_model/Burger.py:66 `np.random.seed(seed)`, :88-89 optional `0.01 + 0.02 * uniform()`,
// :94-95 `randfac1 = normal(size=(32, nsteps))`, `randfac2 = ...`), of which the solver only ever reads rows 1..3,
// columns < stepper (:416-419).  NumPy's legacy stream = MT19937 seeded by init_genrand, 53-bit doubles from two
// 32-bit outputs, and the polar Box-Muller `legacy_gauss` with one cached deviate (third-party: NumPy, pinned
// numpy==1.20.1; restated and pinned against NumPy in oracle/mt19937_oracle.py).
//
// One WARP per seed.  The rejection loop looks sequential, but every candidate pair consumes exactly four 32-bit
// outputs, so candidate j sits at a fixed position of the output stream: the twister block (624 words in shared
// memory) is regenerated in parallel, tempered in parallel, all 156 candidates of the block are tested in parallel,
// and a ballot / popc running count locates the accepted pair behind each wanted draw.  log / sqrt are evaluated only
// for the 6 * stepper draws that are kept (their last bit may differ from glibc's: <= 1 ulp on the table entry; which
// candidates are accepted is decided in exact IEEE arithmetic and is bit-identical).
#include <cuda_runtime.h>
#include <cstdint>
#include <string>

#include "../../include/marlpde_b200.h"

namespace {

constexpr int MT_N = 624, MT_M = 397, WARPS = 4;

__device__ __forceinline__ double u53(unsigned a32, unsigned b32) {
    const double a = (double)(a32 >> 5), b = (double)(b32 >> 6);
    return __ddiv_rn(__dadd_rn(__dmul_rn(a, 67108864.0), b), 9007199254740992.0);
}

__global__ void __launch_bounds__(32 * WARPS) forcing_tables_kernel(const long long* __restrict__ seeds, long long n, int nsteps,
                                                                     int stepper, int nunoise, double* __restrict__ r1,
                                                                     double* __restrict__ r2, double* __restrict__ nu_out) {
    __shared__ unsigned mt_s[WARPS][MT_N];
    __shared__ unsigned out_s[WARPS][MT_N + 4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long e = (long long)blockIdx.x * WARPS + w;
    if (e >= n) return;
    unsigned* mt = mt_s[w];
    unsigned* ob = out_s[w];
    if (lane == 0) {                                  // init_genrand(seed & 0xffffffff)
        unsigned x = (unsigned)(seeds[e] & 0xffffffffLL);
        mt[0] = x;
        for (int i = 1; i < MT_N; ++i) {
            x = 1812433253u * (x ^ (x >> 30)) + (unsigned)i;
            mt[i] = x;
        }
    }
    __syncwarp();
    const int per_tbl = 3 * stepper, nq = 2 * per_tbl;
    int q = 0;                   // next wanted draw (ascending draw index)
    long long acc = 0;           // accepted candidate pairs so far
    int carry = 0;               // outputs left over from the previous block (ob[0..carry))
    bool first = true;
    while (q < nq) {
        // ---- next twister block: mt[kk] = mt[kk+397] ^ (y >> 1) ^ mag(y), y = (mt[kk] & UP) | (mt[kk+1] & LOW) ----
        // chunks of 32 in index order; reads of a chunk complete before its writes, so mt[kk+1] is always the old
        // word and mt[(kk+397) % 624] the old (kk < 227) or the already regenerated (kk >= 227) one, as in genrand
        for (int c0 = 0; c0 < MT_N; c0 += 32) {
            const int kk = c0 + lane;
            unsigned nv = 0;
            if (kk < MT_N) {
                const unsigned y = (mt[kk] & 0x80000000u) | (mt[kk + 1 == MT_N ? 0 : kk + 1] & 0x7fffffffu);
                const int km = kk + MT_M >= MT_N ? kk + MT_M - MT_N : kk + MT_M;
                nv = mt[km] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            __syncwarp();
            if (kk < MT_N) {
                mt[kk] = nv;
                unsigned t = nv;                       // tempering
                t ^= t >> 11;
                t ^= (t << 7) & 0x9d2c5680u;
                t ^= (t << 15) & 0xefc60000u;
                t ^= t >> 18;
                ob[carry + kk] = t;
            }
            __syncwarp();
        }
        int avail = carry + MT_N, start = 0;
        if (first) {
            first = false;
            if (nunoise) {                             // Burger.py:88-89: nu = 0.01 + 0.02 * uniform(), drawn before the tables
                if (lane == 0 && nu_out) nu_out[e] = __dadd_rn(0.01, __dmul_rn(0.02, u53(ob[0], ob[1])));
                start = 2;
            }
        }
        const int ncand = (avail - start) >> 2;
        for (int base = 0; base < ncand && q < nq; base += 32) {
            const int j = base + lane;
            double x1 = 0, x2 = 0, rr = 2.0;
            if (j < ncand) {
                const unsigned* o = ob + start + 4 * j;
                x1 = __dadd_rn(__dmul_rn(2.0, u53(o[0], o[1])), -1.0);
                x2 = __dadd_rn(__dmul_rn(2.0, u53(o[2], o[3])), -1.0);
                rr = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));
            }
            const bool ok = rr < 1.0 && rr != 0.0;
            const unsigned mask = __ballot_sync(0xffffffffu, ok);
            const int cnt = __popc(mask);
            while (q < nq) {
                const int tbl = q / per_tbl, rem = q - tbl * per_tbl, k = rem / stepper, c = rem - k * stepper;
                const long long t = (long long)tbl * 32 * nsteps + (long long)(k + 1) * nsteps + c;      // draw index
                const long long pair = t >> 1;
                if (pair >= acc + cnt) break;
                const int src = __fns(mask, 0, (int)(pair - acc) + 1);
                if (lane == src) {
                    const double f = sqrt(__ddiv_rn(__dmul_rn(-2.0, log(rr)), rr));
                    const double val = (t & 1) ? __dmul_rn(f, x1) : __dmul_rn(f, x2);        // cached deviate = f * x1 comes second
                    (tbl == 0 ? r1 : r2)[(e * 3 + k) * stepper + c] = val;
                }
                ++q;
            }
            acc += cnt;
        }
        const int used = start + 4 * ncand, left = avail - used;
        __syncwarp();
        unsigned keep = 0;
        if (lane < left) keep = ob[used + lane];
        __syncwarp();
        if (lane < left) ob[lane] = keep;
        carry = left;
        __syncwarp();
    }
}

thread_local std::string g_rerr;
int rfail(const std::string& m) { g_rerr = m; return -1; }

}  // namespace

extern "C" {

const char* mpde_rng_last_error(void) { return g_rerr.c_str(); }

int mpde_forcing_tables(const int64_t* seeds_dev, int64_t n, int32_t nsteps, int32_t stepper, int32_t nunoise, double* r1_dev,
                        double* r2_dev, double* nu_dev, void* stream) {
    if (!seeds_dev || !r1_dev || !r2_dev || n < 1) return rfail("forcing_tables: null argument");
    if (nsteps < 1 || stepper < 1 || stepper > nsteps) return rfail("forcing_tables: need 1 <= stepper <= nsteps");
    if (nunoise && !nu_dev) return rfail("forcing_tables: nunoise needs nu_dev");
    const unsigned grid = (unsigned)((n + WARPS - 1) / WARPS);
    forcing_tables_kernel<<<grid, 32 * WARPS, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const long long*>(seeds_dev), (long long)n, nsteps, stepper, nunoise, r1_dev, r2_dev, nu_dev);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return rfail(std::string("forcing_tables: ") + cudaGetErrorString(e));
    return 0;
}

}  // extern "C"
