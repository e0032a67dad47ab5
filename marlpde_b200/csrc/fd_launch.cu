// Finite-difference environments: Diffusion (Diffusion.py:137-216, 238-298), Advection
// (Advection.py:138-213, 235-286), DiffusionError (DiffusionError.py:137-216: the action perturbs the Laplacian
// stencil) and Laplace (Laplace.py:116-166: relaxation with three free stencil entries per agent, one Dirichlet point).  The reference builds a dense N x N matrix per step and
// evaluates M @ u; M is a periodic tridiagonal stencil whose entries are the agents' actions,
// so the kernel applies the 3-point stencil directly.  One warp per environment, the row u and
// the three coefficient rows live in shared memory, `nsub` steps are fused per launch.
#include "dispatch.h"

namespace mpde {

constexpr int EQ_DIFFUSION = 2, EQ_ADVECTION = 3, EQ_DIFFUSION_ERROR = 4, EQ_LAPLACE = 5;

template <typename T>
__global__ void __launch_bounds__(32) fd_step_kernel(const SpectralParams<T> prm, int equation, int implicit) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = prm.N, lane = threadIdx.x;
    const int64_t e = blockIdx.x;
    T* ua = reinterpret_cast<T*>(smem_raw);
    T* ub = ua + N;
    T* lo = ub + N;
    T* di = lo + N;
    T* up = di + N;
    const T dt = prm.dt, dx = prm.dx, nu = prm.nu[e];
    const bool was_live = prm.status[e] == 0;
    int iout = prm.iout[e];
    T tnow = prm.tnow[e];
    for (int n = lane; n < N; n += 32) ua[n] = prm.uprev[e * N + n];

    // ---- stencil rows from the actions -----------------------------------------------------------
    const bool has_act = prm.flags & F_ACTIONS;
    const T* act = prm.actions + e * prm.M;
    for (int n = lane; n < N; n += 32) {
        T l, d, u;
        if (equation == EQ_DIFFUSION) {
            // one global weight a -> (-a/2, a, -a/2) (Diffusion.py:172-178); per point a_k (:186-200)
            const T a = has_act ? act[prm.M == 1 ? 0 : n] : T(-2);
            l = -a / T(2); d = a; u = -a / T(2);
        } else if (equation == EQ_DIFFUSION_ERROR) {
            // the action is the ERROR of the stencil: (1 - a/2, -2 + a, 1 - a/2) (DiffusionError.py:166-190); with ONE agent
            // the two wrap-around entries are M[0,-1] = 1 - ac[0] and M[-1,0] = 1 + ac[2], as the reference writes them (:171-172)
            const T a = has_act ? act[prm.M == 1 ? 0 : n] : T(0);
            l = T(1) - a / T(2); d = T(-2) + a; u = T(1) - a / T(2);
            if (has_act && prm.M == 1) {
                if (n == 0) l = T(1) - (T(1) - a / T(2));
                if (n == N - 1) u = T(1) + (T(1) - a / T(2));
            }
        } else if (equation == EQ_LAPLACE) {
            // row n = i + 1 of agent i: entries on columns i, i + 1, (i + 2) % N; row 0 stays zero (Laplace.py:121-129)
            if (has_act && n >= 1) { l = act[3 * (n - 1)]; d = act[3 * (n - 1) + 1]; u = act[3 * (n - 1) + 2]; }
            else { l = T(0); d = T(0); u = T(0); }
        } else {
            if (!has_act) {           // Lax (Advection.py:142-148), alpha = nu dt / dx as set at construction
                const T al = prm.etd[e];
                l = T(0.5) + T(0.5) * al; d = T(0); u = T(0.5) - T(0.5) * al;
            } else if (prm.M == 2) {  // global: a0 on u_{k-1}, a1 on u_{k+1}, diagonal 1 - sum (Advection.py:163-169)
                l = act[0]; u = act[1]; d = T(1) - (act[0] + act[1]);
            } else {                  // per point: entry 2k on u_{k+1}, 2k+1 on u_{k-1} (:181-194) ...
                u = act[2 * n]; l = act[2 * n + 1];
                d = T(1) - act[2 * n] - act[2 * n + 1];
                if (n == N - 1) { l = act[2 * n]; u = act[2 * n + 1]; }      // ... except the last row (:188-190)
            }
        }
        lo[n] = l; di[n] = d; up[n] = u;
    }
    __syncwarp();

    const int nsub = (prm.flags & F_NO_ADVANCE) ? 0 : prm.nsub;
    bool bad = false;
    T* cur = ua;
    T* nxt = ub;
    for (int it = 0; it < nsub; ++it) {
        if ((equation == EQ_DIFFUSION || equation == EQ_DIFFUSION_ERROR) && !has_act && implicit) {
            // implicit Euler: (I - c Lap) u' = u, periodic tridiagonal (Diffusion.py:142-149), solved by the
            // Thomas algorithm with a Sherman-Morrison correction for the two corner entries
            if (lane == 0) {
                const T c = dt * nu / (dx * dx);
                const T a = -c, b = T(1) + T(2) * c;
                T* cp = nxt;                 // modified super-diagonal
                T* y = lo;                   // solution of A' y = u   (lo/di are free: no actions in this mode)
                T* z = di;                   // solution of A' z = w
                const T gamma = -b;
                // A' = A - w v^T, w = [gamma, 0, ..., 0, a], v = [1, 0, ..., 0, a/gamma]
                T bb0 = b - gamma, bbn = b - a * a / gamma;
                T denom = bb0;
                cp[0] = a / denom;
                y[0] = cur[0] / denom;
                z[0] = gamma / denom;
                for (int i = 1; i < N; ++i) {
                    const T bi = (i == N - 1) ? bbn : b;
                    denom = bi - a * cp[i - 1];
                    cp[i] = a / denom;
                    y[i] = (cur[i] - a * y[i - 1]) / denom;
                    z[i] = (((i == N - 1) ? a : T(0)) - a * z[i - 1]) / denom;
                }
                for (int i = N - 2; i >= 0; --i) {
                    y[i] -= cp[i] * y[i + 1];
                    z[i] -= cp[i] * z[i + 1];
                }
                const T fact = (y[0] + a * y[N - 1] / gamma) / (T(1) + z[0] + a * z[N - 1] / gamma);
                for (int i = 0; i < N; ++i) nxt[i] = y[i] - fact * z[i];
            }
            __syncwarp();
        } else {
            for (int n = lane; n < N; n += 32) {
                const T um = cur[n == 0 ? N - 1 : n - 1], u0 = cur[n], upv = cur[n == N - 1 ? 0 : n + 1];
                T r;
                if ((equation == EQ_DIFFUSION || equation == EQ_DIFFUSION_ERROR) && !has_act) {
                    // explicit Euler, standard Laplacian (Diffusion.py:156-160)
                    const T d2 = (T(-2) * u0 + um + upv) / (dx * dx);
                    r = u0 + dt * nu * d2;
                } else {
                    const T a = lo[n] * um, b = di[n] * u0, c = up[n] * upv;
                    // a dense row-times-vector sums in column order: rows 0 and N-1 wrap
                    const T mv = n == 0 ? (b + c) + a : (n == N - 1 ? (c + a) + b : (a + b) + c);
                    if (equation == EQ_ADVECTION) r = mv;                                           // Advection.py:200
                    else if (equation == EQ_LAPLACE) r = n == 0 ? T(1) : u0 + dt * mv;            // Laplace.py:131-136 (u[0] = 1)
                    else r = u0 + dt * nu * mv / (dx * dx);                                       // Diffusion.py:206, DiffusionError.py:195
                }
                nxt[n] = r;
                bad |= !(fabs((double)r) <= 1.79e308);
            }
            __syncwarp();
        }
        T* tmp = cur; cur = nxt; nxt = tmp;
        iout += 1;
        tnow += dt;
        if (prm.uu_hist && iout < prm.hist_rows)
            for (int n = lane; n < N; n += 32) prm.uu_hist[(e * prm.hist_rows + iout) * N + n] = cur[n];
    }
    const bool blew = __any_sync(0xffffffffu, bad);
    const bool live = was_live && !blew;
    if (nsub > 0) {
        if (was_live && blew && lane == 0) prm.status[e] = 1;
        if (live) {
            for (int n = lane; n < N; n += 32) prm.uprev[e * N + n] = cur[n];
            if (lane == 0) { prm.iout[e] = iout; prm.tnow[e] = tnow; }
        }
    }
    const T inf = T(1) / T(0);
    const int A = prm.A;
    if (equation == EQ_LAPLACE) {
        // force row: prm.truth [ntruth][1][N]; state [u_{i-1}, u_i, u_{i+1}, force_i] for i < N - 1 (Laplace.py:162-166);
        // direct reward -(u_xx - force)^2 on points 1..N-1 (:153-160)
        const T* force = prm.truth ? prm.truth + (prm.truth_map ? prm.truth_map[e] : 0) * prm.truth_rows * N : nullptr;
        const int nA = N - 1;
        if (prm.state_out && force) {
            for (int o = lane; o < 4 * nA; o += 32) {
                const int i = o >> 2, c = o & 3;
                const T val = c == 3 ? force[i] : cur[(i - 1 + c + N) % N];
                prm.state_out[e * 4 * nA + o] = live ? val : inf;
            }
        }
        if (prm.reward_out && prm.reward_mode == REWARD_DIRECT && force) {
            for (int n = 1 + lane; n < N; n += 32) {
                const T um = cur[n - 1], u0 = cur[n], upv = cur[n == N - 1 ? 0 : n + 1];
                const T d2 = (T(-2) * u0 + um + upv) / (dx * dx);
                const T df = d2 - force[n];
                prm.reward_out[e * nA + (n - 1)] = live ? -(df * df) : -inf;
            }
        }
        return;
    }
    if (prm.state_out) {      // getState: u, or per-agent windows with a one-point halo (Diffusion.py:284-298)
        const int seg = A == 1 ? N : N / A + 2;
        const int S = A * seg;
        for (int o = lane; o < S; o += 32) {
            const int a = o / seg, w = o - a * seg;
            const int j = A == 1 ? w : ((a * (N / A) - 1 + w) % N + N) % N;
            prm.state_out[e * S + o] = live ? cur[j] : inf;
        }
    }
    if (prm.reward_out && prm.reward_mode == REWARD_MSE && prm.truth) {
        // getMseReward: -mean((truth - u)^2) per agent section at the current time (Diffusion.py:245-252)
        const int64_t tb = (prm.truth_map ? prm.truth_map[e] : 0) * prm.truth_rows;
        const int64_t row = iout < prm.truth_rows ? iout : prm.truth_rows - 1;
        const T* tr = prm.truth + (tb + row) * N;
        const int W = N / A;
        for (int a = lane; a < A; a += 32) {
            T sum = T(0);
            for (int j = 0; j < W; ++j) { const T d = tr[a * W + j] - cur[a * W + j]; sum += d * d; }
            prm.reward_out[e * A + a] = live ? -(sum / T(W)) : -inf;
        }
    } else if (prm.reward_out && prm.reward_mode == REWARD_DIRECT) {
        // getDirectReward: -(d2u/dx2)^2 / numAgents per grid point (Diffusion.py:275-281)
        for (int n = lane; n < N; n += 32) {
            const T um = cur[n == 0 ? N - 1 : n - 1], u0 = cur[n], upv = cur[n == N - 1 ? 0 : n + 1];
            const T d2 = (T(-2) * u0 + um + upv) / (dx * dx);
            prm.reward_out[e * N + n] = live ? -(d2 * d2) / T(A) : -inf;
        }
    }
}

template <typename T>
__global__ void fd_reset_kernel(const SpectralParams<T> prm, const T* __restrict__ src, const uint8_t* __restrict__ mask) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= prm.B * prm.N) return;
    const int64_t e = i / prm.N;
    if (mask && !mask[e]) return;
    const int n = (int)(i - e * prm.N);
    prm.uprev[i] = src[i];
    if (prm.uu_hist && prm.hist_rows > 0) prm.uu_hist[(e * prm.hist_rows) * prm.N + n] = src[i];
    if (n == 0) { prm.iout[e] = 0; prm.tnow[e] = T(0); prm.status[e] = 0; prm.kprev[e] = T(0); }
}

template <typename T>
int launch_fd(const SpectralParams<T>& p, int equation, bool implicit, cudaStream_t st) {
    const size_t smem = (size_t)5 * p.N * sizeof(T);
    if (smem > 227 * 1024) return -1;
    static size_t configured = 48 * 1024;
    if (smem > configured) {
        if (cudaFuncSetAttribute(fd_step_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        configured = smem;
    }
    fd_step_kernel<T><<<(unsigned)p.B, 32, smem, st>>>(p, equation, implicit ? 1 : 0);
    return 1;
}

template <typename T>
int launch_fd_reset(const SpectralParams<T>& p, const void* src, const uint8_t* mask, cudaStream_t st) {
    const int64_t n = p.B * p.N;
    fd_reset_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, static_cast<const T*>(src), mask);
    return 1;
}

#define INST(T)                                                                       \
    template int launch_fd<T>(const SpectralParams<T>&, int, bool, cudaStream_t);     \
    template int launch_fd_reset<T>(const SpectralParams<T>&, const void*, const uint8_t*, cudaStream_t);
INST(double)
INST(float)

}  // namespace mpde
