// Kernel parameter blocks shared between the C-ABI layer (capi.cu) and the kernels.
#pragma once
#include <cstdint>
#include "common.cuh"

namespace mpde {

enum : int {
    F_DFORCE = 1 << 0,        // actions are a direct forcing (else they scale d2u/dx2, Burger.py:445-450)
    F_FORCING = 1 << 1,       // 3-mode stochastic forcing (Burger.py:410-423)
    F_SSM = 1 << 2,           // static Smagorinsky (Burger.py:337-352)
    F_DSM = 1 << 3,           // dynamic Smagorinsky (Burger.py:354-408)
    F_ACTIONS = 1 << 4,       // an action array was supplied to this call
    F_BASIS_DENSE = 1 << 5,   // basis has >2 non-zeros in some column: use the dense [M,N] table
    F_FORCING_PER_ENV = 1 << 6,
    F_NO_ADVANCE = 1 << 7,    // nsub == 0: only evaluate the state (getState without stepping)
    // 1 << 8: F_KS_UUROW (ks_warp.cuh)
    F_FD = 1 << 9,            // Burger_fd: explicit Euler + finite differences in real space (Burger_fd.py:335-476)
    F_SSMFORCE = 1 << 10,     // Burger_fd(ssmforce=True): the action is a Smagorinsky coefficient (Burger_fd.py:447-455)
};

enum : int { REWARD_NONE = 0, REWARD_SPECTRAL = 1, REWARD_MSE = 2, REWARD_DIRECT = 3 };
enum : int { AUX_RESET_U = 0, AUX_RESET_V = 1, AUX_GET_U = 2 };

// Multi-GPU gather fused into the step kernel's epilogue: besides state_out / reward_out (this rank's slab in its
// own gather buffer) every state / reward store is repeated into this rank's slab of each peer's gather buffer
// (peer-mapped pointers, plain stores over NVLink).  Publishing the step to the peers is NOT done here: a
// system-scope fence in this kernel costs ~6 us on B200, so a 1-thread signal kernel behind the kernel boundary
// (peer.cu) does it off the critical path.
constexpr int MAX_PEERS = 8;
struct PeerSink {
    int n_data = 0;                 // peers that receive a copy of the outputs (other ranks)
    int parity = 0;                 // which copy of the double-buffered gather buffers this launch writes
    long long parity_stride = 0;    // elements of T between the two copies (0 = single-buffered): step s writes copy
                                    // s & 1, locally and on the peers, so a fast rank never overwrites rows a slow
                                    // rank's learner is still reading
    void* state[MAX_PEERS] = {};    // peer p: where this rank's [B,S] state slab lives in p's gather buffer
    void* reward[MAX_PEERS] = {};   // peer p: where this rank's [B,A] reward slab lives
    // NVSwitch multicast (NVLS) alternative: ONE multimem.st per row lands in every rank's buffer (this rank's
    // included), so a rank sends its slab once instead of once per peer.  Set -> replaces the local store and
    // the per-peer loop.
    void* mc_state = nullptr;
    void* mc_reward = nullptr;
    // single-agent state rows: 1 = staged through shared memory and written as whole rows (256 contiguous bytes per 16
    // lanes: what the NVLink write efficiency of an 8-GPU gather wants), 0 = every lane stores its own 16-byte pieces
    // to every destination (shorter epilogue; 64-byte segments per 4-lane team)
    int row_stores = 1;
};

template <typename T>
struct SpectralParams {
    int64_t B;
    int N, M, A, version, stepper, nsub, flags, reward_mode;
    int team_lanes;         // 0 = default team size for this N (mpde_config.team_lanes)
    T dt, dx;
    // read-only tables
    const Cx<T>* tw;        // [N/2]  exp(-2 pi i j / N)
    const T* kwave;         // [N]    dimensional wavenumbers, FFT order (fftfreq)
    const T* nu;            // [B]
    const int* tap_idx;     // [N][2] sparse basis: f_j = w0 a[i0] + w1 a[i1]
    const T* tap_w;         // [N][2]
    const T* basis;         // [M][N] dense basis
    const Cx<T>* fcoef;     // [B or 1][stepper][3] spectrum of the stochastic forcing at k = 1,2,3
    const T* etd;           // KS: [6][N]  E, E2, Q, f1, f2, f3 (FFT order)
    // persistent per-env state (owned by the library)
    Cx<T>* v;               // [B][N/2+1] half spectrum (the Nyquist entry keeps its imaginary part)
    Cx<T>* fn;              // [B][N/2+1] Fn_old of the AB2 scheme
    float* acc;             // [B][N/2+1] running float32 sum of the energy spectrum rows
    int* iout;              // [B] ioutnum
    T* tnow;                // [B] accumulated time (t += dt per step)
    T* kprev;               // [B] kPrevRelErr of the spectral reward
    int* status;            // [B] 0 = running, 1 = truncated (numerical blow-up)
    T* uprev;               // [B][N] previous real-space row (state version 1 / KS float32 row)
    // per-call I/O (caller-owned device buffers)
    const T* actions;       // [B][M]
    T* state_out;           // [B][S]
    T* reward_out;          // [B][A]
    // references
    const double* ek_ref;   // [nref][ek_rows][N/2] time-averaged DNS spectrum rows
    const Cx<T>* ek_pair;   // same shape: (T(ek_ref), T(1) / T(ek_ref)), filled by the library when the reference is set: one
                            // 16-byte load per wavenumber, and the reward's divisions become correction steps
                            // (common.cuh: div_by_rcp)
    int64_t ek_rows;
    const int* ek_map;      // [B] env -> ref index (nullptr: all use 0)
    const T* truth;         // [ntruth][truth_rows][N] DNS truth interpolated on the env grid
    int64_t truth_rows;
    const int* truth_map;   // [B]
    // optional history (caller-owned)
    T* uu_hist;             // [B][hist_rows][N]
    Cx<float>* vv_hist;     // [B][hist_rows][N] complex64, FFT order
    double* ektt_hist;      // [B][hist_rows][N/2+1] running time-average of the spectrum (Ek_ktt)
    int64_t hist_rows;
    PeerSink peer;
};

}  // namespace mpde
