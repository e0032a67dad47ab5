// C ABI of marlpde_b200 (see include/marlpde_b200.h).  Host-side handle management and
// kernel dispatch; no torch, no CPU compute path: every solver call launches sm_100a kernels.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/marlpde_b200.h"
#include "dispatch.h"

namespace {

thread_local std::string g_err;

int fail(const std::string& msg) {
    g_err = msg;
    return -1;
}

#define CU(call)                                                                          \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess)                                                            \
            return fail(std::string(#call) + ": " + cudaGetErrorString(_e));              \
    } while (0)

bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

}  // namespace

using namespace mpde;

struct mpde_env {
    mpde_config cfg{};
    int64_t launches = 0;
    int64_t epoch = 0;      // bumped by every setter: what a step launch looks like may have changed (graph cache key)
    int aux_flags = 0;
    virtual ~mpde_env() {}
    virtual int init() = 0;
    virtual int set_nu(const double* nu, int64_t n) = 0;
    virtual int set_basis(int M, const double* basis) = 0;
    virtual int set_etd(const double* const tabs[6]) = 0;
    virtual int set_forcing(const double* coef, int64_t n) = 0;
    virtual int set_spectrum_ref(const double* ek, int64_t nref, int64_t rows, const int32_t* map) = 0;
    virtual int set_truth(const void* truth, int64_t ntruth, int64_t rows, const int32_t* map) = 0;
    virtual int set_history(void* uu, void* vv, double* ektt, int64_t rows) = 0;
    virtual int reset(const void* src, bool spectral, const uint8_t* mask, cudaStream_t st) = 0;
    virtual int reset_handoff(const void* vsrc, int64_t nsrc, int nsrc_points, const double* ksrc, const int32_t* src_map,
                              const double* offset, const uint8_t* mask, cudaStream_t st) = 0;
    virtual int reset_turbulence(const int64_t* seed, const double* offset, const double* x, const double* amp,
                                 const uint8_t* mask, cudaStream_t st) = 0;
    virtual int step(const void* actions, int nsub, void* state_out, void* reward_out, cudaStream_t st) = 0;
    virtual int step_host(const void* actions, int nsub, void* state_out, void* reward_out, cudaStream_t st) = 0;
    virtual int step_host_packed(const void* actions, int nsub, void* out, cudaStream_t st) = 0;
    virtual int set_peer_output(int n_data, void* const* state, void* const* reward, int64_t parity_stride, void* mc_state,
                                void* mc_reward) = 0;
    virtual int set_peer_local(void* state, void* reward) = 0;
    virtual int set_peer_row_stores(int on) = 0;
    virtual int set_peer_sync(void* const* flag_slots, int n, void* step_dev, const void* my_flags, int nranks, void* expect_dev,
                              void* err, int64_t timeout_us) = 0;
    virtual int step_fused(const void* actions, int nsub, void* state_out, void* reward_out, int async, cudaStream_t st) = 0;
    virtual int peer_join(cudaStream_t st) = 0;
    virtual int compute_sgs(int nURG, int64_t rows, void* sgs, void* alt, void* alt2, cudaStream_t st) = 0;
    virtual int get(int field, void* dst, cudaStream_t st) = 0;
    virtual int set(int field, const void* src, cudaStream_t st) = 0;
    int64_t state_size() const {
        const int N = cfg.N, A = cfg.num_agents, ver = cfg.version;
        if (cfg.equation == MPDE_BURGERS) {
            const int nf = (ver == 1 || ver == 2) ? 2 : 1;
            const int seg = A == 1 ? N : N / A + 2;
            const int tail = (ver == 3 || ver == 4) ? N / 2 : 0;
            return (int64_t)A * (nf * seg + tail);
        }
        if (cfg.equation == MPDE_KS) return 2 * N;
        if (cfg.equation == MPDE_LAPLACE) return 4 * (int64_t)(N - 1);
        return A == 1 ? N : (int64_t)A * (N / A + 2);
    }
};

template <typename T>
struct Env : mpde_env {
    SpectralParams<T> prm{};
    std::vector<void*> owned;
    Cx<T>* ek_rcp_buf = nullptr;    // (value, reciprocal) pairs of the spectrum reference (set_spectrum_ref)
    int64_t ek_rcp_cap = 0;
    int n_forcing = 0;

    template <typename U>
    int dalloc(U** p, size_t count) {
        void* q = nullptr;
        CU(cudaMalloc(&q, count * sizeof(U) + 16));
        CU(cudaMemset(q, 0, count * sizeof(U) + 16));
        owned.push_back(q);
        *p = static_cast<U*>(q);
        return 0;
    }
    template <typename U>
    int upload(U* dst, const std::vector<U>& src) {
        // the tables may be read by step kernels in flight on non-blocking streams, which a legacy-stream copy does
        // not order against: drain the device first (set-up path, once per episode at most)
        CU(cudaDeviceSynchronize());
        CU(cudaMemcpy(dst, src.data(), src.size() * sizeof(U), cudaMemcpyHostToDevice));
        return 0;
    }
    ~Env() override {
        for (const HostGraph& g : host_graphs) cudaGraphExecDestroy(g.exec);
        if (sync.side) { cudaStreamDestroy(sync.side); cudaEventDestroy(sync.ev_fork); cudaEventDestroy(sync.ev_join); }
        for (void* p : owned) cudaFree(p);
    }

    bool spectral() const { return cfg.equation == MPDE_BURGERS || cfg.equation == MPDE_KS; }

    int init() override {
        CU(cudaSetDevice(cfg.device));
        const int N = cfg.N;
        const int64_t B = cfg.nenvs;
        prm.B = B;
        prm.N = N;
        prm.M = cfg.M;
        prm.A = cfg.num_agents;
        prm.version = cfg.version;
        prm.stepper = cfg.stepper;
        prm.reward_mode = cfg.reward_mode;
        prm.team_lanes = cfg.team_lanes;
        prm.dt = (T)cfg.dt;
        prm.dx = (T)(cfg.L / N);
        if (spectral()) {
            // twiddles exp(-2 pi i j / N), j < N/2
            std::vector<Cx<T>> tw(N / 2);
            for (int j = 0; j < N / 2; ++j) {
                const long double a = -2.0L * 3.14159265358979323846264338327950288L * j / N;
                tw[j].re = (T)cosl(a);
                tw[j].im = (T)sinl(a);
            }
            Cx<T>* dtw;
            if (dalloc(&dtw, N / 2)) return -1;
            if (upload(dtw, tw)) return -1;
            prm.tw = dtw;
            // scipy.fftpack.fftfreq(N, d) with d = L / (2 pi N) (Burger.py:161): k_n = n * (1 / (N d))
            const double d = cfg.L / (2 * M_PI * N);
            const double val = 1.0 / (N * d);
            std::vector<T> kw(N);
            for (int n = 0; n < N; ++n) kw[n] = (T)((n < (N + 1) / 2 ? n : n - N) * val);
            if (N % 2 == 0) kw[N / 2] = (T)(-(N / 2) * val);
            T* dk;
            if (dalloc(&dk, N)) return -1;
            if (upload(dk, kw)) return -1;
            prm.kwave = dk;
            const int NH = N / 2 + 1;
            Cx<T>*v, *fn;
            float* acc;
            if (dalloc(&v, B * NH) || dalloc(&fn, B * NH) || dalloc(&acc, B * NH)) return -1;
            prm.v = v;
            prm.fn = fn;
            prm.acc = acc;
        }
        T *nu, *tnow, *kprev, *uprev;
        int *iout, *status;
        if (dalloc(&nu, B) || dalloc(&tnow, B) || dalloc(&kprev, B) || dalloc(&iout, B) || dalloc(&status, B)) return -1;
        if (dalloc(&uprev, B * N)) return -1;
        prm.nu = nu;
        prm.tnow = tnow;
        prm.kprev = kprev;
        prm.iout = iout;
        prm.status = status;
        prm.uprev = uprev;
        {
            int* ti;
            T* twt;
            if (dalloc(&ti, 2 * N) || dalloc(&twt, 2 * N)) return -1;
            prm.tap_idx = ti;
            prm.tap_w = twt;
        }
        if (cfg.flags & MPDE_FORCING) {
            Cx<T>* fc;
            if (dalloc(&fc, (size_t)B * cfg.stepper * 3)) return -1;
            prm.fcoef = fc;
        }
        if (cfg.equation == MPDE_KS) {
            T* etd;
            if (dalloc(&etd, (size_t)6 * N)) return -1;
            prm.etd = etd;
        }
        if (cfg.equation == MPDE_ADVECTION) {      // Courant number alpha[B] (Advection.py:43) lives in the etd slot
            T* al;
            if (dalloc(&al, B)) return -1;
            prm.etd = al;
        }
        return 0;
    }

    int set_nu(const double* nu, int64_t n) override {
        if (n != 1 && n != cfg.nenvs) return fail("set_nu: n must be 1 or nenvs");
        std::vector<T> h(cfg.nenvs);
        for (int64_t i = 0; i < cfg.nenvs; ++i) h[i] = (T)nu[n == 1 ? 0 : i];
        if (cfg.equation == MPDE_ADVECTION) {      // alpha = nu dt / dx (Advection.py:43)
            std::vector<T> al(cfg.nenvs);
            for (int64_t i = 0; i < cfg.nenvs; ++i) al[i] = (T)(nu[n == 1 ? 0 : i] * cfg.dt / (cfg.L / cfg.N));
            if (upload(const_cast<T*>(prm.etd), al)) return -1;
        }
        return upload(const_cast<T*>(prm.nu), h);
    }

    int set_basis(int M, const double* basis) override {
        const int N = cfg.N;
        if (M <= 0 || M > 4096) return fail("set_basis: M must be in 1..4096");
        CU(cudaSetDevice(cfg.device));
        CU(cudaDeviceSynchronize());
        if (M != basis_rows) {
            T* bs;
            if (dalloc(&bs, (size_t)M * N)) return -1;      // old table stays owned until destroy
            prm.basis = bs;
            basis_rows = M;
        }
        cfg.M = M;
        prm.M = M;
        std::vector<int> idx(2 * N, 0);
        std::vector<T> w(2 * N, T(0)), dense((size_t)M * N);
        bool sparse = true;
        for (int j = 0; j < N; ++j) {
            int cnt = 0;
            for (int i = 0; i < M; ++i) {
                const double b = basis[(size_t)i * N + j];
                dense[(size_t)i * N + j] = (T)b;
                if (b != 0.0) {
                    if (cnt < 2) {
                        idx[2 * j + cnt] = i;
                        w[2 * j + cnt] = (T)b;
                    }
                    ++cnt;
                }
            }
            if (cnt > 2) sparse = false;
        }
        basis_dense = !sparse;
        if (upload(const_cast<int*>(prm.tap_idx), idx) || upload(const_cast<T*>(prm.tap_w), w) ||
            upload(const_cast<T*>(prm.basis), dense))
            return -1;
        basis_set = true;
        return 0;
    }
    bool basis_dense = false, basis_set = false;
    int basis_rows = 0;

    int set_etd(const double* const tabs[6]) override {
        if (cfg.equation != MPDE_KS) return fail("set_etdrk4: not a KS environment");
        std::vector<T> h((size_t)6 * cfg.N);
        for (int t = 0; t < 6; ++t)
            for (int n = 0; n < cfg.N; ++n) h[(size_t)t * cfg.N + n] = (T)tabs[t][n];
        etd_set = true;
        return upload(const_cast<T*>(prm.etd), h);
    }
    bool etd_set = false;

    int set_forcing(const double* coef, int64_t n) override {
        if (!(cfg.flags & MPDE_FORCING)) return fail("set_forcing: environment was created without MPDE_FORCING");
        if (n != 1 && n != cfg.nenvs) return fail("set_forcing: n must be 1 or nenvs");
        std::vector<Cx<T>> h((size_t)n * cfg.stepper * 3);
        for (size_t i = 0; i < h.size(); ++i) {
            h[i].re = (T)coef[2 * i];
            h[i].im = (T)coef[2 * i + 1];
        }
        n_forcing = (int)(n == 1 ? 1 : 2);
        return upload(const_cast<Cx<T>*>(prm.fcoef), h);
    }

    int set_spectrum_ref(const double* ek, int64_t nref, int64_t rows, const int32_t* map) override {
        // reciprocals of the table for the step kernels' reward (library-owned; set-up path: drain the device on both sides,
        // the caller's table may have been produced on any stream and step kernels may be in flight on others)
        const int64_t n = (ek && nref > 0 && rows > 0) ? nref * rows * (cfg.N / 2) : 0;
        if (n > ek_rcp_cap) {
            Cx<T>* buf = nullptr;
            CU(cudaDeviceSynchronize());
            CU(cudaMalloc(&buf, (size_t)n * sizeof(Cx<T>)));
            if (ek_rcp_buf) {
                owned.erase(std::remove(owned.begin(), owned.end(), static_cast<void*>(ek_rcp_buf)), owned.end());
                cudaFree(ek_rcp_buf);
            }
            owned.push_back(buf);
            ek_rcp_buf = buf;
            ek_rcp_cap = n;
        }
        if (n > 0) {
            CU(cudaDeviceSynchronize());
            launches += launch_rcp_table<T>(ek, ek_rcp_buf, n, nullptr);
            CU(cudaDeviceSynchronize());
        }
        prm.ek_ref = ek;
        prm.ek_pair = n > 0 ? ek_rcp_buf : nullptr;
        prm.ek_rows = rows;
        prm.ek_map = map;
        return 0;
    }
    int set_truth(const void* truth, int64_t ntruth, int64_t rows, const int32_t* map) override {
        (void)ntruth;
        prm.truth = static_cast<const T*>(truth);
        prm.truth_rows = rows;
        prm.truth_map = map;
        return 0;
    }
    int set_history(void* uu, void* vv, double* ektt, int64_t rows) override {
        prm.uu_hist = static_cast<T*>(uu);
        prm.vv_hist = static_cast<Cx<float>*>(vv);
        prm.ektt_hist = ektt;
        prm.hist_rows = (uu || vv || ektt) ? rows : 0;
        return 0;
    }

    int set_peer_output(int n_data, void* const* state, void* const* reward, int64_t parity_stride, void* mc_state,
                        void* mc_reward) override {
        peer_bound = false;
        peer_steps = 0;
        peer_local_state = peer_local_reward = nullptr;
        prm.peer = PeerSink{};
        if (n_data == 0 && parity_stride == 0 && !mc_state && !mc_reward) return 0;
        if (cfg.equation != MPDE_BURGERS || cfg.N > 256)
            return fail("set_peer_output: the fused gather exists for the warp-resident Burgers kernels (N <= 256) only; "
                        "use mpde_peer_put for the other solvers");
        if (n_data < 0 || n_data > MAX_PEERS) return fail("set_peer_output: at most 8 peers");
        if (parity_stride < 0) return fail("set_peer_output: negative parity_stride");
        PeerSink ps;
        ps.n_data = n_data;
        for (int i = 0; i < n_data; ++i) {
            if (!state || !reward || !state[i] || !reward[i]) return fail("set_peer_output: null peer buffer");
            ps.state[i] = state[i];
            ps.reward[i] = reward[i];
        }
        ps.parity_stride = parity_stride;
        ps.row_stores = peer_row_stores;
        if (mc_state && !mc_reward) return fail("set_peer_output: multicast of the state without the reward");
        if (mc_reward && !mc_state && n_data > 0) return fail("set_peer_output: reward-only multicast takes n_data = 0");
        ps.mc_state = mc_state;
        ps.mc_reward = mc_reward;
        prm.peer = ps;
        peer_bound = true;
        return 0;
    }
    // this rank's own slab (copy 0) of the gather buffers: where mpde_step_host finds the rows it copies to the host
    T *peer_local_state = nullptr, *peer_local_reward = nullptr;
    int set_peer_local(void* state, void* reward) override {
        peer_local_state = static_cast<T*>(state);
        peer_local_reward = static_cast<T*>(reward);
        return 0;
    }
    int peer_row_stores = 1;    // PeerSink::row_stores of the next / current binding
    int set_peer_row_stores(int on) override {
        peer_row_stores = on ? 1 : 0;
        prm.peer.row_stores = peer_row_stores;
        return 0;
    }
    bool peer_bound = false;
    int64_t peer_steps = 0;     // steps enqueued since the gather was bound: its low bit selects the buffer copy

    int reset(const void* src, bool spectral_ic, const uint8_t* mask, cudaStream_t st) override {
        CU(cudaSetDevice(cfg.device));
        if (!src) return fail("reset: null initial condition");
        int rc;
        if (spectral()) {
            rc = launch_spectral_aux<T>(prm, cfg.equation, spectral_ic ? AUX_RESET_V : AUX_RESET_U, src, mask, nullptr, st);
        } else {
            if (spectral_ic) return fail("reset_v: finite-difference environments take u0");
            rc = launch_fd_reset<T>(prm, src, mask, st);
        }
        if (rc > 0) { launches += rc; rc = 0; }
        if (rc < 0) return fail("reset: unsupported N for this equation (power of two, 8..2048)");
        CU(cudaGetLastError());
        return 0;
    }

    // staging field of the device-side IC generators: [B,N] complex (hand-off) or real (turbulence)
    Cx<T>* ic_stage = nullptr;
    int ensure_ic_stage() {
        if (!ic_stage && dalloc(&ic_stage, (size_t)cfg.nenvs * cfg.N)) return -1;
        return 0;
    }
    int reset_handoff(const void* vsrc, int64_t nsrc, int nsrc_points, const double* ksrc, const int32_t* src_map,
                      const double* offset, const uint8_t* mask, cudaStream_t st) override {
        CU(cudaSetDevice(cfg.device));
        if (!spectral()) return fail("reset_handoff: spectral solvers only");
        if (!vsrc || !ksrc || nsrc < 1) return fail("reset_handoff: null source");
        if (nsrc_points < cfg.N) return fail("reset_handoff: the source grid must be at least as fine as the environment grid");
        if (ensure_ic_stage()) return -1;
        launches += launch_handoff<T>(vsrc, nsrc_points, ksrc, src_map, offset, mask, ic_stage, cfg.nenvs, cfg.N, st);
        CU(cudaGetLastError());
        return reset(ic_stage, true, mask, st);
    }
    int reset_turbulence(const int64_t* seed, const double* offset, const double* x, const double* amp, const uint8_t* mask,
                         cudaStream_t st) override {
        CU(cudaSetDevice(cfg.device));
        if (!seed || !x || !amp) return fail("reset_turbulence: null argument");
        if (ensure_ic_stage()) return -1;
        const int rc = launch_turbulence<T>(reinterpret_cast<const long long*>(seed), offset, x, amp, mask, ic_stage, cfg.nenvs, cfg.N,
                                            cfg.L, st);
        if (rc < 0) return fail("reset_turbulence: N <= 2048");
        launches += rc;
        CU(cudaGetLastError());
        return reset(ic_stage, false, mask, st);
    }

    int step(const void* actions, int nsub, void* state_out, void* reward_out, cudaStream_t st) override {
        CU(cudaSetDevice(cfg.device));
        if (nsub < 0) return fail("step: nsub < 0");
        SpectralParams<T> p = prm;
        p.nsub = nsub;
        p.actions = static_cast<const T*>(actions);
        p.state_out = static_cast<T*>(state_out);
        p.reward_out = static_cast<T*>(reward_out);
        p.reward_mode = cfg.reward_mode;
        int flags = 0;
        if (cfg.flags & MPDE_DFORCE) flags |= F_DFORCE;
        if (cfg.flags & MPDE_FORCING) flags |= F_FORCING;
        if (cfg.flags & MPDE_SSM) flags |= F_SSM;
        if (cfg.flags & MPDE_DSM) flags |= F_DSM;
        if (cfg.flags & MPDE_FD) flags |= F_FD;
        if (cfg.flags & MPDE_SSMFORCE) flags |= F_SSMFORCE;
        p.A = cfg.num_agents;
        p.M = cfg.M;
        if (actions) {
            if (cfg.M <= 0) return fail("step: actions given but M == 0 (call setup_basis / set M first)");
            if (spectral() && !basis_set) return fail("step: actions given but no basis was set");
            if ((cfg.equation == MPDE_DIFFUSION || cfg.equation == MPDE_DIFFUSION_ERROR) && cfg.M != 1 && cfg.M != cfg.N)
                return fail("step: Diffusion takes 1 or N actions");
            if (cfg.equation == MPDE_LAPLACE && cfg.M != 3 * (cfg.N - 1)) return fail("step: Laplace takes 3 (N - 1) actions");
            if (cfg.equation == MPDE_ADVECTION && cfg.M != 2 && cfg.M != 2 * cfg.N) return fail("step: Advection takes 2 or 2N actions");
            flags |= F_ACTIONS;
            if (basis_dense) flags |= F_BASIS_DENSE;
        }
        if (cfg.flags & MPDE_FORCING) {
            if (n_forcing == 0) return fail("step: forcing enabled but mpde_set_forcing was never called");
            if (n_forcing == 2) flags |= F_FORCING_PER_ENV;
        }
        if (nsub == 0) flags |= F_NO_ADVANCE;
        if (aux_flags & 1) flags |= (1 << 8);       // F_KS_UUROW
        p.flags = flags;
        const bool peer_step = peer_bound && nsub > 0;
        if (peer_bound && nsub == 0) {
            // getState() / getMseReward() of the current state (episode reset, diagnostics): written to the caller's
            // LOCAL buffers only -- no peer stores, no parity flip, nothing is published
            p.peer = PeerSink{};
        } else if (peer_bound) {
            if (!state_out || !reward_out)
                return fail("step: a fused peer gather is bound (mpde_set_peer_output): an advancing call must write state and reward");
            p.peer.parity = (int)(peer_steps & 1);
            static const int row_override = [] { const char* e = std::getenv("MPDE_PEER_ROW_STORES"); return e ? std::atoi(e) : -1; }();
            if (row_override >= 0) p.peer.row_stores = row_override;       // tuning experiments
        }
        if (reward_out && nsub > 0) {
            if (cfg.reward_mode == MPDE_REWARD_SPECTRAL && !p.ek_ref) return fail("step: spectral reward without mpde_set_spectrum_ref");
            if (cfg.reward_mode == MPDE_REWARD_MSE && !p.truth)
                return fail("step: MSE reward without mpde_set_truth");
        }
        int rc;
        switch (cfg.equation) {
            case MPDE_BURGERS: rc = launch_burgers<T>(p, st); break;
            case MPDE_KS:
                if (!etd_set) return fail("step: KS tables missing (mpde_set_etdrk4)");
                rc = launch_ks<T>(p, st);
                break;
            default: rc = launch_fd<T>(p, cfg.equation, (cfg.flags & MPDE_IMPLICIT) != 0, st); break;
        }
        if (rc == -2) return fail("step: the Smagorinsky closures are only available for N <= 256");
        if (rc == -3) return fail("step: the MSE reward is only available for N <= 256");
        if (rc < 0) return fail("step: unsupported N for this equation (power of two, 8..2048)");
        launches += rc;
        CU(cudaGetLastError());
        if (peer_step) ++peer_steps;      // only a step that really launched flips the parity of the gather copies
        return 0;
    }

    // host-buffer variant: stage through library-owned device buffers (allocated on first use)
    // state and reward staging are ONE device allocation [state | reward]: when the caller's host buffers are laid out
    // the same way (reward right behind the state) both travel in a single D2H copy
    bool packed_out = false;      // set by mpde_step_host_packed for the duration of the call
    T *stage_act = nullptr, *stage_out = nullptr;
    size_t stage_act_n = 0, stage_out_n = 0;
    // The chain H2D -> kernel -> D2H of one (actions, nsub, state, reward) signature is captured ONCE into a CUDA
    // graph and replayed by later calls: one driver call per RL step instead of four (MPDE_HOST_GRAPH=0 disables;
    // a stream that is already being captured by the caller, or the legacy default stream, gets the plain chain).
    struct HostGraph {
        const void* actions; void* state; void* reward; int nsub; int64_t epoch; int kernels; cudaGraphExec_t exec;
        int kind;       // 0 host buffers, 1 host buffers with one packed output copy, 2 fused multi-GPU step (device buffers)
        int parity;
    };
    std::vector<HostGraph> host_graphs;
    int step_host_enqueue(const void* actions, int nsub, void* state_out, void* reward_out, size_t na, size_t ns, size_t nr,
                          cudaStream_t st) {
        // with a fused gather bound the kernel writes this rank's rows into its slab of the gather buffer (copy = parity
        // of the step about to be enqueued); the host copy reads them from there
        T* const stage_state = peer_bound ? peer_local_state : stage_out;
        T* const stage_reward = peer_bound ? peer_local_reward : stage_out + ns;
        const int64_t poff = peer_bound ? (peer_steps & 1) * prm.peer.parity_stride : 0;
        if (na) CU(cudaMemcpyAsync(stage_act, actions, na * sizeof(T), cudaMemcpyHostToDevice, st));
        if (step(na ? stage_act : nullptr, nsub, ns ? stage_state : nullptr, nr ? stage_reward : nullptr, st)) return -1;
        if (ns && nr && packed_out && reward_out == static_cast<T*>(state_out) + ns && stage_reward == stage_state + ns) {
            CU(cudaMemcpyAsync(state_out, stage_state + poff, (ns + nr) * sizeof(T), cudaMemcpyDeviceToHost, st));
            return 0;
        }
        if (ns) CU(cudaMemcpyAsync(state_out, stage_state + poff, ns * sizeof(T), cudaMemcpyDeviceToHost, st));
        if (nr) CU(cudaMemcpyAsync(reward_out, stage_reward + poff, nr * sizeof(T), cudaMemcpyDeviceToHost, st));
        return 0;
    }
    int step_host(const void* actions, int nsub, void* state_out, void* reward_out, cudaStream_t st) override {
        CU(cudaSetDevice(cfg.device));
        if (peer_bound && (!peer_local_state || !peer_local_reward))
            return fail("step_host: a fused peer gather is bound; name this rank's own slab with mpde_set_peer_local first");
        const int parity = peer_bound ? (int)(peer_steps & 1) : 0;
        const size_t B = (size_t)cfg.nenvs;
        const size_t na = actions ? B * (size_t)cfg.M : 0, ns = state_out ? B * (size_t)state_size() : 0;
        const size_t nr = reward_out ? B * (size_t)(cfg.reward_mode == MPDE_REWARD_DIRECT ? (cfg.equation == MPDE_LAPLACE ? cfg.N - 1 : cfg.N)
                                                                                             : cfg.num_agents) : 0;
        if (na > stage_act_n) { if (dalloc(&stage_act, na)) return -1; stage_act_n = na; ++epoch; }
        if (ns + nr > stage_out_n) { if (dalloc(&stage_out, ns + nr + 2)) return -1; stage_out_n = ns + nr; ++epoch; }
        static const bool use_graph = [] { const char* s = std::getenv("MPDE_HOST_GRAPH"); return !(s && s[0] == '0'); }();
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (st) CU(cudaStreamIsCapturing(st, &cs));
        if (!use_graph || !st || st == cudaStreamLegacy || cs != cudaStreamCaptureStatusNone)
            return step_host_enqueue(actions, nsub, state_out, reward_out, na, ns, nr, st);
        {
            int rc = 0;
            if (find_and_launch(HostGraph{actions, state_out, reward_out, nsub, epoch, 0, nullptr, packed_out ? 1 : 0, parity}, st, rc)) return rc;
        }
        return capture_and_launch(HostGraph{actions, state_out, reward_out, nsub, epoch, 0, nullptr, packed_out ? 1 : 0, parity}, st,
                                  [&] { return step_host_enqueue(actions, nsub, state_out, reward_out, na, ns, nr, st); });
    }
    // Capture what `enqueue` puts on `st` into a graph, cache it under `key` and launch it once.  The capture pass only
    // records: the parity counter of a bound gather moves when the graph is LAUNCHED, and is rolled back on any failure.
    template <typename F>
    int capture_and_launch(HostGraph key, cudaStream_t st, F enqueue) {
        const int64_t l0 = launches, ps0 = peer_steps;
        CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue();
        const std::string first_err = g_err;
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(st, &graph);
        key.kernels = (int)(launches - l0);
        launches = l0;
        peer_steps = ps0;
        if (rc != 0 || ce != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            return rc != 0 ? fail(first_err) : fail(std::string("graph capture failed: ") + cudaGetErrorString(ce));
        }
        const cudaError_t ie = cudaGraphInstantiate(&key.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) return fail(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie));
        if (host_graphs.size() >= 8) {
            cudaGraphExecDestroy(host_graphs.front().exec);
            host_graphs.erase(host_graphs.begin());
        }
        host_graphs.push_back(key);
        CU(cudaGraphLaunch(key.exec, st));
        launches += key.kernels;
        if (peer_bound && key.nsub > 0) ++peer_steps;
        return 0;
    }
    bool find_and_launch(const HostGraph& key, cudaStream_t st, int& rc) {
        for (const HostGraph& g : host_graphs)
            if (g.actions == key.actions && g.state == key.state && g.reward == key.reward && g.nsub == key.nsub && g.epoch == epoch &&
                g.kind == key.kind && g.parity == key.parity) {
                const cudaError_t e = cudaGraphLaunch(g.exec, st);
                if (e != cudaSuccess) { rc = fail(std::string("cudaGraphLaunch: ") + cudaGetErrorString(e)); return true; }
                if (peer_bound && g.nsub > 0) ++peer_steps;
                launches += g.kernels;
                rc = 0;
                return true;
            }
        return false;
    }

    // ---- fused multi-GPU step: kernel (+ peer stores) -> publish -> wait, ONE host call ---------------------------------
    struct PeerSync {
        void* slots[16] = {};
        int n = 0, nranks = 0;
        void *step_dev = nullptr, *expect_dev = nullptr, *err = nullptr;
        const void* my_flags = nullptr;
        int64_t timeout_us = 0;
        bool set = false;
        cudaStream_t side = nullptr;
        cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
        bool pending_join = false;
    } sync;
    int set_peer_sync(void* const* flag_slots, int n, void* step_dev, const void* my_flags, int nranks, void* expect_dev, void* err,
                      int64_t timeout_us) override {
        if (!flag_slots && n == 0) { sync.set = false; return 0; }
        if (n < 1 || n > 16 || nranks < 1 || nranks > 16 || !flag_slots || !step_dev || !my_flags || !expect_dev || !err)
            return fail("set_peer_sync: 1..16 flag slots / ranks, counters and an error flag");
        CU(cudaSetDevice(cfg.device));
        for (int i = 0; i < n; ++i) sync.slots[i] = flag_slots[i];
        sync.n = n; sync.nranks = nranks; sync.step_dev = step_dev; sync.my_flags = my_flags; sync.expect_dev = expect_dev;
        sync.err = err; sync.timeout_us = timeout_us;
        if (!sync.side) {
            CU(cudaStreamCreateWithFlags(&sync.side, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&sync.ev_fork, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&sync.ev_join, cudaEventDisableTiming));
        }
        sync.set = true;
        return 0;
    }
    int exchange(cudaStream_t st) {
        if (mpde_peer_exchange_next(sync.slots, sync.n, sync.step_dev, sync.my_flags, sync.nranks, sync.expect_dev, sync.err,
                                    sync.timeout_us, st))
            return fail(std::string("step_fused: ") + mpde_peer_last_error());
        launches += 1;
        return 0;
    }
    int step_fused(const void* actions, int nsub, void* state_out, void* reward_out, int async, cudaStream_t st) override {
        CU(cudaSetDevice(cfg.device));
        if (!peer_bound || !sync.set) return fail("step_fused: bind the gather first (mpde_set_peer_output + mpde_set_peer_sync)");
        if (nsub <= 0) return fail("step_fused: nsub must be positive (use mpde_step for getState)");
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (st) CU(cudaStreamIsCapturing(st, &cs));
        static const bool use_graph = [] { const char* s = std::getenv("MPDE_HOST_GRAPH"); return !(s && s[0] == '0'); }();
        if (async) {
            // publish + wait on the library's side stream, forked behind the step kernel: the caller's stream is free for the
            // next (independent) batch; mpde_peer_join orders a consumer after the gather.  Capturable by the caller.
            if (step(actions, nsub, state_out, reward_out, st)) return -1;
            CU(cudaEventRecord(sync.ev_fork, st));
            CU(cudaStreamWaitEvent(sync.side, sync.ev_fork, 0));
            if (exchange(sync.side)) return -1;
            CU(cudaEventRecord(sync.ev_join, sync.side));
            sync.pending_join = true;
            return 0;
        }
        if (!use_graph || !st || st == cudaStreamLegacy || cs != cudaStreamCaptureStatusNone) {
            if (step(actions, nsub, state_out, reward_out, st)) return -1;
            return exchange(st);
        }
        const HostGraph key{actions, state_out, reward_out, nsub, epoch, 0, nullptr, 2, (int)(peer_steps & 1)};
        int rc = 0;
        if (find_and_launch(key, st, rc)) return rc;
        return capture_and_launch(key, st, [&] { return step(actions, nsub, state_out, reward_out, st) ? -1 : exchange(st); });
    }
    int peer_join(cudaStream_t st) override {
        if (sync.pending_join) {
            CU(cudaStreamWaitEvent(st, sync.ev_join, 0));
            sync.pending_join = false;
        }
        return 0;
    }

    int step_host_packed(const void* actions, int nsub, void* out, cudaStream_t st) override {
        if (!out) return fail("step_host_packed: null output buffer");
        const size_t ns = (size_t)cfg.nenvs * (size_t)state_size();
        const bool has_reward = cfg.reward_mode != MPDE_REWARD_NONE && nsub > 0;
        packed_out = true;
        const int rc = step_host(actions, nsub, out, has_reward ? static_cast<T*>(out) + ns : nullptr, st);
        packed_out = false;
        return rc;
    }

    int compute_sgs(int nURG, int64_t rows, void* sgs, void* alt, void* alt2, cudaStream_t st) override {
        CU(cudaSetDevice(cfg.device));
        if (!spectral()) return fail("compute_sgs: spectral solvers only");
        if (!prm.uu_hist || prm.hist_rows < 2) return fail("compute_sgs: needs the uu history (mpde_set_history) with at least two rows");
        if (rows < 2 || rows > prm.hist_rows) return fail("compute_sgs: rows must be in 2..hist_rows");
        if (nURG < 2 || nURG > 64 || nURG > cfg.N) return fail("compute_sgs: nURG must be in 2..64");
        if (!sgs) return fail("compute_sgs: null output");
        if (rows != prm.hist_rows) return fail("compute_sgs: rows must equal the history length (the reference scans uu.shape[0] rows)");
        const int rc = launch_sgs<T>(prm, prm.uu_hist, rows, nURG, cfg.equation == MPDE_KS ? 1 : 0, sgs, alt, alt2, st);
        if (rc < 0) return fail("compute_sgs: N must be 256, 512, 1024 or 2048");
        launches += rc;
        CU(cudaGetLastError());
        return 0;
    }

    int get(int field, void* dst, cudaStream_t st) override {
        CU(cudaSetDevice(cfg.device));
        const int64_t B = cfg.nenvs;
        const int N = cfg.N, NH = N / 2 + 1;
        auto copy = [&](const void* src, size_t bytes) -> int {
            CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
            return 0;
        };
        switch (field) {
            case MPDE_FIELD_U:
                if (spectral()) {
                    int rc = launch_spectral_aux<T>(prm, cfg.equation, AUX_GET_U, nullptr, nullptr, dst, st);
                    if (rc < 0) return fail("get(U): unsupported N");
                    launches += rc;
                    CU(cudaGetLastError());
                    return 0;
                }
                return copy(prm.uprev, sizeof(T) * B * N);      // FD solvers keep u in the uprev slot
            case MPDE_FIELD_V:
            case MPDE_FIELD_FN_OLD: {
                if (!spectral()) return fail("get: field only exists for spectral solvers");
                const Cx<T>* half = field == MPDE_FIELD_V ? prm.v : prm.fn;
                const int64_t n = B * N;
                unpack_half_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(half, static_cast<Cx<T>*>(dst), B, N);
                launches += 1;
                CU(cudaGetLastError());
                return 0;
            }
            case MPDE_FIELD_U_PREV: return copy(prm.uprev, sizeof(T) * B * N);
            case MPDE_FIELD_EK_SUM:
                if (!spectral()) return fail("get: field only exists for spectral solvers");
                return copy(prm.acc, sizeof(float) * B * NH);
            case MPDE_FIELD_IOUTNUM: return copy(prm.iout, sizeof(int) * B);
            case MPDE_FIELD_T: return copy(prm.tnow, sizeof(T) * B);
            case MPDE_FIELD_KPREV: return copy(prm.kprev, sizeof(T) * B);
            case MPDE_FIELD_STATUS: return copy(prm.status, sizeof(int) * B);
            case MPDE_FIELD_K:
                if (!spectral()) return fail("get: field only exists for spectral solvers");
                return copy(prm.kwave, sizeof(T) * N);
            case MPDE_FIELD_NU: return copy(prm.nu, sizeof(T) * B);
            case MPDE_FIELD_ALPHA:
                if (cfg.equation != MPDE_ADVECTION) return fail("get: field only exists for Advection");
                return copy(prm.etd, sizeof(T) * B);
        }
        return fail("get: unknown field");
    }

    int set(int field, const void* src, cudaStream_t st) override {
        CU(cudaSetDevice(cfg.device));
        const int64_t B = cfg.nenvs;
        const int N = cfg.N, NH = N / 2 + 1;
        auto copy = [&](void* dst, size_t bytes) -> int {
            CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
            return 0;
        };
        switch (field) {
            case MPDE_FIELD_V:
            case MPDE_FIELD_FN_OLD: {
                if (!spectral()) return fail("set: field only exists for spectral solvers");
                Cx<T>* half = field == MPDE_FIELD_V ? prm.v : prm.fn;
                const int64_t n = B * NH;
                pack_half_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(static_cast<const Cx<T>*>(src), half, B, N);
                launches += 1;
                CU(cudaGetLastError());
                return 0;
            }
            case MPDE_FIELD_U:
                if (spectral()) return fail("set(U): spectral solvers derive u from v; use mpde_reset_u or set V");
                return copy(prm.uprev, sizeof(T) * B * N);
            case MPDE_FIELD_U_PREV: return copy(prm.uprev, sizeof(T) * B * N);
            case MPDE_FIELD_EK_SUM: return copy(prm.acc, sizeof(float) * B * NH);
            case MPDE_FIELD_IOUTNUM: return copy(prm.iout, sizeof(int) * B);
            case MPDE_FIELD_T: return copy(prm.tnow, sizeof(T) * B);
            case MPDE_FIELD_KPREV: return copy(prm.kprev, sizeof(T) * B);
            case MPDE_FIELD_STATUS: return copy(prm.status, sizeof(int) * B);
            case MPDE_FIELD_NU: return copy(const_cast<T*>(prm.nu), sizeof(T) * B);
            case MPDE_FIELD_ALPHA:
                if (cfg.equation != MPDE_ADVECTION) return fail("set: field only exists for Advection");
                return copy(const_cast<T*>(prm.etd), sizeof(T) * B);
        }
        return fail("set: field is read-only or unknown");
    }
};

extern "C" {

int mpde_create(const mpde_config* cfg, mpde_env** out) {
    if (!cfg || !out) return fail("create: null argument");
    if (cfg->struct_size != (int32_t)sizeof(mpde_config)) return fail("create: mpde_config size mismatch (ABI)");
    if (cfg->nenvs <= 0) return fail("create: nenvs must be positive");
    if (cfg->N < 4) return fail("create: N must be >= 4");
    if (cfg->equation < MPDE_BURGERS || cfg->equation > MPDE_LAPLACE) return fail("create: unknown equation");
    if ((cfg->equation == MPDE_BURGERS || cfg->equation == MPDE_KS) && cfg->N < 8) return fail("create: spectral solvers need N >= 8");
    const bool spectral = cfg->equation == MPDE_BURGERS || cfg->equation == MPDE_KS;
    if (spectral && !is_pow2(cfg->N)) return fail("create: spectral solvers need a power-of-two N");
    if (spectral && cfg->N > 2048) return fail("create: N > 2048 not supported");
    if (cfg->num_agents < 1 || cfg->N % cfg->num_agents) return fail("create: num_agents must divide N");
    if (cfg->stepper < 1) return fail("create: stepper must be >= 1");
    if ((cfg->flags & MPDE_SSM) && (cfg->flags & MPDE_DSM)) return fail("create: ssm and dsm are exclusive (Burger.py:50)");
    if (cfg->flags & MPDE_FD) {
        if (cfg->equation != MPDE_BURGERS) return fail("create: MPDE_FD is a Burgers option (Burger_fd)");
        if (cfg->N > 256) return fail("create: Burger_fd is available for N <= 256");
        if (cfg->version > 2) return fail("create: Burger_fd has state versions 0, 1 and 2 (Burger_fd.py:590-640)");
    }
    if (cfg->version < 0 || cfg->version > 4) return fail("create: version must be 0..4");
    if (!(cfg->L > 0) || !(cfg->dt > 0)) return fail("create: L and dt must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail("create: no CUDA device -- marlpde_b200 has no CPU path");
    if (cfg->device < 0 || cfg->device >= ndev) return fail("create: bad device ordinal");
    mpde_env* e = nullptr;
    if (cfg->dtype == MPDE_F64) e = new Env<double>();
    else if (cfg->dtype == MPDE_F32) e = new Env<float>();
    else return fail("create: unknown dtype");
    e->cfg = *cfg;
    if (e->init() != 0) {
        delete e;
        return -1;
    }
    *out = e;
    return 0;
}

int mpde_destroy(mpde_env* env) {
    delete env;
    return 0;
}

int64_t mpde_state_size(const mpde_env* env) { return env ? env->state_size() : -1; }
int mpde_set_nu(mpde_env* env, const double* nu, int64_t n) {
    if (env) ++env->epoch; return env && nu ? env->set_nu(nu, n) : fail("null argument"); }
int mpde_set_basis(mpde_env* env, int32_t M, const double* b) {
    if (env) ++env->epoch; return env && b ? env->set_basis(M, b) : fail("null argument"); }
int mpde_set_reward_mode(mpde_env* env, int32_t mode) {
    if (env) ++env->epoch;
    if (!env) return fail("null argument");
    if (mode < 0 || mode > 3) return fail("set_reward_mode: unknown mode");
    env->cfg.reward_mode = mode;
    return 0;
}
int mpde_set_etdrk4(mpde_env* env, const double* E, const double* E2, const double* Q, const double* f1,
                    const double* f2, const double* f3) {
    if (env) ++env->epoch;
    if (!env || !E || !E2 || !Q || !f1 || !f2 || !f3) return fail("null argument");
    const double* tabs[6] = {E, E2, Q, f1, f2, f3};
    return env->set_etd(tabs);
}
int mpde_set_forcing(mpde_env* env, const double* c, int64_t n) {
    if (env) ++env->epoch; return env && c ? env->set_forcing(c, n) : fail("null argument"); }
int mpde_set_spectrum_ref(mpde_env* env, const double* ek, int64_t nref, int64_t rows, const int32_t* map) {
    if (env) ++env->epoch;
    return env && ek ? env->set_spectrum_ref(ek, nref, rows, map) : fail("null argument");
}
int mpde_set_truth(mpde_env* env, const void* t, int64_t nt, int64_t rows, const int32_t* map) {
    if (env) ++env->epoch;
    return env && t ? env->set_truth(t, nt, rows, map) : fail("null argument");
}
int mpde_set_history(mpde_env* env, void* uu, void* vv, double* ektt, int64_t rows) {
    if (env) ++env->epoch;
    return env ? env->set_history(uu, vv, ektt, rows) : fail("null argument");
}
int mpde_reset_u(mpde_env* env, const void* u0, const uint8_t* mask, void* stream) {
    return env ? env->reset(u0, false, mask, static_cast<cudaStream_t>(stream)) : fail("null argument");
}
int mpde_reset_v(mpde_env* env, const void* v0, const uint8_t* mask, void* stream) {
    return env ? env->reset(v0, true, mask, static_cast<cudaStream_t>(stream)) : fail("null argument");
}
int mpde_reset_handoff(mpde_env* env, const void* vsrc_dev, int64_t nsrc, int32_t nsrc_points, const double* ksrc_dev,
                       const int32_t* src_map_dev, const double* offset_dev, const uint8_t* mask_dev, void* stream) {
    return env ? env->reset_handoff(vsrc_dev, nsrc, nsrc_points, ksrc_dev, src_map_dev, offset_dev, mask_dev, static_cast<cudaStream_t>(stream))
               : fail("null argument");
}
int mpde_reset_turbulence(mpde_env* env, const int64_t* seed_dev, const double* offset_dev, const double* x_dev, const double* amp_dev,
                          const uint8_t* mask_dev, void* stream) {
    return env ? env->reset_turbulence(seed_dev, offset_dev, x_dev, amp_dev, mask_dev, static_cast<cudaStream_t>(stream))
               : fail("null argument");
}
int mpde_step(mpde_env* env, const void* actions, int32_t nsub, void* state_out, void* reward_out, void* stream) {
    return env ? env->step(actions, nsub, state_out, reward_out, static_cast<cudaStream_t>(stream)) : fail("null argument");
}
int mpde_step_host(mpde_env* env, const void* actions, int32_t nsub, void* state_out, void* reward_out, void* stream) {
    return env ? env->step_host(actions, nsub, state_out, reward_out, static_cast<cudaStream_t>(stream)) : fail("null argument");
}
int mpde_step_host_packed(mpde_env* env, const void* actions, int32_t nsub, void* out_host, void* stream) {
    return env ? env->step_host_packed(actions, nsub, out_host, static_cast<cudaStream_t>(stream)) : fail("null argument");
}
int mpde_set_peer_output(mpde_env* env, int32_t n_data, void* const* state_ptrs, void* const* reward_ptrs, int64_t parity_stride,
                         void* mc_state, void* mc_reward) {
    if (env) ++env->epoch;
    return env ? env->set_peer_output(n_data, state_ptrs, reward_ptrs, parity_stride, mc_state, mc_reward) : fail("null argument");
}
int mpde_set_peer_local(mpde_env* env, void* local_state, void* local_reward) {
    if (env) ++env->epoch;
    return env ? env->set_peer_local(local_state, local_reward) : fail("null argument");
}
int mpde_set_peer_row_stores(mpde_env* env, int32_t on) {
    if (env) ++env->epoch;
    return env ? env->set_peer_row_stores(on) : fail("null argument");
}
int mpde_set_peer_sync(mpde_env* env, void* const* flag_ptrs, int32_t n, void* step_dev, const void* my_flags_dev, int32_t nranks,
                       void* expect_dev, void* err, int64_t timeout_us) {
    if (env) ++env->epoch;
    return env ? env->set_peer_sync(flag_ptrs, n, step_dev, my_flags_dev, nranks, expect_dev, err, timeout_us) : fail("null argument");
}
int mpde_step_fused(mpde_env* env, const void* actions, int32_t nsub, void* state_out, void* reward_out, int32_t async_gather,
                    void* stream) {
    return env ? env->step_fused(actions, nsub, state_out, reward_out, async_gather, static_cast<cudaStream_t>(stream)) : fail("null argument");
}
int mpde_peer_join(mpde_env* env, void* stream) { return env ? env->peer_join(static_cast<cudaStream_t>(stream)) : fail("null argument"); }
int mpde_compute_sgs(mpde_env* env, int32_t nURG, int64_t rows, void* sgs_out, void* alt_out, void* alt2_out, void* stream) {
    return env ? env->compute_sgs(nURG, rows, sgs_out, alt_out, alt2_out, static_cast<cudaStream_t>(stream)) : fail("null argument");
}
int mpde_get(mpde_env* env, int32_t field, void* dst, void* stream) {
    return env && dst ? env->get(field, dst, static_cast<cudaStream_t>(stream)) : fail("null argument");
}
int mpde_set(mpde_env* env, int32_t field, const void* src, void* stream) {
    return env && src ? env->set(field, src, static_cast<cudaStream_t>(stream)) : fail("null argument");
}
int mpde_set_option(mpde_env* env, int32_t key, int64_t value) {
    if (env) ++env->epoch;
    if (!env) return fail("null argument");
    if (key == MPDE_OPT_KS_UUROW) {
        env->aux_flags = (env->aux_flags & ~1) | (value ? 1 : 0);
        return 0;
    }
    if (key == MPDE_OPT_NUM_AGENTS) {
        if (value < 1 || env->cfg.N % value) return fail("set_option: num_agents must divide N");
        env->cfg.num_agents = (int32_t)value;
        return 0;
    }
    if (key == MPDE_OPT_NUM_ACTIONS) {
        if (env->cfg.equation == MPDE_BURGERS || env->cfg.equation == MPDE_KS)
            return fail("set_option: spectral solvers set M through mpde_set_basis");
        env->cfg.M = (int32_t)value;
        return 0;
    }
    return fail("set_option: unknown key");
}
int64_t mpde_launch_count(const mpde_env* env) { return env ? env->launches : -1; }
const char* mpde_last_error(void) { return g_err.c_str(); }
int mpde_abi_version(void) { return MPDE_ABI_VERSION; }

}  // extern "C"
