/*
 * marlpde_b200 -- C ABI of the B200-native marlpde environment time-stepper.
 *
 * The reference (wadaniel/marlpde) has no FFI: its hot path is the duck-typed Python
 * class API of python/_model/{Burger,KS,Diffusion,Advection}.py, driven one environment
 * and one step per call by python/_model/*_environment.py.  This header is the boundary a
 * maintainer binds instead (ctypes stub in INTEGRATION.md); every entry point names the
 * reference method(s) it replaces for a BATCH of B independent environments.
 *
 * Conventions
 *   - plain C, no torch types: device buffers are raw CUDA device pointers, `stream` is a
 *     cudaStream_t passed as void* (NULL = default stream);
 *   - all per-call buffers are CALLER-owned device memory, row-major with the environment
 *     index first ([B, ...]); the library owns only the persistent solver state allocated
 *     in mpde_create and freed in mpde_destroy;
 *   - element type of real buffers is double (dtype MPDE_F64) or float (MPDE_F32); complex
 *     buffers are interleaved (re, im) pairs of that type;
 *   - every call is asynchronous on `stream`, does no allocation and no synchronisation
 *     (CUDA-graph capturable), except create/destroy/set_* which may synchronise;
 *   - return value 0 = success, negative = error (message via mpde_last_error()).
 *     A numerical blow-up is NOT an error: it sets status[e] = MPDE_TRUNCATED and freezes
 *     environment e (reference: np.seterr(over='raise') + the try/except of
 *     burger_environment.py:158-167,198-201).
 */
#ifndef MARLPDE_B200_H
#define MARLPDE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPDE_ABI_VERSION 1

enum mpde_equation {
    MPDE_BURGERS = 0, MPDE_KS = 1, MPDE_DIFFUSION = 2, MPDE_ADVECTION = 3,
    MPDE_DIFFUSION_ERROR = 4, /* DiffusionError.py:160-216: the action is the ERROR of the Laplacian stencil            */
    MPDE_LAPLACE = 5          /* Laplace.py:116-166: N grid points incl. the Dirichlet point 0 (the class's N + 1), 3 free */
                              /* stencil entries per agent (M = 3 (N - 1)), forcing row via mpde_set_truth ([.,1,N]),    */
                              /* state [u_{i-1}, u_i, u_{i+1}, force_i] (4 (N - 1)), reward MPDE_REWARD_DIRECT (N - 1)    */
};
enum mpde_dtype { MPDE_F64 = 0, MPDE_F32 = 1 };
enum mpde_status { MPDE_RUNNING = 0, MPDE_TRUNCATED = 1 };
enum mpde_reward { MPDE_REWARD_NONE = 0, MPDE_REWARD_SPECTRAL = 1, MPDE_REWARD_MSE = 2, MPDE_REWARD_DIRECT = 3 };

/* configuration flags (constructor keywords of the reference classes) */
#define MPDE_DFORCE   (1 << 0) /* Burger/KS(dforce=True): actions are a direct forcing          */
#define MPDE_FORCING  (1 << 1) /* Burger(forcing=True): 3-mode stochastic forcing                */
#define MPDE_SSM      (1 << 2) /* Burger(ssm=True): static Smagorinsky closure                   */
#define MPDE_DSM      (1 << 3) /* Burger(dsm=True): dynamic Smagorinsky closure                  */
#define MPDE_IMPLICIT (1 << 4) /* Diffusion(implicit=True): implicit-Euler FDstep                */
#define MPDE_FD       (1 << 5) /* Burger_fd: explicit Euler + finite differences (Burger_fd.py:335-476);  */
                               /* warp-resident kernels only (N <= 256), state versions 0, 1 and 2       */
#define MPDE_SSMFORCE (1 << 6) /* Burger_fd(ssmforce=True): actions are Smagorinsky coefficients          */

/* fields for mpde_get / mpde_set */
enum mpde_field {
    MPDE_FIELD_U = 0,       /* [B,N] real      current field  (Burger.u, KS: Re ifft(v))                 */
    MPDE_FIELD_V = 1,       /* [B,N] complex   current spectrum, FFT order (Burger.v / KS.v)             */
    MPDE_FIELD_FN_OLD = 2,  /* [B,N] complex   Burger.Fn_old                                             */
    MPDE_FIELD_U_PREV = 3,  /* [B,N] real      uu[ioutnum-1]                                             */
    MPDE_FIELD_EK_SUM = 4,  /* [B,N/2+1] float32  running sum of Ek_kt rows (Ek_ktt * (ioutnum+1))       */
    MPDE_FIELD_IOUTNUM = 5, /* [B] int32                                                                 */
    MPDE_FIELD_T = 6,       /* [B] real        accumulated time                                          */
    MPDE_FIELD_KPREV = 7,   /* [B] real        kPrevRelErr of the spectral reward                        */
    MPDE_FIELD_STATUS = 8,  /* [B] int32                                                                 */
    MPDE_FIELD_K = 9,       /* [N] real        wavenumber table (Burger.k)                               */
    MPDE_FIELD_NU = 10,     /* [B] real                                                                  */
    MPDE_FIELD_ALPHA = 11   /* [B] real        Advection Courant number nu*dt/dx (Advection.py:43)        */
};

typedef struct mpde_config {
    int32_t struct_size;  /* = sizeof(mpde_config)                                                 */
    int32_t equation;     /* enum mpde_equation                                                    */
    int32_t dtype;        /* enum mpde_dtype                                                       */
    int32_t device;       /* CUDA device ordinal                                                   */
    int64_t nenvs;        /* B                                                                     */
    int32_t N;            /* grid points (power of two for the spectral solvers)                   */
    int32_t M;            /* actions per environment (0 = none yet; mpde_set_basis may change it)  */
    int32_t num_agents;   /* A (numAgents); must divide N                                          */
    int32_t version;      /* Burger state version 0..4 (Burger.py:617-626)                         */
    int32_t stepper;      /* s: forcing column period (Burger.py:416)                              */
    int32_t flags;        /* MPDE_DFORCE | MPDE_FORCING | ...                                      */
    int32_t reward_mode;  /* enum mpde_reward                                                      */
    int32_t team_lanes;   /* 0 = default (N/4 lanes per environment); 4 / 8 / 16 / 32 selects another team size of the
                           * warp-resident Burgers kernels (N = 32: 4 wins from ~8192 environments per GPU).  Variants
                           * differ in the last bits: keep it fixed for runs that must agree bitwise.             */
    double L;             /* domain length                                                         */
    double dt;            /* time step                                                             */
} mpde_config;

typedef struct mpde_env mpde_env;

/* Burger.__init__/__setup_fourier (Burger.py:24-175), KS.__init__ (KS.py:33-137),
 * Diffusion/Advection.__init__: allocate the batched solver state and constant tables. */
int mpde_create(const mpde_config* cfg, mpde_env** out);
int mpde_destroy(mpde_env* env);

/* number of reals per environment written by mpde_step(state_out): len(getState()) */
int64_t mpde_state_size(const mpde_env* env);

/* viscosity per environment (Burger.nu incl. nunoise, Burger.py:87-89); host array of n = 1 or B */
int mpde_set_nu(mpde_env* env, const double* nu_host, int64_t n);

/* setup_basis (Burger.py:177-203 / KS.py:139-164): HOST row-major [M,N] basis matrix.  May be
 * called at any time (the reference sets the basis after construction); synchronises. */
int mpde_set_basis(mpde_env* env, int32_t M, const double* basis_host);

/* which reward mpde_step(reward_out) evaluates (enum mpde_reward) */
int mpde_set_reward_mode(mpde_env* env, int32_t mode);

/* small per-handle switches.  MPDE_OPT_KS_UUROW (value 0/1): KS(dforce=False).step reads the float32
 * history row uu[ioutnum], which only fou2real()/getState() refresh (KS.py:241, 316-320); 1 = the row is
 * current for the FIRST solver step of the next mpde_step call, later rows are zero as in the reference. */
#define MPDE_OPT_KS_UUROW 1
/* MPDE_OPT_NUM_AGENTS: numAgents of the NEXT calls (Diffusion/Advection pass it per call: step(actions, numAgents),
 * getState(numAgents), getMseReward(numAgents)).  MPDE_OPT_NUM_ACTIONS: actions per environment of the FD solvers
 * (Diffusion: 1 or N, Advection: 2 or 2N). */
#define MPDE_OPT_NUM_AGENTS 2
#define MPDE_OPT_NUM_ACTIONS 3
int mpde_set_option(mpde_env* env, int32_t key, int64_t value);

/* KS.__setup_etdrk4 tables (KS.py:127-137): HOST arrays of N doubles each, FFT order */
int mpde_set_etdrk4(mpde_env* env, const double* E, const double* E2, const double* Q,
                    const double* f1, const double* f2, const double* f3);

/* Stochastic forcing (Burger.py:410-421): HOST complex table [n, stepper, 3] (n = 1 or B) holding
 * fft(forcing)[k] for k = 1,2,3 and each column c = ioutnum % stepper:
 *   r1[k,c] * sqrt(2)/L / sqrt(k*s*dt) * (N/2) * exp(i*(2 pi k offset / L + 2 pi r2[k,c])) */
int mpde_set_forcing(mpde_env* env, const double* coef_host, int64_t n);

/* Spectral-reward reference (burger_environment.py:174): DEVICE double [nref, rows, N/2] rows of
 * dns.Ek_ktt[:, :N/2]; env_map DEVICE int32 [B] or NULL (all environments use reference 0).
 * Set-up call: synchronises the device and builds a library-owned table of (value, reciprocal) pairs from ek_dev as it is
 * NOW (the step kernels read that table; call again after changing ek_dev).  ek_dev itself must stay valid as well (the
 * KS and CTA-resident kernels read it directly). */
int mpde_set_spectrum_ref(mpde_env* env, const double* ek_dev, int64_t nref, int64_t rows, const int32_t* env_map_dev);

/* setGroundTruth + getMseReward (Burger.py:322-323, 578-601): DEVICE real [ntruth, rows, N] table of
 * the truth interpolated at the (shifted) grid of the environments, one row per solver step. */
int mpde_set_truth(mpde_env* env, const void* truth_dev, int64_t ntruth, int64_t rows, const int32_t* env_map_dev);

/* Optional per-step history (Burger.py:151-152, 497-498, 555): caller-owned DEVICE buffers, each may
 * be NULL: uu [B, rows, N] real, vv [B, rows, N] complex64, ektt [B, rows, N/2+1] double
 * (= Ek_ktt[:, :N/2+1], the running time-average of the float32 spectrum rows).
 * Row i is written when an environment reaches ioutnum == i (row 0 by reset). */
int mpde_set_history(mpde_env* env, void* uu_dev, void* vv_dev, double* ektt_dev, int64_t rows);

/* IC(u0=...) / IC(v0=...) (Burger.py:205-320, KS.py:166-219): DEVICE [B,N] real or complex.
 * mask_dev: DEVICE uint8 [B] (1 = reset this environment) or NULL = all. */
int mpde_reset_u(mpde_env* env, const void* u0_dev, const uint8_t* mask_dev, void* stream);
int mpde_reset_v(mpde_env* env, const void* v0_dev, const uint8_t* mask_dev, void* stream);

/* Episode reset without a host loop.
 * mpde_reset_handoff: the DNS -> LES spectral hand-off of burger_environment.py:109-112 / ks_environment.py:52-54 for
 *   every (masked) environment e:  v0_e = concat(w[:(N+1)//2], w[-(N-1)//2:]) * N / nsrc_points,
 *   w = vsrc[src_map[e]] * exp(1j * 2 pi * offset[e] * ksrc), followed by IC(v0 = v0_e).
 *   vsrc_dev: DEVICE complex128 [nsrc, nsrc_points] (dns.v0), ksrc_dev: DEVICE double [nsrc_points] (dns.k),
 *   src_map_dev int32 [B] or NULL (all use source 0), offset_dev double [B] or NULL (no shift).
 * mpde_reset_turbulence: IC(case='turbulence') (Burger.py:227-260) for every (masked) environment from its own seed and
 *   offset: seed_dev int64 [B], offset_dev double [B] or NULL, x_dev double [N] (Burger.x), amp_dev double [N] with
 *   amp[k] = sqrt(2 E_k), E_k = 5^(-5/3) for k <= 5 else k^(-5/3). */
int mpde_reset_handoff(mpde_env* env, const void* vsrc_dev, int64_t nsrc, int32_t nsrc_points, const double* ksrc_dev,
                       const int32_t* src_map_dev, const double* offset_dev, const uint8_t* mask_dev, void* stream);
int mpde_reset_turbulence(mpde_env* env, const int64_t* seed_dev, const double* offset_dev, const double* x_dev,
                          const double* amp_dev, const uint8_t* mask_dev, void* stream);

/* Stochastic-forcing tables on the device (Burger.py:66, 88-89, 94-95): for each of the n seeds (DEVICE int64,
 * values in [0, 2^32): NumPy's init_genrand seeding) reproduce `np.random.seed(seed)`; [nunoise: nu = 0.01 + 0.02 *
 * np.random.uniform()]; `randfac1 = np.random.normal(size=(32, nsteps))`; `randfac2 = ...` and return the only entries
 * the solver reads (Burger.py:416-419): r1_dev / r2_dev [n, 3, stepper] = rows 1..3, columns < stepper (DEVICE double),
 * nu_dev [n] when nunoise (else may be NULL).  Which candidates the polar method accepts is bit-identical to NumPy; the
 * kept values agree to the last bit of log() (<= 1 ulp).  One warp per seed, ~2 ms per 4096 seeds at nsteps = 5000. */
int mpde_forcing_tables(const int64_t* seeds_dev, int64_t n, int32_t nsteps, int32_t stepper, int32_t nunoise, double* r1_dev,
                        double* r2_dev, double* nu_dev, void* stream);
const char* mpde_rng_last_error(void);

/* Ground truth for the MSE reward at scale: sample the tensor-product B-spline that
 * setGroundTruth builds (Burger.py:322-323; FITPACK knots tx [ntx], ty [nty], coefficients c [(ntx-kx-1)*(nty-ky-1)],
 * degrees kx, ky in 1..3 -- all DEVICE double arrays) at out[q, i, j] = S(xq[q, j], tq[i]) for nq shifted grids of N points
 * and `rows` times: the [nq, rows, N] table mpde_set_truth consumes (dtype MPDE_F64 / MPDE_F32).  FITPACK bispev
 * semantics (arguments clamped to the spline's domain).  Returns 0 / -1. */
/* Interpolating tensor-product spline of odd degree k (1 = interp2d kind 'linear', 3 = 'cubic') through z[it][ix] = uu_truth on
 * the grids x [mx], t [mt] (all DEVICE double): FITPACK regrid / fpregr with s = 0, i.e. what interp2d builds in setGroundTruth
 * (Burger.py:322-323).  Writes the knots tx [mx+k+1], ty [mt+k+1] and the coefficients c [mx*mt] (c[ix*mt + it], FITPACK order)
 * that mpde_eval_spline_table consumes.  work_dev: >= mt*mx + 7*(mx+mt) doubles.  Coefficients agree with FITPACK's to rounding
 * (banded LU instead of Givens rotations on the same totally positive collocation matrix). */
int mpde_fit_spline(const double* x_dev, int32_t mx, const double* t_dev, int32_t mt, const double* z_dev, int32_t k, double* tx_dev,
                    double* ty_dev, double* c_dev, double* work_dev, void* stream);
int mpde_eval_spline_table(const double* tx_dev, int32_t ntx, const double* ty_dev, int32_t nty, const double* c_dev, int32_t kx,
                           int32_t ky, const double* xq_dev, int64_t nq, int32_t N, const double* tq_dev, int64_t rows, void* out_dev,
                           int32_t dtype, void* stream);

/* step(actions) x nsub + getState + reward (Burger.py:333-499, 604-675; KS.py:230-274, 369-383;
 * Diffusion.py:164-216; Advection.py:154-213; burger_environment.py:148-176).
 *   actions_dev : [B, M] real or NULL (step() without actions)
 *   nsub        : solver steps with these actions held fixed (nIntermediate); 0 = getState only
 *   state_out   : [B, mpde_state_size] real or NULL
 *   reward_out  : [B, A] real or NULL
 * status is kept in the library (MPDE_FIELD_STATUS). */
int mpde_step(mpde_env* env, const void* actions_dev, int32_t nsub, void* state_out, void* reward_out, void* stream);

/* Same as mpde_step for a HOST-side caller: `actions_host`, `state_host`, `reward_host` are (preferably
 * pinned) host buffers; the library stages them through device buffers it owns and enqueues
 * H2D copy -> step kernel -> D2H copies on `stream`, returning immediately.  The results are valid once the
 * stream (or an event recorded after the call) has completed.  This is the call the reference's environment
 * loop maps to when the policy lives on the host (burger_environment.py:140-190: s["Action"] in, s["State"] /
 * s["Reward"] out). */
int mpde_step_host(mpde_env* env, const void* actions_host, int32_t nsub, void* state_host, void* reward_host, void* stream);
/* Same with ONE host output buffer [B*S state | B*A reward] (a single pinned allocation): state and reward come back in a
 * single D2H copy, which is worth ~30 % of the end-to-end rate at B = 4096 (one DMA operation per direction per step).
 * The reward part is written when a reward mode is set and nsub > 0. */
int mpde_step_host_packed(mpde_env* env, const void* actions_host, int32_t nsub, void* out_host, void* stream);

/* A-priori sub-grid-scale term of the recorded history, Burger.compute_Sgs(nURG) (Burger.py:677-736) / KS.compute_Sgs
 * (KS.py:385-409): for every row of the uu history bound with mpde_set_history (rows = its length), filter u and u^2 with the
 * sharp cut |k| > nURG // 2 and write  sgs_out [B, rows, N] = -uh duh/dx + 0.5 d(u^2)h/dx  and, for Burgers (NULL to skip),
 * alt_out [B, rows, N] = duh/dt + uh duh/dx - nu d2uh/dx2 and alt2_out [B, rows, nURG] (the same on the nURG-point grid of the
 * truncated spectrum).  N in {256, 512, 1024, 2048}, nURG <= 64.  Evaluation-time diagnostic (testing mode), not a hot path. */
int mpde_compute_sgs(mpde_env* env, int32_t nURG, int64_t rows, void* sgs_out, void* alt_out, void* alt2_out, void* stream);

/* attribute access (u, v, Fn_old, ioutnum, t, ...): DEVICE destination / source of the natural shape */
int mpde_get(mpde_env* env, int32_t field, void* dst_dev, void* stream);
int mpde_set(mpde_env* env, int32_t field, const void* src_dev, void* stream);

/* introspection for the bench: kernels launched so far by this handle */
int64_t mpde_launch_count(const mpde_env* env);

/* ---- peer-memory gather of per-environment summaries (multi-GPU, one process per GPU, one node) ------------
 * The reference is serial; this is the learner-side exchange of a sharded batch.  Buffers are
 * cudaMalloc'ed by the library, shared with the other ranks of the node through CUDA IPC handles, and written
 * with plain stores over NVLink by mpde_peer_put (no NCCL call, no host round trip per step).
 *   alloc/free   : device buffer that can be exported
 *   export/open  : 64-byte IPC handle of a buffer / map a peer's buffer into this process
 *   put          : copy `nbytes` from `src` into EVERY rank's gather buffer at `dst_offset_bytes`, then publish
 *                  `step` in slot `my_rank` of every rank's flag array (int64 [nranks])
 *   wait         : make `stream` wait until all slots of this rank's flag array are >= step; bounded in TIME:
 *                  a peer that has not published after `timeout_us` microseconds (<= 0: MPDE_PEER_TIMEOUT_S seconds,
 *                  default 30) sets *err = 1 + rank instead of hanging.  A timed-out wait_next / exchange_next does NOT
 *                  advance the expected step, and once *err is set later waits give up after 1 ms, so a dead peer costs
 *                  one timeout.  `err` may be device memory or mapped pinned host memory (mpde_host_flag_alloc), which
 *                  the host can poll without synchronising.
 *   host_flag_*  : int32 [n] of zero-initialised mapped pinned host memory (device-accessible under UVA) */
int mpde_peer_alloc(size_t bytes, void** out);
int mpde_peer_free(void* p);
int mpde_peer_export(void* dev_ptr, void* handle64);
int mpde_peer_open(const void* handle64, void** out);
int mpde_peer_close(void* p);
int mpde_peer_put(const void* src, size_t nbytes, void* const* dst_ptrs, size_t dst_offset_bytes, void* const* flag_ptrs,
                  int32_t my_rank, int32_t nranks, int64_t step, void* counter_dev, void* stream);
int mpde_peer_wait(const void* my_flags_dev, int32_t nranks, int64_t step, void* err, int64_t timeout_us, void* stream);
int mpde_host_flag_alloc(int32_t n_int32, void** out);
int mpde_host_flag_free(void* p);
const char* mpde_peer_last_error(void);

/* Gather FUSED into the step kernel (warp-resident Burgers kernels, N <= 256): after this call every mpde_step of
 * `env` repeats its state / reward stores into `state_ptrs[i]` / `reward_ptrs[i]` (i < n_data: this rank's slab inside
 * the gather buffer of every OTHER rank, peer-mapped through mpde_peer_open; the local slab is the state_out /
 * reward_out of the call).  parity_stride > 0 double-buffers: the s-th step enqueued since this call (0-based, counted
 * by the library on the host) writes every buffer -- the local state_out / reward_out included -- at an offset of
 * (s & 1) * parity_stride ELEMENTS, so a fast rank's next step never lands in rows a slower rank's learner is still
 * reading (a CUDA graph that replays steps of `env` must therefore hold an even number of them).
 * mc_state / mc_reward (both, or mc_reward alone = only the rewards are gathered, the state stays local): this rank's slab addressed through an NVSwitch MULTICAST mapping of the gather
 * buffers (cuMulticast* / torch symmetric memory): one multimem.st per row then reaches every rank, this one included,
 * replacing the local store and the per-peer loop (pass n_data = 0).
 * n_data = 0, parity_stride = 0 and no multicast pointers unbinds.
 * Publishing is stream-ordered behind the step: mpde_peer_signal_next bumps *step_dev (int64, device) and stores it to
 * the n addresses flag_ptrs[j] = &flag_array_of_rank_j[my rank] (this rank included); mpde_peer_wait_next makes the
 * stream wait until every slot of this rank's flag array has reached ++(*expect_dev).  All counters live on the device,
 * so step + signal + wait replay from a CUDA graph; neither needs to sit on the step kernels' own stream. */
int mpde_set_peer_output(mpde_env* env, int32_t n_data, void* const* state_ptrs, void* const* reward_ptrs, int64_t parity_stride,
                         void* mc_state, void* mc_reward);
/* this rank's OWN slab (copy 0) inside its gather buffer: lets mpde_step_host / mpde_step_host_packed run with the gather
 * bound (the kernel writes the rows there -- through the multicast address when one is bound -- and the D2H copy reads
 * the copy of the current parity).  local_reward = local_state + B*S for the packed single-copy path. */
int mpde_set_peer_local(mpde_env* env, void* local_state, void* local_reward);
/* How the single-agent state rows travel in the fused gather: on != 0 (default) = staged through shared memory and written
 * as whole rows (256 contiguous bytes per 16 lanes: the NVLink write efficiency an 8-GPU gather into one learner needs:
 * 0.53 -> 0.60-0.66 TB/s of ingress); 0 = every lane stores its own 16-byte pieces to every destination (shorter epilogue:
 * 7.9 instead of 8.1 us per step on 2 GPUs, where the links are not the bound).  Sticky across mpde_set_peer_output. */
int mpde_set_peer_row_stores(mpde_env* env, int32_t on);
int mpde_peer_signal_next(void* const* flag_ptrs, int32_t n, void* step_dev, void* stream);
int mpde_peer_wait_next(const void* my_flags_dev, int32_t nranks, void* expect_dev, void* err, int64_t timeout_us, void* stream);
/* signal_next + wait_next as ONE kernel launch */
int mpde_peer_exchange_next(void* const* flag_ptrs, int32_t n, void* step_dev, const void* my_flags_dev, int32_t nranks,
                            void* expect_dev, void* err, int64_t timeout_us, void* stream);

/* One RL step of this rank's shard INCLUDING the gather, as ONE host call (the multi-GPU form of the loop body of
 * burger_environment.py:148-176): step kernel with the fused peer stores -> publish -> wait for every rank's rows.
 *   mpde_set_peer_sync : once, after mpde_set_peer_output -- the arguments mpde_peer_exchange_next takes.
 *   mpde_step_fused    : async_gather = 0: kernel and exchange run in stream order on `stream`, replayed from a CUDA graph
 *                        cached per (buffers, nsub, parity) -- one driver call per RL step; work enqueued behind it sees
 *                        the gathered rows of ALL ranks.  async_gather = 1: the exchange runs on a side stream the
 *                        library owns, forked behind the step kernel, so `stream` is free for the next independent
 *                        batch; mpde_peer_join(env, s) makes stream s wait for the last exchange (the caller may
 *                        capture step_fused(async) ... peer_join into a CUDA graph of its own).
 * A step that fails to launch does not flip the parity of the double-buffered gather copies.
 * mpde_step(nsub = 0) stays available with a gather bound (getState / getMseReward of the current state at episode
 * reset): it writes the caller's local buffers only and publishes nothing. */
int mpde_set_peer_sync(mpde_env* env, void* const* flag_ptrs, int32_t n, void* step_dev, const void* my_flags_dev, int32_t nranks,
                       void* expect_dev, void* err, int64_t timeout_us);
int mpde_step_fused(mpde_env* env, const void* actions_dev, int32_t nsub, void* state_out, void* reward_out, int32_t async_gather,
                    void* stream);
int mpde_peer_join(mpde_env* env, void* stream);

const char* mpde_last_error(void);
int mpde_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MARLPDE_B200_H */
