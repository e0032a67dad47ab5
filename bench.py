#!/usr/bin/env python3
"""Benchmark of the marlpde environment time-stepper hot path (BASELINE.json metric:
env-steps/s, batched Burgers LES N=32 x 4096 envs per GPU, fp64, stochastic forcing,
spectral reward, nIntermediate = 10 solver steps per RL step).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo (CUDA)
  python bench.py --impl reference ...                             # CPU arm: the reference's own Burger class

One "step" = one RL step of one batch = ONE kernel launch: 10 ABCN solver steps with the actions held
fixed + getState + spectral reward (burger_environment.py:148-176).  Prints ONE JSON line (driver contract).

Timed region: the K steps are captured once into ONE CUDA graph (pool of independent batches in rotation, so the
working set exceeds L2) and replayed; multi-GPU runs use the same graph with the gather fused into the step
kernel (mpde_step_fused) -- whatever K the driver asks for, no step is launched from Python.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ----------------------------------------------------------------------------- workload
B_PER_GPU = 4096           # BASELINE.json configs[1]
N, M, NSUB = 32, 32, 10
L_DOM, DT, NU, TEND = 2 * np.pi, 1e-3, 0.02, 5.0
POOL = 24                  # independent batches rotated so the working set exceeds L2
STABLE_SEEDS = (50, 59, 81, 89)
# ALGORITHMIC bytes per environment per LAUNCH (Burgers C2, per-step-I/O figure of one state round
# trip; with NSUB fused sub-steps the state makes that round trip once per launch, so bytes per env-step = 1912 / NSUB):
#   read  actions M*8 + v,Fn_old 2*(N+2)*8 + forcing coefficients 6*8 + Ek sums (N/2)*8
#   write v,Fn_old 2*(N+2)*8 + Ek sums (N/2)*8 + state S*8 + reward A*8          = 8*(32+136+6+32+32+1) = 1912 B
BYTES_PER_ENV_LAUNCH = 8 * (M + 4 * (N + 2) + 6 + N + N + 1)
FLOPS_PER_ENV_STEP = 2600           # algorithmic fp64 flops of one Burgers N=32 solver step
BYTES_C5 = 8 * (M + 4 * (N + 2) + 3 * N + N)                  # 2368 B per env per launch
BYTES_KS = 8 * (64 + 2 * (64 + 2) + 64 + 128 + 1)             # 3112 B per env per launch
FLOPS_KS_STEP, FLOPS_DNS_STEP = 15000, 82000
REF_DIR = os.path.join(ROOT, "baseline", "_ref", "_model")     # tools/install_ref.sh (git-ignored, travels with gpurun)


def kernel_facts():
    """Machine / profile facts measured by the tools (not constants of this file): profiles/kernel_facts.json is written
    by tools/ncu_summary.py (ncu --set full DRAM bytes per launch) and tools/microbench.cu (FP64 DFMA peak)."""
    p = os.path.join(ROOT, "profiles", "kernel_facts.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.th.join(timeout=2)
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def spectrum_table(seed=0):
    """Synthetic stand-in for dns.Ek_ktt[:, :N/2] (5001 rows): same shape / indexing as burger_environment.py:174."""
    return np.abs(np.random.default_rng(seed).normal(1.0, 0.1, (int(TEND / DT) + 1, N // 2))) * 1e-3 + 1e-6


# ----------------------------------------------------------------------------- CPU arm
def _cpu_worker_reference(args):
    """One process = ONE environment of the UNMODIFIED reference class (baseline/_ref/_model/Burger.py), driven the way
    burger_environment.environment drives it in spectral-reward mode (burger_environment.py:134-192): nIntermediate x
    sgs.step(actions), sgs.getState(), sgs.compute_Ek() (rescans the whole history, as the reference does), kRelErr,
    reward; a FloatingPointError (np.seterr raise, Burger.py:8) or the end of the episode starts a new episode with a
    new Burger object, as Korali would."""
    idx, seconds = args
    os.environ["OMP_NUM_THREADS"] = "1"
    sys.path.insert(0, REF_DIR)
    import Burger as RB
    rng = np.random.default_rng(idx)
    ref = spectrum_table()
    gs = N

    def new_env(ep):
        sgs = RB.Burger(L=L_DOM, N=N, dt=DT, nu=NU, tend=TEND, case="turbulence", forcing=True, dforce=False,
                        seed=42 + idx + 1000 * ep, version=0, noise=0., s=1)
        sgs.setup_basis(M, "hat")
        return sgs

    actions = rng.uniform(0.02, 0.1, M).tolist()
    episodes, rl_done, prev = 0, 0, 0.
    sgs = new_env(0)
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        try:
            for _ in range(NSUB):
                sgs.step(actions)
            state = sgs.getState()                                                            # noqa: F841
            sgs.compute_Ek()
            k = np.mean(((np.abs(ref[sgs.ioutnum, 1:gs // 2] - sgs.Ek_ktt[sgs.ioutnum, 1:gs // 2])) / ref[sgs.ioutnum, 1:gs // 2]) ** 2)
            reward = prev - k                                                                 # noqa: F841
            prev = k
            rl_done += 1
            if sgs.ioutnum + NSUB > sgs.nsteps:
                raise FloatingPointError("episode over")
        except FloatingPointError:
            episodes += 1
            prev = 0.
            sgs = new_env(episodes)
    return rl_done * NSUB, time.perf_counter() - t0, episodes


def _cpu_worker_port(args):
    """Fallback when baseline/_ref is absent: the repo's numpy restatement (oracle/), one environment per process,
    one step() per Python call.  Environments are re-initialised before they blow up (alive check every RL step)."""
    idx, seconds = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle.burger_oracle import BurgerOracle, forcing_tables, turbulence_ic
    from oracle.common import grid, spectral_rel_err
    rng = np.random.default_rng(idx)
    o = BurgerOracle(B=1, L=L_DOM, N=N, dt=DT, nu=NU, forcing=True, dforce=False)
    o.setup_basis(M, "hat")
    sd = STABLE_SEEDS[idx % len(STABLE_SEEDS)]
    r1, r2 = forcing_tables(sd, int(TEND / DT))
    o.set_forcing_tables(r1[:, :1], r2[:, :1])
    u0 = turbulence_ic(grid(L_DOM, N), L_DOM, N, 0.0, sd)[None]
    o.IC(u0=u0)
    ref = spectrum_table()
    acts = rng.uniform(0.02, 0.1, (1, M))
    done, episodes, t0 = 0, 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for _ in range(NSUB):
            o.step(acts)
        o.state()
        spectral_rel_err(ref[min(o.ioutnum, 5000)], o.Ek_ktt_row()[0], N)
        done += 1
        if o.ioutnum >= 3000 or not np.all(np.abs(o.u) < 1e3):          # new episode before anything overflows
            o.IC(u0=u0)
            episodes += 1
    return done * NSUB, time.perf_counter() - t0, episodes


def cpu_run(seconds, cores=None):
    cores = cores or len(os.sched_getaffinity(0))
    kind = "reference" if os.path.exists(os.path.join(REF_DIR, "Burger.py")) else "port"
    worker = _cpu_worker_reference if kind == "reference" else _cpu_worker_port
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(worker, [(i, seconds) for i in range(cores)])
    steps = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    what = ("the UNMODIFIED reference Burger class (baseline/_ref/_model, tools/install_ref.sh) driven as burger_environment.py:134-192 "
            "does: 10 x step(actions) + getState + compute_Ek (whole-history rescan) + spectral reward per RL step"
            if kind == "reference" else
            "the repo's numpy restatement (oracle/; baseline/_ref absent): 10 x step + state + spectral reward per RL step")
    sample = (f"{cores} single-environment processes (one per host core, OMP_NUM_THREADS=1) x {seconds:.0f} s of the bench workload "
              f"(Burgers N=32, forcing, eddy-viscosity actions, M=32 hat basis); {what}; "
              f"{sum(r[2] for r in res)} episodes restarted")
    return steps / busy, cores, kind, sample


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    seconds = float(min(60.0, max(4.0, args.cpu_seconds)))
    value, cores, kind, sample = cpu_run(seconds)
    line = {
        "impl": "reference", "metric": "env_steps_per_s", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * cores * NSUB / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "one step of this arm = one RL step (10 solver steps) of `cores` environments, one per host core; the sample is "
                "time-boxed, --steps/--warmup only label the line",
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus):
    # identical in both arms (the driver compares the config of the reference arm with ours)
    return {"workload": (f"Burgers LES N={N} x {B_PER_GPU} envs/GPU (BASELINE configs[1]), fp64, M={M} hat basis, "
                         f"eddy-viscosity actions (dforce=False), 3-mode stochastic forcing with per-env seeds (42 + global env id), "
                         f"per-env turbulence ICs, spectral reward vs one shared DNS N=512, "
                         f"nIntermediate={NSUB} solver steps per RL step; one step = one RL step of one batch"),
            "envs_per_gpu": B_PER_GPU, "N": N, "M": M, "n_intermediate": NSUB, "global_envs": B_PER_GPU * n_gpus,
            "l2": f"rotating pool of {POOL} independent batches per GPU (state working set > 126 MB L2)",
            "parallelism": "single GPU" if n_gpus == 1 else f"env-sharded x{n_gpus}, state+reward rows gathered to the learner rank each RL step"}


# ----------------------------------------------------------------------------- GPU arm
_DNS_SPECTRUM = {}


def dns_spectrum(torch, device):
    """Reward reference of the bench workload: time-averaged spectrum rows Ek_ktt[:, :N/2] of ONE shared DNS
    (N = 512, forcing, seed 42), generated here by the library's own DNS kernels (setup_dns_default of
    burger_environment.py:11-16).  Falls back to the synthetic table if that DNS does not survive the 5000 steps."""
    key = str(device)
    if key not in _DNS_SPECTRUM:
        from marlpde_b200 import Burger
        dns = Burger(L=L_DOM, N=512, dt=DT, nu=NU, tend=TEND, case="turbulence", forcing=True, seed=42, nenvs=1, device=device,
                     history=True)
        dns.simulate()
        ok = int((dns.status != 0).sum()) == 0
        tab = dns._ektt[0, :, :N // 2].clone() if ok else None
        if ok and not bool(torch.isfinite(tab).all() and (tab[:, 1:] > 0).all()):
            ok = False
        _DNS_SPECTRUM[key] = (tab.cpu().numpy() if ok else spectrum_table(), "DNS N=512" if ok else "synthetic table (DNS blew up)")
        del dns
    return _DNS_SPECTRUM[key]


def make_batch(torch, device, batch_id, B=None, team_lanes=0, spec=None, per_env_seeds=True):
    """One batch of the bench workload (BASELINE configs[1]): every environment has its own forcing seed
    42 + global id (tables generated on the device, Burger.py:66,94-95) and its own turbulence IC (Burger.py:227-260);
    spectral reward against one shared DNS N = 512."""
    from marlpde_b200 import Burger
    B = B or B_PER_GPU
    if per_env_seeds:
        seeds = 42 + batch_id * B + np.arange(B)
    else:
        # explanatory sweeps run each batch for > 100 RL steps: the forced N=32 LES blows up for most forcing seeds within
        # ~600 solver steps (the reference's own physics); these four stay bounded under a positive eddy viscosity
        seeds = np.array(STABLE_SEEDS)[(np.arange(B) + batch_id) % len(STABLE_SEEDS)]
    env = Burger(L=L_DOM, N=N, dt=DT, nu=NU, tend=TEND, case="turbulence", forcing=True, dforce=False, seed=seeds,
                 nenvs=B, device=device, history=False, team_lanes=team_lanes)
    env.setup_basis(M, "hat")
    env.set_spectrum_reference(dns_spectrum(torch, device)[0] if spec is None else spec)
    return env


def make_batch_c5(torch, device, seed0, B=8192, **_):
    """BASELINE configs[4] per GPU: MARL Burgers N=32, 32 per-gridpoint agents (state windows of 3, one action each),
    MSE reward against a shared truth table, 4-lane teams (the large-batch kernel)."""
    from marlpde_b200 import Burger
    seeds = np.array(STABLE_SEEDS)[(np.arange(B) + seed0) % len(STABLE_SEEDS)]
    env = Burger(L=L_DOM, N=N, dt=DT, nu=NU, tend=TEND, case="turbulence", forcing=False, dforce=False, seed=seeds, version=0,
                 numAgents=N, nenvs=B, device=device, history=False, team_lanes=4)
    env.setup_basis(M, "hat")
    env.set_truth_table(np.random.default_rng(seed0).normal(1.0, 0.3, (int(TEND / DT) + 1, N))[None])
    return env


def time_graph(torch, launches, reps=3, warm=2, chains=1):
    """Capture the list of callables `launches` (one kernel launch each, on independent batches) into a CUDA graph;
    best-of-`reps` ms per launch.  chains > 1: consecutive launches alternate between `chains` streams inside the graph
    (independent batches in flight, as in the headline measurement)."""
    for fn in launches:
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    streams = [torch.cuda.Stream() for _ in range(chains)] if chains > 1 else []
    with torch.cuda.graph(g):
        cap = torch.cuda.current_stream()
        for cs in streams:
            cs.wait_stream(cap)
        for i, fn in enumerate(launches):
            if streams:
                with torch.cuda.stream(streams[i % chains]):
                    fn()
            else:
                fn()
        for cs in streams:
            cap.wait_stream(cs)
    for _ in range(warm):
        g.replay()
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(100_000)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / len(launches)
        best = ms if best is None else min(best, ms)
    del g
    return best


def pin_cores(local, world):
    """One disjoint slice of the host cores per rank (all ranks of a node otherwise share the same affinity mask and
    their launch threads migrate over each other)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // world
        if world > 1 and per >= 1:
            os.sched_setaffinity(0, set(cores[local * per:(local + 1) * per]))
            return per
    except Exception:
        pass
    return None


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores_per_rank = pin_cores(local, world)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    K, W = args.steps, args.warmup
    pool = max(1, args.pool)
    make = make_batch_c5 if args.workload == "c5" else make_batch
    B = 8192 if args.workload == "c5" else B_PER_GPU
    if args.workload == "c5":
        envs = [make(torch, device, rank * pool + i) for i in range(pool)]
    else:
        envs = [make(torch, device, rank * pool + i, team_lanes=args.lanes) for i in range(pool)]
    rng = np.random.default_rng(rank)
    # eddy-viscosity coefficients: one value per (environment, action), positive (a stabilising closure)
    acts_host = torch.from_numpy(rng.uniform(0.02, 0.1, (pool, B, M))).pin_memory()
    acts = acts_host.to(device)
    S = envs[0]._state_size
    RW = envs[0]._reward_buf.shape[1]
    gathers = []
    fused = world > 1 or args.fused_single      # --fused-single: 1-GPU diagnostic of the fused-gather overheads
    if fused:
        # Learner-side gather FUSED into the step kernel: every rank's kernel stores its state + reward rows straight
        # into every rank's (double-buffered) gather buffer over NVLink (one multimem.st per row on NVSwitch) and a
        # 1-thread kernel behind it publishes the step / waits for the peers.  No NCCL call, no host work per step.
        from marlpde_b200.dist import PeerGather
        for env in envs:
            pg = PeerGather(B * (S + RW), torch.float64, device, copies=2)
            pg.fuse(env, B, S, RW, gather_state=not args.rewards_only, learner=0 if args.gather == "learner" else None)
            gathers.append(pg)

    def one_step(i):
        k = i % pool
        if fused:
            # ONE library call: step kernel -> [side stream: publish + wait]; the main stream is free for the next batch
            envs[k].step_n_fused(acts[k], NSUB, async_gather=True)
            gathers[k].step += 1
        else:
            envs[k].step_n(acts[k], NSUB)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def capture(nsteps, chains, first=0):
        """CUDA graph of RL steps first .. first+nsteps-1 of the pool rotation.  chains > 1: consecutive (independent)
        batches alternate between `chains` streams inside the graph, so one batch's prologue / tail overlaps another
        batch's sub-step loop (each chain serialises its own kernels through programmatic dependent launch)."""
        l0 = sum(e.launch_count for e in envs)
        g_ = torch.cuda.CUDAGraph()
        cstreams = [torch.cuda.Stream(device=device) for _ in range(chains)] if chains > 1 else []
        with torch.cuda.graph(g_):
            cap = torch.cuda.current_stream()
            for cs in cstreams:
                cs.wait_stream(cap)
            for i in range(first, first + nsteps):
                if chains > 1:
                    with torch.cuda.stream(cstreams[i % chains]):
                        one_step(i)
                else:
                    one_step(i)
            for cs in cstreams:
                cap.wait_stream(cs)
            if fused:
                for e in envs:
                    e.peer_join()           # the graph ends when every batch's gather has completed
        n_k = sum(e.launch_count for e in envs) - l0
        torch.cuda.synchronize()
        return g_, n_k

    # independent batches in flight inside the graph: 4 saturate the chip in steady state; a short run (K <= pool, every batch
    # steps once) gains another 4 % from 10, which shortens the fill / drain of the pipeline (profiles/r2_lanes_chains.md)
    want_chains = args.chains if args.chains > 0 else (10 if K <= pool else 4)
    chains = max(1, min(want_chains, pool, K))
    if K > pool and pool % chains:
        chains = 1                          # a batch must stay on one chain (its launches are ordered by its stream)
    # every batch once outside any graph (module load, lazy set-up, first-use allocations)
    for i in range(pool):
        one_step(i)
    if fused:
        for e in envs:
            e.peer_join()
    sync()

    def new_episode():
        """Untimed: put every batch back at t = 0 (the forced N=32 LES only stays bounded for about one episode)."""
        for e in envs:
            e.IC(case="turbulence")
        sync()

    # The timed region is ONE replay of a graph holding exactly K steps (K <= 960), else whole rotations + a tail graph.
    seg = K if K <= 960 else pool * 8
    graph, per_graph = capture(seg, chains)
    tail_graph, per_tail = (None, 0)
    if K % seg:
        tail_graph, per_tail = capture(K % seg, chains)

    def run_timed():
        launched = 0
        for _ in range(K // seg):
            graph.replay()
            launched += per_graph
        if tail_graph is not None:
            tail_graph.replay()
            launched += per_tail
        return launched

    sampler = ClockSampler(local) if rank == 0 else None      # covers warm-up + timed + e2e regions
    # warm-up: (a) ~3000 steps of the same graph so the SM clock has ramped before anything is timed, (b) a fresh episode,
    # (c) the W steps the driver asks for (rounded up to whole replays of the graph)
    # (the number of replays is a function of K only: with a gather every rank must replay the same number of steps)
    ramp_steps = 0
    for _ in range(max(2, 3000 // seg)):
        graph.replay()
        ramp_steps += seg
    torch.cuda.synchronize()
    new_episode()
    warm_steps = 0
    while warm_steps < max(W, 3):
        graph.replay()
        warm_steps += seg
    sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # a ~100 us spin kernel keeps the (synchronised, idle) stream busy while the host enqueues the start event and the graph,
    # so the bracket [ev0, ev1] holds the K steps on the device and not the host's graph-launch latency
    torch.cuda._sleep(200_000)
    ev0.record()
    launches = run_timed()
    ev1.record()
    sync()
    ms = ev0.elapsed_time(ev1)
    for g in gathers:
        g.check()
    alive_frac = float(np.mean([float((e.status == 0).double().mean()) for e in envs]))

    # explanatory extra (not the headline): the same K steps strictly one batch after the other (one chain)
    ms1 = None
    if chains > 1 and not args.quick:
        new_episode()
        g1, _ = capture(seg, 1)
        for _ in range(2):
            g1.replay()
        sync()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(200_000)
        e2.record()
        g1.replay()
        e3.record()
        sync()
        ms1 = e2.elapsed_time(e3) / seg
        del g1

    # ---- shard invariance, driver-side evidence (untimed): rank 0 recomputes rows of other ranks locally ----------
    gather_parity = None
    if world > 1:
        new_episode()
        envs[0].step_n_fused(acts[0], NSUB)              # kernel -> publish -> wait, in stream order
        gathers[0].step += 1
        sync()
        gathers[0].check()
        if rank == 0:
            cur = gathers[0].current()
            ok = True
            for r in sorted({1, world // 2, world - 1}):
                twin = make(torch, device, r * pool)                               # rank r's batch 0, rebuilt here
                a_r = torch.from_numpy(np.random.default_rng(r).uniform(0.02, 0.1, (pool, B, M))[0]).to(device)
                st, rw = twin.step_n(a_r, NSUB)
                torch.cuda.synchronize()
                rows = slice(0, B, max(1, B // 64))
                got_rw = cur[r][B * S:].view(B, RW)
                ok = ok and bool(torch.equal(got_rw[rows], rw[rows]))
                if not args.rewards_only:
                    ok = ok and bool(torch.equal(cur[r][:B * S].view(B, S)[rows], st[rows]))
                del twin
            gather_parity = ok
        sync()
    new_episode()

    # ---- end to end through the public API with HOST buffers --------------------------------
    # Every RL step of every batch: pinned-host actions -> H2D -> step_n (one launch) -> D2H of state and
    # reward -> the host waits for them before that batch gets its next actions.  The learner keeps
    # `depth` independent batches in flight (marlpde_b200.pipeline.HostPipeline) so PCIe transfers of one
    # batch overlap the kernel of another; each batch's own action->state chain stays strictly serial.
    # Multi-GPU: the kernel still stores its rows into every rank's gather buffer; each rank PUBLISHES behind its step
    # (1-thread kernel on the slot's stream, no wait) and only the learner rank waits for the peers, on its own stream.
    from marlpde_b200.pipeline import HostPipeline
    depth = min(pool, max(1, args.depth))
    learner = torch.cuda.Stream(device=device) if (fused and rank == 0) else None

    def publish(k, st, rw):
        gathers[k].step += 1
        gathers[k].signal_next()
        return None

    def publish_on(k, stream):
        gathers[k].step += 1
        gathers[k].signal_next(stream)

    pipe = HostPipeline(envs[:depth], NSUB, post_step=publish if fused else None, post_step_on=publish_on if fused else None)
    for k in range(depth):
        pipe.act_host[k].copy_(acts_host[k])
    # e2e steps: at least 10 per batch in flight -- a 20-step run with 8 batches in flight would time the fill and the
    # drain of the pipeline, not its rate -- and at most 40 per batch (the forced LES stays bounded for ~50); reported
    # as e2e.steps
    Ke = min(max(K, 10 * depth), 40 * depth)
    checksum = 0.0

    stamps = []
    rw_np = [r.numpy() for r in pipe.reward_host]     # numpy views of the pinned result buffers

    def e2e_round(n):
        nonlocal checksum
        stamps.clear()
        for i in range(n):
            k = i % depth
            pipe.collect(k)                           # results of this batch's previous step are on the host
            checksum += float(rw_np[k][0, 0])         # the host really reads them
            stamps.append(time.perf_counter())
            pipe.submit(k)                            # next actions for this batch (already in pinned memory)
            if learner is not None:
                learner.wait_event(pipe.done[k])
                gathers[k].wait_next(learner)
        pipe.drain()
        if learner is not None:
            learner.synchronize()

    e2e_round(max(3 * depth, Ke))       # untimed: graph captures of every slot, host caches and clocks settled
    new_episode()                       # the timed round starts from fresh episodes (<= 40 RL steps per batch)
    t0 = time.perf_counter()
    e2e_round(Ke)
    torch.cuda.synchronize()
    e2e_local = time.perf_counter() - t0
    # rate inside the round (host time stamps at every collect): first and second half, in us per step
    half = len(stamps) // 2
    e2e_halves = [round((stamps[half] - stamps[0]) / half * 1e6, 2), round((stamps[-1] - stamps[half]) / (len(stamps) - 1 - half) * 1e6, 2)] if half > 1 else None
    sync()
    e2e_s = e2e_local
    clocks = sampler.stop() if sampler else None
    # PCIe floor of this box (untimed extra): the same pinned buffers and slot streams moving the same bytes per step in
    # both directions with no kernel in between -- what e2e can at best reach here
    floor_us = None
    if world == 1:
        n_floor = 40 * depth
        floor_dev = torch.empty(pipe.out_host[0].numel(), dtype=pipe.out_host[0].dtype, device=device)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        for i in range(n_floor):
            k = i % depth
            pipe.done[k].synchronize()
            with torch.cuda.stream(pipe.streams[k]):
                pipe.act_dev[k].copy_(pipe.act_host[k], non_blocking=True)
                pipe.out_host[k].copy_(floor_dev, non_blocking=True)
                pipe.done[k].record()
        torch.cuda.synchronize()
        floor_us = (time.perf_counter() - t1) / n_floor * 1e6
    for g in gathers:
        g.poll()

    per_rank_e2e = None
    if world > 1:
        t = torch.tensor([ms, e2e_s * 1e3, ms1 or 0.0, -alive_frac], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1]) / 1e3
        ms1 = float(t[2]) if ms1 is not None else None
        alive_frac = -float(t[3])
        allt = [torch.zeros(1, device=device, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allt, torch.tensor([e2e_local * 1e6 / Ke], device=device, dtype=torch.float64))
        per_rank_e2e = [round(float(x), 2) for x in allt]

    if rank == 0:
        total_envs = B * world
        value = total_envs * NSUB * K / (ms * 1e-3)
        peak, how = peaks()
        facts = kernel_facts()
        per_launch_s = ms * 1e-3 / K
        bytes_env = BYTES_C5 if args.workload == "c5" else BYTES_PER_ENV_LAUNCH
        achieved = B * bytes_env / per_launch_s / 1e9
        fp64_peak = float(facts.get("fp64_peak_tflops", 33.2))
        tf = B * NSUB * FLOPS_PER_ENV_STEP / per_launch_s / 1e12
        cfg = workload_config(world)
        if args.workload == "c5":
            cfg["workload"] = (f"MARL Burgers N={N} x {B} envs/GPU (BASELINE configs[4]), fp64, {N} per-gridpoint agents (state windows "
                               f"of 3, one eddy-viscosity action each), MSE reward vs a shared truth table, nIntermediate={NSUB}; 4-lane teams")
            cfg["envs_per_gpu"], cfg["global_envs"] = B, B * world
        line = {
            "metric": "env_steps_per_s", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg,
            "timing": {"how": f"{K} steps = ONE replay of a CUDA graph holding exactly {K} launches (pool rotation)" if K <= 960 else
                              f"{K // seg} replays of a {seg}-step graph + a {K % seg}-step tail graph",
                       "batches_in_flight": chains,
                       "warmup_steps_run": int(warm_steps), "clock_ramp_steps": int(ramp_steps),
                       "note": "the step kernels of consecutive (independent) batches alternate between `batches_in_flight` "
                               "streams inside the graph; device time by CUDA events around the replay (barrier + synchronize, then a "
                               "100 us spin kernel ahead of the start event hides the host's graph-launch latency), max over ranks"},
            "e2e": {"value": total_envs * NSUB * Ke / e2e_s, "unit": "env-steps/s",
                    "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
                    "steps": Ke, "batches_in_flight": depth, "us_per_step_per_rank": per_rank_e2e,
                    "us_per_step": e2e_s / Ke * 1e6, "us_per_step_first_second_half": e2e_halves, "pcie_floor_us_per_step": floor_us,
                    "cores_per_rank": cores_per_rank,
                    "note": "per RL step of a batch: pinned host actions -> H2D -> step_n -> D2H state+reward -> host waits; "
                            "independent batches overlap (HostPipeline); timed over `steps` = min(max(K, 10 x batches in "
                            "flight), 40 x batches in flight) RL steps, so that a short K does not time the pipeline's fill and drain" +
                            ("; multi-GPU: rows also stored into every rank's gather buffer, each rank publishes behind its step, "
                             "only the learner rank (0) waits for the peers" if fused else "")},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": facts.get("traffic_bytes_per_launch", {}).get("c5" if args.workload == "c5" else "c2"),
                         "traffic_source": facts.get("traffic_source"), "peak_source": how,
                         "kernel": "burgers_warp_kernel<double,32,4,FORCING|ACTIONS,HOT>" if args.workload != "c5"
                                   else "burgers_warp_kernel<double,32,4,generic>",
                         "bytes_per_launch": B * bytes_env,
                         "launch_us": per_launch_s * 1e6,
                         "launch_us_one_batch_at_a_time": None if ms1 is None else ms1 * 1e3,
                         "note": f"algorithmic bytes = {bytes_env} B per env per launch (SURVEY 8d) x {B} envs; launch_us = timed region / "
                                 "launches (kernels of independent batches overlap); the 10 solver steps fused into one launch keep the "
                                 "state on chip, so the launch is bound by FP64 + shuffle issue and their latencies, not by HBM: even at "
                                 "the FP64 peak the launch would take 3.2 us = 0.37 of the HBM roofline (see roofline_fp64, sweep, DESIGN.md 5)"},
            "roofline_fp64": {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak,
                              "peak_source": facts.get("fp64_peak_source", "measured (tools/microbench.cu DFMA loop, profiles/r1_microbench_b200.md)"),
                              "note": "2.6 kflop per env-step (SURVEY 8d) x 40960 env-steps per launch"},
            "all_envs_alive": alive_frac == 1.0, "alive_fraction": alive_frac,
            "reward_reference": dns_spectrum(torch, device)[1] if args.workload != "c5" else "truth table (MSE)",
        }
        if world > 1:
            line["gather_parity"] = gather_parity
            line["transport"] = (("NVSwitch multicast (multimem.st), %s memory" if gathers[0].multicast else "unicast peer stores, %s memory")
                                 % gathers[0].backend) + (", rewards only" if args.rewards_only else "") + \
                                (", gathered to rank 0 only" if args.gather == "learner" else ", gathered to every rank")
            nbytes = B * (RW + (0 if args.rewards_only else S)) * 8
            line["gather_bytes_per_step"] = {"per_rank_slab": nbytes, "ingress_learner": nbytes * (world - 1 if args.gather == "learner" else world),
                                             "ingress_other_ranks": 0 if args.gather == "learner" else nbytes * world,
                                             "ingress_gbs_at_this_rate": nbytes * (world - 1 if args.gather == "learner" else world) / (ms * 1e-3 / K) / 1e9}
        if world == 1 and not args.quick:
            del pipe
            envs.clear()
            torch.cuda.empty_cache()
            try:
                line["sweep"] = batch_sweep(torch, device)
            except Exception as e:          # the extras never take the headline down
                line["sweep"] = {"error": repr(e)}
            try:
                line["other_configs"] = other_configs(torch, device)
            except Exception as e:
                line["other_configs"] = {"error": repr(e)}
        if world == 1 and not args.no_cpu:
            v, cores, kind, sample = cpu_run(args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": kind, "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- driver-run extras (N = 1)
def batch_sweep(torch, device):
    """Bench workload at other batch sizes / fused-step counts (the HBM roofline is the bound of the per-step-I/O
    regime, the FP64 roofline of the fused regime).  Each entry: one CUDA graph over a pool of batches larger than L2."""
    peak, _ = peaks()
    fp64_peak = float(kernel_facts().get("fp64_peak_tflops", 33.2))
    out = []
    spec = spectrum_table()
    for B, nsub, lanes in ((4096, 10, 4), (4096, 10, 8), (8192, 10, 4), (32768, 10, 4), (32768, 1, 4), (131072, 1, 4)):
        per_batch = B * 1912
        pool = max(2, -(-160_000_000 // per_batch))
        if B <= 8192:
            pool = -(-pool // 4) * 4
        envs = [make_batch(torch, device, 7 + i, B=B, team_lanes=lanes, spec=spec, per_env_seeds=False) for i in range(pool)]
        a = torch.from_numpy(np.random.default_rng(B).uniform(0.02, 0.1, (B, M))).to(device)
        reps = max(1, 40 // pool)
        launches = [(lambda e=e: e.step_n(a, nsub)) for _ in range(reps) for e in envs]
        chains = 4 if (B <= 8192 and pool % 4 == 0) else 1          # small batches: independent batches in flight, as the headline
        ms = time_graph(torch, launches, chains=chains)
        alive = all(int((e.status != 0).sum()) == 0 for e in envs)
        gbs = per_batch / (ms * 1e-3) / 1e9
        tf = B * nsub * FLOPS_PER_ENV_STEP / (ms * 1e-3) / 1e12
        out.append({"envs": B, "n_sub": nsub, "lanes_per_env": lanes, "batches_in_flight": chains, "launch_us": ms * 1e3,
                    "env_steps_per_s": B * nsub / (ms * 1e-3), "hbm_frac": gbs / peak, "fp64_frac": tf / fp64_peak,
                    "all_envs_alive": alive})
        del envs
        torch.cuda.empty_cache()
    return out


OTHER_CHAINS = int(os.environ.get("MPDE_OTHER_CHAINS", "2"))     # independent batches in flight for configurations 3 and 5 (4 measures the same)


def other_configs(torch, device, only=None):
    """BASELINE configs[2], [3], [4] on one GPU, measured in-process after the headline (each < 1 s of GPU time)."""
    from marlpde_b200 import Burger, KS
    peak, _ = peaks()
    fp64_peak = float(kernel_facts().get("fp64_peak_tflops", 33.2))
    res = {}
    rng = np.random.default_rng(0)
    # ---- C3: KS L=22 N=64 x 8192, M=64 hat basis, 10 ETDRK4 steps + state per launch
    B, n, m = 8192, 64, 64
    pool = [KS(L=22, N=n, dt=0.25, nsteps=100000, nenvs=B, u0=rng.normal(0, 1e-3, (B, n)), history=False, device=device)
            for _ in range(8)]
    for k in pool:
        k.setup_basis(m, "hat")
    a = torch.as_tensor(rng.normal(0, 1e-3, (B, m)), device=device)

    ms = time_graph(torch, [(lambda k=k: k.step_n(a, 10, want_reward=False)) for k in pool], chains=OTHER_CHAINS)
    res["c3_ks_n64_x8192"] = {"launch_us": ms * 1e3, "batches_in_flight": OTHER_CHAINS, "value": B * 10 / (ms * 1e-3), "unit": "env-steps/s", "n_sub": 10,
                              "hbm_frac": B * BYTES_KS / (ms * 1e-3) / 1e9 / peak,
                              "fp64_frac": B * 10 * FLOPS_KS_STEP / (ms * 1e-3) / 1e12 / fp64_peak,
                              "alive": all(int((k.status != 0).sum()) == 0 for k in pool)}
    del pool
    torch.cuda.empty_cache()
    if only == 'c3':
        return res
    # ---- C4: Burgers DNS N=1024 x 512, 500 steps per launch, u / v / Ek history rows every step
    B, n, steps = 512, 1024, 500
    dns = Burger(L=L_DOM, N=n, dt=DT, nu=NU, nsteps=steps, case="turbulence", seed=100 + np.arange(B) % 4, nenvs=B, history=True,
                 device=device)

    def dns_run():
        dns.IC(case="turbulence", on_device=True)
        dns.step_n(None, steps, want_state=False, want_reward=False)
    dns_run()
    torch.cuda.synchronize()
    best = None
    for _ in range(2):
        dns.IC(case="turbulence", on_device=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dns.step_n(None, steps, want_state=False, want_reward=False)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1)
        best = t if best is None else min(best, t)
    hist_bytes = B * steps * (n * 8 + n * 8 + (n // 2 + 1) * 8)
    res["c4_dns_n1024_x512"] = {"launch_us": best * 1e3, "value": B * steps / (best * 1e-3), "unit": "env-steps/s", "n_sub": steps,
                                "history": "uu f64 + vv complex64 + Ek_ktt f64 rows every step",
                                "hbm_frac": hist_bytes / (best * 1e-3) / 1e9 / peak,
                                "fp64_frac": B * steps * FLOPS_DNS_STEP / (best * 1e-3) / 1e12 / fp64_peak,
                                "alive": int((dns.status != 0).sum()) == 0}
    del dns
    torch.cuda.empty_cache()
    # ---- C5 per GPU: MARL Burgers N=32 x 8192, 32 agents, MSE reward, 4-lane teams
    B = 8192
    pool = [make_batch_c5(torch, device, 42 + i, B=B) for i in range(12)]
    a5 = torch.as_tensor(rng.uniform(0.0, 0.02, (B, M)), device=device)

    ms = time_graph(torch, [(lambda e=e: e.step_n(a5, NSUB)) for e in pool], chains=OTHER_CHAINS)
    res["c5_marl_n32_x8192_per_gpu"] = {"launch_us": ms * 1e3, "batches_in_flight": OTHER_CHAINS, "value": B * NSUB / (ms * 1e-3), "unit": "env-steps/s", "n_sub": NSUB,
                                        "hbm_frac": B * BYTES_C5 / (ms * 1e-3) / 1e9 / peak,
                                        "fp64_frac": B * NSUB * FLOPS_PER_ENV_STEP / (ms * 1e-3) / 1e12 / fp64_peak,
                                        "alive": all(int((e.status != 0).sum()) == 0 for e in pool)}
    del pool
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=240)
    ap.add_argument("--warmup", type=int, default=24)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pool", type=int, default=POOL)
    ap.add_argument("--depth", type=int, default=8, help="e2e: independent batches in flight")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--quick", action="store_true", help="skip the sweep / other-config / single-chain extras")
    ap.add_argument("--fused-single", action="store_true", help="diagnostic: bind the fused gather on one GPU")
    ap.add_argument("--chains", type=int, default=0, help="independent batches in flight inside the replayed graph (0 = auto: 4, or 10 for a run of K <= pool steps)")
    ap.add_argument("--lanes", type=int, default=0, help="lanes per environment (0 = library default)")
    ap.add_argument("--gather", default="learner", choices=["all", "learner"],
                    help="multi-GPU: rows gathered to every rank (all-gather) or to rank 0 only (SURVEY 8e: gather semantics suffice)")
    ap.add_argument("--rewards-only", action="store_true", help="multi-GPU: gather only the rewards (configs[4] wording)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"], help="c5: BASELINE configs[4] per GPU (MARL, 8192 envs)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
