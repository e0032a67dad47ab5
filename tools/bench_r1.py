#!/usr/bin/env python3
"""Benchmark of the marlpde environment time-stepper hot path (BASELINE.json metric:
env-steps/s, batched Burgers LES N=32 x 4096 envs per GPU, fp64, stochastic forcing,
spectral reward, nIntermediate = 10 solver steps per RL step).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo (CUDA)
  python bench.py --impl reference ...                             # CPU reference arm (numpy port)

One "step" = one RL step of the whole batch = ONE kernel launch: 10 ABCN solver steps with
the actions held fixed + getState + spectral reward (burger_environment.py:148-176).
Prints ONE JSON line (see the driver contract in the task description).
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

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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
# (the kernel's own layout moves a little less: float32 Ek sums, no u_prev row for state version 0 -> 1844 B)
BYTES_PER_ENV_LAUNCH = 8 * (M + 4 * (N + 2) + 6 + N + N + 1)
FLOPS_PER_ENV_STEP = 2600           # algorithmic fp64 flops of one Burgers N=32 solver step
FP64_PEAK_TFLOPS = 33.2             # measured on this pool's B200 with tools/microbench.cu (profiles/r1_microbench_b200.md)
# dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full (profiles/r1_ncu_summary_final.md): the reads
# are the cold-cache state + actions; the 3.6 MB of results are still dirty in the 126 MB L2 when the replay ends
NCU_TRAFFIC_BYTES_PER_LAUNCH = 3.951e6


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


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


# ----------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    """One process = one reference-style environment stepped one solver step per Python call
    (the reference has no batching: burger_environment.py:134-192)."""
    seed, seconds, rl_steps = args
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle.burger_oracle import BurgerOracle, forcing_tables, turbulence_ic
    from oracle.common import grid, spectral_rel_err
    rng = np.random.default_rng(seed)
    o = BurgerOracle(B=1, L=L_DOM, N=N, dt=DT, nu=NU, forcing=True, dforce=False)
    o.setup_basis(M, "hat")
    sd = STABLE_SEEDS[seed % len(STABLE_SEEDS)]
    r1, r2 = forcing_tables(sd, int(TEND / DT))
    o.set_forcing_tables(r1[:, :1], r2[:, :1])
    o.IC(u0=turbulence_ic(grid(L_DOM, N), L_DOM, N, 0.0, sd)[None])
    ref = np.abs(rng.normal(1.0, 0.1, (5001, N // 2))) * 1e-3 + 1e-6
    acts = np.full((1, M), rng.uniform(0.05, 0.1))
    prev, done, t0 = 0.0, 0, time.perf_counter()
    while True:
        for _ in range(NSUB):
            o.step(acts)
        o.state()
        err = spectral_rel_err(ref[min(o.ioutnum, 5000)], o.Ek_ktt_row()[0], N)
        prev = err
        done += 1
        if o.ioutnum >= 4000:
            o.IC(u0=turbulence_ic(grid(L_DOM, N), L_DOM, N, 0.0, sd)[None])
        if (rl_steps and done >= rl_steps) or (not rl_steps and time.perf_counter() - t0 >= seconds):
            break
    return done * NSUB, time.perf_counter() - t0


def cpu_run(seconds=None, rl_steps=None, cores=None):
    cores = cores or len(os.sched_getaffinity(0))
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(i, seconds, rl_steps) for i in range(cores)])
        wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return steps / busy, cores, steps, wall


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: every step = `cores` environments x 1 RL step (10 solver steps), all host cores busy
    cores = len(os.sched_getaffinity(0))
    total_rl = args.steps + args.warmup
    per_proc = max(1, min(total_rl, 400))
    value, cores, steps, wall = cpu_run(rl_steps=per_proc, cores=cores)
    sample = (f"{cores} single-env numpy-port processes (one per host core) x {per_proc} RL steps x {NSUB} solver steps, "
              "Burgers N=32 forcing+eddy action+spectral reward, one step() per Python call as in the reference")
    line = {
        "impl": "reference", "metric": "env_steps_per_s", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * cores * NSUB / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus, transport=None):
    what = (f"Burgers LES N={N} x {B_PER_GPU} envs/GPU (BASELINE configs[1]), fp64, M={M} hat basis, "
            f"eddy-viscosity actions (dforce=False), 3-mode stochastic forcing, spectral reward, "
            f"nIntermediate={NSUB} solver steps per RL step; one step = one RL step of the batch")
    if WORKLOAD == "c5":
        what = (f"MARL Burgers N={N} x {B_PER_GPU} envs/GPU (BASELINE configs[4]), fp64, {N} per-gridpoint agents (state windows "
                f"of 3, one eddy-viscosity action each), MSE reward vs a shared truth table, nIntermediate={NSUB}; 4-lane teams")
    return {"workload": what,
            "envs_per_gpu": B_PER_GPU, "N": N, "M": M, "n_intermediate": NSUB, "global_envs": B_PER_GPU * n_gpus,
            "l2": f"rotating pool of {POOL} independent batches per GPU (state working set > 126 MB L2)",
            "parallelism": (f"env-sharded x{n_gpus}, state+reward gathered to every rank per RL step by {transport or 'peer'} "
                            "stores fused into the step kernel (NVLink / NVSwitch, no NCCL call)") if n_gpus > 1 else "single GPU"}


# ----------------------------------------------------------------------------- GPU arm
WORKLOAD = "c2"            # --workload c5 switches to BASELINE configs[4] per GPU (diagnostic; the bench line is c2)


def make_batch_c5(torch, device, seed0):
    """BASELINE configs[4] per GPU: MARL Burgers N=32, 32 per-gridpoint agents (state windows of 3, one action each),
    MSE reward against a shared truth table, 4-lane teams (the large-batch kernel)."""
    from marlpde_b200 import Burger
    seeds = np.array(STABLE_SEEDS)[(np.arange(B_PER_GPU) + seed0) % len(STABLE_SEEDS)]
    env = Burger(L=L_DOM, N=N, dt=DT, nu=NU, tend=TEND, case="turbulence", forcing=False, dforce=False, seed=seeds, version=0,
                 numAgents=N, nenvs=B_PER_GPU, device=device, history=False, team_lanes=4)
    env.setup_basis(M, "hat")
    env.set_truth_table(np.random.default_rng(seed0).normal(1.0, 0.3, (int(TEND / DT) + 1, N))[None])
    return env


def make_batch(torch, device, seed0):
    if WORKLOAD == "c5":
        return make_batch_c5(torch, device, seed0)
    from marlpde_b200 import Burger
    # forced N=32 LES blows up for most forcing seeds within ~10^3 steps (the reference's own physics);
    # these four stay bounded for a whole episode under a positive eddy viscosity, so every
    # environment stays alive (= does all its arithmetic) during the timed region
    seeds = np.array(STABLE_SEEDS)[(np.arange(B_PER_GPU) + seed0) % len(STABLE_SEEDS)]
    env = Burger(L=L_DOM, N=N, dt=DT, nu=NU, tend=TEND, case="turbulence", forcing=True, dforce=False, seed=seeds,
                 nenvs=B_PER_GPU, device=device, history=False)
    env.setup_basis(M, "hat")
    rng = np.random.default_rng(seed0)
    env.set_spectrum_reference(np.abs(rng.normal(1.0, 0.1, (int(TEND / DT) + 1, N // 2))) * 1e-3 + 1e-6)
    return env


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    K, W = args.steps, args.warmup
    pool = max(1, args.pool)
    envs = [make_batch(torch, device, 42 + 16 * i + 1000 * rank) for i in range(pool)]
    rng = np.random.default_rng(rank)
    # one eddy-viscosity coefficient per environment, replicated over its M actions
    acts_host = torch.from_numpy(np.repeat(rng.uniform(0.05, 0.1, (pool, B_PER_GPU, 1)), M, axis=2).copy()).pin_memory()
    acts = acts_host.to(device)
    S = envs[0]._state_size
    RW = envs[0]._reward_buf.shape[1]
    gathers = []
    fused = world > 1 or args.fused_single      # --fused-single: 1-GPU diagnostic of the fused-gather overheads
    if fused:
        # Learner-side gather FUSED into the step kernel: every rank's kernel stores its state + reward rows straight
        # into every rank's (double-buffered) gather buffer over NVLink and publishes a step flag; a 1-CTA wait kernel
        # is the consumer side.  No NCCL call and no host work per step (marlpde_b200.dist.PeerGather.fuse).
        from marlpde_b200.dist import PeerGather
        for env in envs:
            pg = PeerGather(B_PER_GPU * (S + RW), torch.float64, device, copies=2)
            pg.fuse(env, B_PER_GPU, S, RW, gather_state=not args.rewards_only)
            gathers.append(pg)

    side = torch.cuda.Stream(device=device) if fused else None

    def one_step(i, join=True):
        k = i % pool
        st, rw = envs[k].step_n(acts[k], NSUB)
        if fused and not args.no_wait:
            # consumer side of the gather (all ranks' rows of this step have landed in this rank's buffer): it orders
            # the LEARNER after the step, not the next batch's step kernel, so it runs on a forked stream and the
            # step kernels stay back to back (programmatic dependent launch); joined once per rotation / step
            main = torch.cuda.current_stream()
            gathers[k].step += 1
            side.wait_stream(main)
            with torch.cuda.stream(side):
                gathers[k].exchange_next()  # behind the kernel boundary: publish this rank's rows, wait for the peers' 
            if join:
                main.wait_stream(side)
        return st, rw

    def drain():
        pass

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The pool rotation (one RL step of each of the `pool` batches) is captured once into a CUDA graph and replayed:
    # the launch loop is host-bound otherwise (~14 us of Python per step_n call vs a ~16 us kernel).
    graph, per_graph = None, 0
    rot = pool * (2 if fused else 1)        # steps per graph: both copies of the double-buffered gather when fused
    extra = 1 if fused and not args.no_wait else 0      # signal+wait kernel per step

    def capture(chains, nsteps=None):
        """Graph of one rotation (or of the first `nsteps` steps of it).  chains > 1: independent batches alternate between `chains` streams inside the graph, so
        one batch's tail (and, multi-GPU, its gather stores draining over NVLink) overlaps the next batch's kernel; a single
        chain serialises them (each kernel waits, through programmatic dependent launch, for its COMPLETE predecessor)."""
        l_before = sum(e.launch_count for e in envs)
        g_ = torch.cuda.CUDAGraph()
        cstreams = [torch.cuda.Stream(device=device) for _ in range(chains)] if chains > 1 else []
        with torch.cuda.graph(g_):
            cap = torch.cuda.current_stream()
            for cs in cstreams:
                cs.wait_stream(cap)
            nst = rot if nsteps is None else nsteps
            for i in range(nst):
                if chains > 1:
                    with torch.cuda.stream(cstreams[i % chains]):
                        one_step(i, join=False)
                else:
                    one_step(i, join=(i == nst - 1))
            for cs in cstreams:
                cap.wait_stream(cs)
            if chains > 1 and fused and not args.no_wait:
                cap.wait_stream(side)
        n_k = (sum(e.launch_count for e in envs) - l_before) + extra * nst
        for g in gathers:                   # the capture pass only recorded: no step was published
            g.step -= nst // pool
        torch.cuda.synchronize()
        return g_, n_k

    chains = max(1, args.chains)
    if pool % chains:
        chains = 1
    if args.graph:
        for i in range(rot):                # warm every batch before capture
            one_step(i)
        sync()
        graph, per_graph = capture(chains)
    # single GPU: the K % rot steps that do not fill a rotation get their own (shorter) graph instead of Python launches
    tail_graph, tail_len, per_tail = None, 0, 0
    if args.graph and not fused and K % rot:
        tail_len = K % rot
        tail_graph, per_tail = capture(1, tail_len)

    def run_steps(first, n):
        """n RL steps starting at rotation index `first` (a multiple of pool when the graph is used)."""
        launched = 0
        if graph is not None:
            reps, n = divmod(n, rot)
            for _ in range(reps):
                graph.replay()
            launched += reps * per_graph
            for g in gathers:
                g.step += reps * (rot // pool)
            first += reps * rot
            if tail_graph is not None and n == tail_len and first % pool == 0:
                tail_graph.replay()
                launched += per_tail
                n = 0
        l0 = sum(e.launch_count for e in envs)
        for i in range(n):
            one_step(first + i)
        launched += sum(e.launch_count for e in envs) - l0 + extra * n
        return launched

    sampler = ClockSampler(local) if rank == 0 else None      # covers warm-up + timed + e2e regions
    Wr = -(-W // rot) * rot if graph is not None else W        # whole rotations keep the graph aligned
    run_steps(0, Wr)
    sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    launches = run_steps(Wr, K)
    ev1.record()
    sync()
    ms = ev0.elapsed_time(ev1)
    for g in gathers:
        g.check()
    alive = all(int((e.status != 0).sum()) == 0 for e in envs)

    def new_episode():
        """Untimed: put every batch back at t = 0 (the forced N=32 LES only stays bounded for about one episode)."""
        for e in envs:
            e.IC(case="turbulence")
        sync()

    # extra (not the headline): the same K steps with TWO independent batches in flight inside the graph
    ms2 = None
    if args.graph and chains == 1 and pool % 2 == 0 and K >= rot:
        new_episode()
        graph1, per1 = graph, per_graph
        graph, per_graph = capture(2)
        run_steps(0, rot)
        sync()
        ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev2.record()
        run_steps(0, K - K % rot)
        ev3.record()
        sync()
        ms2 = ev2.elapsed_time(ev3) / (K - K % rot)
        graph, per_graph = graph1, per1
        alive = alive and all(int((e.status != 0).sum()) == 0 for e in envs)
    new_episode()

    # ---- end to end through the public API with HOST buffers --------------------------------
    # Every RL step of every batch: pinned-host actions -> H2D -> step_n (one launch) -> D2H of state and
    # reward -> the host waits for them before that batch gets its next actions.  The learner keeps
    # `depth` independent batches in flight (marlpde_b200.pipeline.HostPipeline) so PCIe transfers of one
    # batch overlap the kernel of another; each batch's own action->state chain stays strictly serial.
    from marlpde_b200.pipeline import HostPipeline
    depth = min(pool, max(1, args.depth))
    drain()

    def gather(k, st, rw):            # N > 1: the fused gather stays part of every step; copy out this step's rows
        g = gathers[k]
        g.step += 1
        g.exchange_next()
        if st is None:                  # host path of the library already copied this rank's rows out
            return None
        mine = g.current()[rank]
        return mine[:B_PER_GPU * S].view(B_PER_GPU, S), mine[B_PER_GPU * S:].view(B_PER_GPU, RW)

    pipe = HostPipeline(envs[:depth], NSUB, post_step=gather if fused else None)
    for k in range(depth):
        pipe.act_host[k].copy_(acts_host[k])
    Ke = max(depth, min(K, 2000))
    checksum = 0.0

    def e2e_round(n):
        nonlocal checksum
        for i in range(n):
            k = i % depth
            st_h, rw_h = pipe.collect(k)              # results of this batch's previous step are on the host
            checksum += float(rw_h[0, 0])             # the host really reads them
            pipe.submit(k)                            # next actions for this batch (already in pinned memory)
        pipe.drain()

    e2e_round(3 * depth)
    sync()
    t0 = time.perf_counter()
    e2e_round(Ke)
    sync()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None

    if world > 1:
        t = torch.tensor([ms, e2e_s * 1e3, ms2 or 0.0], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1]) / 1e3
        ms2 = float(t[2]) if ms2 is not None else None
        ok = torch.tensor([1 if alive else 0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        alive = bool(ok.item())

    if rank == 0:
        total_envs = B_PER_GPU * world
        value = total_envs * NSUB * K / (ms * 1e-3)
        peak, how = peaks()
        per_launch_s = ms * 1e-3 / K
        achieved = B_PER_GPU * BYTES_PER_ENV_LAUNCH / per_launch_s / 1e9
        line = {
            "metric": "env_steps_per_s", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(world, ("multicast (multimem.st, %s memory)" if gathers[0].multicast else "unicast peer (%s memory)")
                                      % gathers[0].backend if gathers else None),
            "e2e": {"value": total_envs * NSUB * Ke / e2e_s, "unit": "env-steps/s",
                    "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
                    "steps": Ke, "batches_in_flight": depth,
                    "note": "per RL step of a batch: pinned host actions -> H2D -> step_n -> D2H state+reward -> host waits; "
                            "independent batches overlap (HostPipeline)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH, "peak_source": how,
                         "kernel": "burgers_warp_kernel<double,32,8,FORCING|ACTIONS,HOT>",
                         "bytes_per_launch": B_PER_GPU * BYTES_PER_ENV_LAUNCH,
                         "launch_us": per_launch_s * 1e6,
                         "note": "algorithmic bytes = 1912 B per env per launch (SURVEY 8d) x 4096 envs; the 10 solver "
                                 "steps fused into one launch keep the state on chip, so the launch is bound by FP64 + "
                                 "shuffle issue and their latencies, not by HBM (see roofline_fp64, DESIGN.md 5, profiles/)"},
            "roofline_fp64": {"bound": "fp64", "achieved": B_PER_GPU * NSUB * FLOPS_PER_ENV_STEP / per_launch_s / 1e12,
                              "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                              "frac": B_PER_GPU * NSUB * FLOPS_PER_ENV_STEP / per_launch_s / 1e12 / FP64_PEAK_TFLOPS,
                              "peak_source": "measured (tools/microbench.cu DFMA loop)",
                              "note": "2.6 kflop per env-step (SURVEY 8d) x 40960 env-steps per launch"},
            "all_envs_alive": alive,
            "two_batches_in_flight": None if ms2 is None else {
                "ms_per_step": ms2, "value": total_envs * NSUB / (ms2 * 1e-3), "unit": "env-steps/s",
                "note": "same steps with two independent batches alternating between two streams inside the replayed graph "
                        "(one batch's tail overlaps the next batch's kernel); not used for value / roofline"},
        }
        if world == 1 and not args.no_cpu:
            v, cores, steps, wall = cpu_run(seconds=args.cpu_seconds)
            line["cpu_baseline"] = {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                                    "sample": f"{cores} single-env numpy-port processes x {args.cpu_seconds:.0f} s of the same "
                                              f"workload (N=32, forcing, eddy action, spectral reward, one step() per call)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pool", type=int, default=POOL)
    ap.add_argument("--depth", type=int, default=8, help="e2e: independent batches in flight")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--fused-single", action="store_true", help="diagnostic: bind the fused gather on one GPU")
    ap.add_argument("--no-wait", action="store_true", help="diagnostic: skip the consumer-side wait kernels")
    ap.add_argument("--no-graph", dest="graph", action="store_false", help="launch every step from Python instead of "
                    "replaying the captured pool rotation")
    ap.add_argument("--chains", type=int, default=1, help="independent batches in flight inside the replayed graph")
    ap.add_argument("--rewards-only", action="store_true", help="multi-GPU: gather only the rewards (configs[4] wording)")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"], help="c5: BASELINE configs[4] per GPU (MARL, 8192 envs)")
    args = ap.parse_args()
    if args.workload == "c5":
        global WORKLOAD, B_PER_GPU, BYTES_PER_ENV_LAUNCH
        WORKLOAD, B_PER_GPU = "c5", 8192
        BYTES_PER_ENV_LAUNCH = 8 * (M + 4 * (N + 2) + 3 * N + N)        # 2368 B
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
