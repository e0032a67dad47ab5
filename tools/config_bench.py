"""Throughput of the other BASELINE configurations (device-resident inputs, CUDA events, 1 GPU):
C3 KS N=64 x 8192, C4 Burgers DNS N=1024 x 512 with u history, C5 MARL Burgers N=32 x 8192 (A=32, MSE reward)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from marlpde_b200 import Burger, KS
dev = torch.device('cuda', 0)
TWO_PI = 2 * np.pi

def timed(label, fn, reps, units):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{label:74s} {ms * 1e3:9.1f} us per call  {units / (ms * 1e-3):.3e} env-steps/s", flush=True)

rng = np.random.default_rng(0)
# ---- C3
B, N, M = 8192, 64, 64
pool = [KS(L=22, N=N, dt=0.25, nsteps=100000, nenvs=B, u0=rng.normal(0, 1e-3, (B, N)), history=False) for _ in range(4)]
for k in pool: k.setup_basis(M, 'hat')
a = torch.as_tensor(rng.normal(0, 1e-3, (B, M)), device=dev)
i = [0]
def ks_step(n):
    def f():
        pool[i[0] % 4].step_n(a, n, want_reward=False); i[0] += 1
    return f
timed("C3 KS N=64 x 8192, M=64, 1 ETDRK4 step per launch + state", ks_step(1), 200, B * 1)
timed("C3 KS N=64 x 8192, M=64, 10 ETDRK4 steps per launch + state", ks_step(10), 100, B * 10)
del pool
# ---- C5
B, N, A = 8192, 32, 32
seeds = 42 + np.arange(B) % 4
pool = []
truth = rng.normal(1.0, 0.3, (5001, N))
for q in range(4):
    e = Burger(L=TWO_PI, N=N, dt=1e-3, nu=0.02, tend=5, case="turbulence", forcing=False, dforce=False, seed=seeds, version=0,
               numAgents=A, nenvs=B, history=False)
    e.setup_basis(32, 'hat'); e.set_truth_table(truth[None]); pool.append(e)
a5 = torch.as_tensor(rng.uniform(0.0, 0.02, (B, 32)), device=dev)
def c5():
    pool[i[0] % 4].step_n(a5, 10); i[0] += 1
timed("C5 MARL Burgers N=32 x 8192, A=32 agents, MSE reward, 10 steps per launch", c5, 200, B * 10)
import os
os.environ["MPDE_TS"] = "4"
timed("C5 same, 4-lane teams (MPDE_TS=4, shared-memory-transposed FFT)", c5, 200, B * 10)
del os.environ["MPDE_TS"]
del pool
# ---- C4
B, N = 512, 1024
dns = Burger(L=TWO_PI, N=N, dt=1e-3, nu=0.02, nsteps=500, case="turbulence", seed=100 + np.arange(B) % 4, nenvs=B, history=True)
def c4():
    dns.IC(case="turbulence", on_device=True); dns.step_n(None, 500, want_state=False, want_reward=False)
timed("C4 Burgers DNS N=1024 x 512, 500 steps per launch, u/v/Ek history rows every step", c4, 3, B * 500)
dns2 = Burger(L=TWO_PI, N=N, dt=1e-3, nu=0.02, nsteps=500, case="turbulence", seed=100 + np.arange(B) % 4, nenvs=B, history=False)
def c4b():
    dns2.step_n(None, 500, want_state=False, want_reward=False)
timed("C4 same without history", c4b, 3, B * 500)
