"""Episode-reset cost: one device hand-off launch for the whole batch vs the per-environment host loop it replaces."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from marlpde_b200 import Burger
from marlpde_b200.burger_environment import _truncated_v0
B, N, Nd = 4096, 32, 512
dns = [Burger(L=2 * np.pi, N=Nd, dt=1e-3, nu=0.02, nsteps=10, case="turbulence", seed=50 + i, history=False) for i in range(4)]
sgs = Burger(L=2 * np.pi, N=N, dt=1e-3, nu=0.02, nsteps=10, case="zero", nenvs=B, history=False)
off = np.random.default_rng(0).normal(0, 0.3, B)
dmap = np.arange(B) % 4
v0 = torch.stack([d.v0.reshape(-1) for d in dns]); k = dns[0].k
def dev():
    sgs.IC_handoff(v0, k, src_map=dmap, offsets=off)
def host():
    sgs.IC(v0=np.stack([_truncated_v0(dns[dmap[e]], off[e], N) for e in range(B)]))
for name, fn, reps in (("device hand-off (IC_handoff)", dev, 20), ("host loop (previous BurgerEnvBatch.reset)", host, 2)):
    fn(); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / reps
    print(f"{name:45s} {dt * 1e3:9.3f} ms per reset of {B} environments  ({B / dt:.3e} resets/s)")
seeds = np.arange(B)
big = Burger(L=2 * np.pi, N=N, dt=1e-3, nu=0.02, nsteps=10, case="zero", seed=seeds, nenvs=B, history=False)
for name, flag, reps in (("turbulence IC on device", True, 10), ("turbulence IC host loop", False, 1)):
    big.IC(case="turbulence", on_device=flag); torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): big.IC(case="turbulence", on_device=flag)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / reps
    print(f"{name:45s} {dt * 1e3:9.3f} ms per reset of {B} environments  ({B / dt:.3e} resets/s)")
