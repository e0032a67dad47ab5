"""KS N=64 x 8192 (BASELINE config 3): team-size variants and fused steps per launch."""
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from marlpde_b200 import KS
dev = torch.device('cuda', 0)
rng = np.random.default_rng(0)
B, N, M = 8192, 64, 64
a = torch.as_tensor(rng.normal(0, 1e-3, (B, M)), device=dev)
import itertools
variants = [(32, 1), (16, 1), (16, 6), (16, 8)] if len(sys.argv) < 2 else [(16, int(sys.argv[1]))]
for ts, minb in variants:
    os.environ["MPDE_KS_TS"] = str(ts)
    os.environ["MPDE_KS_MINB"] = str(minb)
    pool = [KS(L=22, N=N, dt=0.25, nsteps=100000, nenvs=B, u0=rng.normal(0, 1e-3, (B, N)), history=False) for _ in range(4)]
    for k in pool: k.setup_basis(M, 'hat')
    for n in (1, 4, 10):
        for k in pool: k.step_n(a, n, want_reward=False)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for k in pool: k.step_n(a, n, want_reward=False)
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): g.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 80 * 1e3
        print(f"KS N=64 x {B}  TS={ts:2d} minb={minb}  {n:2d} steps/launch  {us:8.1f} us  {B * n / us * 1e-3:6.3f} Genv-steps/s", flush=True)
    del pool
