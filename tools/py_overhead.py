"""Host-side cost of one step_n call (tiny batch: the kernel itself is ~free)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from marlpde_b200 import Burger
env = Burger(N=32, nsteps=5000, case='sinus', forcing=False, dforce=False, nenvs=2, history=False)
env.setup_basis(32, 'hat')
a = torch.zeros(2, 32, dtype=torch.float64, device='cuda')
for _ in range(100): env.step_n(a, 1)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5000): env.step_n(a, 1)
torch.cuda.synchronize(); print('us per step_n call (B=2):', (time.perf_counter() - t) / 5000 * 1e6)
