"""Run the C4 workload (Burgers DNS N=1024 x 512, 500 steps per launch with history) a few times: timing + ncu target."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from marlpde_b200 import Burger
B, n, steps = 512, 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 500
hist = (sys.argv[2] != '0') if len(sys.argv) > 2 else True
dns = Burger(L=2 * np.pi, N=n, dt=1e-3, nu=0.02, nsteps=steps, case="turbulence", seed=100 + np.arange(B) % 4, nenvs=B, history=hist, device='cuda:0')
for rep in range(3):
    dns.IC(case="turbulence", on_device=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dns.step_n(None, steps, want_state=False, want_reward=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"rep {rep}: {ms:.3f} ms per {steps} steps, {ms / steps * 1e3:.2f} us per step, {B * steps / (ms * 1e-3):.3e} env-steps/s, alive={int((dns.status == 0).sum())}", flush=True)
