"""Ground-truth tables for many distinct shifts: device sampling of the spline (mpde_eval_spline_table) vs the host loop."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from marlpde_b200.hostmath import TruthInterpolant
L, Nd, rows, N = 2 * np.pi, 512, 5001, 32
rng = np.random.default_rng(0)
xd = np.linspace(0, L, Nd, endpoint=False)
tt = np.concatenate(([0.], np.cumsum(np.full(rows - 1, 1e-3))))
uu = 1.0 + np.sin(2 * xd[None, :] + 3 * tt[:, None]) + 0.05 * rng.normal(size=(rows, Nd))
x = np.linspace(0, L, N, endpoint=False)
f = TruthInterpolant(xd, tt, uu, kind="cubic")
t0 = time.perf_counter(); f._spline(); t_fit = time.perf_counter() - t0
def grids(n):
    g = x[None, :] + rng.normal(0, 0.4, n)[:, None]
    g[g > L] -= L; g[g < 0] += L
    return g
dev = torch.device("cuda", 0)
f.rows_device(grids(4), tt, dev, torch.float64); torch.cuda.synchronize()
for n in (64, 1024):
    g = grids(n)
    t0 = time.perf_counter(); out = f.rows_device(g, tt, dev, torch.float64); torch.cuda.synchronize(); td = time.perf_counter() - t0
    m = min(n, 8)
    t0 = time.perf_counter(); host = np.stack([f.rows(g[i], tt) for i in range(m)]); th = (time.perf_counter() - t0) / m * n
    err = float(np.max(np.abs(out[:m].cpu().numpy() - host)))
    print(f"{n:5d} shifts x {rows} rows x {N} points: device {td * 1e3:8.2f} ms, host loop {th * 1e3:9.1f} ms (extrapolated from {m}), "
          f"max |diff| {err:.1e}; host spline fit once: {t_fit * 1e3:.0f} ms", flush=True)
