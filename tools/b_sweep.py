"""Launch time vs batch size for each team-size variant (bench workload, 10 fused sub-steps, CUDA-graph replay)."""
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
print("B      " + "".join(f"   TS={ts:<2d} us (Genv-steps/s)" for ts in (16, 8, 4)), flush=True)
for B in (512, 2048, 4096, 8192, 16384, 32768):
    bench.B_PER_GPU = B
    pool = 4 if B > 8192 else 8
    envs = [bench.make_batch(torch, dev, 42 + i) for i in range(pool)]
    a = torch.full((B, 32), 0.07, dtype=torch.float64, device=dev)
    row = f"{B:<7d}"
    for ts in (16, 8, 4):
        os.environ["MPDE_TS"] = str(ts)
        for e in envs:
            e.IC(case='turbulence')
            e.step_n(a, 10)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for e in envs:
                e.step_n(a, 10)
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): g.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / (reps * pool) * 1e3
        row += f"   {us:8.2f} ({B * 10 / us * 1e-3:6.2f})   "
        del g
    print(row, flush=True)
    del envs
