"""Kernel time vs number of fused sub-steps (fixed launch cost vs per-sub-step cost)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
bench.B_PER_GPU = B
dev = torch.device('cuda', 0)
envs = [bench.make_batch(torch, dev, 42 + i) for i in range(4)]
a = torch.full((B, 32), 0.07, dtype=torch.float64, device=dev)
for nsub in (0, 1, 2, 5, 10, 20, 40):
    for e in envs:
        e.IC(case='turbulence')
    reps = 40
    for i in range(8): envs[i % 4].step_n(a, nsub)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for i in range(reps): envs[i % 4].step_n(a, nsub)
    ev[1].record(); torch.cuda.synchronize()
    print(f"B={B} nsub={nsub:3d}  {ev[0].elapsed_time(ev[1]) / reps * 1e3:8.2f} us per launch")
