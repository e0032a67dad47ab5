"""Run the bench workload's step kernel at a chosen batch size / fused-step count a few times (timing + ncu target).
usage: python tools/step_run.py B nsub [pool] [reps]"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import bench

B, nsub = int(sys.argv[1]), int(sys.argv[2])
pool = int(sys.argv[3]) if len(sys.argv) > 3 else max(2, -(-160_000_000 // (B * 1912)))
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device('cuda:0')
spec = bench.spectrum_table()
envs = [bench.make_batch(torch, dev, 7 + i, B=B, team_lanes=4, spec=spec, per_env_seeds=False) for i in range(pool)]
a = torch.from_numpy(np.random.default_rng(B).uniform(0.02, 0.1, (B, bench.M))).to(dev)
for rep in range(reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for e in envs:
        e.step_n(a, nsub)
    e1.record(); torch.cuda.synchronize()
    print(f"rep {rep}: {e0.elapsed_time(e1) / pool * 1e3:.2f} us per launch (B={B}, nsub={nsub}, pool={pool})", flush=True)
