"""Run step_n with several nsub values (for `ncu --metrics gpu__time_duration.sum`)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
bench.B_PER_GPU = B
dev = torch.device('cuda', 0)
envs = [bench.make_batch(torch, dev, 42 + i) for i in range(8)]
a = torch.full((B, 32), 0.07, dtype=torch.float64, device=dev)
torch.cuda.synchronize()
for nsub in (0, 1, 2, 10, 20):
    for i in range(8):
        envs[i].step_n(a, nsub)
    torch.cuda.synchronize()
print("done")
