"""PCIe floor of the end-to-end step on this box: 1 MB of actions H2D + 1.03 MB of state/reward D2H per RL step, duplex,
several copies in flight, no kernel -- what `e2e` can at best reach (tools/gpu_r2p.sh)."""
import sys, time, torch
sys.path.insert(0, '.')
dev = torch.device('cuda', 0)
nb_in, nb_out, depth, steps = 4096 * 32 * 8, 4096 * 33 * 8, 8, 2000
hin = [torch.empty(nb_in, dtype=torch.uint8).pin_memory() for _ in range(depth)]
hout = [torch.empty(nb_out, dtype=torch.uint8).pin_memory() for _ in range(depth)]
din = [torch.empty(nb_in, dtype=torch.uint8, device=dev) for _ in range(depth)]
dout = [torch.empty(nb_out, dtype=torch.uint8, device=dev) for _ in range(depth)]
streams = [torch.cuda.Stream() for _ in range(depth)]
evs = [torch.cuda.Event() for _ in range(depth)]
for mode in ("h2d", "d2h", "both"):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        k = i % depth
        evs[k].synchronize()
        with torch.cuda.stream(streams[k]):
            if mode in ("h2d", "both"):
                din[k].copy_(hin[k], non_blocking=True)
            if mode in ("d2h", "both"):
                hout[k].copy_(dout[k], non_blocking=True)
            evs[k].record()
    torch.cuda.synchronize()
    us = (time.perf_counter() - t0) / steps * 1e6
    print(f"{mode:5s}: {us:6.2f} us per step -> at best {4096 * 10 / us * 1e6:.3e} env-steps/s", flush=True)
