"""Which pair of {H2D copy, step kernel, D2H copy} serialises in the 4-stream end-to-end pipeline?  Run on a GPU box."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench
dev = torch.device('cuda', 0)
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 4
envs = [bench.make_batch(torch, dev, 42 + 16 * i) for i in range(depth)]
streams = [torch.cuda.Stream() for _ in range(depth)]
B, M = 4096, 32
ah = [torch.full((B, M), 0.07, dtype=torch.float64).pin_memory() for _ in range(depth)]
ad = [torch.full((B, M), 0.07, dtype=torch.float64, device=dev) for _ in range(depth)]
sh = [torch.empty((B, 32), dtype=torch.float64).pin_memory() for _ in range(depth)]
big = [torch.empty(1 << 20, dtype=torch.float32, device=dev) for _ in range(depth)]
n = 1500
def run(label, body):
    for i in range(3 * depth): body(i)
    torch.cuda.synchronize(); t = time.perf_counter()
    for i in range(n): body(i)
    t1 = time.perf_counter() - t
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / n * 1e6
    print(f"{label:70s} {dt:7.2f} us/step (issue {t1 / n * 1e6:6.2f})", flush=True)
def mk(h2d, kern, d2h, dummy=False):
    def body(i):
        k = i % depth
        with torch.cuda.stream(streams[k]):
            if h2d: ad[k].copy_(ah[k], non_blocking=True)
            if kern:
                if dummy:
                    for _ in range(2): big[k].mul_(1.0001)
                else:
                    envs[k].step_n(ad[k], 10)
            if d2h: sh[k].copy_(envs[k]._state_buf, non_blocking=True)
    return body
run("kernel only", mk(0, 1, 0))
run("H2D only", mk(1, 0, 0))
run("D2H only", mk(0, 0, 1))
run("H2D + D2H", mk(1, 0, 1))
run("H2D + kernel", mk(1, 1, 0))
run("kernel + D2H", mk(0, 1, 1))
run("H2D + kernel + D2H", mk(1, 1, 1))
run("dummy torch kernel only", mk(0, 1, 0, True))
run("H2D + dummy + D2H", mk(1, 1, 1, True))
# one stream for all copies in each direction + events (explicit 3-stage software pipeline)
s_in, s_k, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
ev_in = [torch.cuda.Event() for _ in range(depth)]; ev_k = [torch.cuda.Event() for _ in range(depth)]; ev_out = [torch.cuda.Event() for _ in range(depth)]
def staged(i):
    k = i % depth
    s_in.wait_event(ev_k[k])            # actions buffer free again
    with torch.cuda.stream(s_in):
        ad[k].copy_(ah[k], non_blocking=True); ev_in[k].record(s_in)
    s_k.wait_event(ev_in[k]); s_k.wait_event(ev_out[k])
    with torch.cuda.stream(s_k):
        envs[k].step_n(ad[k], 10); ev_k[k].record(s_k)
    s_out.wait_event(ev_k[k])
    with torch.cuda.stream(s_out):
        sh[k].copy_(envs[k]._state_buf, non_blocking=True); ev_out[k].record(s_out)
run("3 engine-streams (in / kernel / out) + events", staged)
