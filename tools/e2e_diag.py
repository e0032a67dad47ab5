"""Where does the end-to-end (host buffers) RL step spend its time?  Run on a GPU box."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench
from marlpde_b200.pipeline import HostPipeline
dev = torch.device('cuda', 0)
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 4
envs = [bench.make_batch(torch, dev, 42 + 16 * i) for i in range(depth)]
pipe = HostPipeline(envs, bench.NSUB)
for k in range(depth):
    pipe.act_host[k].fill_(0.07)
n = 2000
def run(label, body, n=n):
    for i in range(3 * depth): body(i)
    torch.cuda.synchronize(); t = time.perf_counter()
    for i in range(n): body(i)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / n * 1e6
    print(f"{label:60s} {dt:8.2f} us/step   {4096 * 10 / dt * 1e6:.3e} env-steps/s", flush=True)
    return dt
ts = [0.0, 0.0]
def full(i):
    k = i % depth
    t0 = time.perf_counter(); pipe.collect(k); t1 = time.perf_counter(); pipe.submit(k); t2 = time.perf_counter()
    ts[0] += t1 - t0; ts[1] += t2 - t1
run("pipeline submit/collect (as bench)", full)
print(f"   host time in collect {ts[0] / (n + 3 * depth) * 1e6:.2f} us, in submit {ts[1] / (n + 3 * depth) * 1e6:.2f} us")
def submit_only(i):
    pipe.submit(i % depth)
run("submit only, never wait (GPU-side throughput of the chains)", submit_only)
streams = pipe.streams
st_dev = [torch.empty((4096, 32), dtype=torch.float64, device=dev) for _ in range(depth)]
def copies(i):
    k = i % depth
    with torch.cuda.stream(streams[k]):
        pipe.act_dev[k].copy_(pipe.act_host[k], non_blocking=True)
        pipe.state_host[k].copy_(st_dev[k], non_blocking=True)
run("copies only (1 MB H2D + 1 MB D2H per step), multi-stream", copies)
def h2d(i):
    k = i % depth
    with torch.cuda.stream(streams[k]):
        pipe.act_dev[k].copy_(pipe.act_host[k], non_blocking=True)
run("H2D only 1 MB", h2d)
def d2h(i):
    k = i % depth
    with torch.cuda.stream(streams[k]):
        pipe.state_host[k].copy_(st_dev[k], non_blocking=True)
run("D2H only 1 MB", d2h)
a = torch.full((4096, 32), 0.07, dtype=torch.float64, device=dev)
def kern(i):
    k = i % depth
    with torch.cuda.stream(streams[k]):
        envs[k].step_n(a, bench.NSUB)
run("kernel only, multi-stream (device buffers)", kern)
def kern1(i):
    envs[i % depth].step_n(a, bench.NSUB)
run("kernel only, one stream", kern1)
big_h = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); big_d = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
def bigcopy(i): big_d.copy_(big_h, non_blocking=True)
dt = run("64 MB H2D", bigcopy, 50); print(f"   -> {64 * 1.048576 / dt * 1e3:.1f} GB/s")
def bigcopy2(i): big_h.copy_(big_d, non_blocking=True)
dt = run("64 MB D2H", bigcopy2, 50); print(f"   -> {64 * 1.048576 / dt * 1e3:.1f} GB/s")
