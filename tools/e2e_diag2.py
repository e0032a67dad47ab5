"""Host-side cost of the end-to-end submit path, piece by piece.  Run on a GPU box."""
import ctypes as C, os, sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench
from marlpde_b200.pipeline import HostPipeline
dev = torch.device('cuda', 0)
depth = 4
envs = [bench.make_batch(torch, dev, 42 + 16 * i) for i in range(depth)]
pipe = HostPipeline(envs, bench.NSUB)
for k in range(depth):
    pipe.act_host[k].fill_(0.07)
lib = envs[0]._lib
n = 2000
def timeit(label, body, n=n, sync_each=False):
    for i in range(3 * depth): body(i)
    torch.cuda.synchronize(); host = 0.0; t00 = time.perf_counter()
    for i in range(n):
        t0 = time.perf_counter(); body(i); host += time.perf_counter() - t0
    t_issue = time.perf_counter() - t00
    torch.cuda.synchronize(); tot = time.perf_counter() - t00
    print(f"{label:66s} host {host / n * 1e6:7.2f} us/call   issue loop {t_issue / n * 1e6:7.2f}   total {tot / n * 1e6:7.2f} us/step", flush=True)
e = envs[0]
h = e._h
def raw_call(i):
    k = i % depth
    ek = envs[k]
    lib.mpde_step_host(ek._h, pipe.act_host[k].data_ptr(), 10, pipe.state_host[k].data_ptr(), pipe.reward_host[k].data_ptr(),
                       pipe.streams[k].cuda_stream)
print("MPDE_HOST_GRAPH =", os.environ.get("MPDE_HOST_GRAPH", "1"))
timeit("raw ctypes mpde_step_host, 4 streams", raw_call)
def raw_call1(i):
    lib.mpde_step_host(h, pipe.act_host[0].data_ptr(), 10, pipe.state_host[0].data_ptr(), pipe.reward_host[0].data_ptr(),
                       pipe.streams[0].cuda_stream)
timeit("raw ctypes mpde_step_host, 1 stream/1 env", raw_call1)
def py_call(i):
    k = i % depth
    envs[k].step_n_host(pipe.act_host[k], 10, pipe.state_host[k], pipe.reward_host[k], stream=pipe.streams[k])
timeit("Burger.step_n_host, 4 streams", py_call)
def submit(i): pipe.submit(i % depth)
timeit("HostPipeline.submit (step_n_host + event record)", submit)
def full(i):
    k = i % depth
    pipe.collect(k); pipe.submit(k)
timeit("collect + submit", full)
a = torch.full((4096, 32), 0.07, dtype=torch.float64, device=dev)
def dev_call(i):
    k = i % depth
    lib.mpde_step(envs[k]._h, a.data_ptr(), 10, envs[k]._state_buf.data_ptr(), envs[k]._reward_buf.data_ptr(), pipe.streams[k].cuda_stream)
timeit("raw ctypes mpde_step (device buffers), 4 streams", dev_call)
def dev_call1(i):
    k = i % depth
    lib.mpde_step(envs[k]._h, a.data_ptr(), 10, envs[k]._state_buf.data_ptr(), envs[k]._reward_buf.data_ptr(), pipe.streams[0].cuda_stream)
timeit("raw ctypes mpde_step (device buffers), 1 stream", dev_call1)
def ev(i): pipe.done[i % depth].record(pipe.streams[i % depth])
timeit("event record only", ev)
