"""BASELINE config 5 (north-star target): MARL Burgers N=32, 32 per-gridpoint agents, MSE reward, 8192 environments per GPU
sharded over the ranks of a torchrun job, state [B,96] + reward [B,32] gathered to every rank each RL step by the fused
multicast stores.  Prints env-steps/s of the whole job (device timing, max over ranks)."""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
from marlpde_b200 import Burger
from marlpde_b200 import dist as mdist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B_PER, N, A, NSUB = 8192, 32, 32, 10
team = int(os.environ.get("C5_TEAM", "4"))
rng = np.random.default_rng(0)
truth = rng.normal(1.0, 0.3, (5001, N))
def factory(n, ids):
    e = Burger(L=2 * np.pi, N=N, dt=1e-3, nu=0.02, tend=5, case="turbulence", forcing=False, dforce=False, seed=42 + ids % 4,
               version=0, numAgents=A, nenvs=n, history=False, team_lanes=team, device=dev)
    e.setup_basis(32, 'hat'); e.set_truth_table(truth[None])
    return e
pool = 4
sbs = [mdist.ShardedBatch(B_PER * world, factory, transport="fused" if world > 1 else "nccl") for _ in range(pool)]
acts = torch.as_tensor(rng.uniform(0.0, 0.02, (B_PER, 32)), device=dev)
def step(i):
    return sbs[i % pool].step_n(acts, NSUB, async_gather=True)
for i in range(3 * pool): step(i)
for sb in sbs: sb.wait()
if world > 1: dist.barrier()
torch.cuda.synchronize()
K = 400
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(K): step(i)
for sb in sbs: sb.wait()
e1.record()
if world > 1: dist.barrier()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
gs, gr = sbs[(K - 1) % pool].views()
ok = bool(torch.isfinite(gr).all())
if rank == 0:
    us = float(ms) / K * 1e3
    peer = sbs[0]._peer
    print(f"C5 x{world}: {B_PER} envs/GPU, team_lanes={team}, {us:.1f} us per RL step, {B_PER * world * NSUB / us * 1e-3:.2f} Genv-steps/s total, "
          f"gather={'multicast' if peer is not None and peer.multicast else ('unicast' if peer is not None else 'none')}, finite={ok}, "
          f"gathered state {tuple(gs.shape)} reward {tuple(gr.shape)}", flush=True)
for sb in sbs:
    if sb._peer is not None: sb._peer.check(); sb._peer.close()
if world > 1: dist.destroy_process_group()
