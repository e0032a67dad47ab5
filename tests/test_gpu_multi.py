"""Shard invariance on real GPUs: a batch sharded over 2 ranks (NCCL all-gather of
state + reward) is bitwise identical to the same batch on one GPU.  Needs >= 2 GPUs (gpurun --gpus 2)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

B, N, M = 64, 32, 32


def _factory(n, ids):
    from marlpde_b200 import Burger
    env = Burger(N=N, dt=1e-3, nu=0.02, tend=1.0, case="turbulence", forcing=True, dforce=False, seed=50 + (ids % 4) * 9,
                 nenvs=n, history=False)
    env.setup_basis(M, "hat")
    ref = np.abs(np.random.default_rng(0).normal(1.0, 0.1, (1001, N // 2))) * 1e-3 + 1e-6
    env.set_spectrum_reference(ref)
    return env


def _acts():
    return torch.as_tensor(np.random.default_rng(1).uniform(0.0, 0.05, (B, M)))


def _worker(rank, ws, port, q, transport="nccl"):
    import torch.distributed as dist
    from marlpde_b200 import dist as mdist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", rank))
    try:
        sb = mdist.ShardedBatch(B, _factory, transport=transport)
        a = _acts().cuda()
        for _ in range(3):
            gs, gr = sb.step_n(a, 10)
        torch.cuda.synchronize()
        if transport == "p2p":
            sb._peer.check()
        q.put((rank, gs.cpu().numpy(), gr.cpu().numpy()))
        if transport == "p2p":
            sb._peer.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("transport", ["nccl", "p2p"])
def test_two_gpu_shards_equal_one_gpu_bitwise(transport):
    import torch.multiprocessing as mp
    ws, port = 2, 29500 + os.getpid() % 400 + (50 if transport == "p2p" else 0)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q, transport)) for r in range(ws)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=300) for _ in range(ws)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    env = _factory(B, np.arange(B))
    a = _acts().cuda()
    for _ in range(3):
        st, rw = env.step_n(a, 10)
    for o in outs:
        assert np.array_equal(o[1], st.cpu().numpy())
        assert np.array_equal(o[2], rw.cpu().numpy())


def test_peer_gather_single_rank_roundtrip():
    """PeerGather degenerates to a local copy + flag handshake on one GPU (runs in the 1-GPU suite)."""
    from marlpde_b200.dist import PeerGather
    pg = PeerGather(1000, torch.float64, "cuda:0")
    x = torch.arange(1000, dtype=torch.float64, device="cuda:0")
    for k in range(3):
        pg.put(x + k)
        pg.wait()
        torch.cuda.synchronize()
        pg.check()
        assert torch.equal(pg.gathered[0], x + k)
    pg.close()
