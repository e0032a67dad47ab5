"""world_size = 2 gloo tests (CPU) of the sharding / gather logic used for multi-GPU runs."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from marlpde_b200 import dist as mdist


def test_shard_ranges_partition_the_batch():
    for B, R in [(8, 1), (8, 2), (4096, 8), (65536, 8)]:
        covered = []
        for r in range(R):
            lo, hi = mdist.shard_range(B, r, R)
            covered += list(range(lo, hi))
        assert covered == list(range(B))
    with pytest.raises(ValueError):
        mdist.shard_range(10, 0, 4)
    seeds = np.arange(16) + 42
    assert np.array_equal(mdist.shard(seeds, 16, 1, 4), seeds[4:8])
    assert mdist.shard(3.0, 16, 1, 4) == 3.0
    shared = np.zeros((32, 5))
    assert mdist.shard(shared, 16, 1, 4) is shared


class _FakeEnv:
    """Deterministic stand-in for a GPU batch: state/reward are functions of the GLOBAL env id only."""

    def __init__(self, n, ids):
        self.ids = torch.as_tensor(ids, dtype=torch.float64)

    def step_n(self, a, n=1):
        st = torch.stack([self.ids * 10 + k for k in range(3)], dim=1) + (0 if a is None else a.sum(dim=1, keepdim=True))
        rw = (self.ids * 0.5).unsqueeze(1) * n
        return st, rw


def _worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        B = 8
        sb = mdist.ShardedBatch(B, _FakeEnv)
        acts = torch.arange(B * 2, dtype=torch.float64).reshape(B, 2)
        gs, gr = sb.step_n(acts, n=3)
        q.put((rank, sb.lo, sb.hi, gs.numpy(), gr.numpy()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_matches_single_rank():
    ws, port = 2, 29000 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=120) for _ in range(ws)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = _FakeEnv(8, np.arange(8))
    acts = torch.arange(16, dtype=torch.float64).reshape(8, 2)
    st, rw = single.step_n(acts, n=3)
    assert [(o[1], o[2]) for o in outs] == [(0, 4), (4, 8)]
    for o in outs:                    # every rank holds the full batch in global env order, bit-exact
        assert np.array_equal(o[3], st.numpy())
        assert np.array_equal(o[4], rw.numpy())
