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
    if transport == "fused-ipc":            # force the CUDA-IPC / unicast path (the default tries NVSwitch multicast)
        os.environ["MPDE_PEER_BACKEND"] = "ipc"
        transport = "fused"
    if transport.endswith("-rows"):         # whole-row state stores staged through shared memory (the default from 6 ranks up)
        os.environ["MPDE_PEER_ROW_STORES"] = "1"
        transport = transport[:-5]
    from marlpde_b200 import dist as mdist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", rank))
    try:
        gather_to = None
        if transport == "fused-learner":    # gather to rank 0 only: the other ranks publish but neither receive nor wait
            transport, gather_to = "fused", 0
        sb = mdist.ShardedBatch(B, _factory, transport=transport, gather_to=gather_to)
        a = _acts().cuda()
        for _ in range(3):
            gs, gr = sb.step_n(a, 10)
        torch.cuda.synchronize()
        if transport in ("p2p", "fused"):
            sb._peer.check()
            print(f"[rank {rank}] transport={transport} backend={sb._peer.backend} multicast={sb._peer.multicast}", flush=True)
        q.put((rank, gs.cpu().numpy(), gr.cpu().numpy()))
        if transport in ("p2p", "fused"):
            sb._peer.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("transport", ["nccl", "p2p", "fused", "fused-ipc", "fused-learner", "fused-rows", "fused-learner-rows"])
def test_two_gpu_shards_equal_one_gpu_bitwise(transport):
    import torch.multiprocessing as mp
    ws, port = 2, 29500 + os.getpid() % 400 + {"nccl": 0, "p2p": 50, "fused": 100, "fused-ipc": 150, "fused-learner": 250,
                                               "fused-rows": 300, "fused-learner-rows": 350}[transport]
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
        rows = slice(0, B)
        if transport.startswith("fused-learner") and o[0] != 0:      # a non-learner rank only holds its own rows
            rows = slice(o[0] * B // ws, (o[0] + 1) * B // ws)
        assert np.array_equal(o[1][rows], st.cpu().numpy()[rows])
        assert np.array_equal(o[2][rows], rw.cpu().numpy()[rows])


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


def test_fused_gather_single_rank_matches_plain_step():
    """transport="fused" on one GPU: the step kernel writes into the (double-buffered) gather buffer and publishes the
    step itself; results equal the plain step_n of an identical batch, bitwise, over odd and even steps."""
    from marlpde_b200 import dist as mdist
    torch.cuda.set_device(0)
    sb = mdist.ShardedBatch(B, _factory, transport="fused")
    ref = _factory(B, np.arange(B))
    a = _acts().cuda()
    for k in range(5):
        gs, gr = sb.step_n(a, 7)
        st, rw = ref.step_n(a, 7)
        torch.cuda.synchronize()
        sb._peer.check()
        assert torch.equal(gs, st) and torch.equal(gr, rw), k
    assert int(sb._peer._steps_dev[0]) == 5 and int(sb._peer._steps_dev[1]) == 5      # published / awaited
    sb._peer.close()


def test_host_step_graph_replay_matches_plain_chain():
    """mpde_step_host replays a cached CUDA graph (H2D -> kernel -> D2H) after its first call on a stream: same bits
    as the device-buffer path."""
    torch.cuda.set_device(0)
    env, ref = _factory(B, np.arange(B)), _factory(B, np.arange(B))
    a_host = _acts().pin_memory()
    a_dev = a_host.cuda()
    S = env._state_buf.shape[1]
    st_h = torch.empty((B, S), dtype=torch.float64).pin_memory()
    rw_h = torch.empty((B, 1), dtype=torch.float64).pin_memory()
    stream = torch.cuda.Stream()
    l0 = env.launch_count
    for k in range(4):
        env.step_n_host(a_host, 10, st_h, rw_h, stream=stream)
        stream.synchronize()
        st, rw = ref.step_n(a_dev, 10)
        torch.cuda.synchronize()
        assert torch.equal(st_h, st.cpu()) and torch.equal(rw_h, rw.cpu()), k
    assert env.launch_count - l0 == 4


def _marl_factory(n, ids):
    from marlpde_b200 import Burger
    env = Burger(N=N, dt=1e-3, nu=0.02, tend=1.0, case="turbulence", forcing=False, dforce=False, seed=50 + (ids % 4) * 9,
                 version=0, numAgents=N, nenvs=n, history=False)
    env.setup_basis(M, "hat")
    env.set_truth_table(np.random.default_rng(3).normal(1.0, 0.3, (1001, N))[None])
    return env


def _marl_worker(rank, ws, port, q):
    import torch.distributed as dist
    from marlpde_b200 import dist as mdist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", rank))
    try:
        sb = mdist.ShardedBatch(B, _marl_factory, transport="fused")
        a = _acts().cuda()
        for _ in range(3):
            gs, gr = sb.step_n(a, 10)
        torch.cuda.synchronize()
        sb._peer.check()
        q.put((rank, gs.cpu().numpy(), gr.cpu().numpy()))
        sb._peer.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_marl_mse_fused_gather_bitwise():
    """BASELINE config 5 in small: per-gridpoint agents (state windows through the shared-memory gather path, MSE reward per
    agent), sharded over 2 GPUs with the fused gather == the same batch on one GPU, bitwise."""
    import torch.multiprocessing as mp
    ws, port = 2, 29500 + os.getpid() % 400 + 200
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_marl_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=300) for _ in range(ws)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    torch.cuda.set_device(0)
    env = _marl_factory(B, np.arange(B))
    a = _acts().cuda()
    for _ in range(3):
        st, rw = env.step_n(a, 10)
    for o in outs:
        assert o[1].shape == (B, 3 * N) and o[2].shape == (B, N)
        assert np.array_equal(o[1], st.cpu().numpy()) and np.array_equal(o[2], rw.cpu().numpy())


def test_host_pipeline_with_fused_gather_single_rank():
    """mpde_step_host with a fused gather bound (mpde_set_peer_local): the kernel writes this rank's rows into its slab of the
    (double-buffered) gather buffer and the packed D2H copy reads them from there -- same bits as the plain device step, and the
    gathered copy holds the same rows."""
    from marlpde_b200 import dist as mdist
    from marlpde_b200.pipeline import HostPipeline
    torch.cuda.set_device(0)
    sb = mdist.ShardedBatch(B, _factory, transport="fused")
    ref = _factory(B, np.arange(B))

    def post(k, st, rw):
        sb._peer.step += 1
        sb._peer.exchange_next()

    pipe = HostPipeline([sb.env], 10, post_step=post)
    a = _acts()
    a_dev = a.cuda()
    for k in range(5):
        pipe.submit(0, a)
        st_h, rw_h = pipe.collect(0)
        st, rw = ref.step_n(a_dev, 10)
        torch.cuda.synchronize()
        sb._peer.check()
        assert torch.equal(st_h, st.cpu()) and torch.equal(rw_h, rw.cpu()), k
        S = st.shape[1]
        cur = sb._peer.current()[0]
        assert torch.equal(cur[:B * S].view(B, S), st) and torch.equal(cur[B * S:].view(B, 1), rw), k
    sb._peer.close()


def test_peer_wait_timeout_is_reported_not_silent():
    """a peer that never publishes must surface as an error.  The wait is bounded in time, does not advance the
    expected step, later waits give up quickly, and the host sees the flag without synchronising (mapped pinned memory)."""
    import time
    from marlpde_b200.dist import PeerGather
    torch.cuda.set_device(0)
    pg = PeerGather(64, torch.float64, "cuda:0", timeout_s=0.05, copies=2)
    pg.poll()
    pg.wait_next()                          # nobody published step 1
    torch.cuda.synchronize()
    assert int(pg._steps_dev[1]) == 0       # expected step NOT advanced
    with pytest.raises(RuntimeError, match="timed out waiting for rank 0"):
        pg.poll()
    t0 = time.perf_counter()
    for _ in range(5):
        pg.wait_next()
    torch.cuda.synchronize()
    assert time.perf_counter() - t0 < 0.2   # 5 further waits cost ~1 ms each, not 5 timeouts
    with pytest.raises(RuntimeError):
        pg.check()
    pg.close()


def test_fused_gather_keeps_getstate_and_reset_working():
    """with a fused gather bound, nsub == 0 calls (getState at episode reset) write the local buffers only and
    do not flip the parity of the gather copies; multi-episode loops therefore run in fused mode."""
    from marlpde_b200 import dist as mdist
    torch.cuda.set_device(0)
    sb = mdist.ShardedBatch(B, _factory, transport="fused")
    ref = _factory(B, np.arange(B))
    a = _acts().cuda()
    for episode in range(2):
        sb.env.IC(case="turbulence")
        ref.IC(case="turbulence")
        s0 = sb.env.getState()
        assert torch.equal(s0, ref.getState())
        for k in range(3):                  # odd number of steps: the next episode starts on the other parity
            gs, gr = sb.step_n(a, 5)
            st, rw = ref.step_n(a, 5)
            torch.cuda.synchronize()
            assert torch.equal(gs, st) and torch.equal(gr, rw), (episode, k)
    sb._peer.check()
    sb._peer.close()


def test_failed_fused_step_does_not_flip_parity():
    """A step that cannot launch (spectral reward without a reference) must leave the parity counter alone."""
    from marlpde_b200 import Burger, dist as mdist
    from marlpde_b200.dist import PeerGather
    torch.cuda.set_device(0)
    env = Burger(N=N, dt=1e-3, nu=0.02, tend=1.0, case="turbulence", forcing=False, dforce=False, seed=50, nenvs=B, history=False)
    env.setup_basis(M, "hat")
    S = env._state_buf.shape[1]
    pg = PeerGather(B * (S + 1), torch.float64, "cuda:0", copies=2)
    pg.fuse(env, B, S, 1)
    env._lib.mpde_set_reward_mode(env._h, 1)                 # spectral reward, but no reference table bound
    a = _acts().cuda()
    with pytest.raises(RuntimeError, match="spectral reward without"):
        env.step_n_fused(a, 5)
    ref = np.abs(np.random.default_rng(0).normal(1.0, 0.1, (1001, N // 2))) * 1e-3 + 1e-6
    env.set_spectrum_reference(ref)
    env.step_n_fused(a, 5)
    pg.step += 1
    torch.cuda.synchronize()
    # the first successful step wrote copy 0
    cur = pg.current()[0]
    assert torch.isfinite(cur).all() and float(cur[:B * S].abs().max()) > 0
    assert float(pg._all[1].abs().max()) == 0.0
    pg.close()
