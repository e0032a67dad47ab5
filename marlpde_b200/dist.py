"""Environment-level data parallelism: contiguous sharding of the environment axis over the
ranks of a ``torch.distributed`` job (one process per GPU) and the gather of per-environment
state / reward summaries to the learner.

Environments never interact, so the solver needs no collective at all; the only
communication is ONE all-gather per RL step of the ``[B/R, S]`` state and ``[B/R, A]`` reward
slabs (NCCL over NVLink on GPUs, gloo on CPU for the tests).  Rank r owns the global
environments ``[r*B/R, (r+1)*B/R)`` and every per-environment parameter (seed, offset, initial
condition) is a function of the GLOBAL environment id, so gathered results are bitwise
independent of the number of ranks.
"""
import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_global, rank=None, world_size=None):
    """Contiguous block of global environment ids owned by ``rank`` (n_global % world_size == 0)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    if n_global % world_size:
        raise ValueError(f"{n_global} environments do not divide over {world_size} ranks")
    per = n_global // world_size
    return rank * per, (rank + 1) * per


def shard(array, n_global=None, rank=None, world_size=None):
    """Rows of a per-environment array (numpy or tensor, leading axis = global env id) owned by this
    rank; scalars / shared arrays (leading axis != n_global) pass through."""
    if n_global is None:
        n_global = len(array)
    if np.ndim(array) == 0 or len(array) != n_global:
        return array
    lo, hi = shard_range(n_global, rank, world_size)
    return array[lo:hi]


def gather_envs(local, out=None):
    """All-gather a ``[B/R, ...]`` slab into ``[B, ...]`` in global environment order on every rank."""
    rank, ws = world()
    if ws == 1:
        if out is None:
            return local
        out.copy_(local)
        return out
    local = local.contiguous()
    if out is None:
        out = torch.empty((local.shape[0] * ws,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local)
    return out


class _RawDeviceMemory:
    """Expose a raw device pointer to torch through the CUDA array interface (no copy, no ownership)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerGather:
    """All-gather of one fixed-size slab per rank through peer-mapped memory (NVLink stores issued by a small
    CUDA kernel, no NCCL call per step).  See include/marlpde_b200.h, "peer-memory gather".

    put(src)  : enqueue the copy of this rank's slab into every rank's ``gathered[rank]`` + publish the step
    wait()    : make the current stream wait until every rank's slab of the current step has landed here
    gathered  : [world_size, chunk_elems] tensor living in this rank's exported buffer

    Fused mode (``copies=2`` + ``fuse(env, ...)``): the Burgers step kernel itself stores state and reward into every
    rank's buffer and publishes the step (include/marlpde_b200.h, mpde_set_peer_output); ``wait_next()`` is the
    consumer side and ``current()`` the [world_size, chunk_elems] copy the last step wrote (double-buffered).
    """

    def __init__(self, chunk_elems, dtype, device, timeout_s=None, copies=1, backend="auto"):
        import ctypes as C
        from . import _lib as LB
        self._C, self._lib = C, LB.lib()
        self.rank, self.ws = world()
        self.device = torch.device(device)
        item = torch.empty((), dtype=dtype).element_size()
        self.chunk_bytes = (chunk_elems * item + 15) // 16 * 16
        self.chunk_elems, self.dtype = chunk_elems, dtype
        # waits are bounded in time: None -> MPDE_PEER_TIMEOUT_S (30 s); a timed-out wait raises on the next host call
        self.timeout_us = 0 if timeout_s is None else max(1, int(float(timeout_s) * 1e6))
        self.copies = int(copies)
        self._copy_bytes = self.ws * self.chunk_bytes
        gbytes = self.copies * self._copy_bytes
        total = gbytes + 16 * ((self.ws * 8 + 15) // 16)
        # Backing memory: "symm" = torch symmetric memory (plumbing: cuMem + fabric handles), which also maps the
        # buffers of all ranks behind ONE NVSwitch multicast address when the box supports NVLS; "ipc" = cudaMalloc +
        # CUDA IPC handles (unicast peer pointers only).  "auto" tries symm and falls back to ipc on every rank alike.
        import os
        backend = os.environ.get("MPDE_PEER_BACKEND", backend)          # auto | symm | ipc (tests / tuning)
        self.backend, self.multicast_base, self._symm, self.multicast = "ipc", 0, None, False
        self._peer_bases, self._opened = [], []
        if backend in ("auto", "symm") and self.ws > 1:
            ok = 0
            try:
                import torch.distributed._symmetric_memory as symm
                t = symm.empty(total, dtype=torch.uint8, device=self.device)
                hdl = symm.rendezvous(t, dist.group.WORLD)
                ok = 1
            except Exception as e:           # no cuMem/fabric support in this container, old torch, ...
                if backend == "symm":
                    raise
                self._symm_error = repr(e)
            flag = torch.tensor([ok], device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 1:
                t.zero_()
                self._symm = (t, hdl)
                self.backend = "symm"
                self._base = t.data_ptr()
                self._peer_bases = [int(p) for p in hdl.buffer_ptrs]
                assert self._peer_bases[self.rank] == self._base
                self.multicast_base = int(getattr(hdl, "multicast_ptr", 0) or 0)
                mc = torch.tensor([1 if self.multicast_base else 0], device=self.device)
                dist.all_reduce(mc, op=dist.ReduceOp.MIN)
                if int(mc.item()) == 0:
                    self.multicast_base = 0
        if self.backend == "ipc":
            base = C.c_void_p()
            self._check(self._lib.mpde_peer_alloc(total, C.byref(base)))
            self._base = base.value
            handle = (C.c_ubyte * 64)()
            self._check(self._lib.mpde_peer_export(self._base, handle))
            handles = [None] * self.ws
            if self.ws > 1:
                dist.all_gather_object(handles, bytes(handle))
            for r in range(self.ws):
                if r == self.rank:
                    self._peer_bases.append(self._base)
                else:
                    p = C.c_void_p()
                    buf = (C.c_ubyte * 64).from_buffer_copy(handles[r])
                    self._check(self._lib.mpde_peer_open(buf, C.byref(p)))
                    self._peer_bases.append(p.value)
                    self._opened.append(p.value)
        arr = C.c_void_p * self.ws
        self._dst = arr(*[b for b in self._peer_bases])
        self._flags = arr(*[b + gbytes for b in self._peer_bases])
        self._my_flags = self._base + gbytes
        raw = torch.as_tensor(_RawDeviceMemory(self._base, gbytes), device=self.device)
        self._all = raw.view(self.copies, self.ws, self.chunk_bytes).view(dtype)[:, :, :chunk_elems]
        self.gathered = self._all[0]
        self._item = item
        self._steps_dev = torch.zeros(4, dtype=torch.int64, device=self.device)    # [0] published, [1] awaited
        self._steps_ptr, self._expect_ptr = self._steps_dev.data_ptr(), self._steps_dev.data_ptr() + 8
        self._fused = None
        self.learner = None
        self._wait_flags, self._wait_n, self._n_slots = self._my_flags, self.ws, self.ws
        self._counter = torch.zeros(4, dtype=torch.int32, device=self.device)
        # error flag in mapped pinned host memory: the wait kernels write it, the host polls it without synchronising
        ep = C.c_void_p()
        self._check(self._lib.mpde_host_flag_alloc(4, C.byref(ep)))
        self._err_ptr = ep.value
        self._err_host = (C.c_int32 * 4).from_address(ep.value)
        self.step = 0
        if self.ws > 1:
            dist.barrier()

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError("marlpde_b200 peer: " + self._lib.mpde_peer_last_error().decode())

    def _stream(self, stream=None):
        """``stream``: a torch stream (or None = the current stream of this device)."""
        return self._C.c_void_p((stream if stream is not None else torch.cuda.current_stream(self.device)).cuda_stream)

    def put(self, src):
        """Store this rank's slab into every rank's buffer and publish the step.  With ``copies=2`` step s lands in copy
        (s - 1) & 1, so a fast rank's next put never overwrites rows a slower rank's learner is still reading (a rank
        can be at most one step ahead: its wait() needs every peer's flag of the previous step)."""
        assert src.is_cuda and src.is_contiguous() and src.numel() * src.element_size() <= self.chunk_bytes
        self.poll()
        self.step += 1
        nbytes = (src.numel() * src.element_size() + 15) // 16 * 16
        off = ((self.step - 1) & 1) * self._copy_bytes if self.copies == 2 else 0
        self._check(self._lib.mpde_peer_put(src.data_ptr(), nbytes, self._dst, off + self.rank * self.chunk_bytes, self._flags,
                                            self.rank, self.ws, self.step, self._counter.data_ptr(), self._stream()))

    def wait(self):
        self._check(self._lib.mpde_peer_wait(self._my_flags, self.ws, self.step, self._err_ptr, self.timeout_us,
                                             self._stream()))

    def fuse(self, env, n_local, S, A, use_multicast=True, gather_state=True, learner=None, row_stores=None):
        """Let ``env``'s step kernel write its [n_local,S] state and [n_local,A] reward straight into this rank's slab
        of every rank's buffer and publish the step itself (no put kernel, no NCCL).

        ``learner=r``: GATHER instead of all-gather -- every rank stores its rows locally and into rank r's buffer only
        (plain peer stores over NVLink); rank r waits for all ranks, the others only publish.  At 8 GPUs an all-gather
        makes every rank ingest 8 slabs per step, a gather only the learner ("gather-style semantics suffice").

        ``row_stores``: how single-agent state rows are written (``mpde_set_peer_row_stores``): whole 256-byte rows staged
        through shared memory (what the ingress-bound gather of 6 and more ranks wants) or direct 16-byte pieces per lane
        (shorter epilogue; faster while the links are not the bound).  Default: by the number of ranks."""
        if row_stores is None:
            row_stores = self.ws > 4
        rc = self._lib.mpde_set_peer_row_stores(env._h, 1 if row_stores else 0)
        if rc != 0:
            raise RuntimeError("marlpde_b200: " + self._lib.mpde_last_error().decode())
        C = self._C
        assert n_local * (S + A) == self.chunk_elems
        slab = self.rank * self.chunk_bytes
        mine = self._all[0, self.rank]
        env.bind_output(mine[:n_local * S].view(n_local, S), mine[n_local * S:].view(n_local, A))
        self.learner = learner
        if learner is None:
            others = [r for r in range(self.ws) if r != self.rank]
        else:
            others = [learner] if self.rank != learner else []
        arr_o = C.c_void_p * max(1, len(others))
        st = arr_o(*[self._peer_bases[r] + slab for r in others])
        rw = arr_o(*[self._peer_bases[r] + slab + n_local * S * self._item for r in others])
        flag_base = self.copies * self._copy_bytes
        # where this rank publishes its step counter: slot [rank] of every consumer's flag array (always its own too)
        consumers = list(range(self.ws)) if learner is None else sorted({learner, self.rank})
        arr_f = C.c_void_p * len(consumers)
        self._flag_slots = arr_f(*[self._peer_bases[r] + flag_base + 8 * self.rank for r in consumers])
        self._n_slots = len(consumers)
        # what this rank waits for: every rank's slot (all-gather, or the learner of a gather) or only its own
        if learner is None or self.rank == learner:
            self._wait_flags, self._wait_n = self._my_flags, self.ws
        else:
            self._wait_flags, self._wait_n = self._my_flags + 8 * self.rank, 1
        stride = self._copy_bytes // self._item if self.copies == 2 else 0
        mc_state = mc_reward = None
        import os
        if learner is None and self.multicast_base and use_multicast and os.environ.get("MPDE_MULTICAST", "1") != "0":
            # one multimem.st per row reaches every rank (this one included): no per-peer stores at all
            mc_reward = self.multicast_base + slab + n_local * S * self._item
            # gather_state=False: only the rewards travel (BASELINE configs[4]: "allgather of rewards"); the state rows
            # are written into this rank's own slab only
            mc_state = self.multicast_base + slab if gather_state else None
            others = []
        if not gather_state and mc_reward is None:
            raise RuntimeError("gather_state=False needs the multicast path (NVSwitch); use the full gather otherwise")
        self.multicast = mc_reward is not None
        rc = self._lib.mpde_set_peer_output(env._h, len(others), st, rw, stride, mc_state, mc_reward)
        if rc != 0:
            raise RuntimeError("marlpde_b200: " + self._lib.mpde_last_error().decode())
        base = self._base + slab            # this rank's own slab, copy 0: lets step_n_host run with the gather bound
        rc = self._lib.mpde_set_peer_local(env._h, base, base + n_local * S * self._item)
        if rc != 0:
            raise RuntimeError("marlpde_b200: " + self._lib.mpde_last_error().decode())
        # publish / wait side of the fused step (mpde_step_fused: kernel -> publish -> wait as one host call)
        rc = self._lib.mpde_set_peer_sync(env._h, self._flag_slots, self._n_slots, self._steps_dev.data_ptr(), self._wait_flags,
                                          self._wait_n, self._steps_dev[1:].data_ptr(), self._err_ptr, self.timeout_us)
        if rc != 0:
            raise RuntimeError("marlpde_b200: " + self._lib.mpde_last_error().decode())
        env._peer_host_ok = True
        env._peer_gather = self
        self._fused = env
        env._state_at = env._reward_at = -1

    def signal_next(self, stream=None):
        """Enqueue (behind the step kernel in stream order) the publication of one more step to every rank."""
        self._check(self._lib.mpde_peer_signal_next(self._flag_slots, self._n_slots, self._steps_ptr, self._stream(stream)))

    def exchange_next(self):
        """signal_next() + wait_next() as one kernel launch."""
        self._check(self._lib.mpde_peer_exchange_next(self._flag_slots, self._n_slots, self._steps_ptr, self._wait_flags,
                                                      self._wait_n, self._expect_ptr, self._err_ptr, self.timeout_us,
                                                      self._stream()))

    def wait_next(self, stream=None):
        """The stream (default: the current one) waits until every rank has published one more step than the last
        wait_next() saw."""
        self._check(self._lib.mpde_peer_wait_next(self._wait_flags, self._wait_n, self._expect_ptr, self._err_ptr,
                                                  self.timeout_us, self._stream(stream)))

    def current(self):
        """[world_size, chunk_elems] copy written by step number ``self.step`` (1-based host count of fused steps)."""
        return self._all[(self.step - 1) & 1] if self.copies == 2 else self._all[0]

    def poll(self):
        """Raise if a wait kernel has timed out so far (reads the mapped host flag: no synchronisation, ~50 ns)."""
        e = int(self._err_host[0])
        if e:
            raise RuntimeError(f"peer gather timed out waiting for rank {e - 1}: the gathered rows are stale "
                               f"(timeout {self.timeout_us / 1e6 if self.timeout_us else 'MPDE_PEER_TIMEOUT_S'} s)")

    def check(self):
        """poll() after draining the device: the definitive answer for everything enqueued so far."""
        torch.cuda.synchronize(self.device)
        self.poll()

    def close(self):
        if getattr(self, "_base", None):
            torch.cuda.synchronize(self.device)
            if self._fused is not None:
                self._lib.mpde_set_peer_output(self._fused._h, 0, None, None, 0, None, None)
                self._lib.mpde_set_peer_sync(self._fused._h, None, 0, None, None, 0, None, None, 0)
                self._fused._peer_gather = None
                self._fused = None
            if self.ws > 1:
                dist.barrier()
            for p in self._opened:
                self._lib.mpde_peer_close(p)
            if self.backend == "ipc":
                self._lib.mpde_peer_free(self._base)
            self._symm = None
            self._base = None
            if self._err_ptr:
                self._err_host = None
                self._lib.mpde_host_flag_free(self._err_ptr)
                self._err_ptr = None


class ShardedBatch:
    """A batch of ``n_global`` environments split over the ranks of the current process group.

    ``factory(n_local, ids)`` builds this rank's environments from the GLOBAL ids it owns, e.g.
    ``lambda n, ids: Burger(N=32, seed=42 + ids, nenvs=n, ...)``.  ``step_n`` advances the local shard
    with one kernel launch; the kernel writes state and reward into ONE flat send buffer, which a
    single all-gather delivers to every rank (the learner reads ``states`` / ``rewards`` views).
    """

    def __init__(self, n_global, factory, transport="nccl", gather_to=None):
        """``gather_to=r`` (transport "fused" only): rows are gathered to rank r alone (the learner); the other ranks see
        their own rows in ``views()``."""
        self.rank, self.world_size = world()
        self.transport = transport
        self.gather_to = gather_to
        self.n_global = int(n_global)
        self.lo, self.hi = shard_range(self.n_global, self.rank, self.world_size)
        self.ids = np.arange(self.lo, self.hi)
        self.env = factory(self.hi - self.lo, self.ids)
        self._flat = self._gflat = None
        if hasattr(self.env, "bind_output"):
            nl = self.hi - self.lo
            S, A = self.env._state_buf.shape[1], self.env._reward_buf.shape[1]
            self._S, self._A = S, A
            buf = self.env._state_buf
            self._flat = torch.zeros(nl * (S + A), dtype=buf.dtype, device=buf.device)
            self.env.bind_output(self._flat[:nl * S].view(nl, S), self._flat[nl * S:].view(nl, A))
            if transport == "fused":
                # the step kernel writes straight into every rank's (double-buffered) gather buffer
                self._peer = PeerGather(nl * (S + A), buf.dtype, buf.device, copies=2)
                self._peer.fuse(self.env, nl, S, A, learner=gather_to)
                self._gflat = self._peer.gathered
            elif transport == "p2p":
                self._peer = PeerGather(nl * (S + A), buf.dtype, buf.device, copies=2)
                self._gflat = self._peer.gathered
            else:
                self._peer = None
                self._gflat = torch.zeros((self.world_size, nl * (S + A)), dtype=buf.dtype, device=buf.device)
        self._g_state = self._g_reward = None
        self._work = None

    def local(self, per_env):
        return shard(per_env, self.n_global, self.rank, self.world_size)

    def wait(self):
        """Block the current stream until the last asynchronous gather has landed."""
        if self._work == "fused":
            self._peer.wait_next()
            self._work = None
        elif self._work == "p2p":
            self._peer.wait()
            self._work = None
        elif self._work is not None:
            self._work.wait()
            self._work = None

    def views(self):
        """(states [R, B/R, S], rewards [R, B/R, A]) views of the gathered buffer, global env order."""
        nl = self.hi - self.lo
        if self._peer is not None:
            self._peer.poll()                         # a timed-out gather raises here instead of handing out stale rows
        g = self._peer.current() if self._peer is not None else self._gflat
        return g[:, :nl * self._S].view(self.world_size, nl, self._S), g[:, nl * self._S:].view(self.world_size, nl, self._A)

    def step_n(self, actions_global_or_local, n=1, async_gather=False, **kw):
        a = actions_global_or_local
        if a is not None and len(a) == self.n_global and self.world_size > 1:
            a = a[self.lo:self.hi]
        self.wait()                                   # the send buffer is about to be overwritten
        if self.transport == "fused" and self._flat is not None and not kw and self.env._reward_enabled():
            # ONE library call: step kernel (rows stored straight into every rank's buffer) -> publish -> wait, replayed
            # from a CUDA graph cached per parity (mpde_step_fused)
            self._peer.poll()
            self.env.step_n_fused(a, n, async_gather=async_gather)
            self._peer.step += 1
            gs, gr = self.views()
            return gs.reshape(self.n_global, self._S), gr.reshape(self.n_global, self._A)
        st, rw = self.env.step_n(a, n, **kw)
        if self._flat is not None and st is not None and rw is not None:
            if self.transport == "fused":
                self._peer.step += 1
                self._peer.signal_next()
                self._work = "fused"
                if not async_gather:
                    self.wait()
            elif self._peer is not None:
                self._peer.put(self._flat)
                self._work = "p2p"
                if not async_gather:
                    self.wait()
            elif self.world_size == 1:
                self._gflat[0].copy_(self._flat)
            else:
                self._work = dist.all_gather_into_tensor(self._gflat.view(-1), self._flat, async_op=True)
                if not async_gather:
                    self.wait()
            gs, gr = self.views()
            return gs.reshape(self.n_global, self._S), gr.reshape(self.n_global, self._A)
        gs = gr = None
        if st is not None:
            if self._g_state is None or self._g_state.shape[1:] != st.shape[1:]:
                self._g_state = torch.empty((self.n_global,) + tuple(st.shape[1:]), dtype=st.dtype, device=st.device)
            gs = gather_envs(st, self._g_state)
        if rw is not None:
            if self._g_reward is None:
                self._g_reward = torch.empty((self.n_global,) + tuple(rw.shape[1:]), dtype=rw.dtype, device=rw.device)
            gr = gather_envs(rw, self._g_reward)
        return gs, gr
