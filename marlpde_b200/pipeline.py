"""Host <-> device pipelining for a learner that keeps several independent environment batches in
flight (the usual asynchronous vectorised-environment pattern).

Each slot owns one environment batch, one CUDA stream and pinned host buffers.  ``submit(k, actions)``
enqueues, on slot k's stream: H2D copy of the actions -> ``step_n`` (one kernel launch) -> D2H copies of
state and reward; ``collect(k)`` waits for that slot only.  While slot k's results travel back over PCIe,
slot k+1's kernel runs and slot k+2's actions travel in -- every batch still sees its own strictly serial
action -> step -> state chain.
"""
import torch


class HostPipeline:
    def __init__(self, envs, n_sub, post_step=None, post_step_on=None):
        self.envs, self.n_sub = list(envs), int(n_sub)
        self.post_step = post_step          # e.g. the all-gather to the learner rank, enqueued after the step
        # post_step_on(k, stream): the same hook for callers that enqueue on an explicit stream -- used instead of
        # post_step(k, None, None) on the pre-bound path, without making the slot's stream current first
        self.post_step_on = post_step_on
        e0 = self.envs[0]
        dev, dt = e0.device, e0.dtype
        B, M, S, A = e0.nenvs, e0.M, e0._state_buf.shape[1], e0._reward_buf.shape[1]
        self.streams = [torch.cuda.Stream(device=dev) for _ in self.envs]
        # whatever set the environments up (IC / reset kernels, table uploads) ran on the caller's current stream: order every
        # slot stream behind it before the first submit
        cur = torch.cuda.current_stream(dev)
        for st_ in self.streams:
            st_.wait_stream(cur)
        self.done = [torch.cuda.Event() for _ in self.envs]
        self.act_host = [torch.empty((B, M), dtype=dt).pin_memory() for _ in self.envs]
        self.act_dev = [torch.empty((B, M), dtype=dt, device=dev) for _ in self.envs]
        # state and reward share ONE pinned buffer [B*S | B*A]: the library then returns both in a single D2H copy
        self.out_host = [torch.empty(B * (S + A), dtype=dt).pin_memory() for _ in self.envs]
        self.state_host = [o[:B * S].view(B, S) for o in self.out_host]
        self.reward_host = [o[B * S:].view(B, A) for o in self.out_host]
        self.h2d_bytes = self.act_host[0].numel() * self.act_host[0].element_size()
        self.d2h_bytes = (self.state_host[0].numel() + self.reward_host[0].numel()) * self.state_host[0].element_size()
        self.pending = [False] * len(self.envs)
        self._fast = [None] * len(self.envs)     # pre-bound step calls (Burger.host_step_closure), built at the first submit

    def submit(self, k, actions_host=None):
        """Start one RL step of batch k with the host-side ``actions`` ([B,M]; None = reuse the pinned buffer)."""
        if actions_host is not None:
            self.act_host[k].copy_(torch.as_tensor(actions_host))
        go = self._fast[k]
        if go is not None:                      # steady state: one pre-bound library call + the event
            go()
            if self.post_step_on is not None:
                self.post_step_on(k, self.streams[k])
            elif self.post_step is not None:
                with torch.cuda.stream(self.streams[k]):
                    self.post_step(k, None, None)
            self.done[k].record(self.streams[k])
            self.pending[k] = True
            return
        env = self.envs[k]
        if hasattr(env, "step_n_host") and (self.post_step is None or getattr(env, "_peer_host_ok", False)):
            # one library call enqueues H2D -> kernel -> D2H on this slot's stream (a cached CUDA graph)
            packed = self.out_host[k] if (env._spec_ref is not None or env._truth_shift is not None) else None
            env.step_n_host(self.act_host[k], self.n_sub, self.state_host[k], self.reward_host[k], stream=self.streams[k],
                            packed_out=packed)
            if packed is not None and hasattr(env, "host_step_closure"):
                self._fast[k] = env.host_step_closure(self.act_host[k], self.n_sub, packed, self.streams[k])
            if self.post_step is not None:      # fused multi-GPU gather: publish / wait behind the step on the same stream
                with torch.cuda.stream(self.streams[k]):
                    self.post_step(k, None, None)
            self.done[k].record(self.streams[k])
        else:
            with torch.cuda.stream(self.streams[k]):
                self.act_dev[k].copy_(self.act_host[k], non_blocking=True)
                st, rw = self.envs[k].step_n(self.act_dev[k], self.n_sub)
                if self.post_step is not None:
                    res = self.post_step(k, st, rw)       # may return the (state, reward) tensors to copy out instead
                    if res is not None:
                        st, rw = res
                self.state_host[k].copy_(st, non_blocking=True)
                if rw is not None:
                    self.reward_host[k].copy_(rw, non_blocking=True)
                self.done[k].record(self.streams[k])
        self.pending[k] = True

    def collect(self, k):
        """Wait for batch k's step; returns (state [B,S], reward [B,A]) pinned host tensors."""
        if self.pending[k]:
            self.done[k].synchronize()
            self.pending[k] = False
        return self.state_host[k], self.reward_host[k]

    def drain(self):
        for k in range(len(self.envs)):
            self.collect(k)
