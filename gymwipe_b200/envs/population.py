"""
A population of env batches on one GPU, stepped as one vector env.

A batch (``CounterTrafficEnv(num_envs=n)``) owns one state allocation and is stepped by one launch of
the fused step kernel; a population is a list of such batches -- the unit a learner with millions of
envs per GPU works with, and the unit ``bench.py`` calls a step.  The device-resident form simply steps
the batches one after the other on the current stream (the launches overlap through programmatic
dependent launch); the host-buffer form goes through ``gw_step_host_compact_many``: every batch reads its
own pinned action buffer and writes its own pinned result buffer in place, and the call synchronises once.
"""
import ctypes as C

import torch

from gymwipe_b200 import _native as N


class EnvPopulation:
    def __init__(self, envs):
        assert len(envs) >= 1 and all(e.device == envs[0].device for e in envs)
        self.envs = list(envs)
        self.device = envs[0].device
        self.num_envs = sum(e.num_envs for e in envs)
        self._lib = N.lib()
        self.streams = 3                                # side streams of the device-resident step()
        self._side = None
        self._handles = (C.c_void_p * len(envs))(*[e._handle for e in envs])
        for e in self.envs[1:]:
            e.share_stats(self.envs[0])                 # one statistics vector per population (gw_share_stats)

    def __len__(self):
        return len(self.envs)

    def reset(self):
        return [e.reset() for e in self.envs]

    def step(self, actions):
        """
        One ``env.step`` of every batch; ``actions``: one action (dict of int32 CUDA tensors) per batch; returns
        the per-batch step tuples.  The batches are independent, so their launches are spread over
        ``self.streams`` side streams forked from / joined into the current stream (events; under CUDA-graph
        capture: parallel branches): the tail of one batch's kernel overlaps the next batch's -- a launch of
        65,536 envs is 0.86 waves of the step kernel, two in flight keep the SMs' issue slots busier (measured
        +19 % steady-state throughput, ``profiles/README.md``).  From the caller's point of view everything
        is ordered on the current stream.
        """
        S = min(self.streams, len(self.envs))
        if S <= 1:
            return [e.step(a) for e, a in zip(self.envs, actions)]
        if self._side is None or len(self._side) != S:
            self._side = [torch.cuda.Stream(device=self.device) for _ in range(S)]
        cur = torch.cuda.current_stream(self.device)
        fork = torch.cuda.Event()
        fork.record(cur)
        for st in self._side:
            st.wait_event(fork)
        out = [None] * len(self.envs)
        for k, st in enumerate(self._side):
            with torch.cuda.stream(st):
                for b in range(k, len(self.envs), S):
                    out[b] = self.envs[b].step(actions[b])
            done = torch.cuda.Event()
            done.record(st)
            cur.wait_event(done)
        return out

    def stats(self, clear=True, out=None):
        return self.envs[0].stats(clear=clear, out=out)

    def check(self):
        for e in self.envs:
            e.check()

    @staticmethod
    def pointer_array(buffers):
        """ctypes array of the data pointers of pinned tensors / arrays (build once per buffer set)."""
        ptrs = [b.data_ptr() if torch.is_tensor(b) else b.ctypes.data for b in buffers]
        return (C.c_void_p * len(ptrs))(*ptrs)

    def step_host_compact(self, action_ptrs, result_ptrs):
        """One ``env.step`` of every batch from HOST buffers (``gw_step_host_compact_many``): ``action_ptrs`` /
        ``result_ptrs`` are :meth:`pointer_array` s of pinned uint8 ``[n, 2]`` action and uint32 ``[n]`` result
        buffers, one per batch.  Synchronises once; afterwards every result buffer holds its batch's packed
        words (``CounterTrafficEnv.unpack_compact``)."""
        rc = self._lib.gw_step_host_compact_many(self._handles, len(self.envs), action_ptrs, result_ptrs,
                                                 torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            N.check(rc)

    def step_host_tiny(self, action_ptrs, result_ptrs):
        """:meth:`step_host_compact` in the smallest wire format (``gw_step_host_tiny_many``): pinned uint8 ``[n]``
        action bytes (``device << 7 | duration``) and uint16 ``[n]`` result words per batch
        (``CounterTrafficEnv.unpack_tiny``)."""
        rc = self._lib.gw_step_host_tiny_many(self._handles, len(self.envs), action_ptrs, result_ptrs,
                                              torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            N.check(rc)

    def close(self):
        for e in self.envs[1:]:
            e.share_stats(None)
        for e in self.envs:
            e.close()
