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
        self._handles = (C.c_void_p * len(envs))(*[e._handle for e in envs])
        for e in self.envs[1:]:
            e.share_stats(self.envs[0])                 # one statistics vector per population (gw_share_stats)

    def __len__(self):
        return len(self.envs)

    def reset(self):
        return [e.reset() for e in self.envs]

    def step(self, actions):
        """``actions``: one action (dict of int32 CUDA tensors) per batch; returns the per-batch step tuples."""
        return [e.step(a) for e, a in zip(self.envs, actions)]

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

    def close(self):
        for e in self.envs[1:]:
            e.share_stats(None)
        for e in self.envs:
            e.close()
