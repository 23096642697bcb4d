"""
Multi-GPU plumbing: env batches shard across the GPUs of one box -- every GPU steps an
independent, contiguous slice of the global env-id range with NO data-path collective (the
reference cannot even host two envs in one process).  The only exchange
is the reduction of the per-step reward / episode statistics that feed the learner
(``agents/dqn_counter_traffic.py:70``): an all-reduce(sum) of the 64-byte vector the step
kernel's epilogue (K5) produces, issued on a side stream so that it never blocks stepping.

RNG keys (mode M masks) use GLOBAL env ids (``env_id_offset``), so results do not depend on
the number of GPUs.
"""
import os

import torch
import torch.distributed as dist


def shard_range(total_envs, rank, world_size):
    """Contiguous global env-id range ``[begin, end)`` of ``rank``; sizes differ by at most 1."""
    base, rem = divmod(int(total_envs), int(world_size))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def bind_to_gpu_numa(device_index):
    """
    Restricts this process to the CPUs next to its GPU (NVML's CPU affinity of the device, intersected with the
    CPUs the process may use): host buffers pinned AFTER the call are first touched -- hence allocated -- on the
    GPU's own NUMA node, so the kernels' in-place reads / writes of host memory (``gw_step_host_compact*``) do not
    cross the socket interconnect.  Matters when several ranks of one box drive the host link at once.  Returns the
    CPU list, or None if NVML / the affinity is unavailable (nothing is changed then).  ``GYMWIPE_B200_NO_BIND=1``
    switches it off.
    """
    if os.environ.get("GYMWIPE_B200_NO_BIND") or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + uuid)
        ncpu = os.cpu_count() or 1
        words = (ncpu + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        near = {c for c in range(ncpu) if (mask[c // 64] >> (c % 64)) & 1}
        cpus = sorted(near & os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def init_from_env(backend=None):
    """Initialise ``torch.distributed`` from torchrun's environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kwargs["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local_rank


class StatsReducer:
    """
    Sums the statistics vector over all ranks without stalling the step loop: ``submit`` copies
    the rank-local vector into a slot of a small ring on the caller's stream, records an event,
    and launches the all-reduce on a side stream; ``result`` waits for the oldest outstanding one.
    Works with CUDA tensors (NCCL) and CPU tensors (gloo, used by the CPU tests).
    """

    def __init__(self, device, width=8, depth=4, group=None):
        self.device = torch.device(device)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buf = torch.zeros((depth, width), dtype=torch.float64, device=self.device)
        self.depth = depth
        self.head = 0
        self.pending = []
        self.cuda = self.device.type == "cuda"
        self.side = torch.cuda.Stream(device=self.device) if self.cuda else None

    def next_slot(self):
        """The tensor the next statistics vector should be written into (no extra copy)."""
        return self.buf[self.head % self.depth]

    def submit(self, local_stats=None):
        """Queue the all-reduce of ``local_stats`` (or of the slot from :meth:`next_slot`)."""
        slot = self.buf[self.head % self.depth]
        self.head += 1
        if local_stats is not None and local_stats.data_ptr() != slot.data_ptr():
            slot.copy_(local_stats, non_blocking=True)
        work, ev = None, None
        if self.world > 1:
            if self.cuda:
                ready = torch.cuda.Event()
                ready.record(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(self.side):
                    self.side.wait_event(ready)
                    work = dist.all_reduce(slot, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                    ev = torch.cuda.Event()
                    ev.record(self.side)
            else:
                work = dist.all_reduce(slot, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self.pending.append((slot, work, ev))
        if len(self.pending) >= self.depth:
            return self.result()
        return None

    def result(self):
        """Global sums of the oldest outstanding submission (None if nothing is pending)."""
        if not self.pending:
            return None
        slot, work, ev = self.pending.pop(0)
        if work is not None:
            work.wait()
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)
        return slot.clone()

    def drain(self):
        out = []
        while self.pending:
            out.append(self.result())
        return out
