"""Per-kernel start / end times INSIDE a CUDA-graph replay of the bench's launch pattern (gw_debug_stamps:
every step launch records the globaltimer when its first block starts, when its first block passes the grid
dependency -- i.e. the previous launch has completed -- and when its last block ends).  Shows that the kernel
itself fits in the per-launch time the bench reports, and how far programmatic dependent launch overlaps
consecutive launches.

    python profiles/scripts/kernel_stamps.py [steady|productive] [batches] [rounds]
"""
import ctypes as C, json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import gymwipe_b200
from gymwipe_b200 import _native as N

regime = sys.argv[1] if len(sys.argv) > 1 else "steady"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 16
R = int(sys.argv[3]) if len(sys.argv) > 3 else 8
n = 65536
dev = torch.device("cuda", 0)
envs = [gymwipe_b200.make('CounterTraffic-v0', num_envs=n, device=dev, env_id_offset=b * n, strict=False) for b in range(M)]
for e in envs:
    e.reset()
g = torch.Generator(device=dev).manual_seed(1)
rows = 256
a_dev = torch.randint(0, 2, (rows, n), generator=g, device=dev, dtype=torch.int32)
a_dur = torch.randint(0, 20, (rows, n), generator=g, device=dev, dtype=torch.int32)
burn = 128 if regime == "steady" else 2
j = 0
for t in range(burn):
    for e in envs:
        e.step({"device": a_dev[j % rows], "duration": a_dur[j % rows]}); j += 1
torch.cuda.synchronize()
L = N.lib()
stamps = [torch.zeros((R, 4), dtype=torch.int64, device=dev) for _ in range(M)]
init = torch.tensor([-1, -1, 0, 0], dtype=torch.int64, device=dev)            # ~0 = 0xFFFF... as int64 -1


def clear():
    for s in stamps:
        s.copy_(init.expand(R, 4))


stream = torch.cuda.Stream(device=dev)
with torch.cuda.stream(stream):
    for b, e in enumerate(envs):
        N.check(L.gw_debug_stamps(e._handle, stamps[b].data_ptr(), R))
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=stream):
        for r in range(R):
            for e in envs:
                e.step({"device": a_dev[j % rows], "duration": a_dur[j % rows]}); j += 1
    for e in envs:
        N.check(L.gw_debug_stamps(e._handle, None, 0))
    clear()
    gr.replay()                                          # graph upload, untimed
    torch.cuda.synchronize()
    best = None
    for rep in range(5):
        clear()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        gr.replay()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        st = torch.stack(stamps, dim=1).reshape(R * M, 4).cpu().numpy().astype(np.uint64)   # launch order: round-major
        if best is None or ms < best[0]:
            best = (ms, st)
ms, st = best
t0 = st[0, 0]
start = (st[:, 0] - t0).astype(np.float64) / 1e3
dep = (st[:, 1] - t0).astype(np.float64) / 1e3
end = (st[:, 2] - t0).astype(np.float64) / 1e3
nl = R * M
out = {"regime": regime, "batches": M, "launches": nl, "event_timed_us_per_launch": 1e3 * ms / nl,
       "span_first_start_to_last_end_us": float(end[-1]), "span_us_per_launch": float(end[-1]) / nl,
       "kernel_start_to_end_us": {"median": float(np.median(end - start)), "min": float((end - start).min()), "max": float((end - start).max())},
       "work_after_dependency_to_end_us": {"median": float(np.median(end - dep)), "min": float((end - dep).min()), "max": float((end - dep).max())},
       "start_before_previous_end_us (PDL overlap)": {"median": float(np.median(end[:-1] - start[1:])), "min": float((end[:-1] - start[1:]).min())},
       "end_to_end_interval_us": {"median": float(np.median(np.diff(end))), "max": float(np.diff(end).max())},
       "first_12_launches_us": [[round(float(a), 2), round(float(b), 2), round(float(c), 2)] for a, b, c in zip(start[:12], dep[:12], end[:12])]}
print(json.dumps(out))
