"""Experiment r4s: envs stepped in the order of their assignment duration (GW_STEP_ORDER build, gw_set_order) in the
productive regime -- population of M batches over 3 streams, T steps captured into CUDA graphs like bench.py.
argv: ordered(0/1) [M] [T]"""
import ctypes as C, json, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import gymwipe_b200
from gymwipe_b200 import _native as N
from gymwipe_b200.envs import EnvPopulation
ordered = int(sys.argv[1]) if len(sys.argv) > 1 else 0
M = int(sys.argv[2]) if len(sys.argv) > 2 else 48
T = int(sys.argv[3]) if len(sys.argv) > 3 else 32
steady = int(sys.argv[4]) if len(sys.argv) > 4 else 0
n = 65536
dev_t = torch.device("cuda", 0)
pop = EnvPopulation([gymwipe_b200.make('CounterTraffic-v0', num_envs=n, device=dev_t, env_id_offset=b * n, strict=False) for b in range(M)])
pop.reset()
L = N.lib()
if ordered:
    L.gw_set_order.restype = C.c_int
    L.gw_set_order.argtypes = [C.c_void_p, C.c_void_p]
g = torch.Generator(device=dev_t).manual_seed(1234)
ROWS = 61
a_dev = torch.randint(0, 2, (ROWS, n), generator=g, device=dev_t, dtype=torch.int32)
a_dur = torch.randint(0, 20, (ROWS, n), generator=g, device=dev_t, dtype=torch.int32)
perm = [torch.argsort(a_dur[r], stable=True).to(torch.int32).contiguous() for r in range(ROWS)] if ordered else None
stream = torch.cuda.Stream(device=dev_t)
counter = [0]
def pop_step():
    j = counter[0]
    if ordered:
        for b, e in enumerate(pop.envs):
            L.gw_set_order(e._handle, perm[(j + b) % ROWS].data_ptr())
    pop.step([{"device": a_dev[(j + b) % ROWS], "duration": a_dur[(j + b) % ROWS]} for b in range(M)])
    counter[0] = j + M
def capture(count):
    out = []
    for _ in range(count):
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=stream):
            pop_step()
        out.append(gr)
    return out
def timed(fn):
    torch.cuda.synchronize(dev_t)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); fn(); e1.record(stream)
    torch.cuda.synchronize(dev_t)
    return e0.elapsed_time(e1)
with torch.cuda.stream(stream):
    if steady:
        for _ in range(128):
            pop_step()
    gp = capture(T)
    ms = timed(lambda: [gr.replay() for gr in gp]) / (T * M)
pop.check()
st = pop.stats().cpu()
print(json.dumps({"ordered": ordered, "steady": steady, "M": M, "T": T, "us_per_launch": 1e3 * ms, "env_steps_per_s": n / (ms * 1e-3),
                  "reward_sum": float(st[0]), "deliveries": float(st[1] + st[2])}))
