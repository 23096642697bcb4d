"""BER-memo statistics of the step kernel (instrumented build, -DGW_MEMO_STATS): evaluations and
second-level hits per step, and the distinct received-power residues across the batch."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import gymwipe_b200
from gymwipe_b200 import _native as N

n, T = 65536, 400
env = gymwipe_b200.make('CounterTraffic-v0', num_envs=n, strict=False)
env.reset()
g = torch.Generator(device="cuda").manual_seed(1)
dev = torch.randint(0, 2, (T, n), generator=g, device="cuda", dtype=torch.int32)
dur = torch.randint(0, 20, (T, n), generator=g, device="cuda", dtype=torch.int32)
L = N.lib()
out = (ctypes.c_ulonglong * 2)()
prev = (0, 0)
for t in range(T):
    env.step({"device": dev[t], "duration": dur[t]})
    if t % 50 == 49 or t < 3:
        torch.cuda.synchronize()
        L.gw_debug_memo_stats(out)
        print("step %3d: evaluations +%d, second-level hits +%d" % (t, out[0] - prev[0], out[1] - prev[1]))
        prev = (out[0], out[1])
p = env.read_state(1).cpu()
for d in range(3):
    u, c = torch.unique(p[d], return_counts=True)
    print("device", d, "distinct received-power values:", len(u), [(float(x).hex(), int(k)) for x, k in zip(u[:6], c[:6])])
