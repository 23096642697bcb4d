"""Counters of the warp-specialised fed-mask step kernel (instrumented build, -DGW_WS_DEBUG) on configs[2]:
scanner cycles (total / waiting for bulk copies / issuing), pieces, idle polls; event-warp iterations."""
import ctypes, json, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import bench, gymwipe_b200
from gymwipe_b200 import _native as N

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
L = N.lib()
L.gw_debug_ws.restype = ctypes.c_int
out = (ctypes.c_ulonglong * 16)()
r = bench.cfg3_long_packet(dev, bench.measured_peak()[0], steps=steps)
torch.cuda.synchronize()
L.gw_debug_ws(out)
v = list(out)
launches = max(1, v[9] // 293)
names = ["scan_cycles", "scan_wait_cycles", "pieces", "idle_polls", "issue_cycles", "ev_iters", "ev_allwait_iters", "ev_cycles",
         "posts", "scanner_warps", "event_warp_rounds", "open_issue", "tma_issue", "open_cons", "wait_plus_compute", "publish"]
d = {n: v[i] for i, n in enumerate(names)}
d["per_scanner_warp"] = {n: v[i] / max(1, v[9]) for i, n in enumerate(names[:5])}
d["per_piece"] = {n: v[i] / max(1, v[2]) for i, n in list(enumerate(names))[11:16]}
d["per_event_warp"] = {n: v[i] / max(1, v[10]) for i, n in list(enumerate(names))[5:9]}
d["ms_per_step"] = r["ms_per_step"]
print(json.dumps(d, indent=1))
