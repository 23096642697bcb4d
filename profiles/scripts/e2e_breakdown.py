"""Where the end-to-end step (gw_step_host_compact, pinned buffers) spends its time."""
import sys, os, time, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import gymwipe_b200

n, T = 65536, 300
env = gymwipe_b200.make('CounterTraffic-v0', num_envs=n, strict=False)
env.reset()
g = torch.Generator(device="cuda").manual_seed(1)
dev = torch.randint(0, 2, (T, n), generator=g, device="cuda", dtype=torch.int32)
dur = torch.randint(0, 20, (T, n), generator=g, device="cuda", dtype=torch.int32)
for t in range(128):
    env.step({"device": dev[t], "duration": dur[t]})
act = torch.stack([dev, dur], dim=2).to(torch.uint8).cpu().pin_memory()
res = torch.empty(n, dtype=torch.int32).pin_memory()
for t in range(20):
    env.step_host_compact(act[t], res)
torch.cuda.synchronize()
t0 = time.perf_counter()
for t in range(T):
    env.step_host_compact(act[t], res)
dt = (time.perf_counter() - t0) / T
print("python step_host_compact: %.1f us/step" % (dt * 1e6))
lib, h = env._lib, env._handle
st = torch.cuda.current_stream().cuda_stream
ptrs = [act[t].data_ptr() for t in range(T)]
rp = res.data_ptr()
t0 = time.perf_counter()
for t in range(T):
    lib.gw_step_host_compact(h, ptrs[t], rp, st)
dt = (time.perf_counter() - t0) / T
print("raw ctypes gw_step_host_compact: %.1f us/step" % (dt * 1e6))
# device-resident step + sync, for comparison (launch + kernel + sync, no host data)
o = [dev[t].contiguous() for t in range(T)]
t0 = time.perf_counter()
for t in range(T):
    env.step({"device": dev[t], "duration": dur[t]})
    torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / T
print("device-resident step + synchronize: %.1f us/step" % (dt * 1e6))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
d_act = act[0].cuda()
tot = 0.0
for t in range(50):
    e0.record(); lib.gw_step_host_compact(h, ptrs[t], rp, st); e1.record(); torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
print("event span around the zero-copy call: %.1f us" % (tot / 50 * 1e3))
