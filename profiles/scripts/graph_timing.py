"""Pure device time of the mode-R step kernel: T steps captured in one CUDA graph (no launch gaps,
L2-warm), default scenario.  GYMWIPE_B200_NO_MACRO=1 disables the macro events (A/B)."""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import gymwipe_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
T = 64
env = gymwipe_b200.make('CounterTraffic-v0', num_envs=n, strict=False)
env.reset()
g = torch.Generator(device="cuda").manual_seed(1)
dev = torch.randint(0, 2, (T, n), generator=g, device="cuda", dtype=torch.int32)
dur = torch.randint(0, 20, (T, n), generator=g, device="cuda", dtype=torch.int32)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for rep in range(3):                        # 192 steps: into the degenerate regime
        for t in range(T):
            env.step({"device": dev[t], "duration": dur[t]})
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        for t in range(T):
            env.step({"device": dev[t], "duration": dur[t]})
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(4):
        graph.replay()
    e1.record(s)
    torch.cuda.synchronize()
us = e0.elapsed_time(e1) / (4 * T) * 1e3
print("n=%d  graph-replayed step: %.2f us  -> %.3e env-steps/s (L2-warm, no launch gaps)" % (n, us, n / us * 1e6))
