"""Device-timed env-steps/s of the mode-R step kernel versus envs per launch (default scenario, steady
state: 160 untimed steps first; launches replayed from a CUDA graph; two batches alternate so that for
the larger sizes the state does not stay in L2)."""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import gymwipe_b200

for n in (16384, 65536, 131072, 262144, 524288, 1048576, 2097152):
    envs = [gymwipe_b200.make('CounterTraffic-v0', num_envs=n, strict=False) for _ in range(2)]
    g = torch.Generator(device="cuda").manual_seed(1)
    T = 32
    dev = torch.randint(0, 2, (T, n), generator=g, device="cuda", dtype=torch.int32)
    dur = torch.randint(0, 20, (T, n), generator=g, device="cuda", dtype=torch.int32)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for e in envs:
            e.reset()
            for t in range(160):
                e.step({"device": dev[t % T], "duration": dur[t % T]})
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=s):
            for t in range(T):
                envs[t % 2].step({"device": dev[t], "duration": dur[t]})
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(4):
            gr.replay()
        e1.record(s)
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (4 * T) * 1e3
    print("n=%8d  %.2f us/launch  %.3e env-steps/s  HBM-frac(193 B) %.3f" % (n, us, n / us * 1e6, 193 * n / us * 1e6 / 6531.6e9))
    for e in envs:
        e.close()
    del envs, gr
    torch.cuda.empty_cache()
