"""Device-timed env-steps/s of the mode-R step kernel versus batch size (default scenario)."""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import gymwipe_b200

for n in (16384, 65536, 148 * 512, 262144, 148 * 512 * 4, 1048576, 148 * 512 * 16):
    env = gymwipe_b200.make('CounterTraffic-v0', num_envs=n, strict=False)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    T = 160
    dev = torch.randint(0, 2, (T, n), generator=g, device="cuda", dtype=torch.int32)
    dur = torch.randint(0, 20, (T, n), generator=g, device="cuda", dtype=torch.int32)
    for t in range(8):
        env.step({"device": dev[t], "duration": dur[t]})
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(8, T):
        env.step({"device": dev[t], "duration": dur[t]})
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (T - 8)
    print("n=%8d  ms/step %.4f  env-steps/s %.3e  HBM-frac(193B) %.3f" % (n, ms, n / ms * 1e3, 193 * n / ms * 1e3 / 6531.6e9))
    del env
    torch.cuda.empty_cache()
