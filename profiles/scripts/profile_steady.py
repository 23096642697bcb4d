"""Driver for ncu captures of the step kernel without thousands of skipped launches: two batches of
65,536 envs stepped alternately; launches 0..2*BURN-1 bring them to the steady state (or stay in the
productive regime for small BURN), the launches after that are the ones to capture:

    ncu --set full --import-source on --clock-control none -k regex:step_kernel \
        --launch-skip $((2*BURN)) --launch-count 2 -o out python profiles/scripts/profile_steady.py BURN
"""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import gymwipe_b200

burn = int(sys.argv[1]) if len(sys.argv) > 1 else 130
n = 65536
envs = [gymwipe_b200.make('CounterTraffic-v0', num_envs=n, strict=False) for _ in range(2)]
for e in envs:
    e.reset()
g = torch.Generator(device="cuda").manual_seed(1)
T = burn + 8
dev = torch.randint(0, 2, (T, n), generator=g, device="cuda", dtype=torch.int32)
dur = torch.randint(0, 20, (T, n), generator=g, device="cuda", dtype=torch.int32)
for t in range(T):
    for e in envs:
        e.step({"device": dev[t], "duration": dur[t]})
torch.cuda.synchronize()
for e in envs:
    e.check()
print("ok")
