"""Throughput of a population step when the (independent) batches are spread over S streams: device-timed us per
65,536-env launch, steady state and productive regime, CUDA graphs with S parallel chains."""
import json, os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import gymwipe_b200

M = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = 65536
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
rows = 509
a_dev = torch.randint(0, 2, (rows, n), generator=g, device=dev, dtype=torch.int32)
a_dur = torch.randint(0, 20, (rows, n), generator=g, device=dev, dtype=torch.int32)
out = {}
for regime in ("productive", "steady"):
    for S in (1, 2, 3, 4, 6, 8):
        envs = [gymwipe_b200.make('CounterTraffic-v0', num_envs=n, device=dev, env_id_offset=b * n, strict=False) for b in range(M)]
        for e in envs:
            e.reset()
        j = 0
        burn = 128 if regime == "steady" else 2
        for t in range(burn):
            for e in envs:
                e.step({"device": a_dev[j % rows], "duration": a_dur[j % rows]}); j += 1
        torch.cuda.synchronize()
        main = torch.cuda.Stream(device=dev)
        side = [torch.cuda.Stream(device=dev) for _ in range(S)]
        R = 8
        with torch.cuda.stream(main):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=main):
                for r in range(R):
                    fork = torch.cuda.Event()
                    fork.record(main)
                    for k in range(S):
                        side[k].wait_event(fork)
                    for b, e in enumerate(envs):
                        with torch.cuda.stream(side[b % S]):
                            e.step({"device": a_dev[j % rows], "duration": a_dur[j % rows]}); j += 1
                    for k in range(S):
                        ev = torch.cuda.Event()
                        ev.record(side[k])
                        main.wait_event(ev)
            if regime == "steady":
                gr.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main)
            gr.replay()
            e1.record(main)
            torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / (R * M)
        out["%s_S%d" % (regime, S)] = {"us_per_launch": us, "env_steps_per_s": n / (us * 1e-6)}
        for e in envs:
            e.check(); e.close()
        del envs, gr
        torch.cuda.empty_cache()
print(json.dumps(out))
