"""configs[2] probe: device-timed ms per 65,536-env step of the mode-M (fed masks) kernel, the same scenario in
mode R, and the statistics of the timed steps.  GYMWIPE_B200_LIB selects a kernel variant."""
import json, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import bench, gymwipe_b200

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 24
dev = torch.device("cuda", 0)
r = bench.cfg3_long_packet(dev, bench.measured_peak()[0], steps=steps)
out = {"lib": os.environ.get("GYMWIPE_B200_LIB", "product"), "mode_m_fed": {k: v for k, v in r.items() if k != "workload"}}
if len(sys.argv) > 2 and sys.argv[2] == "modeR":
    n = 65536
    sc = {"assignment_duration_factor": 10000, "bands": [{"frequency": 2.4e9, "bandwidth": 22e6, "devices": [
        {"role": "sender", "x": 0.0, "y": 2.0, "mult": 1, "payload": 1500, "interval": 0.001, "dest": 1},
        {"role": "sender", "x": 0.0, "y": -2.0, "mult": 3, "payload": 1500, "interval": 0.001, "dest": 0},
        {"role": "rrm", "x": 0.0, "y": 0.0},
        {"role": "jammer", "x": 6.0, "y": 0.0, "interval": 0.05, "delay": 0.003, "power": 0.0, "hdr": 13, "payload": 200}]}]}
    env = gymwipe_b200.make('CounterTraffic-v0', num_envs=n, device=dev, scenario=sc, mode="reference", strict=False)
    env.reset()
    g = torch.Generator(device=dev).manual_seed(7)
    a_dev = torch.randint(0, 2, (steps + 4, n), generator=g, device=dev, dtype=torch.int32)
    a_dur = torch.randint(12, 20, (steps + 4, n), generator=g, device=dev, dtype=torch.int32)
    for t in range(4):
        env.step({"device": a_dev[t], "duration": a_dur[t]})
    env.stats()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(4, steps + 4):
        env.step({"device": a_dev[t], "duration": a_dur[t]})
    e1.record()
    torch.cuda.synchronize()
    st = env.stats().cpu().numpy()
    out["mode_r_same_scenario"] = {"ms_per_step": e0.elapsed_time(e1) / steps, "transmissions_per_step": float(st[6]) / steps,
                                   "deliveries_per_step": float(st[1] + st[2]) / steps}
print(json.dumps(out))
