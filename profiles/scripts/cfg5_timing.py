"""configs[4] (networked inverted pendulum, in-kernel RK4 plant): device-timed ms per step."""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import bench
r = bench.cfg5_pendulum(torch.device("cuda", 0), steps=32)
print({k: v for k, v in r.items() if k != "workload"})
