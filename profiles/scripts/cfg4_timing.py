"""configs[3] (4 bands x 4 devices, per-env positions, mode R): device-timed ms per step."""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import bench
r = bench.cfg4_multiband(torch.device("cuda", 0), steps=32)
print({k: v for k, v in r.items() if k != "workload"})
