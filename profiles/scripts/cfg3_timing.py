"""configs[2] (1500-byte payloads, fed per-bit masks, PHY-only interferer, mode M): device-timed ms per step."""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import bench
r = bench.cfg3_long_packet(torch.device("cuda", 0), steps=24)
print({k: v for k, v in r.items() if k != "workload"})
