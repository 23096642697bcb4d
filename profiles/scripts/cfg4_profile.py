"""Driver for ncu captures of configs[3] (4 bands x 4 devices, per-env positions, mode R):
    ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 20 --launch-count 1 ... python profiles/scripts/cfg4_profile.py 24
"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import bench
r = bench.cfg4_multiband(torch.device("cuda", 0), steps=int(sys.argv[1]) if len(sys.argv) > 1 else 24)
print(r["ms_per_step"])
