"""Driver for ncu captures of configs[2] (fed masks + index): one gw_set_masks (mask_index_kernel) and a few steps.

    ncu --set full --import-source on --clock-control none -k regex:'mask_index_kernel|step_kernel' ... python profiles/scripts/cfg3_profile.py
"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import bench
r = bench.cfg3_long_packet(torch.device("cuda", 0), bench.measured_peak()[0], steps=int(sys.argv[1]) if len(sys.argv) > 1 else 4)
print(r["ms_per_step"])
