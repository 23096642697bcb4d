"""Runs the K3 mask-scan roofline measurement alone (profiling target for ncu)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import bench  # noqa: E402

if __name__ == "__main__":
    order = sys.argv[1] if len(sys.argv) > 1 else "random"
    print(json.dumps(bench.mask_scan_roofline(torch.device("cuda:0"), bench.measured_peak()[0], order)))
