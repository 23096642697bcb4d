"""Driver for ncu captures of grid_run_kernel: a mobile (or static) batch of the reference's benchmark grid, one short
warm-up run and one run of DURATION simulated seconds (the launch to capture: --launch-skip 1 --launch-count 1)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from gymwipe_b200.envs import SendingDeviceGrid
mobile = (sys.argv[1] if len(sys.argv) > 1 else "mobile") == "mobile"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dur = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
grid = SendingDeviceGrid(n, 20, device="cuda:0", mobile=mobile, max_moves=1002, seed=100)
grid.runSimulation(0.02)
torch.cuda.synchronize()
grid.runSimulation(dur)
torch.cuda.synchronize()
print("ok", float(grid.now[0]))
