"""The reference's benchmark grid at different numbers of grids per launch (occupancy of grid_run_kernel)."""
import json, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import bench
for n in [int(a) for a in sys.argv[1:]] or [4096, 32768]:
    r = bench.grid_benchmark(torch.device("cuda", 0), n_envs=n)
    print(json.dumps({"n_envs": n, "static_s": r["static"]["device_seconds_per_simulated_second"], "static_grid_s_per_s": r["static"]["grid_seconds_per_second"],
                      "mobile_s": r["mobile"]["device_seconds_per_simulated_second"], "mobile_grid_s_per_s": r["mobile"]["grid_seconds_per_second"],
                      "raises": r["mobile"]["grids_where_the_reference_raises"]}))
