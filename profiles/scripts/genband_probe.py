"""General band engine (genband_step_kernel): ms per step of bench.general_band for the library named by
GYMWIPE_B200_LIB (launch-configuration variants), optionally other sender / interferer counts."""
import json, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import bench
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nj = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
mode = sys.argv[4] if len(sys.argv) > 4 else "reference"
r = bench.general_band(torch.device("cuda", 0), n=n, ns=ns, nj=nj, mode=mode, with_mode_m=False)
print(json.dumps({"lib": os.path.basename(os.environ.get("GYMWIPE_B200_LIB", "default")), "ns": ns, "nj": nj, "n": n, "mode": mode,
                  "ms_per_step": r["ms_per_step"], "env_steps_per_s": r["env_steps_per_s"],
                  "tx_per_env_step": r["transmissions_per_env_step"], "deliveries_per_env_step": r["deliveries_per_env_step"]}))
