set -x
mkdir -p gpurun_out/r4h
timeout 600 python -m pytest tests/test_general_band.py tests/test_grid.py -m gpu -q 2>&1 | tail -5
timeout 600 python profiles/scripts/grid_sweep.py 4096 18944 > gpurun_out/r4h/grid_sweep.jsonl 2>&1
cat gpurun_out/r4h/grid_sweep.jsonl
timeout 120 python profiles/scripts/genband_probe.py 8 4 65536 reference > gpurun_out/r4h/probe.jsonl
timeout 120 python profiles/scripts/genband_probe.py 8 16 65536 reference >> gpurun_out/r4h/probe.jsonl
cat gpurun_out/r4h/probe.jsonl
