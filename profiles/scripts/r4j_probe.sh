set -x
mkdir -p gpurun_out/r4j
timeout 600 python -m pytest tests/test_grid.py tests/test_general_band.py -m gpu -q 2>&1 | tail -3
timeout 900 python profiles/scripts/grid_sweep.py 4096 18944 > gpurun_out/r4j/grid_sweep.jsonl 2>&1
cat gpurun_out/r4j/grid_sweep.jsonl
