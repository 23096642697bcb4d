set -x
mkdir -p gpurun_out/r2l
python -m pytest tests/test_grid.py tests/test_gpu_dqn.py tests/test_construction_cpu.py -x -q > gpurun_out/r2l/pytest_new.log 2>&1; tail -5 gpurun_out/r2l/pytest_new.log
python - > gpurun_out/r2l/grid_bench.json 2> gpurun_out/r2l/grid_bench.err <<'PY'
import json, torch, bench
print(json.dumps(bench.grid_benchmark(torch.device("cuda", 0))))
PY
cat gpurun_out/r2l/grid_bench.json; tail -3 gpurun_out/r2l/grid_bench.err
