set -x
mkdir -p gpurun_out/r4i
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r4i/pytest_gpu.log 2>&1; tail -3 gpurun_out/r4i/pytest_gpu.log
timeout 600 python profiles/scripts/grid_sweep.py 4096 > gpurun_out/r4i/grid_sweep.jsonl 2>&1
cat gpurun_out/r4i/grid_sweep.jsonl
timeout 120 python profiles/scripts/genband_probe.py 8 4 65536 reference > gpurun_out/r4i/probe.jsonl
timeout 120 python profiles/scripts/genband_probe.py 8 4 65536 mask_philox >> gpurun_out/r4i/probe.jsonl
cat gpurun_out/r4i/probe.jsonl
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r4i/bench.json 2> gpurun_out/r4i/bench.err; tail -c 300 gpurun_out/r4i/bench.json; tail -3 gpurun_out/r4i/bench.err
