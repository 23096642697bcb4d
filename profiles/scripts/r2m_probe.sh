set -x
mkdir -p gpurun_out/r2m
python -m pytest tests -m gpu -x -q > gpurun_out/r2m/pytest_gpu.log 2>&1; tail -5 gpurun_out/r2m/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2m/bench_k20.json 2> gpurun_out/r2m/bench_k20.err; tail -3 gpurun_out/r2m/bench_k20.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2m/bench_ref.json 2> gpurun_out/r2m/bench_ref.err
