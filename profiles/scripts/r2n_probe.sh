set -x
mkdir -p gpurun_out/r2n
python -m pytest tests -m gpu -x -q > gpurun_out/r2n/pytest_gpu.log 2>&1; tail -5 gpurun_out/r2n/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2n/bench_k20.json 2> gpurun_out/r2n/bench_k20.err; tail -3 gpurun_out/r2n/bench_k20.err
