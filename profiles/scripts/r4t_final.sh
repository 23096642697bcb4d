set -x
mkdir -p gpurun_out/r4t
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r4t/pytest_gpu.log 2>&1; tail -4 gpurun_out/r4t/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r4t/bench_n1.json 2> gpurun_out/r4t/bench_n1.err; tail -c 200 gpurun_out/r4t/bench_n1.json; tail -2 gpurun_out/r4t/bench_n1.err
