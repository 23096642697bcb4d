set -x
mkdir -p gpurun_out/r4m
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r4m/pytest_gpu.log 2>&1; tail -4 gpurun_out/r4m/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r4m/bench_n1.json 2> gpurun_out/r4m/bench_n1.err; tail -c 200 gpurun_out/r4m/bench_n1.json; tail -2 gpurun_out/r4m/bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r4m/bench_reference_arm.json 2> gpurun_out/r4m/bench_reference_arm.err; tail -c 200 gpurun_out/r4m/bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r4m/launches_bench.csv python bench.py --batches 16 --steps 2 --warmup 3 --no-extras > gpurun_out/r4m/ncu_bench.log 2>&1; tail -1 gpurun_out/r4m/ncu_bench.log | cut -c1-200
