set -x
mkdir -p gpurun_out/r2h
python -m pytest tests/test_gpu_abi.py tests/test_gpu_mode_m.py -x -q > gpurun_out/r2h/pytest_new.log 2>&1; tail -5 gpurun_out/r2h/pytest_new.log
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2h/bench_k20.json 2> gpurun_out/r2h/bench_k20.err; tail -3 gpurun_out/r2h/bench_k20.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 262 --launch-count 2 -o gpurun_out/r2h/modeR_steady python profiles/scripts/profile_steady.py 130 > gpurun_out/r2h/ncu_steady.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 6 --launch-count 2 -o gpurun_out/r2h/modeR_productive python profiles/scripts/profile_steady.py 3 > gpurun_out/r2h/ncu_productive.log 2>&1
tail -2 gpurun_out/r2h/ncu_steady.log gpurun_out/r2h/ncu_productive.log
