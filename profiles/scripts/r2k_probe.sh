set -x
mkdir -p gpurun_out/r2k
python -m pytest tests -m gpu -x -q > gpurun_out/r2k/pytest_gpu.log 2>&1; tail -5 gpurun_out/r2k/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2k/bench_k20.json 2> gpurun_out/r2k/bench_k20.err; tail -3 gpurun_out/r2k/bench_k20.err
python bench.py --steps 20 --warmup 5 --streams 1 --no-extras --no-cpu-baseline > gpurun_out/r2k/bench_k20_s1.json 2> gpurun_out/r2k/bench_k20_s1.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 262 --launch-count 2 -o gpurun_out/r2k/modeR_steady python profiles/scripts/profile_steady.py 130 > gpurun_out/r2k/ncu_steady.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 12 --launch-count 2 -o gpurun_out/r2k/cfg3_step_kernel python profiles/scripts/cfg3_probe.py 12 > gpurun_out/r2k/ncu_cfg3.log 2>&1
python profiles/scripts/kernel_stamps.py steady 16 8 > gpurun_out/r2k/stamps_steady.json 2>/dev/null
python profiles/scripts/kernel_stamps.py productive 16 8 > gpurun_out/r2k/stamps_productive.json 2>/dev/null
