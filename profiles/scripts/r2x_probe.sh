set -x
mkdir -p gpurun_out/r2x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2x/pytest_gpu.log 2>&1; tail -4 gpurun_out/r2x/pytest_gpu.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 20 --launch-count 1 -f -o gpurun_out/r2x/cfg4_step python profiles/scripts/cfg4_profile.py 24 > gpurun_out/r2x/ncu_cfg4.log 2>&1; tail -2 gpurun_out/r2x/ncu_cfg4.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 5 --launch-count 1 -f -o gpurun_out/r2x/cfg3_step python profiles/scripts/cfg3_profile.py 4 > gpurun_out/r2x/ncu_cfg3.log 2>&1; tail -2 gpurun_out/r2x/ncu_cfg3.log
