set -x
mkdir -p gpurun_out/r4f
timeout 300 python -m pytest tests/test_general_band.py tests/test_construction_cpu.py -m gpu -q 2>&1 | tail -3
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 20 --launch-count 1 -f -o gpurun_out/r4f/cfg4 python profiles/scripts/cfg4_profile.py 24 > gpurun_out/r4f/ncu_cfg4.log 2>&1; tail -2 gpurun_out/r4f/ncu_cfg4.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 8 --launch-count 1 -f -o gpurun_out/r4f/productive python profiles/scripts/profile_steady.py 3 > gpurun_out/r4f/ncu_prod.log 2>&1; tail -2 gpurun_out/r4f/ncu_prod.log
