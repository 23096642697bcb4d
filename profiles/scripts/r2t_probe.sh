set -x
mkdir -p gpurun_out/r2t
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 20 --launch-count 1 -f -o gpurun_out/r2t/cfg4_step python profiles/scripts/cfg4_profile.py 24 > gpurun_out/r2t/ncu_cfg4.log 2>&1; tail -3 gpurun_out/r2t/ncu_cfg4.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 12 --launch-count 1 -f -o gpurun_out/r2t/modeR_productive python profiles/scripts/profile_steady.py 6 > gpurun_out/r2t/ncu_prod.log 2>&1; tail -3 gpurun_out/r2t/ncu_prod.log
