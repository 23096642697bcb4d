set -x
mkdir -p gpurun_out/r2s
# steady-state mode-R step kernel: full capture with source
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 260 --launch-count 2 -f -o gpurun_out/r2s/modeR_steady python profiles/scripts/profile_steady.py 130 > gpurun_out/r2s/ncu_modeR.log 2>&1; tail -3 gpurun_out/r2s/ncu_modeR.log
# configs[2]: the index kernel (launch 2 of 6) and a step kernel in the timed loop
timeout 600 ncu --set full --import-source on --clock-control none -k regex:mask_index_kernel --launch-skip 2 --launch-count 1 -f -o gpurun_out/r2s/mask_index python profiles/scripts/cfg3_profile.py 4 > gpurun_out/r2s/ncu_idx.log 2>&1; tail -3 gpurun_out/r2s/ncu_idx.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 5 --launch-count 2 -f -o gpurun_out/r2s/cfg3_step python profiles/scripts/cfg3_profile.py 4 > gpurun_out/r2s/ncu_cfg3.log 2>&1; tail -3 gpurun_out/r2s/ncu_cfg3.log
ls -la gpurun_out/r2s
