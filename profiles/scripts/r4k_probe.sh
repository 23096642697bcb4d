set -x
mkdir -p gpurun_out/r4k
timeout 900 ncu --set full --import-source on --clock-control none -k regex:grid_run_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/r4k/grid_mobile python profiles/scripts/grid_profile.py mobile 4096 0.05 > gpurun_out/r4k/ncu.log 2>&1; tail -2 gpurun_out/r4k/ncu.log
