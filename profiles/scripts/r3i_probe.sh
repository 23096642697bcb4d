set -x
mkdir -p gpurun_out/r3i
timeout 900 python profiles/scripts/grid_sweep.py 4096 18944 37888 2>&1 | tail -4
timeout 600 ncu --set full --import-source on --clock-control none -k regex:grid_run_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/r3i/grid python profiles/scripts/grid_sweep.py 4096 > gpurun_out/r3i/ncu.log 2>&1; tail -2 gpurun_out/r3i/ncu.log
