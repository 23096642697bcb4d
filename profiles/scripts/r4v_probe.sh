set -x
mkdir -p gpurun_out/r4v
timeout 600 ncu --set full --import-source on --clock-control none -k regex:genband_step_kernel --launch-skip 12 --launch-count 1 -f -o gpurun_out/r4v/genband python profiles/scripts/genband_probe.py 8 4 > gpurun_out/r4v/ncu.log 2>&1; tail -2 gpurun_out/r4v/ncu.log
