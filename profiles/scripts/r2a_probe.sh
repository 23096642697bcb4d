set -x
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python profiles/scripts/cfg3_probe.py 24 modeR > gpurun_out/r2a/probe_product.json 2> gpurun_out/r2a/probe_product.err
GYMWIPE_B200_LIB=$PWD/gymwipe_b200/lib/variants/lib_noscan.so python profiles/scripts/cfg3_probe.py 24 > gpurun_out/r2a/probe_noscan.json 2> gpurun_out/r2a/probe_noscan.err
cat gpurun_out/r2a/*.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 12 --launch-count 2 -o gpurun_out/r2a/cfg3_step_kernel python profiles/scripts/cfg3_probe.py 12 > gpurun_out/r2a/ncu.log 2>&1
tail -5 gpurun_out/r2a/ncu.log
ls -la gpurun_out/r2a
