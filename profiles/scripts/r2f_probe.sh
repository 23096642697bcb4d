set -x
mkdir -p gpurun_out/r2f
python -m pytest tests/test_gpu_mode_m.py -x -q > gpurun_out/r2f/pytest_mode_m.log 2>&1; tail -3 gpurun_out/r2f/pytest_mode_m.log
python profiles/scripts/cfg3_probe.py 24 > gpurun_out/r2f/probe_product.json 2> gpurun_out/r2f/probe_product.err
for v in noscan nopf long1 long8; do
GYMWIPE_B200_LIB=$PWD/gymwipe_b200/lib/variants/lib_$v.so python profiles/scripts/cfg3_probe.py 24 > gpurun_out/r2f/probe_$v.json 2> gpurun_out/r2f/probe_$v.err
done
cat gpurun_out/r2f/*.json
timeout 600 ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 12 --launch-count 1 -o gpurun_out/r2f/cfg3_step_kernel python profiles/scripts/cfg3_probe.py 12 > gpurun_out/r2f/ncu.log 2>&1
tail -3 gpurun_out/r2f/ncu.log
