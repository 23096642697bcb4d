set -x
mkdir -p gpurun_out/r4e
timeout 300 python -m pytest tests/test_general_band.py -m gpu -q 2>&1 | tail -3
timeout 120 python profiles/scripts/genband_probe.py 8 4 > gpurun_out/r4e/probe.jsonl
for v in gen_r80 gen_r64 gen_b128r128; do GYMWIPE_B200_LIB=$PWD/gymwipe_b200/lib/variants/lib_$v.so timeout 120 python profiles/scripts/genband_probe.py 8 4 >> gpurun_out/r4e/probe.jsonl; done
timeout 120 python profiles/scripts/genband_probe.py 3 0 >> gpurun_out/r4e/probe.jsonl
timeout 120 python profiles/scripts/genband_probe.py 8 16 >> gpurun_out/r4e/probe.jsonl
timeout 120 python profiles/scripts/genband_probe.py 8 4 262144 >> gpurun_out/r4e/probe.jsonl
cat gpurun_out/r4e/probe.jsonl
timeout 600 ncu --set full --import-source on --clock-control none -k regex:genband_step_kernel --launch-skip 12 --launch-count 1 -f -o gpurun_out/r4e/genband python profiles/scripts/genband_probe.py 8 4 > gpurun_out/r4e/ncu.log 2>&1; tail -2 gpurun_out/r4e/ncu.log
