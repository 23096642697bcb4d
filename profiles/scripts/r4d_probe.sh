set -x
mkdir -p gpurun_out/r4d
python -m pytest tests/test_general_band.py -m gpu -q 2>&1 | tail -3
python profiles/scripts/genband_probe.py 8 4 > gpurun_out/r4d/probe.jsonl
for v in gen_b128 gen_b32 gen_r80; do GYMWIPE_B200_LIB=$PWD/gymwipe_b200/lib/variants/lib_$v.so python profiles/scripts/genband_probe.py 8 4 >> gpurun_out/r4d/probe.jsonl; done
python profiles/scripts/genband_probe.py 3 0 >> gpurun_out/r4d/probe.jsonl
python profiles/scripts/genband_probe.py 8 16 >> gpurun_out/r4d/probe.jsonl
python profiles/scripts/genband_probe.py 8 4 262144 >> gpurun_out/r4d/probe.jsonl
cat gpurun_out/r4d/probe.jsonl
timeout 600 ncu --set full --import-source on --clock-control none -k regex:genband_step_kernel --launch-skip 12 --launch-count 1 -f -o gpurun_out/r4d/genband python profiles/scripts/genband_probe.py 8 4 > gpurun_out/r4d/ncu.log 2>&1; tail -2 gpurun_out/r4d/ncu.log
