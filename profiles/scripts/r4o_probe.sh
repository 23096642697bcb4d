set -x
mkdir -p gpurun_out/r4o
timeout 900 python -m pytest tests/test_general_band.py tests/test_construction_cpu.py -m gpu -q 2>&1 | tail -4
timeout 120 python profiles/scripts/genband_probe.py 8 4 65536 reference > gpurun_out/r4o/probe.jsonl
timeout 300 python profiles/scripts/genband_probe.py 8 4 65536 mask_philox >> gpurun_out/r4o/probe.jsonl
timeout 120 python profiles/scripts/genband_probe.py 3 0 65536 reference >> gpurun_out/r4o/probe.jsonl
timeout 120 python profiles/scripts/genband_probe.py 8 16 65536 reference >> gpurun_out/r4o/probe.jsonl
cat gpurun_out/r4o/probe.jsonl
