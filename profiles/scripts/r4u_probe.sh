set -x
mkdir -p gpurun_out/r4u
timeout 900 python -m pytest tests/test_general_band.py tests/test_gpu_abi.py -m gpu -q 2>&1 | tail -8
timeout 120 python profiles/scripts/genband_probe.py 8 4 65536 reference > gpurun_out/r4u/probe.jsonl; cat gpurun_out/r4u/probe.jsonl
