set -x
mkdir -p gpurun_out/r2y
timeout 600 python -m pytest tests/test_gpu_pendulum.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
timeout 300 python profiles/scripts/cfg5_timing.py 2>&1 | tail -1
