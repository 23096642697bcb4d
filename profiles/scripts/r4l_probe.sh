set -x
mkdir -p gpurun_out/r4l
timeout 900 python -m pytest tests/test_general_band.py tests/test_grid.py tests/test_gpu_abi.py tests/test_construction_cpu.py -m gpu -q 2>&1 | tail -15
