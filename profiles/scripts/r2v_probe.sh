set -x
mkdir -p gpurun_out/r2v
timeout 600 python -m pytest tests/test_gpu_abi.py -x -q > gpurun_out/r2v/pytest_abi.log 2>&1; tail -12 gpurun_out/r2v/pytest_abi.log
