set -x
mkdir -p gpurun_out/r2u
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2u/pytest_gpu.log 2>&1; tail -8 gpurun_out/r2u/pytest_gpu.log
timeout 300 python profiles/scripts/cfg4_profile.py 64 > gpurun_out/r2u/cfg4.txt 2>&1; tail -2 gpurun_out/r2u/cfg4.txt
timeout 300 python profiles/scripts/cfg5_timing.py > gpurun_out/r2u/cfg5.txt 2>&1; tail -3 gpurun_out/r2u/cfg5.txt
