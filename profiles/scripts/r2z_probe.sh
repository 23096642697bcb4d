set -x
mkdir -p gpurun_out/r2z
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2z/pytest_gpu.log 2>&1; tail -4 gpurun_out/r2z/pytest_gpu.log
echo SIMMEMO; timeout 300 python profiles/scripts/cfg4_profile.py 64 2>&1 | tail -1
echo NOMEMO; GYMWIPE_B200_NO_MEMO=1 timeout 300 python profiles/scripts/cfg4_profile.py 64 2>&1 | tail -1
