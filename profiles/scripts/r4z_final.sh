set -x
mkdir -p gpurun_out/r4z
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r4z/pytest_gpu.log 2>&1; tail -4 gpurun_out/r4z/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
