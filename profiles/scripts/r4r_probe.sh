set -x
timeout 900 python -m pytest tests/test_gpu_dqn.py tests/test_general_band.py tests/test_grid.py -m gpu -q 2>&1 | tail -6
