set -x
mkdir -p gpurun_out/r2r
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2r/pytest_gpu.log 2>&1; tail -8 gpurun_out/r2r/pytest_gpu.log
timeout 200 python profiles/scripts/cfg3_probe.py 24 > gpurun_out/r2r/cfg3_index.json 2> gpurun_out/r2r/cfg3_index.err; python -c "
import json;d=json.load(open('gpurun_out/r2r/cfg3_index.json'))['mode_m_fed'];print('RESULT',d['ms_per_step'],d['roofline']['frac'],d['roofline']['avg_launch_ms'])"; tail -3 gpurun_out/r2r/cfg3_index.err
