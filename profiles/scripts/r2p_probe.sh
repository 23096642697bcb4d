set -x
mkdir -p gpurun_out/r2p
timeout 300 python -m pytest tests/test_gpu_mode_m.py tests/test_gpu_fullsize.py -x -q > gpurun_out/r2p/pytest_mode_m.log 2>&1; tail -15 gpurun_out/r2p/pytest_mode_m.log
timeout 200 python profiles/scripts/cfg3_probe.py 24 > gpurun_out/r2p/cfg3_ws.json 2> gpurun_out/r2p/cfg3_ws.err; cat gpurun_out/r2p/cfg3_ws.json; tail -3 gpurun_out/r2p/cfg3_ws.err
GW_FED_WS=0 timeout 200 python profiles/scripts/cfg3_probe.py 24 > gpurun_out/r2p/cfg3_rounds.json 2> gpurun_out/r2p/cfg3_rounds.err; cat gpurun_out/r2p/cfg3_rounds.json
