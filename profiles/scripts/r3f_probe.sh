set -x
mkdir -p gpurun_out/r3f
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3f/bench.json 2> gpurun_out/r3f/bench.err; tail -3 gpurun_out/r3f/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3f/bench.json').read().strip().splitlines()[-1])
print('RESULT',d['value'],d['productive']['value'],'e2e',d['e2e']['value'],'cfg3',d['cfg3_long_packet_mode_m']['ms_per_step'],'cfg4',d['cfg4_multiband']['ms_per_step'],'cfg5',d['cfg5_pendulum']['ms_per_step'])
PY
