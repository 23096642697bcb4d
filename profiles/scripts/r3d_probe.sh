set -x
mkdir -p gpurun_out/r3d
timeout 600 python -m pytest tests/test_gpu_abi.py tests/test_gpu_parity.py -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r3d/bench_q.json 2> gpurun_out/r3d/bench_q.err; tail -3 gpurun_out/r3d/bench_q.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3d/bench_q.json').read().strip().splitlines()[-1])
print('RESULT',d['value'],d['productive']['value'],'e2e tiny',d['e2e']['value'],'compact',d['e2e']['compact_api']['value'],'single',d['e2e']['single_batch_sync']['value'])
PY
