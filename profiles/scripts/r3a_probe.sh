set -x
mkdir -p gpurun_out/r3a
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r3a/bench_q.json 2> gpurun_out/r3a/bench_q.err; tail -3 gpurun_out/r3a/bench_q.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3a/bench_q.json').read().strip().splitlines()[-1])
print('RESULT',d['value'],d['productive']['value'],d['e2e']['value'])
PY
