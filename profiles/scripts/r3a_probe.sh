set -x
mkdir -p gpurun_out/r3a
for pe in 8 32 64; do
python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --e2e-batches $pe > gpurun_out/r3a/bench_pe$pe.json 2> gpurun_out/r3a/bench_q.err; tail -3 gpurun_out/r3a/bench_q.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r3a/bench_pe$pe.json').read().strip().splitlines()[-1])
print('RESULT pe=$pe',d['value'],d['productive']['value'],d['e2e']['value'])
PY
done
