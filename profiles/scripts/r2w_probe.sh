set -x
mkdir -p gpurun_out/r2w
echo MEMO; timeout 300 python profiles/scripts/cfg4_profile.py 64 2>&1 | tail -1
echo NOMEMO; GYMWIPE_B200_NO_MEMO=1 timeout 300 python profiles/scripts/cfg4_profile.py 64 2>&1 | tail -1
for lib in libgymwipe_b200.so variants/lib_rollall.so; do
GYMWIPE_B200_LIB=gymwipe_b200/lib/$lib timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2w/bench_$(basename $lib).json 2> gpurun_out/r2w/bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2w/bench_$(basename $lib).json').read().strip().splitlines()[-1])
print('RESULT $lib',d['value'],d['productive']['value'],d['e2e']['value'])
PY
done
