set -x
mkdir -p gpurun_out/r4y
for v in b64 b96 b256; do
  GYMWIPE_B200_LIB=$PWD/gymwipe_b200/lib/variants/lib_$v.so timeout 200 python bench.py --steps 10 --warmup 3 --no-extras --batches 96 2> gpurun_out/r4y/$v.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', d['value'], d['us_per_launch'], d['productive']['value'])" >> gpurun_out/r4y/probe.txt
done
timeout 200 python bench.py --steps 10 --warmup 3 --no-extras --batches 96 2> gpurun_out/r4y/base.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('base128', d['value'], d['us_per_launch'], d['productive']['value'])" >> gpurun_out/r4y/probe.txt
cat gpurun_out/r4y/probe.txt
