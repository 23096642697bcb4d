set -x
mkdir -p gpurun_out/r2g
python -m pytest tests -m gpu -x -q > gpurun_out/r2g/pytest_gpu.log 2>&1; tail -5 gpurun_out/r2g/pytest_gpu.log
python profiles/scripts/cfg3_probe.py 24 > gpurun_out/r2g/probe_product.json 2> gpurun_out/r2g/probe_product.err
for v in noscan u2; do
GYMWIPE_B200_LIB=$PWD/gymwipe_b200/lib/variants/lib_$v.so python profiles/scripts/cfg3_probe.py 24 > gpurun_out/r2g/probe_$v.json 2> gpurun_out/r2g/probe_$v.err
done
cat gpurun_out/r2g/probe_*.json
python profiles/scripts/kernel_stamps.py steady 16 8 > gpurun_out/r2g/stamps_steady.json 2> gpurun_out/r2g/stamps_steady.err
python profiles/scripts/kernel_stamps.py productive 16 8 > gpurun_out/r2g/stamps_productive.json 2> gpurun_out/r2g/stamps_productive.err
cat gpurun_out/r2g/stamps_*.json; tail -3 gpurun_out/r2g/stamps_steady.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2g/bench_k20.json 2> gpurun_out/r2g/bench_k20.err; tail -3 gpurun_out/r2g/bench_k20.err
python bench.py --steps 200 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2g/bench_k200.json 2> gpurun_out/r2g/bench_k200.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2g/bench_ref.json 2> gpurun_out/r2g/bench_ref.err
