set -x
mkdir -p gpurun_out/r3h
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
(time python bench.py) > gpurun_out/r3h/bench_default.json 2> gpurun_out/r3h/bench_default.err; tail -4 gpurun_out/r3h/bench_default.err
