set -x
mkdir -p gpurun_out/r2j
python -m pytest tests/test_gpu_dqn.py tests/test_construction_cpu.py -x -q > gpurun_out/r2j/pytest_new.log 2>&1; tail -5 gpurun_out/r2j/pytest_new.log
GYMWIPE_B200_LIB=$PWD/gymwipe_b200/lib/variants/lib_mb5.so python profiles/scripts/multistream_probe.py 64 > gpurun_out/r2j/multistream_mb5.json 2> gpurun_out/r2j/multistream_mb5.err; cat gpurun_out/r2j/multistream_mb5.json; tail -3 gpurun_out/r2j/multistream_mb5.err
