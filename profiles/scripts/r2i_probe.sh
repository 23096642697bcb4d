set -x
mkdir -p gpurun_out/r2i
python profiles/scripts/multistream_probe.py 64 > gpurun_out/r2i/multistream.json 2> gpurun_out/r2i/multistream.err; cat gpurun_out/r2i/multistream.json; tail -3 gpurun_out/r2i/multistream.err
