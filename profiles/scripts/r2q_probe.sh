set -x
mkdir -p gpurun_out/r2q
for v in wsdbg; do
GYMWIPE_B200_LIB=gymwipe_b200/lib/variants/lib_$v.so timeout 200 python profiles/scripts/ws_debug.py 8 > gpurun_out/r2q/$v.json 2> gpurun_out/r2q/$v.err; grep -A20 per_scanner gpurun_out/r2q/$v.json; grep ms_per gpurun_out/r2q/$v.json; tail -3 gpurun_out/r2q/$v.err
done
