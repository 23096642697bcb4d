set -x
mkdir -p gpurun_out/r4s
V=$PWD/gymwipe_b200/lib/variants/lib_order.so
timeout 200 python profiles/scripts/order_probe.py 0 48 32 0 > gpurun_out/r4s/probe.jsonl
GYMWIPE_B200_LIB=$V timeout 200 python profiles/scripts/order_probe.py 0 48 32 0 >> gpurun_out/r4s/probe.jsonl
GYMWIPE_B200_LIB=$V timeout 200 python profiles/scripts/order_probe.py 1 48 32 0 >> gpurun_out/r4s/probe.jsonl
GYMWIPE_B200_LIB=$V timeout 200 python profiles/scripts/order_probe.py 1 48 16 1 >> gpurun_out/r4s/probe.jsonl
GYMWIPE_B200_LIB=$V timeout 200 python profiles/scripts/order_probe.py 0 48 16 1 >> gpurun_out/r4s/probe.jsonl
cat gpurun_out/r4s/probe.jsonl
