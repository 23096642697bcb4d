set -x
mkdir -p gpurun_out/r3c
nvidia-smi topo -m > gpurun_out/r3c/topo.txt 2>&1; lscpu | grep -i "numa\|socket\|^CPU(s)" > gpurun_out/r3c/lscpu.txt; python -c "import os; print(len(os.sched_getaffinity(0)))" >> gpurun_out/r3c/lscpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 profiles/scripts/e2e_many.py 64 > gpurun_out/r3c/e2e_bind.json 2> gpurun_out/r3c/e2e_bind.err; cat gpurun_out/r3c/e2e_bind.json; tail -3 gpurun_out/r3c/e2e_bind.err
GYMWIPE_B200_NO_BIND=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 profiles/scripts/e2e_many.py 64 > gpurun_out/r3c/e2e_nobind.json 2> gpurun_out/r3c/e2e_nobind.err; cat gpurun_out/r3c/e2e_nobind.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 profiles/scripts/e2e_many.py 16 > gpurun_out/r3c/e2e_bind16.json 2> gpurun_out/r3c/e2e_bind16.err; cat gpurun_out/r3c/e2e_bind16.json
