set -x
mkdir -p gpurun_out/r4z
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r4z/bench_n2.json 2> gpurun_out/r4z/bench_n2.err; tail -c 300 gpurun_out/r4z/bench_n2.json; tail -3 gpurun_out/r4z/bench_n2.err
