set -x
mkdir -p gpurun_out/r4p
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r4p/bench_n8.json 2> gpurun_out/r4p/bench_n8.err; tail -c 200 gpurun_out/r4p/bench_n8.json; tail -2 gpurun_out/r4p/bench_n8.err
