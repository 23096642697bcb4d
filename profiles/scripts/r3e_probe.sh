set -x
mkdir -p gpurun_out/r3e
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r3e/bench_n8.json 2> gpurun_out/r3e/bench_n8.err; tail -5 gpurun_out/r3e/bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 20 --warmup 5 --no-extras > gpurun_out/r3e/bench_n4.json 2> gpurun_out/r3e/bench_n4.err; tail -5 gpurun_out/r3e/bench_n4.err
