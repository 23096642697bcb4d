set -x
mkdir -p gpurun_out/r3b
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r3b/bench_n8.json 2> gpurun_out/r3b/bench_n8.err; tail -5 gpurun_out/r3b/bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > gpurun_out/r3b/bench_ref_n8.json 2> gpurun_out/r3b/bench_ref_n8.err
