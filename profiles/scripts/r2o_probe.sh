set -x
mkdir -p gpurun_out/r2o
python -m pytest tests -m gpu -x -q > gpurun_out/r2o/pytest_gpu.log 2>&1; tail -3 gpurun_out/r2o/pytest_gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2o/bench_n2.json 2> gpurun_out/r2o/bench_n2.err; tail -5 gpurun_out/r2o/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2o/bench_ref_n2.json 2> gpurun_out/r2o/bench_ref_n2.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2o/bench_n1.json 2> gpurun_out/r2o/bench_n1.err
