"""The e2e leg alone (EnvPopulation.step_host_compact, PE batches per call) under torchrun: per-rank and aggregate
env-steps/s, with / without binding the rank to its GPU's NUMA node (GYMWIPE_B200_NO_BIND=1)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch.distributed as dist
import gymwipe_b200
from gymwipe_b200.distributed import init_from_env, bind_to_gpu_numa
from gymwipe_b200.envs import EnvPopulation

PE = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rank, world, local = init_from_env("nccl")
torch.cuda.set_device(local)
cpus = bind_to_gpu_numa(local)
dev = torch.device("cuda", local)
n = 65536
envs = [gymwipe_b200.make('CounterTraffic-v0', num_envs=n, device=dev, env_id_offset=(rank * PE + k) * n, strict=False) for k in range(PE)]
pop = EnvPopulation(envs)
pop.reset()
rs = np.random.RandomState(rank)
EROWS = 4
h_act = [[torch.as_tensor(np.stack([rs.randint(0, 2, n), rs.randint(0, 20, n)], axis=1).astype(np.uint8)).pin_memory() for _ in range(PE)] for _ in range(EROWS)]
h_res = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(PE)]
act_ptrs = [EnvPopulation.pointer_array(h_act[r]) for r in range(EROWS)]
res_ptrs = EnvPopulation.pointer_array(h_res)
for r in range(140):
    pop.step_host_compact(act_ptrs[r % EROWS], res_ptrs)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
K = 100
t0 = time.perf_counter()
for r in range(K):
    pop.step_host_compact(act_ptrs[r % EROWS], res_ptrs)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
v = torch.tensor([dt], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"world": world, "batches_per_call": PE, "bound_cpus": len(cpus) if cpus else None, "allowed_cpus": len(os.sched_getaffinity(0)),
                      "rank0_env_steps_per_s": n * PE * K / dt, "aggregate_env_steps_per_s": n * PE * K * world / float(v[0])}))
pop.check()
if world > 1:
    dist.destroy_process_group()
