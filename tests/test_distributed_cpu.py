"""
CPU suite, part 4: the multi-GPU host logic on the gloo backend, world_size 2 (no GPU needed):
shard ranges, the asynchronous statistics reducer, and sharding invariance of the per-env
results (checked with the oracle standing in for the device step, keyed by GLOBAL env ids).
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    from gymwipe_b200.distributed import shard_range
    for total in (1, 7, 8, 65536, 1000003):
        for world in (1, 2, 3, 8):
            ranges = [shard_range(total, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RANK"], os.environ["WORLD_SIZE"], os.environ["LOCAL_RANK"] = str(rank), str(world), str(rank)
    from gymwipe_b200.distributed import StatsReducer, init_from_env, shard_range
    import gw_oracle as O
    from gymwipe_b200.scenario import default_scenario_dict
    r, w, _ = init_from_env("gloo")
    assert (r, w) == (rank, world)
    # every rank steps its shard of a 96-env job (mode M: keys depend on the global env id)
    total, T = 96, 40
    rs = np.random.RandomState(0)
    dev = rs.randint(0, 2, size=(T, total)).astype(np.int32)
    dur = rs.randint(0, 20, size=(T, total)).astype(np.int32)
    a, b = shard_range(total, rank, world)
    res = O.run_batch(default_scenario_dict(), dev[:, a:b], dur[:, a:b], mode=O.MODE_M, seed=3,
                      env_id_offset=a, threads=1)
    reducer = StatsReducer("cpu", width=8, depth=3)
    sums = []
    for t in range(T):
        local = torch.zeros(8, dtype=torch.float64)
        local[0] = float(res["reward"][t].sum())
        local[4] = b - a
        out = reducer.submit(local)
        if out is not None:
            sums.append(out)
    sums += reducer.drain()
    assert len(sums) == T
    np.save(os.path.join(tmpdir, "obs_%d.npy" % rank), res["obs"])
    np.save(os.path.join(tmpdir, "sums_%d.npy" % rank), torch.stack(sums).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_stats(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gw_oracle as O
    from gymwipe_b200.scenario import default_scenario_dict
    port = 29000 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    total, T = 96, 40
    rs = np.random.RandomState(0)
    dev = rs.randint(0, 2, size=(T, total)).astype(np.int32)
    dur = rs.randint(0, 20, size=(T, total)).astype(np.int32)
    whole = O.run_batch(default_scenario_dict(), dev, dur, mode=O.MODE_M, seed=3, env_id_offset=0, threads=2)
    obs = np.concatenate([np.load(tmp_path / ("obs_%d.npy" % r)) for r in range(2)], axis=1)
    assert (obs == whole["obs"]).all()                      # results do not depend on the sharding
    s0, s1 = np.load(tmp_path / "sums_0.npy"), np.load(tmp_path / "sums_1.npy")
    assert (s0 == s1).all()                                  # every rank holds the global sums
    assert (s0[:, 4] == total).all()
    assert (s0[:, 0] == whole["reward"][:, :, 0].sum(axis=1)).all()
