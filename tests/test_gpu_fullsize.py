"""
GPU suite at BASELINE.json's full per-GPU sizes for the non-headline configs: size-independent
properties plus a strided sample of envs checked bit for bit against the oracle.
"""
import numpy as np
import pytest
import torch

import gw_oracle as O

pytestmark = pytest.mark.gpu


def _bands(nb, jam_power=10.0):
    bands = []
    for b in range(nb):
        bands.append({"frequency": 2.4e9 + b * 25e6, "bandwidth": 22e6, "devices": [
            {"role": "sender", "x": 0.0, "y": 2.0, "mult": 1, "payload": "counter", "interval": 0.001, "dest": 1},
            {"role": "sender", "x": 0.0, "y": -2.0, "mult": 3, "payload": "counter", "interval": 0.001, "dest": 0},
            {"role": "rrm", "x": 0.0, "y": 0.0},
            {"role": "jammer", "x": 5.0, "y": 5.0, "interval": 0.013 + 0.002 * b, "delay": 0.001 * b, "power": jam_power,
             "hdr": 13, "payload": 60}]})
    return {"assignment_duration_factor": 1000, "bands": bands}


def test_config4_multiband_131072_envs_per_gpu():
    """configs[3]: 16 devices over 4 bands, positions per env; 1 M envs / 8 GPUs = 131 072 per GPU."""
    import gymwipe_b200
    n, T, nb = 131072, 12, 4
    sc = _bands(nb)
    g = torch.Generator(device="cuda").manual_seed(5)
    pos = torch.rand((n, nb, 4, 2), generator=g, device="cuda", dtype=torch.float64) * 8.0 - 4.0
    dev = torch.randint(0, 2, (T, n, nb), generator=g, device="cuda", dtype=torch.int32)
    dur = torch.randint(0, 20, (T, n, nb), generator=g, device="cuda", dtype=torch.int32)
    env = gymwipe_b200.make('CounterTraffic-v0', num_envs=n, scenario=sc, positions=pos, strict=False)
    env.reset()
    env.stats()
    obs_all = torch.empty((T, n, nb), dtype=torch.int64, device="cuda")
    prev_now = torch.zeros(n, dtype=torch.float64, device="cuda")
    for t in range(T):
        o, r, d, _ = env.step({"device": dev[t], "duration": dur[t]})
        obs_all[t] = o
        diff = o - 65536
        assert bool(((diff == -2) | (diff == 0) | (diff == 2)).all())
        now = env.read_state(0)
        # the env clock ends at the LATEST band: at least the longest assignment of the step
        longest = dur[t].max(dim=1).values.double() * 1000e-6
        assert bool((now - prev_now > longest).all())
        prev_now = now
    env.check()
    assert float(env.stats().cpu()[4]) == n * nb * T
    idx = np.arange(0, n, n // 128)
    o = O.run_batch(sc, dev[:, idx].cpu().numpy(), dur[:, idx].cpu().numpy(),
                    pos=np.concatenate([pos[idx].cpu().numpy(), np.zeros((len(idx), nb, 4, 2))], axis=2))
    assert (obs_all[:, idx].cpu().numpy() == o["obs"]).all()
    assert (env.read_state(0)[idx].cpu().numpy() == o["now"][-1]).all()
    assert (env.delivered().reshape(n, nb, 2)[idx].cpu().numpy() == o["counts"][:, :, 1:3]).all()


def test_config3_long_packets_philox_65536_envs():
    """configs[2]: 1500-byte payloads, per-bit Philox masks, a PHY-only interferer (mode M)."""
    import gymwipe_b200
    n, T = 65536, 6
    sc = {"assignment_duration_factor": 10000, "bands": [{"frequency": 2.4e9, "bandwidth": 22e6, "devices": [
        {"role": "sender", "x": 0.0, "y": 2.0, "mult": 1, "payload": 1500, "interval": 0.001, "dest": 1},
        {"role": "sender", "x": 0.0, "y": -2.0, "mult": 3, "payload": 1500, "interval": 0.001, "dest": 0},
        {"role": "rrm", "x": 0.0, "y": 0.0},
        {"role": "jammer", "x": 3.0, "y": 0.0, "interval": 0.05, "delay": 0.003, "power": 0.0, "hdr": 13, "payload": 200}]}]}
    g = torch.Generator(device="cuda").manual_seed(6)
    dev = torch.randint(0, 2, (T, n), generator=g, device="cuda", dtype=torch.int32)
    dur = torch.randint(12, 20, (T, n), generator=g, device="cuda", dtype=torch.int32)
    env = gymwipe_b200.make('CounterTraffic-v0', num_envs=n, scenario=sc, mode="mask_philox", seed=2026,
                            env_id_offset=1 << 33, strict=False)       # global ids beyond 32 bits
    env.reset()
    obs_all = torch.empty((T, n), dtype=torch.int64, device="cuda")
    for t in range(T):
        o, r, d, _ = env.step({"device": dev[t], "duration": dur[t]})
        obs_all[t] = o
    env.check()
    deliv = env.delivered()
    assert int(deliv.sum()) > n            # long packets do get through
    idx = np.arange(0, n, n // 64)
    # the oracle of env k must be keyed with ITS global id: run the sample one env at a time
    for k in idx[:24]:
        ora = O.run_batch(sc, dev[:, k:k + 1].cpu().numpy(), dur[:, k:k + 1].cpu().numpy(), mode=O.MODE_M, seed=2026,
                          env_id_offset=(1 << 33) + int(k), threads=1)
        assert (obs_all[:, k].cpu().numpy() == ora["obs"][:, 0, 0]).all()
        assert (deliv[k].cpu().numpy() == ora["counts"][0, 0, 1:3]).all()
