"""
GPU suite: the C ABI exercised exactly as INTEGRATION.md section 2 shows -- raw ctypes on the shared
library, no env class -- plus the boundary properties VERDICT r1 asked for: handles on two devices in one
process, the [n_sims][2] layout of the compact host actions, the population call and the diagnostics.
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import gw_oracle as O
from util import random_scenario, random_tapes

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gymwipe_b200", "lib", "libgymwipe_b200.so")


class gw_config(C.Structure):            # opaque, as in INTEGRATION.md: filled by gw_default_config
    _fields_ = [("raw", C.c_byte * 2048)]


def _vp(t):
    return C.c_void_p(t.data_ptr())


def test_raw_ctypes_create_step_destroy_matches_oracle():
    """gw_default_config -> gw_create(state = NULL) -> gw_reset -> gw_step x T -> gw_check -> gw_destroy
    through a bare ctypes.CDLL (no argtypes, no wrapper classes): results equal the oracle's."""
    from gymwipe_b200.scenario import default_scenario_dict
    import gymwipe_b200  # noqa: F401  (makes sure the in-tree library is built)
    gymwipe_b200.build()
    lib = C.CDLL(LIB)
    lib.gw_last_error.restype = C.c_char_p
    n, T = 64, 30
    cfg = gw_config()
    assert lib.gw_default_config(C.byref(cfg), C.c_int64(n)) == 0
    h = C.c_void_p()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.gw_create(C.byref(cfg), C.c_int(0), None, C.c_size_t(0), stream, C.byref(h))
    assert rc == 0, lib.gw_last_error()
    rs = np.random.RandomState(11)
    dev = rs.randint(0, 2, size=(T, n)).astype(np.int32)
    dur = rs.randint(0, 20, size=(T, n)).astype(np.int32)
    d_dev, d_dur = torch.as_tensor(dev).cuda(), torch.as_tensor(dur).cuda()
    obs = torch.empty(n, dtype=torch.int64, device="cuda")
    rew = torch.empty(n, dtype=torch.float64, device="cuda")
    done = torch.empty(n, dtype=torch.uint8, device="cuda")
    assert lib.gw_reset(h, None, C.c_int64(0), _vp(obs), stream) == 0
    assert (obs.cpu().numpy() == 65536).all()
    got_obs, got_rew = np.zeros((T, n), np.int64), np.zeros((T, n), np.float64)
    for t in range(T):
        rc = lib.gw_step(h, _vp(d_dev[t]), _vp(d_dur[t]), _vp(obs), _vp(rew), _vp(done), stream)
        assert rc == 0, lib.gw_last_error()
        got_obs[t], got_rew[t] = obs.cpu().numpy(), rew.cpu().numpy()
    assert lib.gw_check(h, stream) == 0
    now = torch.empty(n, dtype=torch.float64, device="cuda")
    assert lib.gw_read_state(h, C.c_int(0), _vp(now), stream) == 0
    ref = O.run_batch(default_scenario_dict(), dev, dur)
    assert (got_obs == ref["obs"][:, :, 0]).all() and (got_rew == ref["reward"][:, :, 0]).all()
    assert (now.cpu().numpy() == ref["now"][-1]).all()
    # an action outside the action space raises the device flag (the reference asserts, counter_traffic.py:147)
    bad = torch.full((n,), 25, dtype=torch.int32, device="cuda")
    assert lib.gw_step(h, _vp(d_dev[0]), _vp(bad), _vp(obs), _vp(rew), _vp(done), stream) == 0
    assert lib.gw_check(h, stream) == -4                      # GW_E_ACTION
    assert b"action" in lib.gw_last_error()
    lib.gw_destroy(h)


def test_compact_host_actions_are_sim_major_pairs():
    """include/gymwipe_b200.h: gw_step_host_compact reads actions as uint8 [n_sims][2] = {device, duration}
    per sim (r1 documented [2][n_sims]); pinned and pageable buffers agree with gw_step."""
    import gymwipe_b200
    n, T = 256, 12
    rs = np.random.RandomState(3)
    dev = rs.randint(0, 2, size=(T, n)).astype(np.int32)
    dur = rs.randint(0, 20, size=(T, n)).astype(np.int32)
    a = gymwipe_b200.make('CounterTraffic-v0', num_envs=n, strict=False)
    b = gymwipe_b200.make('CounterTraffic-v0', num_envs=n, strict=False)
    c = gymwipe_b200.make('CounterTraffic-v0', num_envs=n, strict=False)
    for e in (a, b, c):
        e.reset()
    res_pin = torch.empty(n, dtype=torch.int32).pin_memory()
    res_pag = np.empty(n, dtype=np.uint32)
    for t in range(T):
        o, r, d, _ = a.step({"device": torch.as_tensor(dev[t]).cuda(), "duration": torch.as_tensor(dur[t]).cuda()})
        pairs = np.stack([dev[t], dur[t]], axis=1).astype(np.uint8)          # [n, 2]: pairs[i] = (device_i, duration_i)
        b.step_host_compact(torch.as_tensor(pairs).pin_memory(), res_pin)
        c.step_host_compact(np.ascontiguousarray(pairs), res_pag)
        for res in (res_pin, torch.as_tensor(res_pag.view(np.int32))):
            oo, rr, dd = a.unpack_compact(res)
            assert torch.equal(oo, o.cpu()) and torch.equal(rr, r.cpu()) and torch.equal(dd, d.cpu())
    for e in (a, b, c):
        e.check()


def test_population_host_step_equals_single_batch_steps():
    """gw_step_host_compact_many: one call steps every batch of a population from its own pinned buffers."""
    import gymwipe_b200
    from gymwipe_b200.envs import EnvPopulation
    n, nb, T = 512, 5, 10
    rs = np.random.RandomState(8)
    pop = EnvPopulation([gymwipe_b200.make('CounterTraffic-v0', num_envs=n, env_id_offset=k * n, strict=False) for k in range(nb)])
    ref = [gymwipe_b200.make('CounterTraffic-v0', num_envs=n, env_id_offset=k * n, strict=False) for k in range(nb)]
    pop.reset()
    for e in ref:
        e.reset()
    res = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(nb)]
    res_ptrs = EnvPopulation.pointer_array(res)
    total_reward = 0.0
    for t in range(T):
        acts = [np.stack([rs.randint(0, 2, n), rs.randint(0, 20, n)], axis=1).astype(np.uint8) for _ in range(nb)]
        pinned = [torch.as_tensor(a).pin_memory() for a in acts]
        pop.step_host_compact(EnvPopulation.pointer_array(pinned), res_ptrs)
        for k in range(nb):
            o, r, d, _ = ref[k].step({"device": torch.as_tensor(acts[k][:, 0].astype(np.int32)).cuda(),
                                      "duration": torch.as_tensor(acts[k][:, 1].astype(np.int32)).cuda()})
            oo, rr, dd = ref[k].unpack_compact(res[k])
            assert torch.equal(oo, o.cpu()) and torch.equal(rr, r.cpu())
            total_reward += float(r.sum())
    pop.check()
    st = pop.stats().cpu().numpy()                      # one shared statistics vector for the population
    assert st[4] == n * nb * T and st[0] == total_reward
    pop.close()


def test_tiny_wire_format_equals_the_device_step():
    """gw_step_host_tiny / gw_step_host_tiny_many: 1 action byte in (device << 7 | duration), one 16-bit result word
    out -- the same observations, rewards and done flags as the device-resident step; pinned and pageable buffers."""
    import gymwipe_b200
    from gymwipe_b200.envs import EnvPopulation
    from gymwipe_b200.envs.counter_traffic import CounterTrafficEnv
    n, nb, T = 640, 3, 40
    rs = np.random.RandomState(31)
    pop = EnvPopulation([gymwipe_b200.make('CounterTraffic-v0', num_envs=n, env_id_offset=k * n, strict=False) for k in range(nb)])
    one = gymwipe_b200.make('CounterTraffic-v0', num_envs=n, strict=False)
    ref = [gymwipe_b200.make('CounterTraffic-v0', num_envs=n, env_id_offset=k * n, strict=False) for k in range(nb)]
    pop.reset(); one.reset()
    for e in ref:
        e.reset()
    acts = [torch.zeros(n, dtype=torch.uint8).pin_memory() for _ in range(nb)]
    res = [torch.empty(n, dtype=torch.int16).pin_memory() for _ in range(nb)]
    a_ptrs, r_ptrs = EnvPopulation.pointer_array(acts), EnvPopulation.pointer_array(res)
    pageable_res = np.empty(n, dtype=np.uint16)
    seen = set()
    for t in range(T):
        dev = [rs.randint(0, 2, n) for _ in range(nb)]
        dur = [rs.randint(0, 20, n) for _ in range(nb)]
        for k in range(nb):
            acts[k].copy_(CounterTrafficEnv.pack_tiny_actions(dev[k], dur[k]))
        pop.step_host_tiny(a_ptrs, r_ptrs)
        one.step_host_tiny(acts[0].numpy().copy(), pageable_res)           # pageable buffers: staged copies
        for k in range(nb):
            o, r, d, _ = ref[k].step({"device": torch.as_tensor(dev[k].astype(np.int32)).cuda(),
                                      "duration": torch.as_tensor(dur[k].astype(np.int32)).cuda()})
            oo, rr, dd = CounterTrafficEnv.unpack_tiny(res[k])
            assert torch.equal(oo, o.cpu()) and torch.equal(rr, r.cpu()) and torch.equal(dd, d.cpu()), (t, k)
            seen.update(int(v) for v in oo.unique())
            if k == 0:
                po, pr, pd = CounterTrafficEnv.unpack_tiny(pageable_res.view(np.int16))
                assert torch.equal(po, o.cpu()) and torch.equal(pr, r.cpu())
    assert seen == {65534, 65536, 65538}                                    # both signs of the difference occurred
    pop.check(); one.check()
    pop.close()


def test_population_host_step_replays_its_cached_graph():
    """The same call (same handles, same pinned buffers) is captured once and replayed: the replay reads the
    buffers' NEW contents; a change of the handles (gw_share_stats) or of the buffers takes a fresh capture; more
    distinct calls than the cache holds are evicted and re-captured."""
    import gymwipe_b200
    from gymwipe_b200.envs import EnvPopulation
    n, nb, T = 384, 3, 48
    rs = np.random.RandomState(18)
    pop = EnvPopulation([gymwipe_b200.make('CounterTraffic-v0', num_envs=n, env_id_offset=k * n, strict=False) for k in range(nb)])
    ref = [gymwipe_b200.make('CounterTraffic-v0', num_envs=n, env_id_offset=k * n, strict=False) for k in range(nb)]
    pop.reset()
    for e in ref:
        e.reset()
    res = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(nb)]
    res_ptrs = EnvPopulation.pointer_array(res)
    fixed = [torch.zeros((n, 2), dtype=torch.uint8).pin_memory() for _ in range(nb)]
    fixed_ptrs = EnvPopulation.pointer_array(fixed)
    for t in range(T):
        acts = [np.stack([rs.randint(0, 2, n), rs.randint(0, 20, n)], axis=1).astype(np.uint8) for _ in range(nb)]
        if t % 40 < 4:
            pinned = [torch.as_tensor(a).pin_memory() for a in acts]       # new buffers: a new capture (40 > cache size)
            pop.step_host_compact(EnvPopulation.pointer_array(pinned), res_ptrs)
        else:
            for k in range(nb):
                fixed[k].copy_(torch.as_tensor(acts[k]))                    # same buffers, new contents: a replay
            pop.step_host_compact(fixed_ptrs, res_ptrs)
        if t == 20:
            pop.stats()                                                     # (clears) ... and the handles change:
            pop.envs[1].share_stats(None)                                   # env 1 counts on its own again
        for k in range(nb):
            o, r, d, _ = ref[k].step({"device": torch.as_tensor(acts[k][:, 0].astype(np.int32)).cuda(),
                                      "duration": torch.as_tensor(acts[k][:, 1].astype(np.int32)).cuda()})
            oo, rr, dd = ref[k].unpack_compact(res[k])
            assert torch.equal(oo, o.cpu()) and torch.equal(rr, r.cpu()), t
    pop.check()
    own = pop.envs[1].stats().cpu().numpy()
    assert own[4] == n * (T - 21)                                            # steps 21 .. T-1 went to its own accumulators
    pop.close()


@pytest.mark.skipif(torch.cuda.is_available() and torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_handles_on_two_devices_in_one_process():
    """A jammer scenario needs > 48 KB of dynamic shared memory: the attribute is per device and must be set
    for every handle's device (r1: a process-wide flag set it once)."""
    import gymwipe_b200
    rs = np.random.RandomState(21)
    sc = random_scenario(rs, jammers=1, spread=2.5)
    dev, dur = random_tapes(rs, 20, 32, 1)
    want = O.run_batch(sc, dev, dur)
    for device in ("cuda:0", "cuda:1"):
        env = gymwipe_b200.make('CounterTraffic-v0', num_envs=32, scenario=sc, device=device, strict=False)
        env.reset()
        for t in range(20):
            o, r, d, _ = env.step({"device": torch.as_tensor(dev[t, :, 0]).to(device), "duration": torch.as_tensor(dur[t, :, 0]).to(device)})
            assert (o.cpu().numpy() == want["obs"][t, :, 0]).all()
        env.check()


def test_mask_bytes_statistic_counts_decided_sections():
    """gw_mask_bytes = 4 bytes per 32-bit word holding on-air bits of every decided section; with all-zero
    masks every section passes, so it can be predicted from the transmissions of the oracle's trace."""
    import gymwipe_b200
    rs = np.random.RandomState(5)
    sc = random_scenario(rs, jammers=0, spread=1.5, fixed_payload=300, factor=10000)
    n, T, slots, words = 8, 6, 2, 128
    masks = torch.zeros((n, 1, 4, slots, 4, words), dtype=torch.int32, device="cuda")
    env = gymwipe_b200.make('CounterTraffic-v0', num_envs=n, scenario=sc, mode="mask_fed", strict=False)
    env.set_masks(masks, slots)
    env.reset()
    dev, dur = random_tapes(rs, T, n, 1)
    for t in range(T):
        env.step({"device": torch.as_tensor(dev[t, :, 0]).cuda(), "duration": torch.as_tensor(dur[t, :, 0]).cuda()})
    env.check()
    got = env.mask_bytes()
    ntx = int(env.transmissions().sum())
    assert got > 0 and got % 4 == 0
    # every transmission is heard by the two other devices: a header section (139 on-air bits -> 5 words) and a
    # payload section each; lower / upper bounds from the announcement (1-6 byte payload) and data packet sizes
    assert 2 * ntx * 4 * 5 <= got <= 2 * ntx * 4 * (5 + 110)
    assert env.mask_bytes() == 0                        # cleared by the previous call


def test_debug_stamps_record_every_launch():
    import gymwipe_b200
    from gymwipe_b200 import _native as N
    env = gymwipe_b200.make('CounterTraffic-v0', num_envs=4096, strict=False)
    env.reset()
    K = 6
    st = torch.tensor([[-1, -1, 0, 0]] * K, dtype=torch.int64, device="cuda")
    N.check(N.lib().gw_debug_stamps(env._handle, st.data_ptr(), K))
    a = {"device": torch.zeros(4096, dtype=torch.int32, device="cuda"), "duration": torch.full((4096,), 3, dtype=torch.int32, device="cuda")}
    for _ in range(K + 2):                              # launches beyond the capacity are not stamped
        env.step(a)
    torch.cuda.synchronize()
    N.check(N.lib().gw_debug_stamps(env._handle, None, 0))
    s = st.cpu().numpy().astype(np.uint64)
    assert (s[:, 0] <= s[:, 1]).all() and (s[:, 1] < s[:, 2]).all()
    assert (s[1:, 1] >= s[:-1, 2]).all()                # a launch passes its grid dependency after the previous one ended


def test_raw_ctypes_general_band_engine_matches_oracle():
    """The general band engine through a bare ctypes.CDLL, as a C caller would drive it: a gw_genband_config struct
    declared field by field from include/gymwipe_b200.h, gw_genband_create -> _reset -> _step x T -> _read -> _check
    -> _destroy on a band of 3 senders + RRM + 1 PHY-only sender; results equal the oracle's."""
    import gymwipe_b200
    gymwipe_b200.build()
    lib = C.CDLL(LIB)
    lib.gw_last_error.restype = C.c_char_p

    class gw_genband_config(C.Structure):
        _fields_ = [("abi_version", C.c_int32), ("n_envs", C.c_int64), ("n_senders", C.c_int32), ("n_phy_senders", C.c_int32),
                    ("assignment_duration_factor", C.c_int32), ("max_assign_duration", C.c_int32), ("per_env_positions", C.c_int32),
                    ("mode", C.c_int32), ("seed", C.c_uint64), ("env_id_offset", C.c_int64),
                    ("frequency_hz", C.c_double), ("bandwidth_hz", C.c_double),
                    ("multiplicity", C.c_int32 * 8), ("payload_bytes", C.c_int32 * 8), ("destination", C.c_int32 * 8),
                    ("max_ticks", C.c_int32 * 8), ("receive", C.c_int32 * 8), ("interval", C.c_double * 8),
                    ("phy_interval", C.c_double * 16), ("phy_delay", C.c_double * 16), ("phy_power_dbm", C.c_double * 16),
                    ("phy_header_bytes", C.c_int32 * 16), ("phy_payload_bytes", C.c_int32 * 16)]

    sc = {"assignment_duration_factor": 1000, "bands": [{"frequency": 2.4e9, "bandwidth": 22e6, "devices": [
        {"role": "sender", "x": 2.0, "y": 0.0, "mult": 1, "payload": "counter", "interval": 0.001, "dest": 1},
        {"role": "sender", "x": -1.0, "y": 1.7, "mult": 3, "payload": "counter", "interval": 0.001, "dest": 2},
        {"role": "sender", "x": -1.0, "y": -1.7, "mult": 2, "payload": 30, "interval": 0.001, "dest": 0},
        {"role": "rrm", "x": 0.0, "y": 0.0},
        {"role": "jammer", "x": 4.0, "y": 4.0, "interval": 0.012, "delay": 0.002, "power": 10.0, "hdr": 13, "payload": 50}]}]}
    n, T = 48, 20
    cfg = gw_genband_config()
    cfg.abi_version = lib.gw_abi_version()
    cfg.n_envs, cfg.n_senders, cfg.n_phy_senders = n, 3, 1
    cfg.assignment_duration_factor, cfg.max_assign_duration, cfg.per_env_positions, cfg.mode = 1000, 20, 0, 0
    cfg.frequency_hz, cfg.bandwidth_hz = 2.4e9, 22e6
    for k, (m, p, d) in enumerate([(1, -1, 1), (3, -1, 2), (2, 30, 0)]):
        cfg.multiplicity[k], cfg.payload_bytes[k], cfg.destination[k], cfg.interval[k] = m, p, d, 0.001
    cfg.phy_interval[0], cfg.phy_delay[0], cfg.phy_power_dbm[0] = 0.012, 0.002, 10.0
    cfg.phy_header_bytes[0], cfg.phy_payload_bytes[0] = 13, 50
    pos = torch.tensor([[2.0, 0.0], [-1.0, 1.7], [-1.0, -1.7], [0.0, 0.0], [4.0, 4.0]], dtype=torch.float64, device="cuda")
    h = C.c_void_p()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.gw_genband_create(C.byref(cfg), C.c_int(0), _vp(pos), stream, C.byref(h))
    assert rc == 0, lib.gw_last_error()
    rs = np.random.RandomState(12)
    dev = rs.randint(0, 3, size=(T, n)).astype(np.int32)
    dur = rs.randint(0, 20, size=(T, n)).astype(np.int32)
    ref = O.run_batch(sc, dev, dur)
    d_dev, d_dur = torch.as_tensor(dev).cuda(), torch.as_tensor(dur).cuda()
    obs = torch.empty(n, dtype=torch.int64, device="cuda")
    rew = torch.empty(n, dtype=torch.float64, device="cuda")
    done = torch.empty(n, dtype=torch.uint8, device="cuda")
    assert lib.gw_genband_reset(h, _vp(obs), stream) == 0
    assert (obs.cpu().numpy() == 65536).all()
    now = torch.empty(n, dtype=torch.float64, device="cuda")
    for t in range(T):
        rc = lib.gw_genband_step(h, _vp(d_dev[t]), _vp(d_dur[t]), _vp(obs), _vp(rew), _vp(done), stream)
        assert rc == 0, lib.gw_last_error()
        assert lib.gw_genband_read(h, C.c_int(0), _vp(now), stream) == 0            # GW_GENBAND_FIELD_NOW
        assert (obs.cpu().numpy() == ref["obs"][t, :, 0]).all() and (rew.cpu().numpy() == ref["reward"][t, :, 0]).all(), t
        assert (now.cpu().numpy() == ref["now"][t]).all(), t
    assert lib.gw_genband_check(h, stream) == 0
    deliv = torch.empty((3, n), dtype=torch.float64, device="cuda")
    assert lib.gw_genband_read(h, C.c_int(1), _vp(deliv), stream) == 0              # GW_GENBAND_FIELD_DELIVERED
    assert (deliv.t().cpu().numpy() == ref["counts"][:, 0, 1:4]).all() and ref["counts"][:, 0, 1:4].sum() > 0
    # an action outside the action space is reported by gw_genband_check and leaves that env untouched
    bad = d_dev[0].clone(); bad[5] = 3
    assert lib.gw_genband_step(h, _vp(bad), _vp(d_dur[0]), _vp(obs), _vp(rew), _vp(done), stream) == 0
    assert lib.gw_genband_check(h, stream) == -4                                    # GW_E_ACTION
    lib.gw_genband_destroy(h)
