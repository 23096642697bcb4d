"""
GPU suite (`-m gpu`): the CUDA path, called through the C ABI / the public env API, against
 (a) the golden vectors produced by the unmodified reference,
 (b) the oracle (plain-C restatement) on the same seeded inputs,
 (c) size-independent properties at BASELINE.json's full batch size.
Bit-exact for obs / reward / done / delivery and transmission counts AND the fp64 step end
times (event order); <= 1e-9 relative for BER / attenuation (stated tolerance: 1e-6).
"""
import ctypes as C
import math

import numpy as np
import pytest
import torch

import gw_oracle as O
from util import (GOLDEN_CASES, golden_results, load_golden, random_scenario, random_tapes,
                  tapes_from_golden)

pytestmark = pytest.mark.gpu

BER_RTOL = 1e-9          # north star: 1e-6 relative in fp64; observed ~1e-15


def make_env(scenario=None, n=1, **kw):
    import gymwipe_b200
    return gymwipe_b200.make('CounterTraffic-v0', num_envs=n, scenario=scenario, **kw)


def run_gpu(sc, dev, dur, do_reset=True, pos=None, **kw):
    """dev, dur: [steps, nenv, nb] -> dict like oracle.run_batch"""
    nsteps, nenv, nb = dev.shape
    env = make_env(sc, nenv, strict=False, positions=None if pos is None else torch.as_tensor(pos[:, :, :4, :]).cuda(), **kw)
    if do_reset:
        env.reset()
    d_dev = torch.as_tensor(dev).cuda()
    d_dur = torch.as_tensor(dur).cuda()
    obs = np.zeros((nsteps, nenv, nb), np.int64)
    rew = np.zeros((nsteps, nenv, nb), np.float64)
    done = np.zeros((nsteps, nenv, nb), np.uint8)
    now = np.zeros((nsteps, nenv), np.float64)
    for t in range(nsteps):
        a = {"device": d_dev[t].reshape(env._shape), "duration": d_dur[t].reshape(env._shape)}
        o, r, d, _ = env.step(a)
        obs[t] = o.reshape(nenv, nb).cpu().numpy()
        rew[t] = r.reshape(nenv, nb).cpu().numpy()
        done[t] = d.reshape(nenv, nb).cpu().numpy()
        now[t] = env.read_state(0).cpu().numpy()
    env.check()
    counts = np.zeros((nenv, nb, 3), np.int64)
    counts[:, :, 0] = env.transmissions().cpu().numpy().reshape(nenv, nb)
    counts[:, :, 1:3] = env.delivered().cpu().numpy().reshape(nenv, nb, 2)
    ties = env.read_state(11).cpu().numpy().sum()
    return {"obs": obs, "reward": rew, "done": done, "now": now, "counts": counts, "ties": ties, "env": env}


def assert_same(o, g):
    assert (o["obs"] == g["obs"]).all()
    assert (o["reward"] == g["reward"]).all()
    assert (o["done"] == g["done"]).all()
    assert (o["now"] == g["now"]).all()
    assert (o["counts"][:, :, :3] == g["counts"]).all()


def test_reference_own_test_verbatim():
    """tests/envs/test_counter_traffic.py of the reference, with gymwipe_b200.make for gym.make."""
    import gymwipe_b200
    env = gymwipe_b200.make('CounterTraffic-v0')
    np.random.seed(123)
    env.seed(123)
    observation_center = env.COUNTER_BOUND
    observation, reward, _, _ = env.step({"device": 0, "duration": 3})
    assert observation - observation_center == 2
    assert reward == -2
    observation, reward, _, _ = env.step({"device": 1, "duration": 12})
    assert observation - observation_center == 0
    assert reward == 2
    assert isinstance(observation, int) and isinstance(reward, float)
    assert env.now == 0.017804000036000002


def test_gym_surface():
    env = make_env()
    assert env.action_space.contains({"device": 1, "duration": 19})
    assert not env.action_space.contains({"device": 2, "duration": 0})
    assert env.observation_space.n == 131072
    assert env.seed(5) == [5]
    assert env.reset() == 65536
    assert env.COUNTER_BOUND == 65536 and env.MAX_ASSIGN_DURATION == 20 and env.ASSIGNMENT_DURATION_FACTOR == 1000
    assert len(env.senders) == 2 and env.senders[1].packetMultiplicity == 3
    assert env.deviceIndexToMacDict[0] == bytes([0, 0, 0, 0, 0, 1])
    with pytest.raises(AssertionError):
        env.step({"device": 2, "duration": 3})           # the reference asserts too
    o, r, d, info = env.step({"device": 0, "duration": 15})
    assert info["Latest received values"] == "[2, 0]" and d is False
    assert env.senders[0].counter >= 16


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_cuda_matches_reference_golden(name):
    doc = load_golden(name)
    dev, dur = tapes_from_golden(doc)
    g = run_gpu(doc["scenario"], dev, dur, do_reset=doc["do_reset"])
    obs, rew, done, now = golden_results(doc)
    assert (g["obs"][:, 0, :] == obs).all()
    assert (g["reward"][:, 0, :] == rew).all()
    assert (g["done"][:, 0, :] == done).all()
    assert (g["now"][:, 0] == now).all()
    n_tx = sum(1 for s in doc["steps"] for r in s["records"] if r[0] == "tx")
    n_rx = sum(1 for s in doc["steps"] for r in s["records"] if r[0] == "rx" and r[3] < 2)
    assert g["counts"][0, :, 0].sum() == n_tx
    assert g["counts"][0, :, 1:3].sum() == n_rx


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_cuda_event_trace_matches_reference_golden(name):
    """
    Event-level parity measured ON THE GPU against the reference's own trace: every transmission
    (sender, start, stop, coded bit counts), every RRM delivery and every decider verdict bit-exact
    and in the same order; BER values and expected error sums within 1e-9 relative (stated
    tolerance 1e-6: CUDA libdevice log10/pow differ from glibc by <= 2 ulp).
    """
    from util import canonical_per_band as canonical
    doc = load_golden(name)
    dev, dur = tapes_from_golden(doc)
    nb = dev.shape[2]
    env = make_env(doc["scenario"], 1, strict=False)
    if doc["do_reset"]:
        env.reset()
    worst = 0.0
    for t, s in enumerate(doc["steps"]):
        a = {"device": torch.as_tensor(dev[t]).cuda().reshape(env._shape), "duration": torch.as_tensor(dur[t]).cuda().reshape(env._shape)}
        obs, rew, done, recs = env.step_traced(a)
        got = canonical([r for b in range(nb) for r in recs[b]])
        want = canonical(s["records"])
        assert len(got) == len(want), (name, t)
        for g, w in zip(got, want):
            assert g[0] == w[0] and g[1] == w[1] and g[2] == w[2] and g[3] == w[3], (name, t, g, w)
            if g[0] == "tx":
                assert tuple(g[4:]) == tuple(w[4:]), (name, t, g, w)
            elif g[0] == "ber":
                rel = abs(g[4] - w[4]) / max(abs(w[4]), 1e-300)      # the BER underflows to exactly 0 on very strong links
                worst = max(worst, rel)
                assert rel <= BER_RTOL, (name, t, g, w)
            elif g[0] == "dec":
                assert g[4] == w[4] and g[6] == w[6] and g[7] == w[7], (name, t, g, w)
                rel = abs(g[5] - w[5]) / max(abs(w[5]), 1e-300) if w[5] != 0 else abs(g[5])
                worst = max(worst, rel)
                assert rel <= BER_RTOL, (name, t, g, w)
    print("%s: max relative BER / error-sum deviation vs reference %.3e" % (name, worst))


def test_cuda_matches_oracle_default_4096x128():
    from gymwipe_b200.scenario import default_scenario_dict
    sc = default_scenario_dict()
    rs = np.random.RandomState(21)
    dev, dur = random_tapes(rs, 128, 4096, 1)
    o = O.run_batch(sc, dev, dur)
    g = run_gpu(sc, dev, dur)
    assert_same(o, g)
    assert g["ties"] == 0
    assert o["counts"][:, :, 1:3].sum() > 100000


def test_cuda_matches_oracle_no_reset():
    from gymwipe_b200.scenario import default_scenario_dict
    sc = default_scenario_dict()
    rs = np.random.RandomState(22)
    dev, dur = random_tapes(rs, 64, 1024, 1)
    assert_same(O.run_batch(sc, dev, dur, do_reset=False), run_gpu(sc, dev, dur, do_reset=False))


@pytest.mark.parametrize("seed", range(4))
def test_cuda_matches_oracle_random_scenarios(seed):
    rs = np.random.RandomState(700 + seed)
    spread = [1.5, 2.5, 4.0][seed % 3]
    for sc, nenv, nsteps in [(random_scenario(rs, jammers=0, spread=spread), 256, 100),
                             (random_scenario(rs, jammers=1, spread=spread), 256, 100),
                             (random_scenario(rs, jammers=1, spread=spread, fixed_payload=1500, factor=10000), 64, 40),
                             (random_scenario(rs, nbands=4, jammers=1, spread=spread), 64, 60)]:
        dev, dur = random_tapes(rs, nsteps, nenv, len(sc["bands"]))
        assert_same(O.run_batch(sc, dev, dur), run_gpu(sc, dev, dur))


def test_cuda_per_env_positions():
    rs = np.random.RandomState(33)
    sc = random_scenario(rs, nbands=4, jammers=1, spread=3.0)
    nenv, nsteps = 128, 50
    pos = np.zeros((nenv, 4, 8, 2))
    pos[:, :, :4, :] = rs.uniform(-4, 4, size=(nenv, 4, 4, 2))
    dev, dur = random_tapes(rs, nsteps, nenv, 4)
    o = O.run_batch(sc, dev, dur, pos=pos)
    g = run_gpu(sc, dev, dur, pos=pos)
    assert_same(o, g)
    # attenuation table of env 5, band 2 against the oracle's (reference formula)
    env = g["env"]
    att = env.read_state(8).cpu().numpy()           # [16, nsim]
    loc = dict(sc)
    one = O.Oracle(_with_positions(sc, pos[5]))
    for p in range(4):
        for d in range(4):
            if p != d:
                want = one.attenuation(2, p, d)
                got = att[p * 4 + d, 5 * 4 + 2]
                assert abs(got - want) <= BER_RTOL * abs(want)


@pytest.mark.parametrize("name", ["mobility_seed13", "mobility_inflight_seed17", "mobility_quirks_seed4"])
def test_cuda_moving_devices_match_reference_golden(name):
    """gw_set_positions between steps vs the reference's Position.set (goldens from the reference); in
    the second golden transmissions are on the air when devices move (SimplePhy._onAttenuationChange)."""
    doc = load_golden(name)
    moves = {int(k): [tuple(m) for m in v] for k, v in doc["moves"].items()}
    devs = doc["scenario"]["bands"][0]["devices"]
    pos = torch.zeros((1, 1, 4, 2), dtype=torch.float64)
    for d, dv in enumerate(devs):
        pos[0, 0, d, 0], pos[0, 0, d, 1] = dv["x"], dv["y"]
    env = make_env(doc["scenario"], 1, strict=False, positions=pos.cuda())
    if doc["do_reset"]:
        env.reset()
    for t, s in enumerate(doc["steps"]):
        if t in moves:
            for (_, d, x, y) in moves[t]:
                pos[0, 0, d, 0], pos[0, 0, d, 1] = x, y
            env.set_positions(pos.cuda())
        a = s["action"]
        o, r, dn, _ = env.step({"device": torch.tensor([a["device"]], dtype=torch.int32).cuda(),
                                "duration": torch.tensor([a["duration"]], dtype=torch.int32).cuda()})
        assert int(o[0]) == s["obs"] and float(r[0]) == s["reward"], t
        assert float(env.read_state(0)[0]) == s["now"], t
    n_rx = sum(1 for s in doc["steps"] for r in s["records"] if r[0] == "rx" and r[3] < 2)    # the MAC senders' packets
    assert int(env.delivered().sum()) == n_rx and n_rx > 0


@pytest.mark.parametrize("mode", ["reference", "mask_philox"])
def test_cuda_moving_devices_random_vs_oracle(mode):
    """Per-env random jumps before every other step while a PHY-only sender keeps the band busy."""
    rs = np.random.RandomState(7100)
    sc = random_scenario(rs, jammers=1, spread=2.5)
    sc["bands"][0]["devices"][3]["interval"] = 0.011
    n, T = 12, 50
    dev, dur = random_tapes(rs, T, n, 1)
    devs = sc["bands"][0]["devices"]
    pos = torch.zeros((n, 1, 4, 2), dtype=torch.float64)
    for d, dv in enumerate(devs):
        pos[:, 0, d, 0], pos[:, 0, d, 1] = dv["x"], dv["y"]
    moves = [{} for _ in range(n)]
    for e in range(n):
        for t in range(1, T, 2):
            ds = sorted(set(int(v) for v in rs.randint(4, size=int(rs.randint(1, 4)))))
            moves[e][t] = [(0, d, float(rs.uniform(-3, 3)), float(rs.uniform(-3, 3))) for d in ds]
    kw = {} if mode == "reference" else {"mode": mode, "seed": 91}
    env = make_env(sc, n, strict=False, positions=pos.cuda(), **kw)
    env.reset()
    got_obs = np.zeros((T, n), np.int64)
    got_now = np.zeros((T, n))
    for t in range(T):
        if t % 2 == 1:
            for e in range(n):
                for (_, d, x, y) in moves[e][t]:
                    pos[e, 0, d, 0], pos[e, 0, d, 1] = x, y
            env.set_positions(pos.cuda())
        o, r, dn, _ = env.step({"device": torch.as_tensor(dev[t, :, 0]).cuda(), "duration": torch.as_tensor(dur[t, :, 0]).cuda()})
        got_obs[t] = o.cpu().numpy()
        got_now[t] = env.read_state(0).cpu().numpy()
    env.check()
    for e in range(n):
        ora = O.Oracle(sc, mode=O.MODE_R if mode == "reference" else O.MODE_M)
        if mode != "reference":
            ora.use_philox_masks(91, e)
        acts = [{"device": int(dev[t, e, 0]), "duration": int(dur[t, e, 0])} for t in range(T)]
        res = O.run_tape(ora, acts, do_reset=True, moves=moves[e])
        assert [s["obs"] for s in res["steps"]] == list(got_obs[:, e]), e
        assert [s["now"] for s in res["steps"]] == list(got_now[:, e]), e


def test_cuda_moving_devices_corner_cases_vs_oracle():
    """Jumps beyond STANDBY_THRESHOLD, onto another device's position and back, from the first step on
    (lazily created attenuation models): move_kernel vs the oracle, one env per random move sequence."""
    rs = np.random.RandomState(7400)
    sc = random_scenario(rs, jammers=1, spread=3.0)
    sc["bands"][0]["devices"][3]["interval"] = 0.0125
    n, T = 10, 60
    dev, dur = random_tapes(rs, T, n, 1)
    devs = sc["bands"][0]["devices"]
    pos = torch.zeros((n, 1, 4, 2), dtype=torch.float64)
    for d, dv in enumerate(devs):
        pos[:, 0, d, 0], pos[:, 0, d, 1] = dv["x"], dv["y"]
    moves = [{} for _ in range(n)]
    for e in range(n):
        cur = [(dv["x"], dv["y"]) for dv in devs]
        for t in range(1, T, 2):
            d = int(rs.randint(4))
            kind = int(rs.randint(4))
            if kind == 0:
                x, y = float(rs.uniform(4000, 6000)), float(rs.uniform(-10, 10))
            elif kind == 1:
                x, y = cur[int((d + 1 + rs.randint(3)) % 4)]
            else:
                x, y = float(rs.uniform(-3, 3)), float(rs.uniform(-3, 3))
            cur[d] = (x, y)
            moves[e][t] = [(0, d, x, y)]
    env = make_env(sc, n, strict=False, positions=pos.cuda())
    env.reset()
    got_obs = np.zeros((T, n), np.int64)
    got_now = np.zeros((T, n))
    for t in range(T):
        if t % 2 == 1:
            for e in range(n):
                for (_, d, x, y) in moves[e][t]:
                    pos[e, 0, d, 0], pos[e, 0, d, 1] = x, y
            env.set_positions(pos.cuda())
        o, r, dn, _ = env.step({"device": torch.as_tensor(dev[t, :, 0]).cuda(), "duration": torch.as_tensor(dur[t, :, 0]).cuda()})
        got_obs[t] = o.cpu().numpy()
        got_now[t] = env.read_state(0).cpu().numpy()
    env.check()
    for e in range(n):
        acts = [{"device": int(dev[t, e, 0]), "duration": int(dur[t, e, 0])} for t in range(T)]
        res = O.run_tape(O.Oracle(sc), acts, do_reset=True, moves=moves[e])
        assert [s["obs"] for s in res["steps"]] == list(got_obs[:, e]), e
        assert [s["now"] for s in res["steps"]] == list(got_now[:, e]), e


def _with_positions(sc, pos_env):
    import copy
    sc = copy.deepcopy(sc)
    for b, band in enumerate(sc["bands"]):
        for d, dv in enumerate(band["devices"]):
            dv["x"], dv["y"] = float(pos_env[b, d, 0]), float(pos_env[b, d, 1])
    return sc


def test_reset_mid_run_matches_oracle():
    """reset() between steps: counters := 0, queued packets keep their sizes (snapshot ring)."""
    from gymwipe_b200.scenario import default_scenario_dict
    sc = default_scenario_dict()
    rs = np.random.RandomState(5)
    nenv = 64
    acts = [(rs.randint(0, 2, nenv).astype(np.int32), rs.randint(0, 20, nenv).astype(np.int32)) for _ in range(90)]
    oras = [O.Oracle(sc) for _ in range(nenv)]
    env = make_env(sc, nenv, strict=False)
    for i, (dv, du) in enumerate(acts):
        if i in (0, 30, 31, 60):
            for o in oras:
                o.reset()
            env.reset()
        want = [o.step({"device": int(dv[e]), "duration": int(du[e])}) for e, o in enumerate(oras)]
        ob, rw, dn, _ = env.step({"device": torch.as_tensor(dv).cuda(), "duration": torch.as_tensor(du).cuda()})
        assert ob.cpu().tolist() == [w[0] for w in want]
        assert rw.cpu().tolist() == [w[1] for w in want]
        assert env.read_state(0).cpu().tolist() == [o.now for o in oras]
    # partial reset (env_ids)
    ids = [3, 17, 40]
    for e in ids:
        oras[e].reset()
    env.reset(env_ids=ids)
    for dv, du in acts[:20]:
        want = [o.step({"device": int(dv[e]), "duration": int(du[e])}) for e, o in enumerate(oras)]
        ob, rw, _, _ = env.step({"device": torch.as_tensor(dv).cuda(), "duration": torch.as_tensor(du).cuda()})
        assert ob.cpu().tolist() == [w[0] for w in want]
        assert rw.cpu().tolist() == [w[1] for w in want]


def test_ber_and_fspl_kernels_numeric():
    doc = load_golden("arithmetic")
    from gymwipe_b200.networking.physical import ber_from_milliwatts
    from gymwipe_b200.networking.attenuation_models import FsplAttenuation
    v = np.array(doc["ber_mw"])
    got = ber_from_milliwatts(torch.tensor(v[:, 0]).cuda(), torch.tensor(v[:, 1]).cuda()).cpu().numpy()
    assert np.all(np.abs(got - v[:, 2]) <= BER_RTOL * np.abs(v[:, 2]))
    print("max BER rel err vs reference: %.3e" % np.max(np.abs(got - v[:, 2]) / np.abs(v[:, 2])))
    f = np.array(doc["fspl"])
    for freq in np.unique(f[:, 4]):
        m = f[:, 4] == freq
        a = FsplAttenuation.attenuation(*(torch.tensor(f[m, k]).cuda() for k in range(4)), freq).cpu().numpy()
        assert np.all(np.abs(a - f[m, 5]) <= 1e-12 * np.abs(f[m, 5]))
    # live-PHY values of SURVEY appendix C through the received-power state
    env = make_env()
    env.step({"device": 0, "duration": 3})
    srx = env.read_state(9).cpu().numpy()[:, 0]
    assert abs(srx[0 * 4 + 2] - 2.4689797345652838e-05) <= 1e-12 * 2.4689797345652838e-05


def test_action_validation_flag():
    env = make_env(None, 32, strict=False)
    dev = torch.zeros(32, dtype=torch.int32, device="cuda")
    dur = torch.full((32,), 20, dtype=torch.int32, device="cuda")       # 20 is outside Discrete(20)
    env.step({"device": dev, "duration": dur})
    with pytest.raises(ValueError, match="action"):
        env.check()
    env.check()                                                          # flag is cleared


def test_flat_action_like_dqn_processor():
    """agents/dqn_counter_traffic.py:25-33: a -> (a // 20, a % 20)."""
    e1 = make_env(None, 8, strict=False)
    e2 = make_env(None, 8, strict=False)
    a = torch.tensor([0, 3, 19, 20, 25, 39, 12, 33], device="cuda")
    o1, r1, _, _ = e1.step(a)
    o2, r2, _, _ = e2.step({"device": (a // 20).int(), "duration": (a % 20).int()})
    assert torch.equal(o1, o2) and torch.equal(r1, r2)


def test_step_host_end_to_end():
    from gymwipe_b200.scenario import default_scenario_dict
    sc = default_scenario_dict()
    rs = np.random.RandomState(8)
    n = 2048
    dev, dur = random_tapes(rs, 20, n, 1)
    o = O.run_batch(sc, dev, dur)
    env = make_env(sc, n, strict=False)
    env.reset()
    obs = torch.empty(n, dtype=torch.int64).pin_memory()
    rew = torch.empty(n, dtype=torch.float64).pin_memory()
    done = torch.empty(n, dtype=torch.uint8).pin_memory()
    for t in range(20):
        env.step_host(np.ascontiguousarray(dev[t, :, 0]), np.ascontiguousarray(dur[t, :, 0]), obs, rew, done)
        assert (obs.numpy() == o["obs"][t, :, 0]).all()
        assert (rew.numpy() == o["reward"][t, :, 0]).all()


def test_step_host_packed_end_to_end():
    from gymwipe_b200.scenario import default_scenario_dict
    sc = default_scenario_dict()
    rs = np.random.RandomState(18)
    n = 4096
    dev, dur = random_tapes(rs, 20, n, 1)
    o = O.run_batch(sc, dev, dur)
    env = make_env(sc, n, strict=False)
    env.reset()
    act = torch.empty((2, n), dtype=torch.int32).pin_memory()
    res = torch.empty(9 * n, dtype=torch.uint8).pin_memory()
    for t in range(20):
        act[0].copy_(torch.as_tensor(dev[t, :, 0]))
        act[1].copy_(torch.as_tensor(dur[t, :, 0]))
        env.step_host_packed(act, res)
        obs, rew, done = env.unpack_results(res)
        assert (obs.numpy() == o["obs"][t, :, 0]).all()
        assert (rew.numpy().astype(np.float64) == o["reward"][t, :, 0]).all()
        assert not done.numpy().any()


@pytest.mark.parametrize("n,pinned", [(37, True), (1024, False), (3000, True)])
def test_step_host_compact_end_to_end(n, pinned):
    """gw_step_host_compact: uint8 actions in, one packed word per sim out; pinned buffers are read and
    written in place by the kernel, pageable ones are staged."""
    from gymwipe_b200.scenario import default_scenario_dict
    sc = default_scenario_dict()
    rs = np.random.RandomState(33)
    T = 60
    dev, dur = random_tapes(rs, T, n, 1)
    ref = O.run_batch(sc, dev, dur)
    env = make_env(sc, n, strict=False)
    env.reset()
    act = torch.empty((n, 2), dtype=torch.uint8)
    res = torch.empty(n, dtype=torch.int32)
    if pinned:
        act, res = act.pin_memory(), res.pin_memory()
    for t in range(T):
        act[:, 0] = torch.from_numpy(dev[t, :, 0].astype(np.uint8))
        act[:, 1] = torch.from_numpy(dur[t, :, 0].astype(np.uint8))
        res.fill_(-1)
        env.step_host_compact(act, res)
        obs, rew, done = env.unpack_compact(res)
        assert (obs.numpy() == ref["obs"][t, :, 0]).all()
        assert (rew.numpy() == ref["reward"][t, :, 0]).all()
        assert (done.numpy() == ref["done"][t, :, 0].astype(bool)).all()
    env.check()
    assert (env.read_state(0).cpu().numpy() == ref["now"][-1]).all()
    # an action outside the action space is flagged like in gw_step
    act[0, 0] = 7
    env.step_host_compact(act, res)
    with pytest.raises(ValueError):
        env.check()


def test_step_host_compact_async_two_batches_in_flight():
    """gw_step_host_compact_async: two env batches in flight on one stream, results valid after their events."""
    from gymwipe_b200.scenario import default_scenario_dict
    sc = default_scenario_dict()
    rs = np.random.RandomState(34)
    n, T = 777, 40
    tapes = [random_tapes(rs, T, n, 1) for _ in range(2)]
    refs = [O.run_batch(sc, d, u) for d, u in tapes]
    envs = [make_env(sc, n, strict=False) for _ in range(2)]
    acts = [[torch.from_numpy(np.stack([d[t, :, 0], u[t, :, 0]], axis=1).astype(np.uint8)).pin_memory() for t in range(T)]
            for d, u in tapes]
    res = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(2)]
    evs = [torch.cuda.Event() for _ in range(2)]
    for e in envs:
        e.reset()
    for t in range(T):
        for b in range(2):
            if t > 0:
                evs[b].synchronize()
                obs, rew, done = envs[b].unpack_compact(res[b])
                assert (obs.numpy() == refs[b]["obs"][t - 1, :, 0]).all()
                assert (rew.numpy() == refs[b]["reward"][t - 1, :, 0]).all()
            envs[b].step_host_compact_async(acts[b][t], res[b])
            evs[b].record()
    torch.cuda.synchronize()
    for b in range(2):
        envs[b].check()
        assert (envs[b].unpack_compact(res[b])[0].numpy() == refs[b]["obs"][-1, :, 0]).all()
    # pageable buffers are refused (the asynchronous form never stages)
    with pytest.raises(Exception):
        envs[0].step_host_compact_async(np.zeros((n, 2), np.uint8), np.zeros(n, np.int32))


def test_stats_epilogue():
    from gymwipe_b200.scenario import default_scenario_dict
    sc = default_scenario_dict()
    rs = np.random.RandomState(9)
    n, T = 1000, 30
    dev, dur = random_tapes(rs, T, n, 1)
    o = O.run_batch(sc, dev, dur)
    env = make_env(sc, n, strict=False)
    env.reset()
    env.stats()
    tot = np.zeros(8)
    for t in range(T):
        env.step({"device": torch.as_tensor(dev[t, :, 0]).cuda(), "duration": torch.as_tensor(dur[t, :, 0]).cuda()})
        tot += env.stats().cpu().numpy()
    assert tot[0] == o["reward"].sum()
    assert tot[1] == o["counts"][:, 0, 1].sum() and tot[2] == o["counts"][:, 0, 2].sum()
    assert tot[4] == n * T and tot[6] == o["counts"][:, 0, 0].sum()


def test_shared_stats_across_handles():
    """gw_share_stats: two env batches accumulate into one statistics vector."""
    from gymwipe_b200.scenario import default_scenario_dict
    sc = default_scenario_dict()
    rs = np.random.RandomState(19)
    n, T = 300, 20
    tapes = [random_tapes(rs, T, n, 1) for _ in range(2)]
    refs = [O.run_batch(sc, d, u) for d, u in tapes]
    envs = [make_env(sc, n, strict=False) for _ in range(2)]
    for e in envs:
        e.reset()
    envs[1].share_stats(envs[0])
    envs[0].stats()
    for t in range(T):
        for e, (d, u) in zip(envs, tapes):
            e.step({"device": torch.as_tensor(d[t, :, 0]).cuda(), "duration": torch.as_tensor(u[t, :, 0]).cuda()})
    tot = envs[0].stats().cpu().numpy()
    assert tot[0] == sum(r["reward"].sum() for r in refs)
    assert tot[4] == 2 * n * T and tot[6] == sum(r["counts"][:, 0, 0].sum() for r in refs)
    assert envs[1].stats().cpu().numpy()[4] == 0            # the shared vector was cleared by the read above
    envs[1].share_stats(None)
    envs[1].step({"device": torch.as_tensor(tapes[1][0][0, :, 0]).cuda(), "duration": torch.as_tensor(tapes[1][1][0, :, 0]).cuda()})
    assert envs[1].stats().cpu().numpy()[4] == n and envs[0].stats().cpu().numpy()[4] == 0


def test_full_size_properties_65536():
    """BASELINE configs[1]: 65,536 envs.  Size-independent properties + a strided oracle sample."""
    from gymwipe_b200.scenario import default_scenario_dict
    sc = default_scenario_dict()
    n, T = 65536, 128
    g = torch.Generator(device="cuda").manual_seed(0)
    dev = torch.randint(0, 2, (T, n), generator=g, device="cuda", dtype=torch.int32)
    dur = torch.randint(0, 20, (T, n), generator=g, device="cuda", dtype=torch.int32)
    env = make_env(sc, n, strict=False)
    env.reset()
    env.stats()
    obs_all = torch.empty((T, n), dtype=torch.int64, device="cuda")
    rew_all = torch.empty((T, n), dtype=torch.float64, device="cuda")
    prev_abs = torch.zeros(n, dtype=torch.int64, device="cuda")
    for t in range(T):
        o, r, d, _ = env.step({"device": dev[t], "duration": dur[t]})
        obs_all[t], rew_all[t] = o, r
        diff = o - 65536
        assert bool(((diff == -2) | (diff == 0) | (diff == 2)).all())     # appendix B #1
        assert torch.equal(r, (prev_abs - diff.abs()).double())            # reward = change of |difference|
        assert not bool(d.any())
        prev_abs = diff.abs()
    env.check()
    st = env.stats().cpu().numpy()
    assert st[4] == n * T and st[0] == rew_all.sum().item()
    deliv = env.delivered()
    assert st[1] == deliv[:, 0].sum().item() and st[2] == deliv[:, 1].sum().item()
    # identical actions => identical trajectories: envs with the same tape must agree (checksum)
    # and a strided sample of 512 envs equals the oracle bit for bit
    idx = np.arange(0, n, n // 512)
    o = O.run_batch(sc, dev[:, idx].cpu().numpy(), dur[:, idx].cpu().numpy())
    assert (obs_all[:, idx].cpu().numpy() == o["obs"][:, :, 0]).all()
    assert (rew_all[:, idx].cpu().numpy() == o["reward"][:, :, 0]).all()
    assert (env.read_state(0)[idx].cpu().numpy() == o["now"][-1]).all()


def test_long_run_through_counter_saturation():
    """
    Maximum sizes: 7 000 steps of ~10 ms each cross COUNTER_BOUND = 65536 ticks (the counter
    saturates, counter_traffic.py:59-60; packets reach 64 KiB), the queue is in permanent
    drop-oldest overflow, and simulated time passes 70 s (slot alignment at large times).
    """
    from gymwipe_b200.scenario import default_scenario_dict
    sc = default_scenario_dict()
    rs = np.random.RandomState(77)
    n, T = 16, 7000
    dev, dur = random_tapes(rs, T, n, 1)
    dur[:] = np.maximum(dur, 8)                       # long assignments: >= 8 ms per step
    o = O.run_batch(sc, dev, dur, want=("obs", "reward", "now", "counts"))
    env = make_env(sc, n, strict=False)
    env.reset()
    d_dev, d_dur = torch.as_tensor(dev[:, :, 0]).cuda(), torch.as_tensor(dur[:, :, 0]).cuda()
    obs = torch.empty((T, n), dtype=torch.int64, device="cuda")
    for t in range(T):
        ob, rw, dn, _ = env.step({"device": d_dev[t], "duration": d_dur[t]})
        obs[t] = ob
    env.check()
    assert (obs.cpu().numpy() == o["obs"][:, :, 0]).all()
    now = env.read_state(0).cpu().numpy()
    assert (now == o["now"][-1]).all() and now.min() > 70.0
    assert (env.read_state(3).cpu().numpy() == 65536).all()              # SenderDevice.counter saturated
    assert (env.read_state(4).cpu().numpy() == 100).all()                # deque(maxlen=100) is full
    assert (env.transmissions().cpu().numpy() == o["counts"][:, 0, 0]).all()
    assert (env.delivered().cpu().numpy() == o["counts"][:, 0, 1:3]).all()


def test_send_queue_overflow_is_reported():
    """A PHY-only sender whose interval is shorter than its airtime queues SEND commands without
    bound in the reference; both the oracle and the CUDA path report it instead of mis-simulating."""
    sc = {"assignment_duration_factor": 1000, "bands": [{"frequency": 2.4e9, "bandwidth": 22e6, "devices": [
        {"role": "sender", "x": 0.0, "y": 2.0, "mult": 1, "payload": "counter", "interval": 0.001, "dest": 1},
        {"role": "sender", "x": 0.0, "y": -2.0, "mult": 3, "payload": "counter", "interval": 0.001, "dest": 0},
        {"role": "rrm", "x": 0.0, "y": 0.0},
        {"role": "jammer", "x": 3.0, "y": 0.0, "interval": 0.001, "delay": 0.0, "power": 0.0, "hdr": 13, "payload": 200}]}]}
    dev = np.zeros((400, 1, 1), np.int32)
    dur = np.full((400, 1, 1), 19, np.int32)
    with pytest.raises(O.OracleFault):
        O.run_batch(sc, dev, dur)
    env = make_env(sc, 4, strict=False)
    a = {"device": torch.zeros(4, dtype=torch.int32, device="cuda"), "duration": torch.full((4,), 19, dtype=torch.int32, device="cuda")}
    from gymwipe_b200._native import NativeError
    with pytest.raises(NativeError, match="SEND queue overflow"):
        for _ in range(400):
            env.step(a)
        env.check()


def test_cuda_mac_receive_known_answers():
    """The reference's MAC known-answer test (tests/networking/test_stack.py:218-235: 4 / 4 / 8 / 8 / 10 / 10 packets
    received after successive assignment rounds) on the GPU, through the env API: both MACs in receive mode, two
    finite 10-packet bursts, ten alternating 10 ms assignments."""
    from test_core_host import MAC_KAT_RECEIVED, mac_kat_scenario
    import gymwipe_b200
    n = 64
    env = gymwipe_b200.make('CounterTraffic-v0', num_envs=n, scenario=mac_kat_scenario(), strict=False)
    ora = O.Oracle(mac_kat_scenario())
    for t in range(10):
        env.step({"device": torch.full((n,), t % 2, dtype=torch.int32, device="cuda"),
                  "duration": torch.full((n,), 10, dtype=torch.int32, device="cuda")})
        ora.step({"device": t % 2, "duration": 10})
        got = env.received().cpu().numpy()
        assert (got == np.array(MAC_KAT_RECEIVED[t])).all()
        assert (env.read_state(0).cpu().numpy() == ora.now).all()
    env.check()


@pytest.mark.parametrize("seed", range(2))
def test_cuda_receive_mode_and_bursts_match_oracle(seed):
    rs = np.random.RandomState(8900 + seed)
    for jam in (0, 1):
        sc = random_scenario(rs, jammers=jam, spread=2.0)
        sc["bands"][0]["devices"][seed % 2]["receive"] = True
        sc["bands"][0]["devices"][1 - seed % 2]["receive"] = bool(rs.randint(2))
        sc["bands"][0]["devices"][int(rs.randint(2))]["max_ticks"] = int(rs.randint(5, 60))
        nenv, nsteps = 128, 70
        dev, dur = random_tapes(rs, nsteps, nenv, 1)
        o = O.run_batch(sc, dev, dur)
        import gymwipe_b200
        env = gymwipe_b200.make('CounterTraffic-v0', num_envs=nenv, scenario=sc, strict=False)
        env.reset()
        for t in range(nsteps):
            ob, rw, dn, _ = env.step({"device": torch.as_tensor(dev[t, :, 0]).cuda(), "duration": torch.as_tensor(dur[t, :, 0]).cuda()})
            assert (ob.cpu().numpy() == o["obs"][t, :, 0]).all() and (rw.cpu().numpy() == o["reward"][t, :, 0]).all()
        env.check()
        assert (env.read_state(0).cpu().numpy() == o["now"][-1]).all()
        ora = O.Oracle(sc)
        ora.reset()
        for t in range(nsteps):
            ora.step({"device": int(dev[t, 0, 0]), "duration": int(dur[t, 0, 0])})
        assert tuple(env.received()[0].tolist()) == tuple(ora.received()[:2])
