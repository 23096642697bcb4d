"""
General band engine (``gymwipe_b200/csrc/gw_band.cuh``): bands beyond ``CounterTrafficEnv``'s 2 senders + RRM
template -- 3..8 MAC senders, the RRM, 0..16 PHY-only senders.

CPU: the oracle's C restatement and the host build of the CUDA engine against golden vectors produced by the
UNMODIFIED reference (``oracle/gen_golden.py``, cases ``nsenders_*``: the reference's own ``SimpleNetworkDevice`` /
``SimpleRrmDevice`` / ``CounterTrafficInterpreter`` classes composed to larger bands by ``oracle/ref_harness.py``), and
against each other on random scenarios.  GPU: the CUDA engine through the C ABI (``gw_genband_*``) against the
goldens and, at batch sizes the reference could never run, against the oracle.  Integer results, event times and
event order: bit-exact; BER values / expected error sums: 1e-9 relative (stated tolerance 1e-6).
"""
import numpy as np
import pytest

import gw_oracle as O
import hostsim as HS
from util import GOLDEN_NSENDERS, assert_step_records, load_golden, random_scenario_n


def _tapes(doc):
    dev = np.array([[s["action"]["device"]] for s in doc["steps"]], np.int32)
    dur = np.array([[s["action"]["duration"]] for s in doc["steps"]], np.int32)
    return dev, dur


def _oracle_tape(sc, dev, dur, do_reset=True, reset_at=None):
    ora = O.Oracle(sc, trace=True)
    if do_reset and reset_at is None:
        ora.reset()
    ora.take_records()
    steps = []
    for t in range(dev.shape[0]):
        if reset_at is not None and t == reset_at:
            ora.reset()
        obs, rew, done = ora.step({"device": int(dev[t, 0]), "duration": int(dur[t, 0])})
        steps.append({"obs": obs, "reward": rew, "done": done, "now": ora.now, "records": ora.take_records()})
    ntx, nd = ora.counts()
    return steps, ntx, nd, ora.received(), ora.near_ties


def _assert_host_equals(h, steps, ntx, nd, nrecv, ns, label):
    assert h["rc"] == 0, label
    for t, s in enumerate(steps):
        assert h["obs"][t, 0] == s["obs"] and h["reward"][t, 0] == s["reward"] and bool(h["done"][t, 0]) == s["done"], (label, t)
        assert h["now"][t, 0] == s["now"], (label, t, h["now"][t, 0], s["now"])
        assert_step_records(h["records"][t], s["records"], "%s step %d" % (label, t))
    assert h["counts"][0, 0] == ntx, label
    assert list(h["counts"][0, 1:1 + ns]) == list(nd[:ns]), label
    assert list(h["counts"][0, 9:9 + ns]) == list(nrecv[:ns]), label


@pytest.mark.parametrize("name", GOLDEN_NSENDERS)
def test_oracle_and_core_match_reference_golden_nsenders(name):
    doc = load_golden(name)
    sc = doc["scenario"]
    ns = sum(1 for d in sc["bands"][0]["devices"] if d["role"] == "sender")
    dev, dur = _tapes(doc)
    steps, ntx, nd, nrecv, _ = _oracle_tape(sc, dev, dur, do_reset=doc["do_reset"])
    h = HS.gen_run(sc, dev, dur, do_reset=doc["do_reset"])
    assert h["rc"] == 0
    for t, g in enumerate(doc["steps"]):
        # the restatement: every record identical (as in tests/test_oracle_golden.py)
        o = steps[t]
        assert (o["obs"], o["reward"], o["done"], o["now"]) == (g["obs"], g["reward"], g["done"], g["now"]), (name, t)
        assert_step_records(o["records"], g["records"], "oracle %s step %d" % (name, t))
        # the engine
        assert (h["obs"][t, 0], h["reward"][t, 0], bool(h["done"][t, 0]), h["now"][t, 0]) == (g["obs"], g["reward"], g["done"], g["now"]), (name, t)
        assert_step_records(h["records"][t], g["records"], "core %s step %d" % (name, t))
    n_rx = [sum(1 for s in doc["steps"] for r in s["records"] if r[0] == "rx" and r[3] == k) for k in range(ns)]
    n_mrx = [sum(1 for s in doc["steps"] for r in s["records"] if r[0] == "mrx" and r[3] == k) for k in range(ns)]
    assert list(h["counts"][0, 1:1 + ns]) == n_rx and list(h["counts"][0, 9:9 + ns]) == n_mrx
    assert sum(n_rx) > 0


@pytest.mark.parametrize("seed", range(10))
def test_core_general_random_vs_oracle(seed):
    rs = np.random.RandomState(4000 + seed)
    ns, nj = int(rs.randint(3, 9)), int(rs.randint(0, 17))
    sc = random_scenario_n(rs, ns, nj, spread=float(rs.choice([1.5, 2.5, 6.0])), receive=bool(seed % 2), bursts=bool(seed % 3 == 0))
    T = 70
    dev = rs.randint(0, ns, size=(T, 1)).astype(np.int32)
    dur = rs.randint(0, 20, size=(T, 1)).astype(np.int32)
    reset_at = None if seed % 4 else 25                  # env.reset() in the middle of a run: queued packets keep their sizes
    steps, ntx, nd, nrecv, near = _oracle_tape(sc, dev, dur, do_reset=seed % 4 != 1, reset_at=reset_at)
    h = HS.gen_run(sc, dev, dur, do_reset=seed % 4 != 1, reset_at=reset_at)
    _assert_host_equals(h, steps, ntx, nd, nrecv, ns, "seed %d (%d senders, %d PHY-only)" % (seed, ns, nj))


def test_core_general_default_scenario_equals_counter_traffic_env_golden():
    """With two senders the general engine is CounterTrafficEnv: the golden of the reference's own class."""
    doc = load_golden("default_reset_seed0")
    dev, dur = _tapes(doc)
    h = HS.gen_run(doc["scenario"], dev, dur, do_reset=True)
    assert h["rc"] == 0
    for t, g in enumerate(doc["steps"]):
        assert (h["obs"][t, 0], h["reward"][t, 0], bool(h["done"][t, 0]), h["now"][t, 0]) == (g["obs"], g["reward"], g["done"], g["now"]), t
        assert_step_records(h["records"][t], g["records"], "default step %d" % t)


def test_core_general_batch_with_per_env_positions():
    """Several band-sims side by side in the device layout ([word][sim]) with per-env geometries."""
    rs = np.random.RandomState(77)
    ns, nj, nenv, T = 4, 3, 6, 40
    sc = random_scenario_n(rs, ns, nj, spread=2.0, receive=True)
    nd = ns + 1 + nj
    pos = rs.uniform(-2.5, 2.5, size=(nenv, nd, 2))
    dev = rs.randint(0, ns, size=(T, nenv)).astype(np.int32)
    dur = rs.randint(0, 20, size=(T, nenv)).astype(np.int32)
    h = HS.gen_run(sc, dev, dur, pos=pos)
    assert h["rc"] == 0
    for e in range(nenv):
        sce = {"assignment_duration_factor": 1000, "bands": [dict(sc["bands"][0], devices=[dict(d, x=float(pos[e, i, 0]), y=float(pos[e, i, 1]))
                                                                                           for i, d in enumerate(sc["bands"][0]["devices"])])]}
        steps, ntx, ndl, nrecv, _ = _oracle_tape(sce, dev[:, e:e + 1], dur[:, e:e + 1])
        for t, s in enumerate(steps):
            assert h["obs"][t, e] == s["obs"] and h["reward"][t, e] == s["reward"] and h["now"][t, e] == s["now"], (e, t)
        assert h["counts"][e, 0] == ntx and list(h["counts"][e, 1:1 + ns]) == list(ndl[:ns]) and list(h["counts"][e, 9:9 + ns]) == list(nrecv[:ns])


def test_core_general_rejects_actions_outside_the_action_space():
    rs = np.random.RandomState(5)
    sc = random_scenario_n(rs, 3, 0)
    h = HS.gen_run(sc, np.array([[3]], np.int32), np.array([[1]], np.int32))
    assert h["rc"] != 0
    h = HS.gen_run(sc, np.array([[1]], np.int32), np.array([[20]], np.int32))
    assert h["rc"] != 0


# ------------------------------------------------------------------------------------------------------------
# GPU: the CUDA engine through the C ABI
# ------------------------------------------------------------------------------------------------------------

def _gpu_env(sc, n, **kw):
    import gymwipe_b200
    return gymwipe_b200.make('CounterTraffic-v0', num_envs=n, device="cuda:0", scenario=sc, strict=False, **kw)


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN_NSENDERS)
def test_cuda_general_band_matches_reference_golden(name):
    import torch
    from gymwipe_b200.envs import GeneralBandEnv
    doc = load_golden(name)
    sc = doc["scenario"]
    ns = sum(1 for d in sc["bands"][0]["devices"] if d["role"] == "sender")
    env = _gpu_env(sc, 1)
    assert isinstance(env, GeneralBandEnv) and env.n_senders == ns
    if doc["do_reset"]:
        assert env.reset() == doc["reset_obs"]
    worst = 0.0
    for t, g in enumerate(doc["steps"]):
        obs, rew, done, recs = env.step_traced({"device": g["action"]["device"], "duration": g["action"]["duration"]})
        assert (obs, rew, done) == (g["obs"], g["reward"], g["done"]), (name, t)
        assert float(env.now[0]) == g["now"], (name, t)
        worst = max(worst, assert_step_records(recs, g["records"], "%s step %d" % (name, t)))
    n_rx = [sum(1 for s in doc["steps"] for r in s["records"] if r[0] == "rx" and r[3] == k) for k in range(ns)]
    n_mrx = [sum(1 for s in doc["steps"] for r in s["records"] if r[0] == "mrx" and r[3] == k) for k in range(ns)]
    assert env.delivered()[0].tolist() == n_rx and env.received()[0].tolist() == n_mrx
    # the interpreter's receivedValues (payload.value = 2 for every sender the RRM has decoded since the reset)
    assert env.received_values()[0].tolist() == [2 if c > 0 else 0 for c in n_rx]
    info = env.step({"device": 0, "duration": 0})[3]              # (num_envs == 1: the reference's info dict)
    assert set(info) == {"Latest received values"} and info["Latest received values"].startswith("[")
    print("%s: max relative BER / error-sum deviation vs reference %.3e" % (name, worst))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(4))
def test_cuda_general_band_matches_oracle_batch(seed):
    """2048 envs x 48 steps of a random N-sender band against the oracle: obs / reward / done / step end times and
    the delivery counts per sender bit-exact."""
    import torch
    rs = np.random.RandomState(5200 + seed)
    ns, nj = int(rs.randint(3, 9)), int(rs.randint(0, 17))
    sc = random_scenario_n(rs, ns, nj, spread=float(rs.choice([1.5, 2.5])), receive=bool(seed % 2), bursts=bool(seed == 3))
    nenv, T = 2048, 48
    dev = rs.randint(0, ns, size=(T, nenv)).astype(np.int32)
    dur = rs.randint(0, 20, size=(T, nenv)).astype(np.int32)
    o = O.run_batch(sc, dev, dur, do_reset=seed != 2)
    env = _gpu_env(sc, nenv)
    if seed != 2:
        env.reset()
    for t in range(T):
        obs, rew, done, _ = env.step({"device": torch.as_tensor(dev[t]).cuda(), "duration": torch.as_tensor(dur[t]).cuda()})
        assert (obs.cpu().numpy() == o["obs"][t, :, 0]).all(), (seed, t)
        assert (rew.cpu().numpy() == o["reward"][t, :, 0]).all(), (seed, t)
        assert (done.cpu().numpy() == o["done"][t, :, 0].astype(bool)).all(), (seed, t)
        assert (env.now.cpu().numpy() == o["now"][t]).all(), (seed, t)
    env.check()
    assert (env.transmissions().cpu().numpy() == o["counts"][:, 0, 0]).all()
    assert (env.delivered().cpu().numpy() == o["counts"][:, 0, 1:1 + ns]).all()
    assert o["counts"][:, 0, 1:1 + ns].sum() > 1000


@pytest.mark.gpu
def test_cuda_general_band_per_env_positions_and_reset():
    import torch
    rs = np.random.RandomState(91)
    ns, nj, nenv, T = 5, 4, 8, 30
    sc = random_scenario_n(rs, ns, nj, spread=2.0, receive=True)
    nd = ns + 1 + nj
    pos = rs.uniform(-2.5, 2.5, size=(nenv, nd, 2))
    dev = rs.randint(0, ns, size=(T, nenv)).astype(np.int32)
    dur = rs.randint(0, 20, size=(T, nenv)).astype(np.int32)
    env = _gpu_env(sc, nenv, positions=torch.as_tensor(pos))
    got = {k: [] for k in ("obs", "reward", "now")}
    for t in range(T):
        if t == 12:
            env.reset()
        obs, rew, done, _ = env.step({"device": torch.as_tensor(dev[t]).cuda(), "duration": torch.as_tensor(dur[t]).cuda()})
        got["obs"].append(obs.cpu().numpy().copy()); got["reward"].append(rew.cpu().numpy().copy()); got["now"].append(env.now.cpu().numpy())
    env.check()
    deliv, recvd = env.delivered().cpu().numpy(), env.received().cpu().numpy()
    for e in range(nenv):
        sce = {"assignment_duration_factor": 1000, "bands": [dict(sc["bands"][0], devices=[dict(d, x=float(pos[e, i, 0]), y=float(pos[e, i, 1]))
                                                                                           for i, d in enumerate(sc["bands"][0]["devices"])])]}
        steps, ntx, ndl, nrecv, _ = _oracle_tape(sce, dev[:, e:e + 1], dur[:, e:e + 1], do_reset=False, reset_at=12)
        for t, s in enumerate(steps):
            assert got["obs"][t][e] == s["obs"] and got["reward"][t][e] == s["reward"] and got["now"][t][e] == s["now"], (e, t)
        assert list(deliv[e]) == list(ndl[:ns]) and list(recvd[e]) == list(nrecv[:ns])


@pytest.mark.gpu
def test_cuda_general_band_two_senders_equal_the_step_kernel():
    """ns = 2: the general engine and CounterTrafficEnv's fused step kernel are two implementations of the same env."""
    import torch
    import gymwipe_b200
    from gymwipe_b200.envs import GeneralBandEnv
    from gymwipe_b200.scenario import default_scenario_dict
    sc = default_scenario_dict()
    rs = np.random.RandomState(3)
    nenv, T = 1024, 40
    a = gymwipe_b200.make('CounterTraffic-v0', num_envs=nenv, device="cuda:0", strict=False)
    b = GeneralBandEnv(num_envs=nenv, device="cuda:0", scenario=sc, strict=False)
    a.reset(); b.reset()
    for t in range(T):
        act = {"device": torch.as_tensor(rs.randint(0, 2, nenv).astype(np.int32)).cuda(),
               "duration": torch.as_tensor(rs.randint(0, 20, nenv).astype(np.int32)).cuda()}
        oa, ra, da, _ = a.step(act)
        ob, rb, db, _ = b.step(act)
        assert torch.equal(oa.reshape(-1), ob) and torch.equal(ra.reshape(-1), rb), t
    assert torch.equal(a.read_state(0).reshape(-1), b.now)
    assert torch.equal(a.delivered().reshape(nenv, 2), b.delivered())


@pytest.mark.gpu
def test_cuda_general_band_rejects_bad_actions_and_bad_configs():
    import torch
    from gymwipe_b200 import _native as N
    from gymwipe_b200.envs import GeneralBandEnv
    rs = np.random.RandomState(5)
    sc = random_scenario_n(rs, 3, 1)
    env = GeneralBandEnv(num_envs=4, device="cuda:0", scenario=sc, strict=True)
    env.reset()
    ok = {"device": torch.tensor([0, 1, 2, 0], dtype=torch.int32).cuda(), "duration": torch.tensor([1, 2, 3, 4], dtype=torch.int32).cuda()}
    env.step(ok)
    now = env.now.clone()
    with pytest.raises(N.NativeError) as e:
        env.step({"device": torch.tensor([0, 3, 2, 0], dtype=torch.int32).cuda(), "duration": ok["duration"]})
    assert e.value.code == N.GW_E_ACTION
    assert env.now[1] == now[1] and env.faults().tolist() == [0, 0, 0, 0]      # the rejected env was not stepped
    env.step(ok)                                                                 # and the batch goes on
    bad = dict(sc, bands=[dict(sc["bands"][0], devices=[dict(d, dest=0) if d["role"] == "sender" else d for d in sc["bands"][0]["devices"]])])
    with pytest.raises(N.NativeError):
        GeneralBandEnv(num_envs=1, device="cuda:0", scenario=bad)


# ------------------------------------------------------------------------------------------------------------
# mode M: per-bit Philox error masks keyed by (seed; env, sender, transmission, receiver, bit)
# ------------------------------------------------------------------------------------------------------------

GOLDEN_NSENDERS_M = "modeM_nsenders_4s_2p_seed34"


def _oracle_tape_m(sc, dev, dur, seed, env_id):
    ora = O.Oracle(sc, trace=True, mode=O.MODE_M)
    ora.use_philox_masks(seed, env_id)
    ora.reset()
    ora.take_records()
    steps = []
    for t in range(dev.shape[0]):
        obs, rew, done = ora.step({"device": int(dev[t, 0]), "duration": int(dur[t, 0])})
        steps.append({"obs": obs, "reward": rew, "done": done, "now": ora.now, "records": ora.take_records()})
    ntx, nd = ora.counts()
    return steps, ntx, nd, ora.received(), ora.near_ties


def test_oracle_and_core_match_reference_golden_nsenders_mode_m():
    """Reference + MaskedPhy (a SimplePhy subclass that replaces only the error bookkeeping by Philox masks) on a
    band of 4 senders + RRM + 2 PHY-only senders: integer error counts, verdicts, deliveries bit-exact."""
    doc = load_golden(GOLDEN_NSENDERS_M)
    sc, seed, env_id = doc["scenario"], doc["mask_seed"], doc["mask_env_id"]
    dev, dur = _tapes(doc)
    steps, _, _, _, _ = _oracle_tape_m(sc, dev, dur, seed, env_id)
    h = HS.gen_run(sc, dev, dur, mode=1, seed=seed, env_offset=env_id)
    assert h["rc"] == 0
    nfail = 0
    for t, g in enumerate(doc["steps"]):
        o = steps[t]
        assert (o["obs"], o["reward"], o["done"], o["now"]) == (g["obs"], g["reward"], g["done"], g["now"]), t
        assert_step_records(o["records"], g["records"], "oracle step %d" % t)
        assert (h["obs"][t, 0], h["reward"][t, 0], bool(h["done"][t, 0]), h["now"][t, 0]) == (g["obs"], g["reward"], g["done"], g["now"]), t
        assert_step_records(h["records"][t], g["records"], "core step %d" % t)
        nfail += sum(1 for r in g["records"] if r[0] == "dec" and not r[7])
    assert nfail > 10


@pytest.mark.parametrize("seed", range(6))
def test_core_general_mode_m_random_vs_oracle(seed):
    rs = np.random.RandomState(6000 + seed)
    ns, nj = int(rs.randint(3, 9)), int(rs.randint(0, 7))
    sc = random_scenario_n(rs, ns, nj, spread=2.5, receive=bool(seed % 2), bursts=bool(seed % 3 == 0))
    T = 40
    dev = rs.randint(0, ns, size=(T, 1)).astype(np.int32)
    dur = rs.randint(0, 20, size=(T, 1)).astype(np.int32)
    steps, ntx, nd, nrecv, _ = _oracle_tape_m(sc, dev, dur, 900 + seed, 31337 + seed)
    h = HS.gen_run(sc, dev, dur, mode=1, seed=900 + seed, env_offset=31337 + seed)
    _assert_host_equals(h, steps, ntx, nd, nrecv, ns, "mode M seed %d (%d senders, %d PHY-only)" % (seed, ns, nj))


@pytest.mark.gpu
def test_cuda_general_band_mode_m_matches_reference_golden():
    doc = load_golden(GOLDEN_NSENDERS_M)
    sc = doc["scenario"]
    env = _gpu_env(sc, 1, mode="mask_philox", seed=doc["mask_seed"], env_id_offset=doc["mask_env_id"])
    assert env.reset() == doc["reset_obs"]
    for t, g in enumerate(doc["steps"]):
        obs, rew, done, recs = env.step_traced({"device": g["action"]["device"], "duration": g["action"]["duration"]})
        assert (obs, rew, done) == (g["obs"], g["reward"], g["done"]), t
        assert float(env.now[0]) == g["now"], t
        assert_step_records(recs, g["records"], "step %d" % t)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(3))
def test_cuda_general_band_mode_m_matches_oracle_and_is_sharding_invariant(seed):
    """1024 envs x 32 steps in mode M (the warp counts the Philox ranges of its lanes' events together) against the
    oracle with the same keys; the second half of the envs stepped as a separate shard gives the same results."""
    import torch
    rs = np.random.RandomState(6400 + seed)
    ns, nj = int(rs.randint(3, 9)), int(rs.randint(0, 9))
    sc = random_scenario_n(rs, ns, nj, spread=2.0, receive=bool(seed % 2))
    nenv, T, mseed, off = 1024, 32, 4242 + seed, 1000003
    dev = rs.randint(0, ns, size=(T, nenv)).astype(np.int32)
    dur = rs.randint(0, 20, size=(T, nenv)).astype(np.int32)
    o = O.run_batch(sc, dev, dur, mode=O.MODE_M, seed=mseed, env_id_offset=off)
    env = _gpu_env(sc, nenv, mode="mask_philox", seed=mseed, env_id_offset=off)
    shard = _gpu_env(sc, nenv // 2, mode="mask_philox", seed=mseed, env_id_offset=off + nenv // 2)
    env.reset(); shard.reset()
    for t in range(T):
        a = {"device": torch.as_tensor(dev[t]).cuda(), "duration": torch.as_tensor(dur[t]).cuda()}
        obs, rew, done, _ = env.step(a)
        assert (obs.cpu().numpy() == o["obs"][t, :, 0]).all() and (rew.cpu().numpy() == o["reward"][t, :, 0]).all(), (seed, t)
        assert (env.now.cpu().numpy() == o["now"][t]).all(), (seed, t)
        obs2, rew2, _, _ = shard.step({"device": a["device"][nenv // 2:].contiguous(), "duration": a["duration"][nenv // 2:].contiguous()})
        assert torch.equal(obs2, obs[nenv // 2:]) and torch.equal(rew2, rew[nenv // 2:])
    env.check()
    assert (env.delivered().cpu().numpy() == o["counts"][:, 0, 1:1 + ns]).all()
    assert torch.equal(shard.delivered(), env.delivered()[nenv // 2:])


def _max_band(rs):
    """The largest band the engine holds: 8 MAC senders + RRM + 16 PHY-only senders (25 devices)."""
    return random_scenario_n(rs, 8, 16, spread=3.0, receive=True, bursts=True)


def test_core_general_maximum_configuration_vs_oracle():
    rs = np.random.RandomState(8816)
    sc = _max_band(rs)
    T = 40
    dev = rs.randint(0, 8, size=(T, 1)).astype(np.int32)
    dur = rs.randint(0, 20, size=(T, 1)).astype(np.int32)
    steps, ntx, nd, nrecv, _ = _oracle_tape(sc, dev, dur)
    h = HS.gen_run(sc, dev, dur)
    _assert_host_equals(h, steps, ntx, nd, nrecv, 8, "8 senders + RRM + 16 PHY-only senders")
    assert ntx > 50


@pytest.mark.gpu
def test_cuda_general_maximum_configuration_ragged_batch_vs_oracle():
    """25 devices per band, 77 envs (two full warps and a ragged one: the idle lanes of the last warp still take part in
    the warp-wide BER evaluations), modes R and M."""
    import torch
    rs = np.random.RandomState(8817)
    sc = _max_band(rs)
    nenv, T = 77, 24
    dev = rs.randint(0, 8, size=(T, nenv)).astype(np.int32)
    dur = rs.randint(0, 20, size=(T, nenv)).astype(np.int32)
    for mode, omode in (("reference", O.MODE_R), ("mask_philox", O.MODE_M)):
        o = O.run_batch(sc, dev, dur, mode=omode, seed=5, env_id_offset=900)
        env = _gpu_env(sc, nenv, mode=mode, seed=5, env_id_offset=900)
        env.reset()
        for t in range(T):
            obs, rew, done, _ = env.step({"device": torch.as_tensor(dev[t]).cuda(), "duration": torch.as_tensor(dur[t]).cuda()})
            assert (obs.cpu().numpy() == o["obs"][t, :, 0]).all() and (rew.cpu().numpy() == o["reward"][t, :, 0]).all(), (mode, t)
            assert (env.now.cpu().numpy() == o["now"][t]).all(), (mode, t)
        env.check()
        assert (env.transmissions().cpu().numpy() == o["counts"][:, 0, 0]).all()
        assert (env.delivered().cpu().numpy() == o["counts"][:, 0, 1:9]).all()


# ------------------------------------------------------------------------------------------------------------
# devices moving between steps (gw_genband_set_positions)
# ------------------------------------------------------------------------------------------------------------

def _random_moves(rs, sc, steps, start=1, colocate=True):
    """Before every other step one to three devices (ascending index) jump: within the band's area, far beyond
    STANDBY_THRESHOLD, or onto another device's position; busy PHY-only senders keep transmissions on the air."""
    devs = sc["bands"][0]["devices"]
    nd = len(devs)
    for d in devs:
        if d["role"] == "jammer":
            d["interval"] = float(rs.uniform(0.008, 0.02))
    cur = [(d["x"], d["y"]) for d in devs]
    moves = {}
    for t in range(start, steps, 2):
        lst = []
        for d in sorted(set(int(v) for v in rs.randint(nd, size=int(rs.randint(1, 4))))):
            k = int(rs.randint(5))
            if k == 0:
                x, y = float(rs.uniform(4000, 6000)), float(rs.uniform(-10, 10))
            elif k == 1 and colocate:
                x, y = cur[int((d + 1 + rs.randint(nd - 1)) % nd)]
            else:
                x, y = float(rs.uniform(-3, 3)), float(rs.uniform(-3, 3))
            cur[d] = (x, y)
            lst.append((0, d, float(x), float(y)))
        moves[t] = lst
    return moves


def _oracle_tape_moves(sc, dev, dur, moves, mode=O.MODE_R, seed=0, env_id=0):
    ora = O.Oracle(sc, trace=True, mode=mode)
    if mode == O.MODE_M:
        ora.use_philox_masks(seed, env_id)
    tape = [{"device": int(dev[t, 0]), "duration": int(dur[t, 0])} for t in range(dev.shape[0])]
    res = O.run_tape(ora, tape, moves=moves)
    ntx, nd = ora.counts()
    return res["steps"], ntx, nd, ora.received(), ora


@pytest.mark.parametrize("seed", range(8))
def test_core_general_moving_devices_vs_oracle(seed):
    """The engine equals the oracle record for record -- both move a device's links by ascending partner index."""
    rs = np.random.RandomState(7000 + seed)
    ns, nj = int(rs.randint(3, 7)), int(rs.randint(1, 5))
    sc = random_scenario_n(rs, ns, nj, spread=2.5, receive=bool(seed % 2))
    T = 50
    dev = rs.randint(0, ns, size=(T, 1)).astype(np.int32)
    dur = rs.randint(0, 20, size=(T, 1)).astype(np.int32)
    moves = _random_moves(rs, sc, T, start=seed % 2)
    m = O.MODE_M if seed >= 6 else O.MODE_R
    try:
        steps, ntx, nd, nrecv, _ = _oracle_tape_moves(sc, dev, dur, moves, mode=m, seed=55, env_id=7)
    except O.OracleFault:
        pytest.skip("the reference raises in this scenario")
    h = HS.gen_run(sc, dev, dur, moves=moves, mode=1 if m == O.MODE_M else 0, seed=55, env_offset=7)
    _assert_host_equals(h, steps, ntx, nd, nrecv, ns, "moves seed %d" % seed)


def test_oracle_and_core_match_reference_golden_nsenders_mobility():
    from util import GOLDEN_NSENDERS_MOBILITY, assert_mobile_step_records
    doc = load_golden(GOLDEN_NSENDERS_MOBILITY)
    sc = doc["scenario"]
    moves = {int(k): [tuple(m) for m in v] for k, v in doc["moves"].items()}
    dev, dur = _tapes(doc)
    steps, _, _, _, _ = _oracle_tape_moves(sc, dev, dur, moves)
    h = HS.gen_run(sc, dev, dur, moves=moves)
    assert h["rc"] == 0
    for t, g in enumerate(doc["steps"]):
        o = steps[t]
        assert (o["obs"], o["reward"], o["done"], o["now"]) == (g["obs"], g["reward"], g["done"], g["now"]), t
        assert (h["obs"][t, 0], h["reward"][t, 0], bool(h["done"][t, 0]), h["now"][t, 0]) == (g["obs"], g["reward"], g["done"], g["now"]), t
        assert_mobile_step_records(o["records"], g["records"], "oracle step %d" % t)
        assert_mobile_step_records(h["records"][t], g["records"], "core step %d" % t)
        assert_step_records(h["records"][t], o["records"], "core vs oracle step %d" % t)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(3))
def test_cuda_general_band_moving_devices_vs_oracle(seed):
    """GeneralBandEnv.set_positions between steps, 16 envs with the same tape of jumps: step results, step end times and
    delivery counts equal the oracle's; modes R and M."""
    import torch
    rs = np.random.RandomState(7100 + seed)
    ns, nj = int(rs.randint(3, 7)), int(rs.randint(1, 5))
    sc = random_scenario_n(rs, ns, nj, spread=2.5, receive=bool(seed % 2))
    nd, nenv, T = ns + 1 + nj, 16, 40
    dev = np.repeat(rs.randint(0, ns, size=(T, 1)), nenv, axis=1).astype(np.int32)
    dur = np.repeat(rs.randint(0, 20, size=(T, 1)), nenv, axis=1).astype(np.int32)
    moves = _random_moves(rs, sc, T, start=seed % 2, colocate=False)
    m = O.MODE_M if seed == 2 else O.MODE_R
    steps, ntx, ndl, nrecv, _ = _oracle_tape_moves(sc, dev[:, :1], dur[:, :1], moves, mode=m, seed=9, env_id=0)
    pos0 = np.array([[d["x"], d["y"]] for d in sc["bands"][0]["devices"]])
    # every env has the same geometry and the same key (a batch of 16 one-env handles would do the same)
    envs = [_gpu_env(sc, 1, positions=torch.as_tensor(pos0[None]), mode="mask_philox" if m == O.MODE_M else "reference", seed=9,
                     env_id_offset=0)]
    batch = _gpu_env(sc, nenv, positions=torch.as_tensor(np.repeat(pos0[None], nenv, axis=0)), mode="reference")
    envs[0].reset(); batch.reset()
    cur = pos0.copy()
    for t in range(T):
        if t in moves:
            for (_, d, x, y) in moves[t]:
                cur[d] = (x, y)
            envs[0].set_positions(torch.as_tensor(cur[None]))
            batch.set_positions(torch.as_tensor(np.repeat(cur[None], nenv, axis=0)))
        obs, rew, done, _ = envs[0].step({"device": int(dev[t, 0]), "duration": int(dur[t, 0])})
        s = steps[t]
        assert (obs, rew, float(envs[0].now[0])) == (s["obs"], s["reward"], s["now"]), (seed, t)
        ob, rb, _, _ = batch.step({"device": torch.as_tensor(dev[t]).cuda(), "duration": torch.as_tensor(dur[t]).cuda()})
        if m == O.MODE_R:
            assert (ob.cpu().numpy() == s["obs"]).all() and (batch.now.cpu().numpy() == s["now"]).all(), (seed, t)
    assert envs[0].delivered()[0].tolist() == list(ndl[:ns]) and int(envs[0].transmissions()[0]) == ntx


@pytest.mark.gpu
def test_cuda_general_band_mobility_matches_reference_golden():
    """The reference's own trace of a band of 5 senders + RRM + 3 PHY-only senders whose devices jump between steps
    while transmissions are on the air: step results, transmissions, deliveries, verdicts exact; error sums and rates
    within the tolerances that the reference's own set-order nondeterminism allows (util.assert_mobile_step_records)."""
    import torch
    from util import GOLDEN_NSENDERS_MOBILITY, assert_mobile_step_records
    doc = load_golden(GOLDEN_NSENDERS_MOBILITY)
    sc = doc["scenario"]
    moves = {int(k): [tuple(m) for m in v] for k, v in doc["moves"].items()}
    cur = np.array([[d["x"], d["y"]] for d in sc["bands"][0]["devices"]])
    env = _gpu_env(sc, 1, positions=torch.as_tensor(cur[None]))
    assert env.reset() == doc["reset_obs"]
    for t, g in enumerate(doc["steps"]):
        if t in moves:
            for (_, d, x, y) in moves[t]:
                cur[d] = (x, y)
            env.set_positions(torch.as_tensor(cur[None]))
        obs, rew, done, recs = env.step_traced({"device": g["action"]["device"], "duration": g["action"]["duration"]})
        assert (obs, rew, done, float(env.now[0])) == (g["obs"], g["reward"], g["done"], g["now"]), t
        # (the records of the jumps themselves belong to the set_positions call, which is not traced: compare the step's)
        want = [r for r in g["records"] if not (r[0] == "ber" and r[1] == (doc["steps"][t - 1]["now"] if t else 0.0))]
        got = [r for r in recs if not (r[0] == "ber" and r[1] == (doc["steps"][t - 1]["now"] if t else 0.0))]
        assert_mobile_step_records(got, want, "step %d" % t)


# ------------------------------------------------------------------------------------------------------------
# mobility processes DURING the steps (gw_genband_set_movers; the mover of tests/test_benchmark.py:73-85)
# ------------------------------------------------------------------------------------------------------------

def _movers_from_golden(doc, nd):
    K = max(len(m["offsets"]) for m in doc["movers"].values())
    md = -np.ones(nd)
    off = np.zeros((nd, K, 2))
    for i, m in doc["movers"].items():
        md[int(i)] = m["first_delay"]
        off[int(i), :len(m["offsets"])] = np.array(m["offsets"])
    return md, off, float(next(iter(doc["movers"].values()))["interval"])


def _oracle_with_movers(sc, md, off, interval, mode=O.MODE_R, seed=0, env_id=0):
    ora = O.Oracle(sc, trace=True, mode=mode)
    if mode == O.MODE_M:
        ora.use_philox_masks(seed, env_id)
    for d in range(len(md)):
        if md[d] >= 0:
            ora.add_mover(0, d, float(md[d]), interval, off[d])
    return ora


@pytest.mark.parametrize("seed", range(6))
def test_core_general_mobility_processes_vs_oracle(seed):
    rs = np.random.RandomState(7500 + seed)
    ns, nj = int(rs.randint(3, 7)), int(rs.randint(0, 5))
    sc = random_scenario_n(rs, ns, nj, spread=2.5, receive=bool(seed % 2))
    nd, T = ns + 1 + nj, 40
    dev = rs.randint(0, ns, size=(T, 1)).astype(np.int32)
    dur = rs.randint(0, 20, size=(T, 1)).astype(np.int32)
    md = rs.uniform(0, 1e-3, size=nd)
    md[rs.rand(nd) < 0.3] = -1.0                        # some devices stay put
    off = rs.uniform(-.2, .2, size=(nd, 600, 2))
    m = O.MODE_M if seed >= 4 else O.MODE_R
    ora = _oracle_with_movers(sc, md, off, 1e-3, mode=m, seed=55, env_id=7)
    res = O.run_tape(ora, [{"device": int(dev[t, 0]), "duration": int(dur[t, 0])} for t in range(T)])
    ntx, ndl = ora.counts()
    h = HS.gen_run(sc, dev, dur, move_delays=md, offsets=off, mode=1 if m == O.MODE_M else 0, seed=55, env_offset=7)
    _assert_host_equals(h, res["steps"], ntx, ndl, ora.received(), ns, "movers seed %d" % seed)


def test_oracle_and_core_match_reference_golden_nsenders_movers():
    """The reference's own trace of a band of 4 senders + RRM + 2 PHY-only senders whose devices run mobility processes
    (a jump every millisecond, while transmissions are on the air)."""
    from util import GOLDEN_NSENDERS_MOVERS, assert_mobile_step_records
    doc = load_golden(GOLDEN_NSENDERS_MOVERS)
    sc = doc["scenario"]
    nd = len(sc["bands"][0]["devices"])
    md, off, interval = _movers_from_golden(doc, nd)
    dev, dur = _tapes(doc)
    ora = _oracle_with_movers(sc, md, off, interval)
    res = O.run_tape(ora, [s["action"] for s in doc["steps"]])
    h = HS.gen_run(sc, dev, dur, move_delays=md, offsets=off, move_interval=interval)
    assert h["rc"] == 0
    for t, g in enumerate(doc["steps"]):
        o = res["steps"][t]
        assert (o["obs"], o["reward"], o["done"], o["now"]) == (g["obs"], g["reward"], g["done"], g["now"]), t
        assert (h["obs"][t, 0], h["reward"][t, 0], bool(h["done"][t, 0]), h["now"][t, 0]) == (g["obs"], g["reward"], g["done"], g["now"]), t
        assert_mobile_step_records(o["records"], g["records"], "oracle step %d" % t)
        assert_mobile_step_records(h["records"][t], g["records"], "core step %d" % t)
        assert_step_records(h["records"][t], o["records"], "core vs oracle step %d" % t)


@pytest.mark.gpu
def test_cuda_general_band_mobility_processes():
    """GeneralBandEnv.set_movers: the reference golden through the traced kernel, and a batch of 96 envs (every env
    its own tape) against the oracle -- step results, step end times, delivery counts."""
    import torch
    from util import GOLDEN_NSENDERS_MOVERS, assert_mobile_step_records
    doc = load_golden(GOLDEN_NSENDERS_MOVERS)
    sc = doc["scenario"]
    nd = len(sc["bands"][0]["devices"])
    ns = sum(1 for d in sc["bands"][0]["devices"] if d["role"] == "sender")
    md, off, interval = _movers_from_golden(doc, nd)
    pos0 = np.array([[d["x"], d["y"]] for d in sc["bands"][0]["devices"]])
    env = _gpu_env(sc, 1, positions=torch.as_tensor(pos0[None]))
    env.set_movers(torch.as_tensor(md[None]), torch.as_tensor(off[None]), interval)
    assert env.reset() == doc["reset_obs"]
    for t, g in enumerate(doc["steps"]):
        obs, rew, done, recs = env.step_traced({"device": g["action"]["device"], "duration": g["action"]["duration"]})
        assert (obs, rew, done, float(env.now[0])) == (g["obs"], g["reward"], g["done"], g["now"]), t
        assert_mobile_step_records(recs, g["records"], "step %d" % t)
    # a batch, every env its own delays / tapes / actions
    rs = np.random.RandomState(7600)
    nenv, T = 96, 24
    mds = rs.uniform(0, 1e-3, size=(nenv, nd))
    mds[rs.rand(nenv, nd) < 0.3] = -1.0
    offs = rs.uniform(-.2, .2, size=(nenv, nd, 400, 2))
    dev = rs.randint(0, ns, size=(T, nenv)).astype(np.int32)
    dur = rs.randint(0, 20, size=(T, nenv)).astype(np.int32)
    batch = _gpu_env(sc, nenv, positions=torch.as_tensor(np.repeat(pos0[None], nenv, axis=0)))
    batch.set_movers(torch.as_tensor(mds), torch.as_tensor(offs), 1e-3)
    batch.reset()
    got_obs, got_now = [], []
    for t in range(T):
        obs, rew, done, _ = batch.step({"device": torch.as_tensor(dev[t]).cuda(), "duration": torch.as_tensor(dur[t]).cuda()})
        got_obs.append(obs.cpu().numpy()); got_now.append(batch.now.cpu().numpy())
    batch.check()
    deliv = batch.delivered().cpu().numpy()
    for e in range(0, nenv, 7):
        ora = _oracle_with_movers(sc, mds[e], offs[e], 1e-3)
        res = O.run_tape(ora, [{"device": int(dev[t, e]), "duration": int(dur[t, e])} for t in range(T)])
        for t, s in enumerate(res["steps"]):
            assert got_obs[t][e] == s["obs"] and got_now[t][e] == s["now"], (e, t)
        assert list(deliv[e]) == list(ora.counts()[1][:ns])


def test_core_general_long_run_through_counter_saturation():
    """7,000 steps (~75 simulated seconds) of a 3-sender band: the counters saturate at COUNTER_BOUND = 65,536, the
    queues sit at their capacity of 100 with drop-oldest, packets outgrow every window (the reference's degenerate
    steady state) -- step results, step end times and counts equal the oracle's throughout, with
    a reset() on the way."""
    rs = np.random.RandomState(99)
    sc = random_scenario_n(rs, 3, 1, spread=2.0)
    for d in sc["bands"][0]["devices"][:3]:
        d["payload"] = "counter"
        d["interval"] = 0.001
    T = 7000
    dev = rs.randint(0, 3, size=(T, 1)).astype(np.int32)
    dur = rs.randint(0, 20, size=(T, 1)).astype(np.int32)
    ora = O.Oracle(sc)
    ora.reset()
    obs, now = np.zeros(T, np.int64), np.zeros(T)
    for t in range(T):
        if t == 4000:
            ora.reset()
        o, r, d = ora.step({"device": int(dev[t, 0]), "duration": int(dur[t, 0])})
        obs[t], now[t] = o, ora.now
    h = HS.gen_run(sc, dev, dur, reset_at=None, do_reset=True, trace=False)
    h2 = HS.gen_run(sc, dev[:4000], dur[:4000], trace=False)
    assert h2["rc"] == 0 and (h2["obs"][:, 0] == obs[:4000]).all() and (h2["now"][:, 0] == now[:4000]).all()
    # (the host driver resets once, before `reset_at`: the second half is checked with the reset at step 4000)
    h3 = HS.gen_run(sc, dev, dur, do_reset=False, reset_at=4000, trace=False)
    ora2 = O.Oracle(sc)
    obs2, now2 = np.zeros(T, np.int64), np.zeros(T)
    for t in range(T):
        if t == 4000:
            ora2.reset()
        o, r, d = ora2.step({"device": int(dev[t, 0]), "duration": int(dur[t, 0])})
        obs2[t], now2[t] = o, ora2.now
    assert h3["rc"] == 0 and (h3["obs"][:, 0] == obs2).all() and (h3["now"][:, 0] == now2).all()
    ntx, nd = ora2.counts()
    assert h3["counts"][0, 0] == ntx and list(h3["counts"][0, 1:4]) == list(nd[:3])
    assert now2[-1] > 65.0 and h["rc"] == 0


def test_core_general_mode_m_is_sharding_invariant():
    """The error masks are keyed by the GLOBAL env id: two shards with their own env offsets give the results of the
    whole batch (what lets the env batch be split over GPUs with no data-path collective)."""
    rs = np.random.RandomState(31)
    sc = random_scenario_n(rs, 4, 2, spread=2.0, receive=True)
    nenv, T = 10, 24
    dev = rs.randint(0, 4, size=(T, nenv)).astype(np.int32)
    dur = rs.randint(0, 20, size=(T, nenv)).astype(np.int32)
    whole = HS.gen_run(sc, dev, dur, mode=1, seed=17, env_offset=1000, trace=False)
    lo = HS.gen_run(sc, dev[:, :6], dur[:, :6], mode=1, seed=17, env_offset=1000, trace=False)
    hi = HS.gen_run(sc, dev[:, 6:], dur[:, 6:], mode=1, seed=17, env_offset=1006, trace=False)
    for key in ("obs", "reward", "now"):
        assert (np.concatenate([lo[key], hi[key]], axis=1) == whole[key]).all(), key
    assert (np.concatenate([lo["counts"], hi["counts"]], axis=0) == whole["counts"]).all()
    other = HS.gen_run(sc, dev[:, 6:], dur[:, 6:], mode=1, seed=17, env_offset=0, trace=False)
    assert not (other["counts"] == hi["counts"]).all() or not (other["now"] == hi["now"]).all()     # another key, other masks
