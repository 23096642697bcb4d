"""
GPU suite, mode M (per-bit error masks): the CUDA path against the golden vectors of the
reference + MaskedPhy subclass, against the oracle with the same Philox keys, and with FED masks
(both sides read the same mask words).  Plus the standalone K3 popcount kernel and Philox KATs.
"""
import numpy as np
import pytest
import torch

import gw_oracle as O
from util import GOLDEN_CASES_M, golden_results, load_golden, random_scenario, random_tapes, tapes_from_golden

pytestmark = pytest.mark.gpu


def make_env(scenario, n, **kw):
    import gymwipe_b200
    return gymwipe_b200.make('CounterTraffic-v0', num_envs=n, scenario=scenario, strict=False, **kw)


def run_gpu(env, dev, dur):
    nsteps, nenv, nb = dev.shape
    d_dev, d_dur = torch.as_tensor(dev).cuda(), torch.as_tensor(dur).cuda()
    obs = np.zeros((nsteps, nenv, nb), np.int64)
    rew = np.zeros((nsteps, nenv, nb), np.float64)
    now = np.zeros((nsteps, nenv), np.float64)
    for t in range(nsteps):
        o, r, d, _ = env.step({"device": d_dev[t].reshape(env._shape), "duration": d_dur[t].reshape(env._shape)})
        obs[t], rew[t] = o.reshape(nenv, nb).cpu().numpy(), r.reshape(nenv, nb).cpu().numpy()
        now[t] = env.read_state(0).cpu().numpy()
    env.check()
    counts = np.zeros((nenv, nb, 3), np.int64)
    counts[:, :, 0] = env.transmissions().cpu().numpy().reshape(nenv, nb)
    counts[:, :, 1:3] = env.delivered().cpu().numpy().reshape(nenv, nb, 2)
    return {"obs": obs, "reward": rew, "now": now, "counts": counts}


def assert_same(o, g):
    assert (o["obs"] == g["obs"]).all()
    assert (o["reward"] == g["reward"]).all()
    assert (o["now"] == g["now"]).all()
    assert (o["counts"][:, :, :3] == g["counts"]).all()


@pytest.mark.parametrize("name", GOLDEN_CASES_M)
def test_mode_m_matches_reference_golden(name):
    doc = load_golden(name)
    dev, dur = tapes_from_golden(doc)
    env = make_env(doc["scenario"], 1, mode="mask_philox", seed=doc["mask_seed"], env_id_offset=doc["mask_env_id"])
    if doc["do_reset"]:
        env.reset()
    g = run_gpu(env, dev, dur)
    obs, rew, done, now = golden_results(doc)
    assert (g["obs"][:, 0, :] == obs).all() and (g["reward"][:, 0, :] == rew).all()
    assert (g["now"][:, 0] == now).all()
    n_rx = sum(1 for s in doc["steps"] for r in s["records"] if r[0] == "rx" and r[3] < 2)
    assert g["counts"][0, :, 1:3].sum() == n_rx


@pytest.mark.parametrize("seed", range(3))
def test_mode_m_philox_matches_oracle(seed):
    rs = np.random.RandomState(400 + seed)
    from gymwipe_b200.scenario import default_scenario_dict
    cases = [(default_scenario_dict(), 512, 100),
             (random_scenario(rs, jammers=1, spread=2.5), 256, 80),
             (random_scenario(rs, jammers=1, spread=2.0, fixed_payload=1500, factor=10000), 64, 24),
             (random_scenario(rs, nbands=4, jammers=1, spread=2.5), 32, 40)]
    for sc, nenv, nsteps in cases:
        dev, dur = random_tapes(rs, nsteps, nenv, len(sc["bands"]))
        o = O.run_batch(sc, dev, dur, mode=O.MODE_M, seed=5 + seed, env_id_offset=70000)
        env = make_env(sc, nenv, mode="mask_philox", seed=5 + seed, env_id_offset=70000)
        env.reset()
        assert_same(o, run_gpu(env, dev, dur))


def test_mode_m_sharding_invariance():
    """RNG keys use GLOBAL env ids: two shards of 128 envs equal one batch of 256."""
    from gymwipe_b200.scenario import default_scenario_dict
    sc = default_scenario_dict()
    rs = np.random.RandomState(3)
    dev, dur = random_tapes(rs, 60, 256, 1)
    whole = make_env(sc, 256, mode="mask_philox", seed=9)
    whole.reset()
    g = run_gpu(whole, dev, dur)
    for k in range(2):
        part = make_env(sc, 128, mode="mask_philox", seed=9, env_id_offset=128 * k)
        part.reset()
        gp = run_gpu(part, dev[:, 128 * k:128 * (k + 1)], dur[:, 128 * k:128 * (k + 1)])
        assert (gp["obs"] == g["obs"][:, 128 * k:128 * (k + 1)]).all()
        assert (gp["counts"] == g["counts"][128 * k:128 * (k + 1)]).all()


def _bernoulli_masks(rs, nenv, slots, words):
    # Bernoulli(p) bits, p different per receiver so that some packets fail and some pass
    p = rs.uniform(0.02, 0.3, size=(nenv, 1, 4, slots, 4, 1))
    bits = rs.random_sample((nenv, 1, 4, slots, 4, words * 32)) < p
    masks = np.packbits(bits.reshape(-1, 8)[:, ::-1], axis=1).reshape(nenv, 1, 4, slots, 4, words * 4)
    return np.ascontiguousarray(masks.view("<u4").reshape(nenv, 1, 4, slots, 4, words))


@pytest.mark.parametrize("index", ["1", "0"])
def test_mode_m_fed_masks_match_oracle(index, monkeypatch):
    """Both sides are fed the SAME mask words; long packets (1500-byte payloads) with a jammer.  index = 1: the
    counts come from the prefix-count index gw_set_masks builds (step_kernel<MODE_M_FEDX>); 0: the step kernel
    scans the mask words itself (step_kernel<MODE_M_FED>, GW_FED_INDEX=0)."""
    monkeypatch.setenv("GW_FED_INDEX", index)
    rs = np.random.RandomState(55)
    sc = random_scenario(rs, jammers=1, spread=2.0, fixed_payload=1500, factor=10000)
    nenv, nsteps, slots, words = 48, 24, 4, 512
    masks = _bernoulli_masks(rs, nenv, slots, words)
    dev, dur = random_tapes(rs, nsteps, nenv, 1)
    o = O.run_batch(sc, dev, dur, mode=O.MODE_M, fed_words=masks, fed_slots=slots)
    env = make_env(sc, nenv, mode="mask_fed")
    env.set_masks(torch.as_tensor(masks.view(np.int32)).cuda(), slots)
    env.reset()
    g = run_gpu(env, dev, dur)
    assert_same(o, g)
    assert 0 < o["counts"][:, :, 1:3].sum() < o["counts"][:, :, 0].sum()


def test_mode_m_fed_rows_longer_than_a_superblock():
    """8000-byte payloads: 85,461 on-air bits per packet = 668 groups of 128 bits, more than one 512-group
    superblock of the prefix-count index (its uint32 level), windows of up to 0.95 s."""
    rs = np.random.RandomState(77)
    sc = random_scenario(rs, jammers=1, spread=2.0, fixed_payload=8000, factor=50000)
    nenv, nsteps, slots, words = 6, 10, 2, 2688
    masks = _bernoulli_masks(rs, nenv, slots, words)
    dev, dur = random_tapes(rs, nsteps, nenv, 1)
    dur = np.maximum(dur, 14)
    o = O.run_batch(sc, dev, dur, mode=O.MODE_M, fed_words=masks, fed_slots=slots)
    env = make_env(sc, nenv, mode="mask_fed")
    env.set_masks(torch.as_tensor(masks.view(np.int32)).cuda(), slots)
    env.reset()
    assert_same(o, run_gpu(env, dev, dur))
    assert o["counts"][:, :, 0].sum() > 0


@pytest.mark.parametrize("words", [64, 512, 2052, 4096])
def test_mask_index_counts_match_numpy(words):
    """gw_mask_index_count (the look-up the step kernel does) against numpy popcounts over arbitrary bit ranges:
    rows shorter than, equal to and longer than one 512-group superblock (2048 words)."""
    from gymwipe_b200 import _native as N
    from gymwipe_b200.scenario import default_scenario_dict
    rs = np.random.RandomState(words)
    nenv, slots = 3, 2
    rows = nenv * 4 * slots * 4
    m = rs.randint(0, 2 ** 32, size=(rows, words), dtype=np.uint64).astype(np.uint32)
    m &= rs.randint(0, 2 ** 32, size=(rows, words), dtype=np.uint64).astype(np.uint32)
    m[1] = 0xFFFFFFFF                    # a full row: the largest counts
    env = make_env(default_scenario_dict(), nenv, mode="mask_fed")
    env.set_masks(torch.as_tensor(m.view(np.int32).reshape(nenv, 1, 4, slots, 4, words)).cuda(), slots)
    n = 6000
    ridx = rs.randint(0, rows, n).astype(np.int64)
    k0 = rs.randint(0, words * 32, n).astype(np.int32)
    k1 = np.minimum(words * 32, k0 + rs.randint(0, 2 * words * 32, n)).astype(np.int32)
    k1[:50] = k0[:50]                     # empty ranges
    k0[50:80], k1[50:80] = 0, words * 32  # whole rows
    k0[80:110] = (k0[80:110] >> 7) << 7   # group-aligned starts
    k1[110:140] = np.maximum(k0[110:140], (k1[110:140] >> 7) << 7)
    ridx[140:150] = 1
    bits = np.unpackbits(m.view(np.uint8).reshape(rows, -1), axis=1, bitorder="little")
    cs = np.concatenate([np.zeros((rows, 1), np.int64), np.cumsum(bits, axis=1)], axis=1)
    want = cs[ridx, k1] - cs[ridx, k0]
    out = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_rows, d_k0, d_k1 = torch.as_tensor(ridx).cuda(), torch.as_tensor(k0).cuda(), torch.as_tensor(k1).cuda()
    N.check(N.lib().gw_mask_index_count(env._handle, d_rows.data_ptr(), d_k0.data_ptr(), d_k1.data_ptr(),
                                        out.data_ptr(), n, torch.cuda.current_stream().cuda_stream))
    assert (out.cpu().numpy().astype(np.int64) == want.astype(np.int64)).all()


def test_k3_count_bit_errors_kernel():
    """gw_count_bit_errors against numpy popcounts over arbitrary bit ranges."""
    from gymwipe_b200 import _native as N
    rs = np.random.RandomState(1)
    rows, words = 300, 512
    m = rs.randint(0, 2 ** 32, size=(rows, words), dtype=np.uint64).astype(np.uint32)
    n = 4000
    ridx = rs.randint(0, rows, n).astype(np.int64)
    k0 = rs.randint(0, words * 32, n).astype(np.int32)
    k1 = np.minimum(words * 32, k0 + rs.randint(0, 17000, n)).astype(np.int32)
    k1[:50] = k0[:50]                     # empty ranges
    k0[50:60], k1[50:60] = 0, words * 32  # whole rows
    bits = np.unpackbits(m.view(np.uint8).reshape(rows, -1), axis=1, bitorder="little")
    cs = np.concatenate([np.zeros((rows, 1), np.int64), np.cumsum(bits, axis=1)], axis=1)
    want = cs[ridx, k1] - cs[ridx, k0]
    dm = torch.as_tensor(m.view(np.int32)).cuda()
    out = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_rows, d_k0, d_k1 = torch.as_tensor(ridx).cuda(), torch.as_tensor(k0).cuda(), torch.as_tensor(k1).cuda()
    N.check(N.lib().gw_count_bit_errors(dm.data_ptr(), words, d_rows.data_ptr(), d_k0.data_ptr(), d_k1.data_ptr(),
                                        out.data_ptr(), n, torch.cuda.current_stream().cuda_stream))
    assert (out.cpu().numpy().astype(np.int64) == want.astype(np.int64)).all()


def test_philox_known_answers_device():
    from gymwipe_b200 import _native as N
    ctr = np.array([[0, 0, 0, 0], [0xffffffff] * 4, [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344]], np.uint32)
    key = np.array([[0, 0], [0xffffffff] * 2, [0xa4093822, 0x299f31d0]], np.uint32)
    want = np.array([[0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8], [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd],
                     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]], np.uint32)
    c, k = torch.as_tensor(ctr.view(np.int32)).cuda(), torch.as_tensor(key.view(np.int32)).cuda()
    out = torch.zeros((3, 4), dtype=torch.int32, device="cuda")
    N.check(N.lib().gw_philox4x32(c.data_ptr(), k.data_ptr(), out.data_ptr(), 3, torch.cuda.current_stream().cuda_stream))
    assert (out.cpu().numpy().view(np.uint32) == want).all()


def test_mode_m_fed_moving_devices_match_oracle():
    """Fed masks with devices jumping between steps: a position change cuts the running segment
    (move_kernel counts it and restarts the segment), the decision counts the rest -- the receivers of one
    transmission then decide over DIFFERENT ranges (the kernel's 'own scans' path)."""
    rs = np.random.RandomState(9300)
    sc = random_scenario(rs, jammers=1, spread=2.5)
    sc["bands"][0]["devices"][3]["interval"] = float(rs.uniform(0.008, 0.02))
    nenv, nsteps, slots, words = 1, 70, 3, 64
    dev, dur = random_tapes(rs, nsteps, nenv, 1)
    dur = np.minimum(dur, 9)
    p = rs.uniform(0.02, 0.3, size=(nenv, 1, 4, slots, 4, 1))
    bits = rs.random_sample((nenv, 1, 4, slots, 4, words * 32)) < p
    masks = np.packbits(bits.reshape(-1, 8)[:, ::-1], axis=1).reshape(nenv, 1, 4, slots, 4, words * 4)
    masks = np.ascontiguousarray(masks.view("<u4").reshape(nenv, 1, 4, slots, 4, words))
    moves = {}
    for t in range(1, nsteps, 2):
        devs = sorted(set(int(v) for v in rs.randint(4, size=int(rs.randint(1, 4)))))
        moves[t] = [(0, d, float(rs.uniform(-3, 3)), float(rs.uniform(-3, 3))) for d in devs]
    acts = [{"device": int(dev[t, 0, 0]), "duration": int(dur[t, 0, 0])} for t in range(nsteps)]
    ora = O.Oracle(sc, mode=O.MODE_M)
    ora.use_fed_masks(masks, slots)
    res = O.run_tape(ora, acts, do_reset=True, moves=moves)
    pos = np.array([[[[d["x"], d["y"]] for d in sc["bands"][0]["devices"]]]], np.float64)      # [1, 1, 4, 2]
    env = make_env(sc, nenv, mode="mask_fed", positions=torch.as_tensor(pos).cuda())
    env.set_masks(torch.as_tensor(masks.view(np.int32)).cuda(), slots)
    env.reset()
    for t in range(nsteps):
        if t in moves:
            for (b, d, x, y) in moves[t]:
                pos[0, b, d] = (x, y)
            env.set_positions(torch.as_tensor(pos).cuda())
        o, r, dn, _ = env.step({"device": torch.as_tensor(dev[t, :, 0]).cuda(), "duration": torch.as_tensor(dur[t, :, 0]).cuda()})
        assert int(o[0]) == res["steps"][t]["obs"] and float(r[0]) == res["steps"][t]["reward"]
        assert float(env.read_state(0)[0]) == res["steps"][t]["now"]
    env.check()
