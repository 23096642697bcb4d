"""
CPU suite, part 1: the oracle (plain-C restatement) against the golden vectors that were
produced by the UNMODIFIED reference (oracle/gen_golden.py).  Bit-exact on everything:
obs / reward / done / step end time / every transmission / every BER value / every decider
input and verdict / every RRM delivery.
"""
import numpy as np
import pytest

import gw_oracle as O
from util import GOLDEN_CASES, GOLDEN_CASES_M, canonical, load_golden


@pytest.mark.parametrize("name", GOLDEN_CASES + GOLDEN_CASES_M)
def test_restatement_matches_reference_trace(name):
    doc = load_golden(name)
    if doc.get("mode", "R") == "M":
        ora = O.Oracle(doc["scenario"], trace=True, mode=O.MODE_M)
        ora.use_philox_masks(doc["mask_seed"], doc["mask_env_id"])
    else:
        ora = O.Oracle(doc["scenario"], trace=True)
    if doc["do_reset"]:
        assert ora.reset() == doc["reset_obs"]
    ora.take_records()
    for i, s in enumerate(doc["steps"]):
        fb = ora.step(s["action"])
        if isinstance(fb, list):
            obs, rew, done = [f[0] for f in fb], [f[1] for f in fb], [f[2] for f in fb]
        else:
            obs, rew, done = fb
        assert obs == s["obs"], (name, i)
        assert rew == s["reward"], (name, i)
        assert done == s["done"], (name, i)
        assert ora.now == s["now"], (name, i)          # bit-exact fp64 time
        assert canonical(ora.take_records()) == canonical(s["records"]), (name, i)
    if doc.get("mode", "R") == "R":
        assert ora.near_ties == 0          # no decider threshold is within rounding distance


@pytest.mark.parametrize("name", ["mobility_seed13", "mobility_inflight_seed17", "mobility_quirks_seed4"])
def test_restatement_matches_reference_with_moving_devices(name):
    """Devices move between steps (Position.set in the reference); goldens from the reference.  In the
    second case a PHY-only sender is regularly on the air when devices move: the reference's
    SimplePhy._onAttenuationChange path (mid-packet power change, error count, BER re-evaluation)."""
    doc = load_golden(name)
    moves = {int(k): [tuple(m) for m in v] for k, v in doc["moves"].items()}
    ora = O.Oracle(doc["scenario"], trace=True)
    res = O.run_tape(ora, [s["action"] for s in doc["steps"]], do_reset=doc["do_reset"], moves=moves)
    for i, (a, b) in enumerate(zip(doc["steps"], res["steps"])):
        assert a["obs"] == b["obs"] and a["reward"] == b["reward"] and a["now"] == b["now"], i
        assert canonical(a["records"]) == canonical(b["records"]), i
    if name == "mobility_inflight_seed17":
        # the case does exercise mid-packet changes: BER records stamped with a step's START time
        starts = [0.0] + [s["now"] for s in doc["steps"][:-1]]
        hits = sum(1 for s, t0 in zip(doc["steps"], starts) for r in s["records"] if r[0] == "ber" and r[1] == t0)
        assert hits >= 10, hits


def test_reference_known_answer():
    """tests/envs/test_counter_traffic.py:25-34 of the reference."""
    ora = O.Oracle()
    obs, reward, _ = ora.step({"device": 0, "duration": 3})
    assert obs - 65536 == 2 and reward == -2
    obs, reward, _ = ora.step({"device": 1, "duration": 12})
    assert obs - 65536 == 0 and reward == 2


def test_survey_appendix_c_vectors():
    """cumulative transmissions / deliveries / step end times."""
    ora = O.Oracle()
    ora.reset()
    acts = [(0, 15), (1, 0), (1, 3), (1, 9), (1, 18), (0, 6), (0, 12), (0, 1), (0, 7), (1, 14), (0, 17), (1, 13)]
    cum_tx = [7, 8, 10, 15, 22, 25, 30, 31, 34, 37, 43, 45]
    cum_deliv = [[6, 0], [6, 0], [6, 1], [6, 5], [6, 11], [8, 11], [12, 11], [12, 11], [14, 11], [14, 13], [19, 13], [19, 14]]
    nows = [0.016442000036, 0.017564000028, 0.021926000034000002, 0.032288000034, 0.05173000003599999,
            0.059092000033999996, 0.072534000036, 0.074896000034, 0.08325800003399998, 0.098700000036,
            0.117142000036, 0.13158400003600002]
    for (d, u), tx, dl, t in zip(acts, cum_tx, cum_deliv, nows):
        ora.step({"device": d, "duration": u})
        n_tx, n_deliv = ora.counts()
        assert n_tx == tx and n_deliv[:2] == dl and ora.now == t


def test_arithmetic_vectors():
    doc = load_golden("arithmetic")
    L = O.lib()
    for x, q in doc["q"]:
        assert L.gwo_q_function(x) == q
    for sd, nd, ber in doc["ber_dbm"]:
        assert L.gwo_ber_bpsk(sd, nd, 133.33333e3) == ber
    import math
    for s_mw, n_mw, ber in doc["ber_mw"]:
        assert L.gwo_ber_bpsk(10 * math.log10(s_mw), 10 * math.log10(n_mw), 133.33333e3) == ber
    for ax, ay, bx, by, f, att in doc["fspl"]:
        assert L.gwo_fspl(ax, ay, bx, by, f) == att
    for k, v in doc["maxBer"].items():
        a, b = k.split("/")
        assert L.gwo_max_correctable_ber(int(a), int(b)) == v
    assert L.gwo_thermal_noise_mw(22e6) == doc["thermal_mw"]


def test_philox_known_answers_oracle():
    """Random123 kat_vectors for philox4x32-10 (the oracle's own restatement of the algorithm)."""
    L = O.lib()
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        c, k, o = np.array(ctr, np.uint32), np.array(key, np.uint32), np.zeros(4, np.uint32)
        L.gwo_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
        assert tuple(int(x) for x in o) == want


def test_mode_m_differs_from_mode_r_as_documented():
    """Mode M counts errors since the last CHANGE and once: the 4 m link that mode R fails
    through double counting is decoded in mode M."""
    a = O.Oracle(mode=O.MODE_R)
    b = O.Oracle(mode=O.MODE_M)
    b.use_philox_masks(1, 0)
    for o in (a, b):
        o.reset()
        o.step({"device": 0, "duration": 15})
    assert a.counts()[1][0] > 0 and b.counts()[1][0] > 0


def test_degenerate_regime():
    """after ~1 s of simulated time nothing fits a window any more."""
    rs = np.random.RandomState(0)
    ora = O.Oracle()
    ora.reset()
    for _ in range(400):
        ora.step({"device": int(rs.randint(2)), "duration": int(rs.randint(20))})
    tx0, d0 = ora.counts()
    for _ in range(200):
        obs, rew, _ = ora.step({"device": int(rs.randint(2)), "duration": int(rs.randint(20))})
        assert rew == 0.0
    tx1, d1 = ora.counts()
    assert d1 == d0 and tx1 - tx0 == 200            # announcements only
