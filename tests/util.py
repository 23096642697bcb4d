"""Shared helpers of the test-suite (golden fixtures, canonical record order, random scenarios)."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

GOLDEN_CASES = ["kat_reference_test", "default_reset_seed0", "default_noreset_seed1", "positions_seed3",
                "jammer_seed5", "longpacket_seed7", "multiband_seed9", "mac_receive_kat", "receive_bursts_seed21"]


GOLDEN_CASES_M = ["modeM_jammer_seed11", "modeM_default_seed12"]


def load_golden(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


def canonical(records):
    """
    Same-instant callbacks of DIFFERENT PHYs run in Python-set order in the reference
    (gymwipe/simtools.py:255) and touch only their own PHY: compare ber/dec records per
    (band, device) subsequence and tx/rx records as one global subsequence.
    """
    records = [tuple(r) for r in records]
    glob = [r for r in records if r[0] in ("tx", "rx", "mrx")]
    per = {}
    for r in records:
        if r[0] in ("ber", "dec"):
            per.setdefault((r[2], r[3]), []).append(r)
    out = list(glob)
    for k in sorted(per):
        out += per[k]
    return out


def canonical_per_band(records):
    """Like :func:`canonical`, with the tx/rx records grouped per band as well: the bands of an
    env are independent simulations, only their order WITHIN a band is meaningful."""
    records = [tuple(r) for r in records]
    out = []
    for b in sorted({r[2] for r in records}):
        out += canonical([r for r in records if r[2] == b])
    return out


def tapes_from_golden(doc):
    """-> dev, dur int32 arrays [steps, 1, nbands]"""
    nb = len(doc["scenario"]["bands"])
    steps = doc["steps"]
    dev = np.zeros((len(steps), 1, nb), np.int32)
    dur = np.zeros((len(steps), 1, nb), np.int32)
    for t, s in enumerate(steps):
        acts = s["action"] if isinstance(s["action"], list) else [s["action"]]
        for b, a in enumerate(acts):
            dev[t, 0, b] = a["device"]
            dur[t, 0, b] = a["duration"]
    return dev, dur


def golden_results(doc):
    nb = len(doc["scenario"]["bands"])
    steps = doc["steps"]
    obs = np.array([s["obs"] if nb > 1 else [s["obs"]] for s in steps], np.int64)
    rew = np.array([s["reward"] if nb > 1 else [s["reward"]] for s in steps], np.float64)
    done = np.array([s["done"] if nb > 1 else [s["done"]] for s in steps], np.uint8)
    now = np.array([s["now"] for s in steps], np.float64)
    return obs, rew, done, now


def random_scenario(rs, nbands=1, jammers=1, fixed_payload=None, spread=3.0, factor=1000):
    """Random geometry / traffic / jammer scenario (same generator as oracle/check_restatement.py)."""
    bands = []
    for b in range(nbands):
        devs = []
        for k in range(2):
            devs.append({"role": "sender", "x": float(rs.uniform(-spread, spread)),
                         "y": float(rs.uniform(-spread, spread)), "mult": int(rs.randint(1, 4)),
                         "payload": "counter" if fixed_payload is None else int(fixed_payload),
                         "interval": 0.001, "dest": 1 - k})
        devs.append({"role": "rrm", "x": float(rs.uniform(-spread, spread)),
                     "y": float(rs.uniform(-spread, spread))})
        for j in range(jammers):
            payload = int(rs.randint(12, 200))
            airtime = (13 + payload) * 8 / 99999.9975
            devs.append({"role": "jammer", "x": float(rs.uniform(-spread, spread)),
                         "y": float(rs.uniform(-spread, spread)),
                         "interval": float(airtime * rs.uniform(1.3, 6.0)),
                         "delay": float(rs.uniform(0, 1e-2)),
                         "power": float(rs.choice([0.0, 10.0, 20.0])), "hdr": 13, "payload": payload})
        bands.append({"frequency": 2.4e9 + b * 25e6, "bandwidth": 22e6, "devices": devs})
    return {"assignment_duration_factor": factor, "bands": bands}


def random_tapes(rs, nsteps, nenv, nb):
    dev = rs.randint(0, 2, size=(nsteps, nenv, nb)).astype(np.int32)
    dur = rs.randint(0, 20, size=(nsteps, nenv, nb)).astype(np.int32)
    return dev, dur


GOLDEN_NSENDERS = ["nsenders_5s_3p_seed31", "nsenders_8s_6p_seed32", "nsenders_3s_16p_seed33"]
GOLDEN_NSENDERS_MOBILITY = "nsenders_mobility_5s_3p_seed35"
GOLDEN_NSENDERS_MOVERS = "nsenders_movers_4s_2p_seed36"

BER_RTOL = 1e-9          # north star: 1e-6 relative in fp64; observed ~1e-15 (libm / libdevice pow, log10 differ by <= 2 ulp)


def assert_step_records(got, want, label=""):
    """Trace records of one step against the reference's (or the oracle's): same records in the same canonical
    order; transmissions, deliveries, decider inputs (section, bit count) and verdicts bit-exact, BER values and
    expected error sums within ``BER_RTOL`` relative.  Returns the largest relative deviation seen."""
    got, want = canonical(got), canonical(want)
    assert len(got) == len(want), (label, len(got), len(want))
    worst = 0.0
    for g, w in zip(got, want):
        assert g[:4] == w[:4], (label, g, w)
        if g[0] == "tx":
            assert tuple(g[4:]) == tuple(w[4:]), (label, g, w)
        elif g[0] == "ber":
            rel = abs(g[4] - w[4]) / max(abs(w[4]), 1e-300)
            worst = max(worst, rel)
            assert rel <= BER_RTOL, (label, g, w)
        elif g[0] == "dec":
            assert g[4] == w[4] and g[6] == w[6] and g[7] == w[7], (label, g, w)
            rel = abs(g[5] - w[5]) / max(abs(w[5]), 1e-300) if w[5] != 0 else abs(g[5])
            worst = max(worst, rel)
            assert rel <= BER_RTOL, (label, g, w)
    return worst


def assert_mobile_step_records(got, want, label="", err_rtol=5e-2, ber_rtol=1e-4):
    """Trace records of one step against the REFERENCE's when devices moved while several transmissions were on the
    air: a moving device's attenuation models are notified in Python-set order in the reference (``simtools.py:255``:
    by object hash -- two runs of the reference itself differ), every notification charges the running reception
    with the errors since the last RESET (appendix B #5) at the rate of that moment, so error sums depend on that
    order (up to 3e-2 between two runs of the reference itself) and the final rate of an instant at the 1e-8 level (more for astronomically small
    rates).  Transmissions, deliveries, decider inputs (section, bit count) and verdicts: exact."""
    got, want = [tuple(r) for r in got], [tuple(r) for r in want]
    assert [r for r in got if r[0] in ("tx", "rx", "mrx")] == [r for r in want if r[0] in ("tx", "rx", "mrx")], label
    key = lambda r: (r[3], r[1], r[4])
    gd, wd = sorted([r for r in got if r[0] == "dec"], key=key), sorted([r for r in want if r[0] == "dec"], key=key)
    assert len(gd) == len(wd), label
    for a, b in zip(gd, wd):
        assert a[:5] == b[:5] and a[6:] == b[6:], (label, a, b)
        assert abs(a[5] - b[5]) <= err_rtol * max(abs(a[5]), abs(b[5]), 1e-300), (label, a, b)
    gg, wg = {}, {}
    for recs, g in ((got, gg), (want, wg)):
        for r in recs:
            if r[0] == "ber":
                g.setdefault((r[3], r[1]), []).append(r[4])
    assert sorted(gg) == sorted(wg), label
    for k in wg:
        assert len(gg[k]) == len(wg[k]), (label, k)
        a, b = gg[k][-1], wg[k][-1]
        # (0.49..0.5: the S ~ N regime, where `sd <= nd` decides between exactly 0.5 and the formula on a noise power
        # that is mostly rounding residue -- order-dependent in the reference itself)
        assert abs(a - b) <= ber_rtol * max(abs(a), abs(b), 1e-30) or (min(a, b) >= 0.45 and abs(a - b) <= 0.02), (label, k, a, b)


def random_scenario_n(rs, ns, nj, spread=2.5, factor=1000, receive=False, bursts=False):
    """One band with ``ns`` MAC senders, the RRM and ``nj`` PHY-only senders (same generator as
    ``oracle/check_restatement.py::random_scenario_n``, which pins the oracle against the live reference on it)."""
    devs = []
    for k in range(ns):
        dest = int((k + 1 + rs.randint(ns - 1)) % ns)
        d = {"role": "sender", "x": float(rs.uniform(-spread, spread)), "y": float(rs.uniform(-spread, spread)),
             "mult": int(rs.randint(1, 4)), "payload": "counter" if rs.rand() < 0.6 else int(rs.randint(1, 60)),
             "interval": float(rs.choice([0.001, 0.001, 0.0007, 0.0013])), "dest": dest}
        if receive and rs.rand() < 0.6:
            d["receive"] = True
        if bursts and rs.rand() < 0.3:
            d["max_ticks"] = int(rs.randint(5, 60))
        devs.append(d)
    devs.append({"role": "rrm", "x": float(rs.uniform(-spread, spread)), "y": float(rs.uniform(-spread, spread))})
    for j in range(nj):
        payload = int(rs.randint(12, 120))
        airtime = (13 + payload) * 8 / 99999.9975
        devs.append({"role": "jammer", "x": float(rs.uniform(-spread, spread)), "y": float(rs.uniform(-spread, spread)),
                     "interval": float(airtime * rs.uniform(2.0, 9.0) * max(1, nj)), "delay": float(rs.uniform(0, 1e-2)),
                     "power": float(rs.choice([0.0, 10.0, 20.0])), "hdr": 13, "payload": payload})
    return {"assignment_duration_factor": factor, "bands": [{"frequency": 2.4e9, "bandwidth": 22e6, "devices": devs}]}


GOLDEN_GRIDS = ["grid_static_n8", "grid_static_n20", "grid_mobile_n8", "grid_mobile_n20"]


def assert_grid_records(got, want, mobile, has_ber, label=""):
    """
    Trace records of a grid run against the reference's (or the oracle's).  Transmissions: identical.  Decider
    records: same time / device / section / bit count / verdict, error sum within 1e-9 relative -- with moving
    devices 1e-2: a moving device's attenuation models are notified in Python-set order in the reference
    (``simtools.py:255``: by object hash, not reproducible between two runs of the reference itself), each
    notification charges the running reception with the errors since the last RESET (appendix B #5), so the sum
    depends on that order (observed up to 1e-3 relative).  BER records: per (device, time) the same number of evaluations and
    the same final value (1e-9); with moving devices the intermediate values depend on the same order.
    """
    from collections import OrderedDict
    got, want = [tuple(r) for r in got], [tuple(r) for r in want]
    assert [r for r in got if r[0] == "tx"] == [r for r in want if r[0] == "tx"], label
    tol = 1e-2 if mobile else 1e-9
    gd, wd = [r for r in got if r[0] == "dec"], [r for r in want if r[0] == "dec"]
    key = lambda r: (r[3], r[1], r[4])
    gd, wd = sorted(gd, key=key), sorted(wd, key=key)
    assert len(gd) == len(wd), (label, len(gd), len(wd))
    for a, b in zip(gd, wd):
        assert a[:5] == b[:5] and a[6:] == b[6:], (label, a, b)
        assert abs(a[5] - b[5]) <= tol * max(abs(a[5]), abs(b[5]), 1e-300), (label, a, b)
    if not has_ber:
        return

    def groups(recs):
        out = OrderedDict()
        for r in recs:
            if r[0] == "ber":
                out.setdefault((r[3], r[1]), []).append(r[4])
        return out
    gg, wg = groups(got), groups(want)
    assert list(sorted(gg)) == list(sorted(wg)), label
    for k in wg:
        assert len(gg[k]) == len(wg[k]), (label, k)
        a, b = gg[k][-1], wg[k][-1]
        assert abs(a - b) <= 1e-9 * max(abs(a), abs(b), 1e-300), (label, k, a, b)
        if not mobile:
            for a, b in zip(gg[k], wg[k]):
                assert abs(a - b) <= 1e-9 * max(abs(a), abs(b), 1e-300), (label, k, a, b)
