"""
Grids of PHY-only senders with in-step mobility -- the reference's own benchmark scenario
(``tests/test_benchmark.py:20-91``).

CPU: the oracle's C restatement and the host build of the CUDA engine (``gw_grid.cuh``) against golden vectors
produced by the UNMODIFIED reference (``oracle/gen_golden_grid.py``), and against each other on random grids.
GPU: the CUDA engine through the C ABI against the goldens and, at batch sizes the reference could never run,
against the oracle.
"""
import numpy as np
import pytest

import gw_oracle as O
import hostsim as HS
from util import GOLDEN_GRIDS, assert_grid_records, load_golden

MOVE_INTERVAL = 1e-3


def _oracle_run(doc_or_sc, durations, move_delays=None, offsets=None):
    sc = doc_or_sc["scenario"] if "scenario" in doc_or_sc else doc_or_sc
    ora = O.Oracle(sc, trace=True)
    if offsets is not None:
        for i in range(len(sc["bands"][0]["devices"])):
            ora.add_mover(0, i, float(move_delays[i]), MOVE_INTERVAL, np.asarray(offsets[i]))
    now, recs = [], []
    for d in durations:
        ora.run_for(d)
        now.append(ora.now)
        recs.append(ora.take_records())
    return now, recs


def _random_grid(rs, n, mobile, total):
    devs = [{"role": "jammer", "x": float(rs.uniform(-3, 3)), "y": float(rs.uniform(-3, 3)),
             "interval": float(rs.uniform(0.004, 0.012)), "delay": float(rs.uniform(0, 0.01)),
             "power": float(rs.choice([0.0, 20.0, 40.0])), "hdr": 13, "payload": int(rs.randint(8, 60))} for _ in range(n)]
    sc = {"assignment_duration_factor": 1000, "bands": [{"frequency": 2.4e9, "bandwidth": 22e6, "devices": devs}]}
    md = off = None
    if mobile:
        md = rs.uniform(0, MOVE_INTERVAL, size=n)
        off = rs.uniform(-.2, .2, size=(n, int(total / MOVE_INTERVAL) + 2, 2))
    return sc, md, off


@pytest.mark.parametrize("name", GOLDEN_GRIDS)
def test_oracle_and_core_match_reference_golden_grid(name):
    doc = load_golden(name)
    mobile, has_ber = doc["mobile"], any(r[0] == "ber" for run in doc["records"] for r in run)
    md, off = (doc["move_delays"], np.array(doc["offsets"])) if mobile else (None, None)
    now, recs = _oracle_run(doc, doc["durations"], md, off)
    assert now == doc["now"]
    h = HS.grid_run(doc["scenario"], doc["durations"], md, off)
    assert h["rc"] == 0 and h["now"] == doc["now"]
    for k in range(len(doc["durations"])):
        assert_grid_records(recs[k], doc["records"][k], mobile, has_ber, "oracle " + name)
        assert_grid_records(h["records"][k], doc["records"][k], mobile, has_ber, "core " + name)


@pytest.mark.parametrize("seed", range(4))
def test_core_grid_random_vs_oracle(seed):
    """Random positions / intervals / powers / packet sizes, 1..24 devices, static and mobile: the host build of
    the engine equals the literal restatement record for record (same notification order: exact)."""
    rs = np.random.RandomState(5100 + seed)
    for n, mobile in ((1, False), (int(rs.randint(2, 7)), False), (int(rs.randint(7, 25)), False),
                      (int(rs.randint(2, 7)), True), (int(rs.randint(7, 25)), True)):
        durations = [0.03, 0.05]
        sc, md, off = _random_grid(rs, n, mobile, sum(durations))
        now, recs = _oracle_run(sc, durations, md, off)
        h = HS.grid_run(sc, durations, md, off)
        assert h["rc"] == 0 and h["now"] == now
        for a, b in zip(h["records"], recs):
            assert_grid_records(a, b, False, True, "n=%d mobile=%s" % (n, mobile))
        ntx = sum(1 for run in recs for r in run if r[0] == "tx")
        assert int(h["stats"][:, 0].sum()) == ntx


# ---------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------

def _gpu_grid_from_doc(doc, num_envs=1):
    import torch
    from gymwipe_b200.envs import SendingDeviceGrid
    n = doc["n"]
    devs = doc["scenario"]["bands"][0]["devices"]
    pos = torch.tensor([[d["x"], d["y"]] for d in devs], dtype=torch.float64).expand(num_envs, n, 2)
    delays = torch.tensor(doc["delays"], dtype=torch.float64).expand(num_envs, n)
    kw = {}
    if doc["mobile"]:
        off = torch.tensor(doc["offsets"], dtype=torch.float64)
        kw = {"offsets": off.expand(num_envs, *off.shape), "move_delays": torch.tensor(doc["move_delays"], dtype=torch.float64).expand(num_envs, n)}
    return SendingDeviceGrid(num_envs, n, positions=pos.contiguous(), delays=delays.contiguous(), **{k: v.contiguous() for k, v in kw.items()})


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN_GRIDS)
def test_cuda_grid_matches_reference_golden(name):
    doc = load_golden(name)
    has_ber = any(r[0] == "ber" for run in doc["records"] for r in run)
    grid = _gpu_grid_from_doc(doc, num_envs=3)            # three identical grids: all equal the reference
    for k, d in enumerate(doc["durations"]):
        recs = grid.run_traced(d)
        assert (grid.now.cpu().numpy() == doc["now"][k]).all()
        for e in range(3):
            assert_grid_records(recs[e], doc["records"][k], doc["mobile"], has_ber, "cuda " + name)
    if doc["mobile"]:
        p = grid.positions().cpu().numpy()
        want = np.array(doc["positions"][-1])
        assert np.allclose(p[0, :, 0], want[:, 0], rtol=0, atol=1e-12) and np.allclose(p[1, :, 0], want[:, 1], rtol=0, atol=1e-12)
    grid.check()


@pytest.mark.gpu
def test_cuda_grid_batch_matches_oracle():
    """256 different mobile 20-device grids (per-env delays, offsets, positions) for 60 ms: clock, per-device
    statistics and -- for a sample -- the traces equal the oracle's."""
    import torch
    from gymwipe_b200.envs import SendingDeviceGrid
    rs = np.random.RandomState(77)
    ne, n, total = 256, 20, 0.06
    jumps = int(total / MOVE_INTERVAL) + 2
    pos = rs.uniform(-3, 3, size=(ne, n, 2))
    delays = rs.uniform(0, 0.01, size=(ne, n))
    md = rs.uniform(0, MOVE_INTERVAL, size=(ne, n))
    off = rs.uniform(-.2, .2, size=(ne, n, jumps, 2))
    grid = SendingDeviceGrid(ne, n, positions=torch.as_tensor(pos), delays=torch.as_tensor(delays),
                             move_delays=torch.as_tensor(md), offsets=torch.as_tensor(off))
    recs = grid.run_traced(total, cap=40000)
    st = grid.stats().cpu().numpy()
    now = grid.now.cpu().numpy()
    for e in list(range(6)) + [ne - 1]:
        devs = [{"role": "jammer", "x": float(pos[e, i, 0]), "y": float(pos[e, i, 1]), "interval": 1e-2,
                 "delay": float(delays[e, i]), "power": 40.0, "hdr": 13, "payload": 26} for i in range(n)]
        sc = {"assignment_duration_factor": 1000, "bands": [{"frequency": 2.4e9, "bandwidth": 22e6, "devices": devs}]}
        onow, orecs = _oracle_run(sc, [total], md[e], off[e])
        assert now[e] == onow[0]
        assert_grid_records(recs[e], orecs[0], False, True, "env %d" % e)
        assert int(st[0, :, e].sum()) == sum(1 for r in orecs[0] if r[0] == "tx")
        ok = sum(1 for r in orecs[0] if r[0] == "dec" and r[4] == 1 and r[7])
        assert int(st[3, :, e].sum()) == ok
    assert (now == now[0]).all() and st[0].sum() > 0


@pytest.mark.gpu
def test_cuda_grid_default_fixture_and_untraced_run():
    """Defaults = the reference's fixtures (grid positions, random delays from a seed); the untraced run gives
    the same statistics as the traced one."""
    from gymwipe_b200.envs import SendingDeviceGrid
    a = SendingDeviceGrid(64, 8, mobile=True, max_moves=80, seed=5)
    b = SendingDeviceGrid(64, 8, mobile=True, max_moves=80, seed=5)
    a.runSimulation(0.03)
    a.runSimulation(0.02)
    b.run_traced(0.03)
    b.run_traced(0.02)
    assert (a.stats() == b.stats()).all() and (a.now == b.now).all() and float(a.now[0]) == 0.05
    assert (a.positions() == b.positions()).all()
    assert int(a.stats()[0].sum()) >= 64 * 8 * 3
    with pytest.raises(ValueError):
        SendingDeviceGrid(4, 25)
