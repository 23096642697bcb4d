"""
ORACLE TEST INFRASTRUCTURE -- not product code, never imported by gymwipe_b200.

The reference's own benchmark scenario (``tests/test_benchmark.py:20-91``) driven on the dependency shims and
traced: a grid of PHY-only ``SendingDevice`` s (40 dBm, one 13 + 26 byte packet every ``SEND_INTERVAL`` after a
per-device initial delay), optionally with the mobility processes of the ``mobile_device_grid`` fixture, run
with ``SimMan.runSimulation``.  ``SendingDevice`` is imported from the reference's test module UNMODIFIED; the
fixtures' ``random.uniform`` draws (initial delays, mover delays, position offsets) are replaced by tapes so
that the C restatement and the CUDA engine can be fed the same numbers.

Only usable where ``/root/reference`` exists (the build container).
"""
import importlib.util
import os
import sys
from math import sqrt

import numpy as np

import ref_harness as H

SEND_INTERVAL = 1e-2        # tests/test_benchmark.py:17
MOVE_INTERVAL = 1e-3        # tests/test_benchmark.py:18


def _benchmark_module():
    H.setup_paths()
    name = "_gymwipe_reference_test_benchmark"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(H.REF, "tests", "test_benchmark.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def canonical(records):
    """Same-instant callbacks of DIFFERENT PHYs run in Python-set order in the reference (simtools.py:255) and
    touch only their own PHY: ber / dec records are compared per device, tx records as one sequence."""
    records = [tuple(r) for r in records]
    out = [r for r in records if r[0] in ("tx", "rx")]
    per = {}
    for r in records:
        if r[0] in ("ber", "dec"):
            per.setdefault((r[2], r[3]), []).append(r)
    for k in sorted(per):
        out += per[k]
    return out


def grid_tapes(n, seed, mobile=False, duration=1.0):
    """The random draws of the fixtures as tapes: initial send delays ~U(0, SEND_INTERVAL) per device; with
    ``mobile``: mover delays ~U(0, MOVE_INTERVAL) and position offsets ~U(-.2, .2)^2 per jump."""
    rs = np.random.RandomState(seed)
    tapes = {"delays": rs.uniform(0, SEND_INTERVAL, size=n)}
    if mobile:
        jumps = int(duration / MOVE_INTERVAL) + 2
        tapes["move_delays"] = rs.uniform(0, MOVE_INTERVAL, size=n)
        tapes["offsets"] = rs.uniform(-.2, .2, size=(n, jumps, 2))
    return tapes


def grid_positions(n):
    """``device_grid`` (tests/test_benchmark.py:63-69): device i at (i / cols, i % cols), cols = int(sqrt(n))."""
    cols = int(sqrt(n)) if n > 0 else 1
    return [(i / cols, float(i % cols)) for i in range(n)]


def run_reference_grid(n, tapes, durations, record_ber=True):
    """Builds the grid from the reference's classes and runs ``SimMan.runSimulation(d)`` for every ``d`` in
    ``durations``; returns ``{"now": [...], "records": [[...], ...]}`` (records per run, Tracer tuple format)."""
    bm = _benchmark_module()
    from gymwipe.networking.attenuation_models import FsplAttenuation
    from gymwipe.networking.physical import FrequencyBand
    from gymwipe.simtools import SimMan
    tracer = H.Tracer()
    tracer.record_ber = record_ber
    tracer.install()
    SimMan.init()
    band = FrequencyBand([FsplAttenuation])
    tracer.band_index[band] = 0
    devices = []
    for i, (x, y) in enumerate(grid_positions(n)):
        d = bm.SendingDevice(i, x, y, band, SEND_INTERVAL, float(tapes["delays"][i]))
        tracer.phy_index[d._phy] = (0, i)
        tracer.device_index[d] = (0, i)
        devices.append(d)
    if "offsets" in tapes:
        def mover(d, i):                                # tests/test_benchmark.py:75-82 with the draws from the tapes
            yield SimMan.timeout(float(tapes["move_delays"][i]))
            initialPos = d.position
            k = 0
            while True:
                xOffset, yOffset = (float(v) for v in tapes["offsets"][i, k])
                k += 1
                d.position.set(initialPos.x + xOffset, initialPos.y + yOffset)
                yield SimMan.timeout(MOVE_INTERVAL)
        for i, d in enumerate(devices):
            SimMan.process(mover(d, i))
    out = {"now": [], "records": [], "positions": []}
    for dur in durations:
        SimMan.runSimulation(dur)
        out["now"].append(SimMan.now)
        out["records"].append(tracer.take())
        out["positions"].append([(d.position.x, d.position.y) for d in devices])
    return out


def grid_scenario(n, tapes):
    """The same grid as a scenario dict of the restatement / the CUDA engine: n PHY-only senders, no RRM."""
    devs = []
    for i, (x, y) in enumerate(grid_positions(n)):
        devs.append({"role": "jammer", "x": x, "y": y, "interval": SEND_INTERVAL, "delay": float(tapes["delays"][i]),
                     "power": 40.0, "hdr": 13, "payload": 26})
    return {"assignment_duration_factor": 1000, "bands": [{"frequency": 2.4e9, "bandwidth": 22e6, "devices": devs}]}


def run_oracle_grid(n, tapes, durations):
    import gw_oracle as O
    ora = O.Oracle(grid_scenario(n, tapes), trace=True)
    if "offsets" in tapes:
        for i in range(n):
            ora.add_mover(0, i, float(tapes["move_delays"][i]), MOVE_INTERVAL, tapes["offsets"][i])
    out = {"now": [], "records": [], "positions": []}
    for dur in durations:
        ora.run_for(dur)
        out["now"].append(ora.now)
        out["records"].append(ora.take_records())
    return out


if __name__ == "__main__":
    import time
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    for n, mobile in ((2, False), (8, False), (20, False), (8, True), (20, True)):
        tapes = grid_tapes(n, 100 + n, mobile, 0.2)
        t0 = time.time()
        ref = run_reference_grid(n, tapes, [0.05, 0.15])
        t1 = time.time()
        ora = run_oracle_grid(n, tapes, [0.05, 0.15])
        ok = ref["now"] == ora["now"]
        for a, b in zip(ref["records"], ora["records"]):
            ca, cb = canonical(a), canonical(b)
            if ca != cb:
                ok = False
                for k, (x, y) in enumerate(zip(ca, cb)):
                    if x != y:
                        print("  first difference at record", k, x, y)
                        break
                print("  lengths", len(ca), len(cb))
        print("n=%d mobile=%s: reference %.2fs, %d records, restatement %s" % (n, mobile, t1 - t0, sum(len(r) for r in ref["records"]), "EQUAL" if ok else "DIFFERENT"))
