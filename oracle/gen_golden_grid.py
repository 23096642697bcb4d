#!/usr/bin/env python
"""
ORACLE TEST INFRASTRUCTURE -- golden vectors of the reference's benchmark grids (``tests/test_benchmark.py:20-91``:
PHY-only ``SendingDevice`` s, static and with mobility processes), produced by running the UNMODIFIED reference
on the shims through ``oracle/ref_grid.py``:

    python oracle/gen_golden_grid.py       # rewrites tests/golden/grid_*.json

A fixture holds the grid size, the tapes that replace the fixtures' ``random.uniform`` draws (initial send
delays, mover delays, position offsets), the durations passed to ``SimMan.runSimulation`` and, per run, the
clock and the trace records (transmissions, decider inputs and verdicts; BER values for the small grids).
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
sys.path.insert(0, HERE)

import ref_grid as G  # noqa: E402

CASES = {
    # name: (n, mobile, seed, durations, keep BER records)
    "grid_static_n8": (8, False, 31, [0.04, 0.06], True),
    "grid_static_n20": (20, False, 32, [0.03, 0.03], False),
    "grid_mobile_n8": (8, True, 33, [0.03, 0.03], True),
    "grid_mobile_n20": (20, True, 34, [0.025, 0.015], False),
}


def main():
    for name, (n, mobile, seed, durations, keep_ber) in CASES.items():
        tapes = G.grid_tapes(n, seed, mobile, sum(durations))
        ref = G.run_reference_grid(n, tapes, durations)
        doc = {"n": n, "mobile": mobile, "seed": seed, "durations": durations,
               "delays": [float(x) for x in tapes["delays"]],
               "scenario": G.grid_scenario(n, tapes), "now": ref["now"], "positions": ref["positions"],
               "records": [[list(r) for r in run if keep_ber or r[0] != "ber"] for run in ref["records"]]}
        if mobile:
            doc["move_delays"] = [float(x) for x in tapes["move_delays"]]
            doc["offsets"] = tapes["offsets"].tolist()
        path = os.path.join(OUT, name + ".json")
        with open(path, "w") as f:
            json.dump(doc, f, separators=(",", ":"))
        print(name, "%d records, %d bytes" % (sum(len(r) for r in doc["records"]), os.path.getsize(path)))


if __name__ == "__main__":
    main()
