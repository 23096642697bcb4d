#!/usr/bin/env python
"""
ORACLE TEST INFRASTRUCTURE -- not product code.

Pins the C restatement (``oracle/gw_oracle.c``) against the reference ITSELF:
both replay the same action tapes; every trace record (transmission start /
stop / bit counts, every BER value, every decider input and verdict, every RRM
delivery) and every step result (obs, reward, done, step end time) must be
bit-identical.  BER values are additionally compared with a relative tolerance
switch for libm differences (none expected: both sides call glibc).

Runs only where ``/root/reference`` exists (the build container):

    python oracle/check_restatement.py            # default env + random scenarios
    python oracle/check_restatement.py --selfcheck  # harness ScenarioEnv == CounterTrafficEnv
"""
import argparse
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import gw_oracle as O  # noqa: E402
import ref_harness as H  # noqa: E402


def canonical(records):
    glob = [r for r in records if r[0] in ("tx", "rx", "mrx")]
    per = {}
    for r in records:
        if r[0] in ("ber", "dec"):
            per.setdefault((r[2], r[3]), []).append(r)
    out = list(glob)
    for k in sorted(per):
        out += per[k]
    return out


def compare(ref, ora, label, ber_rtol=0.0):
    """Returns the number of mismatching steps (prints the first one)."""
    bad = 0
    assert ref["reset_obs"] == ora["reset_obs"], (label, ref["reset_obs"], ora["reset_obs"])
    for i, (a, b) in enumerate(zip(ref["steps"], ora["steps"])):
        ok = (a["obs"] == b["obs"] and a["reward"] == b["reward"] and a["done"] == b["done"]
              and a["now"] == b["now"])
        # Same-instant callbacks of DIFFERENT PHYs run in Python-set order in the reference
        # (simtools.py:255) and touch only their own PHY, so records are compared per
        # (band, device) subsequence for ber/dec and as one global subsequence for tx/rx.
        ra, rb = canonical(a["records"]), canonical(b["records"])
        if ra != rb:
            ok = False
        if not ok:
            bad += 1
            if bad == 1:
                print("[%s] MISMATCH at step %d action %s" % (label, i, a["action"]))
                print("  ref: obs %s rew %s now %r nrec %d" % (a["obs"], a["reward"], a["now"], len(ra)))
                print("  ora: obs %s rew %s now %r nrec %d" % (b["obs"], b["reward"], b["now"], len(rb)))
                for j in range(max(len(ra), len(rb))):
                    x = ra[j] if j < len(ra) else None
                    y = rb[j] if j < len(rb) else None
                    if x != y:
                        print("  first differing record %d:\n    ref %s\n    ora %s" % (j, x, y))
                        break
    return bad


def compare_mobile(ref, ora, label, err_rtol=5e-2, ber_rtol=1e-4):
    """
    Comparison for devices that move while SEVERAL transmissions are on the air.  A moving device's attenuation models
    are notified in Python-set order in the reference (simtools.py:255: by object hash -- two runs of the reference
    itself differ), every notification charges the running reception with the errors since the last RESET (appendix
    B #5) at the rate of that moment, so error sums (and the rates in between) depend on that order (observed: up to 3e-2 between two runs of the reference),
    the final rate of an instant through the rounding of the power sum at the 1e-8 level (times the exponent of
    exp(-Eb/N0) for rates that are astronomically small: observed 6e-6 at rates of 1e-58).  Compared: step results,
    transmissions, deliveries exactly; decisions: time / device / section / bit count / verdict exactly, error sum
    within `err_rtol`; rates: per (device, instant) the same number of evaluations, the last one within `ber_rtol`.
    """
    bad = 0
    assert ref["reset_obs"] == ora["reset_obs"]
    for i, (a, b) in enumerate(zip(ref["steps"], ora["steps"])):
        ok = (a["obs"], a["reward"], a["done"], a["now"]) == (b["obs"], b["reward"], b["done"], b["now"])
        ra, rb = [tuple(r) for r in a["records"]], [tuple(r) for r in b["records"]]
        ok = ok and [r for r in ra if r[0] in ("tx", "rx", "mrx")] == [r for r in rb if r[0] in ("tx", "rx", "mrx")]
        key = lambda r: (r[3], r[1], r[4])
        da, db = sorted([r for r in ra if r[0] == "dec"], key=key), sorted([r for r in rb if r[0] == "dec"], key=key)
        ok = ok and len(da) == len(db)
        for x, y in zip(da, db):
            ok = ok and x[:5] == y[:5] and x[6:] == y[6:] and abs(x[5] - y[5]) <= err_rtol * max(abs(x[5]), abs(y[5]), 1e-300)
        ga, gb = {}, {}
        for recs, g in ((ra, ga), (rb, gb)):
            for r in recs:
                if r[0] == "ber":
                    g.setdefault((r[3], r[1]), []).append(r[4])
        ok = ok and sorted(ga) == sorted(gb)
        if ok:
            for k in ga:
                x, y = ga[k][-1], gb[k][-1]
                # (a rate of 0.49..0.5 is the S ~ N regime, where `sd <= nd` decides between exactly 0.5 and the formula
                # on a noise power that is mostly rounding residue: order-dependent in the reference itself)
                close = abs(x - y) <= ber_rtol * max(abs(x), abs(y), 1e-30) or (min(x, y) >= 0.45 and abs(x - y) <= 0.02)
                ok = ok and len(ga[k]) == len(gb[k]) and close
        if not ok:
            bad += 1
            if bad == 1:
                print("[%s] MISMATCH at step %d action %s" % (label, i, a["action"]))
    return bad


def run_case_m(scenario, tape, label, seed=77, env_id=12345, moves=None, mobile=False):
    """mode M: reference + MaskedPhy subclass (numpy Philox) vs the restatement (C Philox)."""
    tr = H.Tracer()
    H.setup_paths()
    H.install_masked_phy(lambda band, sender, seq, receiver, k0, k1, ber:
                         H.philox_mask_errors(seed, env_id, band, sender, seq, receiver, k0, k1, ber), tr)
    env = H.ScenarioEnv(scenario, tr)
    ref = H.run_tape(env, tape, tr, moves=moves)
    ora_env = O.Oracle(scenario, trace=True, mode=O.MODE_M)
    ora_env.use_philox_masks(seed, env_id)
    ora = O.run_tape(ora_env, tape, moves=moves)
    bad = compare_mobile(ref, ora, label) if mobile else compare(ref, ora, label)
    ndec = sum(1 for s in ref["steps"] for r in s["records"] if r[0] == "dec")
    nfail = sum(1 for s in ref["steps"] for r in s["records"] if r[0] == "dec" and not r[7])
    print("[%s] steps %d mismatching %d ; decisions %d (failed %d)" % (label, len(tape), bad, ndec, nfail))
    return bad


def run_case(scenario, tape, label, do_reset=True, use_default_class=False, moves=None, mobile=False, movers=None):
    tr = H.Tracer()
    if use_default_class:
        env = H.make_default_env(tr)
    else:
        env = H.ScenarioEnv(scenario, tr, movers=movers)
    ref = H.run_tape(env, tape, tr, do_reset=do_reset, moves=moves)
    ora_env = O.Oracle(scenario, trace=True)
    for i in sorted(movers or {}):
        ora_env.add_mover(0, i, movers[i][0], movers[i][1], movers[i][2])
    ora = O.run_tape(ora_env, tape, do_reset=do_reset, moves=moves)
    bad = compare_mobile(ref, ora, label) if mobile else compare(ref, ora, label)
    ev_ref = sum(s["events"] for s in ref["steps"])
    ev_ora = sum(s["events"] for s in ora["steps"])
    print("[%s] steps %d mismatching %d ; heap pops ref %d / restatement %d" %
          (label, len(tape), bad, ev_ref, ev_ora))
    return bad


def random_scenario(rs, nbands=1, jammers=1, fixed_payload=None, spread=20.0, factor=1000):
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
            # interval > airtime: otherwise the jammer's SEND queue grows without bound
            devs.append({"role": "jammer", "x": float(rs.uniform(-spread, spread)),
                         "y": float(rs.uniform(-spread, spread)),
                         "interval": float(airtime * rs.uniform(1.3, 6.0)),
                         "delay": float(rs.uniform(0, 1e-2)),
                         "power": float(rs.choice([0.0, 10.0, 20.0])), "hdr": 13,
                         "payload": payload})
        bands.append({"frequency": 2.4e9 + b * 25e6, "bandwidth": 22e6, "devices": devs})
    return {"assignment_duration_factor": factor, "bands": bands}


def random_scenario_n(rs, ns, nj, spread=2.5, factor=1000, receive=False, bursts=False):
    """One band with `ns` MAC senders (every sender addresses another one), the RRM and `nj` PHY-only senders --
    beyond CounterTrafficEnv's 2 + 1 template; the interpreter's observation still follows senders 0 and 1
    (counter_traffic.py:75-80: `receivedValues[0] - receivedValues[1]`)."""
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


def nsender_moves(rs, sc, steps, start=1):
    """Before every other step one to three devices (ascending index) jump -- within the band's area or far beyond
    STANDBY_THRESHOLD --; the PHY-only senders are made busy so that transmissions are on the air at step boundaries
    (SimplePhy._onAttenuationChange).  (Jumps ONTO another device's position are pinned on the four-device band of
    --case mobilityquirks: with several equal-power signals from one spot on the air the noise power is signal-minus-
    signal residue, and whether a rate is 0.5 or 0.49 depends on the reference's own set iteration order.)"""
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
            if int(rs.randint(5)) == 0:
                x, y = float(rs.uniform(4000, 6000)), float(rs.uniform(-10, 10))
            else:
                x, y = float(rs.uniform(-3, 3)), float(rs.uniform(-3, 3))
            cur[d] = (x, y)
            lst.append((0, d, float(x), float(y)))
        moves[t] = lst
    return moves


def nsender_movers(rs, nd, jumps=1500, interval=1e-3):
    """{device: (first delay, interval, offsets)}: most devices get a mobility process (jumps of up to +-0.2 m per axis
    every millisecond, as the mobile_device_grid fixture draws them)."""
    out = {}
    for d in range(nd):
        if rs.rand() < 0.75:
            out[d] = (float(rs.uniform(0, interval)), float(interval), rs.uniform(-.2, .2, size=(jumps, 2)))
    return out


def child(args):
    """One case per process: the reference allows one env per process."""
    rs = np.random.RandomState(args.seed)
    if args.case == "default":
        tape = H.random_actions(args.steps, seed=args.seed)
        return run_case(H.default_scenario(), tape, "default seed %d" % args.seed,
                        use_default_class=True)
    if args.case == "default_noreset":
        tape = H.random_actions(args.steps, seed=args.seed)
        return run_case(H.default_scenario(), tape, "default/no-reset seed %d" % args.seed,
                        do_reset=False, use_default_class=True)
    if args.case == "kat":
        tape = [{"device": 0, "duration": 3}, {"device": 1, "duration": 12}]
        return run_case(H.default_scenario(), tape, "reference KAT", do_reset=False,
                        use_default_class=True)
    if args.case == "positions":
        sc = random_scenario(rs, jammers=0, spread=args.spread)
        tape = H.random_actions(args.steps, seed=args.seed + 1000)
        return run_case(sc, tape, "positions seed %d" % args.seed)
    if args.case == "jammer":
        sc = random_scenario(rs, jammers=args.jammers, spread=args.spread)
        tape = H.random_actions(args.steps, seed=args.seed + 2000)
        return run_case(sc, tape, "jammer seed %d" % args.seed)
    if args.case == "long":
        sc = random_scenario(rs, jammers=args.jammers, fixed_payload=1500, spread=args.spread,
                             factor=10000)
        tape = H.random_actions(args.steps, seed=args.seed + 3000)
        return run_case(sc, tape, "long-packet seed %d" % args.seed)
    if args.case == "maskdefault":
        tape = H.random_actions(args.steps, seed=args.seed + 5000)
        return run_case_m(H.default_scenario(), tape, "mode M default seed %d" % args.seed, seed=args.seed + 77)
    if args.case == "maskjammer":
        sc = random_scenario(rs, jammers=1, spread=args.spread)
        tape = H.random_actions(args.steps, seed=args.seed + 6000)
        return run_case_m(sc, tape, "mode M jammer seed %d" % args.seed, seed=args.seed + 78)
    if args.case == "masklong":
        sc = random_scenario(rs, jammers=1, fixed_payload=1500, spread=args.spread, factor=10000)
        tape = H.random_actions(args.steps, seed=args.seed + 7000)
        return run_case_m(sc, tape, "mode M long-packet seed %d" % args.seed, seed=args.seed + 79)
    if args.case in ("mobilityjam", "maskmobilityjam"):
        # devices jump before every other step while a PHY-only sender keeps the band busy:
        # SimplePhy._onAttenuationChange for the transmissions that are on the air
        sc = random_scenario(rs, jammers=1, spread=3.0)
        sc["bands"][0]["devices"][3]["interval"] = float(rs.uniform(0.008, 0.02))
        tape = H.random_actions(args.steps, seed=args.seed + 9000)
        moves = {}
        for t in range(1, args.steps, 2):
            devs = sorted(set(int(v) for v in rs.randint(4, size=int(rs.randint(1, 4)))))
            moves[t] = [(0, d, float(rs.uniform(-3, 3)), float(rs.uniform(-3, 3))) for d in devs]
        if args.case == "mobilityjam":
            return run_case(sc, tape, "mobility with transmissions on the air, seed %d" % args.seed, moves=moves)
        return run_case_m(sc, tape, "mode M mobility with transmissions on the air, seed %d" % args.seed,
                          seed=args.seed + 80, moves=moves)
    if args.case == "mobilityquirks":
        # the corner cases of PositionalAttenuationModel / FsplAttenuation._update: a device jumps beyond
        # STANDBY_THRESHOLD (the model stops following), onto another device's position (the model keeps its
        # value) and back, while a PHY-only sender keeps the band busy
        sc = random_scenario(rs, jammers=1, spread=3.0)
        sc["bands"][0]["devices"][3]["interval"] = float(rs.uniform(0.008, 0.02))
        tape = H.random_actions(args.steps, seed=args.seed + 9500)
        devs = sc["bands"][0]["devices"]
        moves = {}
        for t in range(1, args.steps, 2):
            d = int(rs.randint(4))
            kind = int(rs.randint(4))
            if kind == 0:
                x, y = float(rs.uniform(4000, 6000)), float(rs.uniform(-10, 10))          # far away
            elif kind == 1:
                o = int((d + 1 + rs.randint(3)) % 4)                                       # onto another device
                x, y = float(devs[o]["x"]), float(devs[o]["y"])
            else:
                x, y = float(rs.uniform(-3, 3)), float(rs.uniform(-3, 3))
            devs[d] = dict(devs[d], x=x, y=y)                                               # track for "onto another device"
            moves[t] = [(0, d, x, y)]
        sc0 = random_scenario(np.random.RandomState(args.seed), jammers=1, spread=3.0)      # the scenario as constructed
        sc0["bands"][0]["devices"][3]["interval"] = sc["bands"][0]["devices"][3]["interval"]
        return run_case(sc0, tape, "mobility corner cases, seed %d" % args.seed, moves=moves)
    if args.case == "nsenders":
        # bands beyond the 2 + 1 template: 3..8 MAC senders, 0..6 PHY-only senders, receive mode and bursts
        ns, nj = int(rs.randint(3, 9)), int(rs.randint(0, 7))
        sc = random_scenario_n(rs, ns, nj, spread=args.spread, receive=bool(args.seed % 2), bursts=bool(args.seed % 3 == 0))
        tape = H.random_actions(args.steps, seed=args.seed + 11000, devices=ns)
        return run_case(sc, tape, "%d senders + RRM + %d PHY-only senders, seed %d" % (ns, nj, args.seed))
    if args.case in ("nsendersmobility", "masknsendersmobility"):
        ns, nj = int(rs.randint(3, 7)), int(rs.randint(1, 5))
        sc = random_scenario_n(rs, ns, nj, spread=args.spread, receive=bool(args.seed % 2))
        tape = H.random_actions(args.steps, seed=args.seed + 13000, devices=ns)
        moves = nsender_moves(rs, sc, args.steps, start=args.seed % 2)
        label = "%d senders + RRM + %d PHY-only senders moving between steps, seed %d" % (ns, nj, args.seed)
        if args.case == "nsendersmobility":
            return run_case(sc, tape, label, moves=moves, mobile=True)
        return run_case_m(sc, tape, "mode M, " + label, seed=args.seed + 82, moves=moves, mobile=True)
    if args.case == "nsendersmovers":
        # mobility processes DURING the steps (the mover of tests/test_benchmark.py:73-85) on a band with MACs and an RRM
        ns, nj = int(rs.randint(3, 7)), int(rs.randint(0, 4))
        sc = random_scenario_n(rs, ns, nj, spread=args.spread, receive=bool(args.seed % 2))
        tape = H.random_actions(args.steps, seed=args.seed + 14000, devices=ns)
        movers = nsender_movers(rs, ns + 1 + nj)
        return run_case(sc, tape, "%d senders + RRM + %d PHY-only senders with mobility processes, seed %d" % (ns, nj, args.seed),
                        mobile=True, movers=movers)
    if args.case == "masknsenders":
        ns, nj = int(rs.randint(3, 7)), int(rs.randint(0, 4))
        sc = random_scenario_n(rs, ns, nj, spread=args.spread, receive=bool(args.seed % 2))
        tape = H.random_actions(args.steps, seed=args.seed + 12000, devices=ns)
        return run_case_m(sc, tape, "mode M, %d senders + RRM + %d PHY-only senders, seed %d" % (ns, nj, args.seed), seed=args.seed + 81)
    if args.case == "multiband":
        sc = random_scenario(rs, nbands=4, jammers=1, spread=args.spread)
        tapes = [H.random_actions(args.steps, seed=args.seed + 4000 + b) for b in range(4)]
        tape = [list(x) for x in zip(*tapes)]
        return run_case(sc, tape, "multiband seed %d" % args.seed)
    raise SystemExit("unknown case")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--jammers", type=int, default=1)
    ap.add_argument("--spread", type=float, default=8.0)
    ap.add_argument("--selfcheck", action="store_true")
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()

    H.setup_paths()
    if args.selfcheck:
        tape = H.random_actions(300, seed=0)
        tr = H.Tracer()
        a = H.run_tape(H.make_default_env(tr), tape, tr)
        tr2 = H.Tracer()
        b = H.run_tape(H.ScenarioEnv(H.default_scenario(), tr2), tape, tr2)
        print("ScenarioEnv(default) == CounterTrafficEnv:", a == b)
        return 0 if a == b else 1
    if args.case:
        return 1 if child(args) else 0

    # driver: every case in its own process
    O.build()
    plan = [("kat", 0, 2), ("default", 0, 400), ("default", 1, 400), ("default_noreset", 2, 200)]
    nseeds = 2 if args.quick else 6
    for sd in range(nseeds):
        plan += [("positions", sd, 200), ("jammer", sd, 200), ("long", sd, 40), ("multiband", sd, 80),
                 ("maskdefault", sd, 120), ("maskjammer", sd, 120), ("masklong", sd, 20),
                 ("mobilityjam", sd, 120), ("maskmobilityjam", sd, 80), ("mobilityquirks", sd, 100),
                 ("nsenders", sd, 120), ("masknsenders", sd, 40), ("nsendersmobility", sd, 80),
                 ("masknsendersmobility", sd, 30), ("nsendersmovers", sd, 60)]
    failed = 0
    for case, sd, steps in plan:
        rc = subprocess.call([sys.executable, os.path.abspath(__file__), "--case", case,
                              "--seed", str(sd), "--steps", str(steps),
                              "--jammers", str(args.jammers), "--spread", str(args.spread)])
        failed += 1 if rc else 0
    print("FAILED CASES:", failed)
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
