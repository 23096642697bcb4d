#!/usr/bin/env python
"""
ORACLE TEST INFRASTRUCTURE -- generates the golden vectors of ``tests/golden`` by running the
UNMODIFIED reference (``/root/reference``) on the shims through ``oracle/ref_harness.py``.

The reference cannot travel to the GPU box, so its outputs are committed as small JSON
fixtures together with this script:

    python oracle/gen_golden.py            # rewrites tests/golden/*.json

Every fixture holds the scenario, the action tape and, per step, obs / reward / done / step end
time / heap events and the trace records (transmissions, BER values, decider inputs and
verdicts, RRM deliveries).  One case per process (the reference allows one env per process).
"""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
sys.path.insert(0, HERE)

import ref_harness as H  # noqa: E402
import check_restatement as CR  # noqa: E402

CASES = {
    # name: (kind, seed, steps)
    "kat_reference_test": ("kat", 0, 2),
    "default_reset_seed0": ("default", 0, 160),
    "default_noreset_seed1": ("default_noreset", 1, 60),
    "positions_seed3": ("positions", 3, 60),
    "jammer_seed5": ("jammer", 5, 60),
    "longpacket_seed7": ("long", 7, 12),
    "multiband_seed9": ("multiband", 9, 24),
    "mobility_seed13": ("mobility", 13, 80),
    "mobility_inflight_seed17": ("mobilityjam", 17, 60),
    "mobility_quirks_seed4": ("mobilityquirks", 4, 70),
    "modeM_jammer_seed11": ("maskjammer", 11, 40),
    "modeM_default_seed12": ("maskdefault", 12, 60),
    "mac_receive_kat": ("mackat", 0, 10),
    "receive_bursts_seed21": ("receive", 21, 60),
    # bands beyond CounterTrafficEnv's 2 senders + RRM template (general band engine)
    "nsenders_5s_3p_seed31": ("nsenders:5:3:1:0", 31, 80),
    "nsenders_8s_6p_seed32": ("nsenders:8:6:1:1", 32, 60),
    "nsenders_3s_16p_seed33": ("nsenders:3:16:0:0", 33, 40),
    "modeM_nsenders_4s_2p_seed34": ("masknsenders:4:2:1:0", 34, 40),
    "nsenders_mobility_5s_3p_seed35": ("nsendersmove:5:3:1:0", 35, 60),
    "nsenders_movers_4s_2p_seed36": ("nsendersmovers:4:2:1:0", 36, 40),
}

MASK_SEED, MASK_ENV = 20261018, 4242


def make_case(kind, seed, steps):
    rs = np.random.RandomState(seed)
    do_reset, use_default = True, False
    if kind == "kat":
        sc, tape, do_reset, use_default = H.default_scenario(), [{"device": 0, "duration": 3}, {"device": 1, "duration": 12}], False, True
    elif kind == "default":
        sc, tape, use_default = H.default_scenario(), H.random_actions(steps, seed=seed), True
    elif kind == "default_noreset":
        sc, tape, do_reset, use_default = H.default_scenario(), H.random_actions(steps, seed=seed), False, True
    elif kind == "positions":
        sc, tape = CR.random_scenario(rs, jammers=0, spread=2.5), H.random_actions(steps, seed=seed + 1000)
    elif kind == "jammer":
        sc, tape = CR.random_scenario(rs, jammers=1, spread=2.5), H.random_actions(steps, seed=seed + 2000)
    elif kind == "long":
        sc = CR.random_scenario(rs, jammers=1, fixed_payload=1500, spread=2.0, factor=10000)
        tape = H.random_actions(steps, seed=seed + 3000)
    elif kind == "multiband":
        sc = CR.random_scenario(rs, nbands=4, jammers=1, spread=2.5)
        tapes = [H.random_actions(steps, seed=seed + 4000 + b) for b in range(4)]
        tape = [list(x) for x in zip(*tapes)]
    elif kind == "mobility":
        sc, tape = CR.random_scenario(rs, jammers=0, spread=2.0), H.random_actions(steps, seed=seed + 8000)
    elif kind == "mobilityjam":
        # a PHY-only sender whose transmissions are regularly on the air at step boundaries, when devices move
        sc, tape = CR.random_scenario(rs, jammers=1, spread=2.0), H.random_actions(steps, seed=seed + 8500)
        sc["bands"][0]["devices"][3]["interval"] = 0.0123
        sc["bands"][0]["devices"][3]["payload"] = 120
    elif kind == "mobilityquirks":
        sc, tape = CR.random_scenario(rs, jammers=1, spread=3.0), H.random_actions(steps, seed=seed + 9500)
        sc["bands"][0]["devices"][3]["interval"] = 0.0131
    elif kind == "maskjammer":
        sc, tape = CR.random_scenario(rs, jammers=1, spread=2.5), H.random_actions(steps, seed=seed + 6000)
    elif kind == "maskdefault":
        sc, tape = H.default_scenario(), H.random_actions(steps, seed=seed + 5000)
    elif kind == "mackat":
        # the reference's MAC known-answer test (tests/networking/test_stack.py:134-235) through the env API: devices at
        # (0,0) / (1,1), RRM at (2,2); each sender hands 10 packets to its MAC, one every 1e-4 s (1- / 2-byte payloads),
        # both MACs in receive mode, ten 10 ms assignments alternately: onReceive counts 4/4/8/8/10/10
        sc = {"assignment_duration_factor": 1000, "bands": [{"frequency": 2.4e9, "bandwidth": 22e6, "devices": [
            {"role": "sender", "x": 0.0, "y": 0.0, "mult": 1, "payload": 1, "interval": 1e-4, "dest": 1, "max_ticks": 10, "receive": True},
            {"role": "sender", "x": 1.0, "y": 1.0, "mult": 1, "payload": 2, "interval": 1e-4, "dest": 0, "max_ticks": 10, "receive": True},
            {"role": "rrm", "x": 2.0, "y": 2.0}]}]}
        tape, do_reset = [{"device": t % 2, "duration": 10} for t in range(steps)], False
    elif kind == "receive":
        sc, tape = CR.random_scenario(rs, jammers=1, spread=2.0), H.random_actions(steps, seed=seed + 7000)
        sc["bands"][0]["devices"][0]["receive"] = True
        sc["bands"][0]["devices"][1]["receive"] = True
        sc["bands"][0]["devices"][1]["max_ticks"] = 45
    elif kind.startswith("nsenders:") or kind.startswith("masknsenders:") or kind.startswith("nsendersmove:") or kind.startswith("nsendersmovers:"):
        # ns MAC senders + RRM + nj PHY-only senders, with / without receive mode and finite bursts
        _, ns, nj, rcv, bursts = kind.split(":")
        sc = CR.random_scenario_n(rs, int(ns), int(nj), spread=2.5, receive=bool(int(rcv)), bursts=bool(int(bursts)))
        tape = H.random_actions(steps, seed=seed + 11000, devices=int(ns))
    else:
        raise SystemExit(kind)
    return sc, tape, do_reset, use_default


def child(name):
    kind, seed, steps = CASES[name]
    H.setup_paths()
    sc, tape, do_reset, use_default = make_case(kind, seed, steps)
    # (nsender_moves also makes the PHY-only senders busy: before the env is constructed)
    pre_moves = CR.nsender_moves(np.random.RandomState(seed + 3), sc, steps) if kind.startswith("nsendersmove:") else None
    movers = None
    if kind.startswith("nsendersmovers:"):
        movers = CR.nsender_movers(np.random.RandomState(seed + 4), len(sc["bands"][0]["devices"]), jumps=700)
    tr = H.Tracer()
    mode_m = kind.startswith("mask")
    if mode_m:
        H.install_masked_phy(lambda band, sender, seq, receiver, k0, k1, ber:
                             H.philox_mask_errors(MASK_SEED, MASK_ENV, band, sender, seq, receiver, k0, k1, ber), tr)
    env = H.make_default_env(tr) if use_default else H.ScenarioEnv(sc, tr, movers=movers)
    moves = None
    if kind == "mobility":
        # every 3rd step one device jumps to a new position (between steps: nothing is on the air)
        mrs = np.random.RandomState(seed + 1)
        moves = {t: [(0, int(mrs.randint(3)), float(mrs.uniform(-2.5, 2.5)), float(mrs.uniform(-2.5, 2.5)))]
                 for t in range(2, steps, 3)}
    if kind == "mobilityjam":
        # every other step one or two devices (ascending index) jump: SimplePhy._onAttenuationChange for the
        # transmissions that are on the air at that instant
        mrs = np.random.RandomState(seed + 1)
        moves = {}
        for t in range(1, steps, 2):
            devs = sorted(set(int(v) for v in mrs.randint(4, size=int(mrs.randint(1, 3)))))
            moves[t] = [(0, d, float(mrs.uniform(-2.5, 2.5)), float(mrs.uniform(-2.5, 2.5))) for d in devs]
    if kind.startswith("nsendersmove:"):
        moves = pre_moves
    if kind == "mobilityquirks":
        # jumps beyond STANDBY_THRESHOLD, onto another device's position and back -- from before the first
        # step on, i.e. also while the pair's attenuation model does not exist yet
        mrs = np.random.RandomState(seed + 2)
        cur = [(d["x"], d["y"]) for d in sc["bands"][0]["devices"]]
        moves = {}
        for t in range(0, steps, 2):
            d = int(mrs.randint(4))
            k = int(mrs.randint(4))
            if k == 0:
                x, y = float(mrs.uniform(4000, 6000)), float(mrs.uniform(-10, 10))
            elif k == 1:
                x, y = cur[int((d + 1 + mrs.randint(3)) % 4)]
            else:
                x, y = float(mrs.uniform(-3, 3)), float(mrs.uniform(-3, 3))
            cur[d] = (x, y)
            moves[t] = [(0, d, float(x), float(y))]
    trace = H.run_tape(env, tape, tr, do_reset=do_reset, moves=moves)
    doc = {"name": name, "kind": kind, "seed": seed, "do_reset": do_reset,
           "mode": "M" if mode_m else "R", "mask_seed": MASK_SEED if mode_m else None,
           "mask_env_id": MASK_ENV if mode_m else None,
           "generator": "oracle/gen_golden.py (unmodified reference on oracle/shims)",
           "reference_class": ("gymwipe.envs.CounterTrafficEnv" if use_default else "oracle.ref_harness.ScenarioEnv")
           + (" + oracle.ref_harness.MaskedPhy (SimplePhy subclass)" if mode_m else ""),
           "scenario": sc, "reset_obs": trace["reset_obs"],
           "moves": {str(k): v for k, v in moves.items()} if moves else None,
           "movers": {str(i): {"first_delay": m[0], "interval": m[1], "offsets": np.asarray(m[2]).tolist()}
                      for i, m in movers.items()} if movers else None,
           "steps": [{"action": s["action"], "obs": s["obs"], "reward": s["reward"], "done": s["done"],
                      "now": s["now"], "events": s["events"],
                      "records": [list(r) for r in s["records"]]} for s in trace["steps"]]}
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    print("wrote", name, "steps", len(trace["steps"]),
          "records", sum(len(s["records"]) for s in trace["steps"]))


def arithmetic_vectors():
    """Numeric vectors from the reference's own functions."""
    H.setup_paths()
    from gymwipe.networking import physical as P
    from gymwipe.networking.attenuation_models import FsplAttenuation
    from gymwipe.devices import Device
    rs = np.random.RandomState(42)
    spec = P.FrequencyBandSpec()
    mcs = P.BpskMcs(spec)
    doc = {"generator": "oracle/gen_golden.py::arithmetic_vectors", "q": [], "ber_dbm": [], "ber_mw": [],
           "fspl": [], "maxBer": {}, "thermal_mw": 1.38e-23 * (20.0 + 273.15) * 22e6 * 1000,
           "noise_power_density": P.temperatureToNoisePowerDensity(20.0)}
    for x in [0.1, 0.5, 1.0, 2.0, 3.0, 5.0] + list(rs.uniform(0.01, 8, 40)):
        doc["q"].append([float(x), P.approxQFunction(float(x))])
    th = doc["thermal_mw"]
    for _ in range(200):
        s_mw = float(10 ** rs.uniform(-9, -2))
        n_mw = float(th + (10 ** rs.uniform(-12, -4) if rs.rand() < 0.5 else 0.0))
        sd, nd = P.milliwattsToDbm(s_mw), P.milliwattsToDbm(n_mw)
        doc["ber_mw"].append([s_mw, n_mw, mcs.calculateBitErrorRate(sd, nd)])
    for sd, nd in [(-46.07482474751174, -100.50608334255742), (-52.095424660791366, -100.50608334255742),
                   (-30.0, -90.0), (-60.0, -100.50608334255742), (-40.0, -40.0)]:
        doc["ber_dbm"].append([sd, nd, mcs.calculateBitErrorRate(sd, nd)])
    for d in [1.0, 2 ** 0.5, 2.0, 4.0, 10.0, 100.0]:
        m = FsplAttenuation(spec, Device("a", 0, 0), Device("b", d, 0))
        doc["fspl"].append([0.0, 0.0, float(d), 0.0, 2.4e9, m.attenuation])
    for _ in range(100):
        ax, ay, bx, by = [float(v) for v in rs.uniform(-50, 50, 4)]
        f = float(rs.choice([2.4e9, 2.425e9, 5.0e9]))
        m = FsplAttenuation(P.FrequencyBandSpec(f), Device("a", ax, ay), Device("b", bx, by))
        doc["fspl"].append([ax, ay, bx, by, f, m.attenuation])
    from fractions import Fraction
    for k, n in [(3, 4), (1, 2), (2, 3), (5, 6), (7, 8)]:
        P.Mcs._codeRateToMaxCorrectableBer = {}
        doc["maxBer"]["%d/%d" % (k, n)] = P.BpskMcs(spec, Fraction(k, n)).maxCorrectableBer()
    with open(os.path.join(OUT, "arithmetic.json"), "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    print("wrote arithmetic")


if __name__ == "__main__":
    if len(sys.argv) > 1:
        if sys.argv[1] == "arithmetic":
            arithmetic_vectors()
        else:
            child(sys.argv[1])
    else:
        for name in CASES:
            subprocess.check_call([sys.executable, os.path.abspath(__file__), name])
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "arithmetic"])
