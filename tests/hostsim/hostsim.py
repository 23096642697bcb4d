"""
TEST INFRASTRUCTURE -- ctypes wrapper of the host build of the device simulation core
(``tests/hostsim/gw_hostsim.cpp`` includes ``gymwipe_b200/csrc/gw_core.cuh``).  Lets the
``-m "not gpu"`` suite compare the kernel's event logic with the oracle without a GPU.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_hostsim.so")
SRC = os.path.join(HERE, "gw_hostsim.cpp")
CORE = os.path.join(HERE, "..", "..", "gymwipe_b200", "csrc", "gw_core.cuh")
PEND = os.path.join(HERE, "..", "..", "gymwipe_b200", "csrc", "gw_pendulum.cuh")
GRID = os.path.join(HERE, "..", "..", "gymwipe_b200", "csrc", "gw_grid.cuh")
BAND = os.path.join(HERE, "..", "..", "gymwipe_b200", "csrc", "gw_band.cuh")

MAXDEV, MAXSEND, MAXBAND = 4, 2, 4


class HsBand(C.Structure):
    _fields_ = [("ns", C.c_int32), ("nj", C.c_int32), ("frequency", C.c_double), ("bandwidth", C.c_double),
                ("x", C.c_double * MAXDEV), ("y", C.c_double * MAXDEV), ("power", C.c_double * MAXDEV),
                ("mult", C.c_int32 * MAXSEND), ("payloadRule", C.c_int32 * MAXSEND),
                ("interval", C.c_double * MAXSEND),
                ("maxTicks", C.c_int32 * MAXSEND), ("recv", C.c_int32 * MAXSEND),
                ("jamInterval", C.c_double), ("jamDelay", C.c_double),
                ("jamHdr", C.c_int32), ("jamPay", C.c_int32)]


class HsMove(C.Structure):
    _fields_ = [("step", C.c_int32), ("band", C.c_int32), ("dev", C.c_int32), ("x", C.c_double), ("y", C.c_double)]


class HsScenario(C.Structure):
    _fields_ = [("nbands", C.c_int32), ("factor", C.c_int32), ("mode", C.c_int32), ("seed", C.c_uint64),
                ("band", HsBand * MAXBAND)]


_lib = None


def build(force=False):
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(CORE), os.path.getmtime(PEND), os.path.getmtime(GRID), os.path.getmtime(BAND)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared",
                               "-o", SO, SRC])
    return SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(SO)
        L.hs_run.restype = C.c_int
        L.hs_run.argtypes = [C.POINTER(HsScenario), C.c_int64, C.c_int, C.c_int] + [C.c_void_p] * 9 + [C.c_int64]
        L.hs_run_moves.restype = C.c_int
        L.hs_run_moves.argtypes = [C.POINTER(HsScenario), C.c_int64, C.c_int, C.c_int] + [C.c_void_p] * 9 + [C.c_int64, C.c_void_p, C.c_int]
        L.hs_ber.restype = C.c_double
        L.hs_ber.argtypes = [C.c_double, C.c_double]
        L.hs_fspl.restype = C.c_double
        L.hs_fspl.argtypes = [C.c_double] * 5
        L.hs_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.hs_fmod_slot.restype = C.c_double
        L.hs_fmod_slot.argtypes = [C.c_double]
        L.hs_within_max_ber.restype = C.c_int
        L.hs_within_max_ber.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int]
        L.hs_airtime.restype = C.c_double
        L.hs_airtime.argtypes = [C.c_int]
        L.hs_pendulum_advance.argtypes = [C.c_void_p, C.c_void_p, C.c_double]
        L.hs_set_fed_masks.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.hs_mask_errors.restype = C.c_int64
        L.hs_mask_errors.argtypes = [C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_uint32, C.c_int,
                                     C.c_int64, C.c_int64, C.c_double]
        _lib = L
    return _lib


def scenario_from_dict(d, mode=0, seed=0):
    sc = HsScenario()
    sc.nbands = len(d["bands"])
    sc.factor = int(d.get("assignment_duration_factor", 1000))
    sc.mode = mode
    sc.seed = seed
    for b, bd in enumerate(d["bands"]):
        hb = sc.band[b]
        devs = bd["devices"]
        roles = [x["role"] for x in devs]
        hb.ns = roles.count("sender")
        hb.nj = roles.count("jammer")
        assert roles == ["sender"] * hb.ns + ["rrm"] + ["jammer"] * hb.nj
        hb.frequency = float(bd.get("frequency", 2.4e9))
        hb.bandwidth = float(bd.get("bandwidth", 22e6))
        for i, x in enumerate(devs):
            hb.x[i], hb.y[i] = float(x["x"]), float(x["y"])
            hb.power[i] = float(x.get("power", 0.0)) if x["role"] == "jammer" else 0.0
            if x["role"] == "sender":
                hb.mult[i] = int(x["mult"])
                p = x.get("payload", "counter")
                hb.payloadRule[i] = -1 if p == "counter" else int(p)
                hb.interval[i] = float(x.get("interval", 0.001))
                hb.maxTicks[i] = int(x.get("max_ticks", 0))
                hb.recv[i] = 1 if x.get("receive") else 0
                assert int(x["dest"]) == 1 - i
            elif x["role"] == "jammer":
                hb.jamInterval = float(x["interval"])
                hb.jamDelay = float(x["delay"])
                hb.jamHdr = int(x.get("hdr", 13))
                hb.jamPay = int(x["payload"])
    return sc


def run(scenario, dev_tape, dur_tape, pos=None, do_reset=True, mode=0, seed=0, env_offset=0, macros=True, moves=None,
        fed_words=None, fed_slots=0):
    L = lib()
    L.hs_set_no_macro(-1 if macros else 1)
    if fed_words is not None:           # mode 2: uint32 [nenv][nbands][4][slots][4][words_per_row]
        fed_words = np.ascontiguousarray(fed_words, dtype=np.uint32)
        L.hs_set_fed_masks(fed_words.ctypes.data_as(C.c_void_p), int(fed_slots), int(fed_words.shape[-1]))
    sc = scenario_from_dict(scenario, mode, seed) if isinstance(scenario, dict) else scenario
    nb = sc.nbands
    dev_tape = np.ascontiguousarray(dev_tape, dtype=np.int32)
    dur_tape = np.ascontiguousarray(dur_tape, dtype=np.int32)
    if dev_tape.ndim == 2:
        dev_tape, dur_tape = dev_tape[:, :, None], dur_tape[:, :, None]
    nsteps, nenv, _ = dev_tape.shape
    obs = np.zeros((nsteps, nenv, nb), np.int64)
    rew = np.zeros((nsteps, nenv, nb), np.float64)
    done = np.zeros((nsteps, nenv, nb), np.uint8)
    now = np.zeros((nsteps, nenv), np.float64)
    counts = np.zeros((nenv, nb, 9), np.int64)
    power = np.zeros((nenv, nb, MAXDEV), np.float64)
    if pos is not None:
        pos = np.ascontiguousarray(pos[:, :, :MAXDEV, :], dtype=np.float64)

    def ptr(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)
    # moves: {step: [(band, dev, x, y), ...]} -- devices that jump before that step, in every env
    mv = [(t, b, d, x, y) for t, lst in sorted((moves or {}).items()) for (b, d, x, y) in lst]
    arr = (HsMove * max(len(mv), 1))()
    for k, (t, b, d, x, y) in enumerate(mv):
        arr[k].step, arr[k].band, arr[k].dev, arr[k].x, arr[k].y = int(t), int(b), int(d), float(x), float(y)
    rc = L.hs_run_moves(C.byref(sc), nenv, nsteps, 1 if do_reset else 0, ptr(pos), ptr(dev_tape), ptr(dur_tape),
                        ptr(obs), ptr(rew), ptr(done), ptr(now), ptr(counts), ptr(power), env_offset,
                        C.cast(arr, C.c_void_p), len(mv))
    ms = (C.c_longlong * 2)()
    L.hs_macro_stats(ms)
    return {"macro_tx": int(ms[0]), "macro_tail": int(ms[1]), "rc": rc, "obs": obs, "reward": rew, "done": done, "now": now, "counts": counts, "power": power}


def grid_run(scenario, durations, move_delays=None, offsets=None, move_interval=1e-3, trace_cap=400000):
    """Grid of PHY-only senders (``gw_grid.cuh``) on the host: ``scenario`` is a one-band scenario dict of
    ``jammer`` devices; ``offsets`` float64 ``[n, k, 2]`` (accumulating jumps) switches the mobility processes on.
    Returns ``{"now": [...], "records": [[...], ...], "stats": uint32 [n, 6]}`` (records in Tracer tuple format)."""
    L = lib()
    devs = scenario["bands"][0]["devices"]
    assert all(d["role"] == "jammer" for d in devs)
    n = len(devs)
    f64 = lambda xs: np.ascontiguousarray(xs, dtype=np.float64)
    power, interval = f64([d.get("power", 0.0) for d in devs]), f64([d["interval"] for d in devs])
    hdr = np.ascontiguousarray([d.get("hdr", 13) for d in devs], dtype=np.int32)
    pay = np.ascontiguousarray([d["payload"] for d in devs], dtype=np.int32)
    pos = f64([[d["x"], d["y"]] for d in devs])
    delays = f64([d["delay"] for d in devs])
    mobile = offsets is not None
    off = f64(offsets) if mobile else np.zeros((n, 1, 2))
    md = f64(move_delays) if mobile else np.zeros(n)
    dur = f64(durations)
    now = np.zeros(len(dur))
    trace = np.zeros((trace_cap, 8))
    counts = np.zeros(len(dur), np.int32)
    stats = np.zeros((n, 6), np.uint32)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    L.hs_grid_run.restype = C.c_int
    L.hs_grid_run.argtypes = [C.c_int, C.c_double, C.c_double] + [C.c_void_p] * 8 + [C.c_int, C.c_double, C.c_void_p, C.c_int,
                                                                                      C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    band = scenario["bands"][0]
    rc = L.hs_grid_run(n, float(band.get("frequency", 2.4e9)), float(band.get("bandwidth", 22e6)), ptr(power), ptr(interval),
                       ptr(hdr), ptr(pay), ptr(pos), ptr(delays), ptr(md), ptr(off), off.shape[1] if mobile else 0,
                       float(move_interval), ptr(dur), len(dur), ptr(now), ptr(trace), trace_cap, ptr(counts), ptr(stats))
    assert int(counts.sum()) <= trace_cap, "raise trace_cap"
    records, at = [], 0
    for k in range(len(dur)):
        recs = []
        for r in trace[at:at + counts[k]]:
            kind = int(r[0])
            if kind == 1:
                recs.append(("tx", float(r[1]), 0, int(r[2]), float(r[3]), float(r[4]), float(r[5])))
            elif kind == 2:
                recs.append(("ber", float(r[1]), 0, int(r[2]), float(r[3])))
            elif kind == 3:
                recs.append(("dec", float(r[1]), 0, int(r[2]), int(r[3]), float(r[4]), float(r[5]), bool(r[6])))
        at += counts[k]
        records.append(recs)
    return {"rc": rc, "now": list(now), "records": records, "stats": stats}


def gen_run(scenario, dev_tape, dur_tape, pos=None, do_reset=True, reset_at=None, trace_cap=400000, mode=0, seed=0, env_offset=0,
            moves=None, move_delays=None, offsets=None, move_interval=1e-3, trace=True):
    """General band engine (``gw_band.cuh``) on the host: a one-band scenario dict with any number of senders
    (<= 8), the RRM and PHY-only senders (<= 16); ``dev_tape`` / ``dur_tape`` int32 ``[nsteps, nenv]``; ``pos``
    optional float64 ``[nenv, nd, 2]``.  Returns obs / reward / done / now ``[nsteps, nenv]``, ``counts``
    ``[nenv, 17]`` (transmissions, deliveries per sender, onReceive calls per sender) and the trace records of
    env 0 per step (Tracer tuple format).  ``mode`` 1: per-bit Philox error masks keyed by ``seed`` and the global
    env id ``env_offset + env``."""
    L = lib()
    band = scenario["bands"][0]
    devs = band["devices"]
    roles = [d["role"] for d in devs]
    ns, nj = roles.count("sender"), roles.count("jammer")
    assert len(scenario["bands"]) == 1 and roles == ["sender"] * ns + ["rrm"] + ["jammer"] * nj
    nd = ns + 1 + nj
    ci, cj, cd = np.zeros((6, 8), np.int32), np.zeros((2, 16), np.int32), np.zeros(40, np.float64)
    for k, d in enumerate(devs[:ns]):
        p = d.get("payload", "counter")
        ci[:5, k] = [int(d["mult"]), -1 if p == "counter" else int(p), int(d["dest"]), int(d.get("max_ticks", 0)),
                     1 if d.get("receive") else 0]
        cd[k] = float(d.get("interval", 0.001))
    for j, d in enumerate(devs[ns + 1:]):
        cj[0, j], cj[1, j] = int(d.get("hdr", 13)), int(d["payload"])
        cd[8 + j], cd[24 + j] = float(d["interval"]), float(d["delay"])
    power = np.array([float(d.get("power", 0.0)) if d["role"] == "jammer" else 0.0 for d in devs], np.float64)
    dev_tape = np.ascontiguousarray(dev_tape, dtype=np.int32)
    dur_tape = np.ascontiguousarray(dur_tape, dtype=np.int32)
    nsteps, nenv = dev_tape.shape
    if pos is None:
        p = np.ascontiguousarray([[d["x"], d["y"]] for d in devs], dtype=np.float64)
    else:
        p = np.ascontiguousarray(pos, dtype=np.float64)
        assert p.shape == (nenv, nd, 2)
    obs, rew = np.zeros((nsteps, nenv), np.int64), np.zeros((nsteps, nenv), np.float64)
    done, now = np.zeros((nsteps, nenv), np.uint8), np.zeros((nsteps, nenv), np.float64)
    counts = np.zeros((nenv, 17), np.int64)
    trace_on = trace
    trace, tc = np.zeros((trace_cap if trace_on else 1, 8)), np.zeros(nsteps, np.int32)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    L.hs_gen_run.restype = C.c_int
    L.hs_gen_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_int,
                             C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.c_int] + [C.c_void_p] * 8 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                             C.c_void_p, C.c_void_p, C.c_int, C.c_double]
    # mobility processes: move_delays float64 [nd] (< 0: the device has none), offsets float64 [nd, K, 2] (accumulating jumps)
    md = np.ascontiguousarray(move_delays, dtype=np.float64) if offsets is not None else np.zeros(1)
    off = np.ascontiguousarray(offsets, dtype=np.float64) if offsets is not None else np.zeros((1, 1, 2))
    # moves: {step: [(band, dev, x, y), ...]} -- devices that jump before that step (ascending device index), in every env
    mv = [(t, d, x, y) for t, lst in sorted((moves or {}).items()) for (_, d, x, y) in sorted(lst, key=lambda m: m[1])]
    arr = (HsMove * max(len(mv), 1))()
    for k, (t, d, x, y) in enumerate(mv):
        arr[k].step, arr[k].band, arr[k].dev, arr[k].x, arr[k].y = int(t), 0, int(d), float(x), float(y)
    if reset_at is None:
        reset_at = 0 if do_reset else -1
    rc = L.hs_gen_run(ns, nj, int(scenario.get("assignment_duration_factor", 1000)), float(band.get("frequency", 2.4e9)),
                      float(band.get("bandwidth", 22e6)), ptr(ci), ptr(cj), ptr(cd), ptr(p), 0 if pos is None else 1, ptr(power),
                      int(mode), int(seed), int(env_offset), nenv, nsteps, int(reset_at), ptr(dev_tape), ptr(dur_tape), ptr(obs), ptr(rew), ptr(done), ptr(now),
                      ptr(counts), ptr(trace) if trace_on else None, trace_cap, ptr(tc), C.cast(arr, C.c_void_p), len(mv),
                      ptr(md), ptr(off), off.shape[1] if offsets is not None else 0, float(move_interval))
    assert int(tc.sum()) <= trace_cap, "raise trace_cap"
    return {"rc": rc, "obs": obs, "reward": rew, "done": done, "now": now, "counts": counts,
            "records": records_from_trace(trace, tc)}


def records_from_trace(trace, counts, band=0):
    """Trace records (8 doubles each, ``gw_core.cuh::trace_rec``) per step -> Tracer tuples."""
    out, at = [], 0
    for n in counts:
        recs = []
        for r in trace[at:at + n]:
            kind = int(r[0])
            if kind == 1:
                recs.append(("tx", float(r[1]), band, int(r[2]), float(r[3]), float(r[4]), float(r[5])))
            elif kind == 2:
                recs.append(("ber", float(r[1]), band, int(r[2]), float(r[3])))
            elif kind == 3:
                recs.append(("dec", float(r[1]), band, int(r[2]), int(r[3]), float(r[4]), float(r[5]), bool(r[6])))
            elif kind == 4:
                recs.append(("rx", float(r[1]), band, int(r[2])))
            elif kind == 5:
                recs.append(("mrx", float(r[1]), band, int(r[2])))
        at += n
        out.append(recs)
    return out
