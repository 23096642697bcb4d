"""
ORACLE TEST INFRASTRUCTURE -- not product code, never imported by gymwipe_b200.

ctypes binding of the plain-C restatement (``oracle/gw_oracle.c``).  Used by
``tests/`` (as the checker), ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs.
"""
import ctypes as C
import os
import subprocess
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libgw_oracle.so")

MAXDEV, MAXBAND, MAXTX = 28, 4, 28
BATCH_DEV = 8           # device stride of the batch API's pos / counts arrays
ROLE = {"sender": 1, "rrm": 2, "jammer": 3}
MODE_R, MODE_M = 0, 1
REC_TX, REC_BER, REC_DEC, REC_RX = 1, 2, 3, 4
FAULTS = {1: "heap overflow", 2: "reference would raise KeyError", 3: "reference would fail an assert",
          4: "send queue overflow", 5: "tx pool exhausted", 6: "empty schedule", 7: "internal"}


class DevSpec(C.Structure):
    _fields_ = [("role", C.c_int32), ("x", C.c_double), ("y", C.c_double),
                ("mult", C.c_int32), ("payload_rule", C.c_int32), ("dest", C.c_int32),
                ("interval", C.c_double), ("max_ticks", C.c_int32), ("receive", C.c_int32),
                ("jam_interval", C.c_double), ("jam_delay", C.c_double), ("jam_power", C.c_double),
                ("jam_hdr", C.c_int32), ("jam_payload", C.c_int32)]


class BandSpec(C.Structure):
    _fields_ = [("ndev", C.c_int32), ("frequency", C.c_double), ("bandwidth", C.c_double),
                ("dev", DevSpec * MAXDEV)]


class Scenario(C.Structure):
    _fields_ = [("nbands", C.c_int32), ("factor", C.c_int32), ("mode", C.c_int32),
                ("band", BandSpec * MAXBAND)]


MASK_FN = C.CFUNCTYPE(C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_uint32,
                      C.c_int, C.c_int64, C.c_int64, C.c_double)

_lib = None


def build(force=False):
    """Compile the restatement (``make -C oracle``)."""
    if force or not os.path.exists(LIB_PATH) or (
            os.path.getmtime(LIB_PATH) < max(os.path.getmtime(os.path.join(HERE, f))
                                             for f in ("gw_oracle.c", "gw_oracle.h", "Makefile"))):
        subprocess.check_call(["make", "-C", HERE, "-s"] + (["-B"] if force else []))
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.gwo_create.restype = C.c_void_p
        L.gwo_create.argtypes = [C.POINTER(Scenario)]
        L.gwo_destroy.argtypes = [C.c_void_p]
        L.gwo_default_scenario.argtypes = [C.POINTER(Scenario)]
        L.gwo_set_trace.argtypes = [C.c_void_p, C.c_int]
        L.gwo_set_mask_fn.argtypes = [C.c_void_p, MASK_FN, C.c_void_p, C.c_int64]
        L.gwo_reset.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.gwo_step.restype = C.c_int
        L.gwo_step.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                               C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_uint8)]
        L.gwo_run_for.restype = C.c_int
        L.gwo_run_for.argtypes = [C.c_void_p, C.c_double]
        L.gwo_add_mover.restype = C.c_int
        L.gwo_add_mover.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_double), C.c_int]
        L.gwo_set_position.restype = C.c_int
        L.gwo_set_position.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double]
        L.gwo_now.restype = C.c_double
        L.gwo_now.argtypes = [C.c_void_p]
        L.gwo_popped.restype = C.c_int64
        L.gwo_popped.argtypes = [C.c_void_p]
        L.gwo_fault.restype = C.c_int
        L.gwo_fault.argtypes = [C.c_void_p]
        L.gwo_near_ties.restype = C.c_int64
        L.gwo_near_ties.argtypes = [C.c_void_p]
        L.gwo_counts.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.gwo_received.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
        L.gwo_attenuation.restype = C.c_double
        L.gwo_attenuation.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.gwo_trace_take.restype = C.c_size_t
        L.gwo_trace_take.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_size_t]
        L.gwo_trace_size.restype = C.c_size_t
        L.gwo_trace_size.argtypes = [C.c_void_p]
        L.gwo_run_batch.restype = C.c_int
        L.gwo_run_batch.argtypes = [C.POINTER(Scenario), C.c_int64, C.c_int, C.c_int,
                                    C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int64, C.c_int64]
        L.gwo_use_philox_masks.argtypes = [C.c_void_p, C.c_uint64, C.c_int64]
        L.gwo_use_fed_masks.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64]
        L.gwo_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.gwo_run_batch_timed.restype = C.c_int
        L.gwo_run_batch_timed.argtypes = [C.POINTER(Scenario), C.c_int64, C.c_int, C.c_int] + [C.c_void_p] * 8 + \
            [C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_double)]
        L.gwo_run_batch_m.restype = C.c_int
        L.gwo_run_batch_m.argtypes = [C.POINTER(Scenario), C.c_int64, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int64, C.c_int64, C.c_uint64, C.c_int64, C.c_void_p, C.c_int, C.c_int]
        for name, args in (("gwo_q_function", [C.c_double]),
                           ("gwo_ber_bpsk", [C.c_double] * 3),
                           ("gwo_fspl", [C.c_double] * 5),
                           ("gwo_thermal_noise_mw", [C.c_double]),
                           ("gwo_max_correctable_ber", [C.c_int, C.c_int])):
            f = getattr(L, name)
            f.restype = C.c_double
            f.argtypes = args
        _lib = L
    return _lib


def scenario_from_dict(d, mode=MODE_R):
    """Convert a scenario dict (``oracle/ref_harness.default_scenario`` format)."""
    sc = Scenario()
    sc.nbands = len(d["bands"])
    sc.factor = int(d.get("assignment_duration_factor", 1000))
    sc.mode = mode
    for b, bd in enumerate(d["bands"]):
        bs = sc.band[b]
        bs.ndev = len(bd["devices"])
        bs.frequency = float(bd.get("frequency", 2.4e9))
        bs.bandwidth = float(bd.get("bandwidth", 22e6))
        order = [ROLE[x["role"]] for x in bd["devices"]]
        assert order == sorted(order), "canonical device order is senders, rrm, jammers"
        assert order.count(2) <= 1            # grids of PHY-only senders have no RRM (they are run with run_for)
        for i, x in enumerate(bd["devices"]):
            ds = bs.dev[i]
            ds.role = ROLE[x["role"]]
            ds.x, ds.y = float(x["x"]), float(x["y"])
            if x["role"] == "sender":
                ds.mult = int(x["mult"])
                p = x.get("payload", "counter")
                ds.payload_rule = -1 if p == "counter" else int(p)
                ds.dest = int(x["dest"])
                ds.interval = float(x.get("interval", 0.001))
                ds.max_ticks = int(x.get("max_ticks", 0))
                ds.receive = 1 if x.get("receive") else 0
            elif x["role"] == "jammer":
                ds.jam_interval = float(x["interval"])
                ds.jam_delay = float(x["delay"])
                ds.jam_power = float(x.get("power", 0.0))
                ds.jam_hdr = int(x.get("hdr", 13))
                ds.jam_payload = int(x["payload"])
    return sc


def default_scenario():
    sc = Scenario()
    lib().gwo_default_scenario(C.byref(sc))
    return sc


class OracleFault(RuntimeError):
    pass


class Oracle:
    """One env of the restatement (gym-like)."""

    def __init__(self, scenario=None, trace=False, mode=MODE_R):
        self.L = lib()
        if scenario is None:
            scenario = default_scenario()
        elif isinstance(scenario, dict):
            scenario = scenario_from_dict(scenario, mode)
        scenario.mode = mode
        self.sc = scenario
        self.nb = scenario.nbands
        self.h = self.L.gwo_create(C.byref(scenario))
        if not self.h:
            raise ValueError("bad scenario")
        self.L.gwo_set_trace(self.h, 1 if trace else 0)
        self._mask_cb = None

    def __del__(self):
        try:
            if self.h:
                self.L.gwo_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_mask_fn(self, fn, env_id=0):
        """fn(env, band, sender, seq, receiver, k0, k1, ber) -> int (mode M)."""
        def cb(ctx, env, band, sender, seq, receiver, k0, k1, ber):
            return int(fn(env, band, sender, seq, receiver, k0, k1, ber))
        self._mask_cb = MASK_FN(cb)
        self.L.gwo_set_mask_fn(self.h, self._mask_cb, None, env_id)

    def use_philox_masks(self, seed, env_id=0):
        """Mode M with the built-in Philox4x32-10 mask provider (same keying as the CUDA kernel)."""
        self.L.gwo_use_philox_masks(self.h, int(seed), int(env_id))

    def use_fed_masks(self, words, slots, env_index=0):
        """Mode M with fed masks: uint32 array [nenv][nbands][4][slots][4][words_per_row]."""
        self._fed = np.ascontiguousarray(words, dtype=np.uint32)
        self.L.gwo_use_fed_masks(self.h, self._fed.ctypes.data_as(C.c_void_p), int(slots),
                                 int(self._fed.shape[-1]), int(env_index))

    def reset(self):
        o = (C.c_int64 * MAXBAND)()
        self.L.gwo_reset(self.h, o)
        return [int(o[i]) for i in range(self.nb)] if self.nb > 1 else int(o[0])

    def step(self, action):
        acts = action if isinstance(action, (list, tuple)) else [action]
        dev = (C.c_int32 * MAXBAND)(*[int(a["device"]) for a in acts])
        dur = (C.c_int32 * MAXBAND)(*[int(a["duration"]) for a in acts])
        o = (C.c_int64 * MAXBAND)()
        r = (C.c_double * MAXBAND)()
        d = (C.c_uint8 * MAXBAND)()
        rc = self.L.gwo_step(self.h, dev, dur, o, r, d)
        if rc:
            raise OracleFault(FAULTS.get(rc, str(rc)))
        if self.nb > 1:
            return [(int(o[i]), float(r[i]), bool(d[i])) for i in range(self.nb)]
        return int(o[0]), float(r[0]), bool(d[0])

    def run_for(self, duration):
        """``SimMan.runSimulation(duration)``: advances simulated time by ``duration`` seconds."""
        rc = self.L.gwo_run_for(self.h, float(duration))
        if rc != 0:
            raise OracleFault("gwo_run_for failed (%d)" % rc)

    def add_mover(self, band, dev, first_delay, interval, offsets):
        """Mobility process of ``tests/test_benchmark.py:73-85``; ``offsets`` float64 ``[k, 2]`` (accumulating)."""
        arr = np.ascontiguousarray(offsets, dtype=np.float64).reshape(-1, 2)
        self._keep = getattr(self, "_keep", []) + [arr]
        rc = self.L.gwo_add_mover(self.h, band, dev, float(first_delay), float(interval),
                                  arr.ctypes.data_as(C.POINTER(C.c_double)), int(arr.shape[0]))
        if rc != 0:
            raise OracleFault("gwo_add_mover failed")

    def set_position(self, band, dev, x, y):
        """``device.position.set(x, y)`` between steps; transmissions that are on the air see the
        reference's ``SimplePhy._onAttenuationChange``."""
        if self.L.gwo_set_position(self.h, band, dev, float(x), float(y)) != 0:
            raise OracleFault("the position change hit a condition under which the reference raises")

    @property
    def now(self):
        return float(self.L.gwo_now(self.h))

    @property
    def popped(self):
        return int(self.L.gwo_popped(self.h))

    @property
    def near_ties(self):
        return int(self.L.gwo_near_ties(self.h))

    def counts(self, band=0):
        n_tx = C.c_int64()
        nd = (C.c_int64 * MAXDEV)()
        self.L.gwo_counts(self.h, band, C.byref(n_tx), nd)
        return int(n_tx.value), [int(x) for x in nd]

    def received(self, band=0):
        """Packets handed to ``onReceive`` per device (MAC receive mode)."""
        nr = (C.c_int64 * MAXDEV)()
        self.L.gwo_received(self.h, band, nr)
        return [int(x) for x in nr]

    def attenuation(self, band, i, j):
        return float(self.L.gwo_attenuation(self.h, band, i, j))

    def take_records(self):
        """Trace records in the tuple format of ``ref_harness.Tracer``."""
        n = self.L.gwo_trace_size(self.h)
        buf = np.empty(max(n, 1), dtype=np.float64)
        self.L.gwo_trace_take(self.h, buf.ctypes.data_as(C.POINTER(C.c_double)), n)
        out = []
        for r in buf[:n].reshape(-1, 8):
            k = int(r[0])
            if k == REC_TX:
                out.append(("tx", float(r[1]), int(r[2]), int(r[3]), float(r[4]), float(r[5]), float(r[6])))
            elif k == REC_BER:
                out.append(("ber", float(r[1]), int(r[2]), int(r[3]), float(r[4])))
            elif k == REC_DEC:
                out.append(("dec", float(r[1]), int(r[2]), int(r[3]), int(r[4]), float(r[5]),
                            float(r[6]), bool(r[7])))
            elif k == REC_RX:
                out.append(("rx", float(r[1]), int(r[2]), int(r[3])))
            elif k == 5:
                out.append(("mrx", float(r[1]), int(r[2]), int(r[3])))
        return out


def run_tape(oracle, actions, do_reset=True, moves=None):
    """Same output structure as ``ref_harness.run_tape`` (without ``events``)."""
    out = {"reset_obs": None, "steps": []}
    if do_reset:
        out["reset_obs"] = oracle.reset()
    oracle.take_records()
    for t, a in enumerate(actions):
        for (band, dev, x, y) in (moves or {}).get(t, []):
            oracle.set_position(band, dev, x, y)
        p0 = oracle.popped
        fb = oracle.step(a)
        if isinstance(fb, list):
            obs, rew, done = [f[0] for f in fb], [f[1] for f in fb], [f[2] for f in fb]
        else:
            obs, rew, done = fb
        out["steps"].append({"action": a, "obs": obs, "reward": rew, "done": done,
                             "now": oracle.now, "events": oracle.popped - p0,
                             "records": oracle.take_records()})
    return out


def run_batch(scenario, dev_tape, dur_tape, pos=None, do_reset=True, threads=None,
              want=("obs", "reward", "done", "now", "counts"), mode=MODE_R, seed=0, env_id_offset=0,
              fed_words=None, fed_slots=0, time_from=None):
    """
    Run ``nenv`` independent envs for ``nsteps`` steps with ``threads`` host threads.
    ``dev_tape`` / ``dur_tape``: int32 ``[nsteps, nenv, nbands]`` (or ``[nsteps, nenv]``).
    ``pos``: optional float64 ``[nenv, nbands, BATCH_DEV, 2]``.
    Returns a dict of numpy arrays.
    """
    L = lib()
    if isinstance(scenario, dict):
        scenario = scenario_from_dict(scenario, mode)
    scenario.mode = mode
    if fed_words is not None:
        fed_words = np.ascontiguousarray(fed_words, dtype=np.uint32)
    nb = scenario.nbands
    dev_tape = np.ascontiguousarray(dev_tape, dtype=np.int32)
    dur_tape = np.ascontiguousarray(dur_tape, dtype=np.int32)
    if dev_tape.ndim == 2:
        dev_tape = dev_tape[:, :, None]
        dur_tape = dur_tape[:, :, None]
    nsteps, nenv, nb2 = dev_tape.shape
    assert nb2 == nb
    if pos is not None:
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        assert pos.shape == (nenv, nb, BATCH_DEV, 2)
    res = {}
    if "obs" in want:
        res["obs"] = np.zeros((nsteps, nenv, nb), dtype=np.int64)
    if "reward" in want:
        res["reward"] = np.zeros((nsteps, nenv, nb), dtype=np.float64)
    if "done" in want:
        res["done"] = np.zeros((nsteps, nenv, nb), dtype=np.uint8)
    if "now" in want:
        res["now"] = np.zeros((nsteps, nenv), dtype=np.float64)
    if "counts" in want:
        res["counts"] = np.zeros((nenv, nb, 1 + BATCH_DEV), dtype=np.int64)

    def ptr(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    threads = threads or os.cpu_count() or 1
    threads = max(1, min(threads, nenv))
    bounds = np.linspace(0, nenv, threads + 1).astype(np.int64)
    rcs = [0] * threads

    secs = [0.0] * threads

    def work(i):
        if time_from is not None:
            # mode R only: the seconds each thread spends in steps >= time_from (burn-in excluded per env)
            out = C.c_double(0.0)
            rcs[i] = L.gwo_run_batch_timed(C.byref(scenario), nenv, nsteps, 1 if do_reset else 0,
                                           ptr(pos), ptr(dev_tape), ptr(dur_tape),
                                           ptr(res.get("obs")), ptr(res.get("reward")), ptr(res.get("done")),
                                           ptr(res.get("now")), ptr(res.get("counts")),
                                           int(bounds[i]), int(bounds[i + 1]), int(time_from), C.byref(out))
            secs[i] = out.value
            return
        rcs[i] = L.gwo_run_batch_m(C.byref(scenario), nenv, nsteps, 1 if do_reset else 0,
                                   ptr(pos), ptr(dev_tape), ptr(dur_tape),
                                   ptr(res.get("obs")), ptr(res.get("reward")), ptr(res.get("done")),
                                   ptr(res.get("now")), ptr(res.get("counts")),
                                   int(bounds[i]), int(bounds[i + 1]),
                                   int(seed), int(env_id_offset), ptr(fed_words), int(fed_slots),
                                   0 if fed_words is None else int(fed_words.shape[-1]))

    if threads == 1:
        work(0)
    else:
        ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    for rc in rcs:
        if rc:
            raise OracleFault(FAULTS.get(rc, str(rc)))
    res["threads"] = threads
    if time_from is not None:
        res["seconds"] = max(secs)         # the threads run side by side: the slowest one bounds the batch
    return res
