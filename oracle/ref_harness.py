"""
ORACLE TEST INFRASTRUCTURE -- not product code, never imported by gymwipe_b200.

Drives the UNMODIFIED reference (``/root/reference/gymwipe``) on the dependency
shims of ``oracle/shims`` and records what the hot path did, so that

* the C restatement (``oracle/gw_oracle.c``) can be pinned against the reference
  itself (``oracle/check_restatement.py``), and
* golden input/output vectors can be committed under ``tests/golden``
  (``oracle/gen_golden.py``) -- the reference cannot travel to the GPU box.

Nothing of the reference is modified: tracing is done by wrapping methods at
run time (``FrequencyBand.transmit``, ``SimplePhy._decide``,
``SimplePhy._updateBitErrorRate`` and the interpreter's ``onPacketReceived``).

Two ways to build an env:

``make_default_env()``
    ``gym.make('CounterTraffic-v0')`` -- the reference's own class
    (``gymwipe/envs/counter_traffic.py:20``).

``ScenarioEnv(scenario)``
    the same building blocks (``SimpleNetworkDevice``, ``SimpleRrmDevice``,
    ``SimplePhy``, ``FrequencyBand``; construction order as in
    ``counter_traffic.py:114-133``) composed from a scenario description:
    arbitrary positions, traffic multiplicities, fixed payload sizes, PHY-only
    periodic senders ("jammers", modelled on ``tests/test_benchmark.py:20-50``)
    and several independent frequency bands (config 4).  With the default
    scenario it reproduces ``CounterTrafficEnv`` event for event
    (``oracle/check_restatement.py --selfcheck``).

Only usable where ``/root/reference`` exists (the build container).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GYMWIPE_REFERENCE", "/root/reference")


def setup_paths():
    """Put the shims and the reference on ``sys.path`` (idempotent)."""
    if not os.path.isdir(os.path.join(REF, "gymwipe")):
        raise RuntimeError("reference not available at %s" % REF)
    for p in (REF, os.path.join(HERE, "shims")):
        if p not in sys.path:
            sys.path.insert(0, p)


# --------------------------------------------------------------------------
# default scenario == CounterTrafficEnv (counter_traffic.py:124-133)
# --------------------------------------------------------------------------

def default_scenario():
    return {
        "assignment_duration_factor": 1000,
        "bands": [{
            "frequency": 2.4e9,
            "bandwidth": 22e6,
            "devices": [
                {"role": "sender", "x": 0.0, "y": 2.0, "mult": 1, "payload": "counter",
                 "interval": 0.001, "dest": 1},
                {"role": "sender", "x": 0.0, "y": -2.0, "mult": 3, "payload": "counter",
                 "interval": 0.001, "dest": 0},
                {"role": "rrm", "x": 0.0, "y": 0.0},
            ],
        }],
    }


class Tracer:
    """Collects trace records; installed by wrapping reference methods."""

    def __init__(self):
        self.records = []
        self.phy_index = {}      # SimplePhy -> (band, device index)
        self.band_index = {}     # FrequencyBand -> band
        self.device_index = {}   # Device -> (band, device index)
        self.enabled = True
        self.record_ber = True

    def install(self):
        from gymwipe.networking.physical import FrequencyBand
        from gymwipe.networking.simple_stack import SimplePhy
        from gymwipe.simtools import SimMan
        tracer = self

        if getattr(FrequencyBand, "_oracle_traced", False):
            FrequencyBand._oracle_tracer[0] = tracer
            return
        holder = [tracer]
        FrequencyBand._oracle_tracer = holder
        FrequencyBand._oracle_traced = True

        orig_transmit = FrequencyBand.transmit

        def transmit(self, sender, power, packet, mcsHeader, mcsPayload):
            t = orig_transmit(self, sender, power, packet, mcsHeader, mcsPayload)
            tr = holder[0]
            if tr is not None:
                seqs = tr.__dict__.setdefault("tx_seq", {})
                t._oracle_seq = seqs.get(sender, 0)
                seqs[sender] = t._oracle_seq + 1
            if tr is not None and tr.enabled and sender in tr.device_index:
                band, dev = tr.device_index[sender]
                tr.records.append(("tx", t.startTime, band, dev, t.stopTime,
                                   t.headerBits, t.payloadBits))
            return t
        FrequencyBand.transmit = transmit

        orig_decide = SimplePhy._decide

        def _decide(self, bitErrorSum, totalBits, mcs, logSubject="Data"):
            ok = orig_decide(self, bitErrorSum, totalBits, mcs, logSubject)
            tr = holder[0]
            if tr is not None and tr.enabled and self in tr.phy_index:
                band, dev = tr.phy_index[self]
                tr.records.append(("dec", SimMan.now, band, dev,
                                   0 if logSubject == "Header" else 1,
                                   float(bitErrorSum), float(totalBits), bool(ok)))
            return ok
        SimplePhy._decide = _decide

        orig_update = SimplePhy._updateBitErrorRate

        def _updateBitErrorRate(self, t):
            orig_update(self, t)
            tr = holder[0]
            if tr is not None and tr.enabled and tr.record_ber and self in tr.phy_index:
                band, dev = tr.phy_index[self]
                tr.records.append(("ber", SimMan.now, band, dev,
                                   float(self._receivedBitErrorRate)))
        SimplePhy._updateBitErrorRate = _updateBitErrorRate

    def wrap_interpreter(self, interpreter, band):
        from gymwipe.simtools import SimMan
        orig = interpreter.onPacketReceived
        tracer = self

        def onPacketReceived(senderIndex, receiverIndex, payload):
            if tracer.enabled:
                tracer.records.append(("rx", SimMan.now, band, senderIndex))
            return orig(senderIndex, receiverIndex, payload)
        interpreter.onPacketReceived = onPacketReceived

    def take(self):
        r, self.records = self.records, []
        return r


# --------------------------------------------------------------------------
# mode M on the reference side: a SimplePhy subclass that replaces ONLY the error
# bookkeeping (_countBitErrors / _resetBitErrorCounter) by per-bit error masks;
# MAC / RRM / timing / power bookkeeping stay the reference's
# --------------------------------------------------------------------------

def philox4x32_10(ctr, key):
    """Vectorised Philox4x32-10 (Random123): ctr uint32 [n,4], key uint32 [2] -> uint32 [n,4]."""
    import numpy as np
    c = [ctr[:, i].astype(np.uint64) for i in range(4)]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        n0 = ((p1 >> np.uint64(32)) ^ c[1] ^ k0) & mask
        n1 = p1 & mask
        n2 = ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & mask
        n3 = p0 & mask
        c = [n0, n1, n2, n3]
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return np.stack(c, axis=1).astype(np.uint32)


def philox_mask_errors(seed, env, band, sender, seq, receiver, k0, k1, ber):
    """Number of error flags among on-air bits [k0, k1) (the project's mode-M keying)."""
    import numpy as np
    if k1 <= k0:
        return 0
    thr = int(ber * 4294967296.0)
    blocks = np.arange(k0 >> 2, ((k1 - 1) >> 2) + 1, dtype=np.uint64)
    ctr = np.zeros((len(blocks), 4), np.uint32)
    ctr[:, 0] = blocks.astype(np.uint32)
    ctr[:, 1] = seq
    ctr[:, 2] = sender | (receiver << 8) | (band << 16)
    ctr[:, 3] = env & 0xFFFFFFFF
    key = [(seed & 0xFFFFFFFF) ^ ((env >> 32) & 0xFFFFFFFF), (seed >> 32) & 0xFFFFFFFF]
    w = philox4x32_10(ctr, key).reshape(-1)
    ks = np.arange((k0 >> 2) * 4, (((k1 - 1) >> 2) + 1) * 4)
    sel = (ks >= k0) & (ks < k1)
    return int(np.count_nonzero(w[sel] < thr))


def install_masked_phy(mask_fn, tracer):
    """
    Makes the reference's devices construct ``MaskedPhy`` instead of ``SimplePhy`` (run-time
    rebinding of the name in ``gymwipe.networking.devices``; no source is modified).
    ``mask_fn(band, sender, seq, receiver, k0, k1, ber) -> int``.
    """
    from math import floor
    import gymwipe.networking.devices as devices_mod
    from gymwipe.networking.simple_stack import SimplePhy
    from gymwipe.simtools import SimMan

    class MaskedPhy(SimplePhy):
        def _resetBitErrorCounter(self):
            super()._resetBitErrorCounter()
            self._segT0 = SimMan.now
            self._errInt = 0

        def _receive(self, t):
            if not self._transmitting:
                self._rxT = t
            yield from super()._receive(t)

        def _countBitErrors(self):
            # errors of the on-air bits of the segment that ends now (counted since the last
            # CHANGE, integer): replaces the expected-value accounting of simple_stack.py:180-188
            t = self._rxT
            now = SimMan.now
            rate = self._currentReceiverMcs.bitRate
            k0 = int(floor((self._segT0 - t.startTime) * rate))
            k1 = int(floor((now - t.startTime) * rate))
            if k1 > k0:
                band, sender = tracer.device_index[t.sender]
                _, receiver = tracer.phy_index[self]
                self._errInt += mask_fn(band, sender, t._oracle_seq, receiver, k0, k1,
                                        self._receivedBitErrorRate)
            self._segT0 = now
            self._receivedBitErrorSum = float(self._errInt)

    devices_mod.SimplePhy = MaskedPhy
    import sys
    me = sys.modules[__name__]
    me._MaskedPhy = MaskedPhy
    return MaskedPhy


def _reset_mac_counter():
    from gymwipe.networking.simple_stack import SimpleMac
    # one env per process is the reference's rule; the class-level address
    # counter (simple_stack.py:374-384) overflows after 255 addresses otherwise
    SimpleMac._macCounter = 0


def make_default_env(tracer=None):
    """The reference's own ``CounterTrafficEnv`` via ``gym.make``."""
    setup_paths()
    import gym
    import gymwipe.envs  # noqa: F401  (registers the ids)
    _reset_mac_counter()
    env = gym.make('CounterTraffic-v0')
    if tracer is not None:
        tracer.install()
        tracer.band_index[env.frequencyBand] = 0
        for i, s in enumerate(env.senders):
            tracer.phy_index[s._phy] = (0, i)
            tracer.device_index[s] = (0, i)
        tracer.phy_index[env.rrm._phy] = (0, len(env.senders))
        tracer.device_index[env.rrm] = (0, len(env.senders))
        tracer.wrap_interpreter(env.rrm.interpreter, 0)
    return env


class ScenarioEnv:
    """
    A gym-style env composed from the reference's classes according to a
    scenario dict (see :func:`default_scenario`).  ``step`` takes one
    ``{"device", "duration"}`` dict per band (or a single dict for one band) and
    runs until every band's ASSIGN message is processed.
    """

    COUNTER_BOUND = 65536
    COUNTER_BYTE_LENGTH = 2

    def __init__(self, scenario, tracer=None, movers=None):
        """``movers``: optional ``{device index: (first delay, interval, offsets [k][2])}`` for band 0 -- mobility
        processes after ``tests/test_benchmark.py:73-85`` (the draws come from the tape), started after the whole
        scenario has been constructed, in device order."""
        setup_paths()
        from gymwipe.envs.counter_traffic import CounterTrafficEnv
        from gymwipe.networking.attenuation_models import FsplAttenuation
        from gymwipe.networking.devices import (NetworkDevice, SimpleNetworkDevice,
                                                SimpleRrmDevice)
        from gymwipe.networking.messages import (Message, Packet, SimpleMacHeader,
                                                 SimpleNetworkHeader, StackMessageTypes,
                                                 Transmittable)
        from gymwipe.networking.physical import BpskMcs, FrequencyBand
        from gymwipe.networking.simple_stack import SimplePhy
        from gymwipe.simtools import SimMan

        self.SimMan = SimMan
        self.scenario = scenario
        self.factor = int(scenario.get("assignment_duration_factor", 1000))
        self.tracer = tracer
        if tracer is not None:
            tracer.install()
        _reset_mac_counter()
        SimMan.init()

        outer = self

        class Sender(SimpleNetworkDevice):
            # mirrors CounterTrafficEnv.SenderDevice (counter_traffic.py:37-61) with
            # the interval / payload rule as parameters
            def __init__(self, name, x, y, band, mult, payload, interval, max_ticks=0):
                super().__init__(name, x, y, band)
                self.packetMultiplicity = mult
                self.payloadRule = payload
                self.interval = interval
                self.maxTicks = max_ticks       # 0: forever (the reference's SenderDevice); n: a burst of n ticks
                self.counter = 1
                SimMan.process(self.senderProcess())
                self.destinationMac = None

            def onReceive(self, packet):        # devices.py:99-111: invoked by the receiver loop in receive mode
                if outer.tracer is not None and outer.tracer.enabled:
                    band, dev = outer.tracer.device_index[self]
                    outer.tracer.records.append(("mrx", SimMan.now, band, dev))

            def senderProcess(self):
                assert self.destinationMac is not None
                ticks = 0
                while self.maxTicks == 0 or ticks < self.maxTicks:
                    ticks += 1
                    for _ in range(self.packetMultiplicity):
                        if self.payloadRule == "counter":
                            data = Transmittable(outer.COUNTER_BYTE_LENGTH, self.counter)
                        else:
                            data = Transmittable(outer.COUNTER_BYTE_LENGTH, int(self.payloadRule))
                        self.send(data, self.destinationMac)
                    if self.counter < outer.COUNTER_BOUND:
                        self.counter += 1
                    yield SimMan.timeout(self.interval)

        class Jammer(NetworkDevice):
            # PHY-only periodic sender, after tests/test_benchmark.py:20-50.  Its packets
            # carry a MAC header and a network packet addressed to itself so that an RRM
            # that decodes one can map both addresses (devices.py:165-166 would raise
            # KeyError / AttributeError otherwise, SURVEY.md section 8d cfg 4 "hazard").
            def __init__(self, name, x, y, band, interval, delay, power, hdr, payload, mac):
                super().__init__(name, x, y, band)
                self.macAddr = mac
                import gymwipe.networking.devices as devices_mod
                self._phy = devices_mod.SimplePhy("phy", self, band)    # MaskedPhy in mode M
                mcs = BpskMcs(band.spec)
                assert payload >= 12

                def sender():
                    yield SimMan.timeout(delay)
                    while True:
                        yield SimMan.timeout(interval)
                        header = SimpleMacHeader(mac, mac, flag=0)
                        header.byteSize = hdr
                        packet = Packet(header, Packet(SimpleNetworkHeader(mac, mac),
                                                       Transmittable(None, payload - 12)))
                        signal = Message(StackMessageTypes.SEND,
                                         {"packet": packet, "power": power, "mcs": mcs})
                        self._phy.gates["macIn"].send(signal)
                SimMan.process(sender())

        self.bands = []
        for b, bspec in enumerate(scenario["bands"]):
            band = FrequencyBand([FsplAttenuation], bspec.get("frequency", 2.4e9),
                                 bspec.get("bandwidth", 22e6))
            senders, jammers, rrm_spec, dest = [], [], None, []
            devs = bspec["devices"]
            dev_objs = [None] * len(devs)
            for i, d in enumerate(devs):
                if d["role"] == "sender":
                    s = Sender("Sender %d.%d" % (b, i), d["x"], d["y"], band, d["mult"],
                               d.get("payload", "counter"), d.get("interval", 0.001), int(d.get("max_ticks", 0)))
                    senders.append((i, s))
                    dest.append(d.get("dest"))
                    dev_objs[i] = s
                elif d["role"] == "rrm":
                    assert rrm_spec is None
                    rrm_spec = (i, d)
                elif d["role"] == "jammer":
                    pass
                else:
                    raise ValueError(d["role"])
            # senders are wired before the RRM exists (counter_traffic.py:128-133)
            sender_objs = [s for _, s in senders]
            idx2mac = {k: s.macAddr for k, s in enumerate(sender_objs)}
            dev_to_sender = {i: k for k, (i, _) in enumerate(senders)}
            for k, s in enumerate(sender_objs):
                s.destinationMac = sender_objs[dev_to_sender[dest[k]]].macAddr
            # the RRM knows every device of its band (index = device index); the jammers'
            # addresses are fixed up-front because the RRM is constructed before them
            jam_mac = {i: bytes([0, 0, 0, 0, 1, i]) for i, d in enumerate(devs) if d["role"] == "jammer"}
            for i, m in jam_mac.items():
                idx2mac[i] = m
            interp = CounterTrafficEnv.CounterTrafficInterpreter(_InterpEnvView([None] * len(devs), self))
            ri, rd = rrm_spec
            assert ri == len(sender_objs), "canonical device order is senders, rrm, jammers"
            rrm = SimpleRrmDevice("RRM %d" % b, rd["x"], rd["y"], band, idx2mac, interp)
            dev_objs[ri] = rrm
            for i, d in enumerate(devs):
                if d["role"] == "jammer":
                    assert i > ri
                    j = Jammer("Jammer %d.%d" % (b, i), d["x"], d["y"], band, d["interval"],
                               d["delay"], d.get("power", 0.0), d.get("hdr", 13), d["payload"],
                               jam_mac[i])
                    jammers.append((i, j))
                    dev_objs[i] = j
            if tracer is not None:
                tracer.band_index[band] = b
                for i, o in enumerate(dev_objs):
                    tracer.phy_index[o._phy] = (b, i)
                    tracer.device_index[o] = (b, i)
                tracer.wrap_interpreter(interp, b)
            self.bands.append({"band": band, "senders": sender_objs, "rrm": rrm,
                               "devices": dev_objs, "interp": interp})
        # MAC receive mode (devices.py:70-97): switched on after the whole scenario has been constructed, band by
        # band in device order -- each starts its blocking receive loop as a process
        for b, bspec in enumerate(scenario["bands"]):
            for i, d in enumerate(bspec["devices"]):
                if d["role"] == "sender" and d.get("receive"):
                    self.bands[b]["devices"][i].receiving = True

        def mover(dev, first, interval, offsets):       # tests/test_benchmark.py:75-82 with the draws from the tape
            yield SimMan.timeout(first)
            initialPos = dev.position
            for (xOffset, yOffset) in offsets:
                dev.position.set(initialPos.x + float(xOffset), initialPos.y + float(yOffset))
                yield SimMan.timeout(interval)
        for i in sorted(movers or {}):
            first, interval, offsets = movers[i]
            SimMan.process(mover(self.bands[0]["devices"][i], float(first), float(interval), offsets))

    # gym-like API --------------------------------------------------------
    def reset(self):
        obs = []
        for b in self.bands:
            for s in b["senders"]:
                s.counter = 0
            b["interp"].reset()
            obs.append(b["interp"].getObservation())
        return obs if len(obs) > 1 else obs[0]

    def step(self, action):
        actions = action if isinstance(action, (list, tuple)) else [action]
        assert len(actions) == len(self.bands)
        signals = []
        for b, a in zip(self.bands, actions):
            signals.append(b["rrm"].assignFrequencyBand(int(a["device"]),
                                                        int(a["duration"]) * self.factor))
        for s in signals:
            self.SimMan.runSimulation(s.eProcessed)
        fb = [b["interp"].getFeedback() for b in self.bands]
        return fb if len(fb) > 1 else fb[0]


class _InterpEnvView:
    """What ``CounterTrafficInterpreter`` reads from its env (``senders``, ``COUNTER_BOUND``)."""

    def __init__(self, senders, outer):
        self.senders = senders
        self.COUNTER_BOUND = outer.COUNTER_BOUND


# --------------------------------------------------------------------------
# running action tapes
# --------------------------------------------------------------------------

def run_tape(env, actions, tracer, do_reset=True, moves=None):
    """
    Replays ``actions`` (list of dicts, or list of per-band lists of dicts) and
    returns a trace dict: per-step ``obs/reward/done/now/events`` plus the
    records the tracer collected during that step.
    """
    from gymwipe.simtools import SimMan
    out = {"reset_obs": None, "steps": []}
    if do_reset:
        out["reset_obs"] = env.reset()
    tracer.take()
    for t, a in enumerate(actions):
        for (band, dev, x, y) in (moves or {}).get(t, []):
            # Position.set -> nChange -> FsplAttenuation._update (devices/core.py:75-84)
            env.bands[band]["devices"][dev].position.set(float(x), float(y))
        popped0 = SimMan.env.popped_events
        fb = env.step(a)
        if isinstance(fb, list):
            obs = [int(f[0]) for f in fb]
            rew = [float(f[1]) for f in fb]
            done = [bool(f[2]) for f in fb]
        else:
            obs, rew, done = int(fb[0]), float(fb[1]), bool(fb[2])
        out["steps"].append({
            "action": a,
            "obs": obs, "reward": rew, "done": done,
            "now": float(SimMan.now),
            "events": SimMan.env.popped_events - popped0,
            "records": tracer.take(),
        })
    return out


def random_actions(n, seed=0, devices=2, durations=20):
    """Action tape of cfg 1: ``RandomState(seed)``, ``device=randint(2)`` then
    ``duration=randint(20)`` alternately."""
    import numpy as np
    rs = np.random.RandomState(seed)
    tape = []
    for _ in range(n):
        d = int(rs.randint(devices))
        u = int(rs.randint(durations))
        tape.append({"device": d, "duration": u})
    return tape
