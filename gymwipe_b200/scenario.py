"""
Scenario compiler: device / band descriptors (``gymwipe_b200.networking``) -> the plain
``gw_config`` struct of the C ABI.  Canonical device order on a band is senders, RRM, jammers
(it is the construction order of ``CounterTrafficEnv.__init__``, ``counter_traffic.py:124-133``,
which fixes the SimPy event ids the kernel's tie-breaking reproduces).
"""
from gymwipe_b200 import _native as N

MODES = {"reference": N.GW_MODE_REFERENCE, "R": N.GW_MODE_REFERENCE,
         "mask_philox": N.GW_MODE_MASK_PHILOX, "M": N.GW_MODE_MASK_PHILOX,
         "mask_fed": N.GW_MODE_MASK_FED}


def default_scenario_dict():
    """``CounterTrafficEnv``'s devices (``counter_traffic.py:124-133``) in dict form."""
    return {
        "assignment_duration_factor": 1000,
        "bands": [{
            "frequency": 2.4e9, "bandwidth": 22e6,
            "devices": [
                {"role": "sender", "x": 0.0, "y": 2.0, "mult": 1, "payload": "counter", "interval": 0.001, "dest": 1},
                {"role": "sender", "x": 0.0, "y": -2.0, "mult": 3, "payload": "counter", "interval": 0.001, "dest": 0},
                {"role": "rrm", "x": 0.0, "y": 0.0},
            ],
        }],
    }


def config_from_dict(d, n_envs, mode="reference", seed=0, env_id_offset=0, per_env_positions=False,
                     max_assign_duration=20):
    """Scenario dict (same format as the oracle harness uses) -> ``Config``."""
    cfg = N.Config()
    cfg.abi_version = N.GW_ABI_VERSION
    cfg.n_envs = int(n_envs)
    cfg.env_id_offset = int(env_id_offset)
    cfg.n_bands = len(d["bands"])
    cfg.assignment_duration_factor = int(d.get("assignment_duration_factor", 1000))
    cfg.max_assign_duration = int(max_assign_duration)
    cfg.mode = MODES[mode] if isinstance(mode, str) else int(mode)
    cfg.seed = int(seed)
    cfg.per_env_positions = 1 if per_env_positions else 0
    for b, bd in enumerate(d["bands"]):
        bc = cfg.band[b]
        devs = bd["devices"]
        if len(devs) > N.GW_MAX_DEVICES:
            raise ValueError("at most %d devices per band" % N.GW_MAX_DEVICES)
        bc.n_devices = len(devs)
        bc.frequency_hz = float(bd.get("frequency", 2.4e9))
        bc.bandwidth_hz = float(bd.get("bandwidth", 22e6))
        for i, x in enumerate(devs):
            dc = bc.device[i]
            dc.role = {"sender": N.GW_ROLE_SENDER, "rrm": N.GW_ROLE_RRM, "jammer": N.GW_ROLE_JAMMER}[x["role"]]
            dc.x, dc.y = float(x["x"]), float(x["y"])
            if x["role"] == "sender":
                dc.multiplicity = int(x["mult"])
                p = x.get("payload", "counter")
                dc.payload_bytes = -1 if p == "counter" else int(p)
                dc.interval = float(x.get("interval", 0.001))
                dc.max_ticks = int(x.get("max_ticks", 0))
                dc.receive = 1 if x.get("receive") else 0
                if "dest" in x and int(x["dest"]) != 1 - i:
                    raise ValueError("the two senders of a band address each other")
            elif x["role"] == "jammer":
                dc.jam_interval = float(x["interval"])
                dc.jam_delay = float(x["delay"])
                dc.jam_power_dbm = float(x.get("power", 0.0))
                dc.jam_header_bytes = int(x.get("hdr", 13))
                dc.jam_payload_bytes = int(x["payload"])
    return cfg


def dict_from_bands(bands, assignment_duration_factor=1000):
    """``FrequencyBand`` descriptors (with their registered devices) -> scenario dict."""
    out = {"assignment_duration_factor": assignment_duration_factor, "bands": []}
    for band in bands:
        order = {"sender": 0, "rrm": 1, "jammer": 2}
        devs = sorted(band.devices, key=lambda dv: order[dv._role])
        senders = [dv for dv in devs if dv._role == "sender"]
        entries = []
        for dv in devs:
            e = {"role": dv._role, "x": dv.position.x, "y": dv.position.y}
            if dv._role == "sender":
                e.update(mult=dv.packetMultiplicity, payload=getattr(dv, "payloadRule", "counter"),
                         interval=getattr(dv, "interval", 0.001), dest=1 - senders.index(dv))
            elif dv._role == "jammer":
                e.update(interval=dv.sendInterval, delay=dv.initialDelay, power=dv.power,
                         hdr=dv.headerBytes, payload=dv.payloadBytes)
            entries.append(e)
        out["bands"].append({"frequency": band.spec.frequency, "bandwidth": band.spec.bandwidth,
                             "devices": entries})
    return out


def compile_stack(bands, assignment_duration_factor=1000):
    """
    Traces wired network stacks into the scenario table.

    ``bands``: ``FrequencyBand`` descriptors whose registered devices own ``Module`` s wired with the
    reference's plumbing (``gymwipe_b200.networking.construction``).  The ROLE of a device is read from its
    wiring, not from its class: the device's ``SimplePhy`` on the band is located, its ``"mac"`` port is
    followed through ``Gate.connectTo`` connections (proxy ports in between are passed through, as in
    ``tests/networking/test_stack.py:134-158``) to the first ``SimpleMac`` (a sender with a MAC queue), or
    ``SimpleRrmMac`` (the band's RRM); a PHY whose ``"mac"`` port leads to no MAC is a PHY-only periodic
    sender (``tests/test_benchmark.py:20-50``).  Traffic parameters are the device's attributes
    (``packetMultiplicity`` / ``interval`` / ``payloadRule``; ``sendInterval`` / ``initialDelay`` / ``power`` /
    ``headerBytes`` / ``payloadBytes``; with more than two senders ``destination``: the addressed sender).  Bands of 2
    senders + RRM (+ 1 PHY-only sender) compile to the step kernels' template, larger ones (up to 8 senders and 16
    PHY-only senders) to the general band engine -- ``gymwipe_b200.make('CounterTraffic-v0', scenario=...)`` picks the
    engine.  Raises ``ValueError`` for stacks neither has a table for.
    """
    from gymwipe_b200.networking.construction import Module, Port
    from gymwipe_b200.networking.simple_stack import SimpleMac, SimplePhy, SimpleRrmMac

    def modules_of(device):
        found = []
        for value in vars(device).values():
            if isinstance(value, Module):
                found.append(value)
        return found

    def mac_behind(phy):
        """Breadth-first along the connections leaving the PHY's mac port."""
        seen, frontier = set(), [phy.ports["mac"].output]
        while frontier:
            gate = frontier.pop(0)
            if id(gate) in seen:
                continue
            seen.add(id(gate))
            port = gate._owner if isinstance(gate._owner, Port) else None
            module = port._owner if port is not None else gate._owner
            if isinstance(module, (SimpleMac, SimpleRrmMac)) and module is not phy:
                back = module.ports["phy"].output
                if not _reaches(back, phy.ports["mac"].input):
                    raise ValueError("%r: the MAC's phy port is not connected back to the PHY" % (module,))
                return module
            frontier.extend(gate.connections)
        return None

    out = {"assignment_duration_factor": assignment_duration_factor, "bands": []}
    for band in bands:
        senders, rrms, jammers = [], [], []
        for dv in band.devices:
            phys = [m for m in modules_of(dv) if isinstance(m, SimplePhy) and m.frequencyBand is band]
            if len(phys) != 1:
                raise ValueError("%r needs exactly one SimplePhy on the band (found %d)" % (dv, len(phys)))
            mac = mac_behind(phys[0])
            if isinstance(mac, SimpleRrmMac):
                rrms.append(dv)
            elif isinstance(mac, SimpleMac):
                senders.append(dv)
            else:
                jammers.append(dv)
        if len(rrms) != 1:
            raise ValueError("a band needs exactly one RRM stack (SimplePhy <-> SimpleRrmMac), found %d" % len(rrms))
        # 2 MAC senders + RRM + up to GW_MAX_JAMMERS PHY-only sender(s): the step kernels' template; up to
        # GW_GENBAND_MAX_SENDERS / GW_GENBAND_MAX_PHY_SENDERS: the general band engine (envs/general_band.py)
        if not 2 <= len(senders) <= N.GW_GENBAND_MAX_SENDERS or len(jammers) > N.GW_GENBAND_MAX_PHY_SENDERS:
            raise ValueError("a band holds 2..%d MAC senders + RRM + up to %d PHY-only senders; got %d / %d"
                             % (N.GW_GENBAND_MAX_SENDERS, N.GW_GENBAND_MAX_PHY_SENDERS, len(senders), len(jammers)))
        entries = []
        for i, dv in enumerate(senders):
            if len(senders) == 2:
                dest = 1 - i                        # counter_traffic.py:128-129: the two senders address each other
            else:
                target = getattr(dv, "destination", None)       # the addressed sender: a device object or its index
                if target is None:
                    raise ValueError("%r: with more than two senders every sender needs a `destination`" % (dv,))
                dest = target if isinstance(target, int) else next((k for k, o in enumerate(senders) if o is target), -1)
                if not 0 <= dest < len(senders) or dest == i:
                    raise ValueError("%r: `destination` must be another sender of the band" % (dv,))
            entries.append({"role": "sender", "x": dv.position.x, "y": dv.position.y,
                            "mult": getattr(dv, "packetMultiplicity", 1), "payload": getattr(dv, "payloadRule", "counter"),
                            "interval": getattr(dv, "interval", 0.001), "dest": dest})
        entries.append({"role": "rrm", "x": rrms[0].position.x, "y": rrms[0].position.y})
        for dv in jammers:
            for attr in ("sendInterval", "initialDelay", "payloadBytes"):
                if not hasattr(dv, attr):
                    raise ValueError("%r is a PHY-only sender: it needs the attribute %r" % (dv, attr))
            entries.append({"role": "jammer", "x": dv.position.x, "y": dv.position.y, "interval": dv.sendInterval,
                            "delay": dv.initialDelay, "power": getattr(dv, "power", 0.0),
                            "hdr": getattr(dv, "headerBytes", 13), "payload": dv.payloadBytes})
        out["bands"].append({"frequency": band.spec.frequency, "bandwidth": band.spec.bandwidth, "devices": entries})
    return out


def _reaches(gate, target, _seen=None):
    _seen = set() if _seen is None else _seen
    if gate is target:
        return True
    if id(gate) in _seen:
        return False
    _seen.add(id(gate))
    return any(_reaches(g, target, _seen) for g in gate.connections)
