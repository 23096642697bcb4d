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
