"""
Bands beyond ``CounterTrafficEnv``'s two senders + RRM: ``GeneralBandEnv`` steps a
band of up to 8 MAC senders (``SimpleNetworkDevice`` + the counter traffic process of
``counter_traffic.py:37-61``, optionally in MAC receive mode or with a finite burst), one RRM
(``SimpleRrmDevice`` with ``CounterTrafficInterpreter``) and up to 16 PHY-only periodic senders
(``tests/test_benchmark.py:20-50``) with the env's own protocol -- ``assignFrequencyBand(device, duration)``, run
until the ASSIGN message is processed, ``Interpreter.getFeedback`` (``counter_traffic.py:146-158``) -- for
``num_envs`` independent envs on the GPU (CUDA engine ``gymwipe_b200/csrc/gw_band.cuh``, C ABI ``gw_genband_*``).

The gym surface is ``CounterTrafficEnv``'s: ``reset`` / ``step`` / ``action_space`` (``device`` is
``Discrete(n_senders)``) / ``observation_space``; the interpreter keeps one received value per device and its
observation stays ``receivedValues[0] - receivedValues[1] + COUNTER_BOUND`` (``counter_traffic.py:69-80, 96``).
``gymwipe_b200.make('CounterTraffic-v0', scenario=...)`` returns this class when the scenario does not fit the
step kernel's 2 + 1 (+ 1) template.
"""
import ctypes as C

import torch

from gymwipe_b200 import _native as N
from gymwipe_b200 import spaces
from gymwipe_b200.envs.core import BaseEnv


def fits_step_kernel_template(scenario):
    """True if ``CounterTrafficEnv``'s kernels have tables for every band of the scenario: two MAC senders that
    address each other, the RRM and at most ``GW_MAX_JAMMERS`` PHY-only sender(s)."""
    for bd in scenario["bands"]:
        roles = [d["role"] for d in bd["devices"]]
        if roles.count("sender") != 2 or roles.count("jammer") > N.GW_MAX_JAMMERS:
            return False
        for i, d in enumerate(bd["devices"][:2]):
            if d["role"] != "sender" or int(d.get("dest", 1 - i)) != 1 - i:
                return False
    return True


def _trace_tuples(rows):
    recs = []
    for r in rows:
        k = int(r[0])
        if k == 1:
            recs.append(("tx", float(r[1]), 0, int(r[2]), float(r[3]), float(r[4]), float(r[5])))
        elif k == 2:
            recs.append(("ber", float(r[1]), 0, int(r[2]), float(r[3])))
        elif k == 3:
            recs.append(("dec", float(r[1]), 0, int(r[2]), int(r[3]), float(r[4]), float(r[5]), bool(r[6])))
        elif k == 4:
            recs.append(("rx", float(r[1]), 0, int(r[2])))
        elif k == 5:
            recs.append(("mrx", float(r[1]), 0, int(r[2])))
    return recs


class _LazyInfo(dict):
    """``info`` of a batched step: ``"Latest received values"`` is read back from the device on first access
    (the reference returns ``str(receivedValues)``, ``counter_traffic.py:109-112``)."""

    def __init__(self, env):
        super().__init__()
        self._env = env

    def __missing__(self, key):
        if key == "Latest received values":
            v = self._env.received_values()
            self[key] = v
            return v
        raise KeyError(key)

    def keys(self):
        return ["Latest received values"]

    def __contains__(self, key):
        return key == "Latest received values" or super().__contains__(key)


class GeneralBandEnv(BaseEnv):
    """
    Args:
        num_envs: independent envs in this batch.
        scenario: one-band scenario dict (``gymwipe_b200.scenario`` format): ``devices`` = senders (``mult``,
            ``payload``, ``interval``, ``dest``, optional ``receive`` / ``max_ticks``), then the RRM, then PHY-only
            senders (``role: "jammer"``: ``interval``, ``delay``, ``power``, ``hdr``, ``payload``).
        positions: optional float64 tensor ``[num_envs, n_devices, 2]`` of per-env positions.
        strict: validate actions / faults after every step (costs a device sync); default: ``num_envs == 1``.
        mode: ``"reference"`` (mode R: the reference's expected-value error accounting) or ``"mask_philox"`` (mode M:
            per-bit Philox4x32-10 error masks keyed by ``seed`` and the global env id ``env_id_offset + env`` --
            results do not depend on batch size or sharding).
    """

    COUNTER_INTERVAL = 0.001
    COUNTER_BYTE_LENGTH = 2
    COUNTER_BOUND = 2 ** (8 * COUNTER_BYTE_LENGTH)

    def __init__(self, num_envs=1, device="cuda", scenario=None, positions=None, strict=None, mode="reference", seed=0,
                 env_id_offset=0):
        if not torch.cuda.is_available():
            raise RuntimeError("gymwipe_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if scenario is None or len(scenario["bands"]) != 1:
            raise ValueError("GeneralBandEnv takes a scenario with exactly one band")
        if mode not in ("reference", "R", "mask_philox", "M"):
            raise ValueError("the general band engine offers the modes 'reference' and 'mask_philox'")
        self.mode = "reference" if mode in ("reference", "R") else "mask_philox"
        dev = torch.device(device)
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self.num_envs = int(num_envs)
        self._scalar_api = self.num_envs == 1
        self.strict = self._scalar_api if strict is None else bool(strict)
        self.scenario = scenario
        band = scenario["bands"][0]
        devs = band["devices"]
        roles = [d["role"] for d in devs]
        ns, nj = roles.count("sender"), roles.count("jammer")
        if roles != ["sender"] * ns + ["rrm"] + ["jammer"] * nj:
            raise ValueError("device order on a band is senders, RRM, PHY-only senders")
        if not 2 <= ns <= N.GW_GENBAND_MAX_SENDERS or nj > N.GW_GENBAND_MAX_PHY_SENDERS:
            raise ValueError("2..%d senders and at most %d PHY-only senders per band"
                             % (N.GW_GENBAND_MAX_SENDERS, N.GW_GENBAND_MAX_PHY_SENDERS))
        self.n_senders, self.n_phy_senders, self.n_devices = ns, nj, ns + 1 + nj
        self.ASSIGNMENT_DURATION_FACTOR = int(scenario.get("assignment_duration_factor", BaseEnv.ASSIGNMENT_DURATION_FACTOR))
        super().__init__(None, deviceCount=ns)
        self.observation_space = spaces.Discrete(2 * self.COUNTER_BOUND)

        cfg = N.GenBandConfig()
        cfg.abi_version = N.GW_ABI_VERSION
        cfg.n_envs, cfg.n_senders, cfg.n_phy_senders = self.num_envs, ns, nj
        cfg.assignment_duration_factor = self.ASSIGNMENT_DURATION_FACTOR
        cfg.max_assign_duration = self.MAX_ASSIGN_DURATION
        cfg.per_env_positions = 0 if positions is None else 1
        cfg.mode = N.GW_MODE_REFERENCE if self.mode == "reference" else N.GW_MODE_MASK_PHILOX
        cfg.seed, cfg.env_id_offset = int(seed), int(env_id_offset)
        cfg.frequency_hz, cfg.bandwidth_hz = float(band.get("frequency", 2.4e9)), float(band.get("bandwidth", 22e6))
        for k, d in enumerate(devs[:ns]):
            p = d.get("payload", "counter")
            cfg.multiplicity[k] = int(d["mult"])
            cfg.payload_bytes[k] = -1 if p == "counter" else int(p)
            cfg.destination[k] = int(d["dest"])
            cfg.max_ticks[k] = int(d.get("max_ticks", 0))
            cfg.receive[k] = 1 if d.get("receive") else 0
            cfg.interval[k] = float(d.get("interval", 0.001))
        for j, d in enumerate(devs[ns + 1:]):
            cfg.phy_interval[j], cfg.phy_delay[j] = float(d["interval"]), float(d["delay"])
            cfg.phy_power_dbm[j] = float(d.get("power", 0.0))
            cfg.phy_header_bytes[j], cfg.phy_payload_bytes[j] = int(d.get("hdr", 13)), int(d["payload"])
        if positions is None:
            pos = torch.tensor([[float(d["x"]), float(d["y"])] for d in devs], dtype=torch.float64, device=self.device)
        else:
            pos = torch.as_tensor(positions, dtype=torch.float64, device=self.device).contiguous()
            if tuple(pos.shape) != (self.num_envs, self.n_devices, 2):
                raise ValueError("positions must have shape [num_envs, n_devices, 2]")
        self._lib = N.lib()
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            N.check(self._lib.gw_genband_create(C.byref(cfg), self.device.index, pos.data_ptr(), self._stream(),
                                                C.byref(self._handle)))
        n = self.num_envs
        self._obs = torch.full((n,), self.COUNTER_BOUND, dtype=torch.int64, device=self.device)
        self._reward = torch.zeros(n, dtype=torch.float64, device=self.device)
        self._done = torch.zeros(n, dtype=torch.uint8, device=self.device)
        self._shape_t = (n,)                # result shape of a batched step (the learner's allocation-free loop)
        self.env_id_offset = int(env_id_offset)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def close(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            self._lib.gw_genband_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # gym surface ------------------------------------------------------------------------------------------
    def reset(self):
        """``counter_traffic.py:135-144``: sender counters and the interpreter are reset; time, queues and PHY
        state stay.  Returns the observation(s)."""
        obs = self._obs if self._scalar_api else torch.empty(self.num_envs, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self._lib.gw_genband_reset(self._handle, obs.data_ptr(), self._stream()))
        return int(obs[0]) if self._scalar_api else obs

    def _actions(self, action):
        dev, dur = action["device"], action["duration"]
        if self._scalar_api and not torch.is_tensor(dev):
            assert self.action_space.contains({"device": int(dev), "duration": int(dur)})
        dev = torch.as_tensor(dev, dtype=torch.int32, device=self.device).reshape(self.num_envs).contiguous()
        dur = torch.as_tensor(dur, dtype=torch.int32, device=self.device).reshape(self.num_envs).contiguous()
        return dev, dur

    def step(self, action, out=None):
        """``counter_traffic.py:146-158`` for every env: ``action = {"device": int32 [num_envs], "duration": int32
        [num_envs]}`` (CUDA tensors, or Python ints for ``num_envs == 1``) -> ``(obs, reward, done, info)``.
        ``out``: optional ``(int64, float64, bool)`` tensors of shape ``[num_envs]`` to write the results into (an
        allocation-free loop, as ``CounterTrafficEnv.step``); otherwise fresh tensors are returned."""
        dev, dur = self._actions(action)
        if out is not None:
            obs, reward, done = out
        elif self._scalar_api:
            obs, reward, done = self._obs, self._reward, self._done
        else:
            obs = torch.empty(self.num_envs, dtype=torch.int64, device=self.device)
            reward = torch.empty(self.num_envs, dtype=torch.float64, device=self.device)
            done = torch.empty(self.num_envs, dtype=torch.bool, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self._lib.gw_genband_step(self._handle, dev.data_ptr(), dur.data_ptr(), obs.data_ptr(),
                                              reward.data_ptr(), done.data_ptr(), self._stream()))
        if self.strict:
            self.check()
        if self._scalar_api and out is None:
            return int(obs[0]), float(reward[0]), bool(done[0]), {"Latest received values": str(self.received_values()[0].tolist())}
        return obs, reward, done if done.dtype == torch.bool else done.to(torch.bool), _LazyInfo(self)

    def step_traced(self, action, cap=8192):
        """:meth:`step` plus the event trace per env (tuples as ``CounterTrafficEnv.step_traced``)."""
        dev, dur = self._actions(action)
        n = self.num_envs
        trace = torch.zeros((n, cap, 8), dtype=torch.float64, device=self.device)
        count = torch.zeros(n, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self._lib.gw_genband_step_traced(self._handle, dev.data_ptr(), dur.data_ptr(), self._obs.data_ptr(),
                                                     self._reward.data_ptr(), self._done.data_ptr(), trace.data_ptr(),
                                                     count.data_ptr(), int(cap), self._stream()))
        self.check()
        tr, cn = trace.cpu().numpy(), count.cpu().numpy()
        if (cn > cap).any():
            raise RuntimeError("trace truncated: raise cap (max count %d)" % int(cn.max()))
        recs = [_trace_tuples(tr[i, :cn[i]]) for i in range(n)]
        if self._scalar_api:
            return int(self._obs[0]), float(self._reward[0]), bool(self._done[0]), recs[0]
        return self._obs, self._reward, self._done.to(torch.bool), recs

    def set_positions(self, positions):
        """Devices moving between steps: float64 ``[num_envs, n_devices, 2]``; every env moves its devices one after the
        other by ascending index like successive ``device.position.set(x, y)`` calls (``devices/core.py:75-84``) --
        transmissions that are on the air see ``SimplePhy._onAttenuationChange``.  Needs an env created with
        ``positions=`` (per-env geometries)."""
        pos = torch.as_tensor(positions, dtype=torch.float64, device=self.device).contiguous()
        if tuple(pos.shape) != (self.num_envs, self.n_devices, 2):
            raise ValueError("positions must have shape [num_envs, n_devices, 2]")
        with torch.cuda.device(self.device):
            N.check(self._lib.gw_genband_set_positions(self._handle, pos.data_ptr(), self._stream()))
        if self.strict:
            self.check()

    def set_movers(self, move_delays, offsets, move_interval=1e-3):
        """Mobility processes that move devices DURING the steps (the mover of ``tests/test_benchmark.py:73-85``):
        ``move_delays`` float64 ``[num_envs, n_devices]`` (first delay; negative: no process for that device),
        ``offsets`` float64 ``[num_envs, n_devices, K, 2]`` (the jumps, which accumulate; a process ends with its tape),
        one jump every ``move_interval``.  Needs per-env geometries; once per env object."""
        md = torch.as_tensor(move_delays, dtype=torch.float64, device=self.device).contiguous()
        self._offsets = torch.as_tensor(offsets, dtype=torch.float64, device=self.device).contiguous()
        if tuple(md.shape) != (self.num_envs, self.n_devices) or self._offsets.dim() != 4 or \
                tuple(self._offsets.shape[:2]) != (self.num_envs, self.n_devices) or self._offsets.shape[3] != 2:
            raise ValueError("move_delays [num_envs, n_devices], offsets [num_envs, n_devices, K, 2]")
        with torch.cuda.device(self.device):
            N.check(self._lib.gw_genband_set_movers(self._handle, md.data_ptr(), self._offsets.data_ptr(),
                                                    int(self._offsets.shape[2]), float(move_interval), self._stream()))
            torch.cuda.current_stream(self.device).synchronize()        # `md` is consumed on the stream

    def check(self):
        """Synchronises and raises if an action was outside the action space or an env hit a condition under
        which the reference raises."""
        with torch.cuda.device(self.device):
            N.check(self._lib.gw_genband_check(self._handle, self._stream()))

    # read-back ----------------------------------------------------------------------------------------------
    def _read(self, field, shape):
        out = torch.zeros(shape, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self._lib.gw_genband_read(self._handle, field, out.data_ptr(), self._stream()))
        return out

    @property
    def now(self):
        """``SimMan.now`` of every env."""
        return self._read(N.GW_GENBAND_FIELD_NOW, (self.num_envs,))

    def delivered(self):
        """int64 ``[num_envs, n_senders]``: data packets of each sender decoded by the RRM."""
        return self._read(N.GW_GENBAND_FIELD_DELIVERED, (self.n_senders, self.num_envs)).t().to(torch.int64)

    def received(self):
        """int64 ``[num_envs, n_senders]``: packets handed to ``onReceive`` (MAC receive mode)."""
        return self._read(N.GW_GENBAND_FIELD_RECEIVED, (self.n_senders, self.num_envs)).t().to(torch.int64)

    def received_values(self):
        """int64 ``[num_envs, n_senders]``: the interpreter's ``receivedValues`` (``counter_traffic.py:69-80``)."""
        return self._read(N.GW_GENBAND_FIELD_RECEIVED_VALUES, (self.n_senders, self.num_envs)).t().to(torch.int64)

    def transmissions(self):
        return self._read(N.GW_GENBAND_FIELD_TRANSMISSIONS, (self.num_envs,)).to(torch.int64)

    def faults(self):
        return self._read(N.GW_GENBAND_FIELD_FAULT, (self.num_envs,)).to(torch.int64)

    def ties(self):
        return self._read(N.GW_GENBAND_FIELD_TIES, (self.num_envs,)).to(torch.int64)

    def received_power(self):
        return self._read(N.GW_GENBAND_FIELD_RECEIVED_POWER, (self.n_devices, self.num_envs)).t()

    def queue_lengths(self):
        return self._read(N.GW_GENBAND_FIELD_QUEUE_LENGTH, (self.n_senders, self.num_envs)).t().to(torch.int64)

    def counters(self):
        return self._read(N.GW_GENBAND_FIELD_COUNTER, (self.n_senders, self.num_envs)).t().to(torch.int64)
