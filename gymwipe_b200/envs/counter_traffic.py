"""
Batched, GPU-resident drop-in for ``gymwipe.envs.counter_traffic.CounterTrafficEnv``
(``gymwipe/envs/counter_traffic.py:20-162``).

The gym surface is kept: ``reset`` / ``step`` / ``seed`` / ``render`` / ``action_space`` /
``observation_space``, the class constants, ``senders`` / ``rrm`` / ``frequencyBand`` /
``deviceIndexToMacDict``.  One object holds ``num_envs`` independent envs whose state lives
in one PyTorch CUDA tensor (structure of arrays, see ``gymwipe_b200/csrc/gw_kernels.cu``);
``step`` launches the fused sm_100a step kernel through the C ABI.  With ``num_envs == 1``
and Python-int actions it returns Python scalars and passes the reference's own test
(``tests/envs/test_counter_traffic.py``) verbatim.
"""
import ctypes as C

import numpy as np
import torch

from gymwipe_b200 import _native as N
from gymwipe_b200 import scenario as S
from gymwipe_b200 import spaces
from gymwipe_b200.envs.core import BaseEnv, Interpreter
from gymwipe_b200.networking.attenuation_models import FsplAttenuation
from gymwipe_b200.networking.devices import PhySenderDevice, SimpleNetworkDevice, SimpleRrmDevice
from gymwipe_b200.networking.physical import FrequencyBand


def _require_cuda(device):
    if not torch.cuda.is_available():
        raise RuntimeError("gymwipe_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("gymwipe_b200 envs live on a CUDA device, got %r" % (device,))
    return torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())


class LazyInfo(dict):
    """``info`` of a batched step: values are read back from the device on first access."""

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


class CounterTrafficEnv(BaseEnv):
    """
    Two sender devices that send counter packets to each other and an RRM whose interpreter
    turns the packets it overhears into observations and rewards
    (``counter_traffic.py:20-30``); ``num_envs`` of them, stepped together on the GPU.

    Args:
        num_envs: number of independent envs in this batch (this GPU's shard).
        device: CUDA device of the batch.
        mode: ``"reference"`` (mode R: the reference's deterministic expected-value error
            accounting), ``"mask_philox"`` (mode M: per-bit Philox4x32-10 error masks generated
            in the kernel) or ``"mask_fed"`` (mode M with masks supplied by :meth:`set_masks`).
        seed: Philox seed of mode M.
        env_id_offset: global id of env 0 (sharding across GPUs keeps results invariant).
        scenario: optional scenario dict (``gymwipe_b200.scenario``) replacing the default devices.
        positions: optional float64 tensor ``[num_envs, n_bands, 4, 2]`` of per-env positions.
        strict: validate actions after every step (costs a device sync).  Defaults to True for
            ``num_envs == 1``; out-of-space actions raise like the reference's ``assert``.
    """

    COUNTER_INTERVAL = 0.001

    COUNTER_BYTE_LENGTH = 2

    COUNTER_BOUND = 2 ** (8 * COUNTER_BYTE_LENGTH)

    class SenderDevice(SimpleNetworkDevice):
        """``counter_traffic.py:37-61``: sends ``packetMultiplicity`` packets every ``COUNTER_INTERVAL``."""

        def __init__(self, name, xPos, yPos, frequencyBand, packetMultiplicity, macIndex=0,
                     payloadRule="counter", interval=0.001):
            super().__init__(name, xPos, yPos, frequencyBand, macIndex)
            self.packetMultiplicity = packetMultiplicity
            self.payloadRule = payloadRule
            self.interval = interval
            self.destinationMac = None
            self._env = None
            self._index = None

        @property
        def counter(self):
            """Current counter value(s), read back from the device."""
            v = self._env.read_state(N.GW_FIELD_COUNTER)[self._index]
            return int(v[0]) if self._env._scalar_api else v.to(torch.int64)

    class CounterTrafficInterpreter(Interpreter):
        """``counter_traffic.py:63-112`` -- evaluated inside the step kernel; this is a view."""

        def __init__(self, env):
            self._env = env

        def reset(self):
            self._env.reset()

        def onPacketReceived(self, senderIndex, receiverIndex, payload):
            raise RuntimeError("packets are delivered inside the CUDA step kernel")

        @property
        def receivedValues(self):
            v = self._env.received_values()
            return [int(x) for x in v[0]] if self._env._scalar_api else v

        def getReward(self):
            return self._env._last_reward

        def getObservation(self):
            return self._env._last_obs

        def getDone(self):
            return self._env._last_done

        def getInfo(self):
            return {"Latest received values": str(self.receivedValues)}

    def __init__(self, num_envs=1, device="cuda", mode="reference", seed=0, env_id_offset=0,
                 scenario=None, positions=None, strict=None):
        self.device = _require_cuda(device)
        self.num_envs = int(num_envs)
        self._scalar_api = self.num_envs == 1
        self.strict = self._scalar_api if strict is None else bool(strict)
        self.scenario = scenario if scenario is not None else S.default_scenario_dict()
        self.n_bands = len(self.scenario["bands"])
        self.ASSIGNMENT_DURATION_FACTOR = int(self.scenario.get("assignment_duration_factor",
                                                                BaseEnv.ASSIGNMENT_DURATION_FACTOR))

        # descriptor objects with the reference's names (counter_traffic.py:114-133)
        self.frequencyBands = []
        self.senders = []
        self.rrms = []
        self.jammers = []
        mac_counter = 0
        for b, bd in enumerate(self.scenario["bands"]):
            band = FrequencyBand([FsplAttenuation], bd.get("frequency", 2.4e9), bd.get("bandwidth", 22e6))
            self.frequencyBands.append(band)
            band_senders = []
            for d in bd["devices"]:
                if d["role"] == "sender":
                    mac_counter += 1
                    sd = CounterTrafficEnv.SenderDevice("Sender %d" % mac_counter, d["x"], d["y"], band,
                                                        d["mult"], mac_counter, d.get("payload", "counter"),
                                                        d.get("interval", 0.001))
                    sd._env, sd._index = self, len(band_senders) if b == 0 else None
                    band_senders.append(sd)
            idx2mac = {i: s.macAddr for i, s in enumerate(band_senders)}
            if len(band_senders) == 2:
                band_senders[0].destinationMac = band_senders[1].macAddr
                band_senders[1].destinationMac = band_senders[0].macAddr
            for d in bd["devices"]:
                if d["role"] == "rrm":
                    self.rrms.append(SimpleRrmDevice("RRM", d["x"], d["y"], band, idx2mac,
                                                     CounterTrafficEnv.CounterTrafficInterpreter(self)))
            for d in bd["devices"]:
                if d["role"] == "jammer":
                    self.jammers.append(PhySenderDevice("Jammer", d["x"], d["y"], band, d["interval"], d["delay"],
                                                        d.get("power", 0.0), d.get("hdr", 13), d["payload"]))
            if b == 0:
                self.senders = band_senders
                self.deviceIndexToMacDict = idx2mac
        self.rrm = self.rrms[0]
        super().__init__(self.frequencyBands[0], deviceCount=2)
        self.frequencyBand = self.frequencyBands[0]

        # the observation is latestDifference + COUNTER_BOUND (counter_traffic.py:118-120)
        self.observation_space = spaces.Discrete(2 * CounterTrafficEnv.COUNTER_BOUND)

        # native handle; the env-batch state is a torch tensor
        self._lib = N.lib()
        self._cfg = S.config_from_dict(self.scenario, self.num_envs, mode=mode, seed=seed,
                                       env_id_offset=env_id_offset,
                                       per_env_positions=positions is not None or self._needs_per_env_tables(),
                                       max_assign_duration=self.MAX_ASSIGN_DURATION)
        self._configure_native(self._cfg)
        nbytes = C.c_size_t()
        N.check(self._lib.gw_state_bytes(C.byref(self._cfg), C.byref(nbytes)))
        self.state = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            N.check(self._lib.gw_create(C.byref(self._cfg), self.device.index, self.state.data_ptr(),
                                        nbytes.value, self._stream(), C.byref(self._handle)))
        self._masks = None
        self._positions = None
        if positions is not None:
            self.set_positions(positions)
        self._shape = (self.num_envs,) if self.n_bands == 1 else (self.num_envs, self.n_bands)
        self._shape_t = torch.Size(self._shape)
        self._last_obs = self.COUNTER_BOUND
        self._last_reward = 0.0
        self._last_done = False
        self._stats_out = torch.zeros(8, dtype=torch.float64, device=self.device)

    # ------------------------------------------------------------------ plumbing
    def _needs_per_env_tables(self):
        return False

    def _configure_native(self, cfg):
        """Hook for subclasses (plant envs) to extend the native config before ``gw_create``."""

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def close(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            self._lib.gw_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self):
        """Synchronises and raises if a kernel flagged an out-of-space action or a sim fault."""
        with torch.cuda.device(self.device):
            rc = self._lib.gw_check(self._handle, self._stream())
        if rc == N.GW_E_ACTION:
            raise ValueError(self._lib.gw_last_error().decode())
        N.check(rc)

    def read_state(self, field):
        """Dense float64 read-back of one state field (``GW_FIELD_*``), shaped ``[k, n_sims]``."""
        nsim = self.num_envs * self.n_bands
        rows = {N.GW_FIELD_NOW: None, N.GW_FIELD_RECEIVED_POWER: 4, N.GW_FIELD_NEXT_TICK: 2, N.GW_FIELD_COUNTER: 2,
                N.GW_FIELD_QUEUE_LEN: 2, N.GW_FIELD_N_TRANSMISSIONS: 1, N.GW_FIELD_N_DELIVERED: 2,
                N.GW_FIELD_RECEIVED_VALUES: 2, N.GW_FIELD_ATTENUATION_DB: 16, N.GW_FIELD_RX_POWER_MW: 16,
                N.GW_FIELD_FAULT: 1, N.GW_FIELD_TIES: 1, N.GW_FIELD_TX_SEQ: 4, N.GW_FIELD_PLANT: 8,
                N.GW_FIELD_N_RECEIVED: 2}[field]
        if rows is None:
            out = torch.zeros(self.num_envs, dtype=torch.float64, device=self.device)
        else:
            out = torch.zeros((rows, nsim), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self._lib.gw_read_state(self._handle, field, out.data_ptr(), self._stream()))
        return out

    @property
    def now(self):
        """Simulated time of every env (``SimMan.now``)."""
        t = self.read_state(N.GW_FIELD_NOW)
        return float(t[0]) if self._scalar_api else t

    def received_values(self):
        """``interpreter.receivedValues`` as an int64 tensor ``[n_sims, 2]``."""
        return self.read_state(N.GW_FIELD_RECEIVED_VALUES).t().to(torch.int64)

    def delivered(self):
        """Packets the RRM decoded per sender, int64 ``[n_sims, 2]``."""
        return self.read_state(N.GW_FIELD_N_DELIVERED).t().to(torch.int64)

    def received(self):
        """MAC receive mode (scenario key ``"receive": True`` of a sender = ``SimpleNetworkDevice.receiving = True``):
        packets handed to ``onReceive`` per sender, int64 ``[n_sims, 2]``."""
        return self.read_state(N.GW_FIELD_N_RECEIVED).t().to(torch.int64)

    def transmissions(self):
        return self.read_state(N.GW_FIELD_N_TRANSMISSIONS)[0].to(torch.int64)

    def stats(self, clear=True, out=None):
        """
        Statistics accumulated by the step kernel's epilogue since the last clearing call
        (float64[8] on the device, see ``gw_stats``); ``out`` lets the caller supply the tensor
        (e.g. the all-reduce buffer of ``gymwipe_b200.distributed.StatsReducer``).
        """
        out = self._stats_out if out is None else out
        with torch.cuda.device(self.device):
            N.check(self._lib.gw_stats(self._handle, out.data_ptr(), 1 if clear else 0, self._stream()))
        return out

    def mask_bytes(self, clear=True):
        """Mode ``mask_fed``: bytes of mask words the step kernels scanned since the last clearing call
        (``gw_mask_bytes``; synchronises)."""
        out = C.c_uint64(0)
        with torch.cuda.device(self.device):
            N.check(self._lib.gw_mask_bytes(self._handle, C.byref(out), 1 if clear else 0, self._stream()))
        return int(out.value)

    def share_stats(self, other):
        """
        Accumulate this env's step statistics into ``other``'s vector (``gw_share_stats``): several env
        batches on one GPU then need one :meth:`stats` call in front of the all-reduce.  ``None`` restores
        this env's own accumulators.  Keep ``other`` alive while sharing.
        """
        N.check(self._lib.gw_share_stats(self._handle, other._handle if other is not None else None))
        self._stats_owner = other

    def set_positions(self, positions):
        """
        Per-env device positions ``[num_envs, n_bands, 4, 2]`` (float64, CUDA).  At construction the
        devices are created there; once the env has been stepped the call MOVES them (the reference's
        ``Position.set`` between two ``step`` calls, devices by ascending index): transmissions that
        are on the air see ``SimplePhy._onAttenuationChange`` (``gw_set_positions``).
        """
        p = torch.as_tensor(positions, dtype=torch.float64, device=self.device).contiguous()
        assert tuple(p.shape) == (self.num_envs, self.n_bands, N.GW_MAX_DEVICES, 2)
        self._positions = p
        with torch.cuda.device(self.device):
            N.check(self._lib.gw_set_positions(self._handle, p.data_ptr(), self._stream()))

    def set_masks(self, mask_words, slots):
        """Mode ``mask_fed``: uint32/int32 tensor ``[num_envs, n_bands, 4, slots, 4, words]``."""
        assert mask_words.is_cuda and mask_words.dtype in (torch.int32, torch.uint32)
        assert tuple(mask_words.shape[:5]) == (self.num_envs, self.n_bands, 4, slots, 4)
        self._masks = mask_words.contiguous()
        N.check(self._lib.gw_set_masks(self._handle, self._masks.data_ptr(), int(slots),
                                       int(self._masks.shape[5]), self._stream()))

    # ------------------------------------------------------------------ gym API
    def reset(self, env_ids=None):
        """
        ``counter_traffic.py:135-144``: counters := 0 and interpreter reset (simulated time,
        queues and PHY state are kept, as in the reference).  Returns the observation.
        """
        obs = torch.empty(self._shape, dtype=torch.int64, device=self.device)
        ids_ptr, n = None, 0
        if env_ids is not None:
            ids = torch.as_tensor(env_ids, dtype=torch.int64, device=self.device).contiguous()
            ids_ptr, n = ids.data_ptr(), ids.numel()
            obs.fill_(self.COUNTER_BOUND)
        with torch.cuda.device(self.device):
            N.check(self._lib.gw_reset(self._handle, ids_ptr, n, obs.data_ptr(), self._stream()))
        if self._scalar_api and self.n_bands == 1:
            self._last_obs = int(obs[0])
            return self._last_obs
        return obs

    def _prepare_action(self, action):
        if torch.is_tensor(action) or isinstance(action, np.ndarray):
            # flat action of the DQN agent: a // 20, a % 20 (agents/dqn_counter_traffic.py:25-33)
            a = torch.as_tensor(action, device=self.device)
            action = {"device": a // self.MAX_ASSIGN_DURATION, "duration": a % self.MAX_ASSIGN_DURATION}
        if not isinstance(action, dict) or set(action) != {"device", "duration"}:
            raise ValueError("action must be a dict with the keys 'device' and 'duration'")
        out = []
        for key in ("device", "duration"):
            v = action[key]
            if not torch.is_tensor(v):
                v = torch.as_tensor(np.asarray(v), device=self.device)
            if v.is_floating_point() or v.dtype == torch.bool:
                raise ValueError("action[%r] must be an integer tensor" % key)
            v = v.to(device=self.device, dtype=torch.int32).reshape(self._shape).contiguous()
            out.append(v)
        return out

    def _fast_action(self, action):
        """Zero-copy path: a dict of contiguous int32 CUDA tensors of the batch shape."""
        if type(action) is dict and len(action) == 2:
            dev, dur = action.get("device"), action.get("duration")
            if (torch.is_tensor(dev) and torch.is_tensor(dur) and dev.dtype == torch.int32 and dur.dtype == torch.int32
                    and dev.device == self.device and dur.device == self.device
                    and dev.shape == self._shape_t and dur.shape == self._shape_t
                    and dev.is_contiguous() and dur.is_contiguous()):
                return dev, dur
        return None

    def step(self, action, out=None):
        """
        ``counter_traffic.py:146-158``: assigns the band (``action["device"]``) for
        ``action["duration"] * ASSIGNMENT_DURATION_FACTOR`` slots in every env and simulates
        until the assignment ends.  Returns ``(obs, reward, done, info)``.  ``out``: an optional
        ``(obs int64, reward float64, done bool)`` triple of CUDA tensors of the batch shape that receives the
        results (a loop that steps many times re-uses one triple instead of allocating three tensors per call).
        """
        fast = self._fast_action(action)
        scalar = False
        if fast is not None:
            dev, dur = fast
        else:
            scalar = self._scalar_api and self.n_bands == 1 and isinstance(action, dict) and \
                not torch.is_tensor(action.get("device"))
            if scalar:
                assert self.action_space.contains(action)
            dev, dur = self._prepare_action(action)
        if out is not None:
            obs, reward, done = out
            assert obs.dtype == torch.int64 and reward.dtype == torch.float64 and done.dtype == torch.bool
            assert obs.shape == self._shape_t and reward.shape == self._shape_t and done.shape == self._shape_t
            assert obs.is_contiguous() and reward.is_contiguous() and done.is_contiguous() and obs.device == self.device
        else:
            obs = torch.empty(self._shape, dtype=torch.int64, device=self.device)
            reward = torch.empty(self._shape, dtype=torch.float64, device=self.device)
            done = torch.empty(self._shape, dtype=torch.bool, device=self.device)
        if torch.cuda.current_device() == self.device.index:
            rc = self._lib.gw_step(self._handle, dev.data_ptr(), dur.data_ptr(), obs.data_ptr(),
                                   reward.data_ptr(), done.data_ptr(), torch.cuda.current_stream().cuda_stream)
        else:
            with torch.cuda.device(self.device):
                rc = self._lib.gw_step(self._handle, dev.data_ptr(), dur.data_ptr(), obs.data_ptr(),
                                       reward.data_ptr(), done.data_ptr(), self._stream())
        if rc:
            N.check(rc)
        if self.strict:
            self.check()
        if scalar:
            self._last_obs, self._last_reward = int(obs[0]), float(reward[0])
            self._last_done = bool(done[0])
            return self._last_obs, self._last_reward, self._last_done, self.rrm.interpreter.getInfo()
        self._last_obs, self._last_reward, self._last_done = obs, reward, done
        return obs, reward, done, LazyInfo(self)

    def step_traced(self, action, cap=1024):
        """
        :meth:`step` plus the device-side event trace (``gw_step_traced``): returns
        ``(obs, reward, done, records)`` where ``records[sim]`` is the list of
        ``("tx", t, band, sender, stop, headerBits, payloadBits)`` /
        ``("ber", t, band, receiver, ber)`` / ``("dec", t, band, receiver, section, errSum, bits, ok)`` /
        ``("rx", t, band, senderIndex)`` tuples of that band-sim, in event order -- the read-back
        view of the reference's ``Transmission`` objects and decider verdicts.
        """
        dev, dur = self._prepare_action(action)
        nsim = self.num_envs * self.n_bands
        obs = torch.empty(self._shape, dtype=torch.int64, device=self.device)
        reward = torch.empty(self._shape, dtype=torch.float64, device=self.device)
        done = torch.empty(self._shape, dtype=torch.uint8, device=self.device)
        trace = torch.zeros((nsim, cap, 8), dtype=torch.float64, device=self.device)
        count = torch.zeros(nsim, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self._lib.gw_step_traced(self._handle, dev.data_ptr(), dur.data_ptr(), obs.data_ptr(),
                                             reward.data_ptr(), done.data_ptr(), trace.data_ptr(), count.data_ptr(),
                                             int(cap), self._stream()))
        self.check()
        tr, cn = trace.cpu().numpy(), count.cpu().numpy()
        if (cn > cap).any():
            raise RuntimeError("trace truncated: raise cap (max count %d)" % int(cn.max()))
        records = []
        for i in range(nsim):
            band = i % self.n_bands
            out = []
            for r in tr[i, :cn[i]]:
                k = int(r[0])
                if k == 1:
                    out.append(("tx", float(r[1]), band, int(r[2]), float(r[3]), float(r[4]), float(r[5])))
                elif k == 2:
                    out.append(("ber", float(r[1]), band, int(r[2]), float(r[3])))
                elif k == 3:
                    out.append(("dec", float(r[1]), band, int(r[2]), int(r[3]), float(r[4]), float(r[5]), bool(r[6])))
                elif k == 4:
                    out.append(("rx", float(r[1]), band, int(r[2])))
                elif k == 5:
                    out.append(("mrx", float(r[1]), band, int(r[2])))
            records.append(out)
        return obs, reward, done.bool(), records

    def step_host(self, device, duration, obs, reward, done):
        """
        End-to-end step with HOST buffers (numpy arrays or pinned CPU tensors): actions are
        copied to the GPU, the step kernel runs, results are copied back (``gw_step_host``).
        ``device``/``duration`` int32, ``obs`` int64, ``reward`` float64, ``done`` uint8.
        """
        def ptr(a):
            return a.data_ptr() if torch.is_tensor(a) else a.ctypes.data
        with torch.cuda.device(self.device):
            N.check(self._lib.gw_step_host(self._handle, ptr(device), ptr(duration), ptr(obs), ptr(reward),
                                           ptr(done), self._stream()))

    def step_host_packed(self, actions, results):
        """
        Throughput variant of :meth:`step_host` (``gw_step_host_packed``): ``actions`` is a pinned
        int32 tensor / array ``[2, n_sims]`` (row 0 device, row 1 duration), ``results`` a pinned
        uint8 buffer of ``9 * n_sims`` bytes that receives ``int32 obs | float32 reward | uint8 done``.
        Use :meth:`unpack_results` for typed views.  One copy in, one copy out.
        """
        def ptr(a):
            return a.data_ptr() if torch.is_tensor(a) else a.ctypes.data
        with torch.cuda.device(self.device):
            N.check(self._lib.gw_step_host_packed(self._handle, ptr(actions), ptr(results), self._stream()))

    def step_host_compact_async(self, actions, results):
        """
        :meth:`step_host_compact` without the final synchronisation (``gw_step_host_compact_async``): for
        callers that keep several env batches in flight.  Pinned buffers only; ``results`` is valid once
        the current stream (or an event recorded after this call) has completed.
        """
        a = actions.data_ptr() if torch.is_tensor(actions) else actions.ctypes.data
        r = results.data_ptr() if torch.is_tensor(results) else results.ctypes.data
        rc = self._lib.gw_step_host_compact_async(self._handle, a, r, torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            N.check(rc)

    def step_host_compact(self, actions, results):
        """
        Compact end-to-end step (``gw_step_host_compact``): ``actions`` is a pinned uint8 tensor / array
        ``[n_sims, 2]`` (device, duration), ``results`` a pinned int32 / uint32 buffer of ``n_sims`` words
        that receives ``obs | (reward + 16) << 17 | done << 22``.  2 bytes in, 4 bytes out per sim; pinned
        buffers are read / written in place by the kernel (no copies), pageable ones are staged.  Use
        :meth:`unpack_compact` for typed tensors.
        """
        a = actions.data_ptr() if torch.is_tensor(actions) else actions.ctypes.data
        r = results.data_ptr() if torch.is_tensor(results) else results.ctypes.data
        # (the native call selects the handle's device itself)
        rc = self._lib.gw_step_host_compact(self._handle, a, r, torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            N.check(rc)

    def step_host_tiny(self, actions, results):
        """
        The smallest wire format (``gw_step_host_tiny``): ``actions`` a pinned uint8 tensor / array ``[n_sims]``
        with ``device << 7 | duration`` (:meth:`pack_tiny_actions`), ``results`` a pinned int16 / uint16 buffer of
        ``n_sims`` words that receives ``(obs - COUNTER_BOUND) & 0xFF | (reward + 16) << 8 | done << 13``.
        1 byte in, 2 bytes out per sim; :meth:`unpack_tiny` for typed tensors.
        """
        a = actions.data_ptr() if torch.is_tensor(actions) else actions.ctypes.data
        r = results.data_ptr() if torch.is_tensor(results) else results.ctypes.data
        rc = self._lib.gw_step_host_tiny(self._handle, a, r, torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            N.check(rc)

    @staticmethod
    def pack_tiny_actions(device, duration):
        """uint8 ``device << 7 | duration`` from integer tensors / arrays (device in {0, 1}, duration < 128)."""
        d = torch.as_tensor(device).to(torch.int32)
        return ((d << 7) | torch.as_tensor(duration).to(torch.int32)).to(torch.uint8)

    @staticmethod
    def unpack_tiny(results):
        """``(obs int64, reward float64, done bool)`` from the 16-bit words of :meth:`step_host_tiny`."""
        r = results if torch.is_tensor(results) else torch.from_numpy(results)
        w = r.view(torch.int16).to(torch.int64) & 0xFFFF
        diff = ((w & 0xFF) ^ 0x80) - 0x80                       # sign-extend the low byte
        assert not bool(((w >> 15) & 1).any()), "an observation did not fit the 8-bit difference"
        return diff + 65536, (((w >> 8) & 31) - 16).to(torch.float64), ((w >> 13) & 1).bool()

    @staticmethod
    def unpack_compact(results):
        """``(obs int64, reward float64, done bool)`` from the packed words of :meth:`step_host_compact`."""
        r = results if torch.is_tensor(results) else torch.from_numpy(results)
        w = r.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
        return w & 0x1FFFF, (((w >> 17) & 31) - 16).to(torch.float64), ((w >> 22) & 1).bool()

    def unpack_results(self, results):
        """Typed views (obs int32, reward float32, done uint8) of a packed result buffer."""
        n = self.num_envs * self.n_bands
        r = results if torch.is_tensor(results) else torch.from_numpy(results)
        return r[:4 * n].view(torch.int32), r[4 * n:8 * n].view(torch.float32), r[8 * n:9 * n]

    def render(self, mode='human', close=False):
        """``counter_traffic.py:160-162`` (env 0)."""
        values = [int(x) for x in self.received_values()[0]]
        print("Last Received: {}, difference: {:6d}".format(values, values[1] - values[0]), end='\r')
