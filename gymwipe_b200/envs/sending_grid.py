"""
Batched counterpart of the reference's benchmark scenario (``tests/test_benchmark.py:20-91``): grids of
PHY-only ``SendingDevice`` s -- every device sends a 13 + 26 byte packet at 40 dBm every ``SEND_INTERVAL`` after an
initial delay, every ``SimplePhy`` receives what the others send -- optionally with the mobility processes of
the ``mobile_device_grid`` fixture (a jump of up to +-0.2 m per axis every ``MOVE_INTERVAL`` while transmissions
are on the air), advanced with :meth:`runSimulation` like ``SimMan.runSimulation``.  ``num_envs`` independent
grids are simulated by the CUDA engine of ``gymwipe_b200/csrc/gw_grid.cuh`` (one thread per grid; the device count is
a run-time value up to 24).  The reference draws the delays and offsets with ``random.uniform``; here they are
tensors (given, or drawn from a seeded ``torch.Generator``), so a run can be repeated and compared.
"""
import ctypes as C
from math import sqrt

import torch

from gymwipe_b200 import _native as N

SEND_INTERVAL = 1e-2        # tests/test_benchmark.py:17
MOVE_INTERVAL = 1e-3        # tests/test_benchmark.py:18


def grid_positions(n):
    """``device_grid`` fixture (:63-69): device i at ``(i / cols, i % cols)`` with ``cols = int(sqrt(n))``."""
    cols = int(sqrt(n)) if n > 0 else 1
    return [(i / cols, float(i % cols)) for i in range(n)]


class SendingDeviceGrid:
    """
    Args:
        num_envs: independent grids in this batch.
        n_devices: ``SendingDevice`` s per grid (the fixture is parametrised with 0, 2, ..., 20).
        mobile: add the mobility processes (``mobile_device_grid``); ``max_moves`` jumps per device are drawn.
        positions / delays / move_delays / offsets: optional float64 tensors ``[num_envs, n, 2]`` / ``[num_envs, n]``
            / ``[num_envs, n]`` / ``[num_envs, n, max_moves, 2]`` replacing the defaults (the fixture's grid; draws
            from ``U(0, SEND_INTERVAL)``, ``U(0, MOVE_INTERVAL)``, ``U(-.2, .2)`` with ``seed``).
    """

    def __init__(self, num_envs, n_devices, device="cuda", mobile=False, max_moves=1024, seed=0, positions=None,
                 delays=None, move_delays=None, offsets=None, power=40.0, send_interval=SEND_INTERVAL,
                 header_bytes=13, payload_bytes=26, move_interval=MOVE_INTERVAL, frequency=2.4e9, bandwidth=22e6):
        if not torch.cuda.is_available():
            raise RuntimeError("gymwipe_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device(device)
        self.device = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        self.num_envs, self.n_devices = int(num_envs), int(n_devices)
        if not 1 <= self.n_devices <= N.GW_GRID_MAX_DEVICES:
            raise ValueError("n_devices must be in 1..%d" % N.GW_GRID_MAX_DEVICES)
        n, ne = self.n_devices, self.num_envs
        g = torch.Generator(device=self.device).manual_seed(int(seed))

        def f64(t, shape):
            t = torch.as_tensor(t, dtype=torch.float64, device=self.device).contiguous()
            assert tuple(t.shape) == shape, (tuple(t.shape), shape)
            return t
        if positions is None:
            positions = torch.tensor(grid_positions(n), dtype=torch.float64).expand(ne, n, 2)
        self._positions = f64(positions, (ne, n, 2))
        if delays is None:
            delays = torch.rand((ne, n), generator=g, device=self.device, dtype=torch.float64) * send_interval
        self._delays = f64(delays, (ne, n))
        self.mobile = bool(mobile) or offsets is not None
        self._move_delays = self._offsets = None
        self.max_moves = 0
        if self.mobile:
            if offsets is None:
                offsets = torch.rand((ne, n, int(max_moves), 2), generator=g, device=self.device, dtype=torch.float64) * 0.4 - 0.2
            self.max_moves = int(torch.as_tensor(offsets).shape[2])
            self._offsets = f64(offsets, (ne, n, self.max_moves, 2))
            if move_delays is None:
                move_delays = torch.rand((ne, n), generator=g, device=self.device, dtype=torch.float64) * move_interval
            self._move_delays = f64(move_delays, (ne, n))
        cfg = N.GridConfig()
        cfg.abi_version = N.GW_ABI_VERSION
        cfg.n_envs, cfg.n_devices = ne, n
        cfg.frequency_hz, cfg.bandwidth_hz = float(frequency), float(bandwidth)
        for d in range(n):
            cfg.power_dbm[d] = float(power)
            cfg.send_interval[d] = float(send_interval)
            cfg.header_bytes[d] = int(header_bytes)
            cfg.payload_bytes[d] = int(payload_bytes)
        cfg.move_interval = float(move_interval)
        cfg.max_moves = self.max_moves
        self._lib = N.lib()
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            N.check(self._lib.gw_grid_create(
                C.byref(cfg), self.device.index, self._positions.data_ptr(), self._delays.data_ptr(),
                self._move_delays.data_ptr() if self.mobile else None, self._offsets.data_ptr() if self.mobile else None,
                self._stream(), C.byref(self._handle)))

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def close(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            self._lib.gw_grid_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def runSimulation(self, duration):
        """``SimMan.runSimulation(duration)`` for every grid of the batch (``gw_grid_run``)."""
        with torch.cuda.device(self.device):
            N.check(self._lib.gw_grid_run(self._handle, float(duration), self._stream()))

    def run_traced(self, duration, cap=65536):
        """:meth:`runSimulation` plus the event trace: per grid the list of ``("tx", t, 0, sender, stop, headerBits,
        payloadBits)`` / ``("ber", t, 0, receiver, ber)`` / ``("dec", t, 0, receiver, section, errSum, bits, ok)``."""
        ne = self.num_envs
        trace = torch.zeros((ne, cap, 8), dtype=torch.float64, device=self.device)
        count = torch.zeros(ne, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self._lib.gw_grid_run_traced(self._handle, float(duration), trace.data_ptr(), count.data_ptr(),
                                                 int(cap), self._stream()))
        self.check()
        tr, cn = trace.cpu().numpy(), count.cpu().numpy()
        if (cn > cap).any():
            raise RuntimeError("trace truncated: raise cap (max count %d)" % int(cn.max()))
        out = []
        for i in range(ne):
            recs = []
            for r in tr[i, :cn[i]]:
                k = int(r[0])
                if k == 1:
                    recs.append(("tx", float(r[1]), 0, int(r[2]), float(r[3]), float(r[4]), float(r[5])))
                elif k == 2:
                    recs.append(("ber", float(r[1]), 0, int(r[2]), float(r[3])))
                elif k == 3:
                    recs.append(("dec", float(r[1]), 0, int(r[2]), int(r[3]), float(r[4]), float(r[5]), bool(r[6])))
            out.append(recs)
        return out

    def check(self):
        with torch.cuda.device(self.device):
            N.check(self._lib.gw_grid_check(self._handle, self._stream()))

    def _read(self, field, shape):
        out = torch.zeros(shape, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            N.check(self._lib.gw_grid_read(self._handle, field, out.data_ptr(), self._stream()))
        return out

    @property
    def now(self):
        """Simulated time of every grid (``SimMan.now``)."""
        return self._read(N.GW_GRID_FIELD_NOW, (self.num_envs,))

    def stats(self):
        """int64 ``[6, n_devices, num_envs]``: transmissions started; headers decoded / failed; payloads decoded /
        failed (as a receiver); BER evaluations."""
        return self._read(N.GW_GRID_FIELD_STATS, (6, self.n_devices, self.num_envs)).to(torch.int64)

    def positions(self):
        """Current device positions ``[2, n_devices, num_envs]``."""
        return self._read(N.GW_GRID_FIELD_POSITIONS, (2, self.n_devices, self.num_envs))

    def faults(self):
        """int64 ``[num_envs]``: 0, or the condition under which the reference raises for that grid (3: the
        reference's ``assert noisePower >= 0``; 2: ``KeyError``); a faulted grid stops simulating."""
        return self._read(N.GW_GRID_FIELD_FAULT, (self.num_envs,)).to(torch.int64)

    def received_power(self):
        return self._read(N.GW_GRID_FIELD_RECEIVED_POWER, (self.n_devices, self.num_envs))
