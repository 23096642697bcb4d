"""
Loader and ctypes declarations of the C-ABI library (``include/gymwipe_b200.h``).

The library is built IN-TREE (``gymwipe_b200/lib/libgymwipe_b200.so``) from
``gymwipe_b200/csrc/gw_kernels.cu`` with ``nvcc -gencode arch=compute_100a,code=sm_100a``.
There is no CPU fallback: if the library is missing it is compiled, and if that is
impossible (or no CUDA device is present when an env is constructed) an error is raised.
"""
import ctypes as C
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.environ.get("GYMWIPE_B200_LIB") or os.path.join(LIB_DIR, "libgymwipe_b200.so")
INCLUDE = os.path.join(HERE, "..", "include", "gymwipe_b200.h")

GW_ABI_VERSION = 2
GW_MAX_BANDS, GW_MAX_DEVICES, GW_MAX_SENDERS, GW_MAX_JAMMERS = 4, 4, 2, 1
GW_OK, GW_E_INVALID, GW_E_CUDA, GW_E_STATE, GW_E_ACTION, GW_E_SIMFAULT = 0, -1, -2, -3, -4, -5
GW_MODE_REFERENCE, GW_MODE_MASK_PHILOX, GW_MODE_MASK_FED = 0, 1, 2
GW_ROLE_SENDER, GW_ROLE_RRM, GW_ROLE_JAMMER = 1, 2, 3
(GW_FIELD_NOW, GW_FIELD_RECEIVED_POWER, GW_FIELD_NEXT_TICK, GW_FIELD_COUNTER, GW_FIELD_QUEUE_LEN,
 GW_FIELD_N_TRANSMISSIONS, GW_FIELD_N_DELIVERED, GW_FIELD_RECEIVED_VALUES, GW_FIELD_ATTENUATION_DB,
 GW_FIELD_RX_POWER_MW, GW_FIELD_FAULT, GW_FIELD_TIES, GW_FIELD_TX_SEQ, GW_FIELD_PLANT, GW_FIELD_N_RECEIVED) = range(15)
GW_PLANT_NONE, GW_PLANT_SLIDING_PENDULUM = 0, 1

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false",
              "-lineinfo", "-Xcompiler", "-fPIC", "-shared"]


class DeviceConfig(C.Structure):
    _fields_ = [("role", C.c_int32), ("x", C.c_double), ("y", C.c_double),
                ("multiplicity", C.c_int32), ("payload_bytes", C.c_int32), ("interval", C.c_double),
                ("max_ticks", C.c_int32), ("receive", C.c_int32),
                ("jam_interval", C.c_double), ("jam_delay", C.c_double), ("jam_power_dbm", C.c_double),
                ("jam_header_bytes", C.c_int32), ("jam_payload_bytes", C.c_int32)]


class BandConfig(C.Structure):
    _fields_ = [("n_devices", C.c_int32), ("frequency_hz", C.c_double), ("bandwidth_hz", C.c_double),
                ("device", DeviceConfig * GW_MAX_DEVICES)]


class PendulumConfig(C.Structure):
    _fields_ = [("cart_mass", C.c_double), ("pendulum_mass", C.c_double), ("arm_length", C.c_double),
                ("gravity", C.c_double), ("motor_fmax", C.c_double), ("motor_kservo", C.c_double),
                ("motor_v_init", C.c_double), ("dt_max", C.c_double),
                ("kp", C.c_double), ("ki", C.c_double), ("kd", C.c_double), ("mobility", C.c_int32)]


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("n_envs", C.c_int64), ("env_id_offset", C.c_int64),
                ("n_bands", C.c_int32), ("assignment_duration_factor", C.c_int32),
                ("max_assign_duration", C.c_int32), ("mode", C.c_int32), ("seed", C.c_uint64),
                ("per_env_positions", C.c_int32), ("band", BandConfig * GW_MAX_BANDS),
                ("plant", C.c_int32), ("pendulum", PendulumConfig)]


GW_GRID_MAX_DEVICES = 24
GW_GRID_FIELD_NOW, GW_GRID_FIELD_STATS, GW_GRID_FIELD_POSITIONS, GW_GRID_FIELD_RECEIVED_POWER, GW_GRID_FIELD_FAULT = range(5)


class GridConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("n_envs", C.c_int64), ("n_devices", C.c_int32),
                ("frequency_hz", C.c_double), ("bandwidth_hz", C.c_double),
                ("power_dbm", C.c_double * GW_GRID_MAX_DEVICES), ("send_interval", C.c_double * GW_GRID_MAX_DEVICES),
                ("header_bytes", C.c_int32 * GW_GRID_MAX_DEVICES), ("payload_bytes", C.c_int32 * GW_GRID_MAX_DEVICES),
                ("move_interval", C.c_double), ("max_moves", C.c_int32)]


GW_GENBAND_MAX_SENDERS, GW_GENBAND_MAX_PHY_SENDERS = 8, 16
GW_GENBAND_MAX_DEVICES = GW_GENBAND_MAX_SENDERS + 1 + GW_GENBAND_MAX_PHY_SENDERS
(GW_GENBAND_FIELD_NOW, GW_GENBAND_FIELD_DELIVERED, GW_GENBAND_FIELD_RECEIVED, GW_GENBAND_FIELD_TRANSMISSIONS,
 GW_GENBAND_FIELD_FAULT, GW_GENBAND_FIELD_RECEIVED_POWER, GW_GENBAND_FIELD_QUEUE_LENGTH, GW_GENBAND_FIELD_COUNTER,
 GW_GENBAND_FIELD_TIES, GW_GENBAND_FIELD_RECEIVED_VALUES) = range(10)


class GenBandConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("n_envs", C.c_int64), ("n_senders", C.c_int32), ("n_phy_senders", C.c_int32),
                ("assignment_duration_factor", C.c_int32), ("max_assign_duration", C.c_int32),
                ("per_env_positions", C.c_int32), ("mode", C.c_int32), ("seed", C.c_uint64), ("env_id_offset", C.c_int64),
                ("frequency_hz", C.c_double), ("bandwidth_hz", C.c_double),
                ("multiplicity", C.c_int32 * GW_GENBAND_MAX_SENDERS), ("payload_bytes", C.c_int32 * GW_GENBAND_MAX_SENDERS),
                ("destination", C.c_int32 * GW_GENBAND_MAX_SENDERS), ("max_ticks", C.c_int32 * GW_GENBAND_MAX_SENDERS),
                ("receive", C.c_int32 * GW_GENBAND_MAX_SENDERS), ("interval", C.c_double * GW_GENBAND_MAX_SENDERS),
                ("phy_interval", C.c_double * GW_GENBAND_MAX_PHY_SENDERS), ("phy_delay", C.c_double * GW_GENBAND_MAX_PHY_SENDERS),
                ("phy_power_dbm", C.c_double * GW_GENBAND_MAX_PHY_SENDERS),
                ("phy_header_bytes", C.c_int32 * GW_GENBAND_MAX_PHY_SENDERS),
                ("phy_payload_bytes", C.c_int32 * GW_GENBAND_MAX_PHY_SENDERS)]


class NativeError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("gymwipe_b200 native error %d: %s" % (code, message))
        self.code = code


def _sources():
    return [os.path.join(CSRC, f) for f in ("gw_kernels.cu", "gw_core.cuh", "gw_pendulum.cuh", "gw_grid.cuh", "gw_band.cuh")] + [INCLUDE]


def needs_build():
    if os.environ.get("GYMWIPE_B200_LIB"):
        return False            # an explicitly chosen build (kernel-variant experiments)
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in _sources())


def build(force=False, verbose=False):
    """Compile the CUDA library for sm_100a (nvcc cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("gymwipe_b200: %s is missing and nvcc is not available to build it; "
                           "there is no CPU fallback" % LIB_PATH)
    os.makedirs(LIB_DIR, exist_ok=True)
    # several ranks (torchrun) may find the library stale at the same time: one of them compiles, into
    # a temporary file that is renamed into place, the others wait on the lock and then find it fresh
    import fcntl
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return LIB_PATH
            tmp = "%s.tmp.%d" % (LIB_PATH, os.getpid())
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
                ["-o", tmp, os.path.join(CSRC, "gw_kernels.cu")]
            try:
                subprocess.check_call(cmd)
                os.replace(tmp, LIB_PATH)
            finally:
                if os.path.exists(tmp):
                    os.unlink(tmp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


_lib = None

_VP = C.c_void_p
_SIGNATURES = {
    "gw_abi_version": (C.c_int, []),
    "gw_last_error": (C.c_char_p, []),
    "gw_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "gw_default_config": (C.c_int, [C.POINTER(Config), C.c_int64]),
    "gw_state_bytes": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_size_t)]),
    "gw_create": (C.c_int, [C.POINTER(Config), C.c_int, _VP, C.c_size_t, _VP, C.POINTER(_VP)]),
    "gw_destroy": (None, [_VP]),
    "gw_set_positions": (C.c_int, [_VP, _VP, _VP]),
    "gw_reset": (C.c_int, [_VP, _VP, C.c_int64, _VP, _VP]),
    "gw_step": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "gw_step_traced": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, C.c_int32, _VP]),
    "gw_step_host": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "gw_step_host_packed": (C.c_int, [_VP, _VP, _VP, _VP]),
    "gw_step_host_compact": (C.c_int, [_VP, _VP, _VP, _VP]),
    "gw_step_host_compact_async": (C.c_int, [_VP, _VP, _VP, _VP]),
    "gw_step_host_compact_many": (C.c_int, [C.POINTER(_VP), C.c_int32, C.POINTER(_VP), C.POINTER(_VP), _VP]),
    "gw_step_host_tiny": (C.c_int, [_VP, _VP, _VP, _VP]),
    "gw_step_host_tiny_many": (C.c_int, [C.POINTER(_VP), C.c_int32, C.POINTER(_VP), C.POINTER(_VP), _VP]),
    "gw_check": (C.c_int, [_VP, _VP]),
    "gw_stats": (C.c_int, [_VP, _VP, C.c_int, _VP]),
    "gw_share_stats": (C.c_int, [_VP, _VP]),
    "gw_debug_stamps": (C.c_int, [_VP, _VP, C.c_int64]),
    "gw_mask_bytes": (C.c_int, [_VP, C.POINTER(C.c_uint64), C.c_int, _VP]),
    "gw_read_state": (C.c_int, [_VP, C.c_int, _VP, _VP]),
    "gw_set_masks": (C.c_int, [_VP, _VP, C.c_int32, C.c_int32, _VP]),
    "gw_mask_index_count": (C.c_int, [_VP, _VP, _VP, _VP, _VP, C.c_int64, _VP]),
    "gw_fspl_attenuation": (C.c_int, [_VP, _VP, _VP, _VP, C.c_double, _VP, C.c_int64, _VP]),
    "gw_ber_bpsk": (C.c_int, [_VP, _VP, _VP, C.c_int64, _VP]),
    "gw_count_bit_errors": (C.c_int, [_VP, C.c_int32, _VP, _VP, _VP, _VP, C.c_int64, _VP]),
    "gw_philox4x32": (C.c_int, [_VP, _VP, _VP, C.c_int64, _VP]),
    "gw_grid_create": (C.c_int, [C.POINTER(GridConfig), C.c_int, _VP, _VP, _VP, _VP, _VP, C.POINTER(_VP)]),
    "gw_grid_destroy": (None, [_VP]),
    "gw_grid_run": (C.c_int, [_VP, C.c_double, _VP]),
    "gw_grid_run_traced": (C.c_int, [_VP, C.c_double, _VP, _VP, C.c_int32, _VP]),
    "gw_grid_read": (C.c_int, [_VP, C.c_int, _VP, _VP]),
    "gw_grid_check": (C.c_int, [_VP, _VP]),
    "gw_genband_create": (C.c_int, [C.POINTER(GenBandConfig), C.c_int, _VP, _VP, C.POINTER(_VP)]),
    "gw_genband_destroy": (None, [_VP]),
    "gw_genband_set_positions": (C.c_int, [_VP, _VP, _VP]),
    "gw_genband_set_movers": (C.c_int, [_VP, _VP, _VP, C.c_int32, C.c_double, _VP]),
    "gw_genband_reset": (C.c_int, [_VP, _VP, _VP]),
    "gw_genband_step": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "gw_genband_step_traced": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, C.c_int32, _VP]),
    "gw_genband_read": (C.c_int, [_VP, C.c_int, _VP, _VP]),
    "gw_genband_check": (C.c_int, [_VP, _VP]),
    "gw_policy_boltzmann": (C.c_int, [_VP, C.c_int32, C.c_int32, _VP, C.c_int64, C.c_float, C.c_double, C.c_double, C.c_double,
                                      C.c_uint64, C.c_uint64, C.c_int64, _VP, _VP, _VP, _VP, _VP]),
    "gw_max_correctable_ber": (C.c_double, [C.c_int, C.c_int]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """The loaded C-ABI library (built on first use if the in-tree .so is stale or absent)."""
    global _lib
    if _lib is None:
        if needs_build():
            build()
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        if L.gw_abi_version() != GW_ABI_VERSION:
            raise RuntimeError("gymwipe_b200: ABI version mismatch")
        _lib = L
    return _lib


def check(rc):
    if rc != GW_OK:
        raise NativeError(rc, lib().gw_last_error().decode("utf-8", "replace"))
    return rc
