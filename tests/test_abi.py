"""
CPU suite, part 3: the C-ABI library loads and exports every symbol include/gymwipe_b200.h
declares; host-side logic (scenario validation, error reporting, spaces) works without a GPU;
the product refuses to run without CUDA (no CPU fallback).
"""
import ctypes as C
import os
import re

import pytest

from gymwipe_b200 import _native as N
from gymwipe_b200 import scenario as S
from gymwipe_b200 import spaces

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gymwipe_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gw_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = N.lib()
    names = declared_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(L, name), name
    assert sorted(N.EXPORTED_SYMBOLS) == names       # the ctypes table covers the header
    assert L.gw_abi_version() == N.GW_ABI_VERSION


def test_cubin_is_sm_100a():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", N.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_default_config_and_state_size():
    L = N.lib()
    cfg = N.Config()
    assert L.gw_default_config(C.byref(cfg), 65536) == 0
    assert cfg.n_bands == 1 and cfg.band[0].n_devices == 3
    assert cfg.band[0].device[1].multiplicity == 3 and cfg.band[0].device[2].role == N.GW_ROLE_RRM
    n = C.c_size_t()
    assert L.gw_state_bytes(C.byref(cfg), C.byref(n)) == 0
    assert 1000 * 65536 < n.value < 2000 * 65536


def test_config_validation_errors():
    L = N.lib()
    cfg = S.config_from_dict(S.default_scenario_dict(), 16)
    n = C.c_size_t()
    cfg.n_bands = 3
    assert L.gw_state_bytes(C.byref(cfg), C.byref(n)) == N.GW_E_INVALID
    assert b"n_bands" in L.gw_last_error()
    cfg = S.config_from_dict(S.default_scenario_dict(), 16)
    cfg.band[0].device[0].role = N.GW_ROLE_RRM
    assert L.gw_state_bytes(C.byref(cfg), C.byref(n)) == N.GW_E_INVALID
    cfg = S.config_from_dict(S.default_scenario_dict(), 0)
    assert L.gw_state_bytes(C.byref(cfg), C.byref(n)) == N.GW_E_INVALID
    cfg = S.config_from_dict(S.default_scenario_dict(), 4)
    cfg.abi_version = 99
    assert L.gw_state_bytes(C.byref(cfg), C.byref(n)) == N.GW_E_INVALID


def test_max_correctable_ber_host():
    L = N.lib()
    assert L.gw_max_correctable_ber(3, 4) == 0.25
    assert L.gw_max_correctable_ber(1, 2) == 0.5
    assert L.gw_max_correctable_ber(7, 8) == 0.125


def test_spaces_follow_gym_semantics():
    import numpy as np
    space = spaces.Dict({"device": spaces.Discrete(2), "duration": spaces.Discrete(20)})
    assert space.contains({"device": 0, "duration": 19})
    assert space.contains({"device": np.int64(1), "duration": np.int32(3)})
    assert not space.contains({"device": 2, "duration": 3})
    assert not space.contains({"device": 0, "duration": 20})
    assert not space.contains({"device": 0})
    assert not space.contains({"device": 0.0, "duration": 1})
    assert not space.contains([0, 1])


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import gymwipe_b200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gymwipe_b200.make('CounterTraffic-v0')
    # the ABI itself also refuses
    L = N.lib()
    cfg = S.config_from_dict(S.default_scenario_dict(), 4)
    h = C.c_void_p()
    assert L.gw_create(C.byref(cfg), 0, None, 0, None, C.byref(h)) == N.GW_E_CUDA


def test_product_does_not_import_oracle():
    """The package must not reference oracle/ or tests/hostsim (parity claims depend on it)."""
    pkg = os.path.join(ROOT, "gymwipe_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "gw_oracle" not in text and "hostsim" not in text.replace("tests/hostsim", ""), f


def test_compact_result_word_layout():
    """The packed result word of gw_step_host_compact (include/gymwipe_b200.h: GW_COMPACT_*) and its
    Python decoder agree: obs in bits 0..16, reward + 16 in bits 17..21, done in bit 22."""
    import numpy as np
    import torch
    from gymwipe_b200.envs.counter_traffic import CounterTrafficEnv
    obs = np.array([0, 65534, 65536, 65538, 131071], np.int64)
    rew = np.array([-10, -2, 0, 2, 10], np.int64)
    done = np.array([0, 1, 0, 1, 1], np.int64)
    words = (obs | ((rew + 16) << 17) | (done << 22)).astype(np.uint32).view(np.int32)
    o, r, d = CounterTrafficEnv.unpack_compact(torch.from_numpy(words.copy()))
    assert o.tolist() == obs.tolist() and r.tolist() == [float(x) for x in rew] and d.tolist() == [bool(x) for x in done]
    text = open(os.path.join(ROOT, "include", "gymwipe_b200.h")).read()
    assert "#define GW_COMPACT_OBS(w)    ((int32_t)((w) & 0x1FFFFu))" in text
    assert "#define GW_COMPACT_REWARD(w) ((int32_t)(((w) >> 17) & 31u) - 16)" in text
    assert "#define GW_COMPACT_DONE(w)   ((int32_t)(((w) >> 22) & 1u))" in text


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (the CPU arm the driver times next to the GPU arm) prints one JSON line
    with the contract's keys and a plausible value even for a handful of steps."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "4",
                          "--warmup", "3"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["steps"] == 4 and line["unit"] == "env-steps/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert 1e4 < line["value"] < 1e9                  # a CPU rate, not a timing artefact
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0


def test_ctypes_structs_mirror_the_header_layout(tmp_path):
    """include/gymwipe_b200.h compiled as plain C (it is the boundary a C caller sees): sizeof and field offsets of
    gw_config, gw_grid_config and gw_genband_config equal those of the ctypes mirrors in gymwipe_b200/_native.py."""
    import subprocess
    from gymwipe_b200 import _native as N
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "gymwipe_b200.h"
int main(void) {
    printf("gw_config %zu %zu %zu %zu\n", sizeof(gw_config), offsetof(gw_config, seed), offsetof(gw_config, band), offsetof(gw_config, plant));
    printf("gw_grid_config %zu %zu %zu %zu\n", sizeof(gw_grid_config), offsetof(gw_grid_config, power_dbm), offsetof(gw_grid_config, header_bytes), offsetof(gw_grid_config, max_moves));
    printf("gw_genband_config %zu %zu %zu %zu %zu %zu\n", sizeof(gw_genband_config), offsetof(gw_genband_config, seed),
           offsetof(gw_genband_config, frequency_hz), offsetof(gw_genband_config, multiplicity), offsetof(gw_genband_config, interval),
           offsetof(gw_genband_config, phy_payload_bytes));
    return 0;
}
''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = {l.split()[0]: [int(x) for x in l.split()[1:]] for l in subprocess.check_output([str(exe)], text=True).splitlines()}
    C_ = N.Config
    assert got["gw_config"] == [C.sizeof(C_), C_.seed.offset, C_.band.offset, C_.plant.offset]
    G = N.GridConfig
    assert got["gw_grid_config"] == [C.sizeof(G), G.power_dbm.offset, G.header_bytes.offset, G.max_moves.offset]
    B = N.GenBandConfig
    assert got["gw_genband_config"] == [C.sizeof(B), B.seed.offset, B.frequency_hz.offset, B.multiplicity.offset, B.interval.offset,
                                        B.phy_payload_bytes.offset]
