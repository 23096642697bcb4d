"""
CPU suite, part 6: the plant integrator of config 5 (gymwipe_b200/csrc/gw_pendulum.cuh, compiled
for the host) against an independent DOP853 integration of the same equations.
STATED TOLERANCE after 1 s of simulated time with <= 1 ms RK4 sub-steps: theta and x within 1e-5
(rad, m) with the velocity motor active (its force clamp is a kink that costs RK4 an order near the
switching instants; observed ~1e-6), within 1e-9 with the motor off.  (No reference oracle exists
for the plant: parity unpinned.)
"""
import ctypes as C

import numpy as np
import pytest

import hostsim as HS
from pendulum_model import default_params, integrate

TOL = 1e-5
TOL_FREE = 1e-9


@pytest.mark.parametrize("th0,vt,fmax", [(0.05, 0.1, 22.0), (-0.2, -0.5, 22.0), (0.3, 0.0, 0.0), (1.0, 2.0, 5.0)])
def test_rk4_plant_matches_dop853(th0, vt, fmax):
    P = default_params()
    P["fmax"] = fmax
    L = HS.lib()
    params = np.array([P["M"], P["m"], P["l"], P["g"], P["fmax"], P["kservo"], 1e-3], np.float64)
    state = np.array([0.0, 0.0, th0, 0.0, vt, 0.0, 0.0, 0.0], np.float64)
    # advance in irregular pieces, as the event loop does (ticks, deliveries)
    rs = np.random.RandomState(1)
    t = 0.0
    while t < 1.0:
        t = min(1.0, t + rs.uniform(1e-4, 3e-3))
        L.hs_pendulum_advance(params.ctypes.data_as(C.c_void_p), state.ctypes.data_as(C.c_void_p), t)
    want = integrate([0.0, 0.0, th0, 0.0], 0.0, 1.0, P["M"], P["m"], P["l"], P["g"], P["fmax"], P["kservo"], vt)[:, -1]
    tol = TOL_FREE if fmax == 0.0 else TOL
    assert abs(state[0] - want[0]) < tol and abs(state[2] - want[2]) < tol
    assert abs(state[1] - want[1]) < 10 * tol and abs(state[3] - want[3]) < 10 * tol


def test_free_pendulum_energy():
    """Motor off (fMax = 0): total mechanical energy is conserved to 1e-9 relative over 2 s."""
    P = default_params()
    L = HS.lib()
    params = np.array([P["M"], P["m"], P["l"], P["g"], 0.0, P["kservo"], 1e-3], np.float64)
    state = np.array([0.0, 0.0, 0.4, 0.0, 0.0, 0.0, 0.0, 0.0], np.float64)

    def energy(s):
        x, v, th, om = s[:4]
        # pendulum at (x - l sin th, l cos th)
        vx = v - P["l"] * np.cos(th) * om
        vy = -P["l"] * np.sin(th) * om
        return 0.5 * P["M"] * v * v + 0.5 * P["m"] * (vx * vx + vy * vy) + P["m"] * P["g"] * P["l"] * np.cos(th)
    e0 = energy(state)
    for k in range(1, 2001):
        L.hs_pendulum_advance(params.ctypes.data_as(C.c_void_p), state.ctypes.data_as(C.c_void_p), k * 1e-3)
    assert abs(energy(state) - e0) < 1e-9 * abs(e0)
