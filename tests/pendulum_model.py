"""
Independent statement of the sliding-pendulum equations (tests only): integrated with SciPy's
DOP853 at tight tolerances, it is the analytic yardstick for the plant integrator of config 5
(there is no reference oracle for the plant).
"""
import numpy as np
from scipy.integrate import solve_ivp


def rhs(t, y, M, m, l, g, fmax, kservo, vt):
    x, v, th, om = y
    F = np.clip(kservo * (vt - v), -fmax, fmax)
    sn, cs = np.sin(th), np.cos(th)
    ax = (F - m * sn * (l * om * om - g * cs)) / (M + m * sn * sn)
    ath = (g * sn + ax * cs) / l
    return [v, ax, om, ath]


def integrate(y0, t0, t1, M, m, l, g, fmax, kservo, vt, t_eval=None):
    sol = solve_ivp(rhs, (t0, t1), y0, method="DOP853", rtol=1e-12, atol=1e-14, max_step=1e-3,
                    args=(M, m, l, g, fmax, kservo, vt), t_eval=t_eval)
    return sol.y


def default_params():
    mass = 2500.0 * 4.0 / 3.0 * np.pi * 0.05 ** 3
    return dict(M=mass, m=mass, l=1.0, g=9.81, fmax=22.0, kservo=2 * mass / 5e-3)
