// gymwipe_b200 -- sliding inverted pendulum plant (config 5) for the fused step kernel.
//
// Reference: gymwipe/plants/sliding_pendulum.py:15-155 (two spheres of density 2500 and radius
// 0.05 m -- 1.309 kg each -- wagon at (0,1,0), pendulum at (0,2,0): arm 1 m; slider joint on x with a
// velocity motor ParamVel (initially 0.1) / ParamFMax = 22; hinge about z; gravity 9.81),
// gymwipe/plants/core.py:16-59 (the world is stepped by the elapsed simulated time),
// gymwipe/control/inverted_pendulum.py:16-69 (PID law, kp = 1, ki = kd = 0, sent every 10 ms),
// gymwipe/envs/inverted_pendulum.py:26-119 (obs = int(degrees(angle)), reward = |180 - degrees(angle)|).
//
// The reference delegates the dynamics to the native ODE library (py3ode, not vendored, not
// installed) and its env cannot be constructed: there is NO oracle for this
// plant -- PARITY UNPINNED.  The model below is therefore this project's own statement of the same
// mechanical system, checked against an independent high-accuracy integration of the same equations
// (tests/test_pendulum_host.py; stated tolerance after 1 s: 1e-5 rad / m with the motor active,
// 1e-9 with the motor off):
//
//   cart mass M, point mass m on a massless arm l, theta from the upward vertical (positive =
//   leaning towards -x), horizontal motor force F on the cart:
//       x''     = (F - m sin(theta) (l theta'^2 - g cos(theta))) / (M + m sin^2(theta))
//       theta'' = (g sin(theta) + x'' cos(theta)) / l
//   velocity motor: F = clamp(kServo (vTarget - x'), -fMax, +fMax)   (ODE's ParamVel / ParamFMax)
//   integrator: classical RK4, sub-steps of at most dtMax between the events that need the state.
#pragma once

#include "gw_core.cuh"

namespace gw {

struct PendulumParams {
    double M, m, l, g;          // 1.30899..., 1.30899..., 1.0, 9.81
    double fMax, kServo;        // 22 N; (M + m) / 5 ms
    double dtMax;               // 1e-3 s
    double kp, ki, kd;          // PID gains of the controller (reference: 1, 0, 0)
    double vInit;               // initial motor velocity (reference: 0.1)
    double frequency;           // band frequency (attenuation of the moving devices)
    double ctrlX, ctrlY, rrmX, rrmY, sensorY, actuatorY;   // positions; sensor / actuator x = wagon x
    int mobility;               // 1: sensor / actuator positions follow the wagon
};

struct PendulumState {
    double x, v, th, om;        // wagon position / velocity, pendulum angle / rate
    double vTarget;             // motor velocity set by the actuator
    double tPlant;              // time the state is valid for
    double ctrlAngleDeg;        // controller's latest received angle (degrees)
    double lastError;           // PID memory
};

GW_HD void pendulum_deriv(const PendulumParams &Q, double v, double th, double om, double vTarget,
                          double &ax, double &ath)
{
    double F = Q.kServo * (vTarget - v);
    F = F > Q.fMax ? Q.fMax : (F < -Q.fMax ? -Q.fMax : F);
#if defined(__CUDA_ARCH__)
    double sn, cs;
    sincos(th, &sn, &cs);               // one range reduction for both (24 % of the plant kernel's instructions were sin + cos)
#else
    const double sn = sin(th), cs = cos(th);
#endif
    ax = (F - Q.m * sn * (Q.l * om * om - Q.g * cs)) / (Q.M + Q.m * sn * sn);
    ath = (Q.g * sn + ax * cs) / Q.l;
}

GW_HD void pendulum_rk4(const PendulumParams &Q, PendulumState &S, double h)
{
    double a1, b1, a2, b2, a3, b3, a4, b4;
    pendulum_deriv(Q, S.v, S.th, S.om, S.vTarget, a1, b1);
    pendulum_deriv(Q, S.v + 0.5 * h * a1, S.th + 0.5 * h * S.om, S.om + 0.5 * h * b1, S.vTarget, a2, b2);
    pendulum_deriv(Q, S.v + 0.5 * h * a2, S.th + 0.5 * h * (S.om + 0.5 * h * b1), S.om + 0.5 * h * b2, S.vTarget, a3, b3);
    pendulum_deriv(Q, S.v + h * a3, S.th + h * (S.om + 0.5 * h * b2), S.om + h * b3, S.vTarget, a4, b4);
    const double v0 = S.v, om0 = S.om;
    S.x += h * (v0 + h * (a1 + a2 + a3) / 6.0);
    S.th += h * (om0 + h * (b1 + b2 + b3) / 6.0);
    S.v += h * (a1 + 2 * a2 + 2 * a3 + a4) / 6.0;
    S.om += h * (b1 + 2 * b2 + 2 * b3 + b4) / 6.0;
}

// OdePlant.updateState (plants/core.py:38-49): integrate up to the current simulated time.
// Out of line on the device: the transition function reaches it from three places (sensor tick, actuator
// delivery, link refresh), and three inlined copies of RK4 -- twelve sin / cos evaluations and as many divisions --
// made the plant kernel 9.5 k instructions, 252 registers and instruction-fetch bound (ncu: stall_no_instruction
// 12.7 per issued instruction).
GW_HD_COLD void pendulum_advance(const PendulumParams &Q, PendulumState &S, double now)
{
    const double dt = now - S.tPlant;
    if (!(dt > 0)) return;
    int n = (int)ceil(dt / Q.dtMax);
    if (n < 1) n = 1;
    const double h = dt / n;
    for (int i = 0; i < n; ++i) pendulum_rk4(Q, S, h);
    S.tPlant = now;
}

GW_HD void pendulum_init(const PendulumParams &Q, PendulumState &S)
{
    S.x = 0; S.v = 0; S.th = 0; S.om = 0; S.vTarget = Q.vInit; S.tPlant = 0; S.ctrlAngleDeg = 0; S.lastError = 0;
}

// Plant policy of the transition function (see NoPlant in gw_core.cuh).  Devices: 0 = AngleSensor
// (sends the angle to the controller every ms), 1 = controller (sends a velocity to the actuator
// every 10 ms), 2 = RRM, 3 = WagonActuator (receives only).  `Vals` stores the packet values.
template <class Vals, class SrxOut>
struct PendulumPlant {
    static constexpr bool active = true;
    const PendulumParams &Q;
    PendulumState &S;
    Vals vals;
    SrxOut srxOut;
    GW_HD PendulumPlant(const PendulumParams &q, PendulumState &s, Vals v, SrxOut o) : Q(q), S(s), vals(v), srxOut(o) {}

    GW_HD void put_value(int k, uint64_t tick, double v) { vals.put(k, (uint32_t)tick & (uint32_t)(kRingSlots - 1), v); }
    GW_HD double get_value(int k, uint64_t tick) { return vals.get(k, (uint32_t)tick & (uint32_t)(kRingSlots - 1)); }

    GW_HD double tick_value(int k, double now)
    {
        if (k == 0) {                       // AngleSensor._sensor (sliding_pendulum.py:131-135)
            pendulum_advance(Q, S, now);
            return S.th;
        }
        // InvertedPendulumPidController.control (control/inverted_pendulum.py:45-69), sp = 0
        const double angle = S.ctrlAngleDeg;
        const double error = angle < 0 ? -angle : angle;
        const double pid = Q.kp * error + Q.ki * (error + S.lastError) + Q.kd * (error - S.lastError);
        S.lastError = error;
        return angle < 0 ? pid : (angle > 0 ? -pid : 0.0);
    }

    GW_HD void delivered(int d, int p, double value, double now)
    {
        if (d == 0 && p == 1) {             // controller.onReceive: angle in degrees (:39-41)
            S.ctrlAngleDeg = value * (180.0 / 3.141592653589793);
        } else if (d == 1 && p == 3) {      // actuator.onReceive -> plant.setMotorVelocity (:152-153)
            pendulum_advance(Q, S, now);
            S.vTarget = value;
        }
    }

    GW_HD void position(int dev, double &px, double &py) const
    {
        if (dev == 0) { px = S.x; py = Q.sensorY; }
        else if (dev == 1) { px = Q.ctrlX; py = Q.ctrlY; }
        else if (dev == 2) { px = Q.rrmX; py = Q.rrmY; }
        else { px = S.x; py = Q.actuatorY; }
    }

    // received powers of every other device from sender d, at the start of d's transmission
    // (SimplePhy._onNewTransmission evaluates the attenuation for the current positions): written into the
    // band-sim's table through `srxOut`; `srx` is the view the transition function reads that table through
    template <class SRX>
    GW_HD void refresh_links(int d, double now, const SRX &)
    {
        if (!Q.mobility) return;
        pendulum_advance(Q, S, now);
        double dx, dy;
        position(d, dx, dy);
        for (int p = 0; p < 4; ++p) {
            if (p == d) continue;
            double px, py;
            position(p, px, py);
            srxOut.put(p * 4 + d, rx_power_mw(0.0, fspl_db(px, py, dx, dy, Q.frequency)));
        }
    }
};

}  // namespace gw
