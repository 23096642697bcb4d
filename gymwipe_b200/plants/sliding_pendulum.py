"""
Declarative descriptor (a parameter holder, by design: SURVEY.md section 8b) with the names of ``gymwipe/plants/sliding_pendulum.py`` (and ``plants/core.py``): the sliding inverted
pendulum as a parameter set.  The reference builds an ODE world (py3ode) of two spheres, a slider
joint with a velocity motor and a hinge (``sliding_pendulum.py:24-55``); here the same mechanical
system is integrated by the step kernel (``gymwipe_b200/csrc/gw_pendulum.cuh``).
"""
from math import pi


class SlidingPendulum:
    """Wagon + pendulum on a motorised slider (``sliding_pendulum.py:15-114``)."""

    SPHERE_DENSITY = 2500.0         # ode.Mass.setSphere(2500, 0.05), :27,:34
    SPHERE_RADIUS = 0.05
    ARM_LENGTH = 1.0                # wagon at (0,1,0), pendulum at (0,2,0), :29,:36
    GRAVITY = 9.81                  # plants/core.py:34
    MOTOR_FMAX = 22.0               # slider.setParam(ode.ParamFMax, 22), :53
    MOTOR_INITIAL_VELOCITY = 0.1    # slider.setParam(ode.ParamVel, 0.1), :52

    def __init__(self, motor_time_constant=5e-3, max_step=1e-3):
        mass = self.SPHERE_DENSITY * 4.0 / 3.0 * pi * self.SPHERE_RADIUS ** 3
        self.cart_mass = mass
        self.pendulum_mass = mass
        self.arm_length = self.ARM_LENGTH
        self.gravity = self.GRAVITY
        self.motor_fmax = self.MOTOR_FMAX
        self.motor_kservo = (self.cart_mass + self.pendulum_mass) / motor_time_constant
        self.motor_v_init = self.MOTOR_INITIAL_VELOCITY
        self.maxStepSize = max_step


class AngleSensor:
    """``sliding_pendulum.py:116-135``: samples the angle every ``sampleInterval`` and sends it."""

    def __init__(self, sampleInterval=0.001, payloadBytes=8):
        self.sampleInterval = sampleInterval
        self.payloadBytes = payloadBytes


class WagonActuator:
    """``sliding_pendulum.py:137-155``: sets the motor velocity to the received value."""
