"""
Batched networked inverted-pendulum env (config 5), mirror of
``gymwipe/envs/inverted_pendulum.py:58-119``: an agent assigns the frequency band to the angle
sensor (device 0) or the controller (device 1); the sensor's packets carry the pendulum angle to the
controller, the controller's packets carry a motor velocity to the actuator on the wagon.

The reference env cannot be constructed (``SimMan.env`` setter recursion, ``simtools.py:39-42``) and
delegates the dynamics to the un-vendored ODE library, so there is NO oracle for this path: PARITY
UNPINNED.  This env keeps the reference's structure and constants and states its own model
(``gymwipe_b200/csrc/gw_pendulum.cuh``); deviations from the reference source are listed in
DESIGN.md section 10 (packet sizes are fixed at 8 bytes instead of ``byteSize = angle``; the
controller starts at t = 0 instead of after 1 s and always sends a command; devices are in receive
mode; attenuation follows the wagon at transmission start, not mid-packet).

Network behaviour (who transmits when, what is decoded) uses exactly the transition function of
``CounterTrafficEnv`` and is checked against the oracle with the equivalent traffic scenario.
"""
from math import degrees

import torch

from gymwipe_b200 import _native as N
from gymwipe_b200 import spaces
from gymwipe_b200.control.inverted_pendulum import InvertedPendulumPidController
from gymwipe_b200.envs.counter_traffic import CounterTrafficEnv
from gymwipe_b200.plants.sliding_pendulum import AngleSensor, SlidingPendulum


def pendulum_scenario(sensor, controller):
    """Devices of ``InvertedPendulumEnv.__init__`` (``inverted_pendulum.py:68-96``) as a scenario dict."""
    return {"assignment_duration_factor": 1000, "bands": [{"frequency": 2.4e9, "bandwidth": 22e6, "devices": [
        # AngleSensor at (wagon x, 0) -> controller
        {"role": "sender", "x": 0.0, "y": 0.0, "mult": 1, "payload": sensor.payloadBytes,
         "interval": sensor.sampleInterval, "dest": 1},
        # InvertedPendulumPidController at (0, -1) -> actuator
        {"role": "sender", "x": 0.0, "y": -1.0, "mult": 1, "payload": controller.payloadBytes,
         "interval": controller.controlInterval, "dest": 0},
        {"role": "rrm", "x": 0.0, "y": 1.0},
        # WagonActuator at (wagon x, 0): receives only (a PHY that never sends)
        {"role": "jammer", "x": 0.0, "y": 0.0, "interval": 1e30, "delay": 0.0, "power": 0.0, "hdr": 13, "payload": 1},
    ]}]}


class InvertedPendulumEnv(CounterTrafficEnv):
    """See module docstring.  ``step`` returns ``obs = int(degrees(angle))``, ``reward = |180 - degrees(angle)|``."""

    def __init__(self, num_envs=1, device="cuda", plant=None, controller=None, sensor=None, mobility=True,
                 strict=None):
        self.plant = plant or SlidingPendulum()
        self.controller = controller or InvertedPendulumPidController()
        self.sensor = sensor or AngleSensor()
        self._mobility = bool(mobility)
        super().__init__(num_envs=num_envs, device=device, mode="reference",
                         scenario=pendulum_scenario(self.sensor, self.controller), strict=strict)
        # Observation depends on plant angle (inverted_pendulum.py:72-73)
        self.observation_space = spaces.Discrete(180)

    def _needs_per_env_tables(self):
        return True

    def _configure_native(self, cfg):
        cfg.plant = N.GW_PLANT_SLIDING_PENDULUM
        p, c = self.plant, self.controller
        pc = cfg.pendulum
        pc.cart_mass, pc.pendulum_mass, pc.arm_length, pc.gravity = p.cart_mass, p.pendulum_mass, p.arm_length, p.gravity
        pc.motor_fmax, pc.motor_kservo, pc.motor_v_init = p.motor_fmax, p.motor_kservo, p.motor_v_init
        pc.dt_max = p.maxStepSize
        pc.kp, pc.ki, pc.kd = c.kp, c.ki, c.kd
        pc.mobility = 1 if self._mobility else 0

    def plant_state(self):
        """float64 ``[8, num_envs]``: x, v, theta, omega, motor target velocity, plant time,
        controller's angle estimate (deg), PID memory."""
        return self.read_state(N.GW_FIELD_PLANT)

    def reset(self, env_ids=None):
        """``inverted_pendulum.py:98-102``: returns the observation; nothing is reset."""
        th = self.plant_state()[2]
        obs = torch.rad2deg(th).to(torch.int64)
        return int(obs[0]) if self._scalar_api else obs

    def step(self, action):
        out = super().step(action)
        if self._scalar_api and isinstance(out[0], int):
            th = float(self.plant_state()[2, 0])
            return out[0], out[1], out[2], {"Sensor angle": degrees(th)}
        return out
