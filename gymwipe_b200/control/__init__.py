"""Declarative descriptors with the names of ``gymwipe/control`` (parameter holders; the control law runs inside the CUDA step kernel)."""
from gymwipe_b200.control.inverted_pendulum import InvertedPendulumPidController

__all__ = ["InvertedPendulumPidController"]
