"""Mirror of ``gymwipe/control``."""
from gymwipe_b200.control.inverted_pendulum import InvertedPendulumPidController

__all__ = ["InvertedPendulumPidController"]
