"""Declarative descriptors with the names of ``gymwipe/plants``: parameter holders -- the dynamics run inside the CUDA step kernel."""
from gymwipe_b200.plants.sliding_pendulum import AngleSensor, SlidingPendulum, WagonActuator

__all__ = ["SlidingPendulum", "AngleSensor", "WagonActuator"]
