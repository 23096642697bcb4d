"""ORACLE TEST INFRASTRUCTURE -- ``simpy==3.0.11``-compatible engine (see core.py)."""
from simpy.core import Environment, Infinity, EmptySchedule, StopSimulation
from simpy.events import (Event, Timeout, Process, Initialize, Condition,
                          AllOf, AnyOf, Interrupt, NORMAL, URGENT, PENDING)
from simpy import rt

__version__ = "3.0.11-oracle-shim"
__all__ = ["Environment", "Event", "Timeout", "Process", "AllOf", "AnyOf",
           "Interrupt", "rt"]
