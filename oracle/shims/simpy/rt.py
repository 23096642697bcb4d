"""ORACLE TEST INFRASTRUCTURE -- import stub for ``simpy.rt`` (the reference only
imports the name, ``gymwipe/simtools.py:13``; wall-clock pacing is irrelevant
to the oracle, so this is the plain environment)."""
from simpy.core import Environment


class RealtimeEnvironment(Environment):
    def __init__(self, initial_time=0, factor=1.0, strict=True):
        super().__init__(initial_time)
        self.factor = factor
        self.strict = strict
