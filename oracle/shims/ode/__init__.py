"""ORACLE TEST INFRASTRUCTURE -- import stub for py3ode (``gymwipe/plants/core.py:5``,
``gymwipe/plants/sliding_pendulum.py:7``): only needed so that ``import gymwipe.envs``
succeeds; the pendulum plant has no oracle."""

environment = None
ParamVel = 0
ParamFMax = 1


class World:
    def setGravity(self, g):
        self.gravity = g

    def step(self, dt):
        raise NotImplementedError("ode stub: the pendulum plant has no oracle")


class Body:
    def __init__(self, world):
        raise NotImplementedError("ode stub")


class Mass:
    pass


class SliderJoint:
    def __init__(self, world):
        raise NotImplementedError("ode stub")


class HingeJoint:
    def __init__(self, world):
        raise NotImplementedError("ode stub")
