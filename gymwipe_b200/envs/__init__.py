"""
Mirror of ``gymwipe/envs/__init__.py:1-14``: the registered env ids and a ``make`` that
instantiates them (gym itself is not a dependency).
"""
from gymwipe_b200.envs.counter_traffic import CounterTrafficEnv
from gymwipe_b200.envs.core import BaseEnv, Interpreter
from gymwipe_b200.envs.general_band import GeneralBandEnv, fits_step_kernel_template
from gymwipe_b200.envs.inverted_pendulum import InvertedPendulumEnv
from gymwipe_b200.envs.population import EnvPopulation
from gymwipe_b200.envs.sending_grid import SendingDeviceGrid

registry = {
    'CounterTraffic-v0': CounterTrafficEnv,
    'InvertedPendulum-v0': InvertedPendulumEnv,
}


def register(id, entry_point):
    registry[id] = entry_point


def make(id, **kwargs):
    """``gym.make(id)`` for the ids this package registers; keyword arguments reach the env."""
    if id not in registry:
        raise KeyError("No registered env with id: {}".format(id))
    cls = registry[id]
    # a CounterTraffic scenario beyond the step kernel's 2 senders + RRM (+ 1 PHY-only sender) template -- more MAC
    # senders, more PHY-only senders -- is stepped by the general band engine behind the same gym surface
    sc = kwargs.get("scenario")
    if cls is CounterTrafficEnv and sc is not None and len(sc["bands"]) == 1 and not fits_step_kernel_template(sc):
        return GeneralBandEnv(**kwargs)
    return cls(**kwargs)


__all__ = ["CounterTrafficEnv", "InvertedPendulumEnv", "EnvPopulation", "SendingDeviceGrid", "GeneralBandEnv", "BaseEnv", "Interpreter", "make", "register", "registry"]
