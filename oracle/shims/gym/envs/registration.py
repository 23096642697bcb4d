"""ORACLE TEST INFRASTRUCTURE -- ``register(id, entry_point)`` / ``make(id)``: without
``max_episode_steps`` gym 0.12.5 instantiates the entry point with no wrapper."""
import importlib

from gym import error

_registry = {}


def register(id, entry_point=None, **kwargs):
    _registry[id] = (entry_point, kwargs)


def make(id, **kwargs):
    if id not in _registry:
        raise error.UnregisteredEnv("No registered env with id: {}".format(id))
    entry_point, reg_kwargs = _registry[id]
    if callable(entry_point):
        cls = entry_point
    else:
        mod_name, attr_name = entry_point.split(":")
        cls = getattr(importlib.import_module(mod_name), attr_name)
    return cls(**kwargs)
