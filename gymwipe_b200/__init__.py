"""
gymwipe_b200 -- B200-native batched simulator for Gym-WiPE's per-step wireless hot path.

    import gymwipe_b200
    env = gymwipe_b200.make('CounterTraffic-v0', num_envs=65536)
    obs = env.reset()
    obs, reward, done, info = env.step({"device": dev, "duration": dur})   # int32 CUDA tensors

Importing the package does not need a GPU; constructing an env does (no CPU fallback).
"""
from gymwipe_b200 import _native
from gymwipe_b200.envs import CounterTrafficEnv, make, register

__version__ = "0.1.0"


def build(force=False, verbose=False):
    """Compile the in-tree CUDA library for sm_100a."""
    return _native.build(force=force, verbose=verbose)


__all__ = ["CounterTrafficEnv", "make", "register", "build", "__version__"]
