"""
ORACLE TEST INFRASTRUCTURE -- not product code.

Stand-in for the pieces of ``gym==0.12.5`` (``/root/reference/Pipfile.lock:39-45``;
third-party, not vendored, not installed here) that the reference touches:
``gym.Env``, ``gym.spaces.{Discrete,Dict}``, ``gym.utils.seeding.np_random`` and
``gym.envs.registration.register`` / ``gym.make``
(``gymwipe/envs/core.py:4-7,39-42,51``; ``gymwipe/envs/__init__.py:1-14``).
"""
from gym import error, spaces, utils
from gym.core import Env
from gym.envs.registration import make, register

__version__ = "0.12.5-oracle-shim"
