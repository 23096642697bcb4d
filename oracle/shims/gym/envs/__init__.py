"""ORACLE TEST INFRASTRUCTURE -- gym.envs."""
from gym.envs import registration
