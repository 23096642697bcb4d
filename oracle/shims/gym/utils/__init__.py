"""ORACLE TEST INFRASTRUCTURE -- gym.utils."""
from gym.utils import seeding
