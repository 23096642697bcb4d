"""
Mirror of ``gymwipe/envs/core.py``: ``BaseEnv`` (action space = device x assignment duration)
and the ``Interpreter`` interface.  In the reference the interpreter is a Python object called
back for every packet the RRM decodes; here ``CounterTrafficInterpreter``'s arithmetic is fused
into the step kernel and the object is a read-back view.
"""
from abc import ABC, abstractmethod

import numpy as np

from gymwipe_b200 import spaces


class BaseEnv:
    """``envs/core.py:14-57``."""

    metadata = {'render.modes': ['human']}

    MAX_ASSIGN_DURATION = 20  # * ASSIGNMENT_DURATION_FACTOR time slots

    ASSIGNMENT_DURATION_FACTOR = 1000

    def __init__(self, frequencyBand, deviceCount):
        self.frequencyBand = frequencyBand
        self.deviceCount = deviceCount
        self.action_space = spaces.Dict({
            "device": spaces.Discrete(deviceCount),
            "duration": spaces.Discrete(self.MAX_ASSIGN_DURATION),
        })
        self.seed()

    def seed(self, seed=None):
        """``envs/core.py:46-52``: stores an RNG nothing on the hot path consumes; returns ``[seed]``."""
        if seed is None:
            seed = int(np.random.SeedSequence().entropy % (2 ** 32))
        self.np_random = np.random.RandomState(seed % (2 ** 32))
        return [seed]

    def render(self, mode='human', close=False):
        """Renders the environment to stdout."""

    def close(self):
        pass


class Interpreter(ABC):
    """``envs/core.py:59-159``."""

    @abstractmethod
    def onPacketReceived(self, senderIndex, receiverIndex, payload):
        """Invoked in the reference whenever the RRM receives a packet."""

    def onFrequencyBandAssignment(self, deviceIndex, duration):
        """Invoked in the reference whenever the RRM assigns the frequency band."""

    @abstractmethod
    def getReward(self):
        """Reward that depends on the last channel assignment."""

    @abstractmethod
    def getObservation(self):
        """Observation of the system's state."""

    def getDone(self):
        return False

    def getInfo(self):
        return {}

    def getFeedback(self):
        return self.getObservation(), self.getReward(), self.getDone(), self.getInfo()

    def reset(self):
        """Invoked when the environment is reset."""
