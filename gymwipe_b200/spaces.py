"""
The two gym spaces the reference's env surface uses (``gym.spaces.Discrete`` / ``Dict``,
gym 0.12.5; ``gymwipe/envs/core.py:39-42``, ``gymwipe/envs/counter_traffic.py:120``), so that
``action_space`` / ``observation_space`` keep their meaning without a gym dependency
(gym is not installable in the target image).  Batched ``contains`` accepts tensors.
"""
from collections import OrderedDict

import numpy as np
import torch


class Space:
    def __contains__(self, x):
        return self.contains(x)


class Discrete(Space):
    def __init__(self, n):
        assert n >= 0
        self.n = int(n)
        self.shape = ()
        self.dtype = np.int64
        self._rng = np.random.RandomState()

    def seed(self, seed=None):
        self._rng.seed(seed)

    def sample(self):
        return int(self._rng.randint(self.n))

    def contains(self, x):
        if isinstance(x, bool):
            return False
        if isinstance(x, int):
            return 0 <= x < self.n
        if isinstance(x, (np.generic, np.ndarray)):
            if x.dtype.kind not in "iu":
                return False
            return bool(np.all((x >= 0) & (x < self.n)))
        if torch.is_tensor(x):
            if x.is_floating_point() or x.dtype == torch.bool:
                return False
            return bool(((x >= 0) & (x < self.n)).all())
        return False

    def __repr__(self):
        return "Discrete(%d)" % self.n

    def __eq__(self, other):
        return isinstance(other, Discrete) and self.n == other.n


class Dict(Space):
    def __init__(self, spaces):
        if isinstance(spaces, dict) and not isinstance(spaces, OrderedDict):
            spaces = OrderedDict(sorted(spaces.items()))
        self.spaces = OrderedDict(spaces)

    def seed(self, seed=None):
        for s in self.spaces.values():
            s.seed(seed)

    def sample(self):
        return OrderedDict((k, s.sample()) for k, s in self.spaces.items())

    def contains(self, x):
        if not isinstance(x, dict) or len(x) != len(self.spaces):
            return False
        return all(k in x and s.contains(x[k]) for k, s in self.spaces.items())

    def __repr__(self):
        return "Dict(" + ", ".join("%s:%r" % kv for kv in self.spaces.items()) + ")"
