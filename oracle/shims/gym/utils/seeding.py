"""ORACLE TEST INFRASTRUCTURE -- ``gym.utils.seeding.np_random`` -> (RandomState, seed)."""
import os
import struct

import numpy as np


def np_random(seed=None):
    if seed is not None and not (isinstance(seed, int) and 0 <= seed):
        raise ValueError('Seed must be a non-negative integer or omitted, not {}'.format(seed))
    if seed is None:
        seed = struct.unpack("<Q", os.urandom(8))[0] % (2 ** 32)
    rng = np.random.RandomState()
    rng.seed(seed % (2 ** 32))
    return rng, seed
