import numpy as np


def np_random(seed=None):
    """gym 0.2x: returns (numpy RandomState, seed)."""
    if seed is None:
        seed = int(np.random.SeedSequence().entropy % (2 ** 31))
    return np.random.RandomState(int(seed) % (2 ** 32)), seed
