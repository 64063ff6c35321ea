import numpy as np


def np_random(seed=None):
    # gym 0.12.5 hashes the seed; the reference never reads self.np_random
    # (track_1v1.py:129-132), so only the (rng, seed) return shape matters.
    rng = np.random.RandomState()
    if seed is not None:
        rng.seed(int(seed) % (2 ** 32))
    return rng, seed
