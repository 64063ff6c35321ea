import numpy as np


class Discrete(object):
    def __init__(self, n):
        self.n = n
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def sample(self):
        return np.random.randint(self.n)

    def contains(self, x):
        return 0 <= int(x) < self.n
