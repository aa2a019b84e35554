"""Space stubs: only constructor arguments and the attributes the reference reads."""
import numpy as np


class Space:
    pass


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        self.n = int(n)
        self.start = start
        self.shape = ()
        self.dtype = np.int64


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64, seed=None):
        self.nvec = np.asarray(nvec, dtype=dtype)
        self.shape = self.nvec.shape
        self.dtype = dtype


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        self.shape = tuple(shape)
        self.low = low
        self.high = high
        self.dtype = np.dtype(dtype)


class Dict(Space):
    def __init__(self, spaces=None, seed=None):
        self.spaces = dict(spaces or {})

    def __getitem__(self, k):
        return self.spaces[k]
