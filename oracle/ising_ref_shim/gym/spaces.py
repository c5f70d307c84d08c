"""TEST INFRASTRUCTURE: the two gym.spaces classes the reference Ising env constructs (only `.n` is read)."""


class Discrete(object):
    def __init__(self, n):
        self.n = n


class MultiBinary(object):
    def __init__(self, n):
        self.n = n
