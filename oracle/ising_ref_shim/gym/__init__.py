"""TEST INFRASTRUCTURE: minimal stand-in for gym 0.9.2 (absent here), just enough for the reference's
examples/ising_model/multiagent/environment.py to import and run unmodified: gym.Env forwards the public
step/reset to the subclass's _step/_reset, as gym < 0.10 did."""
from . import spaces  # noqa: F401


class Env(object):
    def step(self, action):
        return self._step(action)

    def reset(self):
        return self._reset()
