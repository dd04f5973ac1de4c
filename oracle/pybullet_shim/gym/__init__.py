"""gym 0.17.3 stand-in (test infrastructure): only what the reference imports."""
from . import spaces, utils, wrappers  # noqa: F401
from .envs.registration import make, register  # noqa: F401


class Env(object):
    metadata = {'render.modes': []}
    reward_range = (-float('inf'), float('inf'))
    spec = None
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode='human'):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return

    @property
    def unwrapped(self):
        return self
