import importlib

from ..wrappers import TimeLimit


class EnvSpec(object):
    def __init__(self, id, entry_point, kwargs=None, max_episode_steps=None):
        self.id, self.entry_point, self.kwargs, self.max_episode_steps = id, entry_point, kwargs or {}, max_episode_steps


class EnvRegistry(object):
    def __init__(self):
        self.env_specs = {}


registry = EnvRegistry()


def register(id, entry_point=None, kwargs=None, max_episode_steps=None, **_):
    registry.env_specs[id] = EnvSpec(id, entry_point, kwargs, max_episode_steps)


def make(id, **kwargs):
    spec = registry.env_specs[id]
    mod_name, attr = spec.entry_point.split(":")
    cls = getattr(importlib.import_module(mod_name), attr)
    kw = dict(spec.kwargs)
    kw.update(kwargs)
    env = cls(**kw)
    env.spec = spec
    if spec.max_episode_steps is not None:
        env = TimeLimit(env, max_episode_steps=spec.max_episode_steps)
    return env
