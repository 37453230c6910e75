"""register / make with the TimeLimit wrapper of max_episode_steps (R/envs/__init__.py:3-24)."""
import importlib

from ..core import Wrapper

registry = {}


class EnvSpec:
    def __init__(self, id, entry_point, max_episode_steps=None, kwargs=None):
        self.id, self.entry_point, self.max_episode_steps, self.kwargs = id, entry_point, max_episode_steps, dict(kwargs or {})


class TimeLimit(Wrapper):
    def __init__(self, env, max_episode_steps):
        super().__init__(env)
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = 0

    def reset(self, **kw):
        self._elapsed_steps = 0
        return self.env.reset(**kw)

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            info["TimeLimit.truncated"] = not done
            done = True
        return obs, reward, done, info


def register(id, entry_point, max_episode_steps=None, kwargs=None, **_):
    registry[id] = EnvSpec(id, entry_point, max_episode_steps, kwargs)


def spec(id):
    return registry[id]


def make(id, **kwargs):
    s = registry[id]
    mod, cls = s.entry_point.split(":")
    env = getattr(importlib.import_module(mod), cls)(**{**s.kwargs, **kwargs})
    env.spec = s
    if s.max_episode_steps is not None:
        env = TimeLimit(env, s.max_episode_steps)
    return env
