import importlib

registry = {}


class TimeLimit(object):
    """gym 0.12.5 wrappers/time_limit.py semantics: done=True once elapsed >= max_episode_steps;
    the counter restarts on reset()."""

    def __init__(self, env, max_episode_steps=None):
        self.env = env
        self.action_space = env.action_space
        self.observation_space = env.observation_space
        self.metadata = getattr(env, 'metadata', {})
        self.reward_range = getattr(env, 'reward_range', None)
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = 0
        self._started = False

    def step(self, action):
        assert self._started, "Cannot call env.step() before calling reset()"
        observation, reward, done, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._max_episode_steps is not None and self._max_episode_steps <= self._elapsed_steps:
            done = True
        return observation, reward, done, info

    def reset(self, **kwargs):
        self._started = True
        self._elapsed_steps = 0
        return self.env.reset(**kwargs)

    def render(self, mode='human', **kwargs):
        return self.env.render(mode, **kwargs)

    def close(self):
        return self.env.close()

    def seed(self, seed=None):
        return self.env.seed(seed)

    @property
    def unwrapped(self):
        return self.env.unwrapped


def register(id, entry_point=None, kwargs=None, max_episode_steps=None, **_):
    registry[id] = dict(entry_point=entry_point, kwargs=kwargs or {}, max_episode_steps=max_episode_steps)


def make(id, **extra):
    spec = registry[id]
    mod_name, attr = spec['entry_point'].split(':')
    cls = getattr(importlib.import_module(mod_name), attr)
    kw = dict(spec['kwargs'])
    kw.update(extra)
    env = cls(**kw)
    if spec['max_episode_steps'] is not None:
        env = TimeLimit(env, max_episode_steps=spec['max_episode_steps'])
    return env
