"""Minimal stand-in for ``gym==0.17.2`` (TEST INFRASTRUCTURE).

The reference imports ``gym`` (``wab_env.py:1-4``) but this image has neither gym nor gymnasium and
no network. Only the names the reference touches are provided: ``Env``, ``ObservationWrapper`` (with
attribute delegation, used at ``wab_env.py:709`` via ``self.game_options``), ``spaces.Discrete / Box /
Tuple``, ``wrappers``, ``logger`` and ``utils.seeding``. Nothing here implements game behaviour.
"""
import sys
import types


class Env:
    metadata = {}
    action_space = None
    observation_space = None

    def reset(self):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def seed(self, seed=None):
        return []

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env
        self.action_space = getattr(env, "action_space", None)
        self.observation_space = getattr(env, "observation_space", None)
        self.metadata = getattr(env, "metadata", {})

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        return self.env.step(action)


class ObservationWrapper(Wrapper):
    def reset(self, **kwargs):
        return self.observation(self.env.reset(**kwargs))

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        return self.observation(obs), reward, done, info

    def observation(self, observation):
        raise NotImplementedError


class Discrete:
    def __init__(self, n):
        self.n = int(n)

    def __eq__(self, other):
        return isinstance(other, Discrete) and other.n == self.n

    def __repr__(self):
        return "Discrete(%d)" % self.n


class Box:
    def __init__(self, low, high, shape=None, dtype=float):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    def __repr__(self):
        return "Box(%r, %r, %r)" % (self.low, self.high, self.shape)


class Tuple:
    def __init__(self, spaces):
        self.spaces = tuple(spaces)

    def __getitem__(self, i):
        return self.spaces[i]

    def __len__(self):
        return len(self.spaces)

    def __repr__(self):
        return "Tuple(%s)" % ", ".join(map(repr, self.spaces))


def install():
    """Register the stub as ``gym`` in ``sys.modules`` (idempotent). Returns the module."""
    if "gym" in sys.modules and getattr(sys.modules["gym"], "_wab_stub", False):
        return sys.modules["gym"]
    gym = types.ModuleType("gym")
    gym._wab_stub = True
    gym.Env, gym.Wrapper, gym.ObservationWrapper = Env, Wrapper, ObservationWrapper
    spaces = types.ModuleType("gym.spaces")
    spaces.Discrete, spaces.Box, spaces.Tuple = Discrete, Box, Tuple
    wrappers = types.ModuleType("gym.wrappers")
    wrappers.Monitor = lambda env, directory=None, force=False, **kw: env
    logger = types.ModuleType("gym.logger")
    logger.INFO, logger.DEBUG, logger.WARN = 20, 10, 30
    logger.set_level = lambda level: None
    utils = types.ModuleType("gym.utils")
    seeding = types.ModuleType("gym.utils.seeding")
    seeding.np_random = lambda seed=None: (None, seed)
    utils.seeding = seeding
    gym.spaces, gym.wrappers, gym.logger, gym.utils = spaces, wrappers, logger, utils
    for name, mod in (
        ("gym", gym),
        ("gym.spaces", spaces),
        ("gym.wrappers", wrappers),
        ("gym.logger", logger),
        ("gym.utils", utils),
        ("gym.utils.seeding", seeding),
    ):
        sys.modules[name] = mod
    return gym
