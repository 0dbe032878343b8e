"""Tier-A oracle for Environment 2.0: the UNMODIFIED reference modules
(``/root/reference/Environment 2.0/{World,Entity,Ostrich,Wolf,Bush,WAB_Environment2,WAB_Environment2_Single}.py``)
loaded with the stub ``gym`` and a keyed ``random.randint`` (TEST INFRASTRUCTURE).

The v2 code draws from Python's global ``random`` at three kinds of sites; each is keyed by identity with
the Philox contract of ``oracle/keyed_rng.py`` (sites 8-10), value = low + ((word * span) >> 32):

    site        key                                   cite
    V2_CREATE   (env, 0, entity id, axis)             WAB_Environment2.py:64-66, :82-84, :100-102
    V2_RESET    (env, episode, entity id, axis)       WAB_Environment2_Single.py:45-46 (inclusive upper bound: a reference bug kept)
    V2_PICK     (env, episode, turn, acting entity)   World.py:112, :125
"""
import importlib.util
import os
import sys
import types

import numpy as np

from . import REFERENCE_DIR, gym_stub
from .. import keyed_rng as kr

V2_DIR = os.path.join(REFERENCE_DIR, "Environment 2.0")
SITE_V2_CREATE, SITE_V2_RESET, SITE_V2_PICK = 8, 9, 10
_MODULES = None


def available():
    return os.path.isfile(os.path.join(V2_DIR, "World.py"))


def keyed_int(seed, env_id, episode, site, turn, entity, axis, low, high):
    """randint(low, high) inclusive from one keyed 32-bit word (multiply-shift range reduction)."""
    w = int(kr._draw(seed, env_id, episode, site, turn, axis, np.int64(entity), np.int64(0)))
    return int(low) + ((w * (int(high) - int(low) + 1)) >> 32)


class _KeyedRandom(types.ModuleType):
    """Stands in for the ``random`` module inside the v2 reference modules."""

    def __init__(self):
        super().__init__("random_proxy")

    def randint(self, low, high):
        f = sys._getframe(1)
        name = f.f_code.co_name
        if name == "default_game_update":                      # World.py:112, :125
            world, acting = f.f_locals["self"], int(f.f_locals["i"])
            k = world._wab
            return keyed_int(k["seed"], k["env_id"], k["episode"], SITE_V2_PICK, world._current_turn, acting, 0, low, high)
        if name == "_get_random_spawn_indices":                # WAB_Environment2_Single.py:45-46
            single = f.f_locals["self"]
            k = single.world._wab
            axis = 1 if "x" in f.f_locals else 0
            return keyed_int(k["seed"], k["env_id"], k["episode"], SITE_V2_RESET, 0, single.id, axis, low, high)
        if name in ("create_ostriches", "create_wolves", "create_bushes", "<listcomp>"):   # WAB_Environment2.py:61-110
            while "self" not in f.f_locals:
                f = f.f_back
            world = f.f_locals["self"]._world
            k = world._wab
            c = k["created"]
            k["created"] = c + 1
            return keyed_int(k["seed"], k["env_id"], 0, SITE_V2_CREATE, 0, c >> 1, c & 1, low, high)
        raise RuntimeError("unkeyed random.randint call from %r" % name)


def load():
    """Import the v2 reference modules (byte-identical sources) with ``random`` replaced."""
    global _MODULES
    if _MODULES is not None:
        return _MODULES
    if not available():
        raise FileNotFoundError(V2_DIR)
    gym_stub.install()
    names = ["Entity", "Bush", "Ostrich", "Wolf", "World", "WAB_Environment2_Single", "WAB_Environment2"]
    saved = {n: sys.modules.get(n) for n in names}
    mods = {}
    sys.path.insert(0, V2_DIR)
    try:
        for n in names:
            sys.modules.pop(n, None)
        for n in names:
            spec = importlib.util.spec_from_file_location(n, os.path.join(V2_DIR, n + ".py"))
            m = importlib.util.module_from_spec(spec)
            sys.modules[n] = m
            spec.loader.exec_module(m)
            mods[n] = m
    finally:
        sys.path.remove(V2_DIR)
        for n in names:
            if saved[n] is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = saved[n]
    proxy = _KeyedRandom()
    for n in ("World", "WAB_Environment2", "WAB_Environment2_Single"):
        mods[n].random = proxy
    _MODULES = mods
    return mods


def make_env(width, height, n_ostriches, n_wolves, n_bushes, game_options=None, seed=0, env_id=0):
    """A reference ``WAB_Environment2`` whose draws are keyed; ``reset_environment`` bumps the episode."""
    mods = load()
    base = mods["WAB_Environment2"].WAB_Environment2
    opts = dict(mods["WAB_Environment2"].default_game_options)
    if game_options:
        opts.update(game_options)

    class KeyedEnvironment2(base):
        def __init__(self):
            super().__init__(width, height, opts)
            self._world._wab = {"seed": int(seed), "env_id": int(env_id), "episode": 0, "created": 0}

        def reset_environment(self):
            self._world._wab["episode"] += 1
            return super().reset_environment()

    env = KeyedEnvironment2()
    env.create_ostriches(n_ostriches)
    env.create_wolves(n_wolves)
    env.create_bushes(n_bushes)
    return env


def hidden_state(env):
    """Per-entity (type, obj x, obj y, table X, table Y, visible, food, role_or_running, status)."""
    out = []
    df = env._world._entities
    for i in range(len(df)):
        row = df.iloc[i]
        obj = row["Entity_Object"]
        t = row["Type"]
        extra = int(obj.role) if t == "Ostrich" else (int(obj.is_running) if t == "Wolf" else int(obj.has_food))
        status = int(getattr(obj, "status", 0))
        out.append((t, int(obj.x), int(obj.y), int(row["X"]), int(row["Y"]), bool(row["Visible"]), float(obj.food), extra, status))
    return out
