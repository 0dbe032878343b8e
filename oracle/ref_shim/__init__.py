"""Tier-A oracle: the UNMODIFIED reference ``wab_env.py`` run under outside-in shims
(TEST INFRASTRUCTURE — never imported by the product path).

The reference pins python 3.8 / gym 0.17.2 / pandas 1.1.2 / numpy 1.19.2 (``Pipfile.lock``); this
image has python 3.12 / pandas 3 / numpy 2 and no gym. The source file is loaded byte-identical from
``/root/reference`` (never copied into the repo) into a private module whose globals ``np`` and ``pd``
are replaced, after import, by two proxies:

* ``pd`` proxy (``_PdProxy``): ``DataFrame`` returns a subclass restoring three pandas-1.x behaviours
  the reference relies on — ``DataFrame.append`` (``wab_env.py:570, 587, 601, 629``), positional axis
  in ``drop(labels, 1)`` (``:58``), and write-through of ``df.col.clip(..., inplace=True)`` (``:310``;
  under pandas-3 copy-on-write it is a silent no-op and food would exceed 1, contradicting the
  reference's own ``observation_space`` ``:223-225``) — plus tolerance for ``dtype=int`` on the action
  table with ``None`` roles (``:150-182``).
* ``np`` proxy (``_NpProxy``): ``np.random.random`` / ``randint`` return *keyed* draws
  (``oracle/keyed_rng.py``) recovered by inspecting the calling frame, so the draw for a cell or
  wolf does not depend on CPython set-iteration order.

``load_reference()`` returns the module; ``make_env()`` builds a bookkeeping subclass instance that
only records (seed, env id, episode) for the key — no game logic is overridden.
"""
import os
import sys
import types

import numpy as _np
import pandas as _pd

from . import gym_stub
from .. import keyed_rng as kr

REFERENCE_DIR = os.environ.get("WAB_REFERENCE_DIR", "/root/reference")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "wab_env.py"))


# --------------------------------------------------------------------------- pandas compat
class _CompatSeries(_pd.Series):
    _metadata = ["_wab_parent"]

    @property
    def _constructor(self):
        return _CompatSeries

    @property
    def _constructor_expanddim(self):
        return _CompatFrame

    def clip(self, lower=None, upper=None, *args, inplace=False, **kwargs):
        parent = getattr(self, "_wab_parent", None)
        if inplace:
            clipped = _pd.Series.clip(self, lower, upper, *args, **kwargs)
            if parent is not None:
                frame, name = parent
                frame[name] = clipped  # pandas<=1.x item-cache write-through (wab_env.py:310)
            return None
        return _pd.Series.clip(self, lower, upper, *args, **kwargs)


class _CompatFrame(_pd.DataFrame):
    @property
    def _constructor(self):
        return _CompatFrame

    @property
    def _constructor_sliced(self):
        return _CompatSeries

    def __getattr__(self, name):
        out = _pd.DataFrame.__getattr__(self, name)
        if isinstance(out, _CompatSeries) and name in self.columns:
            out._wab_parent = (self, name)
        return out

    def append(self, other, ignore_index=False, **kwargs):
        """pandas<2 ``DataFrame.append`` (removed in 2.0)."""
        if isinstance(other, (dict, _pd.Series)):
            # pandas 1.1.2 builds the row from a mixed-type Series: an object-dtype frame. Keeping
            # object dtype lets `food` hold 1 (int) then 0.975 (float) as in the reference (:307).
            row = other if isinstance(other, dict) else other.to_dict()
            other = _CompatFrame([row], dtype=object)
        if len(self.columns) and len(other.columns):
            other = other.reindex(columns=list(self.columns) + [c for c in other.columns if c not in self.columns])
        frames = [f for f in (self, other) if len(f)] or [self]
        return _pd.concat(frames, ignore_index=ignore_index)

    def drop(self, labels=None, *args, **kwargs):
        if args:  # positional axis (pandas<2): drop(labels, 1, inplace=True)
            kwargs["axis"] = args[0]
            args = args[1:]
        return _pd.DataFrame.drop(self, labels, *args, **kwargs)


class _PdProxy(types.ModuleType):
    def __init__(self):
        super().__init__("pandas_proxy")

    def __getattr__(self, name):
        return getattr(_pd, name)

    @staticmethod
    def DataFrame(data=None, *args, **kwargs):
        if isinstance(data, (set, frozenset)):
            data = sorted(data)
        try:
            return _CompatFrame(data, *args, **kwargs)
        except (ValueError, TypeError):
            # action table: dtype=int with None roles -> float column holding NaN (pandas 1.1.2 result)
            kwargs.pop("dtype", None)
            return _CompatFrame(data, *args, **kwargs)

    @staticmethod
    def concat(objs, *args, **kwargs):
        return _pd.concat(objs, *args, **kwargs)


# --------------------------------------------------------------------------- keyed numpy proxy
def _env_key(env):
    # a bare reference env (e.g. built by the reference's own test file) is keyed as (0, 0, 0)
    return getattr(env, "_wab_seed", 0), getattr(env, "_wab_env_id", 0), getattr(env, "_wab_episode", 0)


class _KeyedRandom:
    """Stands in for ``np.random`` inside the reference module only."""

    def random(self, size=None):
        frame = sys._getframe(1)
        site = frame.f_code.co_name
        if site == "generate_n_bush_values":  # wab_env.py:631-635, called from generate_bushes :627
            caller = frame.f_back
            env = caller.f_locals["self"]
            cells = caller.f_locals["new_bushes"]
            seed, eid, ep = _env_key(env)
            words = kr.bush_words(seed, eid, ep, _ints(cells["x"]), _ints(cells["y"]))
        elif site == "initialize_wolves":  # :588-591
            env = frame.f_locals["self"]
            cells = frame.f_locals["new_wolves"]
            seed, eid, ep = _env_key(env)
            opts = env.game_options
            units = kr.init_units(seed, eid, ep, _ints(cells["x"]), _ints(cells["y"]), opts["width"], opts["height"],
                                  opts["chance_wolf_on_square"] / 2)
            return self._checked(units, size, site)
        elif site == "spawn_wolves":  # :571-574
            env = frame.f_locals["self"]
            cells = frame.f_locals["new_wolves"]
            seed, eid, ep = _env_key(env)
            opts = env.game_options
            ox = int(env.ostriches.iloc[0].x)
            oy = int(env.ostriches.iloc[0].y)
            units = kr.spawn_units(
                seed, eid, ep, env.current_turn,
                _ints(cells["x"]) - ox, _ints(cells["y"]) - oy,
                opts["width"], opts["height"], opts["wolf_spawn_margin"], opts["chance_wolf_on_square"] / 2,
            )
            return self._checked(units, size, site)
        elif site == "step":  # despawn, :262-264
            env = frame.f_locals["self"]
            seed, eid, ep = _env_key(env)
            words = kr.despawn_words(seed, eid, ep, env.current_turn, _ints(env.wolves["x"]), _ints(env.wolves["y"]))
        elif site == "spawn_ostriches":  # starting food, :596-597
            env = frame.f_locals["self"]
            return float(kr.to_unit(kr.start_words(*_env_key(env))[0]))
        else:
            raise RuntimeError("unkeyed np.random.random call from %r" % site)
        return self._checked(kr.to_unit(words), size, site)

    @staticmethod
    def _checked(units, size, site):
        if size is None:
            raise RuntimeError("scalar draw at vector site %r" % site)
        n = int(size) if not isinstance(size, tuple) else int(size[0])
        if n != len(units):
            raise RuntimeError("draw count mismatch at %r: %d vs %d" % (site, n, len(units)))
        return units

    def randint(self, low, high=None, size=None):
        frame = sys._getframe(1)
        if frame.f_code.co_name != "spawn_ostriches" or high is not None or low != 2:  # :598-599
            raise RuntimeError("unkeyed np.random.randint call")
        env = frame.f_locals["self"]
        return int(kr.start_words(*_env_key(env))[1] >> 31)

    def seed(self, *_a, **_k):
        pass


def _ints(series):
    return _np.asarray([int(v) for v in series], dtype=_np.int64)


class _NpProxy(types.ModuleType):
    def __init__(self):
        super().__init__("numpy_proxy")
        self.random = _KeyedRandom()

    def __getattr__(self, name):
        return getattr(_np, name)


# --------------------------------------------------------------------------- loader
_MODULE = None


def load_reference(keyed=True):
    """Execute the byte-identical reference source in a private module with shimmed ``pd``/``np``."""
    global _MODULE
    if _MODULE is not None and keyed:
        return _MODULE
    if not reference_available():
        raise FileNotFoundError("reference not present at %s" % REFERENCE_DIR)
    gym_stub.install()
    path = os.path.join(REFERENCE_DIR, "wab_env.py")
    with open(path, "rb") as fh:
        source = fh.read()
    mod = types.ModuleType("wab_env_reference")
    mod.__file__ = path
    exec(compile(source, path, "exec"), mod.__dict__)
    mod.pd = _PdProxy()
    if keyed:
        mod.np = _NpProxy()
        _MODULE = mod
    return mod


def make_env(game_options=None, seed=0, env_id=0):
    """Reference env whose draws are keyed by (seed, env_id, episode). Episode 0 is the reset inside
    ``__init__`` (``wab_env.py:186``); each later ``reset()`` bumps it."""
    mod = load_reference()

    class KeyedWolvesAndBushesEnv(mod.WolvesAndBushesEnv):
        def __init__(self, opts):
            self._wab_seed, self._wab_env_id, self._wab_episode = int(seed), int(env_id), -1
            super().__init__(opts)

        def reset(self):
            self._wab_episode += 1
            return super().reset()

    opts = dict(mod.default_game_options)
    if game_options:
        opts.update(game_options)
    return KeyedWolvesAndBushesEnv(opts)


def hidden_state(env):
    """Hidden state used by the differential tests: ostrich, wolf multiset, bush records."""
    o = env.ostriches.iloc[0]
    wolves = sorted((int(x), int(y)) for x, y in zip(env.wolves["x"], env.wolves["y"]))
    bushes = {(int(x), int(y)): int(f) for x, y, f in zip(env.bushes["x"], env.bushes["y"], env.bushes["food"])}
    return {
        "x": int(o.x), "y": int(o.y), "food": float(o.food), "role": int(o.role),
        "status": int(o.alive_starved_killed), "turn": int(env.current_turn),
        "wolves": wolves, "bushes": bushes,
    }
