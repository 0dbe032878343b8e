"""ctypes front-end of the CPU oracle ``oracle/wab_oracle.c`` (TEST INFRASTRUCTURE, tier B).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; the product package never does.
"""
import ctypes
import os
import subprocess

import numpy as np

from wab_gym_b200.config import GATHERER_TILE_MASK, LOOKOUT_TILE_MASK, GameConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libwab_oracle.so")
MAX_ACTIONS = 8


class OracleConfig(ctypes.Structure):
    _fields_ = [
        ("width", ctypes.c_int32), ("height", ctypes.c_int32), ("max_turns", ctypes.c_int32),
        ("wolf_spawn_margin", ctypes.c_int32), ("n_actions", ctypes.c_int32),
        ("action_dx", ctypes.c_int32 * MAX_ACTIONS), ("action_dy", ctypes.c_int32 * MAX_ACTIONS),
        ("action_role", ctypes.c_int32 * MAX_ACTIONS),
        ("lookout_only", ctypes.c_int32), ("restrict_view", ctypes.c_int32), ("wolves", ctypes.c_int32),
        ("wolves_can_move", ctypes.c_int32), ("god_mode", ctypes.c_int32), ("starting_role", ctypes.c_int32),
        ("starting_food", ctypes.c_double), ("turns_to_fill_food", ctypes.c_double),
        ("turns_to_empty_food", ctypes.c_double), ("chance_wolf_on_square", ctypes.c_double),
        ("wolf_chance_to_despawn", ctypes.c_double),
        ("reward_per_turn", ctypes.c_double), ("reward_for_being_killed", ctypes.c_double),
        ("reward_for_starving", ctypes.c_double), ("reward_for_finishing", ctypes.c_double),
        ("reward_for_eating", ctypes.c_double),
        ("bush_thr", ctypes.POINTER(ctypes.c_uint32)), ("n_bush_thr", ctypes.c_int32),
        ("mask_lookout", ctypes.POINTER(ctypes.c_uint8)), ("mask_gatherer", ctypes.POINTER(ctypes.c_uint8)),
        ("spawn_cdf", ctypes.POINTER(ctypes.c_uint64)), ("init_cdf", ctypes.POINTER(ctypes.c_uint64)),
    ]


class OracleObs(ctypes.Structure):
    _fields_ = [("grids", ctypes.POINTER(ctypes.c_uint8)), ("food", ctypes.c_int32),
                ("role", ctypes.c_int32), ("status", ctypes.c_int32)]


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("wab_oracle.c", "wab2_oracle.c", "wab_oracle.h", "Makefile")]
    if not force and os.path.exists(_LIB_PATH) and os.path.getmtime(_LIB_PATH) >= max(map(os.path.getmtime, srcs)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libwab_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.wab_oracle_create.restype = ctypes.c_void_p
        L.wab_oracle_create.argtypes = [ctypes.POINTER(OracleConfig), ctypes.c_uint64, ctypes.c_uint64]
        L.wab_oracle_destroy.argtypes = [ctypes.c_void_p]
        L.wab_oracle_reset.argtypes = [ctypes.c_void_p, ctypes.POINTER(OracleObs)]
        L.wab_oracle_step.restype = ctypes.c_int
        L.wab_oracle_step.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(OracleObs),
                                      ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)]
        L.wab_oracle_get_state.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 7
        L.wab_oracle_num_wolves.restype = ctypes.c_int32
        L.wab_oracle_num_wolves.argtypes = [ctypes.c_void_p]
        L.wab_oracle_get_wolves.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.wab_oracle_num_bushes.restype = ctypes.c_int32
        L.wab_oracle_num_bushes.argtypes = [ctypes.c_void_p]
        L.wab_oracle_get_bushes.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.wab_oracle_philox.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.wab_oracle_philox2.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]
        L.wab_oracle_ego_proximities.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.wab_oracle_run.restype = ctypes.c_int64
        L.wab_oracle_run.argtypes = [ctypes.POINTER(OracleConfig), ctypes.c_uint64, ctypes.c_int64, ctypes.c_int64,
                                     ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_uint64)]
        _lib = L
    return _lib


class _ConfigHolder:
    """Keeps the numpy buffers an OracleConfig points into alive."""

    def __init__(self, game_options=None):
        self.game = game_options if isinstance(game_options, GameConfig) else GameConfig.from_options(game_options)
        o = self.game.options
        c = OracleConfig()
        c.width, c.height, c.max_turns = int(o["width"]), int(o["height"]), int(o["max_turns"])
        c.wolf_spawn_margin = int(o["wolf_spawn_margin"])
        c.n_actions = self.game.n_actions
        for k, (dx, dy, role) in enumerate(self.game.actions):
            c.action_dx[k], c.action_dy[k], c.action_role[k] = dx, dy, role
        c.lookout_only = int(bool(o["lookout_only"]))
        c.restrict_view = int(bool(o["restrict_view"]))
        c.wolves = int(bool(o["wolves"]))
        c.wolves_can_move = int(bool(o["wolves_can_move"]))
        c.god_mode = int(bool(o.get("god_mode")))
        c.starting_role = -1 if o["starting_role"] is None else int(o["starting_role"])
        c.starting_food = -1.0 if o["starting_food"] is None else float(o["starting_food"])
        c.turns_to_fill_food = float(o["turns_to_fill_food"])
        c.turns_to_empty_food = float(o["turns_to_empty_food"])
        c.chance_wolf_on_square = float(o["chance_wolf_on_square"])
        c.wolf_chance_to_despawn = float(o["wolf_chance_to_despawn"])
        for name in ("reward_per_turn", "reward_for_being_killed", "reward_for_starving",
                     "reward_for_finishing", "reward_for_eating"):
            setattr(c, name, float(o[name]))
        self._thr = np.ascontiguousarray(self.game.bush_thr, dtype=np.uint32)
        c.bush_thr = self._thr.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
        c.n_bush_thr = len(self._thr)
        from . import keyed_rng as kr
        p = o["chance_wolf_on_square"] / 2
        self._scdf = np.asarray(kr.binomial_thresholds(kr.ring_size(c.width, c.height, c.wolf_spawn_margin), p), dtype=np.uint64)
        self._icdf = np.asarray(kr.binomial_thresholds(c.width * c.height, p), dtype=np.uint64)
        c.spawn_cdf = self._scdf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
        c.init_cdf = self._icdf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))
        if c.width == 11 and c.height == 11:
            self._ml = np.ascontiguousarray(LOOKOUT_TILE_MASK, dtype=np.uint8)
            self._mg = np.ascontiguousarray(GATHERER_TILE_MASK, dtype=np.uint8)
            c.mask_lookout = self._ml.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
            c.mask_gatherer = self._mg.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
        self.c = c


class OracleEnv:
    """One CPU oracle environment with the reference's reset()/step() surface."""

    def __init__(self, game_options=None, seed=0, env_id=0):
        self._cfg = _ConfigHolder(game_options)
        self.width, self.height = self._cfg.c.width, self._cfg.c.height
        self.n_actions = self._cfg.c.n_actions
        self._h = lib().wab_oracle_create(ctypes.byref(self._cfg.c), seed, env_id)
        if not self._h:
            raise ValueError("oracle rejected the configuration")
        self._grids = np.zeros((3, self.width, self.height), dtype=np.uint8)
        self._obs = OracleObs()
        self._obs.grids = self._grids.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().wab_oracle_destroy(self._h)
            self._h = None

    def _obs_tuple(self):
        return self._grids.copy(), int(self._obs.food), int(self._obs.role), int(self._obs.status)

    def reset(self):
        lib().wab_oracle_reset(self._h, ctypes.byref(self._obs))
        return self._obs_tuple()

    def step(self, action):
        r = ctypes.c_double()
        d = ctypes.c_int32()
        if lib().wab_oracle_step(self._h, int(action), ctypes.byref(self._obs), ctypes.byref(r), ctypes.byref(d)):
            raise IndexError("single positional indexer is out-of-bounds")
        return self._obs_tuple(), r.value, bool(d.value)

    def ego_proximities(self):
        """(wolf proximities[5], bush proximities[5]) of wab_env.py:637-667 for up, right, down, left, stay."""
        out = np.zeros(10, dtype=np.int32)
        lib().wab_oracle_ego_proximities(self._h, out.ctypes.data)
        return [int(v) for v in out[:5]], [int(v) for v in out[5:]]

    def hidden_state(self):
        x, y, role, status, turn = (ctypes.c_int32() for _ in range(5))
        food = ctypes.c_double()
        ep = ctypes.c_int64()
        L = lib()
        L.wab_oracle_get_state(self._h, *(ctypes.addressof(v) for v in (x, y, food, role, status, turn, ep)))
        nw = L.wab_oracle_num_wolves(self._h)
        wolves = np.zeros((nw, 2), dtype=np.int32)
        L.wab_oracle_get_wolves(self._h, wolves.ctypes.data)
        nb = L.wab_oracle_num_bushes(self._h)
        bushes = np.zeros((nb, 3), dtype=np.int32)
        L.wab_oracle_get_bushes(self._h, bushes.ctypes.data)
        return {
            "x": x.value, "y": y.value, "food": food.value, "role": role.value, "status": status.value,
            "turn": turn.value, "episode": ep.value,
            "wolves": sorted((int(a), int(b)) for a, b in wolves),
            "bushes": {(int(a), int(b)): int(f) for a, b, f in bushes},
        }


def philox(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().wab_oracle_philox(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


def philox2(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    out = np.zeros(2, dtype=np.uint32)
    lib().wab_oracle_philox2(c.ctypes.data, int(key) & 0xFFFFFFFF, out.ctypes.data)
    return out


def run(game_options, seed, n_envs, n_steps, actions, n_threads=0):
    """Lockstep run with reset-on-done; returns (env_steps, checksum). actions: u8[n_steps, n_envs]."""
    holder = _ConfigHolder(game_options)
    a = np.ascontiguousarray(actions, dtype=np.uint8)
    assert a.shape == (n_steps, n_envs)
    cs = ctypes.c_uint64()
    n = lib().wab_oracle_run(ctypes.byref(holder.c), seed, n_envs, n_steps, a.ctypes.data, n_threads, ctypes.byref(cs))
    return int(n), int(cs.value)
