"""ctypes front-end of ``oracle/wab2_oracle.c`` — the CPU oracle of the Environment 2.0 world turn
(TEST INFRASTRUCTURE; same import rules as ``oracle/wab_oracle.py``)."""
import ctypes

import numpy as np

from . import wab_oracle

#: same keys / values as the reference's v2 ``default_game_options`` that the world turn reads
#: ("/root/reference/Environment 2.0/WAB_Environment2.py":9-50)
DEFAULT_V2_OPTIONS = {
    "starting_role": 1, "food_per_bush": 20, "food_given_per_turn": 5, "ostrich_starting_food": 40.0,
    "lookout_view_radius": 9, "gatherer_view_radius": 5, "wolf_starting_food": 20, "wolf_food_for_eating_ostrich": 10,
    "wolf_view_radius": 6,
}


class Config2(ctypes.Structure):
    _fields_ = [("width", ctypes.c_int32), ("height", ctypes.c_int32), ("n_ostriches", ctypes.c_int32),
                ("n_wolves", ctypes.c_int32), ("n_bushes", ctypes.c_int32), ("lookout_view_radius", ctypes.c_int32),
                ("gatherer_view_radius", ctypes.c_int32), ("wolf_view_radius", ctypes.c_int32),
                ("starting_role", ctypes.c_int32), ("ostrich_starting_food", ctypes.c_double),
                ("wolf_starting_food", ctypes.c_double), ("wolf_food_for_eating_ostrich", ctypes.c_double),
                ("food_per_bush", ctypes.c_double), ("food_given_per_turn", ctypes.c_double)]


def make_config(width, height, n_ostriches, n_wolves, n_bushes, game_options=None):
    o = dict(DEFAULT_V2_OPTIONS)
    if game_options:
        o.update({k: v for k, v in game_options.items() if k in o})
    return Config2(width, height, n_ostriches, n_wolves, n_bushes, int(o["lookout_view_radius"]),
                   int(o["gatherer_view_radius"]), int(o["wolf_view_radius"]), int(o["starting_role"]),
                   float(o["ostrich_starting_food"]), float(o["wolf_starting_food"]),
                   float(o["wolf_food_for_eating_ostrich"]), float(o["food_per_bush"]), float(o["food_given_per_turn"]))


_ready = False


def _lib():
    global _ready
    L = wab_oracle.lib()
    if not _ready:
        L.wab2_oracle_create.restype = ctypes.c_void_p
        L.wab2_oracle_create.argtypes = [ctypes.POINTER(Config2), ctypes.c_uint64, ctypes.c_uint64]
        L.wab2_oracle_destroy.argtypes = [ctypes.c_void_p]
        L.wab2_oracle_reset.argtypes = [ctypes.c_void_p]
        L.wab2_oracle_take_action.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
        L.wab2_oracle_get_obs.restype = ctypes.c_int32
        L.wab2_oracle_get_obs.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
        L.wab2_oracle_get_state.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.wab2_oracle_set_state.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]
        L.wab2_oracle_run.restype = ctypes.c_int64
        L.wab2_oracle_run.argtypes = [ctypes.POINTER(Config2), ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int64, ctypes.c_int32,
                                      ctypes.c_int32, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]
        L.wab2_oracle_turn.restype = ctypes.c_int32
        L.wab2_oracle_turn.argtypes = [ctypes.c_void_p]
        _ready = True
    return L


class OracleWorld2:
    """One Environment-2.0 world with the reference's ``get_obs`` / ``take_action`` / ``reset_environment`` surface."""

    def __init__(self, width, height, n_ostriches, n_wolves, n_bushes, game_options=None, seed=0, env_id=0, window_radius=None):
        self.cfg = make_config(width, height, n_ostriches, n_wolves, n_bushes, game_options)
        self.n = n_ostriches + n_wolves + n_bushes
        self.R = window_radius if window_radius is not None else max(self.cfg.lookout_view_radius, self.cfg.gatherer_view_radius,
                                                                      self.cfg.wolf_view_radius)
        self._h = _lib().wab2_oracle_create(ctypes.byref(self.cfg), seed, env_id)

    def __del__(self):
        if getattr(self, "_h", None):
            _lib().wab2_oracle_destroy(self._h)
            self._h = None

    def reset_environment(self):
        _lib().wab2_oracle_reset(self._h)

    def get_obs(self, entity):
        s = 2 * self.R + 1
        planes = np.zeros((3, s, s), dtype=np.uint8)
        internal = np.zeros(5, dtype=np.float64)
        rows = _lib().wab2_oracle_get_obs(self._h, entity, self.R, planes.ctypes.data, internal.ctypes.data)
        return planes, internal, rows

    def take_action(self, entity, action):
        r, d = ctypes.c_double(), ctypes.c_int32()
        _lib().wab2_oracle_take_action(self._h, entity, int(action), ctypes.addressof(r), ctypes.addressof(d))
        return r.value, bool(d.value)

    def state(self):
        out = np.zeros((self.n, 9), dtype=np.float64)
        _lib().wab2_oracle_get_state(self._h, out.ctypes.data)
        return out

    def set_state(self, state9, turn=0):
        """Test hook: overwrite every entity (layout of ``state()``); the reference's KAT worlds are built this way."""
        st = np.ascontiguousarray(state9, dtype=np.float64)
        assert st.shape == (self.n, 9)
        _lib().wab2_oracle_set_state(self._h, st.ctypes.data, int(turn))

    @property
    def turn(self):
        return _lib().wab2_oracle_turn(self._h)


def run(width, height, n_ostriches, n_wolves, n_bushes, game_options, seed, env_id_base, n_envs, episodes, turns, actions,
        window_radius=None, threads=0):
    """Batch checksum run (see ``wab2_oracle_run`` in wab2_oracle.c): ``actions`` u8[episodes*turns, A, n_envs].
    Returns (world turns executed, checksum of every observation/reward/done, checksum of the final entity tables)."""
    import os
    cfg = make_config(width, height, n_ostriches, n_wolves, n_bushes, game_options)
    R = window_radius if window_radius is not None else max(cfg.lookout_view_radius, cfg.gatherer_view_radius, cfg.wolf_view_radius)
    a = np.ascontiguousarray(actions, dtype=np.uint8)
    assert a.shape == (episodes * turns, n_ostriches + n_wolves, n_envs), a.shape
    if threads <= 0:
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    out = np.zeros(2, dtype=np.int64)
    done = _lib().wab2_oracle_run(ctypes.byref(cfg), seed, env_id_base, n_envs, episodes, turns, a.ctypes.data, R, threads, out.ctypes.data)
    return int(done), int(out[0]), int(out[1])
