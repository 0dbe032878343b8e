"""ctypes front-end of tests/hostsim/hostsim.cpp: the kernel logic header compiled for the host
(TEST-ONLY; see the header comment of hostsim.cpp). Built on demand with g++."""
import ctypes
import os
import subprocess

import numpy as np

from wab_gym_b200.config import GameConfig, WabConfigStruct

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(os.path.dirname(_HERE))
_LIB = os.path.join(_HERE, "libhostsim.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, "hostsim.cpp"),
            os.path.join(_REPO, "wab_gym_b200", "csrc", "wab_core.cuh"),
            os.path.join(_REPO, "wab_gym_b200", "csrc", "wab_params.h"),
            os.path.join(_REPO, "wab_gym_b200", "csrc", "wab_features.cuh"),
            os.path.join(_REPO, "wab_gym_b200", "csrc", "wab2_core.cuh"),
            os.path.join(_REPO, "include", "wab_b200.h")]
    if not force and os.path.exists(_LIB) and os.path.getmtime(_LIB) >= max(map(os.path.getmtime, srcs)):
        return _LIB
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", _LIB, srcs[0]])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        L.hostsim_create.restype = ctypes.c_void_p
        L.hostsim_create.argtypes = [ctypes.POINTER(WabConfigStruct), ctypes.c_void_p, ctypes.c_int32, ctypes.c_uint64,
                                     ctypes.c_uint64, ctypes.c_char_p, ctypes.c_int]
        L.hostsim_destroy.argtypes = [ctypes.c_void_p]
        L.hostsim_reset.argtypes = [ctypes.c_void_p] * 6
        L.hostsim_step.argtypes = [ctypes.c_void_p, ctypes.c_int32] + [ctypes.c_void_p] * 8
        L.hostsim_state.argtypes = [ctypes.c_void_p] * 5
        L.hostsim_features.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                       ctypes.c_void_p]
        L.hostsim2_create.restype = ctypes.c_void_p
        L.hostsim2_create.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64]
        L.hostsim2_destroy.argtypes = [ctypes.c_void_p]
        L.hostsim2_reset.argtypes = [ctypes.c_void_p]
        L.hostsim2_observe.restype = ctypes.c_int32
        L.hostsim2_observe.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
        L.hostsim2_act.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
        L.hostsim2_state.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.hostsim_reveal_probe.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int64, ctypes.c_void_p]
        L.hostsim_philox2.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]
        L.hostsim_philox.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]
        _lib = L
    return _lib


class HostSimEnv:
    def __init__(self, game_options=None, seed=0, env_id=0, auto_reset=True, force_f64_food=False, wolf_cap=15):
        self.game = GameConfig.from_options(game_options, auto_reset=auto_reset, force_f64_food=force_f64_food,
                                            wolf_cap=wolf_cap)
        self._cs = self.game.to_struct()
        thr = np.ascontiguousarray(self.game.bush_thr, dtype=np.uint32)
        err = ctypes.create_string_buffer(256)
        self._h = lib().hostsim_create(ctypes.byref(self._cs), thr.ctypes.data, len(thr), seed, env_id, err, 256)
        if not self._h:
            raise ValueError(err.value.decode())
        self._grids = np.zeros((3, 11, 11), dtype=np.uint8)
        self._i = [ctypes.c_int32() for _ in range(6)]
        self._r = ctypes.c_float()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().hostsim_destroy(self._h)
            self._h = None

    def reset(self):
        f, ro, st, ov = self._i[:4]
        lib().hostsim_reset(self._h, self._grids.ctypes.data, *(ctypes.addressof(v) for v in (f, ro, st, ov)))
        return (self._grids.copy(), f.value, ro.value, st.value), ov.value

    def step(self, action):
        f, ro, st, dn, inf, ov = self._i
        lib().hostsim_step(self._h, int(action), self._grids.ctypes.data, ctypes.addressof(f), ctypes.addressof(ro),
                           ctypes.addressof(st), ctypes.addressof(self._r), ctypes.addressof(dn),
                           ctypes.addressof(inf), ctypes.addressof(ov))
        return (self._grids.copy(), f.value, ro.value, st.value), self._r.value, bool(dn.value), inf.value, ov.value

    def hidden_state(self):
        scal = np.zeros(9, dtype=np.int32)
        food = ctypes.c_double()
        wolves = np.zeros((64, 2), dtype=np.int32)
        mask = np.zeros(4, dtype=np.uint32)
        lib().hostsim_state(self._h, scal.ctypes.data, ctypes.addressof(food), wolves.ctypes.data, mask.ctypes.data)
        nw = int(scal[7])
        return {"x": int(scal[0]), "y": int(scal[1]), "food": food.value, "role": int(scal[3]), "status": int(scal[4]),
                "turn": int(scal[5]), "episode": int(scal[6]), "wolves": sorted((int(a), int(b)) for a, b in wolves[:nw]),
                "n_log": int(scal[8]), "bush_mask": mask}


def reveal_probe(seed, thr, iters):
    """(tied cells, mismatching cells, cells checked) of the reveal paths against full-word draws."""
    out = np.zeros(3, dtype=np.int64)
    lib().hostsim_reveal_probe(int(seed), int(thr), int(iters), out.ctypes.data)
    return tuple(int(v) for v in out)


def philox2(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    out = np.zeros(2, dtype=np.uint32)
    lib().hostsim_philox2(c.ctypes.data, int(key) & 0xFFFFFFFF, out.ctypes.data)
    return out


def philox(ctr, key):
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().hostsim_philox(c.ctypes.data, int(key[0]), int(key[1]), out.ctypes.data)
    return out


def features(wolf_grid, bush_grid, food, role, status):
    """PragmaticObsWrapper features (28 bytes) of one observation, through the kernel header."""
    w = np.ascontiguousarray(np.asarray(wolf_grid) != 0, dtype=np.uint8)
    b = np.ascontiguousarray(np.asarray(bush_grid) != 0, dtype=np.uint8)
    out = np.zeros(28, dtype=np.uint8)
    lib().hostsim_features(w.ctypes.data, b.ctypes.data, int(food), int(role), int(status), out.ctypes.data)
    return out


class HostSimWorld2:
    """The Environment-2.0 logic header (wab2_core.cuh) compiled for the host: one world."""

    def __init__(self, width, height, n_ostriches, n_wolves, n_bushes, game_options=None, seed=0, env_id=0):
        from wab_gym_b200.world2 import make_config2
        self.cfg = make_config2(width, height, n_ostriches, n_wolves, n_bushes, game_options)
        self.n = n_ostriches + n_wolves + n_bushes
        self.R = self.cfg.window_radius
        self._h = lib().hostsim2_create(ctypes.byref(self.cfg), seed, env_id)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().hostsim2_destroy(self._h)
            self._h = None

    def reset_environment(self):
        lib().hostsim2_reset(self._h)

    def get_obs(self, entity):
        s = 2 * self.R + 1
        planes = np.zeros((3, s, s), dtype=np.uint8)
        internal = np.zeros(5, dtype=np.int32)
        rows = lib().hostsim2_observe(self._h, entity, planes.ctypes.data, internal.ctypes.data)
        return planes, internal, rows

    def take_action(self, entity, action):
        r, d = ctypes.c_float(), ctypes.c_int32()
        lib().hostsim2_act(self._h, entity, int(action), ctypes.addressof(r), ctypes.addressof(d))
        return r.value, bool(d.value)

    def state(self):
        out = np.zeros((self.n, 9), dtype=np.int32)
        turn = ctypes.c_int32()
        lib().hostsim2_state(self._h, out.ctypes.data, ctypes.addressof(turn))
        return out, turn.value
