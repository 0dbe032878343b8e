"""Host-side image of the reference's ``game_options`` (``/root/reference/wab_env.py:11-39``).

Turns an options dict into the POD ``WabConfig`` of ``include/wab_b200.h``: the action table
(``wab_env.py:149-182``), exact integer thresholds for the keyed 32-bit draws, the bush-value
threshold table (``:631-635``), the f32 reward table (``:328-340``) and, when it can be *proven*
equivalent, the integer food counter that replaces the reference's float64 food (``:307-322, :452``).
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass, field
from fractions import Fraction
from typing import Dict, List, Optional, Tuple

import numpy as np

WAB_ABI_VERSION = 1
WAB_MAX_ACTIONS = 8
WAB_FOOD_F64, WAB_FOOD_INT = 0, 1

#: Same keys and values as the reference's ``default_game_options`` (wab_env.py:11-39).
default_game_options: Dict[str, object] = {
    "reward_per_turn": 0,
    "reward_for_being_killed": -1,
    "reward_for_starving": -1,
    "reward_for_finishing": 1,
    "reward_for_eating": 0.1,
    "gatherer_only": False,
    "lookout_only": True,
    "restrict_view": False,
    "starting_role": 1,
    "max_turns": 80,
    "num_ostriches": 1,
    "height": 11,
    "width": 11,
    "bush_power": 100,
    "max_berries_per_bush": 200,
    "turns_to_fill_food": 8,
    "turns_to_empty_food": 40,
    "starting_food": 1,
    "wolf_spawn_margin": 1,
    "chance_wolf_on_square": 0.001,
    "wolf_chance_to_despawn": 0.05,
    "wolves": True,
    "wolves_can_move": True,
}

#: Blind-spot masks of the reference (wab_env.py:109-123 lookout, :125-139 gatherer); 1 = not visible.
LOOKOUT_TILE_MASK = np.array(
    [[1 if (min(i, 10 - i) + min(j, 10 - j)) < 3 else 0 for j in range(11)] for i in range(11)], dtype=np.uint8
)
GATHERER_TILE_MASK = np.array(
    [[0 if (abs(i - 5) <= 2 and abs(j - 5) <= 2 and abs(i - 5) + abs(j - 5) <= 3) else 1 for j in range(11)]
     for i in range(11)],
    dtype=np.uint8,
)


class WabConfigStruct(ctypes.Structure):
    """ctypes mirror of ``struct WabConfig`` (include/wab_b200.h)."""

    _fields_ = [
        ("abi_version", ctypes.c_int32),
        ("width", ctypes.c_int32),
        ("height", ctypes.c_int32),
        ("max_turns", ctypes.c_int32),
        ("wolf_spawn_margin", ctypes.c_int32),
        ("n_actions", ctypes.c_int32),
        ("action_dx", ctypes.c_int8 * WAB_MAX_ACTIONS),
        ("action_dy", ctypes.c_int8 * WAB_MAX_ACTIONS),
        ("action_role", ctypes.c_int8 * WAB_MAX_ACTIONS),
        ("lookout_only", ctypes.c_uint8),
        ("restrict_view", ctypes.c_uint8),
        ("wolves", ctypes.c_uint8),
        ("wolves_can_move", ctypes.c_uint8),
        ("god_mode", ctypes.c_uint8),
        ("starting_role", ctypes.c_int8),
        ("food_mode", ctypes.c_uint8),
        ("auto_reset", ctypes.c_uint8),
        ("food_int_start", ctypes.c_int32),
        ("food_int_inc", ctypes.c_int32),
        ("food_int_max", ctypes.c_int32),
        ("wolf_cap", ctypes.c_int32),
        ("log_cap", ctypes.c_int32),
        ("food_start", ctypes.c_double),
        ("food_inc", ctypes.c_double),
        ("food_dec", ctypes.c_double),
        ("food_obs_scale", ctypes.c_double),
        ("spawn_cdf", ctypes.c_uint64 * 32),
        ("init_cdf", ctypes.c_uint64 * 32),
        ("thr_keep", ctypes.c_uint64),
        ("reward_table", ctypes.c_float * 8),
        ("mask_lookout", ctypes.c_uint32 * 4),
        ("mask_gatherer", ctypes.c_uint32 * 4),
    ]


def action_table(options) -> List[Tuple[int, int, int]]:
    """(dx, dy, role) rows of ``action_definitions`` (wab_env.py:149-182); role -1 = NaN = keep.
    ``gatherer_only`` takes precedence over ``lookout_only`` exactly as the if/elif there."""
    moves = [(0, 1, -1), (1, 0, -1), (0, -1, -1), (-1, 0, -1)]  # up, right, down, left
    if options["gatherer_only"]:
        return moves + [(0, 0, 1)]
    if options["lookout_only"]:
        return moves + [(0, 0, 0)]
    return moves + [(0, 0, 1), (0, 0, 0)]


def lt_threshold(p: float) -> int:
    """Least integer t with: for 32-bit w, (w * 2**-32 < p)  <=>  (w < t). Exact (rational arithmetic)."""
    t = math.ceil(Fraction(float(p)) * (1 << 32))
    return min(max(t, 0), 1 << 32)


BINOMIAL_TABLE = 32


def binomial_thresholds(n: int, p: float) -> List[int]:
    """Inverse-CDF table of the binomial-first draws: T_k = min(ceil(CDF_{n,p}(k) * 2**64), 2**64 - 1), k < 32, in
    exact rational arithmetic on the double p. With a 64-bit draw v, the number of events among n cells is
    K = #{k : v >= T_k}. Raises when the mass beyond the table is not negligible (absurdly large chances)."""
    n = int(n)
    pf = min(max(Fraction(float(p)), Fraction(0)), Fraction(1))
    out, cdf = [], Fraction(0)
    for k in range(BINOMIAL_TABLE):
        if k <= n:
            cdf += math.comb(n, k) * pf ** k * (1 - pf) ** (n - k)
        out.append(min(int(math.ceil(cdf * (1 << 64))), (1 << 64) - 1))
    if 1 - cdf > Fraction(1, 1 << 40):
        raise ValueError("chance_wolf_on_square is too large for the binomial-first draw table (n=%d, p=%g)" % (n, float(p)))
    return out


def gt_threshold(q: float) -> int:
    """Least integer t with: for 32-bit w, (w * 2**-32 > q)  <=>  (w >= t)."""
    t = math.floor(Fraction(float(q)) * (1 << 32)) + 1
    return min(max(t, 0), (1 << 32) + 1)


def reference_bush_value(words: np.ndarray, bush_power, max_berries) -> np.ndarray:
    """The reference's ``generate_n_bush_values`` formula (wab_env.py:631-635) applied to the uniforms
    ``U = w * 2**-32``: ``np.round(U ** bush_power * max_berries)`` (np.round is half-to-even)."""
    u = np.asarray(words, dtype=np.float64) * (2.0 ** -32)
    return np.round(u ** bush_power * max_berries)


def bush_thresholds(bush_power, max_berries) -> np.ndarray:
    """thr[k-1] = least 32-bit word w whose reference bush value is >= k, for every reachable k.

    Found by bisection on the reference's own floating-point formula (monotone in w), so that device
    code only compares integers and never evaluates ``pow``. ``tests/test_config.py`` checks the
    table against exact big-integer arithmetic and against the formula on both sides of every
    threshold."""
    top = int(reference_bush_value(np.array([0xFFFFFFFF], dtype=np.uint64), bush_power, max_berries)[0])
    if top <= 0:
        return np.zeros(0, dtype=np.uint32)
    ks = np.arange(1, top + 1, dtype=np.float64)
    lo = np.zeros(top, dtype=np.uint64)  # invariant: value(lo) < k  (value(0) = 0)
    hi = np.full(top, 0xFFFFFFFF, dtype=np.uint64)  # invariant: value(hi) >= k
    if reference_bush_value(lo[:1], bush_power, max_berries)[0] >= 1:
        raise ValueError("bush_power / max_berries_per_bush give a non-zero bush at U = 0")
    while np.any(hi - lo > 1):
        mid = (lo + hi) >> np.uint64(1)
        ge = reference_bush_value(mid, bush_power, max_berries) >= ks
        hi = np.where(ge, mid, hi)
        lo = np.where(ge, lo, mid)
    return hi.astype(np.uint32)


def prove_integer_food(turns_to_fill, turns_to_empty, starting_food, max_turns, limit=2_000_000):
    """Exhaustive check that an integer counter in units of 1/turns_to_empty reproduces the reference's
    float64 food (wab_env.py:307-322) on every reachable eat / no-eat path of an episode of
    ``max_turns`` turns: same starvation turn (``food <= 0``) and same ``ceil(food * empty)`` (:452).

    Returns (start_units, inc_units, max_units) or None when not provable (then fp64 food is used)."""
    if starting_food is None:
        return None
    try:
        empty = Fraction(turns_to_empty)
        fill = Fraction(turns_to_fill)
        start = Fraction(starting_food) * empty
    except (TypeError, ValueError):
        return None
    if empty.denominator != 1 or empty <= 0 or empty > 200 or fill <= 0:
        return None
    inc_units = empty / fill
    if inc_units.denominator != 1 or start.denominator != 1 or not (0 <= start <= empty):
        return None
    e_units, inc_units, start_units = int(empty), int(inc_units), int(start)
    inc = 1 / turns_to_fill  # python true division, as the reference
    dec = 1 / turns_to_empty
    frontier = {(float(starting_food), start_units)}
    for _turn in range(int(max_turns)):
        nxt = set()
        for food, c in frontier:
            for eat in (False, True):
                f2, c2 = food, c
                if eat:
                    f2 = min(max(f2 + inc, 0), 1)  # :307-310
                    c2 = min(c2 + inc_units, e_units)
                f2 -= dec  # :316
                c2 -= 1
                if (f2 <= 0) != (c2 <= 0):  # :319-322
                    return None
                if f2 <= 0:
                    continue  # episode over
                if int(math.ceil(f2 * turns_to_empty)) != c2:  # :452
                    return None
                nxt.add((f2, c2))
        if len(nxt) > limit:
            return None
        frontier = nxt
    return start_units, inc_units, e_units


def _mask_words(mask: np.ndarray) -> List[int]:
    bits = 0
    flat = np.asarray(mask, dtype=np.uint8).reshape(-1)
    for c, v in enumerate(flat):
        if v:
            bits |= 1 << c
    return [(bits >> (32 * k)) & 0xFFFFFFFF for k in range(4)]


@dataclass
class GameConfig:
    """Validated options + everything derived from them on the host."""

    options: Dict[str, object]
    actions: List[Tuple[int, int, int]]
    bush_thr: np.ndarray
    spawn_cdf: List[int]
    init_cdf: List[int]
    thr_keep: int
    food_mode: int
    food_int: Tuple[int, int, int]
    reward_table64: np.ndarray = field(repr=False)
    auto_reset: bool = True
    wolf_cap: int = 16
    log_cap: int = 80

    @property
    def n_actions(self) -> int:
        return len(self.actions)

    @classmethod
    def from_options(cls, game_options: Optional[dict] = None, *, auto_reset: bool = True,
                     force_f64_food: bool = False, wolf_cap: int = 16, log_cap: Optional[int] = None) -> "GameConfig":
        opts = dict(default_game_options)
        if game_options:
            opts.update(game_options)  # KeyError semantics: missing keys fall back to defaults here
        width, height = int(opts["width"]), int(opts["height"])
        if width % 2 == 0 or height % 2 == 0:
            raise ValueError("width and height must be odd numbers")  # wab_env.py:147-148
        max_turns = int(opts["max_turns"])
        if not (1 <= max_turns <= 30000):
            raise ValueError("max_turns must be in [1, 30000] (positions are int16 on the device)")
        p_spawn = opts["chance_wolf_on_square"] / 2  # wab_env.py:573, :590
        ring = (width + 2 * int(opts["wolf_spawn_margin"])) * (height + 2 * int(opts["wolf_spawn_margin"])) - width * height
        spawn_cdf = binomial_thresholds(ring, p_spawn)
        init_cdf = binomial_thresholds(width * height, p_spawn)
        thr_keep = gt_threshold(opts["wolf_chance_to_despawn"])
        proof = None
        if auto_reset and not force_f64_food:
            proof = prove_integer_food(opts["turns_to_fill_food"], opts["turns_to_empty_food"],
                                       opts["starting_food"], max_turns)
        if not (0 < float(opts["turns_to_empty_food"]) <= 255):
            raise ValueError("turns_to_empty_food must be in (0, 255] (food observation is a byte)")
        rewards = np.zeros(8, dtype=np.float64)
        terminal = [opts["reward_per_turn"], opts["reward_for_finishing"], opts["reward_for_starving"],
                    opts["reward_for_being_killed"]]
        for ate in (0, 1):
            for outcome in range(4):
                r = 0  # wab_env.py:251
                if ate:
                    r += opts["reward_for_eating"]  # :313
                r += terminal[outcome]  # :328-340
                rewards[ate * 4 + outcome] = r
        if log_cap is None:
            log_cap = min(max_turns, 255)
        if not (1 <= wolf_cap <= 64) or not (1 <= log_cap <= 255):
            raise ValueError("wolf_cap must be in [1, 64] and log_cap in [1, 255]")
        return cls(
            options=opts,
            actions=action_table(opts),
            bush_thr=bush_thresholds(opts["bush_power"], opts["max_berries_per_bush"]),
            spawn_cdf=spawn_cdf,
            init_cdf=init_cdf,
            thr_keep=thr_keep,
            food_mode=WAB_FOOD_INT if proof else WAB_FOOD_F64,
            food_int=proof or (0, 0, 0),
            reward_table64=rewards,
            auto_reset=auto_reset,
            wolf_cap=wolf_cap,
            log_cap=log_cap,
        )

    def to_struct(self) -> WabConfigStruct:
        o = self.options
        s = WabConfigStruct()
        s.abi_version = WAB_ABI_VERSION
        s.width, s.height = int(o["width"]), int(o["height"])
        s.max_turns = int(o["max_turns"])
        s.wolf_spawn_margin = int(o["wolf_spawn_margin"])
        s.n_actions = self.n_actions
        for k, (dx, dy, role) in enumerate(self.actions):
            s.action_dx[k], s.action_dy[k], s.action_role[k] = dx, dy, role
        s.lookout_only = int(bool(o["lookout_only"]))
        s.restrict_view = int(bool(o["restrict_view"]))
        s.wolves = int(bool(o["wolves"]))
        s.wolves_can_move = int(bool(o["wolves_can_move"]))
        s.god_mode = int(bool(o.get("god_mode")))
        s.starting_role = -1 if o["starting_role"] is None else int(o["starting_role"])
        s.food_mode = self.food_mode
        s.auto_reset = int(self.auto_reset)
        s.food_int_start, s.food_int_inc, s.food_int_max = self.food_int
        s.wolf_cap, s.log_cap = self.wolf_cap, self.log_cap
        s.food_start = -1.0 if o["starting_food"] is None else float(o["starting_food"])
        s.food_inc = 1 / o["turns_to_fill_food"]
        s.food_dec = 1 / o["turns_to_empty_food"]
        s.food_obs_scale = float(o["turns_to_empty_food"])
        for k in range(BINOMIAL_TABLE):
            s.spawn_cdf[k], s.init_cdf[k] = self.spawn_cdf[k], self.init_cdf[k]
        s.thr_keep = self.thr_keep
        for k in range(8):
            s.reward_table[k] = float(np.float32(self.reward_table64[k]))
        for k, w in enumerate(_mask_words(LOOKOUT_TILE_MASK)):
            s.mask_lookout[k] = w
        for k, w in enumerate(_mask_words(GATHERER_TILE_MASK)):
            s.mask_gatherer[k] = w
        return s
