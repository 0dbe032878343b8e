"""Shared helpers for the parity tests."""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.basename(p)[len("trace_"):-len(".npz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "trace_*.npz")))


def sized_names():
    """Traces of viewports other than 11x11 / spawn margins other than 1 (oracle/make_golden.py --sized)."""
    return sorted(os.path.basename(p)[len("sized_"):-len(".npz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "sized_*.npz")))


def load_golden(name, sized=False):
    z = np.load(os.path.join(GOLDEN_DIR, ("sized_%s.npz" if sized else "trace_%s.npz") % name))
    meta = json.loads(str(z["meta"]))
    return meta, {k: z[k] for k in z.files if k != "meta"}


def bush_digest(bushes):
    """Same digest as oracle/make_golden.py (order-free hash of {(x, y): food})."""
    h = 0
    for (x, y), f in bushes.items():
        v = ((x & 0xFFFF) | ((y & 0xFFFF) << 16) | ((f & 0xFFFF) << 32)) * 0x9E3779B97F4A7C15 & 0xFFFFFFFFFFFFFFFF
        v ^= v >> 29
        h = (h + v * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    return h


def golden_wolves(tr, t):
    n = int(tr["n_wolves"][t])
    return sorted((int(a), int(b)) for a, b in tr["wolves"][t][:n])


def window_mask_from_bushes(hs):
    """121-bit occupancy the kernel state must hold, from the oracle's bush records."""
    m = 0
    for (bx, by), f in hs["bushes"].items():
        dx, dy = hs["x"] - bx, hs["y"] - by
        if abs(dx) <= 5 and abs(dy) <= 5 and f > 0:
            m |= 1 << ((dx + 5) * 11 + dy + 5)
    return m


def mask_words_to_int(words):
    return int(words[0]) | int(words[1]) << 32 | int(words[2]) << 64 | int(words[3]) << 96


OPTION_SETS = {
    # name -> (overrides, greedy policy?)   (same sets as oracle/make_golden.py, plus a few)
    "defaults": ({}, False),
    "six_actions_random_start": ({"lookout_only": False, "starting_role": None, "starting_food": None}, True),
    "restrict_view": ({"lookout_only": False, "restrict_view": True, "starting_role": None}, True),
    "dense": ({"chance_wolf_on_square": 0.012, "bush_power": 12, "wolf_chance_to_despawn": 0.2,
               "reward_per_turn": 0.25, "reward_for_eating": 0.5}, False),
    "gatherer_static_wolves": ({"gatherer_only": True, "wolves_can_move": False, "chance_wolf_on_square": 0.004}, True),
    "god_mode_short": ({"god_mode": True, "max_turns": 30, "turns_to_fill_food": 4, "turns_to_empty_food": 20,
                        "chance_wolf_on_square": 0.005, "max_berries_per_bush": 3, "bush_power": 30}, True),
    "no_wolves": ({"wolves": False, "starting_food": 0.5, "reward_for_starving": -2.5}, True),
    "tiny_bushes": ({"max_berries_per_bush": 1, "bush_power": 8, "turns_to_fill_food": 2}, True),
}


def pick_action(rng, obs_grids, n_actions, greedy):
    """Random action; the 'greedy' policy mostly stays on a bush so eating / depletion paths are exercised."""
    if greedy and obs_grids[1][5, 5] == 1 and rng.random() < 0.75:
        return 4
    return int(rng.integers(0, n_actions))
