"""Generate the committed golden traces in ``tests/golden/`` from the REAL reference
(TEST INFRASTRUCTURE). Run in the build container, where ``/root/reference`` exists:

    python -m oracle.make_golden            # all option sets, in parallel
    python -m oracle.make_golden defaults   # one set
    python -m oracle.make_golden --sized    # viewports other than 11x11 / spawn margins other than 1

Each trace is produced by the unmodified ``/root/reference/wab_env.py`` under ``oracle/ref_shim``
(keyed draws, pandas-3 compatibility) driven by a seeded action trace, and records after every
``reset()`` / ``step()`` the observation, reward, done flag and the hidden state. The fixtures let the
oracle and the CUDA kernels be checked against the reference where it cannot travel (the GPU box).
"""
import json
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")
WOLF_PAD = 16
SENTINEL = -(2 ** 15)

#: name -> (option overrides, seed, env_id, n_events, policy)
OPTION_SETS = {
    "defaults": ({}, 0, 0, 900, "random"),
    "defaults_greedy": ({}, 3, 4097, 700, "greedy"),
    "six_actions_random_start": (
        {"lookout_only": False, "starting_role": None, "starting_food": None}, 11, 5, 700, "greedy"),
    "restrict_view": ({"lookout_only": False, "restrict_view": True, "starting_role": None}, 5, 77, 600, "greedy"),
    "dense": ({"chance_wolf_on_square": 0.012, "bush_power": 12, "wolf_chance_to_despawn": 0.2,
               "reward_per_turn": 0.25, "reward_for_eating": 0.5}, 9, 1, 600, "random"),
    "gatherer_static_wolves": ({"gatherer_only": True, "wolves_can_move": False, "chance_wolf_on_square": 0.004},
                               2, 123456, 500, "greedy"),
    "god_mode_short": ({"god_mode": True, "max_turns": 30, "turns_to_fill_food": 4, "turns_to_empty_food": 20,
                        "chance_wolf_on_square": 0.01, "max_berries_per_bush": 3, "bush_power": 30},
                       21, 9, 500, "greedy"),
    "no_wolves": ({"wolves": False, "starting_food": 0.5, "reward_for_starving": -2.5}, 4, 31, 400, "greedy"),
}


#: viewports other than 11x11 and spawn margins other than 1 (the reference is generic in both: wab_env.py:25-26, :34,
#: :147-148, :510-576) -> tests/golden/sized_<name>.npz, same record layout
SIZED_SETS = {
    "5x5_m1": ({"width": 5, "height": 5, "chance_wolf_on_square": 0.01, "bush_power": 20}, 31, 3, 400, "greedy"),
    "7x9_m1": ({"width": 7, "height": 9, "chance_wolf_on_square": 0.004, "lookout_only": False, "starting_role": None}, 32, 8, 400, "greedy"),
    "11x11_m2": ({"wolf_spawn_margin": 2, "chance_wolf_on_square": 0.002}, 33, 1, 400, "random"),
    "15x13_m2": ({"width": 15, "height": 13, "wolf_spawn_margin": 2, "chance_wolf_on_square": 0.002, "bush_power": 40}, 34, 21, 350, "greedy"),
    "31x31_m2": ({"width": 31, "height": 31, "wolf_spawn_margin": 2, "chance_wolf_on_square": 0.0005, "max_turns": 60}, 35, 6, 170, "greedy"),
}


def bush_digest(bushes):
    """Order-free 64-bit digest of the bush record map {(x, y): food}."""
    h = 0
    for (x, y), f in bushes.items():
        v = ((x & 0xFFFF) | ((y & 0xFFFF) << 16) | ((f & 0xFFFF) << 32)) * 0x9E3779B97F4A7C15 & 0xFFFFFFFFFFFFFFFF
        v ^= v >> 29
        h = (h + v * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    return h


def trace(name):
    from . import ref_shim

    sized = name in SIZED_SETS
    overrides, seed, env_id, n_events, policy = (SIZED_SETS if sized else OPTION_SETS)[name]
    env = ref_shim.make_env(overrides, seed=seed, env_id=env_id)
    ci, cj = env.game_options["width"] // 2, env.game_options["height"] // 2
    n_act = env.action_space.n
    rng = np.random.default_rng(seed * 7919 + env_id)
    rec = {k: [] for k in ("action", "grids", "food", "role", "status", "reward", "done", "x", "y", "food_f64",
                           "turn", "episode", "n_wolves", "wolves", "n_bushes", "bush_digest")}

    def push(action, obs, reward, done):
        hs = ref_shim.hidden_state(env)
        rec["action"].append(action)
        rec["grids"].append(np.stack([np.asarray(obs[p]) for p in range(3)]).astype(np.uint8))
        rec["food"].append(int(obs[3])); rec["role"].append(int(obs[4])); rec["status"].append(int(obs[5]))
        rec["reward"].append(float(reward)); rec["done"].append(int(bool(done)))
        rec["x"].append(hs["x"]); rec["y"].append(hs["y"]); rec["food_f64"].append(hs["food"])
        rec["turn"].append(hs["turn"]); rec["episode"].append(env._wab_episode)
        w = np.full((WOLF_PAD, 2), SENTINEL, dtype=np.int32)
        assert len(hs["wolves"]) <= WOLF_PAD, (name, len(hs["wolves"]))
        for k, xy in enumerate(hs["wolves"]):
            w[k] = xy
        rec["n_wolves"].append(len(hs["wolves"])); rec["wolves"].append(w)
        rec["n_bushes"].append(len(hs["bushes"])); rec["bush_digest"].append(bush_digest(hs["bushes"]))

    # event 0: the reset inside the constructor (episode 0)
    push(-1, env._get_obs(), 0.0, False)
    obs = rec["grids"][-1]
    done, linger = False, 0
    while len(rec["action"]) < n_events:
        if done and linger == 0:
            o = env.reset()
            push(-1, o, 0.0, False)
            obs, done = rec["grids"][-1], False
            continue
        if policy == "greedy" and obs[1][ci, cj] == 1 and rng.random() < 0.75:
            a = 4  # stay on the bush (and, with 6 actions, become gatherer)
        else:
            a = int(rng.integers(0, n_act))
        o, r, d, _ = env.step(a)
        push(a, o, r, d)
        obs = rec["grids"][-1]
        if d and not done:
            linger = int(rng.integers(0, 3)) if rng.random() < 0.3 else 0  # keep stepping a dead/finished env
        elif done:
            linger -= 1
        done = d
    out = {k: np.asarray(v) for k, v in rec.items()}
    out["bush_digest"] = np.asarray(rec["bush_digest"], dtype=np.uint64)
    out["meta"] = np.array(json.dumps({"name": name, "overrides": overrides, "seed": seed, "env_id": env_id,
                                       "n_actions": n_act, "policy": policy}))
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, ("sized_%s.npz" if sized else "trace_%s.npz") % name), **out)
    return name, len(rec["action"]), int(np.sum(out["done"])), int(np.max(out["n_wolves"]))


def main(argv):
    names = list(SIZED_SETS) if argv == ["--sized"] else (argv or list(OPTION_SETS))
    with mp.get_context("spawn").Pool(min(len(names), os.cpu_count() or 1)) as pool:
        for res in pool.imap_unordered(trace, names):
            print("golden trace %-28s events=%d dones=%d max_wolves=%d" % res, flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
