"""Generate the Environment-2.0 golden fixtures from the REAL reference (TEST INFRASTRUCTURE).

    python -m oracle.make_golden_v2            # tests/golden/v2_trace.npz   (3 small worlds, round 1)
    python -m oracle.make_golden_v2 --long     # tests/golden/v2_long_*.npz  (9 worlds x >= 200 turns, round 2)

Runs "/root/reference/Environment 2.0" unmodified under ``oracle/ref_shim/v2.py`` (stub gym, keyed
``random.randint``) and records, per entity action, the observation list (as one-hot planes + row count),
internal observation, reward, done, and the full entity table after every turn. The long set covers the world
shapes the CUDA kernels special-case (19x21 — the smallest world every 19x19 window fits once, where the strict
``size < entity + radius`` test of World.py:264/:285 bites —, 33x64, 64x64, the BASELINE config-3 and config-4
populations) and non-default radii / roles / food constants; seeds are chosen (with the C oracle, which is cheap)
so that every trace contains kills."""
import json
import os
import sys
import warnings

import numpy as np

# (W, H, ostriches, wolves, bushes, env_id, episodes, turns, option overrides)
LONG_WORLDS = [
    (20, 20, 10, 3, 20, 0, 2, 100, {}),                                                   # BASELINE config 3 population
    (19, 21, 6, 6, 10, 11, 2, 100, {}),
    (21, 19, 5, 8, 6, 2, 2, 100, {"starting_role": 0}),
    (33, 64, 12, 20, 20, 5, 2, 100, {}),
    (64, 64, 6, 20, 30, 3, 2, 100, {"lookout_view_radius": 7, "gatherer_view_radius": 3, "wolf_view_radius": 8}),
    (64, 64, 8, 64, 256, 0, 1, 40, {}),                                                   # BASELINE config 4 population
    (7, 9, 6, 4, 5, 3, 2, 100, {"food_per_bush": 7, "food_given_per_turn": 2}),
    (12, 5, 3, 6, 2, 7, 2, 100, {"wolf_food_for_eating_ostrich": 3, "wolf_starting_food": 5}),
    (40, 11, 8, 10, 12, 9, 2, 100, {"starting_role": 0, "lookout_view_radius": 5, "gatherer_view_radius": 2}),
]
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def long_path(k):
    return os.path.join(GOLDEN_DIR, "v2_long_%d.npz" % k)


def find_seed_with_kills(world):
    """Cheap search with the C oracle for a seed whose trace has kills in every episode (same action stream as the recorder)."""
    import random
    from oracle.wab2_oracle import OracleWorld2
    from tests.test_v2 import actions_for
    W, H, no, nw, nb, env_id, episodes, turns, opts = world
    n = no + nw + nb
    for seed in range(1, 400):
        orc = OracleWorld2(W, H, no, nw, nb, game_options=opts, seed=seed, env_id=env_id)
        rng = random.Random(seed)
        kills = []
        for ep in range(episodes):
            orc.reset_environment()
            for t in range(turns):
                acts = actions_for(rng, no, nw, n)
                for i in range(n):
                    orc.take_action(i, acts[i])
            kills.append(int((orc.state()[:no, 8] == 2).sum()))
        if min(kills) > 0:
            return seed, kills
    raise RuntimeError("no seed with kills found for %r" % (world,))


def make_long(which):
    from tests.test_v2 import record_reference
    for k in which:
        world = LONG_WORLDS[k]
        W, H, no, nw, nb, env_id, episodes, turns, opts = world
        seed, kills = find_seed_with_kills(world)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            rec = record_reference((W, H, no, nw, nb, seed, env_id), episodes, turns, game_options=opts)
        meta = {"world": [W, H, no, nw, nb], "seed": seed, "env_id": env_id, "episodes": episodes, "turns": turns,
                "options": opts, "kills_per_episode": kills, "entity_actions": int(len(rec["reward"]))}
        np.savez_compressed(long_path(k), meta=json.dumps(meta), **rec)
        print("v2 long golden", k, meta, flush=True)


def main():
    if "--long" in sys.argv:
        rest = [int(a) for a in sys.argv[1:] if a.isdigit()]
        return make_long(rest or range(len(LONG_WORLDS)))
    from tests.test_v2 import GOLDEN, WORLDS, record_reference
    episodes, turns = 3, 10
    out = {"n_worlds": np.int64(len(WORLDS)), "worlds": np.asarray(WORLDS, dtype=np.int64), "episodes": np.int64(episodes),
           "turns": np.int64(turns)}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for w, world in enumerate(WORLDS):
            rec = record_reference(world, episodes, turns)
            for k, v in rec.items():
                out["w%d_%s" % (w, k)] = v
            print("v2 golden world", world, "events", len(rec["reward"]), "kills", int((rec["state"][-1][:, 8] == 2).sum()), flush=True)
    os.makedirs(os.path.dirname(GOLDEN), exist_ok=True)
    np.savez_compressed(GOLDEN, **out)


if __name__ == "__main__":
    main()
