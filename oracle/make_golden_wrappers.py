"""Freeze outputs of the reference's consumer-side code into fixtures that travel to the GPU box (TEST INFRASTRUCTURE).

    python -m oracle.make_golden_wrappers

* ``tests/golden/features.npz`` — the live ``PragmaticObsWrapper.observation`` (wab_env.py:726-761, unmodified, under
  oracle/ref_shim) on 2,400 random observation tuples of every density plus the three known-answer inputs of the
  reference's own wab_env_test.py:9-169.
* ``tests/golden/render.npz`` — ``WolvesAndBushesEnv.render`` (wab_env.py:468-502) frames of keyed reference episodes
  (restricted view and full view; alive in both roles, killed, starved), with the action trace that leads to them.
"""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def features():
    from oracle import ref_shim
    from tests.test_features import REFERENCE_KATS, grid
    mod = ref_shim.load_reference()
    wrapper = mod.PragmaticObsWrapper(ref_shim.make_env())
    rng = np.random.default_rng(20261018)
    W, B, S, OUT = [], [], [], []
    cases = []
    for inp, _ in REFERENCE_KATS:
        wolves, bushes, food, role, status = inp
        cases.append((grid(wolves), grid(bushes), food, role, status))
    for trial in range(2400):
        density = rng.choice([0.0, 0.01, 0.02, 0.06, 0.15, 0.3, 0.6, 0.9, 1.0])
        wolves = (rng.random((11, 11)) < density * rng.random()).astype(float)
        bushes = (rng.random((11, 11)) < density).astype(float)
        if trial % 7 == 0:                       # ties on purpose: symmetric pairs around the centre
            i, j = int(rng.integers(0, 11)), int(rng.integers(0, 11))
            bushes[i, j] = bushes[10 - i, 10 - j] = bushes[j, i] = 1.0
            wolves[10 - j, i] = wolves[i, 10 - j] = 1.0
        cases.append((wolves, bushes, int(rng.integers(0, 41)), int(rng.integers(0, 2)), int(rng.integers(0, 3))))
    for wolves, bushes, food, role, status in cases:
        ref = wrapper.observation((wolves.copy(), bushes.copy(), np.zeros((11, 11)), food, role, status, np.zeros((11, 11))))
        out = list(ref[0]) + list(ref[1]) + [int(v) for v in ref[2]] + list(ref[3]) + list(ref[4]) + [int(v) for v in ref[5]] + \
            [int(ref[6]), int(ref[7]), int(ref[8]), int(ref[9])]
        W.append(np.packbits(wolves.astype(np.uint8))); B.append(np.packbits(bushes.astype(np.uint8)))
        S.append([food, role, status]); OUT.append(out)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "features.npz"), wolves=np.asarray(W), bushes=np.asarray(B),
                        scalars=np.asarray(S, dtype=np.uint8), features=np.asarray(OUT, dtype=np.uint8),
                        n_kats=np.int64(len(REFERENCE_KATS)))
    print("features golden:", len(OUT), "rows")


RENDER_CASES = [
    ("restrict_view", {"lookout_only": False, "restrict_view": True, "starting_role": None, "chance_wolf_on_square": 0.006}, 5),
    ("defaults", {}, 3),
]


def render():
    from oracle import ref_shim
    out = {}
    meta = []
    for name, opts, seed in RENDER_CASES:
        env = ref_shim.make_env(opts, seed=seed, env_id=2)
        rng = np.random.default_rng(seed)
        n_actions = env.action_space.n
        actions, frames, frame_steps, kinds = [], [], [], []
        want = {"alive0", "alive1", "killed", "starved"} if name == "restrict_view" else {"alive0", "starved"}
        frames.append(env.render(draw_health=False)); frame_steps.append(0); kinds.append("reset")
        t = 0
        done_any = 0
        while want and t < 3000:
            a = int(rng.integers(0, n_actions))
            if "starved" not in want and env.ostriches.iloc[0].food < 0.3 and rng.random() < 0.8:
                a = 4
            obs, r, done, _ = env.step(a)
            actions.append(a); t += 1
            status, role = int(obs[5]), int(obs[4])
            kind = {0: "alive%d" % role, 1: "starved", 2: "killed"}[status]
            if kind in want and (status != 0 or t % 5 == 3):
                want.discard(kind)
                frames.append(env.render(draw_health=False)); frame_steps.append(t); kinds.append(kind)
                frames.append(env.render(draw_health=True)); frame_steps.append(t); kinds.append(kind + "+health")
            if done:
                env.reset(); actions.append(-1); t += 1   # -1 marks a reset in the trace
                done_any += 1
        out[name + "_actions"] = np.asarray(actions, dtype=np.int8)
        out[name + "_frames"] = np.asarray(frames, dtype=np.uint8)
        out[name + "_frame_steps"] = np.asarray(frame_steps, dtype=np.int64)
        meta.append({"name": name, "options": opts, "seed": seed, "env_id": 2, "kinds": kinds, "missing": sorted(want)})
        print("render golden:", name, kinds, "events", len(actions), "missing", sorted(want))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "render.npz"), meta=json.dumps(meta), **out)




EGO_CASES = [("defaults", {}, False), ("dense", None, False), ("tiny_bushes", None, True)]


def ego():
    """``tests/golden/ego.npz`` — the reference's egocentric observation family (wab_env.py:637-667): after the reset and
    after every step of keyed reference runs, ``_get_wolf_proximities()`` and ``_get_bush_proximities()`` (the latter is
    element 0 of ``WolvesAndBushesEnvEgoCentric._get_obs``, :951-958), plus the actions (-1 = reset after done)."""
    from oracle import ref_shim
    from tests.util import OPTION_SETS, pick_action
    out, meta = {}, []
    for name, opts, _ in EGO_CASES:
        if opts is None:
            opts = OPTION_SETS[name][0]
        greedy = OPTION_SETS[name][1] if name in OPTION_SETS else False
        for env_id in (40, 41):
            env = ref_shim.make_env(opts, seed=17, env_id=env_id)
            env.max_distance = env.game_options["width"] // 2 + env.game_options["height"] // 2 + 1   # :932-934
            rng = np.random.default_rng(env_id)
            obs = env._get_obs()
            actions, prox = [], []
            prox.append([int(v) for v in env._get_wolf_proximities()] + [int(v) for v in env._get_bush_proximities()])
            for t in range(260):
                a = pick_action(rng, [np.asarray(obs[0]), np.asarray(obs[1])], env.action_space.n, greedy)
                obs, r, done, _ = env.step(a)
                actions.append(a)
                prox.append([int(v) for v in env._get_wolf_proximities()] + [int(v) for v in env._get_bush_proximities()])
                if done:
                    obs = env.reset()
                    actions.append(-1)
                    prox.append([int(v) for v in env._get_wolf_proximities()] + [int(v) for v in env._get_bush_proximities()])
            key = "%s_%d" % (name, env_id)
            out[key + "_actions"] = np.asarray(actions, dtype=np.int8)
            out[key + "_prox"] = np.asarray(prox, dtype=np.uint8)
            meta.append({"name": name, "options": opts, "seed": 17, "env_id": env_id, "key": key})
            print("ego golden:", key, len(actions), "events; bush proximities seen:", sorted(set(np.asarray(prox)[:, 5:].ravel().tolist())))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "ego.npz"), meta=json.dumps(meta), **out)


if __name__ == "__main__":
    import sys
    which = sys.argv[1:] or ["features", "render", "ego"]
    for w in which:
        {"features": features, "render": render, "ego": ego}[w]()
