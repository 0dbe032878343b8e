#!/usr/bin/env python
"""Volume parity run for Environment 2.0: the UNMODIFIED reference ("/root/reference/Environment 2.0" under
oracle/ref_shim/v2.py, keyed draws) against the C oracle (oracle/wab2_oracle.c), entity action by entity action,
over many independent random worlds — test infrastructure, build container only.

The reference spends ~10-15 ms per entity action (get_obs + take_action on pandas frames), so a >= 10^6-action
comparison cannot run inside the test suite; this script runs it in the background over several processes and
appends one JSON line per completed world (a partial run is still evidence). ``--summary`` folds the lines.

    python -m oracle.reference_volume_check_v2 --workers 4 --hours 6 --out profiles/r2_reference_parity_volume_v2.jsonl
    python -m oracle.reference_volume_check_v2 --summary profiles/r2_reference_parity_volume_v2.jsonl

Per entity action (every entity, bushes included, in id order, as the reference driver loop does,
Env2Tests.py:46-88): the rows of get_obs as one-hot planes + the row count, the internal observation, reward, done;
after every world turn: the whole entity table (type, object x/y, table X/Y, Visible, food, role, status).
Worlds: random sizes 5..40 plus the shapes the CUDA kernels special-case (19x21: smallest world whose windows fit
once; 33x64; 64x64; 20x20 = BASELINE config 3), random populations, radii, starting role and food constants.
"""
import argparse
import hashlib
import json
import multiprocessing as mp
import os
import random
import sys
import time
import warnings

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

TYPES = {"Ostrich": 0, "Wolf": 1, "Bush": 2}
SPECIAL = [(19, 21), (33, 64), (64, 64), (20, 20), (21, 19), (64, 19)]


def random_world(rng):
    if rng.random() < 0.4:
        W, H = rng.choice(SPECIAL)
    else:
        W, H = rng.randint(5, 40), rng.randint(5, 40)
    crowded = rng.random() < 0.5
    cap = max(3, (W * H) // (6 if crowded else 30))
    no, nw, nb = rng.randint(1, min(12, cap)), rng.randint(1, min(14, cap)), rng.randint(0, min(30, cap))
    opts = {"lookout_view_radius": rng.randint(2, 9), "gatherer_view_radius": rng.randint(1, 7), "wolf_view_radius": rng.randint(1, 8),
            "starting_role": rng.randint(0, 1), "food_per_bush": rng.choice([20, 7, 3, 12]), "food_given_per_turn": rng.choice([5, 2, 4]),
            "wolf_food_for_eating_ostrich": rng.choice([10, 3]), "wolf_starting_food": rng.choice([20, 5, 9]),
            "ostrich_starting_food": float(rng.choice([40, 11]))}
    return W, H, no, nw, nb, opts


def worker(wid, deadline, out_path, lock):
    import numpy as np
    from oracle.ref_shim import v2 as ref_v2
    from oracle.wab2_oracle import OracleWorld2
    warnings.simplefilter("ignore")
    chunk = 0
    while time.time() < deadline:
        rng = random.Random(9_000_011 * wid + chunk)
        W, H, no, nw, nb, opts = random_world(rng)
        seed, env_id = rng.getrandbits(62), rng.getrandbits(31)
        n = no + nw + nb
        R = max(opts["lookout_view_radius"], opts["gatherer_view_radius"], opts["wolf_view_radius"])
        S = 2 * R + 1
        ref = ref_v2.make_env(W, H, no, nw, nb, game_options=opts, seed=seed, env_id=env_id)
        orc = OracleWorld2(W, H, no, nw, nb, game_options=opts, seed=seed, env_id=env_id, window_radius=R)
        sha = hashlib.sha256()
        actions = kills = 0
        error = None
        t0 = time.time()

        def state_equal(tag):
            want = np.array([[TYPES[t], x, y, tx, ty, int(v), f, r, s] for (t, x, y, tx, ty, v, f, r, s) in ref_v2.hidden_state(ref)],
                            dtype=np.float64)
            got = orc.state()
            if not np.array_equal(want, got):
                return "%s: entity table differs" % tag
            sha.update(got.tobytes())
            return None

        error = state_equal("create")
        episodes = rng.randint(1, 3)
        turns = max(2, min(60, 1500 // n))
        for ep in range(episodes):
            if error or time.time() >= deadline:
                break
            ref.reset_environment(); orc.reset_environment()
            error = state_equal("reset %d" % ep)
            for turn in range(turns):
                if error or time.time() >= deadline:
                    break
                for i in range(n):
                    a = rng.randint(0, 5) if i < no else (rng.randint(0, 4) if i < no + nw else 0)
                    df, internal = ref.get_obs(i)
                    planes = np.zeros((3, S, S), np.uint8)
                    for _, row in df.iterrows():
                        planes[TYPES[row["Type"]], int(row["Delta_X"]) + R, int(row["Delta_Y"]) + R] = 1
                    po, io, rows = orc.get_obs(i)
                    if rows != len(df) or not np.array_equal(planes, po):
                        error = "ep %d turn %d entity %d: observation differs" % (ep, turn, i); break
                    if [float(v) for v in internal] + [0.0] * (5 - len(internal)) != [float(v) for v in io]:
                        error = "ep %d turn %d entity %d: internal obs differs" % (ep, turn, i); break
                    rr, rd = ref.take_action(i, a)
                    orr, od = orc.take_action(i, a)
                    if float(rr) != float(orr) or bool(rd) != bool(od):
                        error = "ep %d turn %d entity %d: reward/done differ" % (ep, turn, i); break
                    actions += 1
                    sha.update(np.packbits(po).tobytes())
                if error is None:
                    error = state_equal("ep %d turn %d" % (ep, turn))
            kills += int((orc.state()[:no, 8] == 2).sum())
        line = {"worker": wid, "chunk": chunk, "world": [W, H, no, nw, nb], "options": opts, "seed": seed, "env_id": env_id,
                "entity_actions": actions, "kills": kills, "ok": error is None, "error": error, "sha256": sha.hexdigest(),
                "seconds": round(time.time() - t0, 2)}
        with lock:
            with open(out_path, "a") as f:
                f.write(json.dumps(line) + "\n")
        if error is not None:
            return
        chunk += 1


def summary(path):
    rows = [json.loads(l) for l in open(path) if l.strip()]
    shapes = {}
    for r in rows:
        key = "%dx%d" % tuple(r["world"][:2]) if tuple(r["world"][:2]) in SPECIAL else "random 5..40"
        shapes[key] = shapes.get(key, 0) + r["entity_actions"]
    digest = hashlib.sha256("".join(sorted(r["sha256"] for r in rows)).encode()).hexdigest()
    out = {"subject": "unmodified '/root/reference/Environment 2.0' under oracle/ref_shim/v2.py (keyed draws) vs oracle/wab2_oracle.c, "
                      "every entity action: observation rows (planes + count), internal obs, reward, done; every turn: entity table",
           "worlds": len(rows), "entity_actions": sum(r["entity_actions"] for r in rows), "kills": sum(r["kills"] for r in rows),
           "mismatches": [r for r in rows if not r["ok"]], "all_equal": all(r["ok"] for r in rows),
           "entity_actions_by_world_shape": dict(sorted(shapes.items())),
           "cpu_seconds": round(sum(r["seconds"] for r in rows), 1), "digest_of_world_digests": digest}
    print(json.dumps(out, indent=1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=4)
    ap.add_argument("--worker-offset", type=int, default=0)
    ap.add_argument("--hours", type=float, default=6.0)
    ap.add_argument("--out", default=os.path.join(REPO, "profiles", "r2_reference_parity_volume_v2.jsonl"))
    ap.add_argument("--summary", default=None)
    args = ap.parse_args()
    if args.summary:
        return summary(args.summary)
    deadline = time.time() + args.hours * 3600
    lock = mp.Lock()
    procs = [mp.Process(target=worker, args=(args.worker_offset + w, deadline, args.out, lock)) for w in range(args.workers)]
    for p in procs:
        p.start()
    for p in procs:
        p.join()


if __name__ == "__main__":
    main()
