#!/usr/bin/env python
"""Volume parity run: the UNMODIFIED reference (wab_env.py under oracle/ref_shim, keyed draws) against the C oracle,
step by step, over many independent (option set, seed, env id) streams — test infrastructure, build container only.

The reference does ~15 steps/s per core, so the >= 10^6-step comparison BASELINE.json asks for cannot run inside
the test suite; this script runs it in the background over several processes and appends one JSON line per
completed chunk (so a partial run is still evidence). ``--summary`` folds the lines into one digest.

    python -m oracle.reference_volume_check --workers 6 --hours 3 --out profiles/r1_reference_parity_volume.jsonl
    python -m oracle.reference_volume_check --summary profiles/r1_reference_parity_volume.jsonl

Every step compares the three 11x11 grids, food, role, status, reward (as float64, exact) and done; every tenth
step also the hidden state (position, float64 food, turn, wolf multiset, every bush record).
"""
import argparse
import hashlib
import json
import multiprocessing as mp
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def worker(wid, deadline, out_path, chunk_steps, lock):
    import numpy as np
    from oracle import ref_shim
    from oracle.wab_oracle import OracleEnv
    from tests.test_fuzz_options import random_options
    from tests.util import OPTION_SETS, pick_action
    names = sorted(OPTION_SETS)
    chunk = 0
    while time.time() < deadline:
        rng = np.random.default_rng(7_000_003 * wid + chunk)
        if chunk % 3 == 2:                                   # every third chunk: a randomised option set
            name, greedy = "random_options(%d)" % (wid * 100000 + chunk), bool(chunk & 1)
            opts = random_options(np.random.default_rng(wid * 100000 + chunk))
        else:
            name = names[(chunk // 3 * 2 + chunk % 3 + wid) % len(names)]
            opts, greedy = OPTION_SETS[name]
        seed, env_id = int(rng.integers(0, 2 ** 62)), int(rng.integers(0, 2 ** 31))
        ref = ref_shim.make_env(opts, seed=seed, env_id=env_id)
        orc = OracleEnv(opts, seed=seed, env_id=env_id)
        o_obs, r_obs = orc.reset(), ref._get_obs()
        sha = hashlib.sha256()
        steps = episodes = 0
        error = None
        done = False
        t0 = time.time()

        def same(tag, hidden):
            for p in range(3):
                if not np.array_equal(np.asarray(r_obs[p]).astype(np.uint8), o_obs[0][p]):
                    return "%s: grid %d differs" % (tag, p)
            if (int(r_obs[3]), int(r_obs[4]), int(r_obs[5])) != tuple(o_obs[1:]):
                return "%s: food/role/status differ" % (tag,)
            if hidden:
                hr, ho = ref_shim.hidden_state(ref), orc.hidden_state()
                for k in ("x", "y", "food", "role", "status", "turn", "wolves", "bushes"):
                    if hr[k] != ho[k]:
                        return "%s: hidden %s differs" % (tag, k)
            return None

        error = same("init", True)
        while error is None and steps < chunk_steps and time.time() < deadline:
            if done:
                r_obs, o_obs = ref.reset(), orc.reset()
                episodes += 1
                error = same("reset@%d" % steps, True)
                if error:
                    break
            a = pick_action(rng, o_obs[0], orc.n_actions, greedy)
            r_obs, rr, done, _ = ref.step(a)
            o_obs, orr, od = orc.step(a)
            steps += 1
            if float(rr) != orr or bool(done) != od:
                error = "step %d: reward/done differ (%r vs %r)" % (steps, rr, orr)
                break
            error = same("step %d action %d" % (steps, a), steps % 10 == 0)
            sha.update(o_obs[0].tobytes())
            sha.update(bytes((o_obs[1] & 0xFF, o_obs[2], o_obs[3], int(od))))
        line = {"worker": wid, "chunk": chunk, "options": name, "seed": seed, "env_id": env_id, "steps": steps,
                "episodes": episodes, "ok": error is None, "error": error, "obs_sha256": sha.hexdigest(),
                "seconds": round(time.time() - t0, 2)}
        with lock:
            with open(out_path, "a") as f:
                f.write(json.dumps(line) + "\n")
        if error is not None:
            return
        chunk += 1


def summary(path):
    rows = [json.loads(l) for l in open(path) if l.strip()]
    sets = {}
    for r in rows:
        key = "random_options" if r["options"].startswith("random_options") else r["options"]
        s = sets.setdefault(key, [0, 0])
        s[0] += r["steps"]; s[1] += r["episodes"]
    digest = hashlib.sha256("".join(sorted(r["obs_sha256"] for r in rows)).encode()).hexdigest()
    out = {"subject": "unmodified /root/reference/wab_env.py under oracle/ref_shim (keyed draws) vs oracle/wab_oracle.c, "
                      "every step: grids, food, role, status, reward, done; every 10th step: full hidden state",
           "chunks": len(rows), "steps": sum(r["steps"] for r in rows), "episodes": sum(r["episodes"] for r in rows),
           "mismatches": [r for r in rows if not r["ok"]], "all_equal": all(r["ok"] for r in rows),
           "steps_by_option_set": {k: v[0] for k, v in sorted(sets.items())},
           "cpu_seconds": round(sum(r["seconds"] for r in rows), 1), "digest_of_chunk_digests": digest}
    print(json.dumps(out, indent=1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workers", type=int, default=6)
    ap.add_argument("--worker-offset", type=int, default=0, help="first worker id (a second run appends new streams)")
    ap.add_argument("--hours", type=float, default=3.0)
    ap.add_argument("--chunk-steps", type=int, default=600)
    ap.add_argument("--out", default=os.path.join(REPO, "profiles", "r1_reference_parity_volume.jsonl"))
    ap.add_argument("--summary", default=None)
    args = ap.parse_args()
    if args.summary:
        return summary(args.summary)
    deadline = time.time() + args.hours * 3600
    lock = mp.Lock()
    procs = [mp.Process(target=worker, args=(args.worker_offset + w, deadline, args.out, args.chunk_steps, lock))
             for w in range(args.workers)]
    for p in procs:
        p.start()
    for p in procs:
        p.join()


if __name__ == "__main__":
    main()
