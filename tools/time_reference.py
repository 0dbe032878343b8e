#!/usr/bin/env python
"""Time the REAL reference (unmodified /root/reference/wab_env.py under oracle/ref_shim) on this host:
single process and one process per core (BASELINE.md "CPU-baseline plan"). Build-container only — the
reference cannot travel to the GPU box; the numbers are committed under profiles/.

    python tools/time_reference.py [--steps 150]
"""
import argparse
import json
import multiprocessing as mp
import os
import platform
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def worker(args):
    rank, steps, keyed = args
    try:
        os.sched_setaffinity(0, {rank})
    except Exception:
        pass
    import warnings
    import numpy as np
    warnings.filterwarnings("ignore")
    from oracle import ref_shim
    env = ref_shim.make_env(seed=rank, env_id=rank)
    actions = np.random.default_rng(12345 + rank).integers(0, env.action_space.n, steps)
    env.reset()
    t0 = time.perf_counter()
    for a in actions:
        _, _, done, _ = env.step(int(a))
        if done:
            env.reset()
    return steps / (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=150)
    args = ap.parse_args()
    import numpy
    import pandas
    cores = len(os.sched_getaffinity(0))
    single = worker((0, args.steps, True))
    with mp.get_context("spawn").Pool(cores) as pool:
        per = pool.map(worker, [(r, args.steps, True) for r in range(cores)])
    print(json.dumps({
        "subject": "unmodified /root/reference/wab_env.py WolvesAndBushesEnv (default options, actions uniform 0-4, reset on done), "
                   "oracle/ref_shim active (stub gym, pandas-3 compat frame, keyed np.random)",
        "steps_per_process": args.steps, "single_process_steps_per_s": single, "cores": cores,
        "per_core_steps_per_s": per, "aggregate_steps_per_s": sum(per),
        "python": platform.python_version(), "numpy": numpy.__version__, "pandas": pandas.__version__,
        "cpu": platform.processor() or open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t"),
        "note": "reported baseline, not the target"}))


if __name__ == "__main__":
    main()
