#!/usr/bin/env python
"""Features-only stepping (VecEnv(features=True, emit_grids=False)) against full stepping: T-step launches of n envs.

    python tools/bench_features_only.py [--num-envs 131072] [--steps 60] [--launches 20]
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import torch
    from wab_gym_b200 import VecEnv
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-envs", type=int, default=131072)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--launches", type=int, default=20)
    args = ap.parse_args()
    n, T = args.num_envs, args.steps
    gen = torch.Generator(device="cuda").manual_seed(1)
    acts = torch.randint(0, 5, (T, n), dtype=torch.uint8, device="cuda", generator=gen)
    out = {}
    for name, kw in (("full", dict(features=True)), ("features_only", dict(features=True, emit_grids=False)), ("grids_only", dict())):
        env = VecEnv(n, seed=0, **kw)
        env.reset()
        for _ in range(3):
            env.step_many(acts)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.launches):
            env.step_many(acts)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        out[name] = {"env_steps_per_s": n * T * args.launches / (ms * 1e-3), "us_per_step": ms * 1e3 / (T * args.launches),
                     "kernel": env.step_kernel_name(T)}
        env.close()
    print(json.dumps({"num_envs": n, "steps_per_launch": T, "launches": args.launches, **out}))


if __name__ == "__main__":
    main()
