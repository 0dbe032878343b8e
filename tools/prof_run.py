#!/usr/bin/env python
"""A short, fixed scenario for ncu: L launches of T lockstep steps of n v1 envs (or v2 world turns).

    python tools/prof_run.py --num-envs 131072 --steps 20 --launches 6
    python tools/prof_run.py --v2 config4 --num-envs 16384 --launches 4
"""
import argparse
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--launches", type=int, default=6)
    ap.add_argument("--v2", default="", choices=["", "config3", "config4"])
    ap.add_argument("--features", action="store_true")
    args = ap.parse_args()
    n, T = args.num_envs, args.steps
    if args.v2:
        from wab_gym_b200.world2 import VecWorld2
        W, H, no, nw, nb = (64, 64, 8, 64, 256) if args.v2 == "config4" else (20, 20, 10, 3, 20)
        env = VecWorld2(n, W, H, no, nw, nb, seed=0)
        env.reset_environment()
        gen = torch.Generator(device="cuda").manual_seed(1)
        acts = torch.empty((no + nw, n), dtype=torch.uint8, device="cuda")
        for _ in range(args.launches):
            acts[:no] = torch.randint(0, 6, (no, n), dtype=torch.uint8, device="cuda", generator=gen)
            acts[no:] = torch.randint(0, 5, (nw, n), dtype=torch.uint8, device="cuda", generator=gen)
            env.turn(acts)
        torch.cuda.synchronize()
        print("ok", args.v2, n, args.launches)
        return
    from wab_gym_b200 import VecEnv
    env = VecEnv(n, seed=0, features=args.features)
    gen = torch.Generator(device="cuda").manual_seed(1)
    env.reset()
    # advance into the steady state (episodes of every age) before the profiled launches
    warm = torch.randint(0, env.n_actions, (100, n), dtype=torch.uint8, device="cuda", generator=gen)
    out = env._alloc(max(T, 1))
    for s in range(0, 100, T):
        c = min(T, 100 - s)
        env.step_many(warm[s:s + c], out={k: v[:c] for k, v in out.items()})
    for _ in range(args.launches):
        a = torch.randint(0, env.n_actions, (T, n), dtype=torch.uint8, device="cuda", generator=gen)
        if T == 1:
            env.step(a[0])
        else:
            env.step_many(a, out=out)
    torch.cuda.synchronize()
    print("ok", n, T, args.launches, env.stats())


if __name__ == "__main__":
    main()
