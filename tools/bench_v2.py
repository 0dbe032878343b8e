#!/usr/bin/env python
"""BASELINE config 3: Environment 2.0 World(20, 20) with 10 ostriches, 3 wolves, 20 bushes (Env2Tests.py:7-22),
65,536 lockstep worlds on one B200; unit of work = one world turn (33 sequential entity actions, 13 observations).

    python tools/bench_v2.py [--num-envs 65536] [--turns 200] [--config4]
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import torch
    from wab_gym_b200.world2 import VecWorld2
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-envs", type=int, default=65536)
    ap.add_argument("--turns", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--config4", action="store_true", help="World(64, 64), 8 ostriches, 64 wolves, 256 bushes")
    ap.add_argument("--no-obs", action="store_true")
    args = ap.parse_args()
    dims = (64, 64, 8, 64, 256) if args.config4 else (20, 20, 10, 3, 20)
    W, H, no, nw, nb = dims
    n, A = args.num_envs, no + nw
    env = VecWorld2(n, W, H, no, nw, nb, seed=0, observations=not args.no_obs)
    env.reset_environment()
    gen = torch.Generator(device="cuda").manual_seed(1)
    acts = torch.empty((8, A, n), dtype=torch.uint8, device="cuda")      # entity-major, like every v2 array
    acts[:, :no] = torch.randint(0, 6, (8, no, n), dtype=torch.uint8, device="cuda", generator=gen)
    acts[:, no:] = torch.randint(0, 5, (8, nw, n), dtype=torch.uint8, device="cuda", generator=gen)
    for t in range(args.warmup):
        env.turn(acts[t % 8])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(args.turns):
        env.turn(acts[t % 8])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    S = 2 * env.R + 1
    bytes_per_turn = A * (3 * S * S + 9) + 2 * (no + nw + nb) * 8 if not args.no_obs else A * 9 + 2 * (no + nw + nb) * 8
    turns_per_s = n * args.turns / (ms * 1e-3)
    try:
        peak = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    print(json.dumps({"workload": "v2 World(%d,%d) %d ostriches %d wolves %d bushes, %d lockstep worlds" % (W, H, no, nw, nb, n),
                      "world_turns_per_s": turns_per_s, "entity_steps_per_s": turns_per_s * (no + nw + nb),
                      "ms_per_turn": ms / args.turns, "algorithmic_bytes_per_turn": bytes_per_turn,
                      "achieved_gbs": turns_per_s * bytes_per_turn / 1e9, "peak_gbs": peak,
                      "frac": turns_per_s * bytes_per_turn / 1e9 / peak, "observations": not args.no_obs}))
    env.close()


if __name__ == "__main__":
    main()
