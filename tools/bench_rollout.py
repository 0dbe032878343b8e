#!/usr/bin/env python
"""BASELINE config 5: actor-critic rollout — the reference's policy network (actor_critic.py:54-97) consuming
device-resident PragmaticObsWrapper observations of N v1 environments (32,768 per GPU in the 8-GPU config).

    python tools/bench_rollout.py [--num-envs 32768] [--steps 300] [--graph] [--bf16]
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
    from wab_gym_b200.policy import Policy, Rollout
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-envs", type=int, default=32768)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--bf16", action="store_true")
    ap.add_argument("--no-tc", action="store_true", help="first layer as flatten kernel + library GEMM instead of the tcgen05 kernel")
    ap.add_argument("--no-trunk", action="store_true", help="tcgen05 first layer only; affine2 / affine3 as library GEMMs")
    args = ap.parse_args()
    torch.manual_seed(0)
    env = VecEnv(args.num_envs, seed=0, features=True)
    ro = Rollout(env, Policy(env.flat_dim, env.n_actions), use_graph=args.graph,
                 dtype=torch.bfloat16 if args.bf16 else torch.float32, tc_first_layer=False if (args.no_tc or args.bf16) else None,
                 tc_trunk=False if (args.no_tc or args.bf16 or args.no_trunk) else None)
    ro.run(args.warmup)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ro.run(args.steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = env.stats()
    print(json.dumps({"workload": "config 5: actor-critic rollout, %d v1 envs on 1 GPU, policy 449-128-150-128-{5,1} %s, %s" % (
                          args.num_envs, "bf16" if args.bf16 else "fp32", "CUDA graph" if args.graph else "eager"),
                      "path": ro.describe(), "env_steps_per_s": args.num_envs * args.steps / (ms * 1e-3), "ms_per_step": ms / args.steps,
                      "episodes": st["episodes"], "mean_episode_length": st["steps"] / max(st["episodes"], 1),
                      "finished": st["finished"], "starved": st["starved"], "killed": st["killed"]}))
    env.close()


if __name__ == "__main__":
    main()
