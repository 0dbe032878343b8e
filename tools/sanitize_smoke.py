#!/usr/bin/env python
"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck):

    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import torch
    from wab_gym_b200 import VecEnv
    from wab_gym_b200.world2 import VecWorld2
    gen = torch.Generator(device="cuda").manual_seed(0)
    for lpe in ("1", "4", "8", "16", "32"):
        os.environ["WAB_LPE"] = lpe
        for n in (77, 128):
            env = VecEnv(n, seed=3, features=True, game_options={"chance_wolf_on_square": 0.02, "restrict_view": True, "lookout_only": False})
            env.reset()
            acts = torch.randint(0, 6, (12, n), dtype=torch.uint8, device="cuda", generator=gen)
            for t in range(6):
                env.step(acts[t])
            env.step_many(acts)
            mask = torch.zeros(n, dtype=torch.uint8, device="cuda"); mask[::2] = 1
            env.reset(mask)
            hb = env.alloc_host_buffers()
            env.step_host(hb)
            env.flatten_features(env.last_features)
            torch.cuda.synchronize()
            assert env.stats()["steps"] == n * 19
            env.close()
    os.environ.pop("WAB_LPE")
    w = VecWorld2(70, 7, 9, 6, 4, 5, seed=2)
    w.reset_environment()
    a = torch.randint(0, 5, (10, 70), dtype=torch.uint8, device="cuda", generator=gen)
    for _ in range(5):
        w.turn(a)
    torch.cuda.synchronize()
    w.close()
    print("sanitize smoke ok")


if __name__ == "__main__":
    main()
