"""Scratch: reproduce tests/test_vecenv_gpu.py::test_batch_without_auto_reset_keeps_reporting_done with details."""
import numpy as np, torch
from oracle.wab_oracle import OracleEnv
from tests.util import OPTION_SETS, pick_action
from wab_gym_b200 import VecEnv

overrides, greedy = OPTION_SETS["defaults"]
n, steps, seed = 48, 220, 13
env = VecEnv(n, overrides, seed=seed, auto_reset=False, wolf_cap=15)
print("lpe", env.lib.wab_vec_lanes_per_env(env._h))
oracles = [OracleEnv(overrides, seed=seed, env_id=i) for i in range(n)]
rng = np.random.default_rng(3)
obs = env.reset()
cur = [o.reset() for o in oracles]
done_now = np.zeros(n, bool)
last_mask = None
bad = 0
for t in range(steps):
    g, f, r, s = (x.cpu().numpy() for x in obs)
    st = env.export_state()
    for i in range(n):
        if not (np.array_equal(g[i], cur[i][0]) and (int(f[i]), int(r[i]), int(s[i])) == cur[i][1:]):
            hs = oracles[i].hidden_state()
            print("MISMATCH t", t, "env", i, "in_last_mask", None if last_mask is None else bool(last_mask[i]),
                  "planes", [int((g[i][p] != cur[i][0][p]).sum()) for p in range(3)],
                  "frs", (int(f[i]), int(r[i]), int(s[i])), cur[i][1:],
                  "gpu xy", st["x"][i], st["y"][i], "turn", st["turn"][i], "ep", st["episode"][i], "nw", st["n_wolves"][i],
                  "orc", hs["x"], hs["y"], hs["turn"], hs["episode"], hs["wolves"])
            print(" diff cells", np.argwhere(g[i] != cur[i][0]).tolist())
            print(" gpu wolves", st["wolves"][i][: st["n_wolves"][i]].tolist())
            bad += 1
    if bad:
        break
    last_mask = None
    if t % 7 == 6 and done_now.any():
        mask = torch.from_numpy(done_now.astype(np.uint8)).cuda()
        last_mask = done_now.copy()
        obs = env.reset(mask)
        for i in np.nonzero(done_now)[0]:
            cur[i] = oracles[i].reset()
        done_now[:] = False
        continue
    acts = np.array([pick_action(rng, cur[i][0], env.n_actions, greedy) for i in range(n)], dtype=np.uint8)
    obs, reward, done, _ = env.step(torch.from_numpy(acts).cuda())
    reward, done = reward.cpu().numpy(), done.cpu().numpy()
    for i, o in enumerate(oracles):
        cur[i], rr, d = o.step(int(acts[i]))
        done_now[i] = d
print("done bad", bad)
