"""The reference's egocentric observation family (SURVEY §8f ④; wab_env.py:637-667, :930-958): the oracle (CPU) and
the CUDA kernel (gpu) against proximities recorded from the unmodified reference (``tests/golden/ego.npz``,
``python -m oracle.make_golden_wrappers ego``), and the CUDA kernel against the oracle on a larger batch."""
import json
import os

import numpy as np
import pytest

from oracle.wab_oracle import OracleEnv

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ego.npz")


def cases():
    z = np.load(GOLDEN)
    return [(m, z[m["key"] + "_actions"], z[m["key"] + "_prox"]) for m in json.loads(str(z["meta"]))]


def test_oracle_reproduces_reference_egocentric_proximities():
    seen_far = 0
    for m, actions, prox in cases():
        orc = OracleEnv(m["options"], seed=m["seed"], env_id=m["env_id"])
        orc.reset()
        k = 0
        w, b = orc.ego_proximities()
        assert w + b == prox[k].tolist(), (m["key"], "reset")
        for a in actions:
            if a < 0:
                orc.reset()
            else:
                orc.step(int(a))
            k += 1
            w, b = orc.ego_proximities()
            assert w + b == prox[k].tolist(), (m["key"], k)
        assert k + 1 == len(prox)
        seen_far += int((prox[:, 5:] <= 5).sum())          # nearest bush 6+ squares away: outside the 11x11 window
    assert seen_far > 50


@pytest.mark.gpu
def test_cuda_reproduces_reference_egocentric_proximities():
    import torch
    from wab_gym_b200 import VecEnv
    for m, actions, prox in cases():
        env = VecEnv(3, m["options"], seed=m["seed"], env_id_base=m["env_id"] - 1, auto_reset=False, ego=True, wolf_cap=64)
        env.reset()
        k = 0
        assert env.ego_proximities()[1].tolist() == prox[k].tolist(), (m["key"], "reset")
        mask = torch.tensor([0, 1, 0], dtype=torch.uint8, device="cuda")
        for a in actions:
            if a < 0:
                env.reset(mask)
            else:
                env.step(torch.tensor([4, int(a), 4], dtype=torch.uint8, device="cuda"))
            k += 1
            assert env.ego_proximities()[1].tolist() == prox[k].tolist(), (m["key"], k)
        env.close()


@pytest.mark.gpu
@pytest.mark.parametrize("lpe", [1, 8])
def test_cuda_egocentric_proximities_match_oracle_on_a_batch(lpe, monkeypatch):
    import torch
    from tests.util import OPTION_SETS
    from wab_gym_b200 import VecEnv
    monkeypatch.setenv("WAB_LPE", str(lpe))
    opts = OPTION_SETS["dense"][0]
    n, steps, seed = 70, 220, 23
    env = VecEnv(n, opts, seed=seed, ego=True, wolf_cap=64)
    oracles = [OracleEnv(opts, seed=seed, env_id=i) for i in range(n)]
    env.reset()
    for o in oracles:
        o.reset()
    rng = np.random.default_rng(3)
    for t in range(steps):
        got = env.ego_proximities().cpu().numpy()
        for i, o in enumerate(oracles):
            w, b = o.ego_proximities()
            assert got[i].tolist() == w + b, (t, i)
        acts = rng.integers(0, env.n_actions, n).astype(np.uint8)
        _, _, done, _ = env.step(torch.from_numpy(acts).cuda())
        done = done.cpu().numpy()
        for i, o in enumerate(oracles):
            _, _, d = o.step(int(acts[i]))
            assert d == bool(done[i])
            if d:
                o.reset()
    env.close()
