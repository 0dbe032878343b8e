"""Viewports other than 11 x 11 and spawn margins other than 1 (the reference is generic in both: wab_env.py:25-26, :34,
:147-148, :510-576) on the warp-per-env kernels of wab_generic.cuh: against traces recorded from the unmodified reference
(tests/golden/sized_*.npz), against the oracle on batches (observations, rewards, dones, wolf lists, depletion logs),
and — forced onto the default geometry — against the specialised 11 x 11 kernels."""
import numpy as np
import pytest
import torch

from oracle import wab_oracle
from oracle.wab_oracle import OracleEnv
from tests.util import golden_wolves, load_golden, sized_names

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sized_names())
def test_compat_env_reproduces_sized_reference_trace(name):
    from wab_gym_b200.config import default_game_options
    from wab_gym_b200.env import WolvesAndBushesEnv
    meta, tr = load_golden(name, sized=True)
    env = WolvesAndBushesEnv({**default_game_options, **meta["overrides"]}, seed=meta["seed"], env_id=meta["env_id"])
    assert env.action_space.n == meta["n_actions"]
    obs = env._get_obs()
    for t in range(len(tr["action"])):
        a = int(tr["action"][t])
        if a < 0:
            if t > 0:
                obs = env.reset()
            reward, done = 0.0, False
        else:
            obs, reward, done, _ = env.step(a)
        g = np.stack([np.asarray(obs[p]) for p in range(3)]).astype(np.uint8)
        assert g.shape == tr["grids"][t].shape and np.array_equal(g, tr["grids"][t]), (name, t)
        assert (obs[3], obs[4], obs[5]) == (tr["food"][t], tr["role"][t], tr["status"][t]), (name, t)
        assert reward == tr["reward"][t] and done == bool(tr["done"][t]), (name, t)
    env.close()


SIZES = [  # (width, height, margin, extra options)
    (5, 5, 1, {"chance_wolf_on_square": 0.02, "bush_power": 12}),
    (7, 9, 2, {"chance_wolf_on_square": 0.006, "lookout_only": False, "starting_role": None, "starting_food": None}),
    (11, 11, 2, {"chance_wolf_on_square": 0.004, "restrict_view": True, "lookout_only": False, "starting_role": None}),
    (13, 3, 1, {"chance_wolf_on_square": 0.01, "max_berries_per_bush": 2, "bush_power": 8, "turns_to_fill_food": 2}),
    (21, 17, 2, {"chance_wolf_on_square": 0.003, "bush_power": 30, "max_berries_per_bush": 3, "wolf_chance_to_despawn": 0.2}),
    (31, 31, 2, {"chance_wolf_on_square": 0.002, "max_turns": 40}),
    (7, 7, 2, {"chance_wolf_on_square": 0.05, "god_mode": True, "max_turns": 200, "turns_to_empty_food": 250}),   # > 32 wolves
]


@pytest.mark.parametrize("case", range(len(SIZES)))
@pytest.mark.parametrize("f64", [False, True])
def test_generic_kernels_match_oracle(case, f64):
    from wab_gym_b200 import VecEnv
    W, H, m, extra = SIZES[case]
    opts = {"width": W, "height": H, "wolf_spawn_margin": m, **extra}
    n, steps, seed, base = 37, (150 if W * H < 500 else 90) if case != len(SIZES) - 1 else 190, 50 + case, 700
    env = VecEnv(n, opts, seed=seed, env_id_base=base, wolf_cap=64, force_f64_food=f64)
    assert env.generic_kernels and env.lanes_per_env == 32 and env.view == (W, H)
    oracles = [OracleEnv(opts, seed=seed, env_id=base + i) for i in range(n)]
    rng = np.random.default_rng(case)
    obs = env.reset()
    cur = [o.reset() for o in oracles]
    ci, cj = W // 2, H // 2
    peak = 0
    for t in range(steps + 1):
        g, f, r, s = (x.cpu().numpy() for x in obs)
        st = env.export_state() if t % 10 == 0 else None
        for i, o in enumerate(oracles):
            assert g[i].shape == cur[i][0].shape and np.array_equal(g[i], cur[i][0]), (case, t, i, np.argwhere(g[i] != cur[i][0])[:4])
            assert (int(f[i]), int(r[i]), int(s[i])) == cur[i][1:], (case, t, i)
            if st is not None:
                hs = o.hidden_state()
                nw = int(st["n_wolves"][i])
                peak = max(peak, nw)
                assert (st["x"][i], st["y"][i], st["turn"][i]) == (hs["x"], hs["y"], hs["turn"]), (case, t, i)
                assert sorted((int(a), int(b)) for a, b in st["wolves"][i][:nw]) == hs["wolves"], (case, t, i)
        if t == steps:
            break
        acts = np.array([4 if (cur[i][0][1][ci, cj] == 1 and rng.random() < 0.7) else rng.integers(0, env.n_actions)
                         for i in range(n)], dtype=np.uint8)
        obs, reward, done, _ = env.step(torch.from_numpy(acts).cuda())
        reward, done = reward.cpu().numpy(), done.cpu().numpy()
        for i, o in enumerate(oracles):
            c, rr, d = o.step(int(acts[i]))
            assert np.float32(rr) == reward[i] and d == bool(done[i]), (case, t, i)
            cur[i] = o.reset() if d else c
    stt = env.stats()
    assert stt["steps"] == n * steps and stt["overflows"] == 0 and stt["bad_actions"] == 0
    if case == len(SIZES) - 1:
        assert peak > 32, peak
    env.close()


def test_generic_kernels_on_the_default_geometry_equal_the_specialised_ones(monkeypatch):
    """WAB_GENERIC=1 forces the warp-per-env kernels onto 11 x 11 / margin 1: every output byte of 300 envs x 200 steps
    (step_many), features included, equals the specialised kernels', and so does the statistics vector."""
    from wab_gym_b200 import VecEnv
    opts = {"lookout_only": False, "restrict_view": True, "starting_role": None, "chance_wolf_on_square": 0.004}
    n, steps = 300, 200
    acts = torch.randint(0, 6, (steps, n), dtype=torch.uint8, device="cuda", generator=torch.Generator("cuda").manual_seed(3))
    outs = []
    for generic in (False, True):
        if generic:
            monkeypatch.setenv("WAB_GENERIC", "1")
        env = VecEnv(n, opts, seed=8, features=True, wolf_cap=64)
        assert env.generic_kernels == generic
        o0 = env.reset()
        first = [x.clone() for x in o0] + [env.last_features.clone()]
        o, r, d, info = env.step_many(acts)
        outs.append(first + [x.clone() for x in o] + [r.clone(), d.clone(), info["info"].clone(), info["features"].clone()])
        outs[-1].append(env.stats())
        env.close()
    for a, b in zip(outs[0][:-1], outs[1][:-1]):
        assert torch.equal(a, b)
    assert outs[0][-1] == outs[1][-1]


def test_sized_batch_checksum_against_oracle():
    """15 x 13 viewport, margin 2: 8,192 envs x 64 steps through step_many, checksum of every observation byte."""
    from wab_gym_b200 import VecEnv
    opts = {"width": 15, "height": 13, "wolf_spawn_margin": 2, "chance_wolf_on_square": 0.002}
    n, steps, seed = 8192, 64, 4
    acts = torch.randint(0, 5, (steps, n), dtype=torch.uint8, device="cuda", generator=torch.Generator("cuda").manual_seed(9))
    env = VecEnv(n, opts, seed=seed, wolf_cap=64)
    env.reset()
    o, r, d, _ = env.step_many(acts)
    K = 3 * 15 * 13
    w = torch.arange(1, K + 1, device="cuda", dtype=torch.int64)
    cs = (o.grids.view(steps, n, K).long() * w).sum() + 1000 * o.food.long().sum() + 100000 * o.role.long().sum() \
        + 200000 * o.status.long().sum() + 400000 * d.long().sum()
    n_steps, want = wab_oracle.run(opts, seed, n, steps, acts.cpu().numpy())
    assert n_steps == n * steps and int(cs.item()) == want
    assert env.stats()["overflows"] == 0
    env.close()


def test_unsupported_geometries_fail_loudly():
    from wab_gym_b200 import VecEnv
    with pytest.raises(ValueError):
        VecEnv(4, {"width": 10})                                  # even: ValueError like wab_env.py:147-148
    with pytest.raises(NotImplementedError):
        VecEnv(4, {"width": 33})
    with pytest.raises(NotImplementedError):
        VecEnv(4, {"width": 7, "restrict_view": True})            # the reference's tile masks are 11 x 11 literals
    with pytest.raises(NotImplementedError):
        VecEnv(4, {"wolf_spawn_margin": 3})
