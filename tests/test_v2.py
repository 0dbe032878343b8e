"""Environment 2.0: CPU oracle vs the reference (live, build container only) and vs the committed golden trace;
the kernel logic header (host build) vs the oracle; the CUDA path vs the oracle (gpu)."""
import os
import random
import warnings

import numpy as np
import pytest

from oracle.ref_shim import v2 as ref_v2
from oracle.wab2_oracle import OracleWorld2
from tests.hostsim import HostSimWorld2

TYPES = {"Ostrich": 0, "Wolf": 1, "Bush": 2}
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "v2_trace.npz")
WORLDS = [  # (W, H, ostriches, wolves, bushes, seed, env_id)
    (20, 20, 10, 3, 20, 1, 0),     # Env2Tests.py:7-22 (BASELINE config 3 population)
    (7, 9, 6, 4, 5, 2, 3),         # small and crowded: kills, wrap-around and stale positions every few turns
    (12, 5, 3, 6, 2, 9, 7),
]


def actions_for(rng, no, nw, n):
    return [rng.randint(0, 5) if i < no else (rng.randint(0, 4) if i < no + nw else 0) for i in range(n)]


def window_radius(game_options=None):
    o = {"lookout_view_radius": 9, "gatherer_view_radius": 5, "wolf_view_radius": 6}
    o.update({k: v for k, v in (game_options or {}).items() if k in o})
    return max(o.values())


def record_reference(world, episodes, turns, game_options=None):
    """Run the real reference and record everything the other implementations are compared with."""
    W, H, no, nw, nb, seed, env_id = world
    ref = ref_v2.make_env(W, H, no, nw, nb, game_options=game_options, seed=seed, env_id=env_id)
    n, R = no + nw + nb, window_radius(game_options)
    S = 2 * R + 1
    rng = random.Random(seed)
    rec = {"actions": [], "planes": [], "rows": [], "internal": [], "reward": [], "done": [], "state": []}

    def state():
        return np.array([[TYPES[t], x, y, tx, ty, int(v), f, r, s] for (t, x, y, tx, ty, v, f, r, s) in ref_v2.hidden_state(ref)], dtype=np.float64)

    rec["state"].append(state())
    for ep in range(episodes):
        ref.reset_environment()
        rec["state"].append(state())
        for turn in range(turns):
            acts = actions_for(rng, no, nw, n)
            for i in range(n):
                df, internal = ref.get_obs(i)
                planes = np.zeros((3, S, S), np.uint8)
                for _, row in df.iterrows():
                    planes[TYPES[row["Type"]], int(row["Delta_X"]) + R, int(row["Delta_Y"]) + R] = 1
                rr, rd = ref.take_action(i, acts[i])
                rec["planes"].append(np.packbits(planes)); rec["rows"].append(len(df))
                rec["internal"].append([float(v) for v in internal] + [0.0] * (5 - len(internal)))
                rec["reward"].append(float(rr)); rec["done"].append(int(bool(rd)))
            rec["actions"].append(acts)
            rec["state"].append(state())
    return {k: np.asarray(v) for k, v in rec.items()}


def replay(world, rec, make, episodes, turns, exact_float=True, game_options=None):
    """Drive an implementation with the recorded actions and compare everything."""
    W, H, no, nw, nb, seed, env_id = world
    impl = make(W, H, no, nw, nb, game_options=game_options, seed=seed, env_id=env_id)
    n, R = no + nw + nb, window_radius(game_options)
    S = 2 * R + 1

    def state():
        st = impl.state()
        return np.asarray(st[0] if isinstance(st, tuple) else st, dtype=np.float64)

    si, k = 0, 0
    assert np.array_equal(state(), rec["state"][si]), "create"
    si += 1
    for ep in range(episodes):
        impl.reset_environment()
        assert np.array_equal(state(), rec["state"][si]), ("reset", ep)
        si += 1
        for turn in range(turns):
            acts = rec["actions"][ep * turns + turn]
            for i in range(n):
                planes, internal, rows = impl.get_obs(i)
                want = np.unpackbits(rec["planes"][k])[: 3 * S * S].reshape(3, S, S)
                assert rows == rec["rows"][k] and np.array_equal(planes, want), (ep, turn, i)
                assert [float(v) for v in internal] == list(rec["internal"][k]), (ep, turn, i)
                r, d = impl.take_action(i, int(acts[i]))
                assert float(r) == rec["reward"][k] and int(d) == rec["done"][k], (ep, turn, i)
                k += 1
            assert np.array_equal(state(), rec["state"][si]), ("turn", ep, turn)
            si += 1


@pytest.mark.reference
@pytest.mark.skipif(not ref_v2.available(), reason="reference sources not present")
@pytest.mark.parametrize("world", WORLDS[:2])
def test_oracle_and_kernel_logic_match_reference_live(world):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        rec = record_reference(world, episodes=2, turns=6)
    replay(world, rec, OracleWorld2, 2, 6)
    replay(world, rec, HostSimWorld2, 2, 6)


def _load_golden():
    z = np.load(GOLDEN, allow_pickle=False)
    out = []
    for w in range(int(z["n_worlds"])):
        rec = {k[len("w%d_" % w):]: z[k] for k in z.files if k.startswith("w%d_" % w)}
        out.append((tuple(int(v) for v in z["worlds"][w]), rec, int(z["episodes"]), int(z["turns"])))
    return out


def test_oracle_reproduces_golden_reference_trace():
    for world, rec, episodes, turns in _load_golden():
        replay(world, rec, OracleWorld2, episodes, turns)


def test_kernel_logic_reproduces_golden_reference_trace():
    for world, rec, episodes, turns in _load_golden():
        replay(world, rec, HostSimWorld2, episodes, turns)


@pytest.mark.parametrize("world", WORLDS)
def test_kernel_logic_matches_oracle_long(world):
    W, H, no, nw, nb, seed, env_id = world
    orc = OracleWorld2(W, H, no, nw, nb, seed=seed, env_id=env_id)
    sim = HostSimWorld2(W, H, no, nw, nb, seed=seed, env_id=env_id)
    n = no + nw + nb
    rng = random.Random(5)
    kills = 0
    for ep in range(4):
        orc.reset_environment(); sim.reset_environment()
        for turn in range(40):
            acts = actions_for(rng, no, nw, n)
            for i in range(n):
                po, io, ro = orc.get_obs(i)
                ps, is_, rs = sim.get_obs(i)
                assert ro == rs and np.array_equal(po, ps) and [float(v) for v in is_] == list(io), (ep, turn, i)
                assert orc.take_action(i, acts[i]) == sim.take_action(i, acts[i]), (ep, turn, i)
            so, (ss, st) = orc.state(), sim.state()
            assert np.array_equal(so, ss.astype(np.float64)) and st == orc.turn, (ep, turn)
        kills += int((so[:no, 8] == 2).sum())
    assert kills > 0 or W * H > 200


GPU_WORLDS = WORLDS + [
    (19, 21, 6, 6, 10, 4, 11),       # smallest world the occupancy-plane kernel accepts: every edge case of the wrap rule
    (64, 64, 8, 64, 256, 6, 0),      # BASELINE config 4 population
    (33, 64, 12, 40, 30, 8, 5),
]


@pytest.mark.gpu
@pytest.mark.parametrize("force_thread_kernel", [False, True])
@pytest.mark.parametrize("world", GPU_WORLDS)
def test_vecworld2_matches_oracle(world, force_thread_kernel, monkeypatch):
    import torch
    from wab_gym_b200.world2 import VecWorld2
    W, H, no, nw, nb, seed, base = world
    if force_thread_kernel:
        if no + nw + nb > 200:
            pytest.skip("thread-per-world kernel is too slow to be interesting here")
        monkeypatch.setenv("WAB2_NO_GRID", "1")
    else:
        monkeypatch.setenv("WAB2_GRID", "1")       # occupancy-plane kernel wherever the world allows it
    n_envs, n = (70 if no + nw + nb < 200 else 9), no + nw + nb
    env = VecWorld2(n_envs, W, H, no, nw, nb, seed=seed, env_id_base=base)
    oracles = [OracleWorld2(W, H, no, nw, nb, seed=seed, env_id=base + e) for e in range(n_envs)]
    rng = random.Random(3)
    st, _ = env.export_state()
    for e, o in enumerate(oracles):
        assert np.array_equal(st[e].astype(np.float64), o.state()), ("create", e)
    for ep in range(2):
        env.reset_environment()
        for o in oracles:
            o.reset_environment()
        for turn in range(12):
            acts = np.array([actions_for(rng, no, nw, n)[: no + nw] for _ in range(n_envs)], dtype=np.uint8)
            planes, internal, reward, done = env.turn(torch.from_numpy(acts.T.copy()).cuda())
            planes, internal, reward, done = planes.cpu().numpy(), internal.cpu().numpy(), reward.cpu().numpy(), done.cpu().numpy()
            for e, o in enumerate(oracles):
                for i in range(n):
                    a = int(acts[e][i]) if i < no + nw else 0
                    if i < no + nw:
                        po, io, _ = o.get_obs(i)
                        assert np.array_equal(planes[i, e], po), (ep, turn, e, i)
                        assert [float(v) for v in internal[i, e]] == list(io), (ep, turn, e, i)
                    r, d = o.take_action(i, a)
                    if i < no + nw:
                        assert reward[i, e] == r and bool(done[i, e]) == d, (ep, turn, e, i)
            st, tn = env.export_state()
            for e, o in enumerate(oracles):
                assert np.array_equal(st[e].astype(np.float64), o.state()) and tn[e] == o.turn, (ep, turn, e)
    env.close()


@pytest.mark.gpu
def test_vecworld2_outputs_never_touch_guard_bytes():
    import torch
    from wab_gym_b200 import _lib
    from wab_gym_b200.world2 import VecWorld2
    import ctypes
    n, G = 45, 256
    env = VecWorld2(n, 7, 9, 6, 4, 5, seed=2)
    env.reset_environment()
    raw = torch.full((env.planes_store.numel() + 2 * G,), 0xAB, dtype=torch.uint8, device="cuda")
    # 16-byte aligned interior: G is a multiple of 16
    inner = raw[G:G + env.planes_store.numel()].view(env.planes_store.shape)
    env.planes_store = inner
    acts = torch.randint(0, 5, (env.n_acting, n), dtype=torch.uint8, device="cuda")
    for _ in range(4):
        env.turn(acts)
    torch.cuda.synchronize()
    assert int((raw[:G] != 0xAB).sum()) == 0 and int((raw[-G:] != 0xAB).sum()) == 0
    assert int(inner.max()) <= 1
    env.close()
