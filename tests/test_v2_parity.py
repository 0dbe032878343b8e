"""Environment 2.0 parity at size and against the reference's own material:

* the reference's two known-answer tests of ``_get_visible_objects`` ("Environment 2.0/World_tests.py":5-45 and the
  five valid rows of :49-88), restated as explicit-position worlds for the oracle (CPU) and both CUDA kernels (gpu);
* the long golden traces recorded from the unmodified reference (``oracle/make_golden_v2.py --long``: 9 worlds,
  >= 200 turns each, kills in every trace, the 19x21 / 33x64 / 64x64 shapes) replayed through the oracle, the host
  build of the kernel logic (CPU) and both CUDA kernels (gpu);
* checksum-of-everything runs at the BASELINE sizes (65,536 config-3 worlds; config 4 at 16,384 worlds against the
  oracle and at 131,072 worlds between the two independent CUDA kernels).
"""
import glob
import json
import os

import numpy as np
import pytest

from oracle import wab2_oracle
from oracle.wab2_oracle import OracleWorld2
from tests.hostsim import HostSimWorld2
from tests.test_v2 import replay, window_radius

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LONG = sorted(glob.glob(os.path.join(GOLDEN_DIR, "v2_long_*.npz")))

# ---------------------------------------------------------------------------------------------- reference KATs
# entity order of the device path is ostriches, wolves, bushes; the reference's row ORDER follows creation order and is
# not part of the planes, everything else of the expected tables is kept: (type, Delta_X, Delta_Y)
KAT_NO_WRAP = {   # World_tests.py:5-45 — World(20, 20), view radius 8, observer = the ostrich at (10, 10)
    "radius": 8, "ostriches": [(10, 10)], "wolves": [(5, 5), (15, 15)], "bushes": [(10, 5), (10, 10), (15, 10)],
    "rows": [("Wolf", -5, -5), ("Bush", 0, -5), ("Ostrich", 0, 0), ("Bush", 0, 0), ("Bush", 5, 0), ("Wolf", 5, 5)],
}
KAT_WRAP = {      # World_tests.py:49-88 — observer at (19, 10), radius 10; the second ostrich has already moved to (15, 16).
    # The file's first five expected rows are valid; its `len == 5` is stale: the moved ostrich (-4, +6) is in range too.
    "radius": 10, "ostriches": [(19, 10), (15, 16)], "wolves": [(5, 5), (15, 15)], "bushes": [(10, 10), (15, 10)],
    "rows": [("Wolf", 6, -5), ("Ostrich", 0, 0), ("Bush", -9, 0), ("Bush", -4, 0), ("Wolf", -4, 5), ("Ostrich", -4, 6)],
}
TYPE_ID = {"Ostrich": 0, "Wolf": 1, "Bush": 2}


def kat_state(kat):
    rows = []
    for t, key, food, role in ((0, "ostriches", 40, 0), (1, "wolves", 20, 0), (2, "bushes", 20, 1)):
        for (x, y) in kat[key]:
            rows.append([t, x, y, x, y, 1, food, role, 0])
    return np.asarray(rows, dtype=np.float64)


def kat_expected_planes(kat):
    R = kat["radius"]
    S = 2 * R + 1
    planes = np.zeros((3, S, S), np.uint8)
    for t, dx, dy in kat["rows"]:
        planes[TYPE_ID[t], dx + R, dy + R] = 1
    return planes


def kat_options(kat):   # lookouts (role 0) see `radius`; nobody else needs a window wider than that
    return {"lookout_view_radius": kat["radius"], "gatherer_view_radius": 5, "wolf_view_radius": 6, "starting_role": 0}


@pytest.mark.parametrize("kat", [KAT_NO_WRAP, KAT_WRAP], ids=["no_wrap", "wrap_horizontal"])
def test_reference_visibility_kats_on_the_oracle(kat):
    no, nw, nb = len(kat["ostriches"]), len(kat["wolves"]), len(kat["bushes"])
    w = OracleWorld2(20, 20, no, nw, nb, game_options=kat_options(kat), seed=1, env_id=0, window_radius=kat["radius"])
    w.set_state(kat_state(kat))
    planes, internal, rows = w.get_obs(0)
    assert rows == len(kat["rows"]) and np.array_equal(planes, kat_expected_planes(kat))
    assert list(internal) == [float(kat["ostriches"][0][0]), float(kat["ostriches"][0][1]), 40.0, 0.0, 0.0]


@pytest.mark.reference
@pytest.mark.parametrize("kat", [KAT_NO_WRAP, KAT_WRAP], ids=["no_wrap", "wrap_horizontal"])
def test_reference_visibility_kats_on_the_reference_itself(kat):
    """The restated tables are what the unmodified World._get_visible_objects returns (build container only)."""
    from oracle.ref_shim import v2 as ref_v2
    if not ref_v2.available():
        pytest.skip("reference sources not present")
    import warnings
    mods = ref_v2.load()
    world = mods["World"].World(20, 20, dict(mods["WAB_Environment2"].default_game_options))   # options injected
    ids = {"o": [world.create_ostrich(x, y) for (x, y) in kat["ostriches"]]}
    for (x, y) in kat["wolves"]:
        world.create_wolf(x, y)
    for (x, y) in kat["bushes"]:
        world.create_bush(x, y)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        df = world._get_visible_objects(ids["o"][0], kat["radius"])
    got = sorted((str(r["Type"]), int(r["Delta_X"]), int(r["Delta_Y"])) for _, r in df.iterrows())
    assert got == sorted(kat["rows"])


@pytest.mark.gpu
@pytest.mark.parametrize("grid_kernel", [False, True])
@pytest.mark.parametrize("kat", [KAT_NO_WRAP, KAT_WRAP], ids=["no_wrap", "wrap_horizontal"])
def test_reference_visibility_kats_on_the_gpu(kat, grid_kernel, monkeypatch):
    import torch
    from wab_gym_b200.world2 import VecWorld2
    R = kat["radius"]
    if grid_kernel and 2 * R + 1 > 20:
        pytest.skip("a 21x21 window does not fit the 20x20 world once: the occupancy-plane kernel does not apply")
    monkeypatch.setenv("WAB2_GRID" if grid_kernel else "WAB2_NO_GRID", "1")
    no, nw, nb = len(kat["ostriches"]), len(kat["wolves"]), len(kat["bushes"])
    n = 37                                                    # the same world in every slot of a ragged batch
    env = VecWorld2(n, 20, 20, no, nw, nb, game_options=kat_options(kat), seed=1, window_radius=R)
    assert env.lib.wab2_kernel_kind(env._h) == int(grid_kernel)
    st = np.repeat(kat_state(kat)[None], n, axis=0).astype(np.int32)
    env.import_state(st, np.ones(n, dtype=np.int32))          # turn 1: tables are fresh (no stale-position refresh)
    back, turn = env.export_state()
    assert np.array_equal(back, st) and int(turn.min()) == 1
    acts = torch.full((no + nw, n), 4, dtype=torch.uint8, device="cuda")   # 4 = role 0 for ostriches, no-op for wolves
    planes, internal, reward, done = env.turn(acts)
    want = kat_expected_planes(kat)
    got = planes[0].cpu().numpy()
    for e in range(n):
        assert np.array_equal(got[e], want), e
    assert internal[0, 0].tolist() == [kat["ostriches"][0][0], kat["ostriches"][0][1], 40, 0, 0]
    env.close()


# ---------------------------------------------------------------------------------------------- long golden traces
def load_long(path):
    z = np.load(path, allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    return meta, {k: z[k] for k in z.files if k != "meta"}


def test_long_golden_set_is_complete():
    assert len(LONG) >= 8
    shapes, actions = set(), 0
    for p in LONG:
        meta, rec = load_long(p)
        shapes.add(tuple(meta["world"][:2]))
        actions += meta["entity_actions"]
        assert meta["episodes"] * meta["turns"] >= 40 and min(meta["kills_per_episode"]) > 0, meta
        assert len(rec["reward"]) == meta["entity_actions"]
    assert {(19, 21), (33, 64), (64, 64), (20, 20)} <= shapes and actions >= 50000


@pytest.mark.parametrize("path", LONG, ids=[os.path.basename(p)[:-4] for p in LONG])
def test_oracle_reproduces_long_reference_trace(path):
    meta, rec = load_long(path)
    W, H, no, nw, nb = meta["world"]
    replay((W, H, no, nw, nb, meta["seed"], meta["env_id"]), rec, OracleWorld2, meta["episodes"], meta["turns"],
           game_options=meta["options"])


@pytest.mark.parametrize("path", LONG, ids=[os.path.basename(p)[:-4] for p in LONG])
def test_kernel_logic_reproduces_long_reference_trace(path):
    meta, rec = load_long(path)
    W, H, no, nw, nb = meta["world"]
    replay((W, H, no, nw, nb, meta["seed"], meta["env_id"]), rec, HostSimWorld2, meta["episodes"], meta["turns"],
           game_options=meta["options"])


@pytest.mark.gpu
@pytest.mark.parametrize("grid_kernel", [False, True])
@pytest.mark.parametrize("path", LONG, ids=[os.path.basename(p)[:-4] for p in LONG])
def test_cuda_reproduces_long_reference_trace(path, grid_kernel, monkeypatch):
    """The CUDA path against what the unmodified reference did: every acting entity's observation planes and internal
    observation, reward and done of every entity action, and the whole entity table after every turn."""
    import torch
    from wab_gym_b200.world2 import VecWorld2
    meta, rec = load_long(path)
    W, H, no, nw, nb = meta["world"]
    opts, n, A = meta["options"], no + nw + nb, no + nw
    R = window_radius(opts)
    S = 2 * R + 1
    if grid_kernel and (W < S or H < S):
        pytest.skip("windows do not fit this world once: thread-per-world kernel only")
    if not grid_kernel and n > 200 and meta["turns"] > 20:
        turns = 20           # the entity-scan kernel on the 328-entity world: the first 20 turns are plenty
    else:
        turns = meta["turns"]
    monkeypatch.setenv("WAB2_GRID" if grid_kernel else "WAB2_NO_GRID", "1")
    env = VecWorld2(3, W, H, no, nw, nb, game_options=opts, seed=meta["seed"], env_id_base=meta["env_id"] - 1)   # world 1 of 3
    assert env.lib.wab2_kernel_kind(env._h) == int(grid_kernel)
    st, _ = env.export_state()
    si, k = 0, 0
    assert np.array_equal(st[1].astype(np.float64), rec["state"][si]), "create"
    si += 1
    for ep in range(meta["episodes"]):
        env.reset_environment()
        st, _ = env.export_state()
        assert np.array_equal(st[1].astype(np.float64), rec["state"][si]), ("reset", ep)
        si += 1
        for turn in range(meta["turns"]):
            if turn >= turns:
                break
            acts = rec["actions"][ep * meta["turns"] + turn]
            a = torch.from_numpy(np.repeat(np.asarray(acts[:A], dtype=np.uint8)[:, None], 3, axis=1)).cuda()
            planes, internal, reward, done = env.turn(a)
            planes, internal, reward, done = planes[:, 1].cpu().numpy(), internal[:, 1].cpu().numpy(), reward[:, 1].cpu().numpy(), done[:, 1].cpu().numpy()
            for i in range(A):
                want = np.unpackbits(rec["planes"][k + i])[: 3 * S * S].reshape(3, S, S)
                assert np.array_equal(planes[i], want), (ep, turn, i)
                assert [float(v) for v in internal[i]] == list(rec["internal"][k + i]), (ep, turn, i)
                assert float(reward[i]) == rec["reward"][k + i] and int(done[i]) == rec["done"][k + i], (ep, turn, i)
            k += n
            st, _ = env.export_state()
            assert np.array_equal(st[1].astype(np.float64), rec["state"][si]), ("turn", ep, turn)
            si += 1
        if turns < meta["turns"]:
            break
    env.close()


# ---------------------------------------------------------------------------------------------- checksums at size
def gpu_checksums(env, actions, episodes, turns):
    """The two sums of oracle/wab2_oracle.c:wab2_oracle_run, computed from the tensors wab2_turn writes."""
    import torch
    A, n = env.n_acting, env.num_envs
    K = 3 * (2 * env.R + 1) ** 2
    w = torch.arange(1, K + 1, device="cuda", dtype=torch.float32)
    ent_w = torch.arange(1, A + 1, device="cuda", dtype=torch.int64)
    iw = torch.tensor([7, 11, 13, 17, 19], device="cuda", dtype=torch.int64)
    total = 0
    for ep in range(episodes):
        env.reset_environment()
        for t in range(turns):
            planes, internal, reward, done = env.turn(actions[ep * turns + t])
            per = torch.empty((A, n), dtype=torch.int64, device="cuda")
            for a in range(A):    # float32 dot products are exact here: at most K (K + 1) / 2 < 2^24
                per[a] = (planes[a].reshape(n, K).to(torch.float32) @ w).to(torch.int64)
            per += (internal.to(torch.int64) * iw).sum(-1) + 23 * reward.to(torch.int64) + 29 * done.to(torch.int64)
            total += int((per.sum(1) * ent_w).sum().item())
    st, _ = env.export_state()
    k = np.arange(1, st.shape[1] + 1, dtype=np.int64)[None, :]
    s = st.astype(np.int64)
    state = int((k * (s[:, :, 1] + 3 * s[:, :, 2] + 5 * s[:, :, 3] + 7 * s[:, :, 4] + 11 * s[:, :, 5] + 13 * s[:, :, 6]
                      + 17 * s[:, :, 7] + 19 * s[:, :, 8])).sum())
    return total, state


def v2_actions(no, nw, n, steps, seed):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.empty((steps, no + nw, n), dtype=torch.uint8, device="cuda")
    a[:, :no] = torch.randint(0, 6, (steps, no, n), dtype=torch.uint8, device="cuda", generator=g)
    a[:, no:] = torch.randint(0, 5, (steps, nw, n), dtype=torch.uint8, device="cuda", generator=g)
    return a


_oracle_cache = {}


@pytest.mark.gpu
@pytest.mark.parametrize("grid_kernel", [False, True])
def test_config3_full_size_checksum_against_oracle(grid_kernel, monkeypatch):
    """BASELINE configs[2] at its size: 65,536 worlds x 2 episodes x 50 turns = 6.5 M world turns (85 M observations),
    every observation byte, internal value, reward and done, and the final entity tables, against the CPU oracle."""
    from wab_gym_b200.world2 import VecWorld2
    monkeypatch.setenv("WAB2_GRID" if grid_kernel else "WAB2_NO_GRID", "1")
    dims, n, episodes, turns, seed = (20, 20, 10, 3, 20), 65536, 2, 50, 5
    acts = v2_actions(dims[2], dims[3], n, episodes * turns, 77)
    env = VecWorld2(n, *dims, seed=seed)
    assert env.lib.wab2_kernel_kind(env._h) == int(grid_kernel)
    got = gpu_checksums(env, acts, episodes, turns)
    env.close()
    if "c3" not in _oracle_cache:
        _oracle_cache["c3"] = wab2_oracle.run(*dims, None, seed, 0, n, episodes, turns, acts.cpu().numpy())
    done, want_turns, want_state = _oracle_cache["c3"]
    assert done == n * episodes * turns and got == (want_turns, want_state)


@pytest.mark.gpu
def test_config4_checksums_oracle_and_both_kernels(monkeypatch):
    """BASELINE configs[3] (64x64, 8 + 64 + 256 entities): 16,384 worlds x 40 turns against the CPU oracle (the naive
    oracle does ~12 k such turns/s per core), then the full per-GPU share — 131,072 worlds x 2 episodes x 50 turns,
    944 M observations — on the occupancy-plane kernel, whose first 16,384 worlds x first 40 turns equal the oracle's
    run, against the entity-scan kernel on a 32,768-world shard of the same ids (two independent CUDA algorithms)."""
    import torch
    from wab_gym_b200.world2 import VecWorld2
    dims, seed = (64, 64, 8, 64, 256), 9
    no, nw = dims[2], dims[3]
    n_big, episodes, turns = 131072, 2, 50
    acts = v2_actions(no, nw, n_big, episodes * turns, 78)
    # (a) oracle-sized prefix: worlds 0..16383, episode 1 only, 40 turns
    n_small, t_small = 16384, 40
    a_small = acts[:t_small, :, :n_small].contiguous()
    monkeypatch.setenv("WAB2_GRID", "1")
    env = VecWorld2(n_small, *dims, seed=seed)
    assert env.lib.wab2_kernel_kind(env._h) == 1
    got_small = gpu_checksums(env, a_small, 1, t_small)
    env.close()
    done, want_turns, want_state = wab2_oracle.run(*dims, None, seed, 0, n_small, 1, t_small, a_small.cpu().numpy())
    assert done == n_small * t_small and got_small == (want_turns, want_state)
    # (b) full size on the occupancy-plane kernel vs a shard on the entity-scan kernel
    env = VecWorld2(n_big, *dims, seed=seed)
    shard0, shard_n = 65536, 32768
    monkeypatch.delenv("WAB2_GRID")
    monkeypatch.setenv("WAB2_NO_GRID", "1")
    scan = VecWorld2(shard_n, *dims, seed=seed, env_id_base=shard0)
    assert scan.lib.wab2_kernel_kind(scan._h) == 0
    K = 3 * (2 * env.R + 1) ** 2
    w = torch.arange(1, K + 1, device="cuda", dtype=torch.float32)
    for ep in range(episodes):
        env.reset_environment(); scan.reset_environment()
        for t in range(turns):
            a = acts[ep * turns + t]
            p1, i1, r1, d1 = env.turn(a)
            if t % 10 == 9 or t < 3:       # the scan kernel is 30x slower: compare a sample of turns in full ...
                p2, i2, r2, d2 = scan.turn(a[:, shard0:shard0 + shard_n].contiguous())
                sl = slice(shard0, shard0 + shard_n)
                assert torch.equal(p1[:, sl], p2) and torch.equal(i1[:, sl], i2) and torch.equal(r1[:, sl], r2) and torch.equal(d1[:, sl], d2), (ep, t)
            else:                          # ... and keep the shard in step on the other turns
                scan.turn(a[:, shard0:shard0 + shard_n].contiguous())
            assert int(p1.max()) == 1
        s1, t1 = env.export_state()
        s2, t2 = scan.export_state()
        assert np.array_equal(s1[shard0:shard0 + shard_n], s2) and np.array_equal(t1[shard0:shard0 + shard_n], t2), ep
    env.close(); scan.close()
