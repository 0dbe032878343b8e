"""The reference's consumer-side code frozen into fixtures (``oracle/make_golden_wrappers.py``) against this repo:

* ``features.npz`` — the LIVE ``PragmaticObsWrapper.observation`` (wab_env.py:726-761) on 2,400 random observation
  tuples + the three inputs of the reference's own known-answer tests (wab_env_test.py:9-169): checked against the
  host build of ``wab_features.cuh`` (CPU) and — independently of that build — against the CUDA kernels on the GPU,
  standalone (``wab_pragmatic_features``) and through the 449-column one-hot (``wab_vec_flatten_features``).
* ``render.npz`` — ``WolvesAndBushesEnv.render`` frames (wab_env.py:468-502) of keyed reference episodes, reproduced by
  the gym-compat class on the CUDA path after replaying the recorded action trace.
"""
import json
import os

import numpy as np
import pytest

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_features():
    z = np.load(os.path.join(GOLDEN_DIR, "features.npz"))
    n = z["features"].shape[0]
    wolves = np.unpackbits(z["wolves"], axis=1)[:, :121].reshape(n, 11, 11)
    bushes = np.unpackbits(z["bushes"], axis=1)[:, :121].reshape(n, 11, 11)
    return wolves, bushes, z["scalars"], z["features"], int(z["n_kats"])


def test_golden_feature_set_shape():
    wolves, bushes, scalars, feats, n_kats = load_features()
    assert feats.shape == (2403, 28) and n_kats == 3 and wolves.shape == (2403, 11, 11)
    assert feats[1, 24] == 1 and feats[:, 24].sum() > 200            # standing_on_bush of test_standing_on_bush, and many more
    assert (feats[:, :24].max(axis=0) > 0).all()                     # every feature column is exercised


def test_host_build_of_feature_header_matches_the_live_wrapper_outputs():
    from tests import hostsim
    wolves, bushes, scalars, feats, _ = load_features()
    for i in range(feats.shape[0]):
        got = hostsim.features(wolves[i].astype(float), bushes[i].astype(float), int(scalars[i, 0]), int(scalars[i, 1]), int(scalars[i, 2]))
        assert [int(v) for v in got] == [int(v) for v in feats[i]], i


@pytest.mark.gpu
def test_cuda_feature_kernels_match_the_live_wrapper_outputs():
    import torch
    from wab_gym_b200 import VecEnv
    from wab_gym_b200.vec_env import ObsBatch
    wolves, bushes, scalars, feats, _ = load_features()
    n = feats.shape[0]
    grids = np.zeros((n, 3, 11, 11), dtype=np.uint8)
    grids[:, 0], grids[:, 1], grids[:, 2, 5, 5] = wolves, bushes, 1
    env = VecEnv(8, seed=0)
    obs = ObsBatch(torch.from_numpy(grids).cuda(), *(torch.from_numpy(np.ascontiguousarray(scalars[:, k])).cuda() for k in range(3)))
    got = env.pragmatic_features(obs)
    assert np.array_equal(got.cpu().numpy(), feats)
    # gym.spaces.flatten of the wrapper observation (actor_critic.py:188): one-hot of every Discrete in tuple order
    flat = env.flatten_features(got).cpu().numpy()
    assert flat.shape == (n, 449)
    dims = [12] * 8 + [11] * 4 + [12] * 8 + [11] * 4 + [2, 41, 2, 3]
    want = np.zeros((n, 449), dtype=np.float32)
    col = 0
    for k, d in enumerate(dims):
        want[np.arange(n), col + feats[:, k].astype(np.int64)] = 1
        col += d
    assert col == 328 and np.array_equal(flat, want)                 # the last 121 columns (view mask) are zero without restrict_view
    env.close()


def _render_cases():
    z = np.load(os.path.join(GOLDEN_DIR, "render.npz"))
    return json.loads(str(z["meta"])), z


def test_golden_render_set_has_every_kind_of_frame():
    meta, z = _render_cases()
    kinds = {k for m in meta for k in m["kinds"]}
    assert {"reset", "alive0", "alive1", "killed", "starved", "killed+health"} <= kinds and all(not m["missing"] for m in meta)
    assert z["restrict_view_frames"].shape[1:] == (352, 352, 3)


@pytest.mark.gpu
def test_render_frames_equal_the_reference_frames():
    from wab_gym_b200.config import default_game_options
    from wab_gym_b200.env import WolvesAndBushesEnv
    meta, z = _render_cases()
    for m in meta:
        env = WolvesAndBushesEnv({**default_game_options, **m["options"]}, seed=m["seed"], env_id=m["env_id"])
        actions, frames, steps = z[m["name"] + "_actions"], z[m["name"] + "_frames"], z[m["name"] + "_frame_steps"]
        k = 0
        assert np.array_equal(env.render(draw_health=False), frames[k]), (m["name"], "reset")
        k += 1
        for t, a in enumerate(actions, start=1):
            if a < 0:
                env.reset()
            else:
                env.step(int(a))
            while k < len(steps) and steps[k] == t:
                health = m["kinds"][k].endswith("+health")
                assert np.array_equal(env.render(draw_health=health), frames[k]), (m["name"], m["kinds"][k], t)
                k += 1
        assert k == len(steps)
        env.close()


@pytest.mark.gpu
def test_episode_monitor_on_device_matches_env_statistics(tmp_path):
    import torch
    from wab_gym_b200 import VecEnv
    from wab_gym_b200.monitor import EpisodeMonitor
    n, steps = 512, 200
    env = VecEnv(n, seed=4)
    mon = EpisodeMonitor(n, directory=str(tmp_path), device="cuda", flush_every=64)
    env.reset()
    gen = torch.Generator(device="cuda").manual_seed(2)
    total = 0.0
    for t in range(steps):
        _, reward, done, _ = env.step(torch.randint(0, 5, (n,), dtype=torch.uint8, device="cuda", generator=gen))
        mon.record(reward, done)
        total += float(reward.double().sum())
    st = mon.stats()
    s = env.stats()
    assert len(st["episode_lengths"]) == s["episodes"] > 0
    open_steps = int(mon._len.sum())
    assert sum(st["episode_lengths"]) + open_steps == n * steps
    assert abs(sum(st["episode_rewards"]) + float(mon._ret.sum()) - total) < 1e-6 * n * steps
    assert max(st["episode_lengths"]) <= 80 and mon.close() is not None
    env.close()
