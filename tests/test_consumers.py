"""Consumer-side rows of the path (SURVEY §8f ②, ③) on CPU: the batched ``finish_episode`` loss against a literal
restatement of the reference's per-episode loop (actor_critic.py:128-165), and the Monitor-style stats writer."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from wab_gym_b200.a2c import episodic_actor_critic_loss
from wab_gym_b200.monitor import EpisodeMonitor


def reference_finish_episode_loss(log_probs, values, rewards, gamma, eps):
    """actor_critic.py:128-165 for ONE episode, statement by statement (lists in, loss tensor out)."""
    R = 0
    policy_losses, value_losses, returns = [], [], []
    for r in rewards[::-1]:
        R = r + gamma * R
        returns.insert(0, R)
    returns = torch.tensor(returns)
    returns = (returns - returns.mean()) / (returns.std() + eps)
    for (log_prob, value), R in zip(zip(log_probs, values), returns):
        advantage = R - value.item()
        policy_losses.append(-log_prob * advantage)
        value_losses.append(F.smooth_l1_loss(value, torch.tensor([R])))
    return torch.stack(policy_losses).sum() + torch.stack(value_losses).sum()


def test_batched_finish_episode_equals_the_reference_loop():
    torch.manual_seed(3)
    T, N, gamma = 80, 37, 0.99
    eps = np.finfo(np.float32).eps.item()
    lengths = torch.randint(2, T + 1, (N,))
    lengths[0], lengths[1] = T, 2
    log_probs = (-torch.rand(T, N) * 2).requires_grad_()
    values = torch.randn(T, N, requires_grad=True)
    rewards = torch.tensor(np.random.default_rng(0).choice([0.0, 0.1, 1.0, 1.1, -1.0, -0.9], (T, N)), dtype=torch.float32)
    dones = torch.zeros(T, N, dtype=torch.bool)
    for n in range(N):
        dones[lengths[n] - 1, n] = True
        if lengths[n] + 3 < T:
            dones[lengths[n] + 3, n] = True          # a later episode of an auto-reset env: must be ignored
    loss, info = episodic_actor_critic_loss(log_probs, values, rewards, dones, gamma, eps)
    want = 0
    for n in range(N):
        L = int(lengths[n])
        want = want + reference_finish_episode_loss([log_probs[t, n] for t in range(L)], [values[t, n].view(1) for t in range(L)],
                                                    [float(rewards[t, n]) for t in range(L)], gamma, eps)
    assert abs(float(loss) - float(want)) <= 1e-6 * max(1.0, abs(float(want))) * N      # fp32 sums in a different order
    assert int(info["episodes"]) == N and int(info["skipped_single_step"]) == 0
    g1 = torch.autograd.grad(loss, [log_probs, values], retain_graph=True)
    g2 = torch.autograd.grad(want, [log_probs, values])
    for a, b in zip(g1, g2):
        assert torch.allclose(a, b, rtol=1e-4, atol=2e-5), float((a - b).abs().max())


def test_single_step_and_unfinished_episodes_are_left_out():
    T, N = 6, 3
    dones = torch.zeros(T, N, dtype=torch.bool)
    dones[0, 0] = True                # length-1 episode: the reference's returns.std() would be nan
    dones[4, 1] = True                # a normal episode; env 2 never finishes inside the horizon
    lp, v, r = -torch.ones(T, N), torch.zeros(T, N), torch.ones(T, N)
    loss, info = episodic_actor_critic_loss(lp, v, r, dones, 0.99, 1e-7)
    assert torch.isfinite(loss) and int(info["episodes"]) == 1 and int(info["skipped_single_step"]) == 1
    only = reference_finish_episode_loss([lp[t, 1] for t in range(5)], [v[t, 1].view(1) for t in range(5)], [1.0] * 5, 0.99, 1e-7)
    assert abs(float(loss) - float(only)) < 1e-5


def test_episode_monitor_writes_the_monitor_stats_file(tmp_path):
    n, steps = 5, 40
    mon = EpisodeMonitor(n, directory=str(tmp_path), flush_every=7)
    rng = np.random.default_rng(1)
    want_len, want_ret = [], []
    run_len, run_ret = np.zeros(n, int), np.zeros(n)
    for t in range(steps):
        reward = rng.choice([0.0, 0.1, -1.0], n).astype(np.float32)
        done = rng.random(n) < 0.15
        mon.record(torch.from_numpy(reward), torch.from_numpy(done))
        run_len += 1; run_ret += reward.astype(np.float64)
        for e in range(n):
            if done[e]:
                want_len.append(int(run_len[e])); want_ret.append(float(run_ret[e]))
                run_len[e] = 0; run_ret[e] = 0.0
    path = mon.close()
    st = json.load(open(path))
    assert set(st) == {"initial_reset_timestamp", "timestamps", "episode_lengths", "episode_rewards", "episode_types"}
    assert st["episode_lengths"] == want_len and np.allclose(st["episode_rewards"], want_ret)
    assert st["episode_types"] == ["t"] * len(want_len) and len(st["timestamps"]) == len(want_len)
    manifest = [f for f in os.listdir(tmp_path) if f.endswith(".manifest.json")]
    assert len(manifest) == 1 and json.load(open(tmp_path / manifest[0]))["stats"] == os.path.basename(path)
