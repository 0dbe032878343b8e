"""Batched on-device actor-critic update — the reference's ``finish_episode`` (``actor_critic.py:128-169``)
for N lockstep environments.

The reference trains on ONE environment, one episode per update: discounted Monte-Carlo returns over the episode
(``:139-144``), normalised by the episode's own mean and (unbiased) standard deviation (``:146-147``), policy loss
``-log_prob * (R - value.item())`` (``:150-153``), critic loss ``smooth_l1(value, R)`` (``:156``), both SUMMED over
the episode (``:162``), Adam (``:159-165``); it synchronises with the host at every step (``.item()``, ``:125``) and
walks Python lists per episode (``:139-155``).

``episodic_actor_critic_loss`` is that loss for a batch of episodes at once — the sum over episodes of the
reference's per-episode loss (N episodes accumulated into one optimiser step; with N = 1 it IS ``finish_episode``).
``A2CTrainer.train_iteration`` collects one episode per environment with every tensor on the device: all
environments are reset together, stepped ``max_turns`` times (every v1 episode ends by then, ``wab_env.py:328-340``)
and each environment contributes the steps up to and including its first ``done``; what an auto-reset environment
does after that is masked out. The policy input is the keyed flatten + noise kernel of the rollout
(``actor_critic.py:188-189``). Deliberate difference: an episode of length 1 has no standard deviation (the
reference computes ``nan`` there and poisons its weights); such episodes are left out of the loss and counted.

``mode="horizon"`` keeps the fixed-horizon variant of round 1 (returns restart at episode boundaries, the unfinished
tail is bootstrapped with the critic) — a different estimator, not the reference's update.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from .policy import Policy
from .vec_env import VecEnv


def episodic_actor_critic_loss(log_probs: torch.Tensor, values: torch.Tensor, rewards: torch.Tensor, dones: torch.Tensor,
                               gamma: float, eps: float) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """``finish_episode`` (``actor_critic.py:128-165``) for N episodes laid out as [T, N] tensors: env n's episode is
    steps 0 .. first done (inclusive); later steps are ignored. Returns (loss, diagnostics)."""
    T, N = rewards.shape
    d = dones.to(torch.bool)
    ended_before = torch.cumsum(d.to(torch.int32), 0) - d.to(torch.int32)      # dones strictly before step t
    mask = ended_before == 0                                                   # steps of the first episode
    complete = d.any(0)                                                        # envs whose episode ended inside T
    mask = mask & complete
    m = mask.to(rewards.dtype)
    returns = torch.zeros_like(rewards)
    running = torch.zeros(N, dtype=rewards.dtype, device=rewards.device)
    for t in range(T - 1, -1, -1):                                             # R = r + gamma * R  (:139-144)
        running = (rewards[t] + gamma * running) * m[t]
        returns[t] = running
    length = m.sum(0)
    usable = length >= 2                                                       # std of one sample is nan (:147)
    m = m * usable.to(m.dtype)
    safe_len = length.clamp(min=2)
    mean = (returns * m).sum(0) / safe_len
    var = (((returns - mean) * m) ** 2).sum(0) / (safe_len - 1)                # torch.std: unbiased
    norm = (returns - mean) / (var.sqrt() + eps)                               # :146-147
    advantage = norm - values.detach()                                         # R - value.item(), :150
    policy_loss = (-(log_probs * advantage) * m).sum()                         # :153, :162
    value_loss = (F.smooth_l1_loss(values, norm, reduction="none") * m).sum()  # :156, :162
    return policy_loss + value_loss, {"policy_loss": policy_loss.detach(), "value_loss": value_loss.detach(),
                                      "episodes": usable.sum(), "skipped_single_step": (complete & ~usable).sum(),
                                      "mean_length": (length * usable).sum() / usable.sum().clamp(min=1)}


class A2CTrainer:
    def __init__(self, env: VecEnv, policy: Policy = None, horizon: int = None, gamma: float = 0.99, lr: float = 3e-2,
                 noise: bool = True, mode: str = "episodic"):
        if not env.with_features:
            raise ValueError("A2CTrainer needs VecEnv(features=True)")
        if mode not in ("episodic", "horizon"):
            raise ValueError("mode must be 'episodic' or 'horizon'")
        self.env, self.gamma, self.noise, self.mode = env, float(gamma), noise, mode
        self.horizon = int(horizon) if horizon is not None else int(env.game_options["max_turns"])
        self.policy = (policy or Policy(env.flat_dim, env.n_actions)).to(env.device)
        self.optimizer = torch.optim.Adam(self.policy.parameters(), lr=lr)      # actor_critic.py:103 (lr = 3e-2)
        self.eps = torch.finfo(torch.float32).eps                                # :104
        self._flat = torch.empty(env.num_envs, env.flat_dim, dtype=torch.float32, device=env.device)
        self._ctr = torch.zeros(1, dtype=torch.int64, device=env.device)
        env.reset()

    def _observe(self) -> torch.Tensor:
        # gym.spaces.flatten (:188) + np.random.rand(...) / 100 (:189): the keyed kernel the rollout uses
        self.env.flatten_features_noisy(self.env.last_features, self._flat, 0.01 if self.noise else 0.0, self._ctr)
        self._ctr += 1
        return self._flat.clone()        # autograd keeps a reference to the policy input of every step

    def train_iteration(self) -> Dict[str, torch.Tensor]:
        env, T = self.env, self.horizon
        if self.mode == "episodic":
            env.reset()                                                          # every env starts an episode (:179)
        log_probs, values, rewards, dones = [], [], [], []
        for _ in range(T):
            probs, value = self.policy(self._observe())
            dist = torch.distributions.Categorical(probs)                        # :114
            action = dist.sample()                                               # :117
            log_probs.append(dist.log_prob(action))                              # :120
            values.append(value.squeeze(1))
            _, reward, done, _ = env.step(action.to(torch.uint8))
            rewards.append(reward.clone())
            dones.append(done.clone())
        log_probs, values = torch.stack(log_probs), torch.stack(values)
        rewards, dones = torch.stack(rewards), torch.stack(dones)
        if self.mode == "episodic":
            loss, info = episodic_actor_critic_loss(log_probs, values, rewards, dones, self.gamma, self.eps)
        else:
            with torch.no_grad():
                _, bootstrap = self.policy(self._observe())                      # critic value of the unfinished tail
                running = bootstrap.squeeze(1)
                returns = torch.empty(T, env.num_envs, device=env.device)
                for t in range(T - 1, -1, -1):                                   # R = r + gamma * R, restarted at done
                    running = rewards[t] + self.gamma * running * (~dones[t]).float()
                    returns[t] = running
                returns = (returns - returns.mean()) / (returns.std() + self.eps)
            advantage = returns - values.detach()
            policy_loss = -(log_probs * advantage).sum(0).mean()
            value_loss = F.smooth_l1_loss(values, returns, reduction="none").sum(0).mean()
            loss = policy_loss + value_loss
            info = {"policy_loss": policy_loss.detach(), "value_loss": value_loss.detach()}
        self.optimizer.zero_grad(set_to_none=True)                               # :159
        loss.backward()                                                          # :165
        self.optimizer.step()
        info.update({"loss": loss.detach(), "mean_reward": rewards.mean()})
        return info
