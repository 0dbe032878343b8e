"""Batched on-device actor-critic update — the reference's ``finish_episode`` (``actor_critic.py:128-169``)
for N lockstep environments.

The reference trains on ONE environment, one episode at a time: discounted Monte-Carlo returns
(``:139-144``), normalised (``:146-147``), policy loss ``-log_prob * (R - value.item())`` (``:150-153``),
critic loss ``smooth_l1(value, R)`` (``:156``), summed, Adam (``:159-165``); it synchronises with the host at
every step (``.item()``, ``:125``) and walks Python lists per episode (``:139-155``). Here a fixed horizon of T
lockstep steps of N environments is collected with every tensor on the device (observations come from the
fused feature kernels), returns are computed with a reverse scan that restarts at episode boundaries, the
unfinished tail of each environment is bootstrapped with the critic, and one optimiser step is taken per
horizon. No host synchronisation happens inside ``train_iteration`` except the optional logging read.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

from .policy import Policy
from .vec_env import VecEnv


class A2CTrainer:
    def __init__(self, env: VecEnv, policy: Policy = None, horizon: int = 80, gamma: float = 0.99, lr: float = 3e-3,
                 noise: bool = True):
        if not env.with_features:
            raise ValueError("A2CTrainer needs VecEnv(features=True)")
        self.env, self.horizon, self.gamma, self.noise = env, int(horizon), float(gamma), noise
        self.policy = (policy or Policy(env.flat_dim, env.n_actions)).to(env.device)
        self.optimizer = torch.optim.Adam(self.policy.parameters(), lr=lr)      # actor_critic.py:103
        self.eps = torch.finfo(torch.float32).eps                                # :104
        env.reset()

    def _observe(self) -> torch.Tensor:
        x = self.env.flatten_features(self.env.last_features)                    # gym.spaces.flatten, :188
        if self.noise:
            x = x + torch.rand_like(x) / 100                                     # :189
        return x

    def train_iteration(self) -> Dict[str, torch.Tensor]:
        env, T = self.env, self.horizon
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
        with torch.no_grad():
            _, bootstrap = self.policy(self._observe())                          # critic value of the unfinished tail
            running = bootstrap.squeeze(1)
            returns = torch.empty(T, env.num_envs, device=env.device)
            for t in range(T - 1, -1, -1):                                       # R = r + gamma * R, restarted at done (:139-144)
                running = rewards[t] + self.gamma * running * (~dones[t]).float()
                returns[t] = running
            returns = (returns - returns.mean()) / (returns.std() + self.eps)    # :146-147
        log_probs, values = torch.stack(log_probs), torch.stack(values)
        advantage = returns - values.detach()                                    # R - value.item(), :150
        policy_loss = -(log_probs * advantage).sum(0).mean()                     # :153, summed over time, mean over envs
        value_loss = F.smooth_l1_loss(values, returns, reduction="none").sum(0).mean()   # :156
        loss = policy_loss + value_loss                                          # :162
        self.optimizer.zero_grad(set_to_none=True)                               # :159
        loss.backward()                                                          # :165
        self.optimizer.step()
        return {"loss": loss.detach(), "policy_loss": policy_loss.detach(), "value_loss": value_loss.detach(),
                "mean_reward": torch.stack(rewards).mean()}
