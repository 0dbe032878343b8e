"""Monitor-style episode statistics for a lockstep batch.

The reference wraps its single env in ``gym.wrappers.Monitor(env, directory=outdir, force=True)``
(``wab_env.py:1012-1013``, ``actor_critic.py:47``), whose durable product — video aside — is the stats file
``openaigym.episode_batch.<n>.<pid>.stats.json`` = ``{"initial_reset_timestamp", "timestamps", "episode_lengths",
"episode_rewards", "episode_types"}`` plus a manifest naming it. ``EpisodeMonitor`` writes the same two files for N
lockstep environments: per-env running length and return live on the device next to the env state, finished
episodes are parked in a [K, N] device ring by pure tensor ops (no host synchronisation inside ``record``), and the
ring is read back once every ``flush_every`` steps. Episodes appear in the file in (step, env id) order.
"""
from __future__ import annotations

import json
import os
import time
from typing import List, Optional

import torch


class EpisodeMonitor:
    def __init__(self, num_envs: int, directory: Optional[str] = None, device="cpu", flush_every: int = 256,
                 env_id: str = "WolvesAndBushes-v0", force: bool = True):
        self.num_envs, self.directory, self.env_id = int(num_envs), directory, env_id
        self.device = torch.device(device)
        self.flush_every = int(flush_every)
        self._len = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
        self._ret = torch.zeros(self.num_envs, dtype=torch.float64, device=self.device)
        self._ring_len = torch.zeros((self.flush_every, self.num_envs), dtype=torch.int32, device=self.device)
        self._ring_ret = torch.zeros((self.flush_every, self.num_envs), dtype=torch.float64, device=self.device)
        self._fill = 0
        self.initial_reset_timestamp = time.time()
        self._fill_times: List[float] = []
        self.episode_lengths: List[int] = []
        self.episode_rewards: List[float] = []
        self.timestamps: List[float] = []
        if directory is not None:
            os.makedirs(directory, exist_ok=True)
            if force:                                   # Monitor(force=True) clears earlier monitor files
                for f in os.listdir(directory):
                    if f.startswith("openaigym."):
                        os.remove(os.path.join(directory, f))

    def record(self, reward: torch.Tensor, done: torch.Tensor) -> None:
        """Account one lockstep step (``reward`` f32[N], ``done`` bool/u8[N] as returned by ``VecEnv.step``)."""
        d = done.to(torch.bool)
        self._len += 1
        self._ret += reward.to(torch.float64)
        self._ring_len[self._fill] = torch.where(d, self._len, torch.zeros_like(self._len))
        self._ring_ret[self._fill] = torch.where(d, self._ret, torch.zeros_like(self._ret))
        self._len.masked_fill_(d, 0)
        self._ret.masked_fill_(d, 0.0)
        self._fill_times.append(time.time())
        self._fill += 1
        if self._fill == self.flush_every:
            self.flush()

    def flush(self) -> None:
        """Read the ring back (the only device-to-host transfer) and append its finished episodes."""
        if self._fill == 0:
            return
        lens = self._ring_len[:self._fill].cpu()
        rets = self._ring_ret[:self._fill].cpu()
        steps, envs = torch.nonzero(lens, as_tuple=True)
        for s, e in zip(steps.tolist(), envs.tolist()):
            self.episode_lengths.append(int(lens[s, e]))
            self.episode_rewards.append(float(rets[s, e]))
            self.timestamps.append(self._fill_times[s])
        self._fill = 0
        self._fill_times = []

    def stats(self) -> dict:
        self.flush()
        return {"initial_reset_timestamp": self.initial_reset_timestamp, "timestamps": self.timestamps,
                "episode_lengths": self.episode_lengths, "episode_rewards": self.episode_rewards,
                "episode_types": ["t"] * len(self.episode_lengths)}

    def close(self) -> Optional[str]:
        """Write the stats file and the manifest (what ``Monitor.close`` leaves on disk). Returns the stats path."""
        st = self.stats()
        if self.directory is None:
            return None
        base = "openaigym.episode_batch.0.%d" % os.getpid()
        path = os.path.join(self.directory, base + ".stats.json")
        with open(path, "w") as fh:
            json.dump(st, fh)
        with open(os.path.join(self.directory, "openaigym.manifest.0.%d.manifest.json" % os.getpid()), "w") as fh:
            json.dump({"stats": os.path.basename(path), "videos": [],
                       "env_info": {"gym_version": None, "env_id": self.env_id, "num_envs": self.num_envs}}, fh)
        return path
