"""Actor-critic policy and device-resident rollout (BASELINE config 5).

The reference's consumer of the environment is ``actor_critic.py``: a 4-layer MLP
(``Policy``, ``actor_critic.py:54-97``) fed with ``gym.spaces.flatten`` of the ``PragmaticObsWrapper``
observation plus ``U[0,1)/100`` input noise (``:188-189``), a ``Categorical`` sample per step (``:108-125``,
with a host sync per step through ``.item()``), stepping ONE environment. Here the same network
consumes the flattened features of N environments straight from HBM: features and the 449-column
one-hot come from the CUDA kernels (``wab_features.cuh``), the MLP is plain torch (cuBLAS GEMMs — a
library GEMM, not a hot op of this path), sampling stays on the device and nothing synchronises with
the host inside the loop.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .vec_env import VecEnv, _ptr, _raw_stream


class Policy(nn.Module):
    """Same architecture as the reference's ``Policy`` (``actor_critic.py:54-97``)."""

    def __init__(self, obs_dim: int, n_actions: int):
        super().__init__()
        self.affine1 = nn.Linear(obs_dim, 128)
        self.affine2 = nn.Linear(128, 150)
        self.affine3 = nn.Linear(150, 128)
        self.action_head = nn.Linear(128, n_actions)
        self.value_head = nn.Linear(128, 1)

    def forward(self, x):
        x = F.leaky_relu(self.affine1(x))
        x = F.leaky_relu(self.affine2(x))
        x = F.leaky_relu(self.affine3(x))
        x = torch.clamp(x, -4, 4)
        return F.softmax(self.action_head(x), dim=-1), self.value_head(x)


def policy_tail(policy: "Policy", z3: torch.Tensor, heads: tuple, actions: torch.Tensor, value: Optional[torch.Tensor] = None,
                probs: Optional[torch.Tensor] = None, logp: Optional[torch.Tensor] = None,
                counter: Optional[torch.Tensor] = None, seed: int = 0) -> torch.Tensor:
    """``clamp(leaky_relu(z3)) -> action_head / value_head -> softmax -> Categorical.sample()`` (``actor_critic.py:92-97``,
    ``:114-120``) as ONE kernel of the library (``wab_policy_tail``): ``z3`` f32[N, 128] is ``policy.affine3``'s output
    before its activation, ``heads`` = ``stacked_heads(policy)``. Fills ``actions`` u8[N] (and ``value``, ``probs``,
    ``logp`` when given)."""
    w, b = heads
    if z3.dtype != torch.float32 or not z3.is_contiguous() or z3.shape[1] != 128:
        raise ValueError("z3 must be a contiguous float32 [N, 128] tensor")
    dev = z3.device.index if z3.device.index is not None else torch.cuda.current_device()
    stream = ctypes.c_void_p(_raw_stream(dev)) if _raw_stream is not None else ctypes.c_void_p(torch.cuda.current_stream(z3.device).cuda_stream)
    _lib.check(_lib.load().wab_policy_tail(_ptr(z3), 128, _ptr(w), _ptr(b), z3.shape[0], w.shape[0] - 1, 0.01, -4.0, 4.0,
                                           int(seed) & (2 ** 64 - 1), _ptr(counter), _ptr(actions), _ptr(value), _ptr(probs),
                                           _ptr(logp), stream))
    return actions


def stacked_heads(policy: "Policy") -> tuple:
    """(action_head.weight stacked on value_head.weight f32[A + 1, 128], the two biases f32[A + 1]) — refresh after an
    optimiser step."""
    with torch.no_grad():
        w = torch.cat([policy.action_head.weight, policy.value_head.weight], 0).float().contiguous()
        b = torch.cat([policy.action_head.bias, policy.value_head.bias], 0).float().contiguous()
    return w, b


class Affine1TC:
    """``leaky_relu(policy.affine1(flatten(obs) + noise))`` (``actor_critic.py:59``, ``:88-90``, ``:188-189``) as ONE tcgen05
    kernel of the library (``wab_policy_affine1``): the 449-wide input is generated inside the kernel from the 28 feature
    bytes per environment — same keyed noise as ``VecEnv.flatten_features_noisy`` — and the product runs on the tensor
    cores with bf16 x 3 operand splits and fp32 accumulation, i.e. to fp32 accuracy. ``refresh()`` re-packs the weights
    after an optimiser step."""

    def __init__(self, env: VecEnv, policy: "Policy"):
        if policy.affine1.out_features != 128 or policy.affine1.in_features != env.flat_dim:
            raise ValueError("Affine1TC is built for the reference's affine1: flat_dim -> 128")
        self.env, self.policy, self.lib = env, policy, _lib.load()
        self.packed = torch.empty(int(self.lib.wab_policy_affine1_packed_bytes()), dtype=torch.uint8, device=env.device)
        self.bias = None
        self.refresh()

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.env.device).cuda_stream)

    @torch.no_grad()
    def refresh(self):
        w = self.policy.affine1.weight.detach().float().contiguous()
        b = self.policy.affine1.bias.detach().float().contiguous()
        self.bias = b.clone() if self.bias is None else self.bias.copy_(b)           # in place: captured graphs hold the pointer
        _lib.check(self.lib.wab_policy_affine1_prepare(_ptr(w), w.shape[1], _ptr(self.packed), self._stream()))

    def __call__(self, features: torch.Tensor, out: torch.Tensor, noise_scale: float = 0.01,
                 counter: Optional[torch.Tensor] = None, slope: float = 0.01) -> torch.Tensor:
        if out.dtype != torch.float32 or not out.is_contiguous() or tuple(out.shape) != (features.shape[0], 128):
            raise ValueError("out must be a contiguous float32 [N, 128] tensor")
        _lib.check(self.lib.wab_policy_affine1(self.env._h, _ptr(features), features.shape[0], _ptr(self.packed), _ptr(self.bias),
                                               float(noise_scale), float(slope), _ptr(counter), _ptr(out), self._stream()))
        return out


class PolicyTrunkTC(Affine1TC):
    """The whole trunk ``affine3(leaky_relu(affine2(leaky_relu(affine1(flatten(obs) + noise)))))`` (``actor_critic.py:88-92``)
    as ONE tcgen05 kernel (``wab_policy_trunk``): returns ``z3`` f32[N, 128], the pre-activation output of ``affine3`` that
    ``policy_tail`` takes. The hidden activations never leave the SM."""

    def __init__(self, env: VecEnv, policy: "Policy"):
        if (policy.affine2.in_features, policy.affine2.out_features, policy.affine3.in_features, policy.affine3.out_features) != (128, 150, 150, 128):
            raise ValueError("PolicyTrunkTC is built for the reference's trunk: 128 -> 150 -> 128")
        lib = _lib.load()
        self.packed2 = torch.empty(int(lib.wab_policy_linear_packed_bytes(150, 128)), dtype=torch.uint8, device=env.device)
        self.packed3 = torch.empty(int(lib.wab_policy_linear_packed_bytes(128, 150)), dtype=torch.uint8, device=env.device)
        self.bias2 = self.bias3 = None
        super().__init__(env, policy)

    @torch.no_grad()
    def refresh(self):
        super().refresh()
        p = self.policy
        w2, w3 = p.affine2.weight.detach().float().contiguous(), p.affine3.weight.detach().float().contiguous()
        b2, b3 = p.affine2.bias.detach().float().contiguous(), p.affine3.bias.detach().float().contiguous()
        self.bias2 = b2.clone() if self.bias2 is None else self.bias2.copy_(b2)
        self.bias3 = b3.clone() if self.bias3 is None else self.bias3.copy_(b3)
        _lib.check(self.lib.wab_policy_linear_prepare(_ptr(w2), 150, 128, _ptr(self.packed2), self._stream()))
        _lib.check(self.lib.wab_policy_linear_prepare(_ptr(w3), 128, 150, _ptr(self.packed3), self._stream()))

    def __call__(self, features: torch.Tensor, out: torch.Tensor, noise_scale: float = 0.01,
                 counter: Optional[torch.Tensor] = None, slope: float = 0.01) -> torch.Tensor:
        if out.dtype != torch.float32 or not out.is_contiguous() or tuple(out.shape) != (features.shape[0], 128):
            raise ValueError("out must be a contiguous float32 [N, 128] tensor")
        _lib.check(self.lib.wab_policy_trunk(self.env._h, _ptr(features), features.shape[0], _ptr(self.packed), _ptr(self.bias),
                                             _ptr(self.packed2), _ptr(self.bias2), 150, _ptr(self.packed3), _ptr(self.bias3),
                                             float(noise_scale), float(slope), _ptr(counter), _ptr(out), self._stream()))
        return out


    def forward_sample(self, features: torch.Tensor, heads: tuple, actions: torch.Tensor, value: Optional[torch.Tensor] = None,
                       probs: Optional[torch.Tensor] = None, logp: Optional[torch.Tensor] = None, noise_scale: float = 0.01,
                       counter: Optional[torch.Tensor] = None, seed: int = 0, z3: Optional[torch.Tensor] = None,
                       slope: float = 0.01) -> torch.Tensor:
        """``Policy.forward`` + ``select_action`` (``actor_critic.py:84-97``, ``:108-125``) in ONE launch (``wab_policy_forward``):
        the trunk above, then clamp, both heads, softmax and the Categorical sample of ``policy_tail`` on the accumulators'
        registers. ``heads`` = ``stacked_heads(policy)``; fills ``actions`` u8[N] (and ``value``, ``probs``, ``logp``, ``z3`` when given)."""
        w, b = heads
        _lib.check(self.lib.wab_policy_forward(self.env._h, _ptr(features), features.shape[0], _ptr(self.packed), _ptr(self.bias),
                                               _ptr(self.packed2), _ptr(self.bias2), 150, _ptr(self.packed3), _ptr(self.bias3),
                                               _ptr(w), _ptr(b), w.shape[0] - 1, float(noise_scale), float(slope), -4.0, 4.0,
                                               int(seed) & (2 ** 64 - 1), _ptr(counter), _ptr(actions), _ptr(value), _ptr(probs),
                                               _ptr(logp), _ptr(z3), self._stream()))
        return actions


class Rollout:
    """N-environment rollout with every tensor resident on the device.

    ``step()`` = flatten features -> + U[0,1)/100 noise -> policy -> Categorical sample -> env.step.
    With ``use_graph=True`` one step is captured in a CUDA graph and replayed (the loop is launch-bound
    for small batches)."""

    def __init__(self, env: VecEnv, policy: Optional[Policy] = None, noise: bool = True, use_graph: bool = False,
                 dtype: torch.dtype = torch.float32, fused_tail: Optional[bool] = None, track_reward: bool = False,
                 tc_first_layer: Optional[bool] = None, tc_trunk: Optional[bool] = None):
        if not env.with_features:
            raise ValueError("Rollout needs VecEnv(features=True)")
        self.env, self.noise, self.dtype = env, noise, dtype
        self.policy = (policy or Policy(env.flat_dim, env.n_actions)).to(env.device).to(dtype).eval()
        self.flat = torch.empty(env.num_envs, env.flat_dim, dtype=dtype, device=env.device)   # policy input, its dtype
        self.noise_ctr = torch.zeros(1, dtype=torch.int64, device=env.device)                  # draw counter (graph-safe)
        self.sample_seed = int(torch.initial_seed()) & (2 ** 63 - 1)
        self.actions = torch.zeros(env.num_envs, dtype=torch.uint8, device=env.device)
        self.values = torch.zeros(env.num_envs, dtype=torch.float32, device=env.device)
        self.reward_sum = torch.zeros((), dtype=torch.float64, device=env.device)
        self.track_reward = track_reward         # episode statistics live in env.stats(); summing rewards here is optional
        # fp32: the trunk's three GEMMs stay library calls (cuBLAS), everything after affine3 is one kernel of this library
        self.fused_tail = (dtype == torch.float32) if fused_tail is None else bool(fused_tail)
        if self.fused_tail and dtype != torch.float32:
            raise ValueError("the fused policy tail is an fp32 kernel")
        self.heads = stacked_heads(self.policy) if self.fused_tail else None
        # fp32: the first layer (input generation + 449 x 128 product + activation) is one tcgen05 kernel of this library
        self.tc_first_layer = self.fused_tail if tc_first_layer is None else bool(tc_first_layer)
        if self.tc_first_layer and not self.fused_tail:
            raise ValueError("the tensor-core first layer belongs to the fp32 fused path")
        # ... and so is the whole trunk (the default): features -> z3 without the hidden activations leaving the SM
        self.tc_trunk = self.tc_first_layer if tc_trunk is None else bool(tc_trunk)
        if self.tc_trunk and not self.tc_first_layer:
            raise ValueError("the tensor-core trunk includes the tensor-core first layer")
        self.affine1_tc = (PolicyTrunkTC if self.tc_trunk else Affine1TC)(env, self.policy) if self.tc_first_layer else None
        self.h1 = torch.empty(env.num_envs, 128, dtype=torch.float32, device=env.device) if self.tc_first_layer else None
        env.reset()
        self.graph = None
        if use_graph:
            side = torch.cuda.Stream(device=env.device)
            side.wait_stream(torch.cuda.current_stream(env.device))
            with torch.cuda.stream(side):
                for _ in range(3):
                    self._step_eager()
                side.synchronize()
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph, stream=side):
                    self._step_eager()
            torch.cuda.current_stream(env.device).wait_stream(side)

    @torch.no_grad()
    def _step_eager(self):
        env = self.env
        if self.tc_trunk:
            # Policy.forward + select_action (actor_critic.py:84-97, :108-125, :188-189) in one tcgen05 kernel
            self.noise_ctr += 1
            self.affine1_tc.forward_sample(env.last_features, self.heads, self.actions, value=self.values,
                                           noise_scale=0.01 if self.noise else 0.0, counter=self.noise_ctr, seed=self.sample_seed)
            _, reward, _, _ = env.step(self.actions)
            if self.track_reward:
                self.reward_sum += reward.sum(dtype=torch.float64)
            return
        if self.tc_first_layer:
            # flatten (actor_critic.py:188) + U[0,1)/100 noise (:189) + affine1 + leaky_relu (:88-90): one tcgen05 kernel
            h = self.affine1_tc(env.last_features, self.h1, 0.01 if self.noise else 0.0, self.noise_ctr)
        else:
            # gym.spaces.flatten (actor_critic.py:188) + U[0,1)/100 input noise (:189) + cast, one kernel
            env.flatten_features_noisy(env.last_features, self.flat, 0.01 if self.noise else 0.0, self.noise_ctr)
        self.noise_ctr += 1
        if self.fused_tail:
            p = self.policy
            if not self.tc_first_layer:
                h = F.leaky_relu(p.affine1(self.flat))                                   # :88-90 (cuBLAS)
            z3 = h if self.tc_trunk else p.affine3(F.leaky_relu(p.affine2(h)))           # the trunk kernel already returned z3
            policy_tail(p, z3, self.heads, self.actions, value=self.values, counter=self.noise_ctr, seed=self.sample_seed)
        else:
            probs, value = self.policy(self.flat)
            env.sample_actions(probs, self.actions, self.noise_ctr, seed=self.sample_seed)   # Categorical(probs).sample(), :117-120
            self.values.copy_(value.squeeze(1))
        _, reward, _, _ = env.step(self.actions)
        if self.track_reward:
            self.reward_sum += reward.sum(dtype=torch.float64)

    def describe(self) -> str:
        if self.tc_trunk:
            return ("wab_affine1_tc_kernel<2> (ONE tcgen05 kernel: flatten + noise + affine1..3 + leaky_relu with bf16 x 3 splits and fp32 "
                    "accumulation, activations stay on the SM; then clamp, both heads, softmax, Categorical sample) -> wab_step_kernel; "
                    "one CUDA graph per step")
        if self.tc_first_layer:
            return ("wab_affine1_tc_kernel (flatten + noise + affine1 + leaky_relu: tcgen05, bf16 x 3 splits, fp32 accumulate) -> "
                    "affine2, affine3 fp32 (cuBLAS) + leaky_relu -> wab_policy_tail_kernel (activation, clamp, both heads, softmax, "
                    "Categorical sample) -> wab_step_kernel; one CUDA graph per step")
        if self.fused_tail:
            return ("wab_flatten_noisy_kernel (flatten + noise) -> affine1..3 fp32 (cuBLAS) + leaky_relu -> wab_policy_tail_kernel "
                    "(activation, clamp, both heads, softmax, Categorical sample) -> wab_step_kernel; one CUDA graph per step")
        return "wab_flatten_noisy_kernel -> torch MLP (cuBLAS) -> wab_sample_kernel -> wab_step_kernel; one CUDA graph per step"

    def refresh_weights(self):
        """After an optimiser step: re-pack what the kernels hold of the policy (stacked heads, split tensor-core operands).
        The buffers are updated in place, so a captured graph keeps working."""
        if self.heads is not None:
            w, b = stacked_heads(self.policy)
            self.heads[0].copy_(w); self.heads[1].copy_(b)
        if self.affine1_tc is not None:
            self.affine1_tc.refresh()

    def step(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_eager()

    def run(self, steps: int) -> Dict[str, float]:
        for _ in range(steps):
            self.step()
        return {"env_steps": steps * self.env.num_envs}
