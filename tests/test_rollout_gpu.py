"""configs[4] building blocks on the GPU: the fused policy tail (``wab_policy_tail``: activation, clamp, both heads,
softmax, Categorical sample) against plain torch fp32, and the rollout loop built from it."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _policy(n_actions=5, seed=0):
    from wab_gym_b200.policy import Policy
    torch.manual_seed(seed)
    return Policy(449, n_actions).cuda()


@pytest.mark.parametrize("n_actions", [5, 6])
def test_policy_tail_matches_torch_fp32(n_actions):
    from wab_gym_b200.policy import policy_tail, stacked_heads
    pol = _policy(n_actions)
    with torch.no_grad():
        pol.action_head.weight.mul_(6.0)                     # spread the logits so the softmax is not flat
    n = 4099
    g = torch.Generator(device="cuda").manual_seed(1)
    z3 = torch.randn(n, 128, device="cuda", generator=g) * 3.0          # plenty of values beyond the +-4 clamp
    heads = stacked_heads(pol)
    actions = torch.empty(n, dtype=torch.uint8, device="cuda")
    value, logp = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
    probs = torch.empty(n, n_actions, device="cuda")
    ctr = torch.full((1,), 3, dtype=torch.int64, device="cuda")
    policy_tail(pol, z3, heads, actions, value=value, probs=probs, logp=logp, counter=ctr, seed=11)
    with torch.no_grad():
        x = torch.clamp(F.leaky_relu(z3), -4, 4)                          # actor_critic.py:92-93
        want_p = F.softmax(pol.action_head(x), dim=-1)                    # :96
        want_v = pol.value_head(x).squeeze(1)                             # :97
    assert torch.allclose(probs, want_p, rtol=1e-5, atol=1e-6), float((probs - want_p).abs().max())
    assert torch.allclose(value, want_v, rtol=1e-5, atol=1e-5), float((value - want_v).abs().max())
    assert int(actions.max()) < n_actions
    picked = probs.gather(1, actions.long()[:, None]).squeeze(1)
    assert torch.allclose(logp, picked.log(), rtol=1e-5, atol=1e-5) and float(picked.min()) > 0
    again = torch.empty_like(actions)
    policy_tail(pol, z3, heads, again, counter=ctr, seed=11)
    assert torch.equal(actions, again)                                    # keyed draws: same counter, same actions
    ctr += 1
    policy_tail(pol, z3, heads, again, counter=ctr, seed=11)
    assert not torch.equal(actions, again)


def test_policy_tail_samples_follow_the_probabilities():
    from wab_gym_b200.policy import policy_tail, stacked_heads
    pol = _policy(5, seed=2)
    with torch.no_grad():
        pol.action_head.weight.mul_(4.0)
    n = 400_000
    z3 = torch.randn(1, 128, device="cuda").repeat(n, 1).contiguous()
    heads = stacked_heads(pol)
    actions = torch.empty(n, dtype=torch.uint8, device="cuda")
    probs = torch.empty(n, 5, device="cuda")
    policy_tail(pol, z3, heads, actions, probs=probs, counter=torch.zeros(1, dtype=torch.int64, device="cuda"), seed=5)
    p = probs[0].double().cpu().numpy()
    freq = np.bincount(actions.cpu().numpy(), minlength=5) / n
    assert np.all(np.abs(freq - p) < 5 * np.sqrt(p * (1 - p) / n) + 1e-4), (freq, p)


@pytest.mark.parametrize("use_graph", [False, True])
def test_rollout_runs_on_the_fused_tail(use_graph):
    from wab_gym_b200 import VecEnv
    from wab_gym_b200.policy import Policy, Rollout
    torch.manual_seed(0)
    env = VecEnv(2048, seed=3, features=True)
    ro = Rollout(env, Policy(env.flat_dim, env.n_actions), use_graph=use_graph)
    assert ro.fused_tail and ro.tc_trunk and "wab_affine1_tc_kernel<2>" in ro.describe()     # the default fp32 path: one policy kernel
    before = env.stats()["steps"]
    ro.run(120)
    torch.cuda.synchronize()
    st = env.stats()
    assert st["steps"] - before == 2048 * 120 and st["bad_actions"] == 0 and st["episodes"] > 2048
    assert len(torch.unique(ro.actions)) > 1 and torch.isfinite(ro.values).all()
    env.close()


@pytest.mark.parametrize("noise", [0.0, 0.01])
@pytest.mark.parametrize("n", [128, 1000, 4099])
def test_tensor_core_first_layer_matches_torch_fp32(n, noise):
    """wab_policy_affine1 (tcgen05, bf16 x 3 operand splits) against what it replaces: the materialised policy input
    (flatten + the same keyed noise, ``actor_critic.py:188-189``) through ``Policy.affine1`` and ``leaky_relu`` (``:59``,
    ``:88-90``) in plain torch (fp64 as the yardstick, the library's fp32 GEMM beside it). fp32 accuracy: 2e-6 of the largest sum."""
    from wab_gym_b200 import VecEnv
    from wab_gym_b200.policy import Affine1TC
    env = VecEnv(n, seed=5, features=True)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(n)
    for _ in range(30):                                              # envs of every age, both roles, wolves around
        env.step(torch.randint(0, env.n_actions, (n,), dtype=torch.uint8, device="cuda", generator=g))
    pol = _policy(env.n_actions, seed=3)
    with torch.no_grad():
        pol.affine1.weight.mul_(3.0)                                 # both signs of the activation well populated
    ctr = torch.full((1,), 7, dtype=torch.int64, device="cuda")
    flat = torch.empty(n, env.flat_dim, dtype=torch.float32, device="cuda")
    env.flatten_features_noisy(env.last_features, flat, noise, ctr)
    tc = Affine1TC(env, pol)
    got = torch.full((n, 128), float("nan"), device="cuda")
    tc(env.last_features, got, noise, ctr)
    with torch.no_grad():
        want = F.leaky_relu(F.linear(flat.double(), pol.affine1.weight.double(), pol.affine1.bias.double())).float()
        lib32 = F.leaky_relu(pol.affine1(flat))
    err, err32 = float((got - want).abs().max()), float((lib32 - want).abs().max())
    assert torch.isfinite(got).all()
    scale = float(want.abs().max())                                  # sums of ~50 products of size <= 0.5
    assert err <= 2e-6 * max(scale, 1.0), (err, err32, scale)        # the library's fp32 GEMM itself is ~1.5e-6 off the fp64 result
    if noise:                                                        # a different draw counter -> different noise
        ctr += 1
        other = torch.empty_like(got)
        tc(env.last_features, other, noise, ctr)
        assert float((other - got).abs().max()) > 1e-4
    env.close()


@pytest.mark.parametrize("noise", [0.0, 0.01])
@pytest.mark.parametrize("n", [128, 1000, 4099])
def test_tensor_core_trunk_matches_torch(n, noise):
    """wab_policy_trunk (one tcgen05 kernel: input generation, affine1..3, the two hidden activations) against the
    materialised input through the three ``nn.Linear`` layers of ``Policy`` (``actor_critic.py:59-61``, ``:88-92``) in
    fp64, with the library's fp32 path beside it."""
    from wab_gym_b200 import VecEnv
    from wab_gym_b200.policy import PolicyTrunkTC
    env = VecEnv(n, seed=6, features=True)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(n + 1)
    for _ in range(30):
        env.step(torch.randint(0, env.n_actions, (n,), dtype=torch.uint8, device="cuda", generator=g))
    pol = _policy(env.n_actions, seed=4)
    with torch.no_grad():
        for lin in (pol.affine1, pol.affine2, pol.affine3):
            lin.weight.mul_(2.0)
    ctr = torch.full((1,), 9, dtype=torch.int64, device="cuda")
    flat = torch.empty(n, env.flat_dim, dtype=torch.float32, device="cuda")
    env.flatten_features_noisy(env.last_features, flat, noise, ctr)
    got = torch.full((n, 128), float("nan"), device="cuda")
    PolicyTrunkTC(env, pol)(env.last_features, got, noise, ctr)
    with torch.no_grad():
        d = pol.double()
        want = d.affine3(F.leaky_relu(d.affine2(F.leaky_relu(d.affine1(flat.double()))))).float()
        pol.float()
        lib32 = pol.affine3(F.leaky_relu(pol.affine2(F.leaky_relu(pol.affine1(flat)))))
    err, err32, scale = float((got - want).abs().max()), float((lib32 - want).abs().max()), float(want.abs().max())
    assert torch.isfinite(got).all()
    assert err <= 3e-6 * max(scale, 1.0), (err, err32, scale)
    env.close()


def test_policy_forward_in_one_launch_equals_trunk_then_tail():
    """wab_policy_forward (trunk + tail in one launch) against wab_policy_trunk followed by wab_policy_tail: same z3, same
    value to fp32 rounding of a 128-term sum, same sampled action except where the uniform lands within that rounding of a
    CDF step."""
    from wab_gym_b200 import VecEnv
    from wab_gym_b200.policy import PolicyTrunkTC, policy_tail, stacked_heads
    n = 5000
    env = VecEnv(n, seed=8, features=True)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(2)
    for _ in range(25):
        env.step(torch.randint(0, env.n_actions, (n,), dtype=torch.uint8, device="cuda", generator=g))
    pol = _policy(env.n_actions, seed=5)
    with torch.no_grad():
        pol.action_head.weight.mul_(6.0)
    heads = stacked_heads(pol)
    trunk = PolicyTrunkTC(env, pol)
    ctr = torch.full((1,), 4, dtype=torch.int64, device="cuda")
    z3 = torch.empty(n, 128, device="cuda")
    trunk(env.last_features, z3, 0.01, ctr)
    a_ref = torch.empty(n, dtype=torch.uint8, device="cuda")
    v_ref, lp_ref, p_ref = torch.empty(n, device="cuda"), torch.empty(n, device="cuda"), torch.empty(n, env.n_actions, device="cuda")
    policy_tail(pol, z3, heads, a_ref, value=v_ref, probs=p_ref, logp=lp_ref, counter=ctr, seed=21)
    a, v, lp, p = torch.full_like(a_ref, 255), torch.empty_like(v_ref), torch.empty_like(lp_ref), torch.empty_like(p_ref)
    z3b = torch.empty_like(z3)
    trunk.forward_sample(env.last_features, heads, a, value=v, probs=p, logp=lp, noise_scale=0.01, counter=ctr, seed=21, z3=z3b)
    assert torch.equal(z3b, z3)
    assert torch.allclose(v, v_ref, rtol=1e-5, atol=1e-5) and torch.allclose(p, p_ref, rtol=1e-5, atol=1e-6)
    same = (a == a_ref)
    assert same.float().mean().item() > 0.999 and torch.allclose(lp[same], lp_ref[same], rtol=1e-5, atol=1e-5)
    env.close()


def test_rollout_with_tensor_core_first_layer_equals_library_first_layer():
    """Same seeds, same noise counter: the rollout with the tcgen05 first layer takes the same actions as the one that
    materialises the input and calls the library GEMM (fp32 differences of 1e-6 do not move a sampled action here)."""
    from wab_gym_b200 import VecEnv
    from wab_gym_b200.policy import Rollout
    acts = []
    for tc in (False, True, None):                       # library GEMMs | tcgen05 first layer | tcgen05 trunk (the default)
        torch.manual_seed(0)
        env = VecEnv(2048, seed=9, features=True)
        ro = Rollout(env, _policy(env.n_actions, seed=1), tc_first_layer=tc, tc_trunk=None if tc is None else False)
        seq = []
        for _ in range(12):
            ro.step()
            seq.append(ro.actions.clone())
        acts.append(torch.stack(seq))
        env.close()
    for other in acts[1:]:
        same = (acts[0] == other).float().mean().item()
        assert same > 0.999, same
