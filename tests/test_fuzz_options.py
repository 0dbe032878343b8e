"""Randomised option sets: the kernel logic (host build of wab_core.cuh) against the oracle, and the oracle
against the live reference for a few of them (build container only). Catches rule interactions that the
hand-picked option sets miss (odd food ratios, tiny max_turns, heavy despawn, few berries, every flag)."""
import numpy as np
import pytest

from oracle.wab_oracle import OracleEnv
from tests.hostsim import HostSimEnv
from tests.util import mask_words_to_int, pick_action, window_mask_from_bushes


def random_options(rng):
    empty = int(rng.integers(4, 61))
    fill = int(rng.integers(1, 12))
    mode = int(rng.integers(0, 3))
    return {
        "gatherer_only": mode == 1, "lookout_only": mode == 0,
        "restrict_view": bool(rng.integers(0, 2)), "starting_role": [0, 1, None][int(rng.integers(0, 3))],
        "max_turns": int(rng.integers(3, 130)), "bush_power": float(rng.choice([3, 12, 40, 100, 7.5])),
        "max_berries_per_bush": int(rng.choice([1, 2, 5, 40, 200, 255])),
        "turns_to_fill_food": fill, "turns_to_empty_food": empty,
        "starting_food": [1, 0.5, None, 1.0 / empty * int(rng.integers(1, empty + 1))][int(rng.integers(0, 4))],
        "chance_wolf_on_square": float(rng.choice([0.0, 0.001, 0.004, 0.02])),
        "wolf_chance_to_despawn": float(rng.choice([0.0, 0.05, 0.3, 0.9, 1.0])),
        "wolves": bool(rng.integers(0, 5)), "wolves_can_move": bool(rng.integers(0, 4)),
        "god_mode": bool(rng.integers(0, 4) == 0),
        "reward_per_turn": float(rng.choice([0, 0.25, -0.01])), "reward_for_eating": float(rng.choice([0.1, 0, 1.5])),
        "reward_for_finishing": float(rng.choice([1, 10])), "reward_for_starving": -1, "reward_for_being_killed": float(rng.choice([-1, -3.5])),
    }


@pytest.mark.parametrize("case", range(24))
def test_kernel_logic_matches_oracle_on_random_options(case):
    rng = np.random.default_rng(1000 + case)
    opts = random_options(rng)
    auto_reset = bool(case % 3)
    orc = OracleEnv(opts, seed=case, env_id=case * 17)
    sim = HostSimEnv(opts, seed=case, env_id=case * 17, auto_reset=auto_reset, force_f64_food=not auto_reset)
    oo = orc.reset()
    so, _ = sim.reset()
    done = False
    for n in range(500):
        if done and not auto_reset:
            oo = orc.reset()
            so, _ = sim.reset()
            done = False
        assert np.array_equal(oo[0], so[0]) and oo[1:] == so[1:], (case, n, opts)
        a = pick_action(rng, oo[0], orc.n_actions, greedy=bool(case & 1))
        oo, orr, od = orc.step(a)
        so, sr, sd, info, ovf = sim.step(a)
        assert ovf == 0 and np.float32(orr) == np.float32(sr) and od == sd, (case, n, opts)
        if auto_reset:
            if od:
                oo = orc.reset()
        else:
            done = od
        ho, hs = orc.hidden_state(), sim.hidden_state()
        assert (ho["x"], ho["y"], ho["turn"], ho["wolves"]) == (hs["x"], hs["y"], hs["turn"], hs["wolves"]), (case, n)
        assert mask_words_to_int(hs["bush_mask"]) == window_mask_from_bushes(ho), (case, n)
        if sim.game.food_mode == 0:
            assert ho["food"] == hs["food"], (case, n)


@pytest.mark.reference
@pytest.mark.parametrize("case", [3, 8, 13])
def test_oracle_matches_reference_on_random_options(case):
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference sources not present")
    rng = np.random.default_rng(1000 + case)
    opts = random_options(rng)
    ref = ref_shim.make_env(opts, seed=case, env_id=case * 17)
    orc = OracleEnv(opts, seed=case, env_id=case * 17)
    o_obs = orc.reset()
    done = False
    for n in range(90):
        if done:
            r_obs, o_obs = ref.reset(), orc.reset()
        a = pick_action(rng, o_obs[0], orc.n_actions, greedy=True)
        r_obs, rr, done, _ = ref.step(a)
        o_obs, orr, od = orc.step(a)
        for p in range(3):
            assert np.array_equal(np.asarray(r_obs[p]).astype(np.uint8), o_obs[0][p]), (case, n, p, opts)
        assert (int(r_obs[3]), int(r_obs[4]), int(r_obs[5])) == o_obs[1:] and float(rr) == orr and bool(done) == od, (case, n, opts)
        hr, ho = ref_shim.hidden_state(ref), orc.hidden_state()
        assert hr["food"] == ho["food"] and hr["wolves"] == ho["wolves"] and hr["bushes"] == ho["bushes"], (case, n)
