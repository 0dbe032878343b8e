"""Host logic: thresholds, action tables, food proof (no GPU)."""
import math
from fractions import Fraction

import numpy as np
import pytest

from wab_gym_b200 import config as C


def test_default_options_match_reference_keys():
    assert C.default_game_options["max_turns"] == 80 and C.default_game_options["reward_for_eating"] == 0.1
    assert len(C.default_game_options) == 23
    from oracle import ref_shim
    if ref_shim.reference_available():
        ref = ref_shim.load_reference()
        assert ref.default_game_options == C.default_game_options
        env = ref_shim.make_env()
        assert np.array_equal(env.lookout_tile_mask, C.LOOKOUT_TILE_MASK)
        assert np.array_equal(env.gatherer_tile_mask, C.GATHERER_TILE_MASK)


def test_tile_masks_shape():
    assert C.LOOKOUT_TILE_MASK.sum() == 24 and (1 - C.GATHERER_TILE_MASK).sum() == 21
    assert C.LOOKOUT_TILE_MASK[5, 5] == 0 and C.GATHERER_TILE_MASK[5, 5] == 0


@pytest.mark.parametrize("opts,n,last", [
    ({"gatherer_only": False, "lookout_only": True}, 5, (0, 0, 0)),
    ({"gatherer_only": True, "lookout_only": True}, 5, (0, 0, 1)),     # gatherer_only wins (wab_env.py:149)
    ({"gatherer_only": False, "lookout_only": False}, 6, (0, 0, 0)),
])
def test_action_table(opts, n, last):
    t = C.action_table({**C.default_game_options, **opts})
    assert len(t) == n and t[:4] == [(0, 1, -1), (1, 0, -1), (0, -1, -1), (-1, 0, -1)] and t[-1] == last
    if n == 6:
        assert t[4] == (0, 0, 1)


@pytest.mark.parametrize("p", [0.0005, 0.05, 0.001 / 2, 0.25, 1e-9, 0.999999, 0.0, 2.0 ** -32, 3 * 2.0 ** -32])
def test_probability_thresholds_are_exact(p):
    lt, gt = C.lt_threshold(p), C.gt_threshold(p)
    for w in {0, 1, lt - 2, lt - 1, lt, lt + 1, gt - 2, gt - 1, gt, gt + 1, 2 ** 32 - 1}:
        if 0 <= w < 2 ** 32:
            u = w * 2.0 ** -32
            assert (u < p) == (w < lt), (p, w)
            assert (u > p) == (w >= gt), (p, w)


def test_binomial_first_tables():
    from oracle import keyed_rng as kr
    for n, p in ((48, 0.0005), (121, 0.0005), (48, 0.006), (48, 0.0), (121, 0.01)):
        t = C.binomial_thresholds(n, p)
        assert t == kr.binomial_thresholds(n, p) and len(t) == 32 and all(a <= b for a, b in zip(t, t[1:]))
        assert abs(t[0] / 2 ** 64 - (1 - p) ** n) < 1e-12
    with pytest.raises(ValueError):
        C.binomial_thresholds(48, 0.6)


def test_default_thresholds_values():
    g = C.GameConfig.from_options()
    assert g.spawn_cdf[0] == math.ceil((1 - Fraction(0.0005)) ** 48 * 2 ** 64)
    assert g.thr_keep == math.floor(Fraction(0.05) * 2 ** 32) + 1
    assert g.n_actions == 5 and g.food_mode == C.WAB_FOOD_INT and g.food_int == (40, 5, 40)


@pytest.mark.parametrize("power,maxb", [(100, 200), (12, 200), (30, 3), (8, 1), (2.5, 17)])
def test_bush_thresholds_bracket_the_reference_formula(power, maxb):
    thr = C.bush_thresholds(power, maxb).astype(np.uint64)
    assert len(thr) == maxb and np.all(np.diff(thr.astype(np.int64)) >= 0)
    ks = np.arange(1, maxb + 1)
    assert np.all(C.reference_bush_value(thr, power, maxb) >= ks)
    assert np.all(C.reference_bush_value(thr - np.uint64(1), power, maxb) < ks)
    # random words: table lookup == formula
    rng = np.random.default_rng(0)
    w = np.concatenate([rng.integers(0, 2 ** 32, 200000, dtype=np.uint64),
                        rng.integers(int(thr[0]), 2 ** 32, 200000, dtype=np.uint64)])
    assert np.array_equal(np.searchsorted(thr, w, side="right"), C.reference_bush_value(w, power, maxb).astype(np.int64))


def test_bush_thresholds_equal_exact_integer_arithmetic():
    """For integer powers the rounding boundary (k - 1/2) / max = U**power can be solved exactly:
    value >= k  <=>  w**100 * 400 >= (2k - 1) * 2**3200  (ties cannot occur: the right side is odd * 2**3200 / 400)."""
    thr = C.bush_thresholds(100, 200)
    for k in (1, 2, 3, 50, 100, 150, 199, 200):
        target = (2 * k - 1) * (1 << 3200)
        lo, hi = 0, 1 << 32
        while hi - lo > 1:
            mid = (lo + hi) // 2
            if mid ** 100 * 400 >= target:
                hi = mid
            else:
                lo = mid
        assert int(thr[k - 1]) == hi, k
    assert abs(int(thr[0]) / 2 ** 32 - 0.9418449208830277) < 1e-9     # SURVEY hard part 5


def test_integer_food_proof():
    assert C.prove_integer_food(8, 40, 1, 80) == (40, 5, 40)
    assert C.prove_integer_food(8, 40, None, 80) is None              # random start -> fp64
    assert C.prove_integer_food(7, 40, 1, 80) is None                 # 40/7 not an integer
    assert C.prove_integer_food(8, 40, 1, 400) is None                # SURVEY hard part 4: diverges past turn 121
    assert C.prove_integer_food(8, 40, 1, 121) == (40, 5, 40)
    assert C.prove_integer_food(4, 20, 1, 30) is not None


def test_validation_errors():
    with pytest.raises(ValueError):
        C.GameConfig.from_options({"width": 10})                      # wab_env.py:147-148
    with pytest.raises(ValueError):
        C.GameConfig.from_options({"chance_wolf_on_square": 2.5})
    with pytest.raises(ValueError):
        C.GameConfig.from_options({"max_turns": 10 ** 6})


def test_reward_table_matches_python_sums():
    g = C.GameConfig.from_options()
    assert list(g.reward_table64) == [0, 1, -1, -1, 0.1, 1.1, -0.9, -0.9]
    assert g.reward_table64[6] == 0 + 0.1 + -1
