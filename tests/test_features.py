"""Device-side PragmaticObsWrapper features (wab_gym_b200/csrc/wab_features.cuh) against the
reference's own known-answer tests (wab_env_test.py:9-169, restated as data here so they travel) and,
where the reference is present, against its wrapper on random grids."""
import numpy as np
import pytest

from tests import hostsim


def unpack(f):
    f = [int(v) for v in f]
    return (f[0:4], f[4:8], f[8:12], f[12:16], f[16:20], f[20:24], f[24], f[25], f[26], f[27])


def grid(points):
    g = np.zeros((11, 11))
    for p in points:
        g[p] = 1
    return g


# (wolves, bushes, food, role, status) -> (nearest_wolf, second_wolf, n_wolves, nearest_bush, second_bush, n_bushes, standing)
REFERENCE_KATS = [
    # test_TwoEquidistantBushes, wab_env_test.py:9-65
    (([(5, 5), (6, 6), (4, 4)], [(6, 3), (7, 4), (8, 6), (6, 10)], 40, 0, 0),
     ([0, 0, 0, 0], [0, 10, 10, 0], [1, 1, 1, 1], [0, 0, 9, 10], [0, 0, 10, 9], [0, 2, 4, 2], 0)),
    # test_standing_on_bush, wab_env_test.py:67-111
    (([], [(5, 5)], 40, 0, 0),
     ([0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0], 1)),
    # test_numerous_bushes_and_wolves_with_blindspots, wab_env_test.py:113-169 (grids after the lookout mask)
    (([(2, j) for j in range(1, 10)] + [(i, 6) for i in range(11)],
      [(1, j) for j in range(2, 9)] + [(9, j) for j in range(2, 9)], 40, 0, 0),
     ([0, 10, 0, 0], [0, 10, 10, 0], [10, 10, 5, 4], [0, 0, 7, 0], [7, 0, 0, 0], [7, 6, 7, 6], 0)),
]


@pytest.mark.parametrize("inp,want", REFERENCE_KATS)
def test_reference_known_answers(inp, want):
    wolves, bushes, food, role, status = inp
    got = unpack(hostsim.features(grid(wolves), grid(bushes), food, role, status))
    assert got[:7] == want and got[7:] == (food, role, status)


def test_standing_on_bush_and_empty_planes():
    got = unpack(hostsim.features(grid([]), grid([(5, 5)]), 12, 1, 0))
    assert got[0] == [0, 0, 0, 0] and got[1] == [0, 0, 0, 0] and got[2] == [0, 0, 0, 0]
    assert got[3] == [0, 0, 0, 0] and got[6] == 1 and got[5] == [0, 0, 0, 0]
    got = unpack(hostsim.features(grid([(0, 0)] + [(i, 10) for i in range(11)]), grid([]), 0, 0, 2))
    assert got[2] == [6, 10, 5, 1]          # counts clip at 10 (wab_env.py:734)


def test_against_reference_wrapper_on_random_grids():
    from oracle import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference sources not present")
    mod = ref_shim.load_reference()
    wrapper = mod.PragmaticObsWrapper(ref_shim.make_env())
    rng = np.random.default_rng(0)
    for trial in range(400):
        density = rng.choice([0.0, 0.02, 0.06, 0.3, 0.9])
        wolves = (rng.random((11, 11)) < density * rng.random()).astype(float)
        bushes = (rng.random((11, 11)) < density).astype(float)
        food, role, status = int(rng.integers(0, 41)), int(rng.integers(0, 2)), int(rng.integers(0, 3))
        ref = wrapper.observation((wolves, bushes, np.zeros((11, 11)), food, role, status, np.zeros((11, 11))))
        want = (list(ref[0]), list(ref[1]), [int(v) for v in ref[2]], list(ref[3]), list(ref[4]), [int(v) for v in ref[5]],
                ref[6], ref[7], ref[8], ref[9])
        assert unpack(hostsim.features(wolves, bushes, food, role, status)) == want, trial
