"""Philox4x32-10 known-answer tests (Random123 kat_vectors) on every implementation."""
import numpy as np
import pytest

from oracle import keyed_rng as kr
from oracle import wab_oracle
from tests import hostsim

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_philox_kat_numpy_oracle_hostsim(ctr, key, want):
    assert tuple(int(v) for v in kr.philox4x32_10(*ctr, *key)) == want
    assert tuple(int(v) for v in wab_oracle.philox(ctr, key)) == want
    assert tuple(int(v) for v in hostsim.philox(ctr, key)) == want


KAT2 = [   # Random123 kat_vectors, philox2x32 10
    ((0, 0), 0, (0xFF1DAE59, 0x6CD10DF2)),
    ((0xFFFFFFFF, 0xFFFFFFFF), 0xFFFFFFFF, (0x2C3F628B, 0xAB4FD7AD)),
    ((0x243F6A88, 0x85A308D3), 0x13198A2E, (0xDD7CE038, 0xF62A4C12)),
]


@pytest.mark.parametrize("ctr,key,want", KAT2)
def test_philox2x32_kat_numpy_oracle_hostsim(ctr, key, want):
    assert tuple(int(v) for v in kr.philox2x32_10(ctr[0], ctr[1], key)) == want
    assert tuple(int(v) for v in wab_oracle.philox2(ctr, key)) == want
    assert tuple(int(v) for v in hostsim.philox2(ctr, key)) == want


def test_three_implementations_agree_on_random_counters():
    rng = np.random.default_rng(7)
    ctrs = rng.integers(0, 2 ** 32, (64, 4), dtype=np.uint64)
    key = (0x12345678, 0x9ABCDEF0)
    ref = np.stack(kr.philox4x32_10(ctrs[:, 0], ctrs[:, 1], ctrs[:, 2], ctrs[:, 3], *key), axis=1)
    for i in range(64):
        assert np.array_equal(wab_oracle.philox(ctrs[i], key), ref[i])
        assert np.array_equal(hostsim.philox(ctrs[i], key), ref[i])


def test_ring_index_enumerates_the_48_cells():
    seen = {}
    for dx in range(-6, 7):
        for dy in range(-6, 7):
            if abs(dx) <= 5 and abs(dy) <= 5:
                continue
            seen[int(kr.ring_index(dx, dy, 11, 11, 1))] = (dx, dy)
    assert sorted(seen) == list(range(48))
    assert seen[0] == (-6, -6) and seen[12] == (-6, 6) and seen[13] == (-5, -6) and seen[14] == (-5, 6)
    assert seen[35] == (6, -6) and seen[47] == (6, 6)


@pytest.mark.gpu
def test_philox_on_device():
    import ctypes
    import torch
    from wab_gym_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(3)
    ctrs = rng.integers(0, 2 ** 32, (1000, 4), dtype=np.uint64)
    for c, k, w in KAT:
        ctrs[len(ctrs) - 1] = c
    key = (0xA4093822, 0x299F31D0)
    ctrs[0] = KAT[2][0]
    d_ctr = torch.from_numpy(ctrs.astype(np.uint32).view(np.int32)).cuda()
    d_out = torch.empty_like(d_ctr)
    _lib.check(L.wab_philox_device(ctypes.c_void_p(d_ctr.data_ptr()), key[0], key[1], len(ctrs),
                                   ctypes.c_void_p(d_out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    got = d_out.cpu().numpy().view(np.uint32)
    want = np.stack(kr.philox4x32_10(ctrs[:, 0], ctrs[:, 1], ctrs[:, 2], ctrs[:, 3], *key), axis=1)
    assert np.array_equal(got, want)
    assert tuple(int(v) for v in got[0]) == KAT[2][2]
