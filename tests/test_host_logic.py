"""CPU tests (no GPU) of the library's host logic through its self-test entry points: rank tables and the key map."""

import ctypes
from fractions import Fraction

import numpy as np
import pytest

from iscc_search_b200 import _lib


def _tables(mask, stride=8192):
    rank = np.zeros((33, 257), dtype=np.uint16)
    hmax = np.zeros((33, stride), dtype=np.uint16)
    R = ctypes.c_uint32()
    _lib.check(_lib.lib().isx_selftest_rank_table(mask, _lib.ptr(rank), _lib.ptr(hmax), stride, ctypes.byref(R)))
    return rank, hmax, int(R.value)


@pytest.mark.parametrize("lengths", [(8, 16, 24, 32), (8,), (4, 5, 12, 13, 20, 27, 32), tuple(range(1, 33))])
def test_rank_table_is_the_dense_rank_of_the_exact_rational(lengths):
    mask = 0
    for m in lengths:
        mask |= 1 << (m - 1)
    rank, hmax, R = _tables(mask)
    fracs = sorted({Fraction(h, 8 * m) for m in lengths for h in range(8 * m + 1)})
    assert R == len(fracs)
    if lengths == (8, 16, 24, 32):
        assert R == 385  # 257 + 193 - 65 distinct values
    index = {f: i for i, f in enumerate(fracs)}
    for m in range(1, 33):
        if m in lengths:
            got = rank[m, : 8 * m + 1].astype(int)
            assert [int(x) for x in got] == [index[Fraction(h, 8 * m)] for h in range(8 * m + 1)]
            assert (rank[m, 8 * m + 1:] == 0xFFFF).all()
            # hmax[m][r] = largest h whose rank is <= r: monotone, consistent with rank, and 0/1 -> h=0 only
            for r in (0, 1, R // 3, R - 1):
                h = int(hmax[m, r])
                assert rank[m, h] <= r and (h == 8 * m or rank[m, h + 1] > r)
            assert (np.diff(hmax[m, :R].astype(int)) >= 0).all() and hmax[m, 0] == 0 and hmax[m, R - 1] == 8 * m
        else:
            assert (rank[m] == 0xFFFF).all()


def test_equal_fractions_of_different_lengths_share_a_rank():
    rank, _, _ = _tables((1 << 7) | (1 << 15) | (1 << 31))
    assert rank[8, 16] == rank[16, 32] == rank[32, 64]      # 16/64 = 32/128 = 64/256
    assert rank[8, 1] > rank[32, 3] and rank[8, 1] < rank[32, 5]  # 1/64 = 4/256 sits between 3/256 and 5/256
    assert rank[8, 1] == rank[32, 4]


def test_keymap_randomized_against_reference():
    L = _lib.lib()
    for seed, ops, space in ((1, 200_000, 64), (2, 300_000, 1500), (3, 400_000, 50_000)):
        rc = L.isx_selftest_keymap(ops, seed, space)
        assert rc == 0, L.isx_last_error()


def test_device_arithmetic_helpers_on_the_host():
    # the same template code the scan kernels use (carry-save distance, 2- and 3-word OR-fold bounds), compiled for the host
    L = _lib.lib()
    for seed in (1, 2, 3):
        rc = L.isx_selftest_distance(20_000, seed)
        assert rc == 0, L.isx_last_error()


def test_instance_index_digest_collisions_never_lose_a_row(cpu_stores, monkeypatch):
    """Two different bodies of one asset whose digests collide occupy consecutive slots; removal keeps the chain gap-free."""
    import hashlib

    from iscc_search_b200.instance import InstancePrefixIndex

    real = hashlib.blake2b

    class Colliding:
        def __init__(self, data, digest_size):
            self._d = bytes([7]) * digest_size   # every body gets the same digest

        def digest(self):
            return self._d

    idx = InstancePrefixIndex(store=cpu_stores(0, 16, 32, 0))
    monkeypatch.setattr("iscc_search_b200.instance.hashlib.blake2b", Colliding)
    a, b, c = bytes(range(8)), bytes(range(8, 16)), bytes(range(16, 24))
    idx.add_many([(5, a), (5, b), (5, c), (5, a)])
    assert len(idx) == 3
    assert idx.search(a) == {5: 1.0} and idx.search(b) == {5: 1.0} and idx.search(c) == {5: 1.0}
    assert idx.remove(5, a) == 1 and len(idx) == 2          # head of the chain goes, the others move up
    idx.add(5, b)                                            # still found: no duplicate row
    assert len(idx) == 2
    assert idx.remove(5, c) == 1 and idx.remove(5, c) == 0
    assert idx.search(b) == {5: 1.0} and idx.search(c) == {} and idx.search(a) == {}
    monkeypatch.setattr("iscc_search_b200.instance.hashlib.blake2b", real)
