"""
GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Bar: bit-exact neighbour ids, integer Hamming counts, nbits, and order under (h/n, key) - integer
work, so no tolerance.
"""

import numpy as np
import pytest

from iscc_search_b200 import synth
from iscc_search_b200._lib import Store
from tests.helpers import assert_same_topk, make_store_arrays, oracle_topk

pytestmark = pytest.mark.gpu


def _run(n, q, k, seed, lengths=synth.STANDARD_LENGTHS, qlengths=synth.STANDARD_LENGTHS, use_c=True, thr=None):
    keys, codes, lens = make_store_arrays(n, seed, lengths)
    queries, qlens = synth.make_queries(q, n, seed + 1, seed, qlengths, lengths)
    st = Store(key_bytes=8, max_bytes=32)
    try:
        added = st.add(keys, codes, lens)
        assert added.all() and st.size() == n
        gk, gh, gn, gc, _ = st.search(queries, qlens, k, thr)
        rows, h, nb, counts = oracle_topk(keys, codes, lens, queries, qlens, k, thr, use_c=use_c)
        assert_same_topk(gk, gh, gn, gc, keys, rows, h, nb, counts)
        return st.stats()
    finally:
        st.close()


def test_small_mixed_vs_numpy_oracle(cuda):
    _run(n=3000, q=24, k=10, seed=11, use_c=False)


def test_tiny_store_fewer_rows_than_k(cuda):
    _run(n=7, q=5, k=10, seed=3, use_c=False)


def test_single_row(cuda):
    _run(n=1, q=3, k=1, seed=5, use_c=False)


@pytest.mark.parametrize("length", [8, 16, 24, 32])
def test_uniform_length_buckets(cuda, length):
    _run(n=50_000, q=40, k=25, seed=20 + length, lengths=(length,), qlengths=(length,))


def test_mixed_lengths_all_query_lengths(cuda):
    st = _run(n=200_000, q=256, k=100, seed=7)
    assert st["fallback_queries"] == 0


def test_nonstandard_lengths_byte_granular(cuda):
    # 32-bit-word and odd byte lengths: generic path (last-word mask)
    _run(n=20_000, q=32, k=16, seed=31, lengths=(4, 5, 12, 13, 20, 27, 32), qlengths=(3, 4, 9, 16, 23, 32))


def test_config1_shape_256bit_k10(cuda):
    # BASELINE config 1: 100K x 256-bit, k=10 (1K queries at full size in bench; 200 here)
    _run(n=100_000, q=200, k=10, seed=101, lengths=(32,), qlengths=(32,))


def test_config2_shape_64bit_heavy_ties_k100(cuda):
    # BASELINE config 2 shape (64-bit, k=100) at 1M rows: 65 distinct distances -> key tie-break dominates
    _run(n=1_000_000, q=64, k=100, seed=202, lengths=(8,), qlengths=(8,))


def test_short_query_against_long_rows(cuda):
    # the mainstream request: 64-bit query vs 256-bit rows scored on the first 8 bytes only
    _run(n=120_000, q=64, k=50, seed=303, lengths=(32,), qlengths=(8,))


def test_large_k(cuda):
    _run(n=60_000, q=8, k=2000, seed=404)


def test_threshold_mode(cuda):
    # keep only rows with h/n <= 16/64 (simprint threshold 0.75)
    _run(n=80_000, q=48, k=100, seed=505, lengths=(8,), qlengths=(8,), thr=(16, 64))
    _run(n=30_000, q=16, k=50, seed=506, thr=(1, 4))


def test_many_queries_multiple_tiles(cuda):
    # more queries than one shared-memory tile (2048)
    _run(n=30_000, q=4500, k=5, seed=606, lengths=(8, 32), qlengths=(8, 32))


def test_empty_query_batch_and_empty_store(cuda):
    st = Store(key_bytes=8, max_bytes=32)
    try:
        gk, gh, gn, gc, _ = st.search(np.zeros((0, 32), np.uint8), np.zeros(0, np.uint8), 5)
        assert gk.shape == (0, 5) and gc.shape == (0,)
        q, ql = synth.make_queries(3, 0, 1)
        gk, gh, gn, gc, _ = st.search(q, ql, 5)          # empty store: zero results per query
        assert (gc == 0).all()
        with pytest.raises(ValueError):
            st.search(q, np.array([0, 8, 8], dtype=np.uint8), 5)   # zero-length query
        with pytest.raises(ValueError):
            st.search(q, np.array([33, 8, 8], dtype=np.uint8), 5)  # longer than max_bytes
    finally:
        st.close()


def test_k_at_the_supported_maximum_with_128bit_keys_and_heavy_ties(cuda):
    # k = isx_max_k (more than the whole store), 64-bit codes with only 65 distinct distances: the cut-off tie group is
    # far larger than the shared-memory sort capacity -> radix select on the 16-byte keys
    n = 60_000
    rng = np.random.default_rng(12)
    codes = np.zeros((n, 32), dtype=np.uint8)
    codes[:, :8] = rng.integers(0, 256, size=(n, 8), dtype=np.uint8)
    lens = np.full(n, 8, dtype=np.uint8)
    hi = rng.integers(0, 2**20, size=n, dtype=np.uint64)      # many equal high halves: order decided by the low half
    lo = rng.permutation(n).astype(np.uint64)
    keys = np.zeros((n, 16), dtype=np.uint8)
    keys[:, :8] = hi.astype(">u8").view(np.uint8).reshape(n, 8)
    keys[:, 8:] = lo.astype(">u8").view(np.uint8).reshape(n, 8)
    st = Store(key_bytes=16, max_bytes=8, fixed_len=8)
    try:
        st.add(keys, codes, lens)
        k = st.max_k()
        q, ql = codes[:3].copy(), lens[:3].copy()
        gk, gh, gn, gc, _ = st.search(q, ql, k)
        rows, h, nb, cnt = oracle_topk(hi, codes, lens, q, ql, k, keys_lo=lo)
        assert_same_topk(gk, gh, gn, gc, hi, rows, h, nb, cnt, keys_lo=lo)
    finally:
        st.close()


def test_byte_granular_lengths_with_threshold_and_codes_out(cuda):
    n = 15_000
    keys, codes, lens = make_store_arrays(n, 77, (3, 7, 8, 11, 32))
    queries, qlens = synth.make_queries(20, n, 78, 77, (3, 8, 11, 30), (3, 7, 8, 11, 32))
    st = Store(key_bytes=8, max_bytes=32)
    try:
        st.add(keys, codes, lens)
        gk, gh, gn, gc, gcodes = st.search(queries, qlens, 30, (1, 8), True)
        rows, h, nb, cnt = oracle_topk(keys, codes, lens, queries, qlens, 30, (1, 8))
        assert_same_topk(gk, gh, gn, gc, keys, rows, h, nb, cnt)
        for i in range(len(qlens)):
            assert np.array_equal(gcodes[i, : cnt[i]], codes[rows[i, : cnt[i]]])     # matched stored codes, zero padded
    finally:
        st.close()


def test_k_beyond_the_shared_memory_sort_capacity(cuda):
    # k = 10000 > 8192 (uint64 keys): winners are sorted in global scratch; mixed lengths, ties cut by key
    st = _run(n=300_000, q=6, k=10_000, seed=909)
    assert st["fallback_queries"] == 0
    # 64-bit codes only: a handful of distinct distances, the tie group at the cut is far larger than any buffer slack
    _run(n=200_000, q=4, k=9_000, seed=910, lengths=(8,), qlengths=(8,))


def test_large_k_with_128bit_keys_and_threshold(cuda):
    n, k = 150_000, 6_000
    rng = np.random.default_rng(21)
    codes = np.zeros((n, 32), dtype=np.uint8)
    codes[:, :8] = rng.integers(0, 256, size=(n, 8), dtype=np.uint8)
    lens = np.full(n, 8, dtype=np.uint8)
    hi = rng.integers(0, 2**16, size=n, dtype=np.uint64)
    lo = rng.permutation(n).astype(np.uint64)
    keys = np.zeros((n, 16), dtype=np.uint8)
    keys[:, :8] = hi.astype(">u8").view(np.uint8).reshape(n, 8)
    keys[:, 8:] = lo.astype(">u8").view(np.uint8).reshape(n, 8)
    st = Store(key_bytes=16, max_bytes=8, fixed_len=8)
    try:
        st.add(keys, codes, lens)
        q, ql = codes[:3].copy(), lens[:3].copy()
        for thr in (None, (26, 64)):
            gk, gh, gn, gc, _ = st.search(q, ql, k, thr)
            rows, h, nb, cnt = oracle_topk(hi, codes, lens, q, ql, k, thr, keys_lo=lo)
            assert_same_topk(gk, gh, gn, gc, hi, rows, h, nb, cnt, keys_lo=lo)
    finally:
        st.close()
