"""Shared helpers for the parity tests (oracle side + comparison)."""

import numpy as np

from iscc_search_b200 import synth
from oracle import c_oracle, nphd_oracle


def make_store_arrays(n, seed, lengths=synth.STANDARD_LENGTHS):
    lens = synth.make_lengths(0, n, seed, lengths)
    codes = synth.make_codes(0, n, seed, lens)
    keys = synth.make_keys(0, n, seed)
    return keys, codes, lens


def oracle_topk(keys, codes, lens, queries, qlens, k, thr=None, keys_lo=None, use_c=True):
    """-> (keys uint64[Q,k] (or (hi,lo)), h, nbits, counts) from the CPU oracle."""
    if use_c:
        rows, h, nb, counts = c_oracle.topk(keys, keys_lo, codes, lens, queries, qlens, k, thr)
    else:
        res = nphd_oracle.topk(keys, keys_lo, codes, lens, queries, qlens, k, thr)
        q = len(qlens)
        rows = np.full((q, k), -1, dtype=np.int64)
        h = np.zeros((q, k), dtype=np.uint16)
        nb = np.zeros((q, k), dtype=np.uint16)
        counts = np.zeros(q, dtype=np.uint32)
        for i, (r, hh, nn) in enumerate(res):
            c = len(r)
            rows[i, :c], h[i, :c], nb[i, :c], counts[i] = r, hh, nn, c
    return rows, h, nb, counts


def assert_same_topk(got_keys, got_h, got_nb, got_counts, keys, rows, h, nb, counts, keys_lo=None):
    """Bit-exact comparison: neighbour ids, integer Hamming counts, nbits and order."""
    assert np.array_equal(got_counts.astype(np.int64), counts.astype(np.int64)), "result counts differ"
    for i in range(len(counts)):
        c = int(counts[i])
        if keys_lo is None:
            exp_keys = keys[rows[i, :c]]
            assert np.array_equal(got_keys[i, :c], exp_keys), f"query {i}: keys/order differ"
        else:
            exp = np.stack([keys[rows[i, :c]], keys_lo[rows[i, :c]]], axis=1)
            g = got_keys[i, :c].reshape(c, 16)
            ghi = np.array([int.from_bytes(bytes(r[:8]), "big") for r in g], dtype=np.uint64)
            glo = np.array([int.from_bytes(bytes(r[8:]), "big") for r in g], dtype=np.uint64)
            assert np.array_equal(ghi, exp[:, 0]) and np.array_equal(glo, exp[:, 1]), f"query {i}: keys/order differ"
        assert np.array_equal(got_h[i, :c], h[i, :c]), f"query {i}: hamming differs"
        assert np.array_equal(got_nb[i, :c], nb[i, :c]), f"query {i}: nbits differs"
