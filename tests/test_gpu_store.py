"""
GPU tests of the store semantics and of the two drop-in index classes, restating the reference's
dependency-boundary tests (tests/test_usearch_{search,add,remove,get,contains}.py) against the CUDA path,
plus exactness under heavy ties / candidate overflow and snapshot round trips.
"""

import json
import struct
from pathlib import Path

import numpy as np
import pytest

from iscc_search_b200 import BatchMatches, Matches, ShardedIndex128, ShardedNphdIndex, synth
from iscc_search_b200._lib import Store
from oracle.nphd_oracle import StoreOracle
from tests.helpers import assert_same_topk, make_store_arrays, oracle_topk

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def _hamming_store(vectors_by_key, nbytes):
    st = Store(key_bytes=8, max_bytes=nbytes, fixed_len=nbytes)
    keys = np.array(list(vectors_by_key.keys()), dtype=np.uint64)
    codes = np.zeros((len(keys), 32), dtype=np.uint8)
    for i, v in enumerate(vectors_by_key.values()):
        codes[i, :nbytes] = v
    st.add(keys, codes, np.full(len(keys), nbytes, dtype=np.uint8))
    return st


def test_reference_literal_kats_through_the_c_abi(cuda):
    kats = json.loads((GOLD / "usearch_kats.json").read_text())
    for case in kats["search"]:
        nbytes = len(case["query"])
        st = _hamming_store({k: v for k, v in case["stored"]}, nbytes)
        q = np.zeros((1, 32), dtype=np.uint8)
        q[0, :nbytes] = case["query"]
        keys, h, nb, cnt, _ = st.search(q, np.array([nbytes], dtype=np.uint8), case["count"])
        got = [(int(keys[0, j]), float(np.float32(h[0, j]))) for j in range(cnt[0])]
        assert got == [tuple(e) for e in case["expected"]], case["source"]
        st.close()


def test_nphd_index_add_get_contains_remove_semantics(cuda):
    idx = ShardedNphdIndex(max_dim=256)
    a, b = np.array([178, 204, 60, 240] * 2, dtype=np.uint8), np.array([100, 150, 200, 250] * 4, dtype=np.uint8)
    assert idx.size == 0 and idx.shard_count == 0
    assert idx.search(a, count=3).keys.tolist() == []                      # empty index -> empty Matches
    idx.add(1, a)
    idx.add(1, b)                                                          # duplicate key: silently skipped, first wins
    assert len(idx) == 1 and np.array_equal(idx.get(1), a)
    idx.add([2, 3], [bytes(b), bytes(range(32))])                          # mixed lengths in one batch
    assert np.array_equal(idx.get(3), np.arange(32, dtype=np.uint8))       # unpadded 256-bit body
    assert idx.get(999) is None and idx.get([1, 999])[1] is None
    assert 2 in idx and 999 not in idx and (2**63 - 1) not in idx and 0 not in idx
    assert idx.contains([1, 999, 2]).tolist() == [True, False, True]
    assert idx.contains([]).dtype == bool and len(idx.contains([])) == 0
    assert idx.remove(999) == 0 and idx.remove([1, 999, 2]) == 2 and len(idx) == 1
    assert idx.remove([]) == 0
    idx.add(1, b)                                                          # remove -> re-add works (update pattern)
    assert np.array_equal(idx.get(1), b) and idx.dirty > 0
    idx.add([2**64 - 1, 0], [bytes(a), bytes(a)])                          # keys span the full uint64 range
    assert (2**64 - 1) in idx and 0 in idx
    with pytest.raises(ValueError):
        idx.search(a, count=0)
    with pytest.raises(ValueError):
        idx.add(5, bytes(33))
    idx.close()


def test_nphd_search_return_shapes_and_float_conversion(cuda):
    idx = ShardedNphdIndex(max_dim=256)
    base = bytes([255, 170, 85, 0] * 4)                     # tests/conftest.py:209-228 `similar_units`
    one_bit = bytes([254]) + base[1:]
    inverted = bytes(x ^ 0xFF for x in base)
    idx.add([10, 11, 12], [base, one_bit, inverted])
    m = idx.search(np.frombuffer(base, dtype=np.uint8), count=10)
    assert isinstance(m, Matches) and m.keys.dtype == np.uint64 and m.distances.dtype == np.float32
    assert m.keys.tolist() == [10, 11, 12] and len(m) == 3   # count > size -> all rows
    assert m.distances.tolist() == [0.0, float(np.float32(1) / np.float32(128)), 1.0]
    assert max(0.0, 1.0 - float(m.distances[0])) == 1.0      # identical unit => score 1.0 (index.py:2041-2043)
    assert m[1].key == 11 and m.to_list()[0] == (10, 0.0)
    bm = idx.search(np.stack([np.frombuffer(base, np.uint8), np.frombuffer(inverted, np.uint8)]), count=2)
    assert isinstance(bm, BatchMatches) and bm.keys.shape == (2, 2) and bm.counts.tolist() == [2, 2]
    assert bm[1].keys.tolist() == [12, 11] and len(bm.to_list()) == 4
    # 64-bit query against the 128-bit rows: scored on the first 8 bytes only
    m8 = idx.search(base[:8], count=1)
    assert m8.keys.tolist() == [10] and m8.nbits.tolist() == [64]
    idx.close()


def _ck(asset, offset, size):
    return struct.pack(">Q", asset) + struct.pack("!II", offset, size)


def test_index128_composite_keys_threshold_and_equality_modes(cuda):
    idx = ShardedIndex128(ndim=64)
    rng = np.random.default_rng(1)
    base = rng.integers(0, 256, size=8, dtype=np.uint8)
    flip48 = base.copy()
    bits = np.unpackbits(flip48)
    bits[:48] ^= 1
    flip48 = np.packbits(bits)                                # 48/64 flipped => score 0.25 (simprint_approx.py:308-330)
    keys = [_ck(7, 0, 100), _ck(7, 100, 50), _ck(3, 0, 10), _ck(9, 5, 5)]
    idx.add(keys, [base, flip48, base, base])
    idx.add([keys[0]], [flip48])                              # dup composite key: first wins
    assert len(idx) == 4 and keys[1] in idx and _ck(1, 1, 1) not in idx
    assert np.array_equal(idx.get(keys[0]), base)
    m = idx.search(base, count=10)
    assert isinstance(m, Matches)
    # ties at h=0 ordered by the 16-byte key bytewise: asset 3 < asset 7 < asset 9
    assert [bytes(k) for k in m.keys] == [keys[2], keys[0], keys[3], keys[1]]
    assert m.distances.tolist() == [0.0, 0.0, 0.0, 48.0]
    assert 1.0 - float(m.distances[3]) / 64 == 0.25
    m = idx.search(base, count=10, threshold_bits=16, with_vectors=True)
    assert len(m) == 3 and np.array_equal(m.vectors[0], base)
    assert isinstance(idx.search(base.reshape(1, -1), count=1), Matches)      # (1, n) batch -> bare Matches too
    m = idx.search(np.stack([base, flip48]), count=2, threshold_bits=0)     # equality join, capped at 2
    assert m.counts.tolist() == [2, 1] and [bytes(k) for k in m[0].keys] == [keys[2], keys[0]]
    assert idx.remove([keys[2], _ck(1, 1, 1)]) == 1 and keys[2] not in idx
    assert [bytes(k) for k in idx.search(base, count=1).keys] == [keys[0]]
    with pytest.raises(ValueError):
        idx.add([_ck(1, 0, 0)], [np.zeros(16, dtype=np.uint8)])   # wrong ndim
    idx.close()


def test_random_mutations_track_the_dict_model(cuda):
    rng = np.random.default_rng(42)
    idx, model = ShardedNphdIndex(max_dim=256), StoreOracle()
    pool = rng.integers(0, 2**64, size=600, dtype=np.uint64)
    for step in range(40):
        ks = rng.choice(pool, size=int(rng.integers(1, 60)), replace=True)
        if rng.random() < 0.6:
            vs = [bytes(rng.integers(0, 256, size=int(rng.choice([8, 16, 24, 32])), dtype=np.uint8)) for _ in ks]
            idx.add(ks, vs)
            model.add([int(k) for k in ks], vs)
        else:
            assert idx.remove(ks) == model.remove([int(k) for k in ks])
        assert len(idx) == len(model)
    for key in pool[:100]:
        got, exp = idx.get(int(key)), model.get(int(key))
        assert (got is None) == (exp is None) and (exp is None or bytes(got) == exp)
    # search after the mutations equals the oracle over the surviving rows
    keys, _, codes, lens = model.arrays()
    queries, qlens = synth.make_queries(12, 0, 5)
    gk, gh, gn, gc, _ = idx._store.search(queries, qlens, 20)
    rows, h, nb, cnt = oracle_topk(keys, codes, lens, queries, qlens, 20, use_c=False)
    assert_same_topk(gk, gh, gn, gc, keys, rows, h, nb, cnt)
    idx.close()


def test_mass_duplicates_overflow_fallback_and_radix_tie_select(cuda):
    # 300K rows share ONE code: every row ties at h=0, the candidate buffer overflows, the exact re-scan
    # collects all ties and the key radix-select picks the k smallest keys.
    n, k = 300_000, 10
    keys = synth.make_keys(0, n, 77)
    codes = np.zeros((n, 32), dtype=np.uint8)
    codes[:, :8] = np.arange(8, dtype=np.uint8) + 1
    codes[n // 2:, 0] ^= 0x80                                 # second half: 1 bit away
    lens = np.full(n, 8, dtype=np.uint8)
    st = Store(key_bytes=8, max_bytes=32)
    st.add(keys, codes, lens)
    q = codes[:1].copy()
    ql = np.array([8], dtype=np.uint8)
    gk, gh, gn, gc, _ = st.search(q, ql, k)
    assert st.stats()["fallback_queries"] == 1
    rows, h, nb, cnt = oracle_topk(keys, codes, lens, q, ql, k)
    assert_same_topk(gk, gh, gn, gc, keys, rows, h, nb, cnt)
    # k larger than the first tie group: crosses into the h=1 group, still exact
    k2 = 2000
    q2 = np.concatenate([q, codes[n - 1:]])
    gk, gh, gn, gc, _ = st.search(q2, np.array([8, 8], dtype=np.uint8), k2)
    rows, h, nb, cnt = oracle_topk(keys, codes, lens, q2, np.array([8, 8], dtype=np.uint8), k2)
    assert_same_topk(gk, gh, gn, gc, keys, rows, h, nb, cnt)
    st.close()


def test_ties_larger_than_sort_capacity_without_overflow(cuda):
    # 6000 ties at the cut-off rank (more than the shared-memory sort holds for small k) but within the buffer
    n, k = 6000, 7
    keys = synth.make_keys(0, n, 5)
    codes = np.zeros((n, 32), dtype=np.uint8)
    codes[:, :16] = 0x5A
    lens = np.full(n, 16, dtype=np.uint8)
    st = Store(key_bytes=8, max_bytes=32)
    st.add(keys, codes, lens)
    q = codes[:1].copy()
    ql = np.array([16], dtype=np.uint8)
    gk, gh, gn, gc, _ = st.search(q, ql, k)
    rows, h, nb, cnt = oracle_topk(keys, codes, lens, q, ql, k)
    assert_same_topk(gk, gh, gn, gc, keys, rows, h, nb, cnt)
    st.close()


def test_snapshot_round_trip_and_reload(cuda, tmp_path):
    n = 30_000
    lens = synth.make_lengths(0, n, 8)
    codes = synth.make_codes(0, n, 8, lens)
    keys = synth.make_keys(0, n, 8)
    idx = ShardedNphdIndex(max_dim=256, path=tmp_path / "META_NONE_V0")
    idx.add(keys, [bytes(codes[i, : lens[i]]) for i in range(n)])
    idx.remove(keys[:100])
    assert idx.dirty == n + 100
    idx.save()
    assert idx.dirty == 0
    queries, qlens = synth.make_queries(8, n, 9, 8)
    before = idx._store.search(queries, qlens, 25)
    idx.close()
    idx2 = ShardedNphdIndex(max_dim=256, path=tmp_path / "META_NONE_V0")       # auto-loads the snapshot
    assert idx2.size == n - 100 and int(keys[5]) not in idx2 and int(keys[500]) in idx2
    after = idx2._store.search(queries, qlens, 25)
    for a, b in zip(before[:4], after[:4]):
        assert np.array_equal(a, b)
    assert np.array_equal(idx2.get(int(keys[500])), codes[500, : lens[500]])
    idx2.reset()
    assert idx2.size == 0
    idx2.close()


def test_max_k_limit_is_reported(cuda):
    st = Store(key_bytes=16, max_bytes=8, fixed_len=8)
    assert st.max_k() >= 4096
    with pytest.raises(ValueError, match="exceeds the supported maximum"):
        st.search(np.zeros((1, 32), np.uint8), np.array([8], np.uint8), st.max_k() + 1)
    st.close()


def test_bulk_remove_resolves_move_chains(cuda):
    """One remove call of half the store (tail rows included, so moved rows are removed again later in the batch): the
    net parallel row moves must leave exactly the surviving rows, each with its own code and key."""
    n = 300_000
    keys, codes, lens = make_store_arrays(n, 77)
    st = Store(key_bytes=8, max_bytes=32)
    st.add(keys, codes, lens)
    rng = np.random.default_rng(5)
    gone = rng.permutation(n)[: n // 2]
    gone = np.concatenate([gone, np.arange(n - 5000, n)])          # plus the whole tail: long chains
    gone = np.unique(gone)
    rng.shuffle(gone)
    removed, cnt = st.remove(np.ascontiguousarray(keys[gone]), len(gone))
    assert cnt == len(gone) and removed.all() and st.size() == n - len(gone)
    keep = np.ones(n, dtype=bool)
    keep[gone] = False
    got_codes, got_lens = st.get(np.ascontiguousarray(keys[keep]), int(keep.sum()))
    assert np.array_equal(got_lens, lens[keep]) and np.array_equal(got_codes, codes[keep])
    assert not st.contains(np.ascontiguousarray(keys[gone][:1000]), 1000).any()
    queries, qlens = synth.make_queries(32, n, 78, 77)
    gk, gh, gn, gc, _ = st.search(queries, qlens, 50)
    rows, h, nb, cnt2 = oracle_topk(keys[keep], codes[keep], lens[keep], queries, qlens, 50)
    assert_same_topk(gk, gh, gn, gc, keys[keep], rows, h, nb, cnt2)
    # and the store keeps working: re-add a part of what was removed
    back = gone[:10_000]
    assert st.add(np.ascontiguousarray(keys[back]), np.ascontiguousarray(codes[back]), np.ascontiguousarray(lens[back])).all()
    keep[back] = True
    gk, gh, gn, gc, _ = st.search(queries, qlens, 50)
    rows, h, nb, cnt2 = oracle_topk(keys[keep], codes[keep], lens[keep], queries, qlens, 50)
    assert_same_topk(gk, gh, gn, gc, keys[keep], rows, h, nb, cnt2)
    st.close()


def test_release_scratch_frees_working_memory_and_searches_keep_working(cuda):
    n = 200_000
    keys, codes, lens = make_store_arrays(n, 91)
    st = Store(key_bytes=8, max_bytes=32)
    st.add(keys, codes, lens)
    queries, qlens = synth.make_queries(3000, n, 92, 91)
    first = st.search(queries, qlens, 100)
    freed = st.release_scratch()
    assert freed > 100 * 2**20          # candidate lists of a 3000-query batch
    assert st.release_scratch() == 0
    again = st.search(queries, qlens, 100)
    for a, b in zip(first[:4], again[:4]):
        assert np.array_equal(a, b)
    one = st.search(queries[:1], qlens[:1], 10)        # small-batch path after a release
    rows, h, nb, cnt = oracle_topk(keys, codes, lens, queries[:1], qlens[:1], 10)
    assert_same_topk(one[0], one[1], one[2], one[3], keys, rows, h, nb, cnt)
    st.close()
