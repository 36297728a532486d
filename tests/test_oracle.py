"""
CPU tests (no GPU): the oracle restatements against the reference's literal known-answer vectors
(tests/golden/usearch_kats.json) and against each other (C vs numpy) on seeded inputs.
"""

import json
from pathlib import Path

import numpy as np
import pytest

from iscc_search_b200 import synth
from oracle import c_oracle, nphd_oracle
from oracle.nphd_oracle import StoreOracle

GOLD = Path(__file__).parent / "golden"


def test_usearch_search_kats_numpy_oracle():
    kats = json.loads((GOLD / "usearch_kats.json").read_text())
    for case in kats["search"]:
        st = StoreOracle()
        st.add([k for k, _ in case["stored"]], [bytes(v) for _, v in case["stored"]])
        got = st.search(bytes(case["query"]), case["count"])
        # Hamming metric: distance is the raw bit count as float32 (tests/test_usearch_search.py:122-167)
        assert [(k, float(np.float32(h))) for k, h, _n in got] == [tuple(e) for e in case["expected"]], case["source"]


def test_usearch_search_kats_c_oracle():
    kats = json.loads((GOLD / "usearch_kats.json").read_text())
    for case in kats["search"]:
        keys = np.array([k for k, _ in case["stored"]], dtype=np.uint64)
        codes, lens = nphd_oracle.pad_codes([bytes(v) for _, v in case["stored"]])
        q, ql = nphd_oracle.pad_codes([bytes(case["query"])])
        rows, h, nb, cnt = c_oracle.topk(keys, None, codes, lens, q, ql, case["count"])
        got = [(int(keys[rows[0, j]]), float(np.float32(h[0, j]))) for j in range(cnt[0])]
        assert got == [tuple(e) for e in case["expected"]], case["source"]


def test_score_conversions_match_reference_literals():
    kats = json.loads((GOLD / "usearch_kats.json").read_text())
    sp = kats["simprint_threshold"]
    assert nphd_oracle.simprint_score(sp["flipped_bits"], sp["ndim"]) == sp["expected_score"]
    assert nphd_oracle.unit_score(0, 64) == kats["unit_score_identity"]["expected"]
    assert nphd_oracle.unit_score(256, 256) == 0.0  # clamp at zero (index.py:2043)


def test_count_zero_raises_value_error():
    # tests/test_usearch_search.py:678-685
    st = StoreOracle()
    st.add([1], [b"\x01\x02\x03\x04"])
    with pytest.raises(ValueError):
        st.search(b"\x01\x02\x03\x04", 0)
    with pytest.raises(ValueError):
        c_oracle.topk(np.array([1], dtype=np.uint64), None, np.zeros((1, 32), np.uint8), np.array([4], np.uint8),
                      np.zeros((1, 32), np.uint8), np.array([4], np.uint8), 0)


def test_store_semantics_first_wins_remove_readd():
    st = StoreOracle()
    assert st.add([1, 1, 2], [b"\xb2\xcc\x3c\xf0", b"\x64\x96\xc8\xfa", b"\x01\x02\x03\x04"]) == [True, False, True]
    assert st.get(1) == b"\xb2\xcc\x3c\xf0" and len(st) == 2          # tests/test_usearch_add.py:53-63
    assert st.remove([1, 999]) == 1 and 1 not in st                    # tests/test_usearch_remove.py:118-141
    assert st.add([1], [b"\x64\x96\xc8\xfa"]) == [True] and st.get(1) == b"\x64\x96\xc8\xfa"  # :226-243
    assert (2**63 - 1) not in st and 0 not in st                       # tests/test_usearch_contains.py:214-235
    assert st.get(12345) is None                                       # tests/test_usearch_get.py:47-56


def test_cross_length_prefix_formula():
    # docs/explanation/similarity-search.md:24-32: compare the common prefix, divide by its bit length
    a = bytes([0xFF] * 8 + [0x00] * 24)   # 256-bit stored
    q = bytes([0xFF] * 7 + [0xFE])        # 64-bit query, 1 bit off inside the prefix
    st = StoreOracle()
    st.add([7], [a])
    (key, h, n), = st.search(q, 1)
    assert (key, h, n) == (7, 1, 64)
    assert nphd_oracle.nphd_distance_f32(np.array([h]), np.array([n]))[0] == np.float32(1) / np.float32(64)


@pytest.mark.parametrize("seed,lengths", [(1, synth.STANDARD_LENGTHS), (2, (8,)), (3, (4, 5, 12, 13, 20, 27, 32))])
def test_c_oracle_equals_numpy_oracle(seed, lengths):
    n, q, k = 4000, 24, 17
    lens = synth.make_lengths(0, n, seed, lengths)
    codes = synth.make_codes(0, n, seed, lens)
    keys = synth.make_keys(0, n, seed)
    queries, qlens = synth.make_queries(q, n, seed + 1, seed, lengths, lengths)
    ref = nphd_oracle.topk(keys, None, codes, lens, queries, qlens, k)
    rows, h, nb, cnt = c_oracle.topk(keys, None, codes, lens, queries, qlens, k)
    for i, (r, hh, nn) in enumerate(ref):
        assert cnt[i] == len(r)
        assert np.array_equal(rows[i, : cnt[i]], r) and np.array_equal(h[i, : cnt[i]], hh) and np.array_equal(nb[i, : cnt[i]], nn)


def test_oracles_threshold_and_128bit_keys():
    n, q, k = 3000, 10, 50
    rng = np.random.default_rng(5)
    codes = np.zeros((n, 32), dtype=np.uint8)
    codes[:, :8] = rng.integers(0, 256, size=(n, 8), dtype=np.uint8)
    codes[: n // 3, :8] = codes[0, :8]  # heavy duplicates: ties broken by the 128-bit key
    lens = np.full(n, 8, dtype=np.uint8)
    hi = rng.integers(0, 2**63, size=n, dtype=np.uint64) // np.uint64(1 << 40)  # few distinct asset ids
    lo = rng.permutation(n).astype(np.uint64)
    queries = codes[rng.integers(0, n, size=q)].copy()
    qlens = np.full(q, 8, dtype=np.uint8)
    for thr in (None, (16, 64), (0, 64)):
        ref = nphd_oracle.topk(hi, lo, codes, lens, queries, qlens, k, thr)
        rows, h, nb, cnt = c_oracle.topk(hi, lo, codes, lens, queries, qlens, k, thr)
        for i, (r, hh, nn) in enumerate(ref):
            assert cnt[i] == len(r) and np.array_equal(rows[i, : cnt[i]], r) and np.array_equal(h[i, : cnt[i]], hh)
            if thr is not None:
                assert np.all(hh.astype(np.int64) * thr[1] <= thr[0] * nn.astype(np.int64))


def test_synth_is_deterministic_and_row_addressable():
    a = synth.make_codes(1000, 50, 9)
    b = synth.make_codes(0, 2000, 9)[1000:1050]
    assert np.array_equal(a, b)
    keys = synth.make_keys(0, 100_000, 9)
    assert len(np.unique(keys)) == len(keys)
    lens = synth.make_lengths(0, 40_000, 9)
    assert set(np.unique(lens)) == set(synth.STANDARD_LENGTHS)
    assert np.all(synth.make_codes(0, 100, 9, lens[:100])[np.arange(32)[None, :] >= lens[:100, None]] == 0)


def test_c_synth_generator_equals_the_python_definition():
    """oracle_synth_rows restates iscc_search_b200/synth.py (the definition of the SURVEY 8d data set)."""
    for start, n, seed in ((0, 5000, 1), (123_456_789, 3000, 7), (999_999_000, 1000, 4)):
        khi, klo, codes, lens = c_oracle.synth_rows(start, n, seed)
        assert klo is None
        want_lens = synth.make_lengths(start, n, seed)
        assert np.array_equal(lens, want_lens)
        assert np.array_equal(codes, synth.make_codes(start, n, seed, want_lens))
        assert np.array_equal(khi, synth.make_keys(start, n, seed))
    # simprint flavour: fixed 8-byte codes, chunk-pointer keys, duplicated chunks
    khi, klo, codes, lens = c_oracle.synth_rows(100, 4000, 4, lengths=(8,), key_mode=1, cpa=64, dup_every=16, dup_back=65)
    l8 = np.full(4000, 8, dtype=np.uint8)
    assert np.array_equal(lens, l8)
    assert np.array_equal(codes, synth.make_codes(100, 4000, 4, l8, dup_every=16, dup_back=65))
    hi, lo = synth.make_keys128(100, 4000, 4, 64)
    assert np.array_equal(khi, hi) and np.array_equal(klo, lo)
    assert (codes[111 - 100] == synth.make_codes(111 - 65, 1, 4, l8[:1])[0]).all()  # row 111 repeats row 46
    kb = synth.keys128_bytes(hi, lo)
    assert int.from_bytes(bytes(kb[5, :8]), "big") == int(hi[5]) and int.from_bytes(bytes(kb[5, 8:12]), "big") == ((105 % 64) * 4096)


def test_streaming_synth_topk_equals_in_memory_oracle():
    n, q, k = 70_000, 24, 50
    khi, _, codes, lens = c_oracle.synth_rows(0, n, 3)
    queries, qlens = synth.make_queries(q, n, 5, 3)
    rows, h, nb, cnt = c_oracle.topk(khi, None, codes, lens, queries, qlens, k)
    skhi, _, sh, snb, scnt = c_oracle.synth_topk(n, 3, queries, qlens, k, n_threads=3)
    assert np.array_equal(scnt, cnt)
    assert np.array_equal(skhi, khi[rows]) and np.array_equal(sh, h) and np.array_equal(snb, nb)
    # threshold + 128-bit keys + duplicates, fewer matches than k
    khi, klo, codes, lens = c_oracle.synth_rows(0, 30_000, 4, lengths=(8,), key_mode=1, dup_every=16, dup_back=65)
    qs = codes[[14, 206, 4110]].copy()  # each is repeated 65 rows later
    ql = np.full(3, 8, dtype=np.uint8)
    rows, h, nb, cnt = c_oracle.topk(khi, klo, codes, lens, qs, ql, 40, (0, 64))
    skhi, sklo, sh, snb, scnt = c_oracle.synth_topk(30_000, 4, qs, ql, 40, lengths=(8,), key_mode=1, dup_every=16, dup_back=65,
                                                    max_h_over_n=(0, 64), n_threads=2)
    assert np.array_equal(scnt, cnt) and (cnt >= 2).all()
    for i in range(3):
        c = int(cnt[i])
        assert np.array_equal(skhi[i, :c], khi[rows[i, :c]]) and np.array_equal(sklo[i, :c], klo[rows[i, :c]])
        assert np.array_equal(sh[i, :c], h[i, :c])


@pytest.mark.parametrize("force_scalar", [False, True])
def test_tuned_cpu_arm_equals_the_plain_restatement(force_scalar):
    """oracle_soa_topk (bucketed word planes, AVX-512 VPOPCNTDQ when present) is the timed CPU arm: same answers as `topk`."""
    n, q, k = 60_000, 40, 37
    lens = synth.make_lengths(0, n, 9, (5, 8, 16, 24, 32, 13))
    codes = synth.make_codes(0, n, 9, lens)
    keys = synth.make_keys(0, n, 9)
    queries, qlens = synth.make_queries(q, n, 10, 9, lengths=(8, 16, 24, 32, 7), row_lengths=(5, 8, 16, 24, 32, 13))
    want = c_oracle.topk(keys, None, codes, lens, queries, qlens, k)
    st = c_oracle.SoaStore(keys, None, codes, lens)
    got = st.topk(queries, qlens, k, n_threads=3, force_scalar=force_scalar)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    # threshold + 128-bit keys + duplicates
    khi, klo, codes, lens = c_oracle.synth_rows(0, 30_000, 4, lengths=(8,), key_mode=1, dup_every=16, dup_back=65)
    qs = codes[[14, 206, 4110, 999]].copy()
    qs[3, 2] ^= 0x55
    ql = np.full(4, 8, dtype=np.uint8)
    for thr in ((0, 64), (16, 64), None):
        want = c_oracle.topk(khi, klo, codes, lens, qs, ql, 50, thr)
        got = c_oracle.SoaStore(khi, klo, codes, lens).topk(qs, ql, 50, thr, force_scalar=force_scalar)
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
