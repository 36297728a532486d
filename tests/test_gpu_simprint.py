"""GPU: the full B200SimprintIndex (search on the device, scoring on the host) against the reference-generated fixtures."""

import json
from pathlib import Path

import numpy as np
import pytest

from iscc_search_b200.simprint import B200SimprintIndex, pack_chunk_pointer

pytestmark = pytest.mark.gpu
GOLD = json.loads((Path(__file__).parent / "golden" / "simprint_scoring.json").read_text())


def _check_results(got, expected):
    assert len(got) == len(expected)
    for g, e in zip(got, expected):
        assert g.iscc_id_body.hex() == e["iscc_id_body"] and g.score == e["score"]
        assert g.queried == e["queried"] and g.matches == e["matches"]
        assert [[c.query.hex(), c.match.hex(), c.score, c.offset, c.size, c.freq] for c in g.chunks] == e["chunks"]


def _build(rows, ndim):
    idx = B200SimprintIndex(path=None, ndim=ndim)
    idx.add_raw([bytes.fromhex(k) for k, _ in rows], [np.frombuffer(bytes.fromhex(v), dtype=np.uint8) for _, v in rows])
    return idx


def test_search_raw_matches_reference_fixtures(cuda):
    for case in GOLD["search_raw"]:
        idx = _build(case["rows"], case["ndim"])
        idx.oversampling_factor = case["oversampling"]
        doc_freq = case["doc_freq"]
        got = idx.search_raw([bytes.fromhex(q) for q in case["query"]], limit=case["limit"], threshold=case["threshold"], detailed=True,
                             doc_freq_fn=lambda s: doc_freq.get(bytes(s).hex(), 0), total_assets=case["total_assets"])
        _check_results(got, case["result"])
        idx.close()


def test_search_raw_with_index_doc_frequencies_equals_callback_path(cuda):
    # doc_freq_fn="index": one batched equality join instead of one lookup per simprint - same numbers
    for case in GOLD["search_raw"][:9]:
        idx = _build(case["rows"], case["ndim"])
        idx.oversampling_factor = case["oversampling"]
        doc_freq = case["doc_freq"]
        query = [bytes.fromhex(q) for q in case["query"]]
        assert idx.doc_freqs(query) == {q: doc_freq.get(q.hex(), 0) for q in query}
        got = idx.search_raw(query, limit=case["limit"], threshold=case["threshold"], detailed=True, doc_freq_fn="index",
                             total_assets=case["total_assets"])
        _check_results(got, case["result"])
        idx.close()


def test_search_exact_and_doc_freq_match_reference_fixtures(cuda):
    for case in GOLD["search_exact"]:
        idx = _build(case["rows"], 64)
        query = [bytes.fromhex(q) for q in case["query"]]
        for q in query:
            assert idx.doc_freq(q, case["dup_limit"]) == case["doc_freq"][q.hex()]
        got = idx.search_exact(query, case["total_assets"], case["limit"], case["threshold"], True, case["dup_limit"])
        _check_results(got, case["result"])
        idx.close()


def test_reference_behaviours_restated(cuda):
    # tests/test_indexes_usearch_simprint_approx.py:186-202, 214-230, 257-280, 308-330, 392-414
    idx = B200SimprintIndex(path=None, ndim=64)
    assert idx.search_raw([b"\xaa" * 8]) == [] and idx.size == 0 and idx.shard_count == 0
    a1, a2 = b"\x01" * 8, b"\x02" * 8
    sp1, sp2 = b"\xaa" * 8, b"\x55" * 8
    k1 = pack_chunk_pointer(a1, 0, 100)
    idx.add_raw([k1, k1], [np.frombuffer(sp1, np.uint8), np.frombuffer(sp2, np.uint8)])      # in-batch dup: first wins
    assert idx.size == 1 and k1 in idx
    idx.add_raw([pack_chunk_pointer(a1, 100, 50), pack_chunk_pointer(a2, 0, 10)], [np.frombuffer(sp2, np.uint8), np.frombuffer(sp1, np.uint8)])
    res = idx.search_raw([sp1, sp2], limit=10, threshold=0.9, detailed=True, total_assets=2)
    by_asset = {r.iscc_id_body: r for r in res}
    assert by_asset[a1].matches == 2 and by_asset[a1].score == 1.0           # multi-chunk grouping, both queries matched
    assert by_asset[a2].matches == 1 and by_asset[a2].score == 0.5           # unmatched query simprint penalises coverage
    assert res[0].iscc_id_body == a1
    flipped = bytes([0x55] * 6 + [0xaa] * 2)                                 # 48 of 64 bits differ from sp1 -> score 0.25
    assert idx.search_raw([flipped], limit=10, threshold=0.9) == [] or all(r.score < 1.0 for r in idx.search_raw([flipped], limit=10, threshold=0.9))
    kept = idx.search_raw([flipped], limit=10, threshold=0.0, detailed=True)
    assert any(c.score == 0.25 for r in kept for c in r.chunks)
    idx.remove([k1])
    assert k1 not in idx and idx.size == 2
    assert all(r.iscc_id_body != a1 or r.matches == 1 for r in idx.search_raw([sp1, sp2], limit=10, threshold=0.9, total_assets=2))
    idx.close()


def test_first_of_asset_flags_equal_host_grouping(cuda):
    # device-side grouping (isx_search first_of_asset_out) against the numpy first-occurrence of the asset id
    from iscc_search_b200 import ShardedIndex128

    rng = np.random.default_rng(4)
    n, nq, count = 40_000, 9, 3000
    assets = rng.integers(0, 500, size=n, dtype=np.uint64)
    assets[:50] = np.uint64(2**64 - 1)                       # the asset id that equals the table's empty marker
    keys = np.zeros((n, 16), dtype=np.uint8)
    keys[:, :8] = assets.astype(">u8").view(np.uint8).reshape(n, 8)
    keys[:, 8:] = np.arange(n, dtype=np.uint64).astype(">u8").view(np.uint8).reshape(n, 8)
    vecs = rng.integers(0, 256, size=(n, 8), dtype=np.uint8)
    idx = ShardedIndex128(ndim=64)
    idx.add(keys, vecs)
    res = idx.search(vecs[:nq], count=count, with_first=True)
    assert res.first.shape == (nq, count)
    for i in range(nq):
        c = int(res.counts[i])
        a = np.ascontiguousarray(res.keys[i, :c, :8]).view(">u8").ravel()
        _, first_idx = np.unique(a, return_index=True)
        expect = np.zeros(c, dtype=np.uint8)
        expect[first_idx] = 1
        assert np.array_equal(res.first[i, :c], expect)
    single = idx.search(vecs[0], count=10, with_first=True)
    assert single.first.shape == (10,) and single.first[0] == 1
    idx.close()
