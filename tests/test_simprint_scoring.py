"""
CPU tests (no GPU): the host-side simprint scoring of iscc_search_b200.simprint against fixtures produced
by the reference's own Python (tests/golden/make_golden.py). Neighbours come from the oracle here; the
GPU run of the same fixtures is tests/test_gpu_simprint.py. Floats must be bit-identical.
"""

import json
from pathlib import Path

import numpy as np
import pytest

from iscc_search_b200 import simprint as sp
from oracle.nphd_oracle import StoreOracle

GOLD = json.loads((Path(__file__).parent / "golden" / "simprint_scoring.json").read_text())


def test_calculate_idf_matches_reference():
    for f, t, expected in GOLD["calculate_idf"]:
        assert sp.calculate_idf(f, t) == expected
    # literal values pinned by the reference's tests (tests/test_indexes_simprint_lmdb_ops.py:89-111)
    assert sp.calculate_idf(5, 0) == 0.0 and sp.calculate_idf(0, 10) == pytest.approx(np.log(11.0))


def test_chunk_pointer_round_trip_and_limits():
    for a, o, s, packed in GOLD["pack_chunk_pointer"]:
        assert sp.pack_chunk_pointer(bytes.fromhex(a), o, s).hex() == packed
        assert sp.unpack_chunk_pointer(bytes.fromhex(packed)) == (bytes.fromhex(a), o, s)
    with pytest.raises(ValueError, match="must be 8 bytes"):
        sp.pack_chunk_pointer(b"\x00" * 7, 0, 0)
    with pytest.raises(ValueError, match="exceeds max"):
        sp.pack_chunk_pointer(b"\x00" * 8, 2**32, 0)
    with pytest.raises(ValueError, match="Expected 16 bytes"):
        sp.unpack_chunk_pointer(b"\x00" * 15)


def test_coverage_quality_score_matches_reference():
    for case in GOLD["coverage_quality"]:
        matches = [(bytes.fromhex(a), bytes.fromhex(b), o, s) for a, b, o, s in case["matches"]]
        freqs = {bytes.fromhex(k): v for k, v in case["freqs"].items()}
        assert sp.coverage_quality_score(matches, freqs, case["num_queried"]) == case["score"]


def _check_results(got, expected):
    assert len(got) == len(expected)
    for g, e in zip(got, expected):
        assert g.iscc_id_body.hex() == e["iscc_id_body"] and g.score == e["score"]
        assert g.queried == e["queried"] and g.matches == e["matches"]
        assert [[c.query.hex(), c.match.hex(), c.score, c.offset, c.size, c.freq] for c in g.chunks] == e["chunks"]


def _model(rows):
    st = StoreOracle()
    st.add([bytes.fromhex(k) for k, _ in rows], [bytes.fromhex(v) for _, v in rows])
    return st


def test_exact_join_scoring_matches_reference():
    for case in GOLD["search_exact"]:
        st = _model(case["rows"])
        query = [bytes.fromhex(q) for q in case["query"]]
        per_query = [[k for k, h, _n in st.search(q, case["dup_limit"], key_bytes=16, max_h_over_n=(0, 1))] for q in query]
        for q, keys in zip(query, per_query):
            assert len({k[:8] for k in keys}) == case["doc_freq"][q.hex()]
        _check_results(sp.score_exact(query, per_query, case["limit"], case["threshold"], True), case["result"])


def test_approximate_scoring_matches_reference():
    for case in GOLD["search_raw"]:
        st = _model(case["rows"])
        ndim, nb = case["ndim"], case["ndim"] // 8
        query = [bytes.fromhex(q) for q in case["query"]]
        count = max(1, case["limit"] * case["oversampling"])
        S = len(query)
        keys = np.zeros((S, count, 16), dtype=np.uint8)
        ham = np.zeros((S, count), dtype=np.uint16)
        vec = np.zeros((S, count, nb), dtype=np.uint8)
        counts = np.zeros(S, dtype=np.int64)
        for i, q in enumerate(query):
            hits = st.search(q, count, key_bytes=16)
            counts[i] = len(hits)
            for j, (k, h, _n) in enumerate(hits):
                keys[i, j], ham[i, j], vec[i, j] = np.frombuffer(k, np.uint8), h, np.frombuffer(st.get(k), np.uint8)
        doc_freq = case["doc_freq"]
        from tests.fakes import OracleStore

        got = sp.score_approx(query, keys, ham, vec, counts, ndim, case["limit"], case["threshold"], True,
                              lambda s: doc_freq.get(bytes(s).hex(), 0), case["total_assets"], scorer=OracleStore().score_segments)
        _check_results(got, case["result"])
