"""GPU: unbounded threshold match (isx_match_all) and the INSTANCE bidirectional prefix semantics of index.py:1957-2022."""

import numpy as np
import pytest

from iscc_search_b200 import synth
from iscc_search_b200._lib import Store
from iscc_search_b200.instance import InstancePrefixIndex
from oracle import nphd_oracle
from tests.helpers import make_store_arrays

pytestmark = pytest.mark.gpu


def test_match_all_equals_oracle_threshold_set(cuda):
    n = 120_000
    keys, codes, lens = make_store_arrays(n, 41)
    codes[: n // 50, :8] = codes[0, :8]          # a few thousand rows share the first 64 bits
    st = Store(key_bytes=8, max_bytes=32)
    st.add(keys, codes, lens)
    for qlen, thr in ((8, (0, 1)), (8, (6, 64)), (32, (40, 256)), (16, (0, 1))):
        q = codes[0].copy()
        q[qlen:] = 0
        gk, gh, gn = st.match_all(bytes(q[:qlen]), thr=thr, max_out=64)    # small buffer: forces the grow-and-retry path
        h, nb = nphd_oracle.prefix_hamming(codes, lens, q, qlen)
        sel = h.astype(np.int64) * thr[1] <= thr[0] * nb.astype(np.int64)
        assert len(gk) == int(sel.sum())
        order_g, order_e = np.argsort(gk), np.argsort(keys[sel])
        assert np.array_equal(gk[order_g], keys[sel][order_e])
        assert np.array_equal(gh[order_g], h[sel][order_e]) and np.array_equal(gn[order_g], nb[sel][order_e])
    st.close()


def test_instance_bidirectional_prefix_matches_reference_rules(cuda):
    # tests/test_indexes_usearch_index.py:141-215: any prefix relation in either direction is a 1.0 match
    rng = np.random.default_rng(9)
    full = bytes(rng.integers(0, 256, size=32, dtype=np.uint8))
    other = bytes(rng.integers(0, 256, size=32, dtype=np.uint8))
    idx = InstancePrefixIndex()
    idx.add(1, full)            # 256-bit
    idx.add(2, full[:16])       # 128-bit prefix of the same data
    idx.add(3, full[:8])        # 64-bit prefix
    idx.add(4, other)           # unrelated
    idx.add(5, full[:24])       # 192-bit body: matched forward, NOT in the reference's reverse probes
    idx.add(1, full[:8])        # the same asset also carries a 64-bit unit (second dupsort entry)
    assert len(idx) == 6
    assert idx.search(full) == {1: 1.0, 2: 1.0, 3: 1.0}                   # 256-bit query: itself + 128/64-bit prefixes; 192-bit skipped
    assert idx.search(full[:16]) == {1: 1.0, 2: 1.0, 3: 1.0, 5: 1.0}      # 128-bit query: forward hits 256/192/128, reverse 64
    assert idx.search(full[:8]) == {1: 1.0, 2: 1.0, 3: 1.0, 5: 1.0}       # 64-bit query: everything starting with it
    assert idx.search(other[:8]) == {4: 1.0}
    assert idx.search(bytes(8)) == {}
    assert idx.remove_many([(1, full), (1, full[:8]), (1, other)]) == 2 and idx.search(full) == {2: 1.0, 3: 1.0}
    idx.add_many([(7, full[:8]), (7, other[:8]), (7, full[:8])])   # same length, two bodies; the repeated pair is skipped
    assert len(idx) == 6 and idx.search(other) == {4: 1.0, 7: 1.0} and list(idx.search(full[:8])) == [2, 3, 5, 7]
    idx.close()
