"""
ORACLE - TEST INFRASTRUCTURE ONLY. Never imported by the product path (`iscc_search_b200/`).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this module, and only as the checker / the CPU arm.

CPU restatement (numpy) of the exact NPHD / Hamming top-k path that sits behind the reference's
two call sites:

* `ShardedNphdIndex.search(query, count=limit)`  - /root/reference/iscc_search/indexes/usearch/index.py:2037
* `ShardedIndex128.search(query_vectors, count)` - /root/reference/iscc_search/indexes/simprint/usearch_core.py:165

The arithmetic itself is NOT in the reference tree: it lives in the un-vendored dependencies
`iscc-usearch==0.8.1` (uv.lock:765-777) -> `usearch-iscc==2.24.6` (uv.lock:2491-2499). What is
restated here is the published formula

    NPHD(a, b) = popcount(a[:m] ^ b[:m]) / (8*m),   m = min(len(a), len(b)) bytes
    (docs/explanation/similarity-search.md:24-32, README.md:156-163)

together with the exact (brute-force) linear scan that `usearch.index.Index.search(exact=True)`
performs (existence shown by tests/test_usearch_search.py:588-622).

PARITY STATUS: **partially pinned**. Pinned by the reference's own literal test vectors (restated
in tests/golden/usearch_kats.json): raw Hamming bit counts as float32, uint64 keys, ascending
order, add/remove/get/contains semantics. **Parity unpinned** (no reference test or runnable
binary asserts them; the contract is the documented formula): (i) NPHD values at non-zero
distance, (ii) cross-length NPHD values, (iii) tie order among equal distances - defined HERE as
key ascending, (iv) float rounding of h/n - defined here as float32(h)/float32(n).
"""

import math

import numpy as np

MAX_BYTES = 32

#: common denominator that makes every h/n (n = 8*m, m in 1..32) an exact int64: d' = h * (LCM // n)
LCM_BITS = 8 * math.lcm(*range(1, MAX_BYTES + 1))

_POP8 = np.array([bin(i).count("1") for i in range(256)], dtype=np.uint16)


def pad_codes(vectors):
    # type: (list[bytes]) -> tuple[np.ndarray, np.ndarray]
    """Pack variable-length byte strings into (codes uint8[N,32] zero padded, lens uint8[N])."""
    n = len(vectors)
    codes = np.zeros((n, MAX_BYTES), dtype=np.uint8)
    lens = np.zeros(n, dtype=np.uint8)
    for i, v in enumerate(vectors):
        b = bytes(v)
        if not 1 <= len(b) <= MAX_BYTES:
            raise ValueError(f"code length {len(b)} outside 1..{MAX_BYTES} bytes")
        codes[i, : len(b)] = np.frombuffer(b, dtype=np.uint8)
        lens[i] = len(b)
    return codes, lens


def prefix_hamming(codes, lens, query, qlen):
    # type: (np.ndarray, np.ndarray, np.ndarray, int) -> tuple[np.ndarray, np.ndarray]
    """
    Hamming distance over the common byte prefix of one query against every stored code.

    Formula: docs/explanation/similarity-search.md:24-32 (prefix of min length, bit count).

    :return: (h uint16[N], nbits uint16[N]) with nbits = 8 * min(qlen, len_i)
    """
    m = np.minimum(lens.astype(np.int64), int(qlen))  # bytes compared per row
    x = np.bitwise_xor(codes, query[None, :MAX_BYTES])
    pc = _POP8[x]  # per-byte bit counts, uint16[N,32]
    col = np.arange(MAX_BYTES)[None, :]
    pc = np.where(col < m[:, None], pc, 0)
    h = pc.sum(axis=1).astype(np.uint16)
    return h, (8 * m).astype(np.uint16)


def topk_one(keys_hi, keys_lo, codes, lens, query, qlen, k, max_h_over_n=None):
    # type: (np.ndarray, np.ndarray|None, np.ndarray, np.ndarray, np.ndarray, int, int, tuple[int,int]|None) -> tuple
    """
    Exact top-k of one query. Order: (h/n ascending as exact rational, key ascending).

    Keys are unsigned 64-bit (`keys_lo is None`) or 128-bit big-endian split into (hi, lo).
    `max_h_over_n=(a, b)` keeps only rows with h/n <= a/b (simprint threshold mode).

    :return: (index int64[c], h uint16[c], nbits uint16[c]) with c <= k rows into the store arrays
    """
    if k < 1:
        raise ValueError("`count` must be >= 1")  # tests/test_usearch_search.py:678-685
    if len(lens) == 0:
        e = np.zeros(0, dtype=np.int64)
        return e, e.astype(np.uint16), e.astype(np.uint16)
    h, nb = prefix_hamming(codes, lens, query, qlen)
    dprime = h.astype(np.int64) * (LCM_BITS // nb.astype(np.int64))
    if keys_lo is None:
        order = np.lexsort((keys_hi, dprime))
    else:
        order = np.lexsort((keys_lo, keys_hi, dprime))
    if max_h_over_n is not None:
        a, b = max_h_over_n
        keep = h.astype(np.int64)[order] * b <= a * nb.astype(np.int64)[order]
        order = order[keep]
    order = order[:k]
    return order.astype(np.int64), h[order], nb[order]


def topk(keys_hi, keys_lo, codes, lens, queries, qlens, k, max_h_over_n=None):
    # type: (...) -> list[tuple[np.ndarray, np.ndarray, np.ndarray]]
    """Batch form of `topk_one` (tests/test_usearch_search.py:58-119 shows the batch call shape)."""
    return [
        topk_one(keys_hi, keys_lo, codes, lens, queries[i], int(qlens[i]), k, max_h_over_n)
        for i in range(len(qlens))
    ]


def nphd_distance_f32(h, nbits):
    # type: (np.ndarray, np.ndarray) -> np.ndarray
    """float32 NPHD distance as the reference consumes it (index.py:2041: `float(distance)`)."""
    return (np.asarray(h, dtype=np.float32) / np.asarray(nbits, dtype=np.float32)).astype(np.float32)


def unit_score(h, nbits):
    # type: (int, int) -> float
    """`max(0.0, 1.0 - float(distance))` - /root/reference/iscc_search/indexes/usearch/index.py:2039-2043."""
    return max(0.0, 1.0 - float(np.float32(h) / np.float32(nbits)))


def simprint_score(h, ndim):
    # type: (int, int) -> float
    """`1.0 - (distance / ndim)` in double - /root/reference/iscc_search/indexes/simprint/usearch_core.py:179-182."""
    return 1.0 - (float(h) / ndim)


class StoreOracle:
    """
    Dict model of the key->vector store semantics the reference relies on (single vector per key).

    add: duplicate key silently skipped, first wins   - tests/test_usearch_add.py:53-63
    remove: returns number removed, missing key -> 0  - tests/test_usearch_remove.py:19-48
    get: vector or None                               - tests/test_usearch_get.py:15-56
    contains over full uint64 range                   - tests/test_usearch_contains.py:214-235
    """

    def __init__(self):
        self.rows = {}  # key (int) -> bytes

    def add(self, keys, vectors):
        added = []
        for key, v in zip(keys, vectors):
            if key in self.rows:
                added.append(False)
                continue
            self.rows[key] = bytes(v)
            added.append(True)
        return added

    def remove(self, keys):
        return sum(1 for key in keys if self.rows.pop(key, None) is not None)

    def get(self, key):
        return self.rows.get(key)

    def __contains__(self, key):
        return key in self.rows

    def __len__(self):
        return len(self.rows)

    def arrays(self, key_bytes=8):
        """(keys_hi, keys_lo|None, codes, lens) in insertion order."""
        ks = list(self.rows.keys())
        codes, lens = pad_codes([self.rows[key] for key in ks])
        if key_bytes == 8:
            return np.array(ks, dtype=np.uint64), None, codes, lens
        hi = np.array([int.from_bytes(key[:8], "big") for key in ks], dtype=np.uint64)
        lo = np.array([int.from_bytes(key[8:], "big") for key in ks], dtype=np.uint64)
        return hi, lo, codes, lens

    def search(self, query, k, key_bytes=8, max_h_over_n=None):
        """-> list of (key, h, nbits) best first."""
        hi, lo, codes, lens = self.arrays(key_bytes)
        q = np.zeros(MAX_BYTES, dtype=np.uint8)
        qb = bytes(query)
        q[: len(qb)] = np.frombuffer(qb, dtype=np.uint8)
        idx, h, nb = topk_one(hi, lo, codes, lens, q, len(qb), k, max_h_over_n)
        ks = list(self.rows.keys())
        return [(ks[i], int(hh), int(nn)) for i, hh, nn in zip(idx, h, nb)]
