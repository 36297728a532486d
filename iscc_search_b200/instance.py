"""
INSTANCE-code identity matching as an `NPHD == 0` scan.

The reference keeps INSTANCE units in an LMDB dupsort table and answers a query with a bidirectional
prefix match, every hit scoring 1.0 (/root/reference/iscc_search/indexes/usearch/index.py:1957-2022).
A stored code and a query match bidirectionally iff their common byte prefix is identical, i.e. iff the
prefix Hamming distance is 0 - exactly the store's threshold match with thr = 0/1 (`isx_match_all`,
unbounded output). One asset may carry INSTANCE units of several lengths (index.py:363-366 puts one dupsort
entry per unit), so rows are keyed by (ISCC-ID, length) in a 128-bit-key store.

Parity nuance kept on purpose: the reference's reverse direction only probes the 128- and 64-bit prefixes of
the query (index.py:1989, 2003-2020), so a stored 192-bit body that prefixes a 256-bit query is NOT a hit there;
`search` drops those rows (their compared length is visible in `nbits`).
"""

import struct

import numpy as np

from iscc_search_b200._lib import Store


class InstancePrefixIndex:
    """HBM-resident replacement for the `__instance__` dupsort table: add / remove / bidirectional prefix search."""

    def __init__(self, device=0):
        self._store = Store(device=device, key_bytes=16, max_bytes=32, fixed_len=0)

    @staticmethod
    def _key(iscc_id_key, nbytes):
        return struct.pack(">QQ", int(iscc_id_key), nbytes)

    def add(self, iscc_id_key, instance_code):
        # type: (int, bytes) -> None
        """Register one INSTANCE unit body (8/16/32 bytes) of the asset with integer ISCC-ID key (index.py:363-366)."""
        code = bytes(instance_code)
        codes = np.zeros((1, 32), dtype=np.uint8)
        codes[0, : len(code)] = np.frombuffer(code, dtype=np.uint8)
        key = np.frombuffer(self._key(iscc_id_key, len(code)), dtype=np.uint8).reshape(1, 16).copy()
        self._store.add(key, codes, np.array([len(code)], dtype=np.uint8))

    def remove_asset(self, iscc_id_key):
        # type: (int) -> int
        """Drop every INSTANCE unit of an asset (update path: remove-before-add, index.py:433-437)."""
        keys = np.stack([np.frombuffer(self._key(iscc_id_key, n), dtype=np.uint8) for n in range(1, 33)]).copy()
        return self._store.remove(keys, len(keys))[1]

    def search(self, instance_code):
        # type: (bytes) -> dict[int, float]
        """Same return shape as `_search_instance_unit`: {ISCC-ID key: 1.0} for every bidirectional prefix match."""
        code = bytes(instance_code)
        keys, _h, nbits = self._store.match_all(code, thr=(0, 1))
        results = {}
        qbits = 8 * len(code)
        for k, n in zip(keys, nbits):
            n = int(n)
            # forward hits compare the whole query (n == qbits); reverse hits compare a shorter stored body:
            # the reference only probes 128- and 64-bit stored prefixes (index.py:2003-2020)
            if n < qbits and n not in (64, 128):
                continue
            results[struct.unpack(">Q", bytes(k[:8]))[0]] = 1.0
        return results

    def __len__(self):
        return self._store.size()

    def close(self):
        self._store.close()
