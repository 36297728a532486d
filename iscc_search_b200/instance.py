"""
INSTANCE-code identity matching as an `NPHD == 0` scan.

The reference keeps INSTANCE units in an LMDB dupsort table and answers a query with a bidirectional
prefix match, every hit scoring 1.0 (/root/reference/iscc_search/indexes/usearch/index.py:1957-2022).
A stored code and a query match bidirectionally iff their common byte prefix is identical, i.e. iff the
prefix Hamming distance is 0 - exactly the store's threshold match with thr = 0/1 (`isx_match_all`,
unbounded output). One asset may carry INSTANCE units of several lengths (index.py:363-366 puts one dupsort
entry per unit), so rows are keyed by (ISCC-ID, length, digest of the body) in a 128-bit-key store.

Parity nuance kept on purpose: the reference's reverse direction only probes the 128- and 64-bit prefixes of
the query (index.py:1989, 2003-2020), so a stored 192-bit body that prefixes a 256-bit query is NOT a hit there;
`search` drops those rows (their compared length is visible in `nbits`).
"""

import hashlib
import struct

import numpy as np


def _default_store(device):
    from iscc_search_b200._lib import Store

    return Store(device=device, key_bytes=16, max_bytes=32, fixed_len=0)


class InstancePrefixIndex:
    """HBM-resident replacement for the `__instance__` dupsort table: add / remove / bidirectional prefix search."""

    def __init__(self, device=0, store=None):
        self._store = store if store is not None else _default_store(device)

    MAX_SLOTS = 256

    @staticmethod
    def _key(iscc_id_key, code, slot=0):
        # (ISCC-ID, body) pairs are unique in the dupsort table (index.py:365-366, dupdata=False). The 16-byte row key is
        # ISCC-ID | body length | 48-bit digest of the body | slot. The digest only spreads the bodies of one asset over
        # key space; identity is decided by the stored body itself: `_locate` compares it and walks the slot byte when two
        # different bodies of one asset and length share a digest, so no row is ever dropped or confused (exact, not
        # probabilistic).
        return struct.pack(">QB", int(iscc_id_key), len(code)) + hashlib.blake2b(code, digest_size=6).digest() + bytes([slot])

    def _rows(self, pairs, slots=None):
        n = len(pairs)
        keys = np.zeros((n, 16), dtype=np.uint8)
        codes = np.zeros((n, 32), dtype=np.uint8)
        lens = np.zeros(n, dtype=np.uint8)
        for i, (iscc_id_key, code) in enumerate(pairs):
            code = bytes(code)
            if not 1 <= len(code) <= 32:
                raise ValueError(f"INSTANCE body must be 1..32 bytes, got {len(code)}")
            keys[i] = np.frombuffer(self._key(iscc_id_key, code, 0 if slots is None else slots[i]), dtype=np.uint8)
            codes[i, : len(code)] = np.frombuffer(code, dtype=np.uint8)
            lens[i] = len(code)
        return keys, codes, lens

    def _locate(self, pairs):
        # type: (list[tuple[int, bytes]]) -> tuple[list[int], list[bool]]
        """Per pair: (slot, present) - the slot that holds exactly this body, or the first free slot of its digest chain."""
        slots, present = [0] * len(pairs), [False] * len(pairs)
        todo = list(range(len(pairs)))
        for slot in range(self.MAX_SLOTS):
            if not todo:
                break
            sub = [pairs[i] for i in todo]
            keys, codes, lens = self._rows(sub, [slot] * len(sub))
            got_codes, got_lens = self._store.get(keys, len(sub))
            nxt = []
            for j, i in enumerate(todo):
                slots[i] = slot
                if got_lens[j] == 0:
                    continue                                   # free slot: the body is not stored
                if got_lens[j] == lens[j] and np.array_equal(got_codes[j], codes[j]):
                    present[i] = True                          # this very body
                else:
                    nxt.append(i)                              # digest collision with another body: next slot
            todo = nxt
        if todo:
            raise ValueError("INSTANCE digest chain exhausted")   # 256 bodies of one asset and length with one 48-bit digest
        return slots, present

    def add(self, iscc_id_key, instance_code):
        # type: (int, bytes) -> None
        """Register one INSTANCE unit body (8/16/32 bytes) of the asset with integer ISCC-ID key (index.py:363-366)."""
        self.add_many([(iscc_id_key, instance_code)])

    def add_many(self, pairs):
        # type: (list[tuple[int, bytes]]) -> None
        """One batched add of (ISCC-ID key, body) pairs; pairs already present are skipped (dupdata=False)."""
        if not pairs:
            return
        pairs = list(dict.fromkeys((int(k), bytes(c)) for k, c in pairs))   # in-batch duplicates collapse
        slots, present = self._locate(pairs)
        new = [i for i in range(len(pairs)) if not present[i]]
        first_of_chain, later = {}, []
        for i in new:   # bodies of one batch that share a digest chain go in one after the other (never happens in practice)
            chain = self._key(*pairs[i])[:15]
            if chain in first_of_chain:
                later.append(i)
            else:
                first_of_chain[chain] = i
        new = list(first_of_chain.values())
        if new:
            self._store.add(*self._rows([pairs[i] for i in new], [slots[i] for i in new]))
        for i in later:
            self.add_many([pairs[i]])

    def remove(self, iscc_id_key, instance_code):
        # type: (int, bytes) -> int
        """Drop one (ISCC-ID, body) row - the update path's stale-body delete (index.py:339-348)."""
        return self.remove_many([(iscc_id_key, instance_code)])

    def remove_many(self, pairs):
        # type: (list[tuple[int, bytes]]) -> int
        if not pairs:
            return 0
        pairs = list(dict.fromkeys((int(k), bytes(c)) for k, c in pairs))
        slots, present = self._locate(pairs)
        hit = [i for i in range(len(pairs)) if present[i]]
        if not hit:
            return 0
        keys, _codes, _lens = self._rows([pairs[i] for i in hit], [slots[i] for i in hit])
        removed = self._store.remove(keys, len(keys))[1]
        # keep digest chains gap-free (lookups stop at the first free slot): a body stored behind a removed one moves up
        for i in hit:
            key_id, code = pairs[i]
            slot = slots[i]
            while slot + 1 < self.MAX_SLOTS:
                nxt = np.frombuffer(self._key(key_id, code, slot + 1), dtype=np.uint8).reshape(1, 16).copy()
                got_codes, got_lens = self._store.get(nxt, 1)
                if got_lens[0] == 0:
                    break
                self._store.remove(nxt, 1)
                moved = np.frombuffer(self._key(key_id, code, slot), dtype=np.uint8).reshape(1, 16).copy()
                self._store.add(moved, got_codes[:1].copy(), got_lens[:1].copy())
                slot += 1
        return removed

    def search(self, instance_code):
        # type: (bytes) -> dict[int, float]
        """
        Same return shape as `_search_instance_unit`: {ISCC-ID key: 1.0} for every bidirectional prefix match.
        Insertion order of the dict is ascending ISCC-ID (deterministic; the reference's is LMDB cursor order).
        """
        code = bytes(instance_code)
        keys, _h, nbits = self._store.match_all(code, thr=(0, 1))
        qbits = 8 * len(code)
        hits = set()
        for k, n in zip(keys, nbits):
            n = int(n)
            # forward hits compare the whole query (n == qbits); reverse hits compare a shorter stored body:
            # the reference only probes 128- and 64-bit stored prefixes (index.py:2003-2020)
            if n < qbits and n not in (64, 128):
                continue
            hits.add(struct.unpack(">Q", bytes(k[:8]))[0])
        return {key: 1.0 for key in sorted(hits)}

    def __len__(self):
        return self._store.size()

    def release_scratch(self):
        return self._store.release_scratch()

    def close(self):
        self._store.close()
