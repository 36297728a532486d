"""
TEST-ONLY double of `iscc_search_b200._lib.Store` (the ctypes owner of one C-ABI store handle), answered by the
CPU oracle. It lets the CPU suite exercise the host layers above the C ABI (drop-in classes, simprint scoring,
INSTANCE prefix index, protocol backend) without a GPU. It is never importable from the product package: the
product has no CPU path (tests/test_abi.py checks that nothing under iscc_search_b200/ imports tests/ or oracle/).
"""

import pickle

import numpy as np

from oracle import nphd_oracle


class OracleStore:
    def __init__(self, device=0, key_bytes=8, max_bytes=32, fixed_len=0):
        self.device, self.key_bytes, self.max_bytes, self.fixed_len = device, key_bytes, max_bytes, fixed_len
        self.rows = {}  # key (int | bytes16) -> code bytes, insertion ordered
        self.launches = 0

    def release_scratch(self):
        return 0

    def score_segments(self, seg, rec_qi, rec_sim, rec_idf, q_idf):
        """Test double of isx_score_segments: the same sums as plain Python floats (left to right, no compensation)."""
        out = np.zeros(len(seg) - 1, dtype=np.float64)
        for a in range(len(seg) - 1):
            total, weighted = 0.0, 0.0
            matched = set()
            for i in range(int(seg[a]), int(seg[a + 1])):
                total += float(rec_idf[i])
                weighted += float(rec_idf[i]) * float(rec_sim[i])
                matched.add(int(rec_qi[i]))
            for q in range(len(q_idf)):
                if q not in matched:
                    total += float(q_idf[q])
            out[a] = weighted / total if total > 0 else 0.0
        return out

    # -- helpers
    def _key(self, keys, i):
        return int(keys[i]) if self.key_bytes == 8 else bytes(np.asarray(keys[i], dtype=np.uint8).tobytes())

    def _arrays(self):
        ks = list(self.rows)
        codes, lens = nphd_oracle.pad_codes([self.rows[k] for k in ks]) if ks else (np.zeros((0, 32), np.uint8), np.zeros(0, np.uint8))
        if self.key_bytes == 8:
            return ks, np.array(ks, dtype=np.uint64), None, codes, lens
        hi = np.array([int.from_bytes(k[:8], "big") for k in ks], dtype=np.uint64)
        lo = np.array([int.from_bytes(k[8:], "big") for k in ks], dtype=np.uint64)
        return ks, hi, lo, codes, lens

    # -- lifecycle / rows
    def close(self):
        self.rows = {}

    @property
    def handle(self):
        return self

    def size(self):
        return len(self.rows)

    def device_bytes(self):
        return sum(len(v) for v in self.rows.values())

    def length_mask(self):
        m = 0
        for v in self.rows.values():
            m |= 1 << (len(v) - 1)
        return m

    def max_k(self):
        return 65536

    def clear(self):
        self.rows = {}

    def add(self, keys, codes, lens):
        added = np.zeros(len(lens), dtype=np.uint8)
        for i in range(len(lens)):
            k = self._key(keys, i)
            if k in self.rows:
                continue
            n = int(lens[i])
            if n < 1 or n > self.max_bytes or (self.fixed_len and n != self.fixed_len):
                raise ValueError(f"row {i}: code length {n} not accepted by this store")
            self.rows[k] = bytes(np.asarray(codes[i][:n], dtype=np.uint8).tobytes())
            added[i] = 1
        return added

    def remove(self, keys, n):
        removed = np.zeros(n, dtype=np.uint8)
        for i in range(n):
            if self.rows.pop(self._key(keys, i), None) is not None:
                removed[i] = 1
        return removed, int(removed.sum())

    def contains(self, keys, n):
        return np.array([self._key(keys, i) in self.rows for i in range(n)], dtype=bool)

    def get(self, keys, n):
        codes, lens = np.zeros((n, 32), dtype=np.uint8), np.zeros(n, dtype=np.uint8)
        for i in range(n):
            v = self.rows.get(self._key(keys, i))
            if v is not None:
                codes[i, : len(v)] = np.frombuffer(v, dtype=np.uint8)
                lens[i] = len(v)
        return codes, lens

    def save(self, path):
        with open(path, "wb") as fh:
            pickle.dump(self.rows, fh)

    def load(self, path):
        with open(path, "rb") as fh:
            self.rows = pickle.load(fh)

    # -- search
    def search(self, queries, qlens, k, thr=None, with_codes=False, first_of_asset=None):
        if k < 1:
            raise ValueError("`count` must be >= 1")
        self.launches += 1
        q = len(qlens)
        ks, hi, lo, codes, lens = self._arrays()
        keys = np.zeros((q, k), dtype=np.uint64) if self.key_bytes == 8 else np.zeros((q, k, 16), dtype=np.uint8)
        h = np.zeros((q, k), dtype=np.uint16)
        nb = np.zeros((q, k), dtype=np.uint16)
        counts = np.zeros(q, dtype=np.uint32)
        out_codes = np.zeros((q, k, 32), dtype=np.uint8) if with_codes else None
        for i in range(q):
            query = np.ascontiguousarray(queries[i], dtype=np.uint8)
            rows, hh, nn = nphd_oracle.topk_one(hi, lo, codes, lens, query, int(qlens[i]), k, thr)
            c = len(rows)
            counts[i] = c
            h[i, :c], nb[i, :c] = hh, nn
            seen = set()
            for j, r in enumerate(rows):
                key = ks[int(r)]
                if self.key_bytes == 8:
                    keys[i, j] = key
                else:
                    keys[i, j] = np.frombuffer(key, dtype=np.uint8)
                    if first_of_asset is not None:
                        first_of_asset[i, j] = key[:8] not in seen
                        seen.add(key[:8])
                if out_codes is not None:
                    out_codes[i, j] = codes[int(r)]
        return keys, h, nb, counts, out_codes

    def match_all(self, query, thr=(0, 1), max_out=4096):
        self.launches += 1
        ks, hi, lo, codes, lens = self._arrays()
        qb = bytes(query)
        q = np.zeros(32, dtype=np.uint8)
        q[: len(qb)] = np.frombuffer(qb, dtype=np.uint8)
        if not ks:
            e = np.zeros(0, dtype=np.uint16)
            return (np.zeros(0, dtype=np.uint64) if self.key_bytes == 8 else np.zeros((0, 16), dtype=np.uint8)), e, e
        h, nb = nphd_oracle.prefix_hamming(codes, lens, q, len(qb))
        sel = np.nonzero(h.astype(np.int64) * thr[1] <= thr[0] * nb.astype(np.int64))[0]
        if self.key_bytes == 8:
            keys = np.array([ks[i] for i in sel], dtype=np.uint64)
        else:
            keys = np.array([np.frombuffer(ks[i], dtype=np.uint8) for i in sel], dtype=np.uint8).reshape(-1, 16)
        return keys, h[sel], nb[sel]

    def set_stream(self, cuda_stream):
        pass

    def set_profiling(self, enabled):
        pass

    def stats(self):
        return {"kernel_launches": self.launches}
