"""
Deterministic synthetic ISCC-UNIT / simprint data, shared by tests and bench.py.

Row `i`, 64-bit word `w` of dataset seed `s` is `splitmix64(s ^ (i*4 + w))`, so any row subset can be
regenerated anywhere (CPU oracle, any rank) without materialising the full set. Keys are a bijective
64-bit mix of the row number: unique, unsorted, spanning uint64. Queries are stored rows with a few
random bit flips (true neighbours exist) plus ~10 % pure-random codes.
"""

import numpy as np

_U = np.uint64
STANDARD_LENGTHS = (8, 16, 24, 32)  # bytes: 64/128/192/256-bit ISCC-UNIT bodies (models.py:138)


def splitmix64(x):
    # type: (np.ndarray) -> np.ndarray
    """Vectorised splitmix64 finaliser (bijective on uint64)."""
    with np.errstate(over="ignore"):
        z = (np.asarray(x, dtype=_U) + _U(0x9E3779B97F4A7C15)).astype(_U)
        z = ((z ^ (z >> _U(30))) * _U(0xBF58476D1CE4E5B9)).astype(_U)
        z = ((z ^ (z >> _U(27))) * _U(0x94D049BB133111EB)).astype(_U)
        return (z ^ (z >> _U(31))).astype(_U)


def make_keys(start, n, seed=0):
    # type: (int, int, int) -> np.ndarray
    """Unique uint64 keys for rows start..start+n-1 (bijective mix, so no duplicates)."""
    i = np.arange(start, start + n, dtype=_U)
    return splitmix64(i ^ _U((seed * 0x51ED27 + 0xA5A5A5A5) & 0xFFFFFFFFFFFFFFFF))


def make_lengths(start, n, seed=0, lengths=STANDARD_LENGTHS):
    # type: (int, int, int, tuple) -> np.ndarray
    """Per-row byte length, uniform over `lengths` (25 % each of 64/128/192/256 bits by default)."""
    i = np.arange(start, start + n, dtype=_U)
    r = splitmix64(i ^ _U((seed + 0x1234567) & 0xFFFFFFFFFFFFFFFF))
    return np.asarray(lengths, dtype=np.uint8)[(r % _U(len(lengths))).astype(np.int64)]


def make_keys128(start, n, seed=0, chunks_per_asset=64):
    # type: (int, int, int, int) -> tuple[np.ndarray, np.ndarray]
    """
    Simprint chunk pointers `asset8 | offset4 | size4` (lmdb_ops.py:30-49) as (hi, lo) uint64 halves of the big-endian
    16-byte key: row i is chunk `i % chunks_per_asset` of asset `i // chunks_per_asset`; asset ids are a bijective mix
    (unique, unsorted), offsets are multiples of 4096, size 4096.
    """
    i = np.arange(start, start + n, dtype=_U)
    cpa = _U(chunks_per_asset)
    hi = splitmix64((i // cpa) ^ _U((seed * 0x51ED27 + 0xA5A5A5A5) & 0xFFFFFFFFFFFFFFFF))
    lo = (((i % cpa) * _U(4096)) << _U(32)) | _U(4096)
    return hi, lo


def keys128_bytes(hi, lo):
    # type: (np.ndarray, np.ndarray) -> np.ndarray
    """(hi, lo) halves -> uint8[n, 16] big-endian rows, the form 16-byte keys cross the C ABI in."""
    out = np.empty((len(hi), 16), dtype=np.uint8)
    out[:, :8] = np.asarray(hi, dtype=_U).astype(">u8").view(np.uint8).reshape(-1, 8)
    out[:, 8:] = np.asarray(lo, dtype=_U).astype(">u8").view(np.uint8).reshape(-1, 8)
    return out


def make_codes(start, n, seed=0, lens=None, dup_every=0, dup_back=0):
    # type: (int, int, int, np.ndarray|None, int, int) -> np.ndarray
    """
    uint8[n, 32] codes, zero padded beyond each row's length (all 32 bytes if `lens` is None).
    `dup_every` > 0: row i with i % dup_every == dup_every - 1 (and i >= dup_back) repeats the words of row
    i - dup_back, so exact duplicates exist (chunks shared between assets, for the simprint equality join).
    """
    i = np.arange(start, start + n, dtype=_U)
    if dup_every:
        i = np.where((i % _U(dup_every) == _U(dup_every - 1)) & (i >= _U(dup_back)), i - _U(dup_back), i)
    words = np.empty((n, 4), dtype=_U)
    for w in range(4):
        words[:, w] = splitmix64(_U(seed) ^ (i * _U(4) + _U(w)))
    codes = words.view(np.uint8).reshape(n, 32).copy()
    if lens is not None:
        col = np.arange(32, dtype=np.uint8)[None, :]
        codes[col >= np.asarray(lens, dtype=np.uint8)[:, None]] = 0
    return codes


def make_queries(q, n_rows, seed=0, data_seed=0, lengths=STANDARD_LENGTHS, row_lengths=STANDARD_LENGTHS, mixed_rows=True):
    # type: (int, int, int, int, tuple, tuple, bool) -> tuple[np.ndarray, np.ndarray]
    """
    -> (queries uint8[q,32] zero padded, qlens uint8[q]).

    90 %: a stored row (regenerated from `data_seed`) with r in {0,1,2,4,8,16,32}*(L/8) random bit
    flips inside the query length; 10 %: pure random. Query lengths are uniform over `lengths`.
    """
    rng = np.random.default_rng(seed + 7919)
    qlens = np.asarray(lengths, dtype=np.uint8)[rng.integers(0, len(lengths), size=q)]
    rows = rng.integers(0, max(n_rows, 1), size=q)
    queries = np.zeros((q, 32), dtype=np.uint8)
    flips_per_8 = np.array([0, 1, 2, 4, 8, 16, 32])
    for j in range(q):
        L = int(qlens[j])
        if rng.random() < 0.1 or n_rows == 0:
            queries[j, :L] = rng.integers(0, 256, size=L, dtype=np.uint8)
            continue
        base = make_codes(int(rows[j]), 1, data_seed)[0]
        if mixed_rows:
            rl = int(make_lengths(int(rows[j]), 1, data_seed, row_lengths)[0])
            base[rl:] = rng.integers(0, 256, size=32 - rl, dtype=np.uint8)  # beyond the stored prefix: free bits
        bits = np.unpackbits(base[:L])
        nflip = min(int(flips_per_8[rng.integers(0, len(flips_per_8))]) * (L // 8), 8 * L)
        if nflip:
            pos = rng.choice(8 * L, size=nflip, replace=False)
            bits[pos] ^= 1
        queries[j, :L] = np.packbits(bits)
    return queries, qlens
