"""
ORACLE - TEST INFRASTRUCTURE ONLY (see oracle/nphd_oracle.py header). ctypes binding of
oracle/build/libnphd_oracle.so, the OpenMP C restatement of the reference's exact CPU search.
Used by tests (bigger parity cases than the numpy oracle finishes in seconds) and by bench.py's
`cpu_baseline` / `--impl reference` legs. Never imported from `iscc_search_b200/`.
"""

import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "build" / "libnphd_oracle.so"
_lib = None


def build(force=False):
    # type: (bool) -> Path
    """Compile the C oracle with the committed Makefile (gcc is present here and on the GPU box)."""
    if force or not _SO.exists() or _SO.stat().st_mtime < (_HERE / "nphd_oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(str(_SO))
        u8p, u16p, u32p, u64p, i64p = (ctypes.POINTER(t) for t in (ctypes.c_uint8, ctypes.c_uint16, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int64))
        L.oracle_nphd_topk.restype = ctypes.c_int
        L.oracle_nphd_topk.argtypes = [u8p, u8p, u64p, u64p, ctypes.c_size_t, u8p, u8p, ctypes.c_size_t, ctypes.c_uint32,
                                       ctypes.c_uint32, ctypes.c_uint32, i64p, u16p, u16p, u32p, ctypes.c_int]
        L.oracle_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t)) if a is not None else None


def topk(keys_hi, keys_lo, codes, lens, queries, qlens, k, max_h_over_n=None, n_threads=0):
    # type: (...) -> tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]
    """
    Same contract as `oracle.nphd_oracle.topk`, array form.

    :return: (rows int64[Q,k] (-1 padded), h uint16[Q,k], nbits uint16[Q,k], counts uint32[Q])
    """
    if k < 1:
        raise ValueError("`count` must be >= 1")
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    lens = np.ascontiguousarray(lens, dtype=np.uint8)
    keys_hi = np.ascontiguousarray(keys_hi, dtype=np.uint64)
    keys_lo = None if keys_lo is None else np.ascontiguousarray(keys_lo, dtype=np.uint64)
    queries = np.ascontiguousarray(queries, dtype=np.uint8)
    qlens = np.ascontiguousarray(qlens, dtype=np.uint8)
    n, q = len(lens), len(qlens)
    assert codes.shape == (n, 32) and queries.shape == (q, 32)
    rows = np.empty((q, k), dtype=np.int64)
    h = np.empty((q, k), dtype=np.uint16)
    nb = np.empty((q, k), dtype=np.uint16)
    counts = np.empty(q, dtype=np.uint32)
    tn, td = (0, 0) if max_h_over_n is None else max_h_over_n
    rc = lib().oracle_nphd_topk(_p(codes, ctypes.c_uint8), _p(lens, ctypes.c_uint8), _p(keys_hi, ctypes.c_uint64),
                                _p(keys_lo, ctypes.c_uint64), n, _p(queries, ctypes.c_uint8), _p(qlens, ctypes.c_uint8), q, k,
                                tn, td, _p(rows, ctypes.c_int64), _p(h, ctypes.c_uint16), _p(nb, ctypes.c_uint16),
                                _p(counts, ctypes.c_uint32), n_threads)
    if rc != 0:
        raise ValueError("oracle_nphd_topk: bad arguments")
    return rows, h, nb, counts


def num_threads():
    return int(lib().oracle_num_threads())
