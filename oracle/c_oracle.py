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
        u32, u64 = ctypes.c_uint32, ctypes.c_uint64
        L.oracle_synth_rows.restype = ctypes.c_int
        L.oracle_synth_rows.argtypes = [u64, u64, ctypes.c_size_t, u8p, u32, u32, u32, u32, u32, u64p, u64p, u8p, u8p, ctypes.c_int]
        L.oracle_synth_topk.restype = ctypes.c_int
        L.oracle_synth_topk.argtypes = [u64, u64, u8p, u32, u32, u32, u32, u32, u8p, u8p, ctypes.c_size_t, u32, u32, u32,
                                        u64p, u64p, u16p, u16p, u32p, ctypes.c_int]
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


def synth_rows(start, n, seed, lengths=(8, 16, 24, 32), key_mode=0, cpa=64, dup_every=0, dup_back=0, n_threads=0):
    # type: (...) -> tuple[np.ndarray, np.ndarray|None, np.ndarray, np.ndarray]
    """C copy of iscc_search_b200.synth (make_keys / make_keys128, make_lengths, make_codes): -> (khi, klo|None, codes, lens)."""
    lengths = np.ascontiguousarray(lengths, dtype=np.uint8)
    khi = np.empty(n, dtype=np.uint64)
    klo = np.empty(n, dtype=np.uint64) if key_mode else None
    codes = np.empty((n, 32), dtype=np.uint8)
    lens = np.empty(n, dtype=np.uint8)
    rc = lib().oracle_synth_rows(seed, start, n, _p(lengths, ctypes.c_uint8), len(lengths), key_mode, cpa, dup_every, dup_back,
                                 _p(khi, ctypes.c_uint64), _p(klo, ctypes.c_uint64), _p(codes, ctypes.c_uint8),
                                 _p(lens, ctypes.c_uint8), n_threads)
    if rc != 0:
        raise ValueError("oracle_synth_rows: bad arguments")
    return khi, klo, codes, lens


def synth_topk(n_rows, seed, queries, qlens, k, lengths=(8, 16, 24, 32), key_mode=0, cpa=64, dup_every=0, dup_back=0,
               max_h_over_n=None, n_threads=0):
    # type: (...) -> tuple[np.ndarray, np.ndarray|None, np.ndarray, np.ndarray, np.ndarray]
    """
    Exact top-k over `n_rows` synthetic rows regenerated block by block (nothing materialised: works at 1 B rows).
    -> (keys_hi uint64[Q,k], keys_lo|None, h, nbits, counts); unused slots: key = 2**64-1, h = nbits = 0.
    """
    if k < 1:
        raise ValueError("`count` must be >= 1")
    lengths = np.ascontiguousarray(lengths, dtype=np.uint8)
    queries = np.ascontiguousarray(queries, dtype=np.uint8)
    qlens = np.ascontiguousarray(qlens, dtype=np.uint8)
    q = len(qlens)
    khi = np.empty((q, k), dtype=np.uint64)
    klo = np.empty((q, k), dtype=np.uint64) if key_mode else None
    h = np.empty((q, k), dtype=np.uint16)
    nb = np.empty((q, k), dtype=np.uint16)
    counts = np.empty(q, dtype=np.uint32)
    tn, td = (0, 0) if max_h_over_n is None else max_h_over_n
    rc = lib().oracle_synth_topk(seed, n_rows, _p(lengths, ctypes.c_uint8), len(lengths), key_mode, cpa, dup_every, dup_back,
                                 _p(queries, ctypes.c_uint8), _p(qlens, ctypes.c_uint8), q, k, tn, td, _p(khi, ctypes.c_uint64),
                                 _p(klo, ctypes.c_uint64), _p(h, ctypes.c_uint16), _p(nb, ctypes.c_uint16),
                                 _p(counts, ctypes.c_uint32), n_threads)
    if rc != 0:
        raise ValueError("oracle_synth_topk: bad arguments")
    return khi, klo, h, nb, counts


class SoaStore:
    """
    The tuned CPU arm (oracle_soa_*): length-bucketed word planes + AVX-512 VPOPCNTDQ when the host CPU has it.
    Same results as `topk`; used by bench.py's cpu_baseline / --impl reference legs and checked in tests/test_oracle.py.
    """

    def __init__(self, keys_hi, keys_lo, codes, lens):
        L = lib()
        L.oracle_soa_build.restype = ctypes.c_void_p
        L.oracle_soa_build.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_size_t]
        L.oracle_soa_free.argtypes = [ctypes.c_void_p]
        L.oracle_soa_topk.restype = ctypes.c_int
        L.oracle_soa_topk.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_uint32,
                                      ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        lens = np.ascontiguousarray(lens, dtype=np.uint8)
        keys_hi = np.ascontiguousarray(keys_hi, dtype=np.uint64)
        keys_lo = None if keys_lo is None else np.ascontiguousarray(keys_lo, dtype=np.uint64)
        self._h = L.oracle_soa_build(codes.ctypes.data, lens.ctypes.data, keys_hi.ctypes.data, None if keys_lo is None else keys_lo.ctypes.data, len(lens))
        if not self._h:
            raise ValueError("oracle_soa_build failed")

    @staticmethod
    def isa():
        return "avx512-vpopcntdq" if lib().oracle_soa_isa() else "scalar popcnt"

    def topk(self, queries, qlens, k, max_h_over_n=None, n_threads=0, force_scalar=False):
        queries = np.ascontiguousarray(queries, dtype=np.uint8)
        qlens = np.ascontiguousarray(qlens, dtype=np.uint8)
        q = len(qlens)
        rows = np.empty((q, k), dtype=np.int64)
        h = np.empty((q, k), dtype=np.uint16)
        nb = np.empty((q, k), dtype=np.uint16)
        counts = np.empty(q, dtype=np.uint32)
        tn, td = (0, 0) if max_h_over_n is None else max_h_over_n
        rc = lib().oracle_soa_topk(self._h, queries.ctypes.data, qlens.ctypes.data, q, k, tn, td, rows.ctypes.data, h.ctypes.data,
                                   nb.ctypes.data, counts.ctypes.data, n_threads, 1 if force_scalar else 0)
        if rc != 0:
            raise ValueError("oracle_soa_topk: bad arguments")
        return rows, h, nb, counts

    def close(self):
        if self._h:
            lib().oracle_soa_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def num_threads():
    return int(lib().oracle_num_threads())
