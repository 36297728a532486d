"""
Row-sharded exact search over the GPUs of one box: one process per GPU (`torch.distributed`, NCCL over
NVLink), every rank scans its own rows, local top-k records are all-gathered and merged on the
device with the same total order, so the result is independent of the number of shards
(SURVEY.md section 8e). The reference has no multi-GPU path (manager.py:43-46) - this is new design
behind the same `search` call shape (index.py:2037).

torch is plumbing here: device buffers, the current stream and the NCCL all-gather. All arithmetic
is in libisx_b200.so.
"""

import ctypes

import numpy as np

from iscc_search_b200 import _lib


def owner_of(keys, world):
    # type: (np.ndarray, int) -> np.ndarray
    """Rank that stores each uint64 key: a fixed bijective mix, so placement is balanced and stable."""
    from iscc_search_b200.synth import splitmix64

    return (splitmix64(np.asarray(keys, dtype=np.uint64)) % np.uint64(world)).astype(np.int64)


def global_length_mask(local_mask, dist, group=None, device=None, flags=None):
    # type: (int, object, object, object, object) -> int
    """
    Union over all ranks of a store's length mask (bit L-1 = some code of L bytes is stored). NCCL has no bitwise
    reduction, so 32 presence flags are MAX-reduced. Collective call; on the device path it is also the barrier
    between zeroing the shared histograms and the first remote count (see `ShardedSearcher.search_device`).
    """
    import torch

    host = torch.tensor([(local_mask >> b) & 1 for b in range(32)], dtype=torch.int32)
    if flags is None:
        flags = torch.empty(32, dtype=torch.int32, device=device if device is not None else "cpu")
    flags.copy_(host, non_blocking=False)
    dist.all_reduce(flags, op=dist.ReduceOp.MAX, group=group)
    out = flags.cpu().tolist()
    return sum(1 << b for b in range(32) if out[b])


def record_layout(q, k):
    # type: (int, int) -> tuple[dict, int]
    """Byte offsets of the packed per-rank result buffer {khi, klo, h, n, cnt} and its 16-byte aligned size."""
    qk = q * k
    off = {"khi": 0, "klo": qk * 8, "h": qk * 16, "n": qk * 18, "cnt": qk * 20}
    size = qk * 20 + q * 4
    return off, (size + 15) // 16 * 16


class ShardedSearcher:
    """Search front of one rank's store; `search` is a collective call (all ranks, same arguments)."""

    def __init__(self, store, rank=0, world=1, group=None, device=None, share_thresholds=True, max_queries=16384):
        import torch

        self.torch = torch
        self.store = store
        self.rank, self.world, self.group = rank, world, group
        self.device = torch.device("cuda", store.device) if device is None else device
        self._bufs = {}
        self.shared = False
        self.max_queries = max_queries
        self.profile = False      # record CUDA events around every all-gather (see `gather_ms`)
        self._gather_events = []
        if world > 1 and share_thresholds:
            self._init_sharing(max_queries)

    def _init_sharing(self, max_queries):
        """Exchange CUDA IPC handles of the per-rank home histograms (include/isx.h: isx_share_*), all ranks call this."""
        torch, dist = self.torch, self.torch.distributed
        handle = (ctypes.c_ubyte * 64)()
        _lib.check(_lib.lib().isx_share_init(self.store.handle, self.world, self.rank, max_queries, handle))
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
        every = torch.empty(64 * self.world, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(every, mine, group=self.group)
        every = every.cpu().numpy()
        for r in range(self.world):
            if r != self.rank:
                peer = (ctypes.c_ubyte * 64)(*every[64 * r: 64 * (r + 1)].tolist())
                _lib.check(_lib.lib().isx_share_attach(self.store.handle, r, peer))
        self._flags = torch.zeros(32, dtype=torch.int32, device=self.device)
        dist.barrier(group=self.group)
        self.shared = True

    def _buf(self, name, nbytes):
        t = self._bufs.get(name)
        if t is None or t.numel() < nbytes:
            t = self.torch.empty(max(nbytes, 256), dtype=self.torch.uint8, device=self.device)
            self._bufs[name] = t
        return t

    def search_device(self, d_queries, qlens, k, thr=None):
        # type: (object, np.ndarray, int, tuple|None) -> tuple
        """
        Queries already on the device (uint8[Q,32], zero padded). Returns the merged result as a device
        byte buffer in `record_layout(Q, k)` form plus the layout; nothing is copied to the host.
        """
        torch = self.torch
        q = len(qlens)
        off, size = record_layout(q, k)
        local = self._buf("local", size)
        base = local.data_ptr()
        tn, td = (0, 0) if thr is None else thr
        L = _lib.lib()
        self.store.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        if self.shared:
            # zero the home histograms, then a collective on the same stream: no rank starts emitting before every
            # rank has zeroed (the all-gather at the end of the previous batch already fenced its emissions). The
            # collective carries the stored-length masks: all ranks must rank distances over the SAME length classes,
            # also when a shard lacks a bucket (a rank table built from local rows only would index the peers'
            # histograms at incompatible positions).
            _lib.check(L.isx_share_reset(self.store.handle))
            mask = global_length_mask(self.store.length_mask(), self.torch.distributed, self.group, self.device, self._flags)
            _lib.check(L.isx_share_set_lengths(self.store.handle, mask))
        _lib.check(L.isx_search_device(self.store.handle, d_queries.data_ptr(), 1, _lib.ptr(qlens), q, k, tn, td,
                                       base + off["khi"], base + off["klo"], base + off["h"], base + off["n"],
                                       base + off["cnt"], 0))
        if self.world == 1:
            return local, off, size
        gathered = self._buf("gathered", size * self.world)
        if self.profile:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        self.torch.distributed.all_gather_into_tensor(gathered[: size * self.world], local[:size], group=self.group)
        if self.profile:
            e1.record()
            self._gather_events.append((e0, e1, size * self.world))
        merged = self._buf("merged", size)
        g0, m0 = gathered.data_ptr(), merged.data_ptr()
        _lib.check(L.isx_merge_device(self.store.handle, self.world, q, k, size, g0 + off["khi"], g0 + off["klo"],
                                      g0 + off["h"], g0 + off["n"], g0 + off["cnt"], m0 + off["khi"], m0 + off["klo"],
                                      m0 + off["h"], m0 + off["n"], m0 + off["cnt"], 0))
        return merged, off, size

    def gather_ms(self):
        # type: () -> tuple[float, int]
        """(CUDA-event milliseconds, bytes received) summed over the all-gathers since the last call (needs `profile`)."""
        self.torch.cuda.current_stream(self.device).synchronize()
        ms = sum(e0.elapsed_time(e1) for e0, e1, _ in self._gather_events)
        nbytes = sum(b for _, _, b in self._gather_events)
        self._gather_events = []
        return ms, nbytes

    def search_device_batched(self, d_queries, qlens, k, thr=None, chunk=None, on_chunk=None):
        # type: (object, np.ndarray, int, tuple|None, int|None, object) -> tuple
        """
        `search_device` for batches beyond one exchange: the queries go through in chunks (<= `max_queries`, so the
        thresholds stay shared; and the all-gather buffer stays world x chunk x k records instead of world x Q x k -
        12.8 GB at BASELINE config 5), each chunk's merged records land in one `record_layout(Q, k)` buffer.
        """
        q = len(qlens)
        chunk = min(chunk or self.max_queries, self.max_queries)
        if q <= chunk:
            r = self.search_device(d_queries, qlens, k, thr)
            if on_chunk:
                on_chunk()
            return r
        chunk = -(-q // -(-q // chunk))  # equal chunks
        off, size = record_layout(q, k)
        out = self._buf("batched", size)
        widths = {"khi": 8, "klo": 8, "h": 2, "n": 2}
        for c0 in range(0, q, chunk):
            cn = min(chunk, q - c0)
            buf, coff, _ = self.search_device(d_queries[c0:c0 + cn], np.ascontiguousarray(qlens[c0:c0 + cn]), k, thr)
            for name, w in widths.items():
                out[off[name] + c0 * k * w: off[name] + (c0 + cn) * k * w].copy_(buf[coff[name]: coff[name] + cn * k * w], non_blocking=True)
            out[off["cnt"] + c0 * 4: off["cnt"] + (c0 + cn) * 4].copy_(buf[coff["cnt"]: coff["cnt"] + cn * 4], non_blocking=True)
            if on_chunk:
                on_chunk()
        return out, off, size

    def search(self, queries, qlens, k, thr=None, pinned_in=None, pinned_out=None, with_lo=False):
        # type: (np.ndarray, np.ndarray, int, tuple|None, object, object, bool) -> tuple
        """
        Host in, host out: (keys uint64[Q,k], hamming uint16[Q,k], nbits uint16[Q,k], counts uint32[Q]); with `with_lo`
        (128-bit keys) the low key halves uint64[Q,k] are appended.
        """
        torch = self.torch
        q = len(qlens)
        if pinned_in is None:
            pinned_in = torch.from_numpy(np.ascontiguousarray(queries, dtype=np.uint8))
        d_q = self._buf("queries", q * 32)[: q * 32].view(q, 32)
        d_q.copy_(pinned_in.view(q, 32), non_blocking=True)
        buf, off, size = self.search_device_batched(d_q, qlens, k, thr)
        if pinned_out is None:
            pinned_out = torch.empty(size, dtype=torch.uint8, pin_memory=True)
        pinned_out[:size].copy_(buf[:size], non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        raw = pinned_out[:size].numpy()
        qk = q * k
        keys = raw[off["khi"]: off["khi"] + qk * 8].view(np.uint64).reshape(q, k)
        h = raw[off["h"]: off["h"] + qk * 2].view(np.uint16).reshape(q, k)
        nb = raw[off["n"]: off["n"] + qk * 2].view(np.uint16).reshape(q, k)
        cnt = raw[off["cnt"]: off["cnt"] + q * 4].view(np.uint32)
        if with_lo:
            return keys, h, nb, cnt, raw[off["klo"]: off["klo"] + qk * 8].view(np.uint64).reshape(q, k)
        return keys, h, nb, cnt
