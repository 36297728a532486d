"""
Batched multi-request front door.

The reference answers one unit per `search` call and never batches across requests
(/root/reference/iscc_search/indexes/usearch/index.py:786-806, call stack in SURVEY.md 3.1: FastAPI sync
handlers on a thread pool, one `ShardedNphdIndex.search(query, count=limit)` per unit per request).
On the GPU a scan pass costs the same for 1 or ~3 queries (HBM bound) and little more for dozens,
so concurrent requests are coalesced: callers block in `search`, a single worker thread drains the
queue into ONE batched index call and hands every caller its own `Matches`.

Pure host logic (threads + queue); the index only needs `.search(list_of_vectors, count=) -> BatchMatches|Matches`.
"""

import threading
import time
from collections import deque


class _Request:
    __slots__ = ("vector", "count", "event", "result", "error")

    def __init__(self, vector, count):
        self.vector, self.count = vector, count
        self.event = threading.Event()
        self.result, self.error = None, None


class BatchingFrontDoor:
    """
    Coalesces concurrent single-query searches into batched calls.

    :param index: object with `.search(vectors, count=)`; a batch of one returns a bare Matches (as iscc-usearch does)
    :param max_batch: flush as soon as this many requests wait
    :param max_delay_ms: oldest request waits at most this long for company
    """

    def __init__(self, index, max_batch=256, max_delay_ms=0.5):
        self.index = index
        self.max_batch = max_batch
        self.max_delay = max_delay_ms / 1000.0
        self._q = deque()
        self._cv = threading.Condition()
        self._closed = False
        self.batches = 0
        self.requests = 0
        self._worker = threading.Thread(target=self._run, name="isx-frontdoor", daemon=True)
        self._worker.start()

    def search(self, vector, count=10):
        """Same call shape as `ShardedNphdIndex.search(query, count=limit)`; blocks until the batch it rode in is done."""
        if count < 1:
            raise ValueError("`count` must be >= 1")
        req = _Request(vector, count)
        with self._cv:
            if self._closed:
                raise ValueError("front door is closed")
            self._q.append(req)
            self._cv.notify()
        req.event.wait()
        if req.error is not None:
            raise req.error
        return req.result

    def _take_batch(self):
        with self._cv:
            while not self._q and not self._closed:
                self._cv.wait()
            if not self._q:
                return None
            deadline = time.monotonic() + self.max_delay
            while len(self._q) < self.max_batch and not self._closed:
                left = deadline - time.monotonic()
                if left <= 0:
                    break
                self._cv.wait(left)
            n = min(len(self._q), self.max_batch)
            return [self._q.popleft() for _ in range(n)]

    def _run(self):
        while True:
            batch = self._take_batch()
            if batch is None:
                return
            self.batches += 1
            self.requests += len(batch)
            try:
                count = max(r.count for r in batch)  # one pass with the largest k, each caller gets its own prefix
                res = self.index.search([r.vector for r in batch], count=count)
                for i, r in enumerate(batch):
                    m = res if len(batch) == 1 else res[i]
                    r.result = _truncate(m, r.count)
            except Exception as e:  # every waiter of the failed batch sees the error
                for r in batch:
                    r.error = e
            for r in batch:
                r.event.set()

    def close(self):
        with self._cv:
            self._closed = True
            self._cv.notify_all()
        self._worker.join(timeout=5)


def _truncate(m, count):
    """First `count` entries of a Matches-like object (results are sorted, so a prefix of top-K is top-count)."""
    if len(m) <= count:
        return m
    kw = {}
    for name in ("hamming", "nbits", "vectors"):
        v = getattr(m, name, None)
        kw[name] = None if v is None else v[:count]
    return type(m)(keys=m.keys[:count], distances=m.distances[:count], visited_members=getattr(m, "visited_members", 0),
                   computed_distances=getattr(m, "computed_distances", 0), **kw)
