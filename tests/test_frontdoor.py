"""CPU test of the batching front door with a stand-in index (host logic only: queueing, coalescing, per-caller truncation, errors)."""

import threading

import numpy as np
import pytest

from iscc_search_b200.frontdoor import BatchingFrontDoor
from iscc_search_b200.matches import BatchMatches, Matches


class FakeIndex:
    """Deterministic 'search': result keys encode the query's first byte; records the batch sizes it saw."""

    def __init__(self):
        self.batch_sizes = []

    def search(self, vectors, count=10):
        self.batch_sizes.append(len(vectors))
        if any(bytes(v)[0] == 0xEE for v in vectors):
            raise RuntimeError("boom")
        keys = np.array([[bytes(v)[0] * 1000 + j for j in range(count)] for v in vectors], dtype=np.uint64)
        dist = np.tile(np.arange(count, dtype=np.float32), (len(vectors), 1))
        if len(vectors) == 1:
            return Matches(keys=keys[0], distances=dist[0])
        return BatchMatches(keys=keys, distances=dist, counts=np.full(len(vectors), count, dtype=np.int64))


def test_concurrent_requests_are_coalesced_and_routed_back():
    idx = FakeIndex()
    door = BatchingFrontDoor(idx, max_batch=64, max_delay_ms=50)
    out = {}

    def worker(i):
        out[i] = door.search(bytes([i]) + b"\x00" * 7, count=3 + (i % 4))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(40)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    door.close()
    assert door.requests == 40 and door.batches < 40 and max(idx.batch_sizes) > 1
    for i in range(40):
        m = out[i]
        assert len(m) == 3 + (i % 4) and m.keys.tolist() == [i * 1000 + j for j in range(3 + (i % 4))]


def test_single_request_and_error_propagation():
    idx = FakeIndex()
    door = BatchingFrontDoor(idx, max_batch=8, max_delay_ms=1)
    m = door.search(b"\x07" * 8, count=2)
    assert isinstance(m, Matches) and m.keys.tolist() == [7000, 7001]
    with pytest.raises(RuntimeError, match="boom"):
        door.search(b"\xee" * 8, count=1)
    with pytest.raises(ValueError):
        door.search(b"\x01" * 8, count=0)
    door.close()
    with pytest.raises(ValueError, match="closed"):
        door.search(b"\x01" * 8, count=1)
