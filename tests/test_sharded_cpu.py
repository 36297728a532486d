"""CPU tests (no GPU) of the multi-rank host logic: key placement, packed record layout, gloo all-gather plumbing (world_size 2)."""

import os
import subprocess
import sys
from pathlib import Path

import numpy as np

from iscc_search_b200 import synth
from iscc_search_b200.sharded import owner_of, record_layout

ROOT = Path(__file__).resolve().parent.parent


def test_owner_of_is_balanced_stable_and_total():
    keys = synth.make_keys(0, 200_000, 3)
    for world in (2, 4, 8):
        own = owner_of(keys, world)
        counts = np.bincount(own, minlength=world)
        assert counts.sum() == len(keys) and own.min() >= 0 and own.max() < world
        assert counts.max() / counts.min() < 1.05
        assert np.array_equal(own, owner_of(keys.copy(), world))  # stable: pure function of the key


def test_record_layout_is_aligned_and_disjoint():
    for q, k in ((1, 1), (5, 7), (1024, 100)):
        off, size = record_layout(q, k)
        qk = q * k
        assert off["khi"] % 8 == 0 and off["klo"] % 8 == 0 and off["h"] % 2 == 0 and off["n"] % 2 == 0 and off["cnt"] % 4 == 0
        assert size % 16 == 0 and size >= off["cnt"] + q * 4
        assert off["klo"] == off["khi"] + qk * 8 and off["h"] == off["klo"] + qk * 8


def test_gloo_world2_all_gather_delivers_rank_major_packed_buffers(tmp_path):
    out = tmp_path / "g.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", str(ROOT / "tests" / "sharded_worker.py"), "--backend", "gloo", "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=str(ROOT), env=dict(os.environ))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    d = np.load(out)
    size, world, q, k = int(d["size"]), int(d["world"]), int(d["q"]), int(d["k"])
    off, size2 = record_layout(q, k)
    assert size == size2 and world == 2
    g = d["gathered"]
    for rank in range(world):
        raw = g[rank * size:(rank + 1) * size]
        khi = raw[off["khi"]: off["khi"] + q * k * 8].view(np.uint64)
        assert np.array_equal(khi, np.arange(q * k, dtype=np.uint64) * 10 + rank)
        assert (raw[off["h"]: off["h"] + q * k * 2].view(np.uint16) == rank + 1).all()
        assert (raw[off["cnt"]: off["cnt"] + q * 4].view(np.uint32) == k - rank).all()
    # the length-mask union every rank derives before a search with shared thresholds (same on all ranks)
    want = (1 << 7) | (1 << 15) | (1 << 4)
    assert d["gmasks"].tolist() == [want] * world
