"""ncu driver for the batch regime: cfg3-shaped store built on the device, batches of mixed-length queries.

    python profiles/prof_batch.py --rows 20000000 --queries 10000 --reps 2 [--single 1]
"""
import argparse
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from bench import CONFIGS, build_store  # noqa: E402
from iscc_search_b200 import _lib, synth  # noqa: E402
from iscc_search_b200.sharded import ShardedSearcher  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=20_000_000)
ap.add_argument("--queries", type=int, default=10_000)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--config", default="cfg3")
ap.add_argument("--single", type=int, default=0, help="also run single-query searches (small-batch kernel)")
args = ap.parse_args()
cfg = dict(CONFIGS[args.config], rows=args.rows, queries=args.queries)
dev = torch.device("cuda", 0)
st = _lib.Store(key_bytes=cfg["key_bytes"], max_bytes=cfg["max_bytes"], fixed_len=cfg["fixed_len"])
st.set_profiling(True)
st.set_stream(torch.cuda.current_stream(dev).cuda_stream)
build_store(st, cfg, 1, 0, args.rows, dev, torch)
mixed = len(cfg["lengths"]) > 1
queries, qlens = synth.make_queries(args.queries, args.rows, 2, 1, lengths=cfg["lengths"], row_lengths=cfg["lengths"], mixed_rows=mixed)
searcher = ShardedSearcher(st, 0, 1, None, dev)
dq = torch.from_numpy(queries).to(dev)
for r in range(args.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    searcher.search_device_batched(dq, qlens, args.k, cfg["thr"])
    e1.record()
    torch.cuda.synchronize()
    s = st.stats()
    ms = e0.elapsed_time(e1)
    print(f"batch rep {r}: {ms:.2f} ms, {args.queries * args.rows / ms / 1e9:.1f} Gpairs/s, launches {s['kernel_launches']} (last chunk), cands/query {s['candidates'] / max(1, min(args.queries, 16384)):.0f}")
if args.single:
    for L in cfg["lengths"]:
        i = int(np.flatnonzero(qlens == L)[0])
        d1 = dq[i:i + 1].contiguous()
        for r in range(3):
            searcher.search_device(d1, np.ascontiguousarray(qlens[i:i + 1]), min(args.k, 100))
        s = st.stats()
        print(f"single {8 * L}-bit: search {s['total_ms']:.4f} ms, {s['algo_bytes'] / s['total_ms'] / 1e6:.0f} GB/s, launches {s['kernel_launches']}")
