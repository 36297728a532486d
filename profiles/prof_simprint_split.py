"""f-2 split: B200SimprintIndex.search_raw at the cfg4 shape (S query simprints, count = 2*limit*oversampling) - GPU search
time vs host scoring time.   python profiles/prof_simprint_split.py --rows 100000000 --simprints 256 --limit 100"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from bench import CONFIGS, build_store  # noqa: E402
from iscc_search_b200 import simprint as sp  # noqa: E402
from iscc_search_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=100_000_000)
ap.add_argument("--simprints", type=int, default=256)
ap.add_argument("--limit", type=int, default=100)
ap.add_argument("--threshold", type=float, default=0.75)
args = ap.parse_args()
cfg = dict(CONFIGS["cfg4"], rows=args.rows)
dev = torch.device("cuda", 0)
idx = sp.B200SimprintIndex(None, ndim=64)
store = idx._index._store
store.set_stream(torch.cuda.current_stream(dev).cuda_stream)
build_store(store, cfg, 1, 0, args.rows, dev, torch)
store.set_stream(None)
queries, qlens = synth.make_queries(args.simprints, args.rows, 2, 1, lengths=(8,), row_lengths=(8,), mixed_rows=False)
simprints = [bytes(q[:8]) for q in queries]
for rep in range(3):
    t0 = time.perf_counter()
    res = idx.search_raw(simprints, limit=2 * args.limit, threshold=args.threshold, detailed=True, doc_freq_fn=None, total_assets=args.rows // 64)
    t_all = time.perf_counter() - t0
    prof = getattr(idx, "last_profile", {})
    print(f"rep {rep}: search_raw {t_all * 1e3:.1f} ms total; gpu search {prof.get('gpu_ms', float('nan')):.1f} ms, host scoring {prof.get('score_ms', float('nan')):.1f} ms; "
          f"{len(res)} assets, best {res[0].score if res else None}")
