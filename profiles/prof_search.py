"""Small driver for ncu: build a mixed-length store and run a batched search and single-query scans.

    python profiles/prof_search.py --rows 20000000 --queries 1024 --reps 2
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from iscc_search_b200 import _lib, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=20_000_000)
ap.add_argument("--queries", type=int, default=1024)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--lengths", default="8,16,24,32")
ap.add_argument("--qlengths", default="8,16,24,32")
ap.add_argument("--single", type=int, default=1)
args = ap.parse_args()
lengths = tuple(int(x) for x in args.lengths.split(","))
qlengths = tuple(int(x) for x in args.qlengths.split(","))

st = _lib.Store(key_bytes=8, max_bytes=32)
st.set_profiling(True)
t0 = time.time()
for c0 in range(0, args.rows, 2_000_000):
    cn = min(2_000_000, args.rows - c0)
    lens = synth.make_lengths(c0, cn, 1, lengths)
    st.add(synth.make_keys(c0, cn, 1), synth.make_codes(c0, cn, 1, lens), lens)
print(f"built {st.size()} rows in {time.time() - t0:.1f}s")
queries, qlens = synth.make_queries(args.queries, args.rows, 2, 1, qlengths, lengths)
for r in range(args.reps):
    st.search(queries, qlens, args.k)
    s = st.stats()
    print(f"batch rep {r}: total {s['total_ms']:.3f} ms scan {s['scan_ms']:.3f} ms select {s['select_ms']:.3f} ms launches {s['kernel_launches']} "
          f"cands/query {s['candidates'] / args.queries:.0f} popc/s {s['algo_popc'] / max(s['scan_ms'], 1e-9) / 1e9:.1f} G fallback {s['fallback_queries']}")
if args.single:
    for L in qlengths:
        q1 = np.zeros((1, 32), dtype=np.uint8)
        q1[0, :L] = np.arange(L, dtype=np.uint8) * 37 + 11
        for r in range(args.reps + 1):
            st.search(q1, np.array([L], dtype=np.uint8), args.k)
        s = st.stats()
        print(f"single {8 * L}-bit: total {s['total_ms']:.3f} ms scan {s['scan_ms']:.3f} ms bytes {s['algo_bytes']} "
              f"-> {s['algo_bytes'] / max(s['scan_ms'], 1e-9) / 1e6:.0f} GB/s launches {s['kernel_launches']} cands {s['candidates']}")
st.close()
