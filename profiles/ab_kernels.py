"""A/B timing of scan-kernel variants: one process per library, single-length stores isolate each k_scan<WE,G>.

    python profiles/ab_kernels.py profiles/ab/A_old.so profiles/ab/B_new.so [--rows 10000000]
"""
import argparse
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def child(lib_path, rows, queries, k, reps, mixed_only=False):
    sys.path.insert(0, str(ROOT))
    import numpy as np

    from iscc_search_b200 import _lib, synth

    _lib.SO_PATH = Path(lib_path).resolve()
    out = []
    for L in (() if mixed_only else (8, 16, 24, 32)):
        st = _lib.Store(key_bytes=8, max_bytes=32)
        st.set_profiling(True)
        for c0 in range(0, rows, 2_000_000):
            cn = min(2_000_000, rows - c0)
            lens = synth.make_lengths(c0, cn, 1, (L,))
            st.add(synth.make_keys(c0, cn, 1), synth.make_codes(c0, cn, 1, lens), lens)
        q, ql = synth.make_queries(queries, rows, 2, 1, (L,), (L,))
        ms = []
        for _ in range(reps + 2):
            st.search(q, ql, k)
            ms.append(st.stats()["scan_ms"])
        best = float(np.median(ms[2:]))
        out.append(f"WE={L // 4}: {best:7.3f} ms {rows * queries / best / 1e9:6.1f} Gpairs/s")
        st.close()
    # the bench mix
    st = _lib.Store(key_bytes=8, max_bytes=32)
    st.set_profiling(True)
    for c0 in range(0, rows * 2, 2_000_000):
        lens = synth.make_lengths(c0, 2_000_000, 1)
        st.add(synth.make_keys(c0, 2_000_000, 1), synth.make_codes(c0, 2_000_000, 1, lens), lens)
    q, ql = synth.make_queries(queries, rows * 2, 2, 1)
    ms = []
    for _ in range(reps + 2):
        st.search(q, ql, k)
        ms.append(st.stats()["scan_ms"])
    out.append(f"mixed {2 * rows / 1e6:.0f}M: {float(np.median(ms[2:])):7.3f} ms")
    print(f"{Path(lib_path).name:14s} " + " | ".join(out), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="+")
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--queries", type=int, default=1024)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--mixed-only", action="store_true", help="only the mixed-length store (2 x rows)")
    a = ap.parse_args()
    if a.child:
        child(a.libs[0], a.rows, a.queries, a.k, a.reps, a.mixed_only)
    else:
        for lib in a.libs:
            subprocess.run([sys.executable, __file__, lib, "--child", "--rows", str(a.rows), "--queries", str(a.queries),
                            "--k", str(a.k), "--reps", str(a.reps)] + (["--mixed-only"] if a.mixed_only else []), check=False)
