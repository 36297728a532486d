"""Config-4-shaped check at single-GPU scale: N 64-bit simprints with 128-bit composite keys, one query asset of S
simprints, count = 4000 (limit 100 x 2 x oversampling 20), threshold h <= 16, plus the equality join (h = 0, cap 1000).

    python profiles/prof_simprint.py --rows 100000000 --simprints 256
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from iscc_search_b200 import _lib, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=100_000_000)
ap.add_argument("--simprints", type=int, default=256)
ap.add_argument("--count", type=int, default=4000)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()

st = _lib.Store(key_bytes=16, max_bytes=8, fixed_len=8)
st.set_profiling(True)
t0 = time.time()
CH = 4_000_000
for c0 in range(0, args.rows, CH):
    cn = min(CH, args.rows - c0)
    lens = np.full(cn, 8, dtype=np.uint8)
    codes = synth.make_codes(c0, cn, 4, lens)
    # every 16th row repeats the simprint 64 rows earlier (chunks shared between assets -> exact duplicates exist)
    dup = np.arange(cn) % 16 == 15
    codes[dup] = codes[np.maximum(np.nonzero(dup)[0] - 64, 0)]
    keys = np.zeros((cn, 16), dtype=np.uint8)
    keys[:, :8] = (np.arange(c0, c0 + cn, dtype=np.uint64) // np.uint64(64)).astype(">u8").view(np.uint8).reshape(cn, 8)   # 64 chunks per asset
    keys[:, 8:12] = ((np.arange(c0, c0 + cn, dtype=np.uint64) % np.uint64(64)) * np.uint64(4096)).astype(">u4").view(np.uint8).reshape(cn, 4)
    keys[:, 12:16] = np.array([0, 0, 16, 0], dtype=np.uint8)
    st.add(keys, codes, lens)
print(f"built {st.size()} simprints in {time.time() - t0:.1f}s, device bytes {st.device_bytes() / 1e9:.2f} GB")
S = args.simprints
rows = np.random.default_rng(1).integers(0, args.rows, size=S)
q = np.concatenate([synth.make_codes(int(r), 1, 4, np.array([8], dtype=np.uint8)) for r in rows])
ql = np.full(S, 8, dtype=np.uint8)
for mode, thr, k in (("threshold h<=16", (16, 64), args.count), ("equality h=0", (0, 64), 1000), ("no threshold", None, args.count)):
    for r in range(args.reps):
        t1 = time.time()
        keys, h, nb, cnt, _ = st.search(q, ql, k, thr)
        wall = (time.time() - t1) * 1e3
    s = st.stats()
    print(f"{mode:16s} k={k}: wall {wall:.2f} ms gpu total {s['total_ms']:.2f} scan {s['scan_ms']:.2f} select {s['select_ms']:.2f} "
          f"mean results {cnt.mean():.0f} cands/query {s['candidates'] / S:.0f} fallback {s['fallback_queries']} launches {s['kernel_launches']}")
st.close()
