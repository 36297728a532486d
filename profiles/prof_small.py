"""Single-query / small-batch latency on the cfg3 store (100M mixed rows): library CUDA-event time of the whole search.

    python profiles/prof_small.py [--rows N]          # ISX_SMALL_PATH=0 -> general multi-launch path (A/B)
"""
import argparse
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from bench import CONFIGS, build_store, measured_peaks  # noqa: E402
from iscc_search_b200 import _lib, synth  # noqa: E402
from iscc_search_b200.sharded import ShardedSearcher  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=100_000_000)
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
cfg = dict(CONFIGS["cfg3"], rows=args.rows)
dev = torch.device("cuda", 0)
st = _lib.Store(key_bytes=8, max_bytes=32)
st.set_profiling(True)
st.set_stream(torch.cuda.current_stream(dev).cuda_stream)
build_store(st, cfg, 1, 0, args.rows, dev, torch)
hbm = measured_peaks()[0]
import pynvml  # noqa: E402
pynvml.nvmlInit()
_h = pynvml.nvmlDeviceGetHandleByIndex(0)


def clk():
    return f"sm={pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM)}MHz mem={pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_MEM)}MHz pw={pynvml.nvmlDeviceGetPowerUsage(_h) / 1000:.0f}W reasons={pynvml.nvmlDeviceGetCurrentClocksEventReasons(_h):#x}"

searcher = ShardedSearcher(st, 0, 1, None, dev)
queries, qlens = synth.make_queries(64, args.rows, 2, 1)
print(f"rows={args.rows} ISX_SMALL_PATH={os.environ.get('ISX_SMALL_PATH', '1')} ISX_SMALL_SAMPLE={os.environ.get('ISX_SMALL_SAMPLE', 'default')}")
for name, sel in (("1x64bit", [np.flatnonzero(qlens == 8)[0]]), ("1x128bit", [np.flatnonzero(qlens == 16)[0]]),
                  ("1x192bit", [np.flatnonzero(qlens == 24)[0]]), ("1x256bit", [np.flatnonzero(qlens == 32)[0]]),
                  ("4 mixed", list(range(4))), ("8 mixed", list(range(8))), ("8x64bit", list(np.flatnonzero(qlens == 8)[:8]))):
    q = np.ascontiguousarray(queries[sel])
    ql = np.ascontiguousarray(qlens[sel])
    dq = torch.from_numpy(q).to(dev)
    for _ in range(3):
        searcher.search_device(dq, ql, 100)
    tot, scan, wall = [], [], []
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        searcher.search_device(dq, ql, 100)
        e1.record()
        torch.cuda.synchronize()
        s = st.stats()
        tot.append(s["total_ms"]); scan.append(s["scan_ms"]); wall.append(e0.elapsed_time(e1))
    c = clk()
    s = st.stats()
    gbs = s["algo_bytes"] / (np.median(tot) * 1e-3) / 1e9
    print(f"{name:9s} launches={s['kernel_launches']:2d} search_ms med={np.median(tot):.4f} min={np.min(tot):.4f} scan_ms={np.median(scan):.4f} "
          f"stream_ms={np.median(wall):.4f} algo_GB={s['algo_bytes'] / 1e9:.3f} GB/s={gbs:.0f} frac_search={gbs / hbm:.3f} cands={s['candidates']} {c} first5={[round(x, 3) for x in tot[:5]]}")
# host API latency (isx_search): pinned result block written by the kernel
import time  # noqa: E402
st.set_stream(None)
for name, sel in (("1x64bit", [np.flatnonzero(qlens == 8)[0]]), ("1x256bit", [np.flatnonzero(qlens == 32)[0]])):
    q = np.ascontiguousarray(queries[sel]); ql = np.ascontiguousarray(qlens[sel])
    for _ in range(3):
        st.search(q, ql, 100)
    t = []
    for _ in range(args.reps):
        t0 = time.perf_counter(); st.search(q, ql, 100); t.append((time.perf_counter() - t0) * 1e3)
    print(f"host API {name}: wall ms med={np.median(t):.4f} min={np.min(t):.4f}")
