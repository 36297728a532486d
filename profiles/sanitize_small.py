"""compute-sanitizer target: the new round-2 kernels on small stores (k_scan_small, bulk append, synth, score, moves)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from iscc_search_b200 import _lib, synth  # noqa: E402
from tests.helpers import assert_same_topk, oracle_topk  # noqa: E402

dev = torch.device("cuda", 0)
n = 60_000
st = _lib.Store(key_bytes=8, max_bytes=32)
st.set_stream(torch.cuda.current_stream(dev).cuda_stream)
d_keys = torch.empty(n * 8, dtype=torch.uint8, device=dev)
d_codes = torch.empty(n * 32, dtype=torch.uint8, device=dev)
d_lens = torch.empty(n, dtype=torch.uint8, device=dev)
st.synth_rows_device(5, 0, n, d_keys.data_ptr(), d_codes.data_ptr(), d_lens.data_ptr())
st.add_device(d_keys.data_ptr(), d_codes.data_ptr(), d_lens.data_ptr(), n)
st.set_stream(None)
lens = synth.make_lengths(0, n, 5); codes = synth.make_codes(0, n, 5, lens); keys = synth.make_keys(0, n, 5)
for q, k in ((1, 10), (4, 100), (3, 2048)):
    queries, qlens = synth.make_queries(q, n, 6 + q, 5)
    gk, gh, gn, gc, _ = st.search(queries, qlens, k)
    assert st.stats()["kernel_launches"] == 1
    rows, h, nb, cnt = oracle_topk(keys, codes, lens, queries, qlens, k)
    assert_same_topk(gk, gh, gn, gc, keys, rows, h, nb, cnt)
queries, qlens = synth.make_queries(300, n, 9, 5)          # batch path: lanes, emit-path tighten
gk, gh, gn, gc, _ = st.search(queries, qlens, 20)
rows, h, nb, cnt = oracle_topk(keys, codes, lens, queries, qlens, 20)
assert_same_topk(gk, gh, gn, gc, keys, rows, h, nb, cnt)
gone = np.random.default_rng(1).permutation(n)[: n // 3]
st.remove(np.ascontiguousarray(keys[gone]), len(gone))
keep = np.ones(n, bool); keep[gone] = False
gk, gh, gn, gc, _ = st.search(queries[:4], qlens[:4], 20)
rows, h, nb, cnt = oracle_topk(keys[keep], codes[keep], lens[keep], queries[:4], qlens[:4], 20)
assert_same_topk(gk, gh, gn, gc, keys[keep], rows, h, nb, cnt)
seg = np.array([0, 2, 3], dtype=np.uint32)
print(st.score_segments(seg, np.array([0, 2, 1], np.uint32), np.array([1.0, 0.5, 0.75]), np.array([1.5, 2.0, 1.0]), np.array([1.0, 1.0, 1.0])))
st.close()
print("sanitize target ok")
