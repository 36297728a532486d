import sys, numpy as np
sys.path.insert(0, "/root/repo")
import torch
from bench import CONFIGS, build_store
from iscc_search_b200 import _lib, synth
from iscc_search_b200.sharded import ShardedSearcher
cfg = dict(CONFIGS["cfg3"]); dev = torch.device("cuda", 0)
st = _lib.Store(key_bytes=8, max_bytes=32); st.set_profiling(True)
st.set_stream(torch.cuda.current_stream(dev).cuda_stream)
build_store(st, cfg, 1, 0, cfg["rows"], dev, torch)
queries, qlens = synth.make_queries(64, cfg["rows"], 2, 1)
searcher = ShardedSearcher(st, 0, 1, None, dev)
for L in (8, 32):
    i = np.flatnonzero(qlens == L)[0]
    for rep in range(3):
        st.search(queries[i:i+1], qlens[i:i+1], 100)
        print("host api", L, st.stats()["total_ms"], file=sys.stderr)
    dq = torch.from_numpy(np.ascontiguousarray(queries[i:i+1])).to(dev)
    for rep in range(3):
        searcher.search_device(dq, np.ascontiguousarray(qlens[i:i+1]), 100)
        print("device api", L, st.stats()["total_ms"], file=sys.stderr)
