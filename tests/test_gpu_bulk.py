"""
Device-side bulk load (isx_add_device) and the on-device generator of the synthetic data set (isx_synth_rows_device):
the generator must reproduce iscc_search_b200/synth.py bit for bit (the CPU oracle regenerates the same rows), and a
store filled from device memory must answer exactly like one filled through isx_add - including the keyed operations
that complete the host key map lazily.
"""

import numpy as np
import pytest

from iscc_search_b200 import _lib, synth
from tests.helpers import assert_same_topk, oracle_topk

pytestmark = pytest.mark.gpu


def _gen(store, torch, seed, start, n, key_bytes, **kw):
    dev = torch.device("cuda", 0)
    d_keys = torch.empty(n * key_bytes, dtype=torch.uint8, device=dev)
    d_codes = torch.empty(n * 32, dtype=torch.uint8, device=dev)
    d_lens = torch.empty(n, dtype=torch.uint8, device=dev)
    store.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    store.synth_rows_device(seed, start, n, d_keys.data_ptr(), d_codes.data_ptr(), d_lens.data_ptr(), **kw)
    torch.cuda.synchronize()
    return d_keys, d_codes, d_lens


def test_device_generator_equals_synth_py(cuda):
    import torch

    st = _lib.Store(key_bytes=8, max_bytes=32)
    for start, n, seed in ((0, 10_000, 1), (987_654_321, 5_000, 9)):
        d_keys, d_codes, d_lens = _gen(st, torch, seed, start, n, 8)
        lens = synth.make_lengths(start, n, seed)
        assert np.array_equal(d_lens.cpu().numpy(), lens)
        assert np.array_equal(d_codes.cpu().numpy().reshape(n, 32), synth.make_codes(start, n, seed, lens))
        assert np.array_equal(d_keys.cpu().numpy().view(np.uint64), synth.make_keys(start, n, seed))
    st.close()
    st = _lib.Store(key_bytes=16, max_bytes=8, fixed_len=8)
    n = 6000
    d_keys, d_codes, d_lens = _gen(st, torch, 4, 50, n, 16, lengths=(8,), key_mode=1, cpa=64, dup_every=16, dup_back=65)
    l8 = np.full(n, 8, dtype=np.uint8)
    assert np.array_equal(d_lens.cpu().numpy(), l8)
    assert np.array_equal(d_codes.cpu().numpy().reshape(n, 32), synth.make_codes(50, n, 4, l8, dup_every=16, dup_back=65))
    hi, lo = synth.make_keys128(50, n, 4, 64)
    assert np.array_equal(d_keys.cpu().numpy().reshape(n, 16), synth.keys128_bytes(hi, lo))
    st.close()


def test_bulk_append_mixed_lengths_search_and_lazy_key_map(cuda, tmp_path):
    import torch

    n, q, k, seed = 300_000, 64, 100, 13
    st = _lib.Store(key_bytes=8, max_bytes=32)
    host_n = 5_000   # some rows go in through the host path first: the bulk append continues their segments
    lens = synth.make_lengths(0, n, seed)
    codes = synth.make_codes(0, n, seed, lens)
    keys = synth.make_keys(0, n, seed)
    st.add(keys[:host_n], codes[:host_n], lens[:host_n])
    d_keys, d_codes, d_lens = _gen(st, torch, seed, host_n, n - host_n, 8)
    st.add_device(d_keys.data_ptr(), d_codes.data_ptr(), d_lens.data_ptr(), n - host_n)
    assert st.size() == n
    assert st.length_mask() == (1 << 7) | (1 << 15) | (1 << 23) | (1 << 31)
    queries, qlens = synth.make_queries(q, n, seed + 1, seed)
    gk, gh, gn, gc, _ = st.search(queries, qlens, k)
    rows, h, nb, cnt = oracle_topk(keys, codes, lens, queries, qlens, k)
    assert_same_topk(gk, gh, gn, gc, keys, rows, h, nb, cnt)
    # keyed operations complete the key map from the device keys
    probe = np.concatenate([keys[[0, host_n, n - 1, 123_456]], np.array([12345], dtype=np.uint64)])
    assert st.contains(probe, 5).tolist() == [True, True, True, True, False]
    got_codes, got_lens = st.get(probe, 5)
    assert got_lens.tolist() == [int(lens[0]), int(lens[host_n]), int(lens[n - 1]), int(lens[123_456]), 0]
    assert np.array_equal(got_codes[1], codes[host_n]) and np.array_equal(got_codes[3], codes[123_456])
    # first-wins add on top of bulk rows, remove of bulk rows, search again
    added = st.add(keys[n - 3:], codes[n - 3:], lens[n - 3:])
    assert added.tolist() == [0, 0, 0] and st.size() == n
    rem_idx = np.arange(host_n, n, 7)
    removed, cnt_rm = st.remove(np.ascontiguousarray(keys[rem_idx]), len(rem_idx))
    assert cnt_rm == len(rem_idx) and removed.all()
    keep = np.ones(n, dtype=bool)
    keep[rem_idx] = False
    gk, gh, gn, gc, _ = st.search(queries, qlens, k)
    rows, h, nb, cnt = oracle_topk(keys[keep], codes[keep], lens[keep], queries, qlens, k)
    assert_same_topk(gk, gh, gn, gc, keys[keep], rows, h, nb, cnt)
    # snapshot round trip of a bulk-filled store
    st.save(tmp_path / "bulk.isx")
    st2 = _lib.Store(key_bytes=8, max_bytes=32)
    st2.load(tmp_path / "bulk.isx")
    assert st2.size() == int(keep.sum())
    gk2, gh2, gn2, gc2, _ = st2.search(queries, qlens, k)
    assert np.array_equal(gk2, gk) and np.array_equal(gh2, gh) and np.array_equal(gc2, gc)
    st.close()
    st2.close()


def test_bulk_append_uniform_length_128bit_keys_threshold_and_join(cuda):
    import torch

    n, seed = 200_000, 4
    st = _lib.Store(key_bytes=16, max_bytes=8, fixed_len=8)
    kw = dict(lengths=(8,), key_mode=1, cpa=64, dup_every=16, dup_back=65)
    for c0 in range(0, n, 70_000):   # several appends: spans continue in the same segments
        cn = min(70_000, n - c0)
        d_keys, d_codes, _ = _gen(st, torch, seed, c0, cn, 16, **kw)
        st.add_device(d_keys.data_ptr(), d_codes.data_ptr(), None, cn, uniform_len=8)
    assert st.size() == n
    l8 = np.full(n, 8, dtype=np.uint8)
    codes = synth.make_codes(0, n, seed, l8, dup_every=16, dup_back=65)
    hi, lo = synth.make_keys128(0, n, seed, 64)
    qs = np.ascontiguousarray(codes[[14, 206, 4110, 77_774, 150_014]])
    ql = np.full(5, 8, dtype=np.uint8)
    for kk, thr in ((300, (16, 64)), (1000, (0, 64))):
        gk, gh, gn, gc, _ = st.search(qs, ql, kk, thr)
        rows, h, nb, cnt = oracle_topk(hi, codes, l8, qs, ql, kk, thr, keys_lo=lo)
        assert_same_topk(gk, gh, gn, gc, hi, rows, h, nb, cnt, keys_lo=lo)
        assert (gc >= 2).all() if thr == (0, 64) else True
    st.close()


def test_bulk_append_with_duplicate_keys_is_reported_by_the_first_keyed_call(cuda):
    import torch

    st = _lib.Store(key_bytes=8, max_bytes=32)
    d_keys, d_codes, d_lens = _gen(st, torch, 3, 0, 1000, 8)
    st.add_device(d_keys.data_ptr(), d_codes.data_ptr(), d_lens.data_ptr(), 1000)
    st.add_device(d_keys.data_ptr(), d_codes.data_ptr(), d_lens.data_ptr(), 1000)   # promise broken: same keys again
    assert st.size() == 2000
    with pytest.raises(ValueError, match="not unique"):
        st.contains(synth.make_keys(0, 1, 3), 1)
    st.close()


def test_bulk_append_rejects_lengths_the_index_does_not_accept(cuda):
    import torch

    st = _lib.Store(key_bytes=8, max_bytes=16)
    d_keys, d_codes, d_lens = _gen(st, torch, 3, 0, 1000, 8)   # lengths up to 32 bytes
    with pytest.raises(ValueError, match="outside"):
        st.add_device(d_keys.data_ptr(), d_codes.data_ptr(), d_lens.data_ptr(), 1000)
    assert st.size() == 0
    with pytest.raises(ValueError, match="uniform_len"):
        st.add_device(d_keys.data_ptr(), d_codes.data_ptr(), None, 1000, uniform_len=24)
    st.close()
