#!/usr/bin/env python
"""
bench.py - exact NPHD top-k over 100M mixed-length ISCC-UNITs (BASELINE.json config 3) on 1..8 B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU exact path (restated, oracle/)

Workload (config.workload = "cfg3"): 100M synthetic codes, 25 % each of 64/128/192/256 bit, uint64 keys;
one step = one batch of Q mixed-length queries (default 10 000, the batch BASELINE.json's configs 2 and 3 name) answered exactly, k = 100. With N GPUs the
100M rows are row-sharded (strong scaling), every rank scans its shard, the per-rank top-k records are
all-gathered over NCCL and merged on the device.

The JSON line reports
  value / ms_per_step   queries/s with the query batch already resident in HBM (device results, no D2H)
  e2e                   the same through the public host API: pinned host queries -> H2D -> search
                        (-> all-gather + merge) -> D2H of keys/hamming/nbits/counts, every step
  roofline              k_scan in its HBM-bound regime (one 256-bit query over the whole store), algorithmic
                        bytes = sum_b N_b*min(Lq,L_b) / CUDA-event time of the scan launches, vs MEASURED_PEAKS hbm_gbs
  roofline_popc         k_scan in the batch regime of the timed steps: algorithmic 32-bit popcounts / CUDA-event
                        scan time vs the POPC-pipe peak measured by profiles/microbench (15.91 /clk/SM)
  cpu_baseline          oracle/ C restatement of the reference's exact CPU search on the box's host cores
                        (N=1, rank 0), full 100M rows x a bounded number of queries; also the parity check.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOAD = "cfg3: 100M mixed 64/128/192/256-bit codes (25% each), exact NPHD top-k, row-sharded"
POPC_PER_CLK_PER_SM = 15.91  # measured, profiles/microbench/r01_pipes_b200.txt
HBM_FALLBACK_GBS = 6650.0    # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=100_000_000)
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured"
    return HBM_FALLBACK_GBS, 1965.0, "fallback"


class ClockSampler:
    """
    SM clock / throttle-reason samples taken DURING the timed region (B200_PROFILING.md's clocks line): an NVML
    polling thread in this process (2 ms period; `nvidia-smi -lms` cannot sample a sub-second region), falling back
    to an `nvidia-smi -lms 100` child when NVML is unusable. Only samples between mark_begin() and mark_end() count.
    """

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, uuid=None):
        self.gpu_index, self.uuid = gpu_index, uuid
        self.proc, self.thread, self.stop_flag = None, None, threading.Event()
        self.samples = []   # (t, sm_mhz, reasons bitmask or set)
        self.smax = None
        self.t0 = self.t1 = None
        self.how = None

    def _nvml_handle(self):
        import pynvml

        pynvml.nvmlInit()
        if self.uuid:
            try:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(self.uuid)
            except Exception:
                pass
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = self.gpu_index
        if vis and all(x.strip().isdigit() for x in vis.split(",")):
            idx = int(vis.split(",")[self.gpu_index])
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)

    def start(self):
        try:
            nv, h = self._nvml_handle()
            self.smax = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                     "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}

            def poll():
                while not self.stop_flag.is_set():
                    try:
                        mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                        mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                        self.samples.append((time.perf_counter(), mhz, {n for n, bit in names.items() if mask & bit}))
                    except Exception:
                        pass
                    time.sleep(0.002)

            self.how = "nvml"
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.how = "nvidia-smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            if len(f) < 9:
                continue
            try:
                mhz, smax = float(f[1]), float(f[2])
            except ValueError:
                continue
            self.smax = max(self.smax or 0.0, smax)
            self.samples.append((time.perf_counter(), mhz, {n for n, v in zip(names, f[5:9]) if v.lower().startswith("active")}))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.thread:
            self.thread.join(timeout=2)
        if self.how is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock source (NVML and nvidia-smi unavailable)"], "samples": 0}
        inside = [x for x in self.samples if self.t0 is not None and self.t1 is not None and self.t0 <= x[0] <= self.t1]
        window = "timed region"
        if not inside:   # region shorter than the sampling period of the fallback source: use everything sampled under load
            inside, window = self.samples, "warm-up + timed region"
        reasons = set()
        for _t, _mhz, r in inside:
            reasons |= r
        return {"sm_mhz": float(np.median([x[1] for x in inside])) if inside else None, "sm_max_mhz": self.smax,
                "reasons": sorted(reasons), "samples": len(inside), "source": self.how, "window": window}


def gen_shard(synth, start, n, seed, sink, chunk=2_000_000):
    """Generate rows [start, start+n) chunk by chunk and hand (keys, codes, lens) to `sink`."""
    for c0 in range(start, start + n, chunk):
        cn = min(chunk, start + n - c0)
        lens = synth.make_lengths(c0, cn, seed)
        codes = synth.make_codes(c0, cn, seed, lens)
        keys = synth.make_keys(c0, cn, seed)
        sink(keys, codes, lens)


def cpu_exact_search(args, n_threads=0):
    """Host arrays of the full store + the query batch (reference arm and cpu_baseline leg)."""
    from iscc_search_b200 import synth

    keys = np.empty(args.rows, dtype=np.uint64)
    codes = np.empty((args.rows, 32), dtype=np.uint8)
    lens = np.empty(args.rows, dtype=np.uint8)
    pos = [0]

    def sink(k, c, l):
        i = pos[0]
        keys[i:i + len(k)], codes[i:i + len(k)], lens[i:i + len(k)] = k, c, l
        pos[0] += len(k)

    gen_shard(synth, 0, args.rows, args.seed, sink)
    return keys, codes, lens


def run_reference(args):
    """`--impl reference`: the reference's CPU exact path (C restatement in oracle/, all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from iscc_search_b200 import synth
    from oracle import c_oracle

    c_oracle.build()
    cores = c_oracle.num_threads()
    keys, codes, lens = cpu_exact_search(args)
    queries, qlens = synth.make_queries(args.queries, args.rows, args.seed + 1, args.seed)
    # bounded sample per step: the first `qs` queries of the batch over ALL rows
    t0 = time.perf_counter()
    c_oracle.topk(keys, None, codes, lens, queries[:cores], qlens[:cores], args.k)
    t_probe = time.perf_counter() - t0
    budget = 120.0 / max(args.steps + args.warmup, 1)
    qs = int(max(cores, min(args.queries, cores * max(1, int(budget / max(t_probe, 1e-3))))))
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        c_oracle.topk(keys, None, codes, lens, queries[:qs], qlens[:qs], args.k)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = qs / (ms * 1e-3)
    sample = f"{qs} of the {args.queries} queries x all {args.rows} rows per step"
    emit(({
        "impl": "reference", "metric": "exact NPHD top-k queries/s at 100M codes", "value": value, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32 popcount", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows": args.rows, "queries_per_step": args.queries, "k": args.k,
                   "sample_queries_per_step": qs},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


_REAL_STDOUT = None


def emit(obj):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, warnings) was sent to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # libraries that print to fd 1 (e.g. "NCCL version ...") must not pollute the JSON line
    if args.impl == "reference":
        return run_reference(args)

    import torch

    from iscc_search_b200 import _lib, synth
    from iscc_search_b200.sharded import ShardedSearcher, record_layout

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: iscc_search_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    hbm_peak, sm_max_mhz, peak_kind = measured_peaks()

    # ---- build this rank's shard: rows [rank*M, (rank+1)*M) of the 100M-row data set -------------------
    per = args.rows // world
    start = rank * per
    n_local = per if rank < world - 1 else args.rows - start
    store = _lib.Store(device=local_rank, key_bytes=8, max_bytes=32)
    store.set_profiling(True)
    keep = (world == 1 and not args.no_cpu_baseline)
    host = {"keys": [], "codes": [], "lens": []}
    t_build = time.perf_counter()

    def sink(k, c, l):
        store.add(k, c, l)
        if keep:
            host["keys"].append(k), host["codes"].append(c), host["lens"].append(l)

    gen_shard(synth, start, n_local, args.seed, sink)
    t_build = time.perf_counter() - t_build
    assert store.size() == n_local

    Q, k = args.queries, args.k
    queries, qlens = synth.make_queries(Q, args.rows, args.seed + 1, args.seed)
    searcher = ShardedSearcher(store, rank, world, None, dev)
    d_queries = torch.from_numpy(queries).to(dev)
    pin_in = torch.from_numpy(queries.copy()).pin_memory()
    off, size = record_layout(Q, k)
    pin_out = torch.empty(size, dtype=torch.uint8, pin_memory=True)

    def step_device():
        searcher.search_device(d_queries, qlens, k)
        return store.stats()

    def step_e2e():
        res = searcher.search(queries, qlens, k, None, pin_in, pin_out)
        return res, store.stats()

    # ---- warm-up, then K timed steps (device-resident inputs). The store (2 GB) is far larger than L2. ----
    try:
        gpu_uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        gpu_uuid = None
    sampler = ClockSampler(local_rank, gpu_uuid)  # started before the warm-up so that the first sample is there in time
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches, scan_ms, algo_popc, issued_popc, pairs, cands, fallbacks = 0, 0.0, 0, 0, 0, 0, 0
    sampler.mark_begin()
    e0.record()
    for _ in range(args.steps):
        st = step_device()
        launches += st["kernel_launches"] + (1 if world > 1 else 0)
        scan_ms += st["scan_ms"]
        algo_popc += st["algo_popc"]
        issued_popc += st["issued_popc"]
        pairs += st["pairs"]
        cands += st["candidates"]
        fallbacks += st["fallback_queries"]
    e1.record()
    barrier()
    sampler.mark_end()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = Q / (ms_step * 1e-3)

    # ---- end to end through the host API ----
    for _ in range(2):
        step_e2e()
    barrier()
    e0.record()
    for _ in range(args.steps):
        res, _st = step_e2e()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_e2e = float(t.item()) / args.steps
    clocks = sampler.stop()

    # ---- HBM-bound regime of the same kernel: one 256-bit query (all planes) and one 64-bit query ----
    def scan_one(length_bytes, reps=10):
        q1 = np.zeros((1, 32), dtype=np.uint8)
        q1[0, :length_bytes] = queries[0, :length_bytes] if qlens[0] >= length_bytes else np.arange(length_bytes, dtype=np.uint8) * 37 + 11
        l1 = np.array([length_bytes], dtype=np.uint8)
        dq = torch.from_numpy(q1).to(dev)
        for _ in range(3):
            searcher.search_device(dq, l1, k)
        torch.cuda.synchronize()
        ms, total = [], []
        for _ in range(reps):
            searcher.search_device(dq, l1, k)
            s = store.stats()
            ms.append(s["scan_ms"])
            total.append(s["total_ms"])
        s = store.stats()
        return {"scan_ms": float(np.mean(ms)), "search_ms": float(np.mean(total)), "algo_bytes": int(s["algo_bytes"]),
                "scan_launches": int(s["scan_launches"]), "gbs": s["algo_bytes"] / (np.mean(ms) * 1e-3) / 1e9}

    scan256 = scan_one(32)
    scan64 = scan_one(8)

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    popc_peak = POPC_PER_CLK_PER_SM * 148 * sm_max_mhz * 1e6  # lane-popc/s at max clock
    popc_achieved = algo_popc / (scan_ms * 1e-3) if scan_ms > 0 else 0.0
    out = {
        "metric": "exact NPHD top-k queries/s at 100M codes", "value": value, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32 popcount", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows": args.rows, "rows_per_gpu": n_local, "queries_per_step": Q, "k": k,
                   "l2_policy": "store per GPU (>= 250 MB) exceeds the 126 MB L2; no flush needed",
                   "build_s": round(t_build, 1)},
        "codes_scanned_per_s": value * args.rows,
        "e2e": {"value": Q / (ms_e2e * 1e-3), "unit": "queries/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(Q * 32),
                "d2h_bytes_per_step": int(size)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": scan256["gbs"], "peak": hbm_peak, "unit": "GB/s", "frac": scan256["gbs"] / hbm_peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum over the 4 scan launches of this pass, ncu --set full
                     # at this exact configuration (profiles/r01d_ncu_full_q1_256bit_100M.csv): 2.001 GB + 18.8 MB
                     "traffic": 2019793664 if (world == 1 and args.rows == 100_000_000 and args.seed == 1) else None,
                     "peak_kind": peak_kind, "kernel": "k_scan<WE,G> (all launches of one pass)",
                     "regime": "1 query of 256 bit over this rank's rows: bytes = sum_b N_b*min(32, L_b)",
                     "algo_bytes_per_pass": scan256["algo_bytes"], "scan_ms": scan256["scan_ms"], "search_ms": scan256["search_ms"],
                     "scan_launches": scan256["scan_launches"],
                     "q64": {"achieved": scan64["gbs"], "frac": scan64["gbs"] / hbm_peak, "algo_bytes_per_pass": scan64["algo_bytes"],
                             "scan_ms": scan64["scan_ms"], "search_ms": scan64["search_ms"]}},
        "roofline_popc": {"bound": "popc", "achieved": popc_achieved / 1e12, "peak": popc_peak / 1e12, "unit": "Tpopc32/s",
                          "frac": popc_achieved / popc_peak if popc_peak else None,
                          "peak_kind": f"measured {POPC_PER_CLK_PER_SM}/clk/SM x 148 SM x {sm_max_mhz:.0f} MHz",
                          "regime": f"{Q} queries per step (timed region)", "scan_ms_per_step": scan_ms / args.steps,
                          "note": "achieved counts ALGORITHMIC popcounts (ceil(min(Lq,Lb)/4) per pair); the kernel issues fewer "
                                  "POPC (carry-save adders: 5 per 8 words; OR-fold lower-bound filter: 1 per group of 2-3 words once the "
                                  "threshold is tight), so frac exceeds 1; xu_pipe_frac is the share of the POPC pipe actually issued",
                          "xu_pipe_frac": (issued_popc / (scan_ms * 1e-3)) / popc_peak if scan_ms > 0 else None,
                          "pairs_per_s": pairs / (scan_ms * 1e-3) if scan_ms > 0 else None,
                          "candidates_per_query": cands / max(args.steps * Q, 1), "fallback_queries": int(fallbacks)},
        "clocks": clocks,
    }

    # ---- CPU baseline + parity (N=1 only): the oracle over ALL rows for a bounded number of queries ----
    if keep:
        from oracle import c_oracle

        c_oracle.build()
        keys_h = np.concatenate(host["keys"])
        codes_h = np.concatenate(host["codes"])
        lens_h = np.concatenate(host["lens"])
        cores = c_oracle.num_threads()
        t0 = time.perf_counter()
        rows, h, nb, cnt = c_oracle.topk(keys_h, None, codes_h, lens_h, queries[:cores], qlens[:cores], k)
        t_probe = time.perf_counter() - t0
        (gk, gh, gn, gc), _ = step_e2e()
        ok = bool(np.array_equal(gc[:cores], cnt) and all(
            np.array_equal(gk[i, :cnt[i]], keys_h[rows[i, :cnt[i]]]) and np.array_equal(gh[i, :cnt[i]], h[i, :cnt[i]])
            and np.array_equal(gn[i, :cnt[i]], nb[i, :cnt[i]]) for i in range(cores)))
        qs = int(min(Q, max(cores, cores * int(args.cpu_seconds / max(t_probe, 1e-3)))))
        t0 = time.perf_counter()
        c_oracle.topk(keys_h, None, codes_h, lens_h, queries[:qs], qlens[:qs], k)
        t_cpu = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": qs / t_cpu, "unit": "queries/s", "cores": cores, "kind": "port",
                               "sample": f"{qs} of the {Q} queries x all {args.rows} rows, {t_cpu:.1f} s",
                               "what": "oracle/nphd_oracle.c: OpenMP restatement of the reference's exact CPU scan (not the usearch binary)"}
        out["parity"] = {"checked_queries": cores, "rows": args.rows, "bit_exact": ok}
    else:
        out["cpu_baseline"] = None
    emit(out)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
