#!/usr/bin/env python
"""bench.py -- headline benchmark of the exact-search hot path (BASELINE.json).

A "step" is one pass of the hot path over one batch of synthetic queries: the database is scored against every
query of the batch and the top-k kept.  N=1 headline = configs[1] (FlatIndex 1M x 768 cosine, batch 1024, k=10;
tcgen05 path).  The same JSON line carries `secondary`: the other BASELINE.json configs (C1, C3a on the fp32 scan
and on the routed path, C3b, C4 with the reference's post-filter and with the filter pushed down at 1 % / 50 %, a C5
shard), each with its own roofline, e2e and sampled CPU baseline, and `sustained`: 1000 steps of the headline.

N>1 (`--gpus N`, launched under torchrun): ONE index sharded row-wise over the N GPUs INSIDE libgfi
(gfi_create_sharded), owned by rank 0 -- the reference server is one process with one index behind one lock
(src/server/mod.rs:13-16), so that is the deployment to measure.  The other ranks join the process group, the
barriers and the max-over-ranks reduction and hold no data.  Headline = C2 weak-scaled (every GPU holds 1M rows);
secondary = C3a / C3b strong-scaled (10M rows over N GPUs) and C5 (12.5M rows per GPU: 100M over 8), each naming
its exchange + merge time.  `--sharding ranks` keeps round 1's one-process-per-GPU NCCL all-gather variant.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the reference (oracle/, all host
threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# name: metric, rows per GPU (weak) or in total (strong), d, generator kind, seed, batch q, k
WORKLOADS = {
    "c2": ("cosine", 1_000_000, 768, 1, 3, 1024, 10),
    "c1": ("euclidean", 10_000, 128, 0, 1, 1, 10),
    "c3a": ("dot", 10_000_000, 768, 1, 5, 1, 100),
    "c3a_scan": ("dot", 10_000_000, 768, 1, 5, 1, 100),
    "c3b": ("dot", 10_000_000, 768, 1, 5, 64, 100),
    "c4": ("euclidean", 10_000_000, 384, 0, 6, 1, 10),
    "c5": ("euclidean", 12_500_000, 128, 0, 7, 4096, 10),
    "c4f1": ("euclidean", 10_000_000, 384, 0, 6, 1, 10),
    "c4f50": ("euclidean", 10_000_000, 384, 0, 6, 1, 10),
    "c4p1": ("euclidean", 10_000_000, 384, 0, 6, 1, 10),
    "c4p50": ("euclidean", 10_000_000, 384, 0, 6, 1, 10),
}
FILTER_PCT = {"c4f1": 1, "c4f50": 50, "c4p1": 1, "c4p50": 50}  # eq-filter selectivity
POST_FILTER = {"c4p1", "c4p50"}  # the reference's semantics: search 3k unfiltered, filter the hits on the host
STRONG = {"c3a", "c3a_scan", "c3b", "c4", "c4f1", "c4f50", "c4p1", "c4p50"}  # N>1: total rows fixed, sharded N ways
WL_OPTS = {"c3a_scan": {"tensor_auto": 0}}
NAMES = {
    "c2": "FlatIndex 1M x 768-d cosine, batch 1024, k=10 (BASELINE.json configs[1])",
    "c1": "FlatIndex 10k x 128-d Euclidean, k=10, single queries (configs[0])",
    "c3a": "FlatIndex 10M x 768-d dot, single query, k=100 (configs[2]), cost-model route",
    "c3a_scan": "FlatIndex 10M x 768-d dot, single query, k=100 (configs[2]), fp32 streaming scan",
    "c3b": "FlatIndex 10M x 768-d dot, batch 64, k=100 (configs[2])",
    "c4": "FlatIndex 10M x 384-d Euclidean, single query, k=10, unfiltered scan (configs[3])",
    "c5": "FlatIndex 12.5M x 128-d Euclidean per GPU, batch 4096, k=10 (configs[4] shard)",
    "c4f1": "Filtered search 10M x 384-d Euclidean, eq filter at 1% selectivity, pushed down (device-side filter), k=10 (configs[3])",
    "c4f50": "Filtered search 10M x 384-d Euclidean, eq filter at 50% selectivity, pushed down (device-side filter), k=10 (configs[3])",
    "c4p1": "Filtered search 10M x 384-d Euclidean, eq filter at 1% selectivity, reference post-filter (fetch 3k, filter hits), k=10 (configs[3])",
    "c4p50": "Filtered search 10M x 384-d Euclidean, eq filter at 50% selectivity, reference post-filter (fetch 3k, filter hits), k=10 (configs[3])",
}
METRIC_ID = {"euclidean": 0, "cosine": 1, "dot": 2}
L2_BYTES = 126e6


def total_rows(wl, world):
    n = WORKLOADS[wl][1]
    return n if (wl in STRONG or wl == "c1") else n * world


def metric_name(wl):
    _, n, _, _, _, _, k = WORKLOADS[wl]
    return "queries/sec (exact flat search, k=%d; per %d-row shard searched)" % (k, n)


def config_of(wl, world):
    metric, n, d, _, _, q, k = WORKLOADS[wl]
    tot = total_rows(wl, world)
    db = tot * d * 4 / max(world, 1)
    pol = ("inputs larger than L2 (each GPU's rows >> 126 MB)" if db > 2 * L2_BYTES else
           "database smaller than L2 (%.1f MB): L2-resident on purpose, the reference's own bench size" % (db / 1e6))
    return {"workload": NAMES[wl], "rows_per_gpu": tot // max(world, 1), "dim": d, "metric": metric, "batch": q, "k": k,
            "index_rows_total": tot, "l2_policy": pol,
            "parallelism": f"row-sharded x{world}" if world > 1 else "single GPU"}


def ncu_traffic(wl):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, or None."""
    import glob
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_traffic.json")), reverse=True):
        try:
            t = json.load(open(p)).get(wl)
            if t:
                return t["bytes_per_launch"]
        except Exception:
            pass
    return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons during the timed region: NVML in-process (about a millisecond per
    sample) when it loads, else the `nvidia-smi` query of the profiling recipe (tens of milliseconds per sample)."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.stop_flag, self.rows, self.source = gpu, threading.Event(), [], "nvidia-smi"
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML orders devices by PCI bus id; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = int(vis.split(",")[gpu]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else gpu
            self.nv = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys))
            self.source = "nvml"
        except Exception:
            self.nv = None

    def sample_nvml(self):
        nv, h = self.nv
        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        bits = [nv.nvmlClocksThrottleReasonHwSlowdown, nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                nv.nvmlClocksThrottleReasonSwThermalSlowdown, nv.nvmlClocksThrottleReasonSwPowerCap]
        return [str(sm), str(mx)] + ["Active" if r & b else "Not Active" for b in bits]

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                if self.nv:
                    self.rows.append(self.sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                    self.rows.append([x.strip() for x in out.stdout.strip().split(",")])
            except Exception:
                if self.nv:
                    self.nv, self.source = None, "nvidia-smi"
            self.stop_flag.wait(0.005 if self.nv else 0.05)

    def finish(self):
        self.stop_flag.set()
        self.join(timeout=3)
        return self.summary()

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if len(r) >= 6 and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) >= 6 and r[1].isdigit()]
        reasons = sorted({self.NAMES[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm), "source": self.source}


def host_cores():
    import oracle
    # all the host cores this process may run on (not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1,
    # which would silently time a single-threaded baseline at N > 1; the oracle passes num_threads() explicitly)
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    return max(cores, oracle.max_threads())


def eligible_rows(first, n, pct):
    """The synthetic eq filter: row id -> `pct` % of the rows carry the matching value (a fixed hash of the id)."""
    import numpy as np
    ids_all = np.arange(first, first + n, dtype=np.uint64)
    hsh = (ids_all * np.uint64(2654435761)) >> np.uint64(7)
    return (hsh % np.uint64(100)) < np.uint64(pct)


def cpu_reference_run(wl, steps, warmup, budget_s=15.0, max_row_bytes=8e9):
    """The reference's CPU path (oracle/ restatement: sequential f32 sums, full sort per query), all host threads
    over queries, each step a bounded sample of the workload's queries against (a bounded prefix of) the database.
    Filtered workloads run the reference's own semantics (storage.rs:249-290): unfiltered top-3k, then the filter."""
    import numpy as np
    import oracle
    metric, n, d, kind, seed, q, k = WORKLOADS[wl]
    cores = host_cores()
    kk = 3 * k if wl in FILTER_PCT else k  # search_with_filter over-fetches (storage.rs:257)
    n_cpu = n
    if n * d * 4 > max_row_bytes:  # keep host memory and generation time bounded: scale rows, report it
        n_cpu = int(max_row_bytes / (d * 4))
    # one query costs about n*d*4 cycles on one core; bound one step to budget_s / (steps + warmup)
    per_query_s = max(n_cpu * d * 4 / 2.5e9, 1e-5)
    per_step = budget_s / max(steps + warmup, 1)
    sample_q = int(max(1, min(q, round(per_step * cores / per_query_s))))
    cores = max(1, min(cores, sample_q))
    rows = oracle.gen_rows(seed, 0, n_cpu, d, kind)
    queries = oracle.gen_rows(seed + 1, 0, sample_q, d, kind)
    elig = eligible_rows(0, n_cpu, FILTER_PCT[wl]) if wl in FILTER_PCT else None
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        res = oracle.search_batch(metric, rows, queries, kk, threads=cores)
        if elig is not None:
            res = [[(i, dd) for i, dd in zip(ids_, ds_) if elig[int(i)]][:k] for ids_, ds_ in res]
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    qps = sample_q / t * (n_cpu / n)  # rows scale linearly if the database was truncated
    sample = f"{sample_q} of {q} queries against {n_cpu} of {n} rows x {d}, {cores} threads over queries"
    # What the reference does today: FlatIndex::search is single-threaded and search_batch maps over the queries
    # sequentially (src/storage.rs:306-309), so one core serves the whole batch.  Timed on one query.
    t0 = time.perf_counter()
    oracle.search_batch(metric, rows, queries[:1], kk, threads=1)
    single = 1.0 / (time.perf_counter() - t0) * (n_cpu / n)
    return {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample,
            "single_thread_value": single}, t


def reference_arm(args, wl, rank, world):
    if rank != 0:
        return
    _, n, _, _, _, q, k = WORKLOADS[wl]
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    cpu, t = cpu_reference_run(wl, steps, warmup, budget_s=100.0)
    qps = cpu["value"]
    # N>1: the reference is one process; the box's host cores serve the same per-shard database the GPU arm's
    # `value` is quoted on (queries/s per n-row shard searched), so the two lines compare like with like
    line = {"impl": "reference", "metric": metric_name(wl), "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_of(wl, world),
            "cpu_baseline": cpu,
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU restatement of FlatIndex::search (oracle/), OpenMP over queries on all host cores; the "
                    "reference itself is single-threaded (single_thread_value)"}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
class Bench:
    """One workload on one index (single-GPU handle, or one handle sharded over `devices` inside libgfi)."""

    def __init__(self, wl, devices, opts=(), seed_first=0):
        import numpy as np
        import torch
        import vectordb_from_scratch_b200 as gfi
        from vectordb_from_scratch_b200 import synth
        self.np, self.torch, self.wl, self.devices = np, torch, wl, list(devices)
        metric, n, d, kind, seed, q, k = WORKLOADS[wl]
        self.G = len(self.devices)
        self.n_total = total_rows(wl, self.G)
        self.q, self.k, self.d = q, k, d
        self.dev = torch.device("cuda", self.devices[0])
        torch.cuda.set_device(self.dev)
        t0 = time.perf_counter()
        if self.G == 1:
            idx = gfi.GpuFlatIndex(METRIC_ID[metric], dim=d, device=self.devices[0])
        else:
            idx = gfi.GpuFlatIndex(METRIC_ID[metric], dim=d, devices=self.devices)
        idx.reserve(self.n_total)  # sharded: contiguous id ranges, one per GPU
        idx.add_generated(seed, seed_first, self.n_total, kind, seed_first)
        idx.flush()
        self.first = seed_first
        self.idx = idx
        idx.set_option("profile", 1)
        for name, val in {**WL_OPTS.get(wl, {}), **dict(opts)}.items():
            idx.set_option(name, int(val))
        self.kk = 3 * k if wl in POST_FILTER else k  # VectorStore::search_with_filter fetches 3k (storage.rs:257)
        self.queries_h = synth.gen_rows(seed + 1, 0, q, d, kind)
        self.ks_h = np.full(q, self.kk, dtype=np.uint32)
        self.n_elig = self.n_total
        self.mask_words = None
        self.filter_json = None
        if wl in FILTER_PCT:
            elig = eligible_rows(seed_first, self.n_total, FILTER_PCT[wl])
            self.elig = elig
            self.n_elig = int(elig.sum())
            if wl not in POST_FILTER:
                # the rows' metadata lives in HBM next to them (one dictionary-encoded column), loaded in bulk
                ids = np.arange(seed_first, seed_first + self.n_total, dtype=np.uint64)
                idx.set_metadata_column("tag", ids, ["hit", "miss"], np.where(elig, 0, 1).astype(np.uint32))
                self.filter_json = json.dumps({"op": "eq", "field": "tag", "value": "hit"})
                from vectordb_from_scratch_b200.index import pack_mask
                full = np.zeros(seed_first + self.n_total, dtype=bool)
                full[seed_first:] = elig
                self.mask_words, self.mask_bits = pack_mask(full)
        self.build_s = time.perf_counter() - t0

    def close(self):
        self.idx.close()
        self.torch.cuda.empty_cache()

    # ---- HBM-resident throughput: inputs on the (root) GPU, CUDA events on the launching stream ----
    def device_leg(self, steps, warmup, sampler_gpu=None):
        torch, np, idx, q, kk = self.torch, self.np, self.idx, self.q, self.kk
        dev = self.dev
        # A sharded index takes batches round-robin on four streams: a batch's inputs are ordered behind the previous
        # merge on ITS stream only, so the exchange + merge of one batch overlap the main passes of the next three
        # (on the root GPU a merge gets SMs only between that GPU's own persistent kernels; libgfi cycles through four
        # gather blocks to match).
        n_str = 4 if self.G > 1 else 1
        tss = [torch.cuda.Stream(device=dev) for _ in range(n_str)]
        ts = tss[0]
        outs = []
        with torch.cuda.stream(ts):
            dq = torch.from_numpy(self.queries_h).to(dev)
            dks = torch.from_numpy(self.ks_h.astype(np.int32)).to(dev)
            for _ in range(n_str):
                outs.append((torch.zeros((q, kk), dtype=torch.int64, device=dev),
                             torch.zeros((q, kk), dtype=torch.float32, device=dev),
                             torch.zeros((q,), dtype=torch.int32, device=dev)))
            d_mask_ptr, mask_bits = 0, 0
            if self.mask_words is not None:
                self.mask_t = torch.from_numpy(self.mask_words.view(np.int64)).to(dev)
                d_mask_ptr, mask_bits = self.mask_t.data_ptr(), self.mask_bits
        ts.synchronize()

        def step(i):
            o_ids, o_d, o_c = outs[i % n_str]
            idx.search_device(dq.data_ptr(), q, dks.data_ptr(), kk, o_ids.data_ptr(), o_d.data_ptr(), o_c.data_ptr(),
                              kk, stream=tss[i % n_str].cuda_stream, d_mask=d_mask_ptr, mask_bits=mask_bits)

        for i in range(warmup):
            step(i)
        idx.search_status()
        for t in tss:
            t.synchronize()
        st0 = idx.stats()
        sh0 = self.per_shard_stats()
        sampler = ClockSampler(sampler_gpu) if sampler_gpu is not None else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ts)
        for t in tss[1:]:
            t.wait_event(e0)
        h0 = time.perf_counter()
        for i in range(steps):
            step(i)
        self.host_enqueue_ms = (time.perf_counter() - h0) * 1e3 / steps  # host time to enqueue one step (no sync)
        for t in tss[1:]:
            ev = torch.cuda.Event()
            ev.record(t)
            ts.wait_event(ev)
        e1.record(ts)
        for t in tss:
            t.synchronize()
        idx.search_status()
        ms = e0.elapsed_time(e1)
        clocks = sampler.finish() if sampler else None
        st1 = idx.stats()
        sh1 = self.per_shard_stats()
        # dominant-kernel time per launch on every shard's GPU (the slowest shard sets the step)
        self.shard_kernel_ms = [max(b["tensor_kernel_ns"] - a["tensor_kernel_ns"], b["scan_kernel_ns"] - a["scan_kernel_ns"]) /
                                max(1, max(b["tensor_kernel_count"] - a["tensor_kernel_count"],
                                           b["scan_kernel_count"] - a["scan_kernel_count"])) / 1e6 for a, b in zip(sh0, sh1)]
        dq_keep = (dq, dks, outs)
        self.keep = dq_keep
        return ms, st0, st1, clocks

    def per_shard_stats(self):
        if self.G == 1:
            return []
        out = []
        for g in range(self.G):
            self.idx.set_option("stats_shard", g)
            out.append(self.idx.stats())
        self.idx.set_option("stats_shard", -1)
        return out

    # ---- end to end: host buffers through the public C-ABI call, copies inside the timed region ----
    def e2e_leg(self, steps):
        torch, np, idx, q = self.torch, self.np, self.idx, self.q
        queries_pin = torch.from_numpy(self.queries_h).pin_memory()  # the step's inputs live in pinned host memory
        queries_h = queries_pin.numpy()
        idx.set_option("profile", 0)  # per-kernel event timing is not part of the user's call
        elig_by_id = None
        if self.wl in POST_FILTER:
            elig_by_id = np.zeros(self.first + self.n_total, dtype=bool)
            elig_by_id[self.first:] = self.elig

        def call():
            if self.filter_json is not None:  # pushed down: the filter is evaluated on the GPU over resident columns
                return idx.search_filtered(queries_h, self.ks_h, self.filter_json)
            ids, dist, cnt = idx.search_arrays(queries_h, self.ks_h)
            if elig_by_id is not None:  # reference semantics: filter the 3k hits on the host, keep k
                out = []
                for i in range(q):
                    keep = elig_by_id[ids[i, :cnt[i]].astype(np.int64)]
                    out.append((ids[i, :cnt[i]][keep][:self.k], dist[i, :cnt[i]][keep][:self.k]))
                return out
            return ids, dist, cnt

        for _ in range(3):
            call()
        # single-query workloads: 256 calls, so that the median / p99 call latency means something
        n_calls = 256 if q == 1 else max(4, min(steps, 50))
        lat = []
        t0 = time.perf_counter()
        for _ in range(n_calls):
            t1 = time.perf_counter()
            call()
            lat.append(time.perf_counter() - t1)
        t_serial = (time.perf_counter() - t0) / n_calls
        t_e2e, callers = t_serial, 1
        if q >= 256:
            # Batch workloads: TWO caller threads, each making the same synchronous gfi_search calls (the reference
            # server answers requests from many workers over one index, src/server/mod.rs:13-16).  Every call still
            # copies its queries in and its results out; the copies of one call overlap the kernels of the other.
            import threading
            callers = 2
            per = max(3, n_calls // callers)
            gate = threading.Barrier(callers + 1)
            errs = []

            def worker():
                try:
                    for _ in range(3):  # untimed: the second caller's search context allocates its buffers here
                        call()
                    gate.wait()
                    for _ in range(per):
                        call()
                except Exception as e:  # surfaced below: a failed call must fail the bench, not shorten the run
                    errs.append(e)
                    gate.abort()

            ths = [threading.Thread(target=worker) for _ in range(callers)]
            for t in ths:
                t.start()
            try:
                gate.wait()
            except threading.BrokenBarrierError:
                pass
            t0 = time.perf_counter()
            for t in ths:
                t.join()
            t_e2e = (time.perf_counter() - t0) / (per * callers)
            if errs:
                raise errs[0]
        idx.search_status()
        idx.set_option("profile", 1)
        lat.sort()
        h2d = q * self.d * 4 + q * 4 + (len(self.filter_json) if self.filter_json else 0)
        d2h = q * self.kk * 12 + q * 4 + 64 * self.G
        return {"value": self.G_units() * q / t_e2e, "unit": "queries/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": t_e2e * 1e3,
                "callers": callers, "serial_ms_per_step": t_serial * 1e3,
                "serial_value": self.G_units() * q / t_serial,
                "call_latency_ms": {"median": lat[len(lat) // 2] * 1e3,
                                    "p99": lat[min(len(lat) - 1, int(len(lat) * 0.99))] * 1e3, "best": lat[0] * 1e3,
                                    "calls": len(lat)},
                "api": "gfi_search_filtered (filter JSON; metadata columns resident)" if self.filter_json else
                       "gfi_search + host post-filter of the 3k hits" if self.wl in POST_FILTER else "gfi_search"}

    def G_units(self):
        """Weak-scaled workloads: each of the G GPUs scores the batch against its own n-row shard, so a step is
        G * q (query, shard) searches; strong-scaled ones (total rows fixed): q queries."""
        return 1 if (self.wl in STRONG or self.wl == "c1") else self.G

    def roofline(self, ms_total, st0, st1):
        pk, pk_kind = peaks()
        wl, G = self.wl, self.G
        _, _, d, _, _, q, _ = WORKLOADS[wl]
        n_shard = self.n_total / G
        tk_n = st1["tensor_kernel_count"] - st0["tensor_kernel_count"]
        sk_n = st1["scan_kernel_count"] - st0["scan_kernel_count"]
        tk_ns = st1["tensor_kernel_ns"] - st0["tensor_kernel_ns"]
        sk_ns = st1["scan_kernel_ns"] - st0["scan_kernel_ns"]
        # (with a device-resident mask both kernels are enqueued and the device-side route lets one of them exit at
        # once: the dominant kernel is the one that took the time)
        if tk_n > 0 and tk_ns >= sk_ns:
            kern_ms = tk_ns / tk_n / 1e6
            flops = 2.0 * n_shard * d * q  # algorithmic, per launch (one shard): counted once
            achieved = flops / (kern_ms * 1e-3) / 1e12
            long_step = ms_total > 2000.0
            peak = pk["bf16_tflops_sustained"] if long_step else pk["bf16_tflops"]
            hbm_gbs = n_shard * d * 2 / (kern_ms * 1e-3) / 1e9  # the fp16 shadow rows are read once per launch
            roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": ncu_traffic(wl), "kernel": "gemm_topk_kernel (main pass)", "kernel_ms": kern_ms,
                    "peak_source": f"{pk_kind} MEASURED_PEAKS.json bf16 " + ("sustained" if long_step else "burst"),
                    "hbm_gbs_scanned": hbm_gbs, "frac_of_sustained_peak": achieved / pk["bf16_tflops_sustained"]}
            if (n_shard * d * 2 / 1e9) / pk["hbm_gbs"] > (flops / 1e12) / peak:
                # low arithmetic intensity (few queries per row byte): the launch is bounded by HBM, not the tensor pipe
                roof.update({"bound": "hbm", "achieved": hbm_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                             "frac": hbm_gbs / pk["hbm_gbs"], "tensor_tflops": achieved,
                             "peak_source": f"{pk_kind} MEASURED_PEAKS.json hbm_gbs"})
        else:
            kern_ms = sk_ns / max(sk_n, 1) / 1e6
            pushed = self.filter_json is not None or self.mask_words is not None
            # algorithmic bytes per launch: every ELIGIBLE fp32 row of the shard once (+ the mask bits when filtering)
            nbytes = float(self.n_elig if pushed else self.n_total) / G * d * 4 + (n_shard / 8 if pushed else 0)
            achieved = nbytes / (kern_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / pk["hbm_gbs"], "traffic": ncu_traffic(wl + "_scan") or ncu_traffic(wl),
                    "kernel": "scan_topk_kernel", "kernel_ms": kern_ms,
                    "peak_source": f"{pk_kind} MEASURED_PEAKS.json hbm_gbs",
                    "full_scan_equiv_gbs": n_shard * d * 4 / (kern_ms * 1e-3) / 1e9,
                    "eligible_rows": self.n_elig if pushed else self.n_total}
            if n_shard * d * 4 < L2_BYTES:
                roof["note"] = "database fits in L2: the launch is latency-bound, the HBM fraction is not a kernel-quality figure"
        roof["launches_timed"] = int(max(tk_n, sk_n))
        return roof

    def run(self, steps, warmup, cpu=True, e2e=True, sampler_gpu=None, cpu_budget_s=6.0):
        wl, q = self.wl, self.q
        ms, st0, st1, clocks = self.device_leg(steps, warmup, sampler_gpu)
        ms_per_step = ms / steps
        qps_full = q / (ms_per_step * 1e-3)
        out = {"name": wl, "metric": metric_name(wl), "value": self.G_units() * qps_full, "unit": "queries/s",
               "qps_full_index": qps_full, "n_gpus": self.G, "steps": steps, "warmup": warmup,
               "ms_per_step": ms_per_step, "scaling": "strong" if wl in STRONG and self.G > 1 else "weak",
               "config": config_of(wl, self.G), "roofline": self.roofline(ms, st0, st1),
               "gpu_launches": int(st1["kernel_launches"] - st0["kernel_launches"]),
               "fallback_queries": int(st1["fallback_queries"] - st0["fallback_queries"]),
               "scanned_gbs_fp32_equiv": self.n_total * self.d * 4 / (ms_per_step * 1e-3) / 1e9,
               "index_build_s": self.build_s, "host_enqueue_ms_per_step": self.host_enqueue_ms}
        if self.G > 1:
            mc = st1["merge_count"] - st0["merge_count"]
            out["exchange"] = {"kind": "peer stores from each shard's finalize kernel into the root GPU's gather block "
                                       "over NVLink + merge kernel on the root (no collective call)",
                               "merge_ms": (st1["merge_ns"] - st0["merge_ns"]) / max(mc, 1) / 1e6, "merges": int(mc),
                               "merge_ms_is": "CUDA-event time on the root GPU's stream from the last shard's arrival to the "
                                              "end of the merge kernel: includes waiting for SMs next to the root shard's "
                                              "own persistent kernels (the kernel itself runs ~15 us)",
                               "bytes_per_shard_per_step": int(q * self.kk * 12 + q * 4),
                               "kernel_ms_per_shard": [round(x, 4) for x in self.shard_kernel_ms]}
        if clocks is not None:
            out["clocks"] = clocks
        if e2e:
            out["e2e"] = self.e2e_leg(steps)
        if cpu:
            out["cpu_baseline"], _ = cpu_reference_run(wl, 1, 0, budget_s=cpu_budget_s, max_row_bytes=1.5e9)
        return out


def ingest_bench(device=0, n=1_000_000, d=768):
    """N3: bulk load of the reference's flat vector file (src/persistence/mmap.rs:13-15: u32 dim, u32
    count, then count x dim little-endian f32) through gfi_add_from_file -- file -> pinned double buffer -> H2D ->
    row_stats (exact norms, fp16 shadow rows), chunks pipelined.  Reported next to the two rates that bound it on
    this box: a pinned-host -> device copy of the same bytes (PCIe) and the device-side pass alone on rows that are
    already in HBM (gfi_add_generated: generator + row_stats)."""
    import numpy as np
    import torch
    import vectordb_from_scratch_b200 as gfi
    from vectordb_from_scratch_b200 import synth
    torch.cuda.set_device(device)
    tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else "/tmp"
    path = os.path.join(tmpdir, f"gfi_ingest_{os.getpid()}.bin")
    block = synth.gen_rows(9, 0, 50_000, d, 1)
    try:
        with open(path, "wb") as f:
            f.write(np.array([d, n], dtype="<u4").tobytes())
            for o in range(0, n, block.shape[0]):
                f.write(block[:min(block.shape[0], n - o)].tobytes())
        nbytes = n * d * 4
        idx = gfi.GpuFlatIndex(METRIC_ID["cosine"], dim=d, device=device)
        idx.reserve(n)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        got = idx.add_from_file(path)
        idx.flush()
        t_file = time.perf_counter() - t0
        assert got == n and idx.len() == n
        row = idx.get_vector(n - 1)
        assert np.array_equal(row, block[(n - 1) % block.shape[0]])
        idx.close()
    finally:
        if os.path.exists(path):
            os.remove(path)
    # the PCIe bound: the same bytes from pinned host memory in 64 MB copies
    chunk = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
    dst = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = max(1, nbytes // (64 << 20))
    for _ in range(reps):
        dst.copy_(chunk, non_blocking=True)
    torch.cuda.synchronize()
    pcie_gbs = reps * (64 << 20) / (time.perf_counter() - t0) / 1e9
    # the device-side pass alone (rows produced in HBM by the generator kernel)
    n_dev = 10_000_000
    idx = gfi.GpuFlatIndex(METRIC_ID["cosine"], dim=d, device=device)
    idx.reserve(n_dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    idx.add_generated(9, 0, n_dev, 1, 0)
    idx.flush()
    t_dev = time.perf_counter() - t0
    idx.close()
    torch.cuda.empty_cache()
    pk, _ = peaks()
    dev_bytes = n_dev * d * (4 + 4 + 4 + 2)  # generator writes fp32, row_stats reads it twice (2nd: L2) and writes fp16
    return {"name": "ingest", "metric": "rows/s bulk-loaded from the reference's flat vector file (gfi_add_from_file)",
            "value": n / t_file, "unit": "rows/s", "config": {"workload": f"bulk ingest {n} x {d} f32 from {tmpdir}", "rows": n, "dim": d},
            "file_gbs": nbytes / t_file / 1e9, "seconds": t_file, "pcie_pinned_h2d_gbs": pcie_gbs,
            "frac_of_pcie": nbytes / t_file / 1e9 / pcie_gbs,
            "device_side": {"rows": n_dev, "rows_per_s": n_dev / t_dev, "seconds": t_dev,
                            "hbm_gbs_algorithmic": dev_bytes / t_dev / 1e9, "frac_of_hbm": dev_bytes / t_dev / 1e9 / pk["hbm_gbs"],
                            "what": "gfi_add_generated: generator kernel + row_stats (sequential exact norms, fp16 shadow rows)"}}


def ranks_mode(args, wl, rank, world, local_rank, warmup):
    """Round 1's variant, kept for A/B: one process per GPU, every rank searches its shard, one NCCL all-gather of
    the packed per-rank blocks + merge kernel per step (vectordb-from-scratch_b200/sharded.py)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import vectordb_from_scratch_b200 as gfi
    from vectordb_from_scratch_b200 import synth
    from vectordb_from_scratch_b200.sharded import ShardedSearch, packed_layout, packed_views
    metric, n, d, kind, seed, q, k = WORKLOADS[wl]
    dev = torch.device("cuda", local_rank)
    idx = gfi.GpuFlatIndex(METRIC_ID[metric], dim=d, device=local_rank)
    idx.reserve(n)
    first = rank * n
    idx.add_generated(seed, first, n, kind, first)
    idx.set_option("profile", 1)
    dq = torch.from_numpy(synth.gen_rows(seed + 1, 0, q, d, kind)).to(dev)
    dks = torch.full((q,), k, dtype=torch.int32, device=dev)
    pack = torch.zeros((packed_layout(q, k)[2],), dtype=torch.uint8, device=dev)
    out_ids, out_d, out_c = packed_views(pack, q, k)
    m_ids, m_d, m_c = torch.zeros_like(out_ids), torch.zeros_like(out_d), torch.zeros_like(out_c)
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    def local_search(_q, _k):
        idx.search_device(dq.data_ptr(), q, dks.data_ptr(), k, out_ids.data_ptr(), out_d.data_ptr(),
                          out_c.data_ptr(), k, stream=stream)
        return out_ids, out_d, out_c, pack

    def merge(all_ids, all_d, all_c, _k):
        idx.merge_topk_device(all_ids.data_ptr(), all_d.data_ptr(), all_c.data_ptr(), world, q, k, dks.data_ptr(),
                              m_ids.data_ptr(), m_d.data_ptr(), m_c.data_ptr(), k, stream=stream,
                              shard_stride_bytes=pack.numel())
        return m_ids, m_d, m_c

    sharded = ShardedSearch(local_search, merge, packed=True)
    for _ in range(warmup):
        sharded.search(dq, dks)
    idx.search_status()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sharded.search(dq, dks)
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    idx.search_status()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    if rank == 0:
        print(json.dumps({"metric": metric_name(wl), "value": world * q / (ms_per_step * 1e-3), "unit": "queries/s",
                          "qps_full_index": q / (ms_per_step * 1e-3), "n_gpus": world, "steps": args.steps,
                          "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f16 tensor-core candidate pass + f32 exact rerank",
                          "data": "synthetic", "config": config_of(wl, world),
                          "sharding": "one process per GPU, NCCL all-gather + merge per step (round-1 variant)"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gfi", choices=["gfi", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["ingest"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--secondary", default="auto", help="'auto', 'none' or a comma list of workloads")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--sharding", default="inproc", choices=["inproc", "ranks"])
    ap.add_argument("--opt", action="append", default=[], help="name=value passed to gfi_set_option (experiments)")
    args = ap.parse_args()
    wl = args.workload
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3) if args.impl == "gfi" else args.warmup

    if args.impl == "reference":
        reference_arm(args, wl, rank, world)
        return
    if wl == "ingest":
        if rank == 0:
            print(json.dumps(ingest_bench(local_rank, n=2_000_000)))
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world > 1 and args.sharding == "ranks":
        ranks_mode(args, wl, rank, world, local_rank, warmup)
        dist.destroy_process_group()
        return

    # Two kinds of rendezvous: an NCCL barrier where all ranks arrive together (the contract's bracket), and a gloo
    # (host-side) wait wherever rank 0 is still working -- a pending NCCL barrier is a kernel spinning on the other
    # ranks' GPUs, which are busy serving rank 0's shards.
    host_group = None
    if world > 1:
        import datetime
        host_group = dist.new_group(backend="gloo", timeout=datetime.timedelta(minutes=60))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def host_wait():
        if world > 1:
            dist.barrier(group=host_group)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ONE index over all the GPUs of the job, owned by rank 0 (the reference server is one process); the other ranks
    # take part in the barriers and the max-over-ranks reduction only.
    devices = list(range(world))
    n_vis = torch.cuda.device_count()
    if world > 1 and n_vis < world:
        raise SystemExit(f"rank {rank}: {n_vis} visible GPUs for a {world}-GPU index (in-process sharding needs all "
                         "of them visible to rank 0; use --sharding ranks)")
    opts = dict(o.split("=") for o in args.opt)
    line = None
    if rank == 0:
        b = Bench(wl, devices, opts)
    host_wait()
    barrier()
    if rank == 0:
        line = b.run(args.steps, warmup, cpu=False, e2e=True, sampler_gpu=local_rank)
        ms_local = line["ms_per_step"] * args.steps
    else:
        ms_local = 0.0
    host_wait()
    barrier()
    ms = max_over_ranks(ms_local)
    if rank != 0:
        host_wait()  # secondaries and the sustained run happen on rank 0
        dist.destroy_process_group()
        return

    assert abs(ms - ms_local) < 1e-6
    line.pop("name")
    head = {"metric": line.pop("metric"), "value": line.pop("value"), "unit": line.pop("unit"),
            "qps_full_index": line.pop("qps_full_index"), "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": line.pop("ms_per_step"), "higher_is_better": True, "scaling": line.pop("scaling"),
            "vs_baseline": None,
            "dtype": "f16 tensor-core candidate pass (f32 accumulate) + f32 reference-exact rerank"
                     if line["roofline"]["kernel"].startswith("gemm") else "f32",
            "data": "synthetic"}
    line.pop("n_gpus"), line.pop("steps"), line.pop("warmup")
    head.update(line)
    head["value_definition"] = ("n_gpus * batch / step time: every GPU scores the batch against its own shard of the "
                                "index (weak scaling, the index grows with n_gpus); qps_full_index = batch / step time "
                                "is the rate at which whole queries are answered over the full n_gpus-shard index")
    if world > 1:
        head["sharding"] = ("one index sharded inside libgfi (gfi_create_sharded), owned by rank 0; ranks 1..N-1 join "
                            "the barriers only")

    # ---- sustained: 1000 steps of the same workload (power-capped clocks show up here, not in a 30 ms burst) ----
    if not args.no_sustained:
        ms_s, s0, s1, clk = b.device_leg(1000, 3, sampler_gpu=local_rank)
        rs = b.roofline(ms_s, s0, s1)
        head["sustained"] = {"steps": 1000, "ms_per_step": ms_s / 1000, "value": b.G_units() * b.q / (ms_s / 1000 * 1e-3),
                             "sm_mhz": clk["sm_mhz"], "sm_min_mhz": clk["sm_min_mhz"], "reasons": clk["reasons"],
                             "samples": clk["samples"], "kernel_ms": rs["kernel_ms"], "achieved": rs["achieved"],
                             "unit": rs["unit"], "frac_of_burst_peak": rs["achieved"] / peaks()[0]["bf16_tflops"]
                             if rs["bound"] == "tensor" else rs["frac"],
                             "frac_of_sustained_peak": rs.get("frac_of_sustained_peak")}
    if not args.no_cpu_baseline:
        head["cpu_baseline"], _ = cpu_reference_run(wl, 1, 0, budget_s=15.0)
    b.close()

    # ---- the other BASELINE.json configs, same contract per entry ----
    if args.secondary == "auto":
        sec = (["c3a_scan", "c3a", "c3b", "c4p1", "c4f1", "c4f50", "c5", "c1", "ingest"] if world == 1 else
               ["c3a_scan", "c3b", "c5"]) if wl == "c2" else []
    elif args.secondary == "none":
        sec = []
    else:
        sec = [s for s in args.secondary.split(",") if s]
    out_sec = []
    for s in sec:
        try:
            if s == "ingest":
                out_sec.append(ingest_bench(devices[0]))
                continue
            sb = Bench(s, devices if s != "c1" else devices[:1], opts if s == wl else {})
            st = max(5, min(args.steps, 20 if WORKLOADS[s][5] >= 64 or WORKLOADS[s][1] >= 1_000_000 else 200))
            r = sb.run(st, warmup, cpu=not args.no_cpu_baseline, e2e=True, sampler_gpu=None, cpu_budget_s=4.0)
            sb.close()
            out_sec.append(r)
        except Exception as e:  # a secondary must never take the headline down with it
            out_sec.append({"name": s, "error": f"{type(e).__name__}: {e}"[:300]})
    if out_sec:
        head["secondary"] = out_sec
    print(json.dumps(head), flush=True)
    if world > 1:
        host_wait()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
