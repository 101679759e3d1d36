#!/usr/bin/env python
"""bench.py -- queries/sec at exact top-10 on synthetic 768-d unit vectors (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--workload c2|c3|c4] [--impl reference]

A "step" is one search of a batch of B queries over the whole corpus.
  N = 1 (default)  workload c2 = BASELINE config 2: 1M x 768 fp32 corpus, k=10, one B200 (HBM-bound scan).
  N > 1 (torchrun) workload c4 = BASELINE config 4: 100M x 768 bf16 corpus row-sharded over the N ranks, k=10,
                   the [B, k] candidates exchanged + merged by ONE kernel over NVLink peer memory per batch (NCCL
                   all-gathers + merge kernel when MQVS_XCHG=nccl or peer mapping fails) ("strong": the corpus is
                   fixed, each rank scans 1/N of it).  `--workload c4 --gpus 1` runs the same corpus on one GPU.
`value` times K steps with the queries already in HBM (CUDA events, max over ranks); `e2e` times the public API
with HOST query buffers (pinned) and host results, copies inside the timed region.  `roofline` is the scan
kernel's algorithmic bytes per launch over its average launch duration, measured by CUDA events the library
records on its launch stream.  `cpu_baseline` (rank 0, N=1) and `--impl reference` time the same search as numpy
sgemm + partition on the host cores (chromadb is not installable offline; this is the arithmetic of its
brute-force path), on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CORPUS_SEED = {"c1": 0x5EED0001, "c2": 0x5EED0002, "c3": 0x5EED0003, "c4": 0x5EED0004, "c5": 0x5EED0005}
QUERY_SEED = 0x0BADC0DE
WORKLOADS = {
    "c1": dict(rows=154, dtype="fp32", k=4, batch=1,
               desc="config1: 154 chunks (data/medical_data.txt through the splitter), synthetic 768-d unit vectors, top-4, single query"),
    "c2": dict(rows=1_000_000, dtype="fp32", k=10, batch=64, desc="config2: 1M x 768 fp32, k=10, exact cosine, 1 GPU"),
    "c3": dict(rows=10_000_000, dtype="bf16", k=100, batch=4096, desc="config3: 10M x 768 bf16, k=100, 1 GPU"),
    "c4": dict(rows=100_000_000, dtype="bf16", k=10, batch=64, desc="config4: 100M x 768 bf16 row-sharded, k=10"),
    "c5": dict(rows=10_000_000, dtype="bf16", k=10, batch=16384,
               desc="config5: 10M x 768 bf16 row-sharded + Chroma-style where-filter as a fused bitmask, k=10"),
}
MASK_SEED = 0xF117E4
ELEM = {"fp32": 4, "bf16": 2}


def splitmix64_np(seed: int, idx: np.ndarray) -> np.ndarray:
    """Counter-based hash used for the synthetic metadata column of config 5 (host-side input generation)."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (idx.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synth_rows_np(seed: int, first_row: int, n: int, dim: int = 768) -> np.ndarray:
    """Host copy of the counter-based generator libmqvs uses on the device (mqvs_add_synthetic):
    v[r, c] = uniform[-1, 1) from splitmix64(seed, r * dim + c).  Keeps the measured arm free of oracle/ imports."""
    idx = (np.arange(first_row, first_row + n, dtype=np.uint64)[:, None] * np.uint64(dim)
           + np.arange(dim, dtype=np.uint64)[None, :])
    z = splitmix64_np(seed, idx)
    return (z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 8388608.0) - np.float32(1.0)


def planted_queries_np(query_seed: int, B: int, corpus_seed: int, n_rows: int, dim: int = 768, noise: float = 0.05):
    """B synthetic queries; even-numbered ones are a (known) corpus row plus 5 % noise, so their top-1 is known."""
    q = synth_rows_np(query_seed, 0, B, dim)
    planted = np.full(B, -1, np.int64)
    pick = splitmix64_np(query_seed ^ 0xA5A5A5A5, np.arange(B, dtype=np.uint64)) % np.uint64(max(n_rows, 1))
    for b in range(0, B, 2):
        t = int(pick[b])
        row = synth_rows_np(corpus_seed, t, 1, dim)[0]
        q[b] = row / np.linalg.norm(row) + np.float32(noise) * (q[b] / np.linalg.norm(q[b]))
        planted[b] = t
    return q, planted


def peaks():
    """-> dict(hbm GB/s, tensor TFLOP/s burst + sustained, source)."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm": float(p["hbm_gbs"]), "tf_burst": float(p["bf16_tflops"]),
                "tf_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback (B200_PROFILING.md: 6.65 TB/s, 1.59 PFLOP/s)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu_index = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def host_corpus(workload: str, rows: int):
    import oracle as O
    x = O.synth_rows(CORPUS_SEED[workload], 0, rows, 768)
    st, _ = O.normalize_rows(x, "fp32", 768)
    del x
    return st


def cpu_port_qps(workload: str, rows: int, batch: int, k: int, steps: int, warmup: int):
    """numpy sgemm + partition over the same rows, all host threads; returns (qps, cores, sample description)."""
    import oracle as O
    cores = len(os.sched_getaffinity(0))
    try:  # torchrun exports OMP_NUM_THREADS=1; the CPU legs are meant to use every host core
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except Exception:
        pass
    wl = WORKLOADS[workload]
    st = host_corpus(workload, rows)
    if wl["dtype"] == "bf16":  # same stored values the GPU scans: bf16-rounded rows, upcast
        st = O.stored_as_f32(O.normalize_rows(st, "bf16", 768)[0])
    q, _ = O.planted_queries(QUERY_SEED, batch, CORPUS_SEED[workload], rows, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    for _ in range(warmup):
        O.sgemm_topk(qn, st, k)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.sgemm_topk(qn, st, k)
    dt = time.perf_counter() - t0
    return batch * steps / dt, cores, dt / steps


def cpu_sample_rows(rows: int, batch: int) -> int:
    """Rows the CPU legs scan per step: the whole corpus when that is <= ~1.5 TFLOP and <= 2M rows, else a slice."""
    cap = int(1.5e12 / (2 * 768 * max(batch, 1)))
    return min(rows, max(50_000, min(2_000_000, cap)))


def run_reference(args):
    """--impl reference: the reference's CPU retrieval arithmetic (oracle port; chromadb cannot be installed
    offline here) on the host cores, same config / metric / unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    rows = wl["rows"]
    sample = f"full {rows} rows x {args.batch} queries per step"
    rows_s = cpu_sample_rows(rows, args.batch)
    if rows_s < rows:  # bound host memory/time: scan a slice, scale QPS by the row ratio
        sample = f"{rows_s} of {rows} rows x {args.batch} queries per step, QPS scaled by {rows_s}/{rows}"
    steps, warmup = max(1, min(args.steps, 8)), max(1, min(args.warmup, 2))
    qps, cores, s_per_step = cpu_port_qps(args.workload, rows_s, args.batch, wl["k"], steps, warmup)
    qps *= rows_s / rows
    line = {"impl": "reference", "metric": "queries_per_sec_exact_top%d" % wl["k"], "value": qps, "unit": "queries/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * s_per_step * rows / rows_s,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "batch": args.batch, "k": wl["k"], "rows": wl["rows"],
                       "note": "chromadb/hnswlib not installable offline; numpy sgemm + partition = its brute-force arithmetic"},
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default=None)
    ap.add_argument("--rows", type=int, default=None, help="override the corpus size (debug)")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep", type=str, default=None, help="comma-separated extra batch sizes to report")
    ap.add_argument("--selectivity", type=float, default=0.1, help="c5: fraction of rows the where-filter allows")
    ap.add_argument("--space", choices=["cosine", "l2", "ip"], default="cosine",
                    help="hnsw:space the rows are ranked in; l2 / ip rank the RAW synthetic vectors (norms 15.5-16.5) with the "
                         "stored row norms on the exact SIMT scan (SURVEY.md section 8 f3); the default is the headline config")
    ap.add_argument("--no-scaling-base", action="store_true",
                    help="N=1 default run: skip the extra config-4 (100M x 768 bf16) measurement on this one GPU")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload is None:
        args.workload = "c2" if max(args.gpus, world) == 1 else "c4"
    wl = dict(WORKLOADS[args.workload])
    if args.rows:
        wl["rows"] = args.rows
        WORKLOADS[args.workload]["rows"] = args.rows
    if args.batch is None:
        args.batch = wl["batch"]
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from mediquery_rag_b200 import _native
    from mediquery_rag_b200.sharded import ShardedIndex

    rank, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    k, B, rows = wl["k"], args.batch, wl["rows"]

    sh = ShardedIndex(768, wl["dtype"], rows, metric=args.space)
    t0 = time.perf_counter()
    sh.add_synthetic(CORPUS_SEED[args.workload])
    torch.cuda.synchronize()
    ingest_s = time.perf_counter() - t0

    q_host_np, planted = planted_queries_np(QUERY_SEED, B, CORPUS_SEED[args.workload], rows)
    q_pinned = torch.from_numpy(q_host_np).pin_memory()
    q_dev = q_pinned.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # config 5: a synthetic categorical metadata column source_id in {0..99} (hash of the row id), filtered with the
    # Chroma grammar {"source_id": {"$in": [...]}} through the product's where -> bitmask compiler
    dev_filter = {}

    def build_mask(selectivity):
        """where -> packed bitmask, evaluated ON THE DEVICE (K5 filter_eval_kernel) over the resident source_id column."""
        from mediquery_rag_b200.filters import DeviceFilter, MetadataColumns
        if "f" not in dev_filter:
            lo, hi = sh.lo, sh.hi
            ids = (splitmix64_np(MASK_SEED, np.arange(lo, hi, dtype=np.uint64)) % np.uint64(100)).astype(np.int64)
            cols = MetadataColumns()
            cols.extend_columns(hi - lo, {"source_id": ids})
            dev_filter["f"] = DeviceFilter(sh.index, cols)
        where = {"source_id": {"$in": list(range(max(1, round(selectivity * 100))))}}
        t0 = time.perf_counter()
        mask, allowed = dev_filter["f"].build(where, None)
        dev_filter["ms"] = 1e3 * (time.perf_counter() - t0)
        return mask, allowed, where

    mask_dev, mask_info = None, None
    if args.workload == "c5":
        mask_dev, allowed, where = build_mask(args.selectivity)
        build_mask(args.selectivity)  # second build: the column is resident now, this is the per-query-filter cost
        mask_info = {"where": str(where)[:60] + "...", "selectivity": args.selectivity, "allowed_rows_rank0": allowed,
                     "evaluated": "on device (filter_eval_kernel), column resident in HBM",
                     "filter_build_ms": round(dev_filter["ms"], 3)}

    def step_device(q):
        return sh.search(q, k, mask_dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    # ---- value: inputs resident in HBM
    for _ in range(args.warmup):
        out = step_device(q_dev)
    barrier()
    st0 = sh.index.stats()
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    ms_total, out = timed(lambda: step_device(q_dev), args.steps)
    clocks = sampler.stop() if sampler else None
    st1 = sh.index.stats()
    launches = (st1["kernel_launches"] - st0["kernel_launches"]) + (args.steps if world > 1 else 0)  # + exchange/merge kernel
    scan_launches = st1["scan_launches"] - st0["scan_launches"]
    scan_ms = st1["scan_ms"] - st0["scan_ms"]
    ms_per_step = ms_total / args.steps
    value = B * args.steps / (ms_total / 1e3)

    # correctness guard on the timed configuration: planted neighbours must come first
    ids = out[1].cpu().numpy()
    if args.space != "cosine":  # unit queries against rows of norm ~16: the planted row is not the l2 / ip neighbour
        sc = out[0].cpu().numpy()
        assert (ids >= 0).all() and (np.diff(sc, axis=1) <= 0).all(), "results are not sorted best-first"
    elif mask_dev is None:
        assert all(ids[b, 0] == planted[b] for b in range(0, B, 2)), "planted neighbours not found: result is wrong"
    else:  # every returned row must carry an allowed source_id
        got = ids[ids >= 0].astype(np.uint64)
        assert ((splitmix64_np(MASK_SEED, got) % np.uint64(100)) < max(1, round(args.selectivity * 100))).all()

    # ---- e2e: public API, host buffers; H2D of the queries and D2H of the results inside the timed region
    q_np_pinned = q_pinned.numpy()
    res_s = torch.empty((B, k), dtype=torch.float32).pin_memory()
    res_i = torch.empty((B, k), dtype=torch.int64).pin_memory()

    def step_e2e():
        if world == 1:
            return sh.index.search(q_np_pinned, k, mask_dev)  # mqvs_search: host in, host out, blocking
        qd = q_pinned.to(dev, non_blocking=True) if rank == 0 else torch.empty_like(q_dev)
        dist.broadcast(qd, 0)
        s, i = sh.search(qd, k, mask_dev)
        if rank == 0:
            res_s.copy_(s, non_blocking=True)
            res_i.copy_(i, non_blocking=True)
        torch.cuda.synchronize()
        return res_s, res_i

    for _ in range(args.warmup):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_qps = B * args.steps / float(e2e_s.item())

    # ---- optional sweep over other batch sizes (device-resident), reported beside the headline
    sweep = []
    for b in [int(v) for v in args.sweep.split(",")] if args.sweep else []:
        qb, _ = planted_queries_np(QUERY_SEED, b, CORPUS_SEED[args.workload], rows)
        qb_dev = torch.from_numpy(qb).to(dev)
        for _ in range(3):
            sh.search(qb_dev, k)
        s0 = sh.index.stats()
        ms_b, _ = timed(lambda: sh.search(qb_dev, k), max(5, args.steps // 2))
        s1 = sh.index.stats()
        n_steps = max(5, args.steps // 2)
        scan_b = (s1["scan_ms"] - s0["scan_ms"]) / max(1, s1["scan_launches"] - s0["scan_launches"])
        sweep.append({"batch": b, "qps": b * n_steps / (ms_b / 1e3), "ms_per_step": ms_b / n_steps,
                      "scan_gbps_per_gpu": sh.local_rows * 768 * ELEM[wl["dtype"]] / 1e9 / (scan_b / 1e3) if scan_b else None})

    sel_sweep = []
    if args.workload == "c5":
        for sel in (0.5, 0.1, 0.01):
            m_dev, allowed, _ = build_mask(sel)
            for _ in range(2):
                sh.search(q_dev, k, m_dev)
            n_steps = max(3, args.steps // 2)
            ms_s, _ = timed(lambda: sh.search(q_dev, k, m_dev), n_steps)
            sel_sweep.append({"selectivity": sel, "qps": B * n_steps / (ms_s / 1e3), "ms_per_step": ms_s / n_steps})

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    algo_bytes = sh.local_rows * 768 * ELEM[wl["dtype"]]  # corpus bytes one scan streams on this rank
    avg_launch_ms = scan_ms / max(1, scan_launches)
    kernel = {0: ("scan_simt_kernel" if args.space == "cosine" else f"scan_simt_metric_kernel ({args.space} space, row norms)")
                 + " (+ candidate memset, bracketed together)",
              1: "scan_fused_kernel (single-launch SIMT search)",
              2: "tc_score_kernel (tcgen05 scoring of all corpus segments + shortlist selection + fp32 rescoring, bracketed together)",
              3: "prep_fused_kernel + tc_fused_kernel (one cooperative launch: tcgen05 scoring of the whole corpus, in-kernel "
                 "threshold learning, selection, fp32 rescoring; bracketed together)"}[st1["last_pipeline"]]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(f"{args.workload}_b{B}" + ("" if args.space == "cosine" else f"_{args.space}"))
    # intensity = B flop per corpus byte (bf16) or B/2 (fp32): above the ridge the tensor pipe bounds the search
    ridge = pk["tf_burst"] * 1e3 / pk["hbm"] * (1 if wl["dtype"] == "bf16" else 2)
    if B >= ridge:
        flops = 2.0 * sh.local_rows * 768 * B
        achieved = flops / 1e12 / (avg_launch_ms / 1e3) if avg_launch_ms else 0.0
        roofline = {"bound": "tensor", "achieved": achieved, "peak": pk["tf_burst"], "unit": "TFLOP/s",
                    "frac": achieved / pk["tf_burst"], "traffic": traffic, "peak_source": pk["src"] + " bf16_tflops (burst)",
                    "frac_of_sustained": achieved / pk["tf_sustained"], "kernel": kernel,
                    "algorithmic_flops_per_launch": flops, "avg_launch_ms": avg_launch_ms,
                    "scan_share_of_step": scan_ms / ms_total if ms_total else None}
    else:
        achieved = algo_bytes / 1e9 / (avg_launch_ms / 1e3) if avg_launch_ms else 0.0
        roofline = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s",
                    "frac": achieved / pk["hbm"], "traffic": traffic, "peak_source": pk["src"] + " hbm_gbs",
                    "kernel": kernel, "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": avg_launch_ms,
                    "scan_share_of_step": scan_ms / ms_total if ms_total else None,
                    "frac_of_nominal_8TBps": achieved / 8000.0}

    line = {
        "metric": "queries_per_sec_exact_top%d" % k, "value": value, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "batch": B, "k": k, "rows": rows, "rows_per_gpu": sh.local_rows,
                   "store_dtype": wl["dtype"], "space": args.space,
                   "parallelism": (f"row-shard x{world}, candidate exchange: {sh.exchange}" if world > 1 else "single GPU"),
                   "l2": ("inputs larger than L2 (corpus per GPU %.2f GB vs 126 MB L2); no flush needed" % (algo_bytes / 1e9)
                          if algo_bytes > 126e6 else "corpus fits in L2 (%.1f KB): launch-latency-bound case, no flush" % (algo_bytes / 1e3)),
                   "ingest_s": round(ingest_s, 3), "scan_passes_per_step": scan_launches / args.steps},
        "roofline": roofline,
        "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": B * 768 * 4, "d2h_bytes_per_step": B * k * 12},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    # what the N = 1, 2, 4, 8 lines of one scaling run have in common (the driver compares their `value`s)
    line["scaling_note"] = (
        "value is config 2 (the metric's single-GPU configuration); the 1-GPU point of the config-4 strong-scaling curve "
        "that --gpus 2/4/8 measure is scaling_base_config4_n1.points[batch=64].qps in this line"
        if world == 1 and args.workload == "c2" else
        "strong scaling over the FIXED 100M x 768 bf16 corpus of config 4 (each rank scans 1/N of it); its 1-GPU point is "
        "scaling_base_config4_n1 in the --gpus 1 line (that line's `value` is config 2, a different workload)"
        if args.workload == "c4" else "single workload, no scaling curve")
    if sweep:
        line["sweep"] = sweep
    if mask_info:
        line["config"]["filter"] = mask_info
        line["selectivity_sweep"] = sel_sweep
    # N=1 default run only: the multi-GPU runs use config 4 (strong scaling over a fixed 100M-row corpus), so the
    # same corpus is also measured here on ONE GPU (153.6 GB of the 180 GB) -- the base point of that scaling curve.
    if world == 1 and args.workload == "c2" and not args.rows and not args.no_scaling_base and args.space == "cosine":
        try:
            del sh
            torch.cuda.empty_cache()
            c4 = WORKLOADS["c4"]
            sh4 = ShardedIndex(768, c4["dtype"], c4["rows"])
            sh4.add_synthetic(CORPUS_SEED["c4"])
            base = {"workload": c4["desc"] + " on 1 GPU", "rows": c4["rows"], "k": c4["k"], "points": []}
            for b in (1, c4["batch"]):
                qb, pl = planted_queries_np(QUERY_SEED, b, CORPUS_SEED["c4"], c4["rows"])
                qb_dev = torch.from_numpy(qb).to(dev)
                for _ in range(2):
                    o4 = sh4.search(qb_dev, c4["k"])
                torch.cuda.synchronize()
                assert all(int(o4[1][j, 0]) == pl[j] for j in range(0, b, 2))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    sh4.search(qb_dev, c4["k"])
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                base["points"].append({"batch": b, "qps": b / (ms / 1e3), "ms_per_step": ms,
                                       "scan_gbps": c4["rows"] * 768 * 2 / 1e9 / (ms / 1e3)})
            line["scaling_base_config4_n1"] = base
            del sh4
        except Exception as exc:  # e.g. not enough free HBM on a shared box
            line["scaling_base_config4_n1"] = {"unavailable": str(exc)[:200]}
    if world == 1 and not args.no_cpu_baseline:
        rows_s = cpu_sample_rows(rows, B)
        qps, cores, s_per = cpu_port_qps(args.workload, rows_s, B, k, steps=3, warmup=1)
        line["cpu_baseline"] = {"value": qps * rows_s / rows, "unit": "queries/s", "cores": cores, "kind": "port",
                                "sample": f"numpy sgemm+partition, {rows_s} rows x {B} queries x 3 steps ({s_per:.2f} s/step)"
                                          + ("" if rows_s == rows else f", scaled by {rows_s}/{rows}")}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
