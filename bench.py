#!/usr/bin/env python
"""bench.py -- queries/sec at exact top-10 on synthetic 768-d unit vectors (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--workload c2|c3|c4|c5] [--impl reference]

ONE workload at every N, so the driver's 1 -> 8 curve compares like with like: BASELINE config 4, a 100M x 768 bf16 corpus
(153.6 GB: it fits the 180 GB of one B200), k = 10, headline batch 64.  N = 1 scans it on one GPU; N > 1 (torchrun, one rank
per GPU) row-shards it, every rank scans 1/N and the [B, k] candidates are exchanged + merged by ONE kernel over NVLink peer
memory per batch ("strong" scaling: the corpus is fixed).  A "step" is one search of a batch of B queries over the corpus.

  value     K steps with the queries already in HBM (CUDA events on the launch stream, max over ranks).
  e2e       the same batch through the object the reference holds -- ONE MediQueryVectorStore in ONE process over all N
            devices (`devices=[0..N)`, src/medical_engine.py:52) -- with HOST query buffers and host results, copies inside
            the timed region.  Under torchrun the ranks release their shards first and rank 0 alone drives the N GPUs.
  roofline  the scan kernel's algorithmic bytes (or flops past the ridge) per launch over its average launch duration,
            from CUDA events the library records on its launch stream; peak = MEASURED_PEAKS.json.
  points    the same corpus at B = 1, 1024 and 4096 (north_star: batch-1 HBM fraction, batch-4096 tensor fraction), each
            with its own roofline, clocks and a >= 1 s sustained loop.
  secondary N = 1: config 2 (1M x 768 fp32; B = 1, 64) and config 3 (10M x 768 bf16, B = 4096, k = 100);
            N = 8: config 5 (10M x 768 bf16 + where-filter bitmask, B = 16384).
  check     after the timed region (rank 0): every returned id is re-scored by the CPU oracle from the regenerated row
            (score bits must match), and a 2M-row slice of the corpus is brute-forced on the CPU -- no row of it may beat
            the k-th returned score without being returned.
  cpu_baseline / --impl reference: the reference's CPU retrieval on the host cores over a bounded slice of the same corpus:
            real chromadb when it is importable on the box (baseline/_ref or site-packages), else numpy sgemm + partition
            (= the arithmetic of Chroma's brute-force path); QPS scaled by slice / corpus.
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
METRIC = "queries_per_sec_exact_top%d"


# ------------------------------------------------------------------------------------------------ synthetic inputs
def splitmix64_np(seed: int, idx: np.ndarray) -> np.ndarray:
    """Counter-based hash (host copy of the generator libmqvs runs on the device)."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (idx.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synth_rows_np(seed: int, first_row: int, n: int, dim: int = 768) -> np.ndarray:
    """v[r, c] = uniform[-1, 1) from splitmix64(seed, r * dim + c): what mqvs_add_synthetic stores (before K1)."""
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
    """nvidia-smi clocks / throttle reasons DURING a timed region."""

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
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, windows=None):
        """windows: [(t0, t1)] host-clock intervals the GPU was under load; only samples that arrived inside one of them
        count (a sample describes the ~20 ms before it arrives).  None = every sample."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.03)
        self.proc.terminate()
        rows = [r for t, r in self.rows if len(r) >= 8 and (windows is None or any(a + 0.01 <= t <= b + 0.02 for a, b in windows))]
        if not rows and self.rows:  # a region shorter than one sampling period: the sample right after it
            rows = [r for _, r in self.rows[-1:] if len(r) >= 8]
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in rows if r[3].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_sample_rows(rows: int, batch: int) -> int:
    """Rows the CPU legs scan per step: the whole corpus when that is <= ~1.5 TFLOP and <= 2M rows, else a slice."""
    cap = int(1.5e12 / (2 * 768 * max(batch, 1)))
    return min(rows, max(50_000, min(2_000_000, cap)))


def host_slice(workload: str, rows: int):
    """Rows [0, rows) of the workload's corpus as the GPU stores them (oracle K1: normalised, bf16-rounded where the
    workload says so) -> (stored bits, fp32 view)."""
    import oracle as O
    wl = WORKLOADS[workload]
    x = O.synth_rows(CORPUS_SEED[workload], 0, rows, 768)
    st, _ = O.normalize_rows(x, wl["dtype"], 768)
    del x
    return st, (O.stored_as_f32(st) if wl["dtype"] == "bf16" else st)


def use_all_cores() -> int:
    cores = len(os.sched_getaffinity(0))
    try:  # torchrun exports OMP_NUM_THREADS=1; the CPU legs are meant to use every host core
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except Exception:
        pass
    return cores


def probe_chromadb():
    """The reference's real backend (langchain_chroma -> chromadb -> chroma-hnswlib, src/medical_engine.py:26,52) when the
    box has it: an offline install under baseline/_ref, or site-packages.  -> module or None."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        import chromadb  # type: ignore
        return chromadb
    except Exception:
        return None


def chroma_qps(chromadb, rows_f32: np.ndarray, qn: np.ndarray, k: int, steps: int, warmup: int):
    """Chroma's own query path (HNSW, cosine space) over the slice: collection.query(query_embeddings, n_results=k)."""
    client = chromadb.EphemeralClient()
    coll = client.create_collection("bench", metadata={"hnsw:space": "cosine"})
    n = rows_f32.shape[0]
    for a in range(0, n, 40000):
        b = min(n, a + 40000)
        coll.add(ids=[str(i) for i in range(a, b)], embeddings=rows_f32[a:b].tolist())
    for _ in range(warmup):
        coll.query(query_embeddings=qn.tolist(), n_results=k)
    t0 = time.perf_counter()
    for _ in range(steps):
        coll.query(query_embeddings=qn.tolist(), n_results=k)
    return (time.perf_counter() - t0) / steps


def cpu_arm(workload: str, batch: int, k: int, steps: int, warmup: int, rows_cap=None, want_candidates: bool = False):
    """Times the reference's CPU retrieval on a bounded slice.  -> dict(qps scaled to the full corpus, cores, kind, sample,
    s_per_step) and, optionally, (slice bits, normalised queries, the slice's top-k by the CPU) for the completeness check."""
    import oracle as O
    cores = use_all_cores()
    rows = WORKLOADS[workload]["rows"]
    rows_s = cpu_sample_rows(rows, batch) if rows_cap is None else min(rows, rows_cap)
    st, f32 = host_slice(workload, rows_s)
    q, _ = O.planted_queries(QUERY_SEED, batch, CORPUS_SEED[workload], rows, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    chroma = probe_chromadb()
    kind, how = "port", "numpy sgemm + partition (chromadb not importable on this box: no wheel offline)"
    if chroma is not None and rows_s <= 2_000_000:
        try:
            rows_c = min(rows_s, 200_000)  # HNSW build time bounds the slice
            s_per = chroma_qps(chroma, f32[:rows_c], qn, k, steps, warmup)
            kind, how, rows_s_eff = "reference", f"chromadb {getattr(chroma, '__version__', '?')} HNSW collection.query", rows_c
            out = {"qps": batch / s_per * rows_c / rows, "cores": cores, "kind": kind, "s_per_step": s_per,
                   "sample": f"{how}, {rows_c} of {rows} rows x {batch} queries x {steps} steps, QPS scaled by {rows_c}/{rows}"}
            return (out, (st, qn, None)) if want_candidates else out
        except Exception as exc:  # a broken install: fall back to the port and say so
            how = f"numpy sgemm + partition (chromadb import worked but failed: {str(exc)[:80]})"
    for _ in range(warmup):
        O.sgemm_topk(qn, f32, k)
    t0 = time.perf_counter()
    for _ in range(steps):
        cand = O.sgemm_topk(qn, f32, k)
    s_per = (time.perf_counter() - t0) / steps
    out = {"qps": batch / s_per * rows_s / rows, "cores": cores, "kind": kind, "s_per_step": s_per,
           "sample": f"{how}, {rows_s} of {rows} rows x {batch} queries x {steps} steps ({s_per:.2f} s/step)"
                     + ("" if rows_s == rows else f", QPS scaled by {rows_s}/{rows}")}
    return (out, (st, qn, cand)) if want_candidates else out


def shared_config(args, wl):
    """The part of `config` both arms print identically (the driver compares them)."""
    return {"workload": wl["desc"], "batch": args.batch, "k": wl["k"], "rows": wl["rows"], "store_dtype": wl["dtype"],
            "space": args.space}


def run_reference(args, wl):
    """--impl reference: the reference's CPU retrieval on the host cores, same config / metric / unit / steps; each step
    is a bounded sample (a slice of the corpus), QPS scaled by slice / corpus."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    res = cpu_arm(args.workload, args.batch, wl["k"], max(1, args.steps), max(1, args.warmup))
    line = {"impl": "reference", "metric": METRIC % wl["k"], "value": res["qps"], "unit": "queries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * args.batch / res["qps"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": shared_config(args, wl),
            "cpu_baseline": {"value": res["qps"], "unit": "queries/s", "cores": res["cores"], "kind": res["kind"],
                             "sample": res["sample"]},
            "e2e": {"value": res["qps"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
class Harness:
    """Timing plumbing shared by the headline, the extra points and the secondary workloads."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank, self.local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        self.host_group = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            self.host_group = dist.new_group(backend="gloo")  # host-side barriers: no kernel spins on a GPU while others work
        self.pk = peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps):
        """K calls bracketed by barrier + synchronize on both sides; CUDA events; max over ranks.  -> (ms total, last result)."""
        torch, dist = self.torch, self.dist
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        self.barrier()
        self.last_window = (t0, time.perf_counter())
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    def roofline(self, sh, dtype, B, scan_ms, scan_launches, ms_total, pipeline, space="cosine", traffic=None, streamed=None):
        pk = self.pk
        algo_bytes = sh.local_rows * 768 * ELEM[dtype]
        # fp32 stores stream a bf16 scoring shadow on the tensor pipelines (include/mqvs.h, mqvs_set_shadow): `streamed` = bytes
        # the scan kernels actually read per launch (mqvs_stats.scan_bytes); exact scores still come from the fp32 rows
        shadowed = streamed is not None and streamed < algo_bytes
        if shadowed:
            dtype = "bf16"
        avg = scan_ms / max(1, scan_launches)
        kernel = {0: "scan_simt_kernel (+ candidate memset, bracketed together)",
                  1: "scan_fused_kernel (single-launch SIMT search)",
                  2: ("tc_score_pair_kernel (cta_group::2 CTA pairs: tcgen05 scoring of all corpus segments" if B > 128 else
                      "tc_score_kernel (tcgen05 scoring of all corpus segments") + " + shortlist selection + fp32 rescoring, bracketed together)",
                  3: "prep_fused_kernel + tc_fused_kernel (one cooperative launch: tcgen05 scoring of the whole corpus, in-kernel "
                     "threshold learning, selection, fp32 rescoring; bracketed together)"}[pipeline]
        # intensity = B flop per corpus byte (bf16) or B/2 (fp32): above the ridge the tensor pipe bounds the search
        ridge = pk["tf_burst"] * 1e3 / pk["hbm"] * (1 if dtype == "bf16" else 2)
        share = scan_ms / ms_total if ms_total else None
        if B >= ridge:
            flops = 2.0 * sh.local_rows * 768 * B
            ach = flops / 1e12 / (avg / 1e3) if avg else 0.0
            return {"bound": "tensor", "achieved": ach, "peak": pk["tf_burst"], "unit": "TFLOP/s", "frac": ach / pk["tf_burst"],
                    "traffic": traffic, "peak_source": pk["src"] + " bf16_tflops (burst)", "frac_of_sustained": ach / pk["tf_sustained"],
                    "kernel": kernel, "algorithmic_flops_per_launch": flops, "avg_launch_ms": avg, "scan_share_of_step": share}
        if shadowed:
            # both accountings: `achieved` / `frac` = the bytes actually streamed (the bf16 shadow) over the kernel time, i.e. what
            # the memory system does; `vs_fp32_algorithmic_bytes` = SURVEY section 8(d)'s figure for the fp32 store over the same time
            # (what a scan of the stored fp32 rows would have to sustain -- above the HBM peak is the point of the shadow)
            ach = streamed / 1e9 / (avg / 1e3) if avg else 0.0
            eq = algo_bytes / 1e9 / (avg / 1e3) if avg else 0.0
            return {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"], "traffic": traffic,
                    "peak_source": pk["src"] + " hbm_gbs", "kernel": kernel + " over the bf16 scoring shadow of the fp32 rows",
                    "algorithmic_bytes_per_launch": algo_bytes, "streamed_bytes_per_launch": streamed,
                    "vs_fp32_algorithmic_bytes": {"achieved": eq, "frac": eq / pk["hbm"]},
                    "avg_launch_ms": avg, "scan_share_of_step": share, "frac_of_nominal_8TBps": ach / 8000.0}
        ach = algo_bytes / 1e9 / (avg / 1e3) if avg else 0.0
        return {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"], "traffic": traffic,
                "peak_source": pk["src"] + " hbm_gbs", "kernel": kernel, "algorithmic_bytes_per_launch": algo_bytes,
                "avg_launch_ms": avg, "scan_share_of_step": share, "frac_of_nominal_8TBps": ach / 8000.0}

    def point(self, sh, workload, B, k, steps, warmup, mask=None, sustained_s=1.0, space="cosine"):
        """One (workload, batch) measurement with device-resident queries: W warm-up steps, K timed steps with the clocks
        sampled meanwhile, then a sustained loop of at least `sustained_s` seconds (its own clocks).  -> (dict, last result)."""
        torch = self.torch
        wl = WORKLOADS[workload]
        q_np, planted = planted_queries_np(QUERY_SEED, B, CORPUS_SEED[workload], wl["rows"])
        q_dev = torch.from_numpy(q_np).to(self.dev)
        fn = lambda: sh.search(q_dev, k, mask)  # noqa: E731
        for _ in range(warmup):
            out = fn()
        self.barrier()
        st0 = sh.index.stats()
        sampler = ClockSampler(self.local_rank).start() if self.rank == 0 else None
        ms_total, out = self.timed(fn, steps)
        windows = [self.last_window]
        st1 = sh.index.stats()
        res = {"batch": B, "k": k, "steps": steps, "warmup": warmup, "ms_per_step": ms_total / steps,
               "qps": B * steps / (ms_total / 1e3), "clocks": None,
               "launches": int(st1["kernel_launches"] - st0["kernel_launches"]) + (steps if self.world > 1 else 0),
               "scan_passes_per_step": (st1["scan_launches"] - st0["scan_launches"]) / steps,
               "unrepaired": int(st1["unrepaired"] - st0["unrepaired"]),
               "roofline": self.roofline(sh, wl["dtype"], B, st1["scan_ms"] - st0["scan_ms"],
                                         st1["scan_launches"] - st0["scan_launches"], ms_total, st1["last_pipeline"], space,
                                         streamed=(st1["scan_bytes"] - st0["scan_bytes"]) / max(1, st1["scan_launches"] - st0["scan_launches"]))}
        if sustained_s > 0:
            n_sus = max(steps, int(np.ceil(sustained_s * 1e3 / max(res["ms_per_step"], 1e-3))))
            ms_sus, _ = self.timed(fn, n_sus)
            windows.append(self.last_window)
            res["sustained"] = {"steps": n_sus, "seconds": ms_sus / 1e3, "qps": B * n_sus / (ms_sus / 1e3),
                                "ms_per_step": ms_sus / n_sus}
        if sampler:
            # one nvidia-smi stream over the K timed steps and the sustained loop right after them; only samples that
            # arrived while the GPU was inside one of the two regions count
            res["clocks"] = sampler.stop(windows)
            res["clocks"]["covers"] = "the K timed steps + the sustained loop that follows them (samples taken under load only)"
        assert res["unrepaired"] == 0, "the asynchronous search left flagged queries unrepaired"
        return res, out, (q_np, planted)


def check_results(workload, q_np, planted, scores, ids, k, cpu_slice=None, mask_fn=None):
    """Rank 0, after the timed region: (a) planted neighbours come first; (b) every returned id re-scored by the CPU oracle
    from the regenerated row -- the score bits must match and the rows must be in order; (c) completeness against a CPU
    brute force over a slice of the corpus: no row of the slice may beat the k-th returned (score, id) unless returned."""
    import oracle as O
    wl = WORKLOADS[workload]
    B = q_np.shape[0]
    out = {"planted_first": True, "rescored_ids": 0, "slice_rows": 0, "slice_candidates": 0}
    if mask_fn is None:
        assert all(ids[b, 0] == planted[b] for b in range(0, B, 2)), "planted neighbours not found: result is wrong"
    else:
        got = ids[ids >= 0]
        assert mask_fn(got).all(), "a returned row is not allowed by the filter"
    qn, _ = O.normalize_rows(q_np, "fp32", 768)
    nb = min(B, 64)  # (regenerating rows one by one is host work: a 64-query sample of large batches)
    pick = np.linspace(0, B - 1, nb).astype(int)
    for b in pick:
        row_ids = ids[b][ids[b] >= 0]
        x = np.concatenate([O.synth_rows(CORPUS_SEED[workload], int(r), 1, 768) for r in row_ids])
        st, _ = O.normalize_rows(x, wl["dtype"], 768)
        es, ei = O.brute_force_topk(qn[b:b + 1], st, len(row_ids))  # (ties -> lower position = lower id: the order returned)
        assert np.array_equal(ei[0], np.arange(len(row_ids))), f"query {b}: returned rows are not in exact-score order"
        assert np.array_equal(es[0].view(np.uint32), scores[b][:len(row_ids)].view(np.uint32)), f"query {b}: score bits differ from the oracle"
        out["rescored_ids"] += len(row_ids)
    if cpu_slice is not None and cpu_slice[2] is not None:
        st_slice, qn_s, (cs, ci) = cpu_slice
        out["slice_rows"] = int(st_slice.shape[0])
        for b in pick:
            cand = ci[b][ci[b] >= 0]
            if mask_fn is not None:
                cand = cand[mask_fn(cand)]
            if cand.size == 0:
                continue
            es, ei = O.brute_force_topk(qn[b:b + 1], st_slice[cand], cand.size)  # canonical fp32 scores of the slice's best rows
            kth_s, kth_i = scores[b, k - 1], ids[b, k - 1]
            returned = set(ids[b].tolist())
            for s, j in zip(es[0], ei[0]):
                r = int(cand[j])
                beats = kth_i < 0 or s > kth_s or (s == kth_s and r < kth_i)
                assert not beats or r in returned, f"query {b}: corpus row {r} (score {s}) beats the k-th result but was not returned"
            out["slice_candidates"] += int(cand.size)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="c4")
    ap.add_argument("--rows", type=int, default=None, help="override the corpus size (debug)")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--points", type=str, default=None, help="comma-separated extra batch sizes on the headline corpus (default 1,1024,4096)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary workloads (config 2 / 3 at N=1, config 5 at N=8)")
    ap.add_argument("--no-store-e2e", action="store_true", help="N>1: measure e2e through the per-rank path instead of one "
                                                                "MediQueryVectorStore over all devices")
    ap.add_argument("--selectivity", type=float, default=0.1, help="c5: fraction of rows the where-filter allows")
    ap.add_argument("--space", choices=["cosine", "l2", "ip"], default="cosine",
                    help="hnsw:space the rows are ranked in; l2 / ip rank the RAW synthetic vectors (norms 15.5-16.5); "
                         "the default is the headline config")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.rows:
        wl["rows"] = args.rows
        WORKLOADS[args.workload]["rows"] = args.rows
    if args.batch is None:
        args.batch = wl["batch"]
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args, wl)

    H = Harness(args)
    torch, dist, world, rank, dev = H.torch, H.dist, H.world, H.rank, H.dev
    from mediquery_rag_b200 import MediQueryVectorStore
    from mediquery_rag_b200.filters import DeviceFilter, MetadataColumns
    from mediquery_rag_b200.sharded import ShardedIndex

    k, B, rows = wl["k"], args.batch, wl["rows"]
    headline = args.workload
    secondary = {}

    def source_ids(lo, hi):
        return (splitmix64_np(MASK_SEED, np.arange(lo, hi, dtype=np.uint64)) % np.uint64(100)).astype(np.int64)

    def allowed_fn(sel):
        return lambda got: (splitmix64_np(MASK_SEED, got.astype(np.uint64)) % np.uint64(100)) < max(1, round(sel * 100))

    def build_filter(sh, sel, state):
        """where -> packed bitmask, evaluated ON THE DEVICE (K5 filter_eval_kernel) over the resident source_id column."""
        if "f" not in state:
            cols = MetadataColumns()
            cols.extend_columns(sh.hi - sh.lo, {"source_id": source_ids(sh.lo, sh.hi)})
            state["f"] = DeviceFilter(sh.index, cols)
        where = {"source_id": {"$in": list(range(max(1, round(sel * 100))))}}
        t0 = time.perf_counter()
        mask, allowed = state["f"].build(where, None)
        state["ms"] = 1e3 * (time.perf_counter() - t0)
        return mask, allowed, where

    def run_secondary(name, batches, steps):
        """A secondary workload on its own corpus (freed afterwards): every batch size with roofline, clocks, sustained loop."""
        w = WORKLOADS[name]
        sh = ShardedIndex(768, w["dtype"], w["rows"])
        sh.add_synthetic(CORPUS_SEED[name])
        torch.cuda.synchronize()
        block = {"workload": w["desc"], "rows": w["rows"], "rows_per_gpu": sh.local_rows, "k": w["k"], "store_dtype": w["dtype"],
                 "points": []}
        mask, mask_fn, fstate = None, None, {}
        if name == "c5":
            mask, allowed, where = build_filter(sh, args.selectivity, fstate)
            build_filter(sh, args.selectivity, fstate)  # second build: the column is resident now = the per-filter cost
            mask_fn = allowed_fn(args.selectivity)
            block["filter"] = {"where": str(where)[:60] + "...", "selectivity": args.selectivity, "allowed_rows_rank0": allowed,
                               "evaluated": "on device (filter_eval_kernel), column resident in HBM",
                               "filter_build_ms": round(fstate["ms"], 3)}
        for b in batches:
            res, out, (q_np, planted) = H.point(sh, name, b, w["k"], steps, 3, mask)
            if rank == 0:
                res["check"] = check_results(name, q_np, planted, out[0].cpu().numpy(), out[1].cpu().numpy(), w["k"], None, mask_fn)
            if world == 1 and mask is None:
                # the same batch as a blocking call with HOST buffers (pinned queries in, host results out), wall clock
                q_host = torch.from_numpy(q_np).pin_memory().numpy()
                for _ in range(3):
                    got = sh.index.search(q_host, w["k"])
                t0 = time.perf_counter()
                for _ in range(steps):
                    got = sh.index.search(q_host, w["k"])
                dt = time.perf_counter() - t0
                assert np.array_equal(got[1], out[1].cpu().numpy()) and \
                    np.array_equal(got[0].view(np.uint32), out[0].cpu().numpy().view(np.uint32)), "e2e results differ from the device-resident path"
                res["e2e"] = {"value": b * steps / dt, "unit": "queries/s", "ms_per_step": 1e3 * dt / steps,
                              "h2d_bytes_per_step": b * 768 * 4, "d2h_bytes_per_step": b * w["k"] * 12,
                              "through": "mqvs_search (blocking): pinned host queries -> host results"}
            block["points"].append(res)
        sh.index.close()
        del sh
        torch.cuda.empty_cache()
        return block

    def run_config1():
        """Config 1, the reference's own regime: 154 chunks (the shipped knowledge base's size), ONE query, k = 5
        (src/agents/nodes.py:93), through the object the reference holds -- microseconds per blocking search with host buffers,
        the CPU port beside it.  Launch-latency-bound: one kernel launch (the query rides in its parameter block) + one sync."""
        import oracle as O
        n, kk = 154, 5
        x = synth_rows_np(CORPUS_SEED["c1"], 0, n)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        vs = MediQueryVectorStore(devices=[H.local_rank], store_dtype="fp32")  # (default hnsw:space "l2", as the reference's)
        vs.add_embeddings(([f"chunk {i}" for i in range(n)], x))
        q = (x[77] + 0.05 * synth_rows_np(QUERY_SEED, 0, 1)[0] / 27.7).astype(np.float32)
        reps = 2000
        for _ in range(50):
            vs.search_batch(q[None, :], kk)
        t0 = time.perf_counter()
        for _ in range(reps):
            s_, i_ = vs.search_batch(q[None, :], kk)
        us_batch = (time.perf_counter() - t0) / reps * 1e6
        t0 = time.perf_counter()
        for _ in range(reps):
            docs = vs.similarity_search_by_vector(q, k=kk)
        us_docs = (time.perf_counter() - t0) / reps * 1e6
        st, _ = O.normalize_rows(x, "fp32", 768)
        qn, _ = O.normalize_rows(q[None, :], "fp32", 768)
        es, ei = O.brute_force_topk(qn, st, kk)
        assert np.array_equal(i_, ei) and np.array_equal(s_.view(np.uint32), es.view(np.uint32)) and docs[0].page_content == "chunk 77"
        use_all_cores()
        f32 = np.ascontiguousarray(st)
        for _ in range(50):
            O.sgemm_topk(qn, f32, kk)
        t0 = time.perf_counter()
        for _ in range(reps):
            O.sgemm_topk(qn, f32, kk)
        us_cpu = (time.perf_counter() - t0) / reps * 1e6
        stats = vs._index.stats()
        vs._index.close()
        return {"workload": WORKLOADS["c1"]["desc"], "rows": n, "k": kk, "batch": 1, "store_dtype": "fp32",
                "e2e_us_per_search": round(us_batch, 1), "e2e_qps": round(1e6 / us_batch, 1),
                "e2e_us_per_similarity_search_by_vector": round(us_docs, 1),
                "through": "MediQueryVectorStore.search_batch / similarity_search_by_vector: host query -> ONE launch of the latency kernel "
                           "(scan_fused_inline_kernel: the query travels in the kernel's parameter block; results written to mapped "
                           "pinned memory) -> one stream sync -> host results",
                "launches_per_search": 1, "last_pipeline": stats["last_pipeline"], "bit_exact_vs_oracle": True,
                "cpu_port_us_per_search": round(us_cpu, 1),
                "note": "launch-latency-bound: below ~1e4 rows a search costs one launch + one sync whatever the store holds"}

    do_secondary = not args.no_secondary and headline == "c4" and not args.rows and args.space == "cosine"
    if do_secondary and world == 1:
        try:
            secondary["config1"] = run_config1()
        except Exception as exc:
            secondary["config1"] = {"error": str(exc)[:300]}
        try:
            secondary["config2"] = run_secondary("c2", (1, 64), max(10, args.steps))
            secondary["config3"] = run_secondary("c3", (4096,), 5)
        except Exception as exc:  # e.g. not enough free HBM on a shared box: the headline still runs
            secondary["error"] = str(exc)[:300]

    # ---- headline corpus.  N = 1: behind the MediQueryVectorStore the e2e leg searches through (same handle, no second copy);
    # N > 1: one shard per rank for `value`, the single-process store is built afterwards for `e2e`
    t0 = time.perf_counter()
    store = None
    if world == 1:
        store = MediQueryVectorStore(devices=[H.local_rank], store_dtype=wl["dtype"], capacity_rows=rows,
                                     collection_metadata={"hnsw:space": args.space})
        store.add_synthetic(CORPUS_SEED[headline], rows)
        sh = ShardedIndex(768, wl["dtype"], rows, index_factory=lambda *a: store._index, metric=args.space)
    else:
        sh = ShardedIndex(768, wl["dtype"], rows, metric=args.space)
        sh.add_synthetic(CORPUS_SEED[headline])
    torch.cuda.synchronize()
    ingest_s = time.perf_counter() - t0

    mask, mask_info, mask_fn, fstate = None, None, None, {}
    if headline == "c5":
        mask, allowed, where = build_filter(sh, args.selectivity, fstate)
        build_filter(sh, args.selectivity, fstate)
        mask_fn = allowed_fn(args.selectivity)
        mask_info = {"where": str(where)[:60] + "...", "selectivity": args.selectivity, "allowed_rows_rank0": allowed,
                     "evaluated": "on device (filter_eval_kernel), column resident in HBM", "filter_build_ms": round(fstate["ms"], 3)}

    head, out, (q_np, planted) = H.point(sh, headline, B, k, args.steps, args.warmup, mask, space=args.space)
    head_scores, head_ids = out[0].cpu().numpy(), out[1].cpu().numpy()

    points = []
    extra = [int(v) for v in args.points.split(",")] if args.points else ([1, 1024, 4096] if headline == "c4" else [])
    for b in extra:
        if b == B:
            continue
        per_step_s = 2.0 * sh.local_rows * 768 * b / 1.1e15 + sh.local_rows * 768 * ELEM[wl["dtype"]] / 6.5e12
        steps_b = max(3, min(args.steps, int(2.0 / per_step_s)))
        res, o, (qb, pl) = H.point(sh, headline, b, k, steps_b, 3, mask, space=args.space)
        if rank == 0 and args.space == "cosine":
            res["check"] = check_results(headline, qb, pl, o[0].cpu().numpy(), o[1].cpu().numpy(), k, None, mask_fn)
        points.append(res)

    # ---- e2e: the public object, host buffers; H2D of the queries and D2H of the results inside the timed region
    q_pinned = torch.from_numpy(q_np).pin_memory()
    q_host = q_pinned.numpy()
    e2e_how, e2e_extra = None, {}

    def e2e_loop(fn):
        for _ in range(args.warmup):
            fn()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            got = fn()
        return time.perf_counter() - t0, got

    if world == 1:
        H.barrier()
        if mask is None:
            dt, got = e2e_loop(lambda: store.search_batch(q_host, k))
        else:
            dt, got = e2e_loop(lambda: sh.index.search(q_host, k, mask))
        e2e_qps = B * args.steps / dt
        e2e_how = "MediQueryVectorStore.search_batch (one handle, this GPU): pinned host queries -> mqvs_search -> host results"
        assert np.array_equal(got[1], head_ids) and np.array_equal(got[0].view(np.uint32), head_scores.view(np.uint32)), \
            "e2e results differ from the device-resident path"
    elif args.no_store_e2e or mask is not None:
        res_s = torch.empty((B, k), dtype=torch.float32).pin_memory()
        res_i = torch.empty((B, k), dtype=torch.int64).pin_memory()
        q_dev = torch.empty((B, 768), dtype=torch.float32, device=dev)

        def step():
            qd = q_pinned.to(dev, non_blocking=True) if rank == 0 else q_dev
            dist.broadcast(qd, 0)
            s, i = sh.search(qd, k, mask)
            if rank == 0:
                res_s.copy_(s, non_blocking=True)
                res_i.copy_(i, non_blocking=True)
            torch.cuda.synchronize()

        for _ in range(args.warmup):
            step()
        H.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        H.barrier()
        t = torch.tensor([time.perf_counter() - t0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_qps = B * args.steps / float(t.item())
        e2e_how = "one process per GPU: rank 0 copies pinned queries in, NCCL broadcast, sharded search + peer exchange, results to pinned host"
    else:
        # every rank releases its shard; rank 0 alone then drives all N GPUs through ONE MediQueryVectorStore, the way the
        # reference holds its vector store (one object, one process); the other ranks wait on a host-side barrier
        xchg_kind = sh.exchange
        H.barrier()  # nobody unmaps or frees an inbox a peer's exchange kernel may still be writing
        sh.index.close()
        if sh._xchg is not None:
            sh._xchg.close()
        torch.cuda.empty_cache()
        H.barrier()
        e2e_qps = 0.0
        if rank == 0:
            t0 = time.perf_counter()
            store = MediQueryVectorStore(devices=list(range(world)), store_dtype=wl["dtype"], capacity_rows=rows,
                                         collection_metadata={"hnsw:space": args.space})
            store.add_synthetic(CORPUS_SEED[headline], rows)
            e2e_extra["store_ingest_s"] = round(time.perf_counter() - t0, 3)
            dt, got = e2e_loop(lambda: store.search_batch(q_host, k))
            e2e_qps = B * args.steps / dt
            assert np.array_equal(got[1], head_ids) and np.array_equal(got[0].view(np.uint32), head_scores.view(np.uint32)), \
                "the single-process store over all devices disagrees with the per-rank shards"
            # the same object with device-resident queries (mqvs_search_async on the multi-device handle)
            qd = torch.from_numpy(q_np).to(dev)
            o_s = torch.empty((B, k), dtype=torch.float32, device=dev)
            o_i = torch.empty((B, k), dtype=torch.int64, device=dev)
            stream = torch.cuda.current_stream().cuda_stream
            for _ in range(args.warmup):
                store._index.search_device(qd, k, o_s, o_i, stream=stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                store._index.search_device(qd, k, o_s, o_i, stream=stream)
            e1.record()
            torch.cuda.synchronize()
            e2e_extra["single_process_device_resident_qps"] = B * args.steps / (e0.elapsed_time(e1) / 1e3)
            assert np.array_equal(o_i.cpu().numpy(), head_ids)
            e2e_extra["single_process_stats"] = {kk: vv for kk, vv in store._index.stats().items()
                                                 if kk in ("kernel_launches", "scan_launches", "last_pipeline", "unrepaired")}
            store._index.close()
        dist.barrier(group=H.host_group)
        e2e_how = (f"ONE MediQueryVectorStore(devices=[0..{world})) in ONE process (rank 0; the other ranks released their shards): "
                   "pinned host queries -> mqvs_search on the multi-device handle (per-device scans, results stored into the home "
                   "device over NVLink peer access, one merge kernel) -> host results")
        e2e_extra["per_rank_exchange"] = xchg_kind
        sh = None

    # ---- N = 8: config 5 as a secondary block (after the headline corpus is gone)
    if do_secondary and world == 8:
        try:
            secondary["config5"] = run_secondary("c5", (16384,), 5)
        except Exception as exc:
            secondary["error"] = str(exc)[:300]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tkey = f"{headline}_b{B}" + ("" if args.space == "cosine" else f"_{args.space}")
            tdb = json.load(f)  # ncu dram__bytes_read + write per launch of the dominant kernel; per-N entries where the shard differs
            traffic = tdb.get(f"{tkey}_n{world}", tdb.get(tkey) if world == 1 and not args.rows else None)
    roofline = dict(head["roofline"])
    roofline["traffic"] = traffic
    local_rows = (rows + world - 1) // world
    algo_bytes = local_rows * 768 * ELEM[wl["dtype"]]
    config = shared_config(args, wl)
    line = {
        "metric": METRIC % k, "value": head["qps"], "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config,
        "details": {"rows_per_gpu": local_rows,
                    "parallelism": (f"row-shard x{world}, one process per GPU, candidate exchange: {e2e_extra.get('per_rank_exchange', getattr(sh, 'exchange', 'none'))}"
                                    if world > 1 else "single GPU"),
                    "l2": "inputs larger than L2 (corpus per GPU %.2f GB vs 126 MB L2); no flush needed" % (algo_bytes / 1e9),
                    "ingest_s": round(ingest_s, 3), "scan_passes_per_step": head["scan_passes_per_step"],
                    "scaling_note": "strong scaling over the FIXED corpus of the workload: the same corpus at every N (each rank scans 1/N)"},
        "roofline": roofline,
        "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": B * 768 * 4 * (world if world > 1 and not args.no_store_e2e else 1),
                "d2h_bytes_per_step": B * k * 12, "through": e2e_how, **e2e_extra},
        "gpu_launches": head["launches"], "clocks": head["clocks"], "sustained": head.get("sustained"),
        "points": points,
    }
    if secondary:
        line["secondary"] = secondary
    if mask_info:
        line["details"]["filter"] = mask_info
    # ---- CPU leg (rank 0): baseline at N = 1, and at every N the slice brute force the completeness check needs
    cpu_slice = None
    if not args.no_cpu_baseline and args.space == "cosine":
        res, cpu_slice = cpu_arm(headline, B, k, steps=3, warmup=1, want_candidates=True)
        if world == 1:
            line["cpu_baseline"] = {"value": res["qps"], "unit": "queries/s", "cores": res["cores"], "kind": res["kind"],
                                    "sample": res["sample"]}
    line.setdefault("cpu_baseline", None)
    if args.space == "cosine":
        line["check"] = check_results(headline, q_np, planted, head_scores, head_ids, k, cpu_slice, mask_fn)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
