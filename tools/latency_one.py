#!/usr/bin/env python
"""One small store, single-query searches on the latency path: host microseconds per blocking call (ncu target too).
    python tools/latency_one.py [rows] [dtype] [reps]      MQVS_NO_INLINE_QUERIES=1: the query is copied to the device first"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mediquery_rag_b200 as M
from bench import synth_rows_np
n = int(sys.argv[1]) if len(sys.argv) > 1 else 154
dt = sys.argv[2] if len(sys.argv) > 2 else "fp32"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
idx = M.Index(768, dt); idx.add(synth_rows_np(1, 0, n))
q = synth_rows_np(2, 0, 1)
for _ in range(20):
    s, i = idx.search(q, 5)
t0 = time.perf_counter()
for _ in range(reps):
    s, i = idx.search(q, 5)
us = (time.perf_counter() - t0) / reps * 1e6
print(i, "device ms", idx.stats()["last_search_ms"], "host us per Index.search", round(us, 1),
      "inline queries" if not os.environ.get("MQVS_NO_INLINE_QUERIES") else "queries copied first")
