#!/usr/bin/env python
"""One small store, a few single-query searches on the latency path (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mediquery_rag_b200 as M
from bench import synth_rows_np
n = int(sys.argv[1]) if len(sys.argv) > 1 else 154
dt = sys.argv[2] if len(sys.argv) > 2 else "fp32"
idx = M.Index(768, dt); idx.add(synth_rows_np(1, 0, n))
q = synth_rows_np(2, 0, 1)
for _ in range(20):
    s, i = idx.search(q, 5)
print(i, idx.stats()["last_search_ms"])
