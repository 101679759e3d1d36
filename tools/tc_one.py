#!/usr/bin/env python
"""One tensor-core search configuration, a few repetitions (ncu target)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
import mediquery_rag_b200 as M
n, dtype, B, k = int(sys.argv[1]), sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
path = int(sys.argv[5]) if len(sys.argv) > 5 else 0
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
idx = M.Index(768, dtype, capacity_rows=n); idx.add_synthetic(0x5EED0002, 0, n); idx.set_path(path)
q, planted = O.planted_queries(0x0BADC0DE, B, 0x5EED0002, n, 768)
for _ in range(reps):
    s, i = idx.search(q, k)
st = idx.stats()
print("planted ok", all(i[b, 0] == planted[b] for b in range(0, B, 2)), {k_: st[k_] for k_ in ("last_scan_ms", "last_search_ms", "last_path", "exact_fallbacks")})
