#!/usr/bin/env python
"""Where the tensor pipeline overtakes the SIMT scan on SMALL stores (4k - 64k rows) as the batch grows: blocking host-buffer
searches, microseconds per call, both paths forced (mqvs_set_path 1 / 2), results compared."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
from mediquery_rag_b200 import _native

out = []
SPACE = sys.argv[1] if len(sys.argv) > 1 else "cosine"   # "l2": raw (non-unit) synthetic rows ranked by squared distance
BATCHES = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 4, 8, 16, 64, 256, 1024]
for dtype in ("bf16", "fp32"):
    for n in (4096, 8192, 16384, 32768, 65536, 131072):
        idx = _native.Index(768, dtype, capacity_rows=n)
        if SPACE != "cosine":
            idx.set_metric(SPACE)
        idx.add_synthetic(0x5EED0002, 0, n)
        for B in BATCHES:
            q, _ = O.planted_queries(0x0BADC0DE, B, 0x5EED0002, n, 768)
            row = {"space": SPACE, "dtype": dtype, "rows": n, "B": B}
            res = {}
            for name, path in (("simt", 1), ("tensor", 2), ("auto", 0)):
                idx.set_path(path)
                for _ in range(3):
                    res[name] = idx.search(q, 10)
                reps = 20 if B <= 64 else 5
                t0 = time.perf_counter()
                for _ in range(reps):
                    idx.search(q, 10)
                row[name + "_us"] = round((time.perf_counter() - t0) / reps * 1e6, 1)
                row[name + "_pipeline"] = idx.stats()["last_pipeline"]
            row["same"] = bool(np.array_equal(res["simt"][1], res["tensor"][1]) and np.array_equal(res["simt"][0], res["tensor"][0]))
            out.append(row)
            print(row, flush=True)
        idx.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/small_store_crossover_{SPACE}.json", "w"), indent=1)
