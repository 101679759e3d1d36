#!/usr/bin/env python
"""fp32 stores: blocking host-buffer searches (Index.search) on the exact SIMT scan, on the tensor pipelines streaming the
fp32 rows (kind::tf32) and on the tensor pipelines streaming the bf16 scoring shadow (kind::f16), over store sizes and batch
sizes -- where the automatic path choice should switch (api.cu:use_tensor_path), and what the shadow buys.
    python tools/shadow_crossover.py [rows ...] > gpurun_out/shadow_crossover.json"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mediquery_rag_b200 as M
from bench import planted_queries_np

sizes = [int(v) for v in sys.argv[1:]] or [65536, 131072, 262144, 524288, 1000000]
out = []
for n in sizes:
    idx = M.Index(768, "fp32", capacity_rows=n)
    idx.add_synthetic(0x5EED0002, 0, n)
    for B in (1, 2, 4, 16, 64, 256):
        q, _ = planted_queries_np(0x0BADC0DE, B, 0x5EED0002, n)
        row = {"rows": n, "B": B}
        ref = None
        for name, path, shadow in (("simt_fp32", 1, True), ("tensor_tf32_rows", 2, False), ("tensor_bf16_shadow", 2, True), ("auto", 0, True)):
            if name == "simt_fp32" and B > 16:
                continue
            idx.set_shadow(shadow)
            idx.set_path(path)
            reps = 30 if n <= 300000 else 15
            for _ in range(3):
                s, i = idx.search(q, 10)
            t0 = time.perf_counter()
            for _ in range(reps):
                s, i = idx.search(q, 10)
            row[name + "_us"] = round((time.perf_counter() - t0) / reps * 1e6, 1)
            if name == "auto":
                row["auto_pipeline"] = idx.stats()["last_pipeline"]
            if ref is None:
                ref = (s.copy(), i.copy())
            else:
                assert np.array_equal(i, ref[1]) and np.array_equal(s.view(np.uint32), ref[0].view(np.uint32)), (n, B, name)
        out.append(row)
        print(row, file=sys.stderr, flush=True)
    idx.close()
print(json.dumps(out, indent=1))
