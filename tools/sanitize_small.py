#!/usr/bin/env python
"""Small shapes through every kernel family (compute-sanitizer target): SIMT scan + latency path, single-launch tensor
pipeline, segmented pipeline (cold pass, radix selection, rescoring), device-side filter, peer exchange on one GPU."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mediquery_rag_b200 as M
from mediquery_rag_b200.filters import DeviceFilter, MetadataColumns
from bench import synth_rows_np

n = 40_000
x = synth_rows_np(1, 0, n)
q = synth_rows_np(2, 0, 200)
for dt in ("bf16", "fp32"):
    idx = M.Index(768, dt); idx.add(x)
    ref = None
    for path, B, k in ((1, 3, 5), (3, 5, 10), (2, 1, 10), (2, 100, 10), (4, 100, 10), (2, 200, 100), (2, 64, 128)):
        idx.set_path(path)
        s, i = idx.search(q[:B], k)
        print(dt, path, B, k, i[0, :3], idx.stats()["last_pipeline"], flush=True)
    cols = MetadataColumns(); cols.extend_columns(n, {"g": np.arange(n) % 7, "f": (np.arange(n) % 13) * 0.5})
    mask, allowed = DeviceFilter(idx, cols).build({"$and": [{"g": {"$in": [1, 3]}}, {"f": {"$lt": 4.0}}]}, None)
    idx.set_path(2)
    s, i = idx.search(q[:16], 10, mask)
    print(dt, "filtered", allowed, i[0, :3], flush=True)
    idx.close()
    # Chroma's l2 / ip spaces over the raw rows (norms ~16): K1 with norms, the SIMT scan and the single-launch tensor
    # pipeline with the space transform (one and two query tiles, a filter, n not a multiple of the tile)
    for space in ("l2", "ip"):
        idx = M.Index(768, dt); idx.set_metric(space); idx.add(x[: n - 77])
        for path, B, k in ((1, 3, 5), (2, 2, 10), (2, 100, 10), (2, 200, 100)):
            idx.set_path(path)
            s, i = idx.search(q[:B] * np.float32(0.1), k)
            print(dt, space, path, B, k, i[0, :3], idx.stats()["last_pipeline"], flush=True)
        idx.close()
import torch
W, B, k = 3, 20, 10
ex = [M._native.PeerExchange(0, r, W, max_entries=1024) for r in range(W)]
for e in ex:
    e.connect_ptrs([y.base for y in ex])
sc = [torch.rand((B, k), device="cuda").sort(dim=1, descending=True).values for _ in range(W)]
ids = [(torch.arange(B * k, device="cuda").view(B, k) * W + r) for r in range(W)]
outs = [(torch.empty((B, k), device="cuda"), torch.empty((B, k), dtype=torch.int64, device="cuda")) for _ in range(W)]
torch.cuda.synchronize()
# all ranks in ONE cooperative launch: kernels that spin on each other's flags must not be separate launches on one GPU
M._native.PeerExchange.merge_colocated(ex, sc, ids, [o[0] for o in outs], [o[1] for o in outs],
                                       stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("xchg", bool(torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][1], outs[2][1])))
for e in ex:
    e.close()
print("SANITIZE RUN DONE")
