#!/usr/bin/env python
"""How often does the exactness margin fail on CLUSTERED data (near-duplicate neighbourhoods, the realistic case for a
RAG corpus), i.e. how often must a tensor-path query be repaired by the exact scan?  n rows in clusters of `members`
rows around random centres (cosine to the centre ~ cos_c); queries are members plus noise.  Prints exact_fallbacks and
the time per batch for the current shortlist size (MQVS_KP_MIN widens it)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mediquery_rag_b200 as M

n, members, B, k = 1_000_000, int(sys.argv[1]), 64, 10
cos_c = float(sys.argv[2])
dtype = sys.argv[3] if len(sys.argv) > 3 else "bf16"
rng = np.random.default_rng(0)
sigma = np.sqrt((1.0 / cos_c ** 2 - 1.0) / 256.0)
centres = rng.standard_normal((n // members, 768)).astype(np.float32)
centres /= np.linalg.norm(centres, axis=1, keepdims=True)
idx = M.Index(768, dtype, capacity_rows=n)
for c0 in range(0, n // members, 200):
    c = centres[c0:c0 + 200]
    x = np.repeat(c, members, axis=0) + sigma * rng.uniform(-1, 1, (c.shape[0] * members, 768)).astype(np.float32)
    idx.add(x)
pick = rng.integers(0, n, B)
q = idx.get_rows(pick) + 0.5 * sigma * rng.uniform(-1, 1, (B, 768)).astype(np.float32)
idx.set_path(1)
es, ei = idx.search(q, k)
gap = es[:, k - 1] - idx.search(q, 32)[0][:, 31]
for path in (2, 4):
    idx.set_path(path)
    f0 = idx.stats()["exact_fallbacks"]
    s, i = idx.search(q, k)
    t0 = time.perf_counter()
    for _ in range(5):
        s, i = idx.search(q, k)
    dt = (time.perf_counter() - t0) / 5
    f1 = idx.stats()["exact_fallbacks"]
    ok = np.array_equal(i, ei) and np.array_equal(s.view(np.uint32), es.view(np.uint32))
    print(f"members={members} cos_c={cos_c} {dtype} path={path} kp_min={os.environ.get('MQVS_KP_MIN', '-')}: bit-exact={ok} "
          f"flagged per batch={(f1 - f0) / 6:.1f} of {B}, {dt * 1e3:.2f} ms per batch; median gap score_10 - score_32 = {np.median(gap):.5f}", flush=True)
