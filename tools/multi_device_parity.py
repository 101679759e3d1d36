#!/usr/bin/env python
"""ONE process, several REAL GPUs: mqvs_create(n_dev = W) / MediQueryVectorStore(devices=[0..W)) against the CPU oracle, bit for
bit, and against the one-process-per-GPU layout's cases (tools/multi_gpu_parity.py).  Usage: multi_device_parity.py [W]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
from mediquery_rag_b200 import MediQueryVectorStore, _native
from mediquery_rag_b200.filters import DeviceFilter, MetadataColumns, evaluate

W = int(sys.argv[1]) if len(sys.argv) > 1 else _native.device_count()
devs = list(range(W))
SEED, QSEED = 0x5EED0004, 0x0BADC0DE
results = []
for n, dtype, B, k, frac, space in [(200_003, "bf16", 1, 10, None, "cosine"), (200_003, "bf16", 37, 10, None, "cosine"),
                                    (200_003, "fp32", 64, 100, 0.2, "cosine"), (70, "fp32", 3, 40, None, "cosine"),
                                    (600_000, "bf16", 300, 10, 0.01, "cosine"), (131_072 * W + 5, "bf16", 16, 10, None, "cosine"),
                                    (500_000, "bf16", 200, 10, None, "cosine"), (250_000, "fp32", 256, 10, 0.5, "cosine"),
                                    (300_000, "bf16", 20, 300, None, "cosine"),
                                    (100_003, "fp32", 5, 10, None, "l2"), (100_003, "bf16", 9, 10, 0.3, "ip")]:
    idx = _native.Index(768, dtype, device=devs)
    if space != "cosine":
        idx.set_metric(space)
    idx.add_synthetic(SEED, 0, n)
    q, planted = O.planted_queries(QSEED, B, SEED, n, 768)
    mask = None if frac is None else O.mask_from_bool(np.random.default_rng(3).random(n) < frac)
    t0 = time.perf_counter()
    s, i = idx.search(q, k, mask)
    dt = time.perf_counter() - t0
    x = O.synth_rows(SEED, 0, n, 768)
    st, _ = O.normalize_rows(x, dtype, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    kw = {} if space == "cosine" else dict(metric=space, norms=O.row_norms(x, 768), q_norms=O.row_norms(q, 768))
    es, ei = O.brute_force_topk(qn, st, k, mask, **kw)
    ok = bool(np.array_equal(i, ei) and np.array_equal(s.view(np.uint32), es.view(np.uint32)))
    stt = idx.stats()
    results.append(dict(devices=W, n=n, dtype=dtype, B=B, k=k, mask=frac, space=space, pipeline=stt["last_pipeline"],
                        first_search_ms=round(dt * 1e3, 2), bit_exact=ok))
    print(results[-1], flush=True)
    idx.close()

# host ingest through W PCIe links at once + a device-built filter + persistence that does not depend on W
n = 400_000
x = O.synth_rows(7, 0, n, 768)
vs = MediQueryVectorStore(devices=devs, store_dtype="bf16", persist_directory="/tmp/mqvs_multi_db", auto_persist=False,
                          collection_metadata={"hnsw:space": "cosine"})
vs.delete_collection()
vs = MediQueryVectorStore(devices=devs, store_dtype="bf16", persist_directory="/tmp/mqvs_multi_db", auto_persist=False,
                          collection_metadata={"hnsw:space": "cosine"})
t0 = time.perf_counter()
vs.add_embeddings(([None] * n, x), metadata_columns={"g": np.arange(n) % 9})
ingest_s = time.perf_counter() - t0
vs.persist()
q = x[[5, 123456, 399999]] + 0.01 * x[[6, 7, 8]]
sc, rows = vs.search_batch(q, 10, filter={"g": {"$in": [1, 4]}})
st, _ = O.normalize_rows(x, "bf16", 768)
es, ei = O.brute_force_topk(O.normalize_rows(q, "fp32", 768)[0], st, 10, O.mask_from_bool(np.isin(np.arange(n) % 9, [1, 4])))
ok = bool(np.array_equal(rows, ei) and np.array_equal(sc.view(np.uint32), es.view(np.uint32)))
one = MediQueryVectorStore(device=0, persist_directory="/tmp/mqvs_multi_db", collection_metadata={"hnsw:space": "cosine"})
sc1, rows1 = one.search_batch(q, 10, filter={"g": {"$in": [1, 4]}})
results.append(dict(devices=W, case="vectorstore: host ingest, device filter, persisted on W devices, reopened on 1",
                    host_ingest_gbps=round(n * 3072 / ingest_s / 1e9, 1), bit_exact=ok,
                    reopened_equal=bool(np.array_equal(rows1, rows) and np.array_equal(sc1, sc))))
print(results[-1], flush=True)

# tombstone compaction across the devices: survivors keep their global order and change device; same bits as one device
n = 300_000 + 11
x = O.synth_rows(11, 0, n, 768) * np.float32(1.3)
keep = np.random.default_rng(9).random(n) < 0.5
keep[:20_000] = True
keep[100_000:140_000] = False
multi, single = _native.Index(768, "bf16", device=devs), _native.Index(768, "bf16", device=0)
t_c = {}
for name, idx in (("multi", multi), ("single", single)):
    idx.set_metric("l2")
    idx.add(x)
    t0 = time.perf_counter()
    assert idx.compact(O.mask_from_bool(keep)) == int(keep.sum())
    t_c[name] = time.perf_counter() - t0
every = np.arange(int(keep.sum()))
same_rows = bool(np.array_equal(multi.get_rows(every).view(np.uint32), single.get_rows(every).view(np.uint32))
                 and np.array_equal(multi.get_norms(), single.get_norms()))
xs = x[keep]
q = xs[[3, 77_777, every.size - 1]] + np.float32(0.01) * O.synth_rows(5, 0, 3, 768)
s, i = multi.search(q, 10)
es, ei = O.brute_force_topk(O.normalize_rows(q, "fp32", 768)[0], O.normalize_rows(xs, "bf16", 768)[0], 10, metric="l2",
                            norms=O.row_norms(xs, 768), q_norms=O.row_norms(q, 768))
results.append(dict(devices=W, case="tombstone compaction across devices (l2 space, 300k -> 50 %)", same_rows_as_one_device=same_rows,
                    seconds_multi=round(t_c["multi"], 3), seconds_single=round(t_c["single"], 3),
                    bit_exact=bool(same_rows and np.array_equal(i, ei) and np.array_equal(s.view(np.uint32), es.view(np.uint32)))))
print(results[-1], flush=True)
multi.close()
single.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(results, open(f"gpurun_out/multi_device_parity_w{W}.json", "w"), indent=1)
assert all(r["bit_exact"] for r in results) and all(r.get("reopened_equal", True) for r in results)
print("ALL BIT-EXACT")
