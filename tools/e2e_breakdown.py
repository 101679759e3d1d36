#!/usr/bin/env python
"""Where the host-side microseconds of one search go (config 2, B = 64 by default): the same batch through every layer."""
import ctypes, json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mediquery_rag_b200 import MediQueryVectorStore, _native
import bench as Bm

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dtype = sys.argv[3] if len(sys.argv) > 3 else "fp32"
k, reps = 10, 200
store = MediQueryVectorStore(store_dtype=dtype, capacity_rows=rows, collection_metadata={"hnsw:space": "cosine"})
store.add_synthetic(0x5EED0002, rows)
idx = store._index
q, _ = Bm.planted_queries_np(Bm.QUERY_SEED, B, 0x5EED0002, rows)
q_pin = torch.from_numpy(q).pin_memory().numpy()
q_dev = torch.from_numpy(q).cuda()
o_s = torch.empty((B, k), dtype=torch.float32, device="cuda")
o_i = torch.empty((B, k), dtype=torch.int64, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
lib = _native.load()


def t(fn):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return round((time.perf_counter() - t0) / reps * 1e6, 1)


def dev_async_batch():
    idx.search_device(q_dev, k, o_s, o_i, stream=stream)


def dev_async_sync():
    idx.search_device(q_dev, k, o_s, o_i, stream=stream)
    torch.cuda.synchronize()


def blocking_dev():
    _native.check(lib.mqvs_search(idx._h, ctypes.c_void_p(q_dev.data_ptr()), 1, B, k, None, 0, ctypes.c_void_p(o_s.data_ptr()),
                                  ctypes.c_void_p(o_i.data_ptr()), 1))


out = {"rows": rows, "B": B, "dtype": dtype,
       "us_async_back_to_back (value path)": t(dev_async_batch),
       "us_async_plus_sync_per_call": t(dev_async_sync),
       "us_blocking_device_buffers": t(blocking_dev),
       "us_blocking_host_pinned (Index.search)": t(lambda: idx.search(q_pin, k)),
       "us_blocking_host_pageable (Index.search)": t(lambda: idx.search(q, k)),
       "us_store_search_batch_pinned": t(lambda: store.search_batch(q_pin, k)),
       "last_pipeline": idx.stats()["last_pipeline"]}
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/e2e_breakdown_{rows}_{B}_{dtype}.json", "w"), indent=1)
