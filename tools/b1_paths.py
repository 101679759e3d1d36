#!/usr/bin/env python
"""Batch-1 search on the row-sharded config-4 corpus (run under torchrun): the SIMT single-launch scan (path 1) against
the single-launch tensor pipeline (path 2), same box, same clocks; ms per search = max over ranks, CUDA events."""
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CORPUS_SEED, QUERY_SEED, planted_queries_np
from mediquery_rag_b200.sharded import ShardedIndex

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
W = dist.get_world_size()
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
sh = ShardedIndex(768, "bf16", rows)
sh.add_synthetic(CORPUS_SEED["c4"])
out = {"world": W, "rows": rows, "exchange": sh.exchange, "points": []}
for B in (1, 4):
    q, planted = planted_queries_np(QUERY_SEED, B, CORPUS_SEED["c4"], rows)
    qd = torch.from_numpy(q).to(dev)
    for rep in range(2):
        for path in (1, 2):
            sh.index.set_path(path)
            for _ in range(3):
                s, i = sh.search(qd, 10)
            dist.barrier(); torch.cuda.synchronize()
            assert int(i[0, 0]) == planted[0]
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20
            e0.record()
            for _ in range(n):
                sh.search(qd, 10)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
            mine = float(t.item())
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            allt = [None] * W
            dist.all_gather_object(allt, round(mine, 3))
            pt = {"B": B, "path": {1: "simt", 2: "tensor"}[path], "rep": rep, "ms_max": round(float(t.item()), 3), "ms_per_rank": allt,
                  "qps": round(B / float(t.item()) * 1e3, 1), "GBps_per_gpu_at_max": round(sh.local_rows * 1536 / 1e6 / float(t.item()), 1)}
            out["points"].append(pt)
            if rank == 0:
                print(pt, flush=True)
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/b1_paths_w{W}.json", "w"), indent=1)
dist.destroy_process_group()
