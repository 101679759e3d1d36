#!/usr/bin/env python
"""K1 (fused L2-normalise + pack) throughput on a B200: device-resident input (kernel roofline) and host input (e2e)."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mediquery_rag_b200 as M

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
x_dev = torch.rand((n, 768), device="cuda", dtype=torch.float32) * 2 - 1
out = []
for dtype in ("fp32", "bf16"):
    idx = M.Index(768, dtype, capacity_rows=2 * n)
    idx.add(x_dev[:1000])                       # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    idx.add(x_dev)                              # mqvs_add is blocking (K1 + non-finite check)
    dt = time.perf_counter() - t0
    bytes_moved = n * 768 * (4 + (4 if dtype == "fp32" else 2))
    out.append(dict(path="device input", store=dtype, rows=n, ms=round(dt * 1e3, 3), gbps=round(bytes_moved / dt / 1e9, 1),
                    rows_per_s=round(n / dt)))
    print(out[-1], flush=True)
    idx.close()
m = n // 2
x_host = x_dev[:m].cpu().numpy()
x_pin = torch.empty((m, 768), dtype=torch.float32).pin_memory()
x_pin.copy_(torch.from_numpy(x_host))
for name, src in (("host (pageable) input", x_host), ("host (pinned) input", x_pin.numpy())):
    for devices in ([0], [0, 0]):  # one shard; two logical shards (one host thread per shard)
        idx = M.Index(768, "bf16", device=devices, capacity_rows=n)
        idx.add(src[:1000])
        best = 1e9
        for _ in range(2):
            t0 = time.perf_counter(); idx.add(src); best = min(best, time.perf_counter() - t0)
        out.append(dict(path=name, shards=len(devices), store="bf16", rows=m, ms=round(best * 1e3, 1), rows_per_s=round(m / best),
                        h2d_gbps=round(m * 3072 / best / 1e9, 2)))
        print(out[-1], flush=True)
        idx.close()
idx2 = M.Index(768, "bf16", capacity_rows=100_000_000 // 8)
t0 = time.perf_counter(); idx2.add_synthetic(1, 0, 100_000_000 // 8); dt = time.perf_counter() - t0
out.append(dict(path="device generator + K1 (bench configs)", store="bf16", rows=100_000_000 // 8, ms=round(dt * 1e3, 1), rows_per_s=round(12.5e6 / dt)))
print(out[-1])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/ingest_check.json", "w"), indent=1)
