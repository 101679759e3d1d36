#!/usr/bin/env python
"""Exchange + merge step alone (run under torchrun on W GPUs): the peer-memory kernel vs NCCL all-gathers + merge
kernel, same inputs, results compared, CUDA-event time per call (max over ranks)."""
import json, os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mediquery_rag_b200 import _native
from mediquery_rag_b200.sharded import ShardedIndex

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
W = dist.get_world_size()
sh = ShardedIndex(768, "bf16", 4096 * W)
sh.add_synthetic(1)
out = {"world": W, "exchange": sh.exchange, "exchange_error": sh.exchange_error, "points": []}
for B, k in [(1, 10), (64, 10), (1024, 10), (16384, 10), (256, 100)]:
    g = torch.Generator(device="cpu").manual_seed(100 * rank + B)
    loc_s = torch.rand((B, k), generator=g).sort(dim=1, descending=True).values.to(dev)
    loc_i = (torch.randperm(B * k, generator=g).view(B, k) * W + rank).to(dev)  # globally unique ids
    def nccl():
        all_s = torch.empty((W * B, k), dtype=torch.float32, device=dev)
        all_i = torch.empty((W * B, k), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_s, loc_s)
        dist.all_gather_into_tensor(all_i, loc_i)
        o_s, o_i = torch.empty_like(loc_s), torch.empty_like(loc_i)
        _native.merge_candidates(local, all_s.view(W, B, k), all_i.view(W, B, k), o_s, o_i, stream=torch.cuda.current_stream().cuda_stream)
        return o_s, o_i
    def p2p():
        o_s, o_i = torch.empty_like(loc_s), torch.empty_like(loc_i)
        sh._xchg.merge(loc_s, loc_i, o_s, o_i, stream=torch.cuda.current_stream().cuda_stream)
        return o_s, o_i
    res = {}
    for name, fn in (("nccl", nccl),) + ((("p2p", p2p),) if sh._xchg is not None and B * k <= sh._xchg.max_entries else ()):
        for _ in range(5):
            r = fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 100
        e0.record()
        for _ in range(n):
            r = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = (float(t.item()), r)
    pt = {"B": B, "k": k, "nccl_us": round(res["nccl"][0], 1)}
    if "p2p" in res:
        pt["p2p_us"] = round(res["p2p"][0], 1)
        pt["equal"] = bool(torch.equal(res["p2p"][1][0], res["nccl"][1][0]) and torch.equal(res["p2p"][1][1], res["nccl"][1][1]))
    out["points"].append(pt)
    if rank == 0:
        print(pt, flush=True)
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/xchg_bench_w{W}.json", "w"), indent=1)
dist.destroy_process_group()
