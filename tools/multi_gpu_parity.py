#!/usr/bin/env python
"""Real multi-GPU parity (run under torchrun on W GPUs): the row-sharded path (peer-memory exchange kernel by default,
MQVS_XCHG=nccl for the all-gather form) == the CPU oracle, bit for bit.
Covers N not divisible by W, a shard smaller than k, a where-mask, both store dtypes and both search paths."""
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O
from mediquery_rag_b200.sharded import ShardedIndex

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
W = dist.get_world_size()
SEED, QSEED = 0x5EED0004, 0x0BADC0DE
results = []
# last column: the space rows are ranked in (l2 / ip: over the raw, non-unit synthetic vectors -- norms 15.5-16.5)
for n, dtype, B, k, frac, space in [(200_003, "bf16", 1, 10, None, "cosine"), (200_003, "bf16", 37, 10, None, "cosine"),
                                    (200_003, "fp32", 64, 100, 0.2, "cosine"), (70, "fp32", 3, 40, None, "cosine"),
                                    (300_000, "bf16", 300, 10, 0.01, "cosine"), (131_072 * W + 5, "bf16", 16, 10, None, "cosine"),
                                    (250_000, "bf16", 200, 10, None, "cosine"), (250_000, "fp32", 256, 10, 0.5, "cosine"),
                                    (100_003, "fp32", 5, 10, None, "l2"), (100_003, "bf16", 9, 10, 0.3, "ip")]:
    sh = ShardedIndex(768, dtype, n, metric=space)
    sh.add_synthetic(SEED)
    q, planted = O.planted_queries(QSEED, B, SEED, n, 768)
    mask = None if frac is None else O.mask_from_bool(np.random.default_rng(3).random(n) < frac)
    lm = sh.local_mask(mask)
    lm_dev = None if lm is None else torch.from_numpy(lm.view(np.int32)).to(dev)
    s, i = sh.search(torch.from_numpy(q).to(dev), k, lm_dev)
    torch.cuda.synchronize()
    s, i = s.cpu().numpy(), i.cpu().numpy()
    if rank == 0:
        st, _ = O.normalize_rows(O.synth_rows(SEED, 0, n, 768), dtype, 768)
        qn, _ = O.normalize_rows(q, "fp32", 768)
        if space == "cosine":
            es, ei = O.brute_force_topk(qn, st, k, mask)
        else:
            es, ei = O.brute_force_topk(qn, st, k, mask, metric=space, norms=O.row_norms(O.synth_rows(SEED, 0, n, 768), 768),
                                        q_norms=O.row_norms(q, 768))
        ok = bool(np.array_equal(i, ei) and np.array_equal(s.view(np.uint32), es.view(np.uint32)))
        st_ = sh.index.stats()
        results.append(dict(world=W, n=n, dtype=dtype, B=B, k=k, mask=frac, space=space, bounds=sh.bounds[:3], path=st_["last_path"],
                            pipeline=st_["last_pipeline"], exchange=sh.exchange, exchange_error=sh.exchange_error, bit_exact=ok))
        print(results[-1], flush=True)
    # every rank must hold the same answer
    chk = torch.from_numpy(i.astype(np.int64)).to(dev)
    ref = chk.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(chk, ref), f"rank {rank} disagrees with rank 0"
    del sh
# clustered rows (dense neighbourhoods: the exactness margin fails, the band pass must repair on every shard), host-fed
n, members, B, k = 120_000, 1000, 32, 10
rng = np.random.default_rng(0)
sigma = np.sqrt((1.0 / 0.98 ** 2 - 1.0) / 256.0)
centres = rng.standard_normal((n // members, 768)).astype(np.float32)
centres /= np.linalg.norm(centres, axis=1, keepdims=True)
x = (np.repeat(centres, members, axis=0) + sigma * rng.uniform(-1, 1, (n, 768))).astype(np.float32)
q = (x[rng.integers(0, n, B)] + 0.5 * sigma * rng.uniform(-1, 1, (B, 768))).astype(np.float32)
sh = ShardedIndex(768, "bf16", n)
sh.add_global(x)
sh.index.set_path(2)
s, i = sh.search(torch.from_numpy(q).to(dev), k)
torch.cuda.synchronize()
st_local = sh.index.stats()
counts = [None] * W
dist.all_gather_object(counts, (st_local["band_repairs"], st_local["exact_fallbacks"], st_local["unrepaired"]))
if rank == 0:
    st, _ = O.normalize_rows(x, "bf16", 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    es, ei = O.brute_force_topk(qn, st, k)
    ok = bool(np.array_equal(i.cpu().numpy(), ei) and np.array_equal(s.cpu().numpy().view(np.uint32), es.view(np.uint32)))
    results.append(dict(world=W, n=n, dtype="bf16", B=B, k=k, data="clusters of 1000 at cosine 0.98", exchange=sh.exchange,
                        band_repairs_exact_fallbacks_unrepaired_per_rank=counts, bit_exact=ok))
    print(results[-1], flush=True)
del sh
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(results, open(f"gpurun_out/multi_gpu_parity_w{W}_{os.environ.get('MQVS_XCHG', 'p2p')}.json", "w"), indent=1)
    print("MULTI-GPU PARITY", "OK" if all(r["bit_exact"] for r in results) else "FAILED")
dist.destroy_process_group()
