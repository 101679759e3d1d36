"""The N>1 path on CPU: world_size-2 and -3 `gloo` groups drive mediquery_rag_b200.sharded.ShardedIndex
(shard bounds, id bases, mask split, the one all-gather, merge layout) with the OracleIndex test double and an
oracle merge standing in for the two libmqvs kernels.  Sharded result == unsharded oracle result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
from mediquery_rag_b200.sharded import ShardedIndex, shard_bounds

SEED, QSEED = 0x5EED0004, 0x0BADC0DE


def test_shard_bounds_are_contiguous_word_aligned_and_cover_everything():
    for n, w in ((100_000_000, 8), (10007, 4), (31, 2), (0, 3), (64, 5), (1, 8)):
        b = shard_bounds(n, w)
        assert len(b) == w and b[0][0] == 0 and b[-1][1] == n
        assert all(lo % 32 == 0 or lo == n for lo, _ in b)
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert sum(hi - lo for lo, hi in b) == n
    assert shard_bounds(100_000_000, 8)[0] == (0, 12_500_000)


def _oracle_merge(all_s, all_i, out_s, out_i):
    W, B, k = all_s.shape
    s, i = O.merge_topk(all_s.permute(1, 0, 2).reshape(B, W * k).numpy(), all_i.permute(1, 0, 2).reshape(B, W * k).numpy(), k)
    out_s.copy_(torch.from_numpy(s))
    out_i.copy_(torch.from_numpy(i))


def _worker(rank, world, port, n, B, k, frac, ret, metric="cosine"):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from conftest import OracleIndex
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = ShardedIndex(768, "fp32", n, device=torch.device("cpu"), index_factory=OracleIndex, merge_fn=_oracle_merge,
                          metric=metric)
        sh.add_synthetic(SEED)
        q, _ = O.planted_queries(QSEED, B, SEED, n, 768)
        mask = None if frac is None else O.mask_from_bool(np.random.default_rng(5).random(n) < frac)
        s, i = sh.search(torch.from_numpy(q), k, sh.local_mask(mask))
        ret[rank] = (s.numpy().copy(), i.numpy().copy(), sh.bounds)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,n,B,k,frac", [(2, 5003, 4, 10, None), (3, 70, 3, 40, None), (2, 4000, 5, 10, 0.1)])
def test_sharded_search_equals_unsharded_oracle(world, n, B, k, frac):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, B, k, frac, ret), nprocs=world, join=True)
    st, _ = O.normalize_rows(O.synth_rows(SEED, 0, n, 768), "fp32", 768)
    q, _ = O.planted_queries(QSEED, B, SEED, n, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    mask = None if frac is None else O.mask_from_bool(np.random.default_rng(5).random(n) < frac)
    es, ei = O.brute_force_topk(qn, st, k, mask)
    for rank in range(world):  # every rank holds the full answer
        s, i, bounds = ret[rank]
        assert np.array_equal(i, ei) and np.array_equal(s.view(np.uint32), es.view(np.uint32))
        assert bounds == shard_bounds(n, world)


@pytest.mark.parametrize("metric", ["l2", "ip"])
def test_sharded_search_in_the_l2_and_ip_spaces_over_raw_vectors(metric):
    """Chroma's l2 / ip spaces on non-unit vectors: every shard ranks with its rows' norms; the ranking score of a query
    is comparable across shards, so the merged result equals the unsharded oracle in that space."""
    world, n, B, k = 2, 3001, 4, 10
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, B, k, None, ret, metric), nprocs=world, join=True)
    x = O.synth_rows(SEED, 0, n, 768)
    st, _ = O.normalize_rows(x, "fp32", 768)
    q, _ = O.planted_queries(QSEED, B, SEED, n, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    es, ei = O.brute_force_topk(qn, st, k, metric=metric, norms=O.row_norms(x, 768), q_norms=O.row_norms(q, 768))
    cs, ci = O.brute_force_topk(qn, st, k)
    assert not np.array_equal(ei, ci)  # the space matters on these rows (norms 15.5 - 16.5)
    for rank in range(world):
        s, i, _ = ret[rank]
        assert np.array_equal(i, ei) and np.array_equal(s.view(np.uint32), es.view(np.uint32))
