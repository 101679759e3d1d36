"""Row-sharded search across the GPUs of one box: one process per GPU (torch.distributed, NCCL over NVLink).

SURVEY.md section 8e: rank r owns the contiguous global rows [r*S, min(N, (r+1)*S)) with S = ceil(N / W) rounded
up to a multiple of 32 (so the packed filter mask splits on word boundaries); queries are replicated; each rank
scans its shard and keeps its k best per query (exact fp32 scores, ids already global); the [B, k] candidates of
all ranks are exchanged and a W*k -> k merge (same ordering rule: score descending, ties -> lower global id)
finishes the batch on every rank, so any rank can answer.
Exchange, in order of preference:
  "p2p"   ONE libmqvs kernel per batch (mqvs_xchg_merge): P2P stores into every peer's inbox over NVLink (CUDA IPC
          mappings set up once; the 64-byte handles travel through the process group), flags, wait, merge.
  "nccl"  all_gather_into_tensor of scores and ids + mqvs_merge_candidates -- when peer mappings cannot be set up
          (MQVS_XCHG=nccl forces it), when B*k exceeds the inbox, and on CPU (gloo) in the tests.
torch is plumbing here (device buffers, streams, the process group); the scan, exchange and merge are libmqvs kernels.
"""
from __future__ import annotations

import os
from typing import Any, Callable, Optional, Tuple

import numpy as np

from . import _native


def shard_bounds(n_total: int, world: int) -> list:
    """[(lo, hi)] per rank: contiguous, word-aligned (multiples of 32), possibly empty at the tail."""
    per = (n_total + world - 1) // world if world else 0
    per = (per + 31) // 32 * 32
    return [(min(n_total, r * per), min(n_total, (r + 1) * per)) for r in range(world)]


def _native_merge(all_scores, all_ids, out_scores, out_ids) -> None:
    import torch

    _native.merge_candidates(all_scores.device.index, all_scores, all_ids, out_scores, out_ids,
                             stream=torch.cuda.current_stream().cuda_stream)


class ShardedIndex:
    def __init__(self, dim: int, store_dtype: str, n_total: int, *, group: Any = None, device: Any = None,
                 index_factory: Optional[Callable[..., Any]] = None, merge_fn: Optional[Callable[..., None]] = None,
                 metric: str = "cosine"):
        import torch
        import torch.distributed as dist

        self._torch, self._dist = torch, dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.n_total = int(n_total)
        self.bounds = shard_bounds(self.n_total, self.world)
        self.lo, self.hi = self.bounds[self.rank]
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        dev_index = self.device.index if hasattr(self.device, "index") and self.device.index is not None else 0
        self.index = (index_factory or _native.Index)(dim, store_dtype, dev_index, self.hi - self.lo)
        if metric != "cosine":
            # l2 / ip over raw (non-unit) vectors: every shard keeps its rows' norms; the ranking score of a query is
            # comparable across shards (include/mqvs.h, mqvs_set_metric), so the exchange + merge are unchanged
            self.index.set_metric(metric)
        self._merge = merge_fn or _native_merge
        self.dim = dim
        self.exchange, self._xchg, self.exchange_error = "none" if self.world == 1 else "nccl", None, None
        if (self.world > 1 and merge_fn is None and getattr(self.device, "type", "cpu") == "cuda"
                and os.environ.get("MQVS_XCHG", "p2p") != "nccl"):
            self._setup_peer_exchange(dev_index)

    def _setup_peer_exchange(self, dev_index: int) -> None:
        """Collective: every rank maps every other rank's inbox; all ranks agree on the outcome (else NCCL for all)."""
        torch, dist = self._torch, self._dist
        none = b"\0" * _native.PeerExchange.HANDLE_BYTES
        xchg, handle, err = None, none, None
        try:
            xchg = _native.PeerExchange(dev_index, self.rank, self.world)
            handle = xchg.handle
        except Exception as exc:  # IPC disabled, out of memory, ...
            err = str(exc)
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=self.group)  # every rank takes part, whatever happened above
        ok = 0
        if xchg is not None and all(h != none for h in handles):
            try:
                xchg.connect(handles)
                ok = 1
            except Exception as exc:  # no peer access between two of the devices
                err = str(exc)
        flag = torch.tensor([ok], device=self.device, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)  # also the barrier before the first exchange
        if int(flag.item()) == 1:
            self._xchg, self.exchange = xchg, "p2p"
        else:
            if xchg is not None:
                xchg.close()
            self.exchange_error = err or "a peer could not map the inboxes"

    @property
    def local_rows(self) -> int:
        return self.hi - self.lo

    def add_synthetic(self, seed: int) -> None:
        """Fill this rank's shard with its slice of the global counter-based corpus (rows lo..hi)."""
        if self.local_rows:
            self.index.add_synthetic(seed, self.lo, self.local_rows)

    def add_global(self, x_all: np.ndarray) -> None:
        """Every rank is handed the same [N, dim] host array and keeps its own slice."""
        assert x_all.shape[0] == self.n_total
        if self.local_rows:
            self.index.add(x_all[self.lo:self.hi])

    def local_mask(self, mask_global: Optional[np.ndarray]):
        if mask_global is None:
            return None
        words = np.ascontiguousarray(mask_global[self.lo // 32:(self.hi + 31) // 32], dtype=np.uint32)
        return words if words.size else np.zeros(1, np.uint32)

    def search(self, q, k: int, mask_local=None) -> Tuple[Any, Any]:
        """q: [B, dim] fp32 tensor on this rank's device (same values on every rank).
        -> (scores [B, k], global ids [B, k]) tensors on the device, identical on every rank."""
        torch, dist = self._torch, self._dist
        B = q.shape[0]
        loc_s = torch.empty((B, k), dtype=torch.float32, device=q.device)
        loc_i = torch.empty((B, k), dtype=torch.int64, device=q.device)
        stream = torch.cuda.current_stream().cuda_stream if q.is_cuda else 0
        self.index.search_device(q, k, loc_s, loc_i, mask=mask_local, id_base=self.lo, stream=stream)
        if self.world == 1:
            return loc_s, loc_i
        if self._xchg is not None and B * k <= self._xchg.max_entries and q.is_cuda:
            out_s, out_i = torch.empty_like(loc_s), torch.empty_like(loc_i)
            self._xchg.merge(loc_s, loc_i, out_s, out_i, stream=stream)
            return out_s, out_i
        all_s = torch.empty((self.world * B, k), dtype=torch.float32, device=q.device)
        all_i = torch.empty((self.world * B, k), dtype=torch.int64, device=q.device)
        dist.all_gather_into_tensor(all_s, loc_s, group=self.group)  # concatenated along dim 0: [W*B, k]
        dist.all_gather_into_tensor(all_i, loc_i, group=self.group)
        out_s, out_i = torch.empty_like(loc_s), torch.empty_like(loc_i)
        self._merge(all_s.view(self.world, B, k), all_i.view(self.world, B, k), out_s, out_i)
        return out_s, out_i
