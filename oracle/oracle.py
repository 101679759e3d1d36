"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.  numpy/ctypes face of the CPU restatement.

What each function restates (reference = /root/reference, all of whose vector arithmetic sits in
un-vendored chromadb/hnswlib -> PARITY UNPINNED, see oracle/__init__.py):

* normalize_rows      -- rows as the reference stores them: Ollama `/api/embed` L2-normalised vectors handed to
                         Chroma.from_documents (src/ingest_medical.py:104-110); new backend's K1.
* brute_force_topk    -- vectorstore.similarity_search(query, k) (src/agents/nodes.py:93,
                         src/medical_engine.py:70) as an EXACT scan in the canonical fp32 order.
* referee_topk_f64    -- same, in float64 from the stored values (the independent referee).
* sgemm_topk          -- same, as numpy sgemm + argpartition: the arithmetic of chromadb's brute-force
                         buffer; this is what bench.py times as the CPU baseline.
* eval_where          -- Chroma `where` filter grammar ($eq/$ne/$gt/$gte/$lt/$lte/$in/$nin/$and/$or), naive.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Iterable, Mapping, Sequence

import numpy as np

DIM_ALIGN = 128
N_THREADS = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib() -> ctypes.CDLL:
    """Load (building if needed) oracle/liboracle.so."""
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "canonical.c")
    src2 = os.path.join(_HERE, "hnswlib_restated.c")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(src2)):
        subprocess.run(["make", "-s", "-B", "-C", _HERE, "liboracle.so"], check=True)
    L = ctypes.CDLL(so)
    i64, i32, vp, u64 = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64
    L.orc_normalize_f32.argtypes = [vp, i64, i32, i32, vp]
    L.orc_normalize_f32.restype = i64
    L.orc_normalize_bf16.argtypes = [vp, i64, i32, i32, vp]
    L.orc_normalize_bf16.restype = i64
    L.orc_scores.argtypes = [vp, i32, vp, i32, i64, i32, vp]
    L.orc_scores.restype = None
    L.orc_topk.argtypes = [vp, i32, vp, i32, i64, i32, vp, i32, i64, vp, vp]
    L.orc_topk.restype = None
    L.orc_topk_metric.argtypes = [vp, i32, vp, i32, i64, i32, vp, i32, i64, i32, vp, vp, vp, vp]
    L.orc_topk_metric.restype = None
    L.orc_norms_f32.argtypes = [vp, i64, i32, i32, vp]
    L.orc_norms_f32.restype = None
    L.orc_synth_rows.argtypes = [u64, i64, i64, i32, vp]
    L.orc_synth_rows.restype = None
    L.hnsw_distances.argtypes = [i32, i32, vp, i64, i64, vp, vp]
    L.hnsw_distances.restype = None
    L.hnsw_bruteforce_knn.argtypes = [vp, i64, vp, i32, vp, vp]
    L.hnsw_bruteforce_knn.restype = None
    _LIB = L
    return L


def _p(a: np.ndarray) -> ctypes.c_void_p:
    return ctypes.c_void_p(a.ctypes.data)


def _row_chunks(n: int, align: int = 32, min_rows: int = 2048):
    """Split [0, n) into <= N_THREADS contiguous chunks whose starts are multiples of `align`."""
    t = max(1, min(N_THREADS, n // min_rows))
    step = ((n + t - 1) // t + align - 1) // align * align if n else 0
    return [(a, min(n, a + step)) for a in range(0, n, step)] if n else []


def _pmap(fn, items):
    """ctypes releases the GIL, so plain threads give real parallelism over the C loops."""
    items = list(items)
    if len(items) <= 1:
        return [fn(it) for it in items]
    with ThreadPoolExecutor(max_workers=N_THREADS) as ex:
        return list(ex.map(fn, items))


def pad_dim(dim: int) -> int:
    return (dim + DIM_ALIGN - 1) // DIM_ALIGN * DIM_ALIGN


def normalize_rows(x: np.ndarray, store_dtype: str = "fp32", dim_pad: int | None = None):
    """-> (stored [n, dim_pad] float32 or uint16(bf16 bits), number of non-finite rows)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x[None, :]
    n, dim = x.shape
    dp = dim_pad or pad_dim(dim)
    if store_dtype == "fp32":
        out, fn = np.empty((n, dp), np.float32), lib().orc_normalize_f32
    elif store_dtype == "bf16":
        out, fn = np.empty((n, dp), np.uint16), lib().orc_normalize_bf16
    else:
        raise ValueError(store_dtype)
    bad = _pmap(lambda ab: fn(_p(x[ab[0]:ab[1]]), ab[1] - ab[0], dim, dp, _p(out[ab[0]:ab[1]])), _row_chunks(n, 1))
    return out, int(sum(bad))


def stored_as_f32(stored: np.ndarray) -> np.ndarray:
    if stored.dtype == np.uint16:
        return (stored.astype(np.uint32) << 16).view(np.float32)
    return stored.astype(np.float32, copy=False)


def _dtype_code(stored: np.ndarray) -> int:
    if stored.dtype == np.float32:
        return 0
    if stored.dtype == np.uint16:
        return 1
    raise TypeError(stored.dtype)


def canonical_scores(qn: np.ndarray, stored: np.ndarray) -> np.ndarray:
    """[B, n] fp32 scores in the canonical accumulation order (bit-exact referee for the kernels)."""
    qn = np.ascontiguousarray(qn, np.float32)
    stored = np.ascontiguousarray(stored)
    B, dp = qn.shape
    n = stored.shape[0]
    assert stored.shape[1] == dp
    code = _dtype_code(stored)

    def part(ab):
        a, b = ab
        o = np.empty((B, b - a), np.float32)
        lib().orc_scores(_p(qn), B, _p(stored[a:b]), code, b - a, dp, _p(o))
        return o

    parts = _pmap(part, _row_chunks(n, 1))
    return np.concatenate(parts, axis=1) if parts else np.empty((B, 0), np.float32)


def mask_from_bool(allowed: np.ndarray) -> np.ndarray:
    """bool[n] -> packed uint32 words, bit r&31 of word r>>5 == row r allowed."""
    allowed = np.asarray(allowed, dtype=bool)
    n = allowed.shape[0]
    words = (n + 31) // 32
    padded = np.zeros(words * 32, dtype=np.uint8)
    padded[:n] = allowed
    return np.packbits(padded.reshape(words, 32), axis=1, bitorder="little").view(np.uint32).reshape(words).copy()


METRICS = {"cosine": 0, "l2": 1, "ip": 2}


def row_norms(x: np.ndarray, dim_pad: int | None = None) -> np.ndarray:
    """||x|| of raw rows as K1 measures it (canonical fp32 sum of squares, sqrtf); what the l2 / ip spaces rank with."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x[None, :]
    n, dim = x.shape
    out = np.empty(n, np.float32)
    lib().orc_norms_f32(_p(x), n, dim, dim_pad or pad_dim(dim), _p(out))
    return out


def distances_from_rank_scores(space: str, t: np.ndarray, q_norm: np.ndarray) -> np.ndarray:
    """Chroma's distance (lower = better) from the backend's ranking score t (orc_rank_score) and ||q||, float64:
    cosine 1 - t; ip 1 - ||q|| t = 1 - <q,x>; l2 ||q||^2 - 2 ||q|| t = ||q - x||^2."""
    t = np.asarray(t, np.float64)
    qn = np.asarray(q_norm, np.float64).reshape(-1, *([1] * (t.ndim - 1)))
    qc = np.maximum(qn, 1e-12)
    if space == "cosine":
        return 1.0 - t
    if space == "ip":
        return 1.0 - qc * t
    if space == "l2":
        return qn * qn - 2.0 * qc * t
    raise ValueError(space)


def brute_force_topk(qn, stored, k: int, mask: np.ndarray | None = None, id_base: int = 0, metric: str = "cosine",
                     norms: np.ndarray | None = None, q_norms: np.ndarray | None = None):
    """Exact canonical-fp32 top-k.  mask: packed uint32 words (mask_from_bool) or bool[n] or None.
    metric "l2" / "ip": ranking in that space over the RAW vectors: norms [n] = row_norms(raw rows), q_norms [B] =
    row_norms(raw queries), qn / stored the normalised forms as always; scores out are the ranking scores."""
    code_m = METRICS[metric]
    if code_m:
        norms = np.ascontiguousarray(norms, np.float32)
        q_norms = np.ascontiguousarray(q_norms, np.float32)
        assert norms.shape[0] == stored.shape[0] and q_norms.shape[0] == qn.shape[0]
    qn = np.ascontiguousarray(qn, np.float32)
    stored = np.ascontiguousarray(stored)
    B, dp = qn.shape
    n = stored.shape[0]
    if mask is not None and mask.dtype == bool:
        mask = mask_from_bool(mask)
    if mask is not None:
        mask = np.ascontiguousarray(mask, np.uint32)
        assert mask.shape[0] >= (n + 31) // 32
    code = _dtype_code(stored)

    def part(ab):
        a, b = ab
        s = np.empty((B, k), np.float32)
        i = np.empty((B, k), np.int64)
        lib().orc_topk_metric(_p(qn), B, _p(stored[a:b]), code, b - a, dp,
                              _p(mask[a // 32:]) if mask is not None else None, k, id_base + a, code_m,
                              _p(norms[a:b]) if code_m else None, _p(q_norms) if code_m else None, _p(s), _p(i))
        return s, i

    parts = _pmap(part, _row_chunks(n, 32))
    if not parts:
        return np.full((B, k), -np.inf, np.float32), np.full((B, k), -1, np.int64)
    if len(parts) == 1:
        return parts[0]
    return merge_topk(np.concatenate([p[0] for p in parts], 1), np.concatenate([p[1] for p in parts], 1), k)


def merge_topk(scores: np.ndarray, ids: np.ndarray, k: int):
    """[B, m] candidates (id -1 = empty) -> best k per row, score desc, tie -> lower id (cross-shard merge referee)."""
    B = scores.shape[0]
    out_s = np.full((B, k), -np.inf, scores.dtype)
    out_i = np.full((B, k), -1, np.int64)
    for b in range(B):
        ok = ids[b] >= 0
        s, i = scores[b][ok], ids[b][ok]
        s = np.where(s == 0, np.float32(0), s)
        order = np.lexsort((i, -s))[:k]
        out_s[b, :order.shape[0]] = s[order]
        out_i[b, :order.shape[0]] = i[order]
    return out_s, out_i


def f64_scores(qn: np.ndarray, stored: np.ndarray) -> np.ndarray:
    return np.asarray(qn, np.float64) @ stored_as_f32(stored).astype(np.float64).T


def _select(scores: np.ndarray, k: int, allowed: np.ndarray | None):
    """rows of `scores` -> (top scores, ids), score desc, tie -> lower id, (-inf, -1) fill."""
    B, n = scores.shape
    out_s = np.full((B, k), -np.inf, scores.dtype)
    out_i = np.full((B, k), -1, np.int64)
    idx_all = np.arange(n)
    for b in range(B):
        s = scores[b]
        idx = idx_all if allowed is None else idx_all[allowed]
        sv = s[idx]
        m = min(k, idx.shape[0])
        if m == 0:
            continue
        if idx.shape[0] > 4 * m:
            kth = np.partition(sv, idx.shape[0] - m)[idx.shape[0] - m]
            keep = sv >= kth  # keeps every tie of the k-th value
            idx, sv = idx[keep], sv[keep]
        order = np.lexsort((idx, -sv))[:m]
        out_s[b, :m] = sv[order]
        out_i[b, :m] = idx[order]
    return out_s, out_i


def referee_topk_f64(qn, stored, k: int, allowed: np.ndarray | None = None):
    """float64 referee computed from the STORED values (bf16 rows are upcast)."""
    return _select(f64_scores(qn, stored), k, None if allowed is None else np.asarray(allowed, bool))


def hnswlib_topk(space: str, rows_raw: np.ndarray, q_raw: np.ndarray, k: int, order: str = "simd16",
                 allowed: np.ndarray | None = None):
    """chroma-hnswlib's published arithmetic (oracle/hnswlib_restated.c): its distance functions for `space` ("l2" | "ip" |
    "cosine", the last with normalize_vector on rows and queries) over RAW fp32 vectors, in the scalar or the 16-lane SIMD
    summation order, and bruteforce.h's exhaustive top-k.  -> (distance [B, k] ascending, ids [B, k]); -1 / inf padded."""
    L = lib()
    rows = np.ascontiguousarray(rows_raw, np.float32)
    q = np.ascontiguousarray(q_raw, np.float32).reshape(-1, rows.shape[1])
    n, dim = rows.shape
    code = {"l2": 0, "ip": 1, "cosine": 2}[space]
    allow = None if allowed is None else mask_from_bool(np.asarray(allowed, bool))
    out_d = np.empty((q.shape[0], k), np.float32)
    out_i = np.empty((q.shape[0], k), np.int64)

    def one(b):
        d = np.empty(n, np.float32)
        L.hnsw_distances(code, 1 if order == "simd16" else 0, _p(rows), n, dim, _p(q[b]), _p(d))
        L.hnsw_bruteforce_knn(_p(d), n, None if allow is None else _p(allow), k, _p(out_d[b]), _p(out_i[b]))

    _pmap(one, range(q.shape[0]))
    return out_d, out_i


def sgemm_topk(qn: np.ndarray, rows_f32: np.ndarray, k: int, allowed: np.ndarray | None = None):
    """numpy sgemm + partition: the arithmetic of chromadb's brute-force path; the timed CPU baseline."""
    scores = np.asarray(qn, np.float32) @ rows_f32.T
    return _select(scores, k, None if allowed is None else np.asarray(allowed, bool))


# ---------------------------------------------------------------------------------- synthetic inputs
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64_np(seed: int, idx: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (idx.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def synth_rows(seed: int, first_row: int, n: int, dim: int, use_numpy: bool = False) -> np.ndarray:
    """Counter-based rows v[r, c] = uniform[-1, 1) from splitmix64(seed, r*dim + c)."""
    if use_numpy:
        idx = (np.arange(first_row, first_row + n, dtype=np.uint64)[:, None] * np.uint64(dim)
               + np.arange(dim, dtype=np.uint64)[None, :])
        z = _splitmix64_np(seed, idx)
        return ((z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 8388608.0) - np.float32(1.0))
    out = np.empty((n, dim), np.float32)
    lib().orc_synth_rows(ctypes.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), first_row, n, dim, _p(out))
    return out


def planted_queries(query_seed: int, B: int, corpus_seed: int, n_rows: int, dim: int, noise: float = 0.05):
    """B queries; even-numbered ones are planted: row t (known) + noise * hash-noise, so top-1 is known.

    -> (queries [B, dim] fp32 un-normalised, planted_row [B] int64 with -1 for free queries)."""
    q = synth_rows(query_seed, 0, B, dim)
    planted = np.full(B, -1, np.int64)
    if n_rows > 0:
        pick = _splitmix64_np(query_seed ^ 0xA5A5A5A5, np.arange(B, dtype=np.uint64)) % np.uint64(n_rows)
        for b in range(0, B, 2):
            t = int(pick[b])
            row = synth_rows(corpus_seed, t, 1, dim)[0]
            row = row / np.linalg.norm(row)
            nz = q[b] / np.linalg.norm(q[b])
            q[b] = row + np.float32(noise) * nz
            planted[b] = t
    return q, planted


# ------------------------------------------------------------------------------- where-filter referee
def _cmp(op: str, have: Any, want: Any) -> bool:
    if op == "$eq":
        return have == want
    if op == "$ne":
        return have != want
    if op in ("$in", "$nin"):
        hit = have in list(want)
        return hit if op == "$in" else not hit
    if have is None or isinstance(have, str) != isinstance(want, str):
        return False
    if op == "$gt":
        return have > want
    if op == "$gte":
        return have >= want
    if op == "$lt":
        return have < want
    if op == "$lte":
        return have <= want
    raise ValueError(f"unknown operator {op}")


def _match(where: Mapping[str, Any], md: Mapping[str, Any]) -> bool:
    if len(where) != 1:
        raise ValueError("where expects exactly one key per level")
    (key, val), = where.items()
    if key == "$and":
        return all(_match(w, md) for w in val)
    if key == "$or":
        return any(_match(w, md) for w in val)
    if isinstance(val, Mapping):
        if len(val) != 1:
            raise ValueError("operator expression expects exactly one operator")
        (op, want), = val.items()
    else:
        op, want = "$eq", val
    if key not in md:
        # chroma: a missing field matches only the negative operators
        return op in ("$ne", "$nin")
    return _cmp(op, md[key], want)


def eval_where(where: Mapping[str, Any] | None, metadatas: Sequence[Mapping[str, Any]] | Iterable) -> np.ndarray:
    """Row-by-row evaluation of a Chroma `where` dict -> bool[n]."""
    mds = list(metadatas)
    if not where:
        return np.ones(len(mds), bool)
    return np.array([_match(where, md or {}) for md in mds], dtype=bool)
