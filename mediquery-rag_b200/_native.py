"""ctypes binding of libmqvs.so (include/mqvs.h).  There is NO fallback: if the library is missing or no
B200 is visible, the calls raise -- the product never computes on the CPU."""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmqvs.so")

OK, E_INVALID, E_CUDA, E_NCCL, E_OOM, E_IO = 0, -1, -2, -3, -4, -5
F32, BF16 = 0, 1
DIM_PAD = 768
MAX_K = 1024
MAX_DEVICES = 16
SHARD_BLOCK = 4096  # rows per block of a multi-device handle's block-cyclic placement (include/mqvs.h)
DTYPES = {"fp32": F32, "float32": F32, "f32": F32, "bf16": BF16, "bfloat16": BF16}
METRICS = {"cosine": 0, "l2": 1, "ip": 2}  # Chroma's hnsw:space -> MQVS_METRIC_*


class FilterOp(ctypes.Structure):
    """mqvs_filter_op (include/mqvs.h): one operation of a postfix `where` program."""
    _fields_ = [("op", ctypes.c_int32), ("col", ctypes.c_int32), ("ival", ctypes.c_int64), ("fval", ctypes.c_double),
                ("set_off", ctypes.c_uint32), ("set_len", ctypes.c_uint32)]


OP = {"$eq": 0, "$ne": 1, "$gt": 2, "$gte": 3, "$lt": 4, "$lte": 5, "$in": 6, "$nin": 7, "$contains": 8, "$not_contains": 9,
      "$text_contains": 10, "$text_not_contains": 11,  # where_document leaves over a text column
      "$and": 16, "$or": 17}
COL_I64, COL_F64, COL_TOKENS, COL_TEXT = 0, 1, 2, 3
MAX_NEEDLE = 1024  # MQVS_MAX_NEEDLE


class DeviceMask:
    """A packed allow-bitmask in device memory owned by the index (valid until its next filter_build)."""

    def __init__(self, ptr: int, words: int):
        self._ptr, self.words = int(ptr), int(words)

    def data_ptr(self) -> int:
        return self._ptr


class MqvsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"mqvs error {code}: {msg}")
        self.code = code


class Stats(ctypes.Structure):
    _fields_ = [("rows", ctypes.c_uint64), ("searches", ctypes.c_uint64), ("queries", ctypes.c_uint64),
                ("kernel_launches", ctypes.c_uint64), ("scan_launches", ctypes.c_uint64),
                ("scan_bytes", ctypes.c_uint64), ("scan_ms", ctypes.c_double), ("last_search_ms", ctypes.c_double),
                ("last_scan_ms", ctypes.c_double), ("last_scan_passes", ctypes.c_uint32),
                ("last_path", ctypes.c_uint32), ("exact_fallbacks", ctypes.c_uint64), ("unrepaired", ctypes.c_uint64),
                ("last_pipeline", ctypes.c_uint32), ("reserved_", ctypes.c_uint32), ("band_repairs", ctypes.c_uint64)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


_libs: dict = {}  # row width -> loaded library

# name -> (restype, argtypes); the loader test checks every symbol include/mqvs.h declares is here and exported
_vp, _i, _u64, _i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_int64
SIGNATURES = {
    "mqvs_version": (_i, []),
    "mqvs_dim_pad": (_i, []),
    "mqvs_last_error": (ctypes.c_char_p, []),
    "mqvs_device_count": (_i, [ctypes.POINTER(_i)]),
    "mqvs_create": (_i, [_i, ctypes.POINTER(_i), _i, _i, _u64, ctypes.POINTER(_vp)]),
    "mqvs_destroy": (None, [_vp]),
    "mqvs_count": (_i, [_vp, ctypes.POINTER(_u64)]),
    "mqvs_reserve": (_i, [_vp, _u64]),
    "mqvs_add": (_i, [_vp, _vp, _i, _i, _u64, ctypes.POINTER(_u64)]),
    "mqvs_add_synthetic": (_i, [_vp, _u64, _u64, _u64, ctypes.POINTER(_u64)]),
    "mqvs_search": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _i]),
    "mqvs_search_async": (_i, [_vp, _vp, _i, _i, _vp, _i64, _vp, _vp, _vp]),
    "mqvs_merge_candidates": (_i, [_i, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "mqvs_get_rows": (_i, [_vp, _vp, _u64, _vp]),
    "mqvs_save": (_i, [_vp, ctypes.c_char_p]),
    "mqvs_save_append": (_i, [_vp, ctypes.c_char_p, _u64]),
    "mqvs_load": (_i, [_vp, ctypes.c_char_p]),
    "mqvs_compact": (_i, [_vp, _vp, ctypes.POINTER(_u64)]),
    "mqvs_get_stats": (_i, [_vp, ctypes.POINTER(Stats)]),
    "mqvs_set_path": (_i, [_vp, _i]),
    "mqvs_set_shadow": (_i, [_vp, _i]),
    "mqvs_shadow_bytes": (_i, [_vp, ctypes.POINTER(_u64)]),
    "mqvs_set_metric": (_i, [_vp, _i]),
    "mqvs_norm_deviation": (_i, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "mqvs_get_norms": (_i, [_vp, _u64, _u64, _vp]),
    "mqvs_column_set": (_i, [_vp, _i, _i, _vp, _vp, _u64]),
    "mqvs_column_set_tokens": (_i, [_vp, _i, _vp, _vp, _u64]),
    "mqvs_column_set_text": (_i, [_vp, _i, _vp, _vp, _vp, _u64]),
    "mqvs_filter_build": (_i, [_vp, _vp, _i, _vp, _u64, _vp, ctypes.POINTER(_vp), ctypes.POINTER(_u64)]),
    "mqvs_shard_of_row": (_i, [_i, _u64, ctypes.POINTER(_i), ctypes.POINTER(_u64)]),
    "mqvs_xchg_create": (_i, [_i, _i, _i, _u64, ctypes.POINTER(_vp)]),
    "mqvs_xchg_get_handle": (_i, [_vp, _vp, _i]),
    "mqvs_xchg_connect": (_i, [_vp, _vp]),
    "mqvs_xchg_base": (_i, [_vp, ctypes.POINTER(_vp)]),
    "mqvs_xchg_connect_ptrs": (_i, [_vp, ctypes.POINTER(_vp)]),
    "mqvs_xchg_max_entries": (_i, [_vp, ctypes.POINTER(_u64)]),
    "mqvs_xchg_merge": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp]),
    "mqvs_xchg_merge_colocated": (_i, [ctypes.POINTER(_vp), _i, ctypes.POINTER(_vp), ctypes.POINTER(_vp), _i, _i,
                                       ctypes.POINTER(_vp), ctypes.POINTER(_vp), _vp]),
    "mqvs_xchg_destroy": (None, [_vp]),
}


def lib_path(width: int = DIM_PAD) -> str:
    """Row width is a build-time constant of the kernels (include/mqvs.h, MQVS_DIM_PAD): libmqvs.so is the 768 build, the
    others are libmqvs_w<width>.so (csrc/Makefile, `make WIDTH=...`)."""
    return LIB_PATH if width == DIM_PAD else os.path.join(_HERE, f"libmqvs_w{int(width)}.so")


def built_widths() -> list:
    """Row widths a library has been built for, ascending."""
    out = [DIM_PAD] if os.path.exists(LIB_PATH) else []
    for name in os.listdir(_HERE):
        if name.startswith("libmqvs_w") and name.endswith(".so") and name[9:-3].isdigit():
            out.append(int(name[9:-3]))
    return sorted(set(out))


def width_for(dim: int) -> int:
    """The narrowest built row width that holds `dim` elements (rows are zero-padded to it in HBM)."""
    have = built_widths()
    for w in have:
        if w >= dim:
            return w
    raise ValueError(f"no libmqvs build holds {dim}-dimensional rows (built: {have or 'none'}); build one with "
                     f"`make -C mediquery-rag_b200/csrc WIDTH={(int(dim) + 127) // 128 * 128}`")


def load(width: int = DIM_PAD) -> ctypes.CDLL:
    lib = _libs.get(width)
    if lib is None:
        path = lib_path(width)
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: build it with `python __graft_entry__.py` "
                              "(nvcc, sm_100a); there is no CPU fallback")
        lib = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.mqvs_dim_pad() != width:
            raise ImportError(f"{path} was built for rows of {lib.mqvs_dim_pad()} elements, not {width}")
        _libs[width] = lib
    return lib


def check(rc: int, lib: Optional[ctypes.CDLL] = None) -> None:
    if rc != OK:
        raise MqvsError(rc, (lib or load()).mqvs_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    n = _i(0)
    check(load().mqvs_device_count(ctypes.byref(n)))
    return n.value


def _ptr(a) -> ctypes.c_void_p:
    """numpy array -> host pointer; torch tensor -> data_ptr; None -> NULL."""
    if a is None:
        return ctypes.c_void_p(None)
    if isinstance(a, np.ndarray):
        return ctypes.c_void_p(a.ctypes.data)
    return ctypes.c_void_p(a.data_ptr())


def _is_torch(a) -> bool:
    return a is not None and not isinstance(a, np.ndarray) and hasattr(a, "data_ptr")


def shard_of_row(n_dev: int, row: int):
    """(shard, local row) of a global row on a handle of n_dev shards -- libmqvs's own host arithmetic (no GPU needed)."""
    shard, local = _i(0), _u64(0)
    check(load().mqvs_shard_of_row(int(n_dev), int(row), ctypes.byref(shard), ctypes.byref(local)))
    return shard.value, local.value


class Index:
    """Thin owner of an mqvs_t handle: one row shard on one B200 (`device` = an int), or ONE handle over several shards
    in this process (`device` = a sequence of device ids, e.g. range(8); a device may repeat -- several shards on one GPU).
    Row ids are global in both cases; device tensors handed to a multi-device index live on its first device."""

    def __init__(self, dim: int, store_dtype: str = "fp32", device=0, capacity_rows: int = 0, shadow: Optional[bool] = None):
        """shadow: fp32 stores keep a bf16 scoring copy of their rows for the tensor pipelines (include/mqvs.h,
        mqvs_set_shadow; +50 % HBM, same results); None = the library default (on), False = score the fp32 rows."""
        self._h = _vp(None)
        self.dim = int(dim)
        self.store_dtype = store_dtype
        self.devices = [int(device)] if isinstance(device, (int, np.integer)) else [int(d) for d in device]
        if not 1 <= len(self.devices) <= MAX_DEVICES:
            raise ValueError(f"an index drives 1..{MAX_DEVICES} devices, got {len(self.devices)}")
        self.device = self.devices[0]
        self.metric = "cosine"
        self.width = width_for(self.dim)  # rows are zero-padded to this many elements (the build that serves them)
        self._lib = load(self.width)
        dev = (_i * len(self.devices))(*self.devices)
        self._check(self._lib.mqvs_create(len(self.devices), dev, self.dim, DTYPES[store_dtype], int(capacity_rows),
                                 ctypes.byref(self._h)))
        if shadow is not None:
            self.set_shadow(shadow)

    def _check(self, rc: int) -> None:
        check(rc, self._lib)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.mqvs_destroy(self._h)
            self._h = _vp(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        n = _u64(0)
        self._check(self._lib.mqvs_count(self._h, ctypes.byref(n)))
        return n.value

    def reserve(self, rows: int) -> None:
        self._check(self._lib.mqvs_reserve(self._h, int(rows)))

    def add(self, x) -> int:
        """x: [n, dim] fp32 numpy array (host) or CUDA torch tensor on this index's device.  -> first row id."""
        first = _u64(0)
        if _is_torch(x):
            if not x.is_cuda or x.dtype.itemsize != 4 or not x.is_contiguous() or x.shape[-1] != self.dim:
                raise ValueError("device embeddings must be a contiguous fp32 CUDA tensor [n, dim]")
            n, dev = (x.shape[0] if x.dim() == 2 else 1), 1
        else:
            x = np.ascontiguousarray(x, dtype=np.float32)
            if x.ndim == 1:
                x = x[None, :]
            if x.ndim != 2 or x.shape[1] != self.dim:
                raise ValueError(f"embeddings must be [n, {self.dim}], got {x.shape}")
            n, dev = x.shape[0], 0
        self._check(self._lib.mqvs_add(self._h, _ptr(x), F32, dev, n, ctypes.byref(first)))
        return first.value

    def add_synthetic(self, seed: int, first_index: int, n: int) -> int:
        first = _u64(0)
        self._check(self._lib.mqvs_add_synthetic(self._h, seed & 0xFFFFFFFFFFFFFFFF, int(first_index), int(n), ctypes.byref(first)))
        return first.value

    def search(self, q, k: int, mask=None):
        """q [B, dim] fp32 host array -> (scores [B, k] float32, ids [B, k] int64) host arrays."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [B, {self.dim}], got {q.shape}")
        B = q.shape[0]
        scores = np.empty((B, k), np.float32)
        ids = np.empty((B, k), np.int64)
        mask_dev = 0
        if mask is not None:
            if _is_torch(mask) or isinstance(mask, DeviceMask):
                mask_dev = 1
            else:
                mask = np.ascontiguousarray(mask, dtype=np.uint32)
                if mask.shape[0] < (len(self) + 31) // 32:
                    raise ValueError("mask has fewer than ceil(rows/32) words")
        self._check(self._lib.mqvs_search(self._h, _ptr(q), 0, B, int(k), _ptr(mask), mask_dev, _ptr(scores), _ptr(ids), 0))
        return scores, ids

    def search_device(self, q, k: int, out_scores, out_ids, mask=None, id_base: int = 0, stream: int = 0) -> None:
        """All-device variant on a caller stream (torch tensors; plumbing only).  The C ABI takes raw pointers, so the
        layout is checked here: q [B, dim] float32, out_scores [B, k] float32, out_ids [B, k] int64, all contiguous."""
        B = q.shape[0]
        for name, t, want, shape in (("q", q, "float32", (B, self.dim)), ("out_scores", out_scores, "float32", (B, int(k))),
                                     ("out_ids", out_ids, "int64", (B, int(k)))):
            if not str(t.dtype).endswith(want) or tuple(t.shape) != shape or not t.is_contiguous():
                raise TypeError(f"{name} must be a contiguous {want} tensor of shape {shape}, got {t.dtype} {tuple(t.shape)}")
        self._check(self._lib.mqvs_search_async(self._h, _ptr(q), B, int(k), _ptr(mask), int(id_base), _ptr(out_scores),
                                       _ptr(out_ids), _vp(stream)))

    def get_rows(self, row_ids) -> np.ndarray:
        row_ids = np.ascontiguousarray(row_ids, dtype=np.int64)
        out = np.empty((row_ids.shape[0], self.dim), np.float32)
        self._check(self._lib.mqvs_get_rows(self._h, _ptr(row_ids), row_ids.shape[0], _ptr(out)))
        return out

    def save(self, path: str) -> None:
        self._check(self._lib.mqvs_save(self._h, os.fsencode(path)))

    def save_append(self, path: str, first_row: int) -> None:
        """Append rows [first_row, len) to a row file that holds exactly first_row rows (0 = rewrite everything)."""
        self._check(self._lib.mqvs_save_append(self._h, os.fsencode(path), int(first_row)))

    def load_file(self, path: str) -> None:
        self._check(self._lib.mqvs_load(self._h, os.fsencode(path)))

    def compact(self, keep_words: np.ndarray) -> int:
        """Keep only the rows whose bit is set in the packed bitmap (bit r & 31 of word r >> 5), bit-for-bit, renumbered
        in order.  -> new row count.  Device-resident metadata columns are dropped (upload them again)."""
        words = np.ascontiguousarray(keep_words, dtype=np.uint32)
        if words.shape[0] < (len(self) + 31) // 32:
            raise ValueError("bitmap has fewer than ceil(rows/32) words")
        n = _u64(0)
        self._check(self._lib.mqvs_compact(self._h, _ptr(words), ctypes.byref(n)))
        return n.value

    def column_set(self, col_id: int, values: np.ndarray, present: Optional[np.ndarray] = None) -> None:
        """Upload metadata column `col_id` (int64 or float64 values, optional presence flags) for filter_build."""
        values = np.ascontiguousarray(values)
        if values.dtype == np.int64:
            ctype = COL_I64
        elif values.dtype == np.float64:
            ctype = COL_F64
        else:
            raise ValueError("device columns are int64 or float64")
        pres = None if present is None else np.ascontiguousarray(present, dtype=np.uint8)
        self._check(self._lib.mqvs_column_set(self._h, int(col_id), ctype, _ptr(values), _ptr(pres), values.shape[0]))

    def column_set_tokens(self, col_id: int, offsets: np.ndarray, tokens: np.ndarray) -> None:
        """Upload a token-set column (comma-joined `tags`): CSR offsets [n + 1] (uint64) and token ids (int64)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        tokens = np.ascontiguousarray(tokens, dtype=np.int64)
        if offsets.ndim != 1 or offsets.shape[0] < 1 or int(offsets[-1]) != tokens.shape[0]:
            raise ValueError("offsets must be [n + 1] and end at len(tokens)")
        self._check(self._lib.mqvs_column_set_tokens(self._h, int(col_id), _ptr(offsets), _ptr(tokens), offsets.shape[0] - 1))

    def column_set_text(self, col_id: int, offsets: np.ndarray, data: np.ndarray, present: Optional[np.ndarray] = None) -> None:
        """Upload a text column (the stored documents, for `where_document`): CSR offsets [n + 1] (uint64) into the rows'
        concatenated UTF-8 bytes (uint8); present: optional flags (0 = the row has no document)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        data = np.ascontiguousarray(data, dtype=np.uint8)
        if offsets.ndim != 1 or offsets.shape[0] < 1 or int(offsets[-1]) != data.shape[0]:
            raise ValueError("offsets must be [n + 1] and end at len(data)")
        pres = None if present is None else np.ascontiguousarray(present, dtype=np.uint8)
        if pres is not None and pres.shape[0] != offsets.shape[0] - 1:
            raise ValueError("one presence flag per row")
        self._check(self._lib.mqvs_column_set_text(self._h, int(col_id), _ptr(offsets), _ptr(data), _ptr(pres), offsets.shape[0] - 1))

    def filter_build(self, ops, pool: Optional[np.ndarray] = None, live_words: Optional[np.ndarray] = None):
        """ops: list of FilterOp in postfix order -> (DeviceMask, number of allowed rows)."""
        arr = (FilterOp * max(len(ops), 1))(*ops)
        pool = np.zeros(0, np.int64) if pool is None else np.ascontiguousarray(pool)
        if pool.dtype.itemsize != 8:
            raise ValueError("the constant pool holds 8-byte values")
        live = None if live_words is None else np.ascontiguousarray(live_words, dtype=np.uint32)
        ptr, allowed = _vp(), _u64(0)
        self._check(self._lib.mqvs_filter_build(self._h, arr, len(ops), _ptr(pool), pool.shape[0], _ptr(live), ctypes.byref(ptr),
                                       ctypes.byref(allowed)))
        return DeviceMask(ptr.value or 0, (len(self) + 31) // 32), int(allowed.value)

    def stats(self) -> dict:
        st = Stats()
        self._check(self._lib.mqvs_get_stats(self._h, ctypes.byref(st)))
        return st.as_dict()

    def set_path(self, path: int) -> None:
        self._check(self._lib.mqvs_set_path(self._h, int(path)))

    def set_shadow(self, enable: bool) -> None:
        """fp32 stores: stream a bf16 scoring shadow on the tensor pipelines (default) or the fp32 rows themselves."""
        self._check(self._lib.mqvs_set_shadow(self._h, 1 if enable else 0))

    def shadow_bytes(self) -> int:
        n = _u64(0)
        self._check(self._lib.mqvs_shadow_bytes(self._h, ctypes.byref(n)))
        return n.value

    def set_metric(self, space: str) -> None:
        """Space the rows are ranked in ("cosine" | "l2" | "ip"); a non-cosine space must first be chosen while the
        index is empty (row norms are kept from then on).  Scores of later searches are that space's ranking score."""
        self._check(self._lib.mqvs_set_metric(self._h, METRICS[space]))
        self.metric = space

    def norm_deviation(self) -> float:
        """Largest | ||x|| - 1 | over the raw rows added so far (0: every row was a unit vector)."""
        d = ctypes.c_float(0.0)
        self._check(self._lib.mqvs_norm_deviation(self._h, ctypes.byref(d)))
        return float(d.value)

    def get_norms(self, first_row: int = 0, n: Optional[int] = None) -> np.ndarray:
        n = len(self) - int(first_row) if n is None else int(n)
        out = np.empty(max(n, 0), np.float32)
        self._check(self._lib.mqvs_get_norms(self._h, int(first_row), n, _ptr(out)))
        return out


def merge_candidates(device: int, scores, ids, out_scores, out_ids, stream: int = 0) -> None:
    """scores/ids [W, B, k] device tensors -> out [B, k] device tensors."""
    W, B, k = scores.shape
    check(load().mqvs_merge_candidates(int(device), _ptr(scores), _ptr(ids), W, B, k, _ptr(out_scores), _ptr(out_ids),
                                       _vp(stream)))


class PeerExchange:
    """Cross-shard candidate exchange + merge in one kernel over NVLink peer memory (include/mqvs.h, mqvs_xchg_*).
    Create on every rank, publish `.handle` (64 bytes) to all ranks over any host channel, then `.connect(handles)`
    with the handles in rank order."""

    HANDLE_BYTES = 64

    def __init__(self, device: int, rank: int, world: int, max_entries: int = 1 << 18):
        h = _vp()
        check(load().mqvs_xchg_create(int(device), int(rank), int(world), int(max_entries), ctypes.byref(h)))
        self._h, self.world = h, int(world)
        try:
            buf = ctypes.create_string_buffer(self.HANDLE_BYTES)
            check(load().mqvs_xchg_get_handle(self._h, buf, self.HANDLE_BYTES))
            self.handle = bytes(buf.raw)
            cap = _u64(0)
            check(load().mqvs_xchg_max_entries(self._h, ctypes.byref(cap)))
            self.max_entries = cap.value
        except Exception:
            self.close()
            raise

    def connect(self, handles) -> None:
        handles = [bytes(x) for x in handles]
        if len(handles) != self.world or any(len(x) != self.HANDLE_BYTES for x in handles):
            raise ValueError("expected one 64-byte handle per rank")
        blob = ctypes.create_string_buffer(b"".join(handles), self.HANDLE_BYTES * self.world)
        check(load().mqvs_xchg_connect(self._h, blob))

    @property
    def base(self) -> int:
        p = _vp()
        check(load().mqvs_xchg_base(self._h, ctypes.byref(p)))
        return p.value

    def connect_ptrs(self, bases) -> None:
        """Same-process ranks: the W inbox base pointers (`.base` of every rank's exchange) in rank order."""
        arr = (_vp * self.world)(*[_vp(int(b)) for b in bases])
        check(load().mqvs_xchg_connect_ptrs(self._h, arr))

    def merge(self, scores, ids, out_scores, out_ids, stream: int = 0) -> None:
        """scores/ids [B, k] device tensors of this rank -> out [B, k]: best k of all ranks' candidates."""
        B, k = scores.shape
        check(load().mqvs_xchg_merge(self._h, _ptr(scores), _ptr(ids), B, k, _ptr(out_scores), _ptr(out_ids), _vp(stream)))

    @staticmethod
    def merge_colocated(ranks, scores, ids, out_scores, out_ids, stream: int = 0) -> None:
        """All ranks of one exchange that share ONE device, in a single cooperative launch (include/mqvs.h): lists of W
        exchanges / [B, k] tensors in rank order."""
        W = len(ranks)
        B, k = scores[0].shape
        arr = lambda xs: (_vp * W)(*[_ptr(t) for t in xs])  # noqa: E731
        hs = (_vp * W)(*[r._h for r in ranks])
        check(load().mqvs_xchg_merge_colocated(hs, W, arr(scores), arr(ids), B, k, arr(out_scores), arr(out_ids), _vp(stream)))

    def close(self) -> None:
        if getattr(self, "_h", None):
            load().mqvs_xchg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
