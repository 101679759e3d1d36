"""Chroma `where` filters compiled to the packed allow-bitmask the scan kernels consume.

The reference passes no filter today (/root/reference/src/agents/nodes.py:93 calls similarity_search(q, k=5)),
but the Chroma surface it programs against takes `filter=` and BASELINE.json config 5 exercises it.  Grammar:
{"field": v} (== $eq) | {"field": {"$eq|$ne|$gt|$gte|$lt|$lte": v}} | {"field": {"$in|$nin": [..]}} |
{"field": {"$contains|$not_contains": token}} | {"$and": [..]} | {"$or": [..]}; `where_document` ($contains /
$not_contains over the stored documents, $and / $or) is the same kind of program over a text column (DeviceFilter).
$contains / $not_contains on a METADATA field is token-set membership: the reference stores `tags` as one comma-joined
string per chunk (/root/reference/src/ingest_medical.py:66-67,75-79, e.g. "慢性病，代谢性疾病，发病率"), so "chunks tagged X" is
`{"tags": {"$contains": "X"}}`: the value is split on "," and "，", tokens are stripped, X must equal one of them
(a row without the field, or whose value is not a string, matches only $not_contains).  Evaluation is columnar (one numpy pass per leaf); a field missing from a row
matches only $ne / $nin.  The result is AND-ed with the live-row bitmap and packed little-endian into uint32
words: bit (r & 31) of word (r >> 5) is row r.
"""
from __future__ import annotations

import re
from typing import Any, Dict, Iterable, List, Mapping, Optional, Sequence

import numpy as np

_COMPARE = ("$gt", "$gte", "$lt", "$lte")
_TOKEN_OPS = ("$contains", "$not_contains")
_LEAF_OPS = ("$eq", "$ne", "$in", "$nin") + _COMPARE + _TOKEN_OPS
_TOKEN_SPLIT = re.compile("[,\uff0c]")


def tokens_of(value: Any) -> List[str]:
    """Tokens of a comma-joined metadata value ("," or "，" separated, stripped, empties dropped); [] for non-strings."""
    if not isinstance(value, str):
        return []
    return [t for t in (p.strip() for p in _TOKEN_SPLIT.split(value)) if t]


class FilterError(ValueError):
    pass


class MetadataColumns:
    """Per-field columns over row-aligned metadata dicts, built lazily and extended incrementally."""

    def __init__(self) -> None:
        self._n = 0                                   # rows
        self._rows: Dict[int, Mapping[str, Any]] = {} # row -> its metadata dict; rows without per-row metadata are absent
        self._explicit: Dict[str, np.ndarray] = {}   # bulk columns supplied as arrays (no per-row dicts)
        self._cache: Dict[str, tuple] = {}           # field -> (n_rows_built, values object array, present bool array)
        self._dict_fields: set = set()               # fields that occur in per-row dicts
        self.version = 0                             # bumped by every mutation (device copies of columns key on it)

    def __len__(self) -> int:
        return self._n

    def extend(self, metadatas: Sequence[Optional[Mapping[str, Any]]]) -> None:
        for md in metadatas:
            if md:
                self._dict_fields.update(md.keys())
                self._rows[self._n] = md
            self._n += 1
        self.version += 1

    def extend_empty(self, n: int) -> None:
        """Append n rows without metadata (bulk loads of millions of rows cost nothing per row)."""
        self._n += int(n)
        self.version += 1

    def extend_columns(self, n: int, columns: Mapping[str, np.ndarray]) -> None:
        """Append n rows whose metadata is given column-wise (bulk ingest of millions of rows)."""
        start = self._n
        self._n += n
        for name, col in columns.items():
            col = np.asarray(col)
            if col.shape[0] != n:
                raise ValueError(f"column {name!r} has {col.shape[0]} values for {n} rows")
            old = self._explicit.get(name)
            if old is None:
                old, present = np.zeros(start, dtype=col.dtype), np.zeros(start, dtype=bool)
            else:
                old, present = old
                if old.shape[0] < start:  # rows added since without this column
                    pad = start - old.shape[0]
                    old = np.concatenate([old, np.zeros(pad, old.dtype)])
                    present = np.concatenate([present, np.zeros(pad, bool)])
            self._explicit[name] = (np.concatenate([old, col]), np.concatenate([present, np.ones(n, bool)]))
        for name, (vals, present) in list(self._explicit.items()):
            if vals.shape[0] < start + n:
                pad = start + n - vals.shape[0]
                self._explicit[name] = (np.concatenate([vals, np.zeros(pad, vals.dtype)]),
                                        np.concatenate([present, np.zeros(pad, bool)]))
        self._cache.clear()
        self.version += 1

    def row(self, r: int) -> Dict[str, Any]:
        md = dict(self._rows.get(r, _EMPTY))
        for name, (vals, present) in self._explicit.items():
            if r < present.shape[0] and present[r]:
                md[name] = vals[r].item()
        return md

    def set_row(self, r: int, md: Optional[Mapping[str, Any]]) -> None:
        if md:
            self._rows[r] = md
            self._dict_fields.update(md.keys())
        else:
            self._rows.pop(r, None)
        self._cache.clear()
        self.version += 1

    def column(self, field: str):
        """-> (values, present): values is a typed array when the field is homogeneous, else object."""
        n = self._n
        if field in self._explicit:
            vals, present = self._explicit[field]
            if field not in self._dict_fields and vals.shape[0] == n:
                return vals, present
        hit = self._cache.get(field)
        if hit is not None and hit[0] == n:
            return hit[1], hit[2]
        vals = np.empty(n, dtype=object)
        present = np.zeros(n, dtype=bool)
        if field in self._explicit:
            ev, ep = self._explicit[field]
            vals[: ev.shape[0]] = ev
            present[: ep.shape[0]] = ep
        for r, md in self._rows.items():
            if field in md:
                vals[r] = md[field]
                present[r] = True
        self._cache[field] = (n, vals, present)
        return vals, present


_EMPTY: Mapping[str, Any] = {}


def _leaf(cols: MetadataColumns, field: str, op: str, want: Any) -> np.ndarray:
    vals, present = cols.column(field)
    n = len(cols)
    if op in _TOKEN_OPS:
        if not isinstance(want, str):
            raise FilterError(f"{op} expects a string token")
        want = want.strip()
        hit = np.fromiter((bool(p) and want in tokens_of(v) for v, p in zip(vals, present)), dtype=bool, count=n)
        return hit if op == "$contains" else ~hit
    if op in ("$in", "$nin"):
        if not isinstance(want, (list, tuple, set, frozenset, np.ndarray)):
            raise FilterError(f"{op} expects a list, got {type(want).__name__}")
        wanted = list(want)
        if vals.dtype != object:
            hit = np.isin(vals, np.asarray(wanted)) if wanted else np.zeros(n, bool)
        else:
            ws = set(wanted)
            hit = np.fromiter((p and v in ws for v, p in zip(vals, present)), dtype=bool, count=n)
        hit &= present
        return hit if op == "$in" else ~hit
    if isinstance(want, (list, dict, tuple, set)):
        raise FilterError(f"{op} expects a scalar, got {type(want).__name__}")
    if op in ("$eq", "$ne"):
        if vals.dtype != object:
            hit = (vals == want) if not isinstance(want, str) or vals.dtype.kind in "US" else np.zeros(n, bool)
        else:
            hit = np.fromiter((bool(p and v == want) for v, p in zip(vals, present)), dtype=bool, count=n)
        hit = np.asarray(hit, bool) & present
        return hit if op == "$eq" else ~hit
    # ordered comparisons: only between like kinds (str with str, number with number); missing never matches
    want_is_str = isinstance(want, str)

    def ok(v: Any) -> bool:
        return isinstance(v, str) == want_is_str and v is not None

    if vals.dtype != object and (vals.dtype.kind in "US") == want_is_str:
        v = vals
        hit = {"$gt": v > want, "$gte": v >= want, "$lt": v < want, "$lte": v <= want}[op]
        return np.asarray(hit, bool) & present
    if vals.dtype != object:
        return np.zeros(n, bool)
    import operator as _op

    fn = {"$gt": _op.gt, "$gte": _op.ge, "$lt": _op.lt, "$lte": _op.le}[op]
    return np.fromiter((bool(p and ok(v) and fn(v, want)) for v, p in zip(vals, present)), dtype=bool, count=n)


def evaluate(where: Optional[Mapping[str, Any]], cols: MetadataColumns) -> np.ndarray:
    """Chroma `where` dict -> bool[n_rows]."""
    n = len(cols)
    if not where:
        return np.ones(n, dtype=bool)
    if not isinstance(where, Mapping) or len(where) != 1:
        raise FilterError("where expects a dict with exactly one key per level (use $and / $or to combine)")
    (key, val), = where.items()
    if key in ("$and", "$or"):
        if not isinstance(val, (list, tuple)) or not val:
            raise FilterError(f"{key} expects a non-empty list of filters")
        parts = [evaluate(w, cols) for w in val]
        out = parts[0].copy()
        for p in parts[1:]:
            out = (out & p) if key == "$and" else (out | p)
        return out
    if key.startswith("$"):
        raise FilterError(f"unknown logical operator {key}")
    if isinstance(val, Mapping):
        if len(val) != 1:
            raise FilterError("an operator expression takes exactly one operator")
        (op, want), = val.items()
        if op not in _LEAF_OPS:
            raise FilterError(f"unknown operator {op}")
    else:
        op, want = "$eq", val
    return _leaf(cols, key, op, want)


def evaluate_document(where_document: Optional[Mapping[str, Any]], texts: Sequence[Optional[str]]) -> np.ndarray:
    """Chroma `where_document`: {"$contains": s} | {"$not_contains": s} | {"$and": [..]} | {"$or": [..]} -> bool[n]."""
    n = len(texts)
    if not where_document:
        return np.ones(n, dtype=bool)
    if not isinstance(where_document, Mapping) or len(where_document) != 1:
        raise FilterError("where_document expects a dict with exactly one key per level")
    (op, val), = where_document.items()
    if op in ("$and", "$or"):
        if not isinstance(val, (list, tuple)) or not val:
            raise FilterError(f"{op} expects a non-empty list of filters")
        parts = [evaluate_document(w, texts) for w in val]
        out = parts[0].copy()
        for p in parts[1:]:
            out = (out & p) if op == "$and" else (out | p)
        return out
    if op in ("$contains", "$not_contains"):
        if not isinstance(val, str):
            raise FilterError(f"{op} expects a string")
        hit = np.fromiter((t is not None and val in t for t in texts), dtype=bool, count=n)
        return hit if op == "$contains" else ~hit
    raise FilterError(f"unknown where_document operator {op}")


def texts_csr(texts: Sequence[Optional[str]]):
    """Stored documents -> (offsets uint64 [n + 1], UTF-8 bytes uint8, present bool [n]): the text column of the device-side
    `where_document` (mqvs_column_set_text).  Byte-wise search over UTF-8 finds exactly the substrings `str.__contains__`
    finds (UTF-8 is self-synchronising).  Sequences that know their own layout (vectorstore._RowColumn: millions of
    bulk-loaded rows without a document) provide `utf8_csr` themselves."""
    own = getattr(texts, "utf8_csr", None)
    if own is not None:
        return own()
    n = len(texts)
    present = np.fromiter((t is not None for t in texts), dtype=bool, count=n)
    enc = [t.encode("utf-8", "surrogatepass") if t is not None else b"" for t in texts]
    offsets = np.zeros(n + 1, np.uint64)
    if n:
        np.cumsum(np.fromiter((len(b) for b in enc), dtype=np.uint64, count=n), out=offsets[1:])
    data = np.frombuffer(b"".join(enc), dtype=np.uint8)
    return offsets, data, present


def pack_mask(allowed: np.ndarray) -> np.ndarray:
    """bool[n] -> uint32[ceil(n/32)], bit r&31 of word r>>5 == allowed[r] (padding bits are 0)."""
    allowed = np.asarray(allowed, dtype=bool)
    n = allowed.shape[0]
    words = (n + 31) // 32
    if words == 0:
        return np.zeros(0, np.uint32)
    padded = np.zeros(words * 32, dtype=bool)
    padded[:n] = allowed
    return np.packbits(padded.reshape(words, 32), axis=1, bitorder="little").view(np.uint32).reshape(words).copy()


def compile_mask(where: Optional[Mapping[str, Any]], cols: MetadataColumns, live: Optional[np.ndarray],
                 where_document: Optional[Mapping[str, Any]] = None, texts: Optional[Sequence[Optional[str]]] = None):
    """-> (packed mask or None when every row is allowed, number of allowed rows)."""
    n = len(cols)
    allowed = evaluate(where, cols) if where else None
    if where_document:
        doc_ok = evaluate_document(where_document, texts if texts is not None else [None] * n)
        allowed = doc_ok if allowed is None else (allowed & doc_ok)
    if live is not None and not live.all():
        allowed = live.copy() if allowed is None else (allowed & live)
    if allowed is None:
        return None, n
    return pack_mask(allowed), int(allowed.sum())


# ------------------------------------------------------------------------------------------ device-side evaluation
class NotOnDevice(Exception):
    """This filter (or one of its columns) needs the host evaluator: mixed-type field, bool/None constants, ..."""


def _typed_column(vals: np.ndarray, present: np.ndarray):
    """-> (kind, int64/float64 array, dictionary or None); kind in {"i", "f", "s", "n"}.  Raises NotOnDevice."""
    if not present.any():
        return "n", np.zeros(vals.shape[0], np.int64), None
    if vals.dtype != object:
        k = vals.dtype.kind
        if k in "iu" and vals.dtype.itemsize <= 8 and vals.dtype != np.uint64:
            return "i", vals.astype(np.int64), None
        if k == "f":
            return "f", vals.astype(np.float64), None
        if k in "US":
            uniq, codes = np.unique(vals[present], return_inverse=True)
            out = np.zeros(vals.shape[0], np.int64)
            out[present] = codes
            return "s", out, uniq
        raise NotOnDevice(f"column dtype {vals.dtype}")
    have = [v for v, p in zip(vals, present) if p]
    if not have:  # no row has the field: only the "missing" rule of each operator matters
        return "n", np.zeros(vals.shape[0], np.int64), None
    if all(isinstance(v, str) for v in have):
        uniq = np.unique(np.asarray(have, dtype=str)) if have else np.zeros(0, dtype=str)
        out = np.zeros(vals.shape[0], np.int64)
        if have:
            out[present] = np.searchsorted(uniq, np.asarray(have, dtype=str))
        return "s", out, uniq
    if all(isinstance(v, (int, np.integer)) and not isinstance(v, (bool, np.bool_)) for v in have):
        if any(abs(int(v)) >= 2 ** 63 for v in have):
            raise NotOnDevice("integer beyond int64")
        out = np.zeros(vals.shape[0], np.int64)
        out[present] = [int(v) for v in have]
        return "i", out, None
    if all(isinstance(v, (int, float, np.integer, np.floating)) and not isinstance(v, (bool, np.bool_)) for v in have):
        if any(isinstance(v, (int, np.integer)) and abs(int(v)) > 2 ** 53 for v in have):
            raise NotOnDevice("integer not exact in float64")
        out = np.zeros(vals.shape[0], np.float64)
        out[present] = [float(v) for v in have]
        return "f", out, None
    raise NotOnDevice("mixed-type field")


class DeviceFilter:
    """Compiles Chroma `where` / `where_document` dicts to mqvs_filter_op programs and keeps the referenced metadata
    columns (and, for where_document, the documents' UTF-8 text) resident on the device, re-uploaded when
    MetadataColumns.version moves.  `index` is a _native.Index; `texts` the row-aligned documents (a sequence of
    Optional[str], or a callable returning the current one)."""

    def __init__(self, index: Any, cols: MetadataColumns, texts: Any = None):
        self.index, self.cols = index, cols
        self._texts = texts
        self._resident: Dict[str, tuple] = {}  # field -> (version, column id, kind, dictionary)

    def _text_column(self) -> int:
        """The documents as a CSR text column under its own column id -> column id."""
        key = "\0document"
        texts = self._texts() if callable(self._texts) else self._texts
        if texts is None or not hasattr(self.index, "column_set_text"):
            raise NotOnDevice("no documents / no text columns on this index")
        hit = self._resident.get(key)
        if hit is not None and hit[0] == self.cols.version and hit[2] == len(texts):
            return hit[1]
        if len(texts) != len(self.cols):
            raise NotOnDevice("documents and metadata out of step")
        offsets, data, present = texts_csr(texts)
        col_id = hit[1] if hit is not None else len(self._resident)
        if col_id >= 64:
            raise NotOnDevice("more than 64 filterable fields")
        self.index.column_set_text(col_id, offsets, data, None if present.all() else present)
        self._resident[key] = (self.cols.version, col_id, len(texts), None)
        return col_id

    def _column(self, field: str):
        hit = self._resident.get(field)
        if hit is not None and hit[0] == self.cols.version:
            return hit[1:]
        vals, present = self.cols.column(field)
        kind, typed, uniq = _typed_column(vals, present)
        col_id = hit[1] if hit is not None else len(self._resident)
        if col_id >= 64:
            raise NotOnDevice("more than 64 filterable fields")
        self.index.column_set(col_id, typed, None if present.all() else present)
        self._resident[field] = (self.cols.version, col_id, kind, uniq)
        return col_id, kind, uniq

    def _token_column(self, field: str):
        """Token-set view of a field (comma-joined strings): dictionary + CSR token lists, resident on the device under its
        own column id.  -> (column id, sorted token dictionary)."""
        key = field + "\0tokens"
        hit = self._resident.get(key)
        if hit is not None and hit[0] == self.cols.version:
            return hit[1], hit[3]
        vals, present = self.cols.column(field)
        lists = [tokens_of(v) if p else [] for v, p in zip(vals, present)]
        uniq = np.unique(np.asarray([t for ts in lists for t in ts], dtype=str)) if any(lists) else np.zeros(0, dtype=str)
        offsets = np.zeros(len(lists) + 1, np.uint64)
        np.cumsum([len(ts) for ts in lists], out=offsets[1:])
        flat = [t for ts in lists for t in ts]
        tokens = np.searchsorted(uniq, np.asarray(flat, dtype=str)).astype(np.int64) if flat else np.zeros(0, np.int64)
        col_id = hit[1] if hit is not None else len(self._resident)
        if col_id >= 64:
            raise NotOnDevice("more than 64 filterable fields")
        self.index.column_set_tokens(col_id, offsets, tokens)
        self._resident[key] = (self.cols.version, col_id, "t", uniq)
        return col_id, uniq

    def _const(self, kind: str, uniq, op: str, want: Any):
        """-> (op, ival, fval) for a scalar leaf; raises NotOnDevice where host and device semantics could differ."""
        if kind == "n":
            return op, 0, 0.0
        if isinstance(want, (bool, np.bool_)) or want is None:
            raise NotOnDevice("bool / None constant")
        if kind == "s":
            if not isinstance(want, str):
                raise NotOnDevice("non-string constant against a string field")
            lo, hi = int(np.searchsorted(uniq, want, "left")), int(np.searchsorted(uniq, want, "right"))
            if op in ("$eq", "$ne"):
                return op, (lo if hi > lo else -1), 0.0
            # codes are ranks in the sorted dictionary, so order comparisons carry over
            return {"$gt": ("$gte", hi), "$gte": ("$gte", lo), "$lt": ("$lt", lo), "$lte": ("$lt", hi)}[op] + (0.0,)
        if isinstance(want, str):
            raise NotOnDevice("string constant against a numeric field")
        if kind == "i":
            if isinstance(want, (float, np.floating)):
                if not float(want).is_integer():
                    raise NotOnDevice("fractional constant against an integer field")
                want = int(want)
            if abs(int(want)) >= 2 ** 63:
                raise NotOnDevice("constant beyond int64")
            return op, int(want), 0.0
        return op, 0, float(want)

    def program(self, where: Optional[Mapping[str, Any]], where_document: Optional[Mapping[str, Any]] = None):
        """-> (list of FilterOp, constant pool as uint64 bit patterns): `where`, then `where_document`, AND-ed."""
        from ._native import MAX_NEEDLE, OP, FilterOp

        ops: List[Any] = []
        pool: List[int] = []

        def emit(w: Mapping[str, Any]) -> None:
            if not isinstance(w, Mapping) or len(w) != 1:
                raise FilterError("where expects a dict with exactly one key per level (use $and / $or to combine)")
            (key, val), = w.items()
            if key in ("$and", "$or"):
                if not isinstance(val, (list, tuple)) or not val:
                    raise FilterError(f"{key} expects a non-empty list of filters")
                for i, part in enumerate(val):
                    emit(part)
                    if i:
                        ops.append(FilterOp(OP[key], 0, 0, 0.0, 0, 0))
                return
            if key.startswith("$"):
                raise FilterError(f"unknown logical operator {key}")
            if isinstance(val, Mapping):
                if len(val) != 1:
                    raise FilterError("an operator expression takes exactly one operator")
                (op, want), = val.items()
                if op not in _LEAF_OPS:
                    raise FilterError(f"unknown operator {op}")
            else:
                op, want = "$eq", val
            if op in _TOKEN_OPS:
                if not isinstance(want, str):
                    raise FilterError(f"{op} expects a string token")
                if not hasattr(self.index, "column_set_tokens"):
                    raise NotOnDevice("this index has no token columns")
                col_id, uniq = self._token_column(key)
                want = want.strip()
                at = int(np.searchsorted(uniq, want)) if uniq.shape[0] else 0
                tok = at if at < uniq.shape[0] and uniq[at] == want else -1
                ops.append(FilterOp(OP[op], col_id, tok, 0.0, 0, 0))
                return
            col_id, kind, uniq = self._column(key)
            if op in ("$in", "$nin"):
                if not isinstance(want, (list, tuple, set, frozenset, np.ndarray)):
                    raise FilterError(f"{op} expects a list, got {type(want).__name__}")
                off = len(pool)
                for w1 in list(want):
                    _, iv, fv = self._const(kind, uniq, "$eq", w1)
                    pool.append(int(np.float64(fv).view(np.uint64)) if kind == "f" else int(np.int64(iv).view(np.uint64)))
                ops.append(FilterOp(OP[op], col_id, 0, 0.0, off, len(pool) - off))
                return
            if isinstance(want, (list, dict, tuple, set)):
                raise FilterError(f"{op} expects a scalar, got {type(want).__name__}")
            op2, iv, fv = self._const(kind, uniq, op, want)
            ops.append(FilterOp(OP[op2], col_id, iv, fv, 0, 0))

        def emit_document(w: Mapping[str, Any]) -> None:
            if not isinstance(w, Mapping) or len(w) != 1:
                raise FilterError("where_document expects a dict with exactly one key per level")
            (op, val), = w.items()
            if op in ("$and", "$or"):
                if not isinstance(val, (list, tuple)) or not val:
                    raise FilterError(f"{op} expects a non-empty list of filters")
                for i, part in enumerate(val):
                    emit_document(part)
                    if i:
                        ops.append(FilterOp(OP[op], 0, 0, 0.0, 0, 0))
                return
            if op not in _TOKEN_OPS:
                raise FilterError(f"unknown where_document operator {op}")
            if not isinstance(val, str):
                raise FilterError(f"{op} expects a string")
            needle = val.encode("utf-8", "surrogatepass")
            if not 1 <= len(needle) <= MAX_NEEDLE:
                raise NotOnDevice("empty needle (matches every document) or one beyond the device limit")
            col_id = self._text_column()
            off = len(pool)
            padded = needle + b"\0" * (-len(needle) % 8)
            pool.extend(int(v) for v in np.frombuffer(padded, dtype="<u8"))
            ops.append(FilterOp(OP["$text_contains" if op == "$contains" else "$text_not_contains"], col_id, 0, 0.0, off, len(needle)))

        if where:
            emit(where)
        if where_document:
            emit_document(where_document)
            if where:
                ops.append(FilterOp(OP["$and"], 0, 0, 0.0, 0, 0))
        if len(ops) > 64:
            raise NotOnDevice("filter program longer than 64 operations")
        return ops, np.asarray(pool, dtype=np.uint64)

    def build(self, where: Optional[Mapping[str, Any]], live: Optional[np.ndarray],
              where_document: Optional[Mapping[str, Any]] = None):
        """-> (DeviceMask, allowed rows).  Raises NotOnDevice when the host evaluator must be used."""
        ops, pool = self.program(where, where_document)
        live_words = None if live is None or live.all() else pack_mask(live)
        return self.index.filter_build(ops, pool, live_words)
