"""MediQueryVectorStore -- drop-in for the `langchain_chroma.Chroma` object MediQuery-RAG uses.

Call sites it replaces (file:line under /root/reference):
  src/ingest_medical.py:106-110   Chroma.from_documents(documents=docs, embedding=embeddings, persist_directory=DB_PATH)
  src/medical_engine.py:52        Chroma(persist_directory=DB_PATH, embedding_function=embeddings)
  src/agents/nodes.py:93          vectorstore.similarity_search(search_query, k=5)      -> [Document], best first
  src/medical_engine.py:70        vectorstore.similarity_search(query, k=k)             (k=3)
Swap = change `from langchain_chroma import Chroma` to
`from mediquery_rag_b200 import MediQueryVectorStore as Chroma`; nothing else in the reference changes.

Division of labour: documents, metadata and string ids stay on the host (Chroma keeps them in sqlite); the
vectors live in HBM behind libmqvs.so and every search is an EXACT scan there (no HNSW, no CPU fallback).
Scores come back as cosine similarity; `*_with_score` converts to the distance Chroma would report for the
collection's `hnsw:space` (default "l2", as in the reference, which passes no collection_metadata):
l2 -> 2 - 2s (= ||a-b||^2 on unit vectors), cosine / ip -> 1 - s.  Rows are stored L2-normalised (Ollama's
embedder already returns unit vectors).  In the l2 / ip spaces the device also keeps ||x|| of every raw row; as
long as every row is a unit vector (| ||x|| - 1 | <= 1e-5, the reference's case) the three spaces order rows
identically and the cosine kernels serve all of them; once a non-unit row is stored, searches rank by the space's
own score over the RAW vectors (mqvs_set_metric) and distances are ||q - x||^2 / 1 - <q, x> of the raw vectors.
"""
from __future__ import annotations

import bisect
import json
import os
import threading
import uuid
import warnings
from typing import Any, Callable, Dict, Iterable, List, Mapping, Optional, Sequence, Tuple

import numpy as np

from . import _native
from .documents import HAVE_LANGCHAIN, Document, Embeddings
from .filters import DeviceFilter, MetadataColumns, NotOnDevice, compile_mask

if HAVE_LANGCHAIN:  # pragma: no cover
    from langchain_core.vectorstores import VectorStore as _Base  # type: ignore
else:
    _Base = object

DEFAULT_K = 4  # langchain_chroma's default
_SPACES = ("l2", "cosine", "ip")
_ROWS_FILE, _DOCS_FILE, _MANIFEST, _DEAD_FILE = "rows.mqvs", "docs.jsonl", "manifest.json", "dead.json"
_NORMS_FILE = _ROWS_FILE + ".norms"  # written by mqvs_save next to the rows in the l2 / ip spaces
_UNIT_TOL = 1e-5  # a collection whose raw rows all satisfy | ||x|| - 1 | <= this is ranked as unit vectors


def _distance(space: str, sim: float) -> float:
    return 2.0 - 2.0 * sim if space == "l2" else 1.0 - sim


def _raw_distance(space: str, t: float, q_norm: float) -> float:
    """Chroma's distance over the RAW vectors from the backend's ranking score t (include/mqvs.h, mqvs_set_metric):
    ip: t = <q,x> / ||q|| -> 1 - <q,x>;  l2: t = (||q||^2 - ||q-x||^2) / (2 ||q||) -> ||q-x||^2."""
    qc = max(q_norm, 1e-12)
    return q_norm * q_norm - 2.0 * qc * t if space == "l2" else 1.0 - qc * t


def _relevance(space: str, sim: float) -> float:
    # langchain: cosine -> 1 - d ; l2 -> 1 - d / sqrt(2) ; ip -> d > 0 ? 1 - d : -d
    d = _distance(space, sim)
    if space == "cosine":
        return 1.0 - d
    if space == "l2":
        return 1.0 - d / (2.0 ** 0.5)
    return 1.0 - d if d > 0 else -d


class _RowColumn:
    """Row-aligned host column (string ids, texts) that does not need a Python object per row for bulk-loaded rows:
    a sequence of runs, each either an explicit list or an implicit run (`prefix + str(row)`, or None).  Indexing,
    len and extend are all the store needs."""

    def __init__(self) -> None:
        self._starts: List[int] = []   # first row of every run
        self._runs: List[Any] = []     # list | (count, prefix or None)
        self._n = 0

    def __len__(self) -> int:
        return self._n

    def extend(self, values: Sequence[Any]) -> None:
        if not len(values):
            return
        if self._runs and isinstance(self._runs[-1], list):
            self._runs[-1].extend(values)
        else:
            self._starts.append(self._n)
            self._runs.append(list(values))
        self._n += len(values)

    def append(self, value: Any) -> None:
        self.extend([value])

    def extend_implicit(self, n: int, prefix: Optional[str]) -> None:
        if n > 0:
            self._starts.append(self._n)
            self._runs.append((int(n), prefix))
            self._n += int(n)

    def __getitem__(self, r: int) -> Any:
        if r < 0:
            r += self._n
        if not 0 <= r < self._n:
            raise IndexError(r)
        i = bisect.bisect_right(self._starts, r) - 1
        run = self._runs[i]
        if isinstance(run, list):
            return run[r - self._starts[i]]
        return None if run[1] is None else f"{run[1]}{r}"

    def utf8_csr(self):
        """-> (offsets uint64 [n + 1], UTF-8 bytes, present bool [n]) of a text column (filters.texts_csr): explicit runs are
        encoded row by row, implicit runs without a value cost one slice assignment whatever their length."""
        lengths = np.zeros(self._n, np.uint64)
        present = np.zeros(self._n, bool)
        parts: List[bytes] = []
        for start, run in zip(self._starts, self._runs):
            if isinstance(run, list):
                enc = [v.encode("utf-8", "surrogatepass") if v is not None else b"" for v in run]
                present[start:start + len(run)] = [v is not None for v in run]
            elif run[1] is None:
                continue
            else:
                enc = [f"{run[1]}{r}".encode("utf-8") for r in range(start, start + run[0])]
                present[start:start + run[0]] = True
            lengths[start:start + len(enc)] = [len(b) for b in enc]
            parts.extend(enc)
        offsets = np.zeros(self._n + 1, np.uint64)
        np.cumsum(lengths, out=offsets[1:])
        return offsets, np.frombuffer(b"".join(parts), dtype=np.uint8), present

    def implicit_row(self, value: str) -> Optional[int]:
        """Row behind an implicit id (`prefix + row`), or None."""
        for start, run in zip(self._starts, self._runs):
            if not isinstance(run, list) and run[1] is not None and value.startswith(run[1]):
                tail = value[len(run[1]):]
                if tail.isdigit() and start <= int(tail) < start + run[0]:
                    return int(tail)
        return None


class _CollectionView:
    """`vectorstore._collection`: the part of chromadb's Collection that code written against langchain_chroma reaches for
    (`_collection.count()`, `.query(query_embeddings=..., n_results=...)`, `.get`, `.delete`, `.name`, `.metadata`).  A view
    over the store: nothing is copied, every call lands on the same kernels."""

    def __init__(self, store: "MediQueryVectorStore") -> None:
        self._store = store

    @property
    def name(self) -> str:
        return self._store._collection_name

    @property
    def metadata(self) -> Dict[str, Any]:
        return dict(self._store._collection_metadata)

    def count(self) -> int:
        return len(self._store)

    def get(self, ids=None, where=None, limit=None, offset=None, where_document=None, include=None) -> Dict[str, Any]:
        return self._store.get(ids=ids, where=where, limit=limit, offset=offset, where_document=where_document, include=include)

    def delete(self, ids=None, where=None, where_document=None) -> None:
        self._store.delete(ids=ids, where=where, where_document=where_document)

    def query(self, query_embeddings=None, query_texts=None, n_results: int = 10, where=None, where_document=None,
              include: Optional[List[str]] = None) -> Dict[str, Any]:
        """chromadb's collection.query: one result list PER query (the batch goes through one search), distances in the
        collection's hnsw:space, best first."""
        st = self._store
        if query_embeddings is None:
            if query_texts is None:
                raise ValueError("query needs query_embeddings or query_texts")
            texts = [query_texts] if isinstance(query_texts, str) else list(query_texts)
            if st._embedding_function is None:
                raise ValueError("this store was built without an embedding_function; pass query_embeddings")
            query_embeddings = st._embedding_function.embed_documents(texts)
        q = np.ascontiguousarray(query_embeddings, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        include = list(include) if include is not None else ["metadatas", "documents", "distances"]
        with st._lock:
            scores, rows = st.search_batch(q, n_results, where, where_document)
            ranked = st.rank_space()
            if ranked == "cosine":
                dist = np.where(rows >= 0, (2.0 - 2.0 * scores) if st._space == "l2" else (1.0 - scores), np.inf)
            else:
                qn = np.linalg.norm(q.astype(np.float64), axis=1)[:, None]
                t = scores.astype(np.float64)
                dist = np.where(rows >= 0, qn * qn - 2.0 * np.maximum(qn, 1e-12) * t if ranked == "l2"
                                else 1.0 - np.maximum(qn, 1e-12) * t, np.inf)
            per_query = [[int(r) for r in row if r >= 0] for row in rows]
            out: Dict[str, Any] = {"ids": [[st._ids[r] for r in rr] for rr in per_query], "distances": None, "metadatas": None,
                                   "documents": None, "embeddings": None, "included": include}
            if "distances" in include:
                out["distances"] = [[float(d) for d in dd[:len(rr)]] for dd, rr in zip(dist, per_query)]
            if "metadatas" in include:
                out["metadatas"] = [[st._cols.row(r) for r in rr] for rr in per_query]
            if "documents" in include:
                out["documents"] = [[st._texts[r] for r in rr] for rr in per_query]
            if "embeddings" in include:
                out["embeddings"] = [st.get(ids=[st._ids[r] for r in rr], include=["embeddings"])["embeddings"] for rr in per_query]
        return out


class MediQueryRetriever:
    """What `as_retriever()` returns when langchain_core is absent: `.invoke(query) -> list[Document]`."""

    def __init__(self, store: "MediQueryVectorStore", search_type: str = "similarity",
                 search_kwargs: Optional[Dict[str, Any]] = None):
        if search_type not in ("similarity", "similarity_score_threshold", "mmr"):
            raise ValueError(f"search_type of {search_type} not allowed")
        self.vectorstore, self.search_type, self.search_kwargs = store, search_type, dict(search_kwargs or {})
        if search_type == "similarity_score_threshold" and "score_threshold" not in self.search_kwargs:
            raise ValueError("`score_threshold` must be given in search_kwargs for similarity_score_threshold")

    def invoke(self, query: str, config: Any = None, **kwargs: Any) -> List[Document]:
        kw = {**self.search_kwargs, **kwargs}
        if self.search_type == "similarity":
            return self.vectorstore.similarity_search(query, **kw)
        if self.search_type == "mmr":
            return self.vectorstore.max_marginal_relevance_search(query, **kw)
        thr = kw.pop("score_threshold")
        return [d for d, s in self.vectorstore.similarity_search_with_relevance_scores(query, **kw) if s >= thr]

    get_relevant_documents = invoke


class MediQueryVectorStore(_Base):  # type: ignore[misc]
    def __init__(self, collection_name: str = "langchain", embedding_function: Optional[Embeddings] = None,
                 persist_directory: Optional[str] = None, collection_metadata: Optional[Dict[str, Any]] = None,
                 client_settings: Any = None, client: Any = None, relevance_score_fn: Optional[Callable] = None,
                 create_collection_if_not_exists: bool = True, *, store_dtype: str = "fp32", device: int = 0,
                 devices: Optional[Sequence[int]] = None, dim: Optional[int] = None, capacity_rows: int = 0,
                 auto_persist: bool = True, compact_dead_fraction: float = 0.5, scoring_shadow: Optional[bool] = None,
                 index_factory: Optional[Callable[..., Any]] = None, **_: Any):
        """`devices=[0, 1, ..., 7]` spreads the rows over several GPUs behind this ONE object in this one process (the
        reference holds its store as a module global, src/medical_engine.py:52): block-cyclic row shards, every search
        scans all of them concurrently and merges on the first device; ids, filters, persistence are unchanged.
        `scoring_shadow`: fp32 collections keep a bf16 copy of their rows for the tensor-core scan (half the bytes streamed
        per search, +50 % HBM; results are identical because every returned score is recomputed from the fp32 rows --
        include/mqvs.h, mqvs_set_shadow).  None = the library default (on), False = scan the fp32 rows themselves."""
        self._collection_name = collection_name
        self._embedding_function = embedding_function
        self._persist_directory = persist_directory
        self._collection_metadata = dict(collection_metadata or {})
        space = self._collection_metadata.get("hnsw:space", "l2")
        if space not in _SPACES:
            raise ValueError(f"hnsw:space must be one of {_SPACES}, got {space!r}")
        self._space = space
        self._relevance_fn = relevance_score_fn
        if store_dtype not in _native.DTYPES:
            raise ValueError(f"store_dtype must be one of {sorted(_native.DTYPES)}")
        self._store_dtype = store_dtype
        self._device = list(devices) if devices is not None and len(devices) > 1 else (devices[0] if devices else device)
        self._capacity = capacity_rows
        self._scoring_shadow = scoring_shadow
        self._compact_at = float(compact_dead_fraction)
        # one lock over everything that touches the host tables or the handle's shared filter mask: the a* coroutines run
        # the blocking calls in worker threads, and a device-built mask belongs to the handle until the search has used it
        self._lock = threading.RLock()
        self._foreign_chroma = False
        self._auto_persist = auto_persist
        self._index_factory = index_factory or _native.Index
        self._index: Any = None
        self._dim: Optional[int] = dim
        self._ids = _RowColumn()
        self._id2row: Dict[str, int] = {}
        self._texts = _RowColumn()
        self._cols = MetadataColumns()
        self._live = np.zeros(0, dtype=bool)
        self._live_key, self._live_all = None, True
        self._device_filter: Optional[DeviceFilter] = None
        self._raw_space = space != "cosine"  # keep ||x|| on the device so l2 / ip can rank non-unit vectors
        self._persisted_rows = 0            # rows already in rows.mqvs / docs.jsonl (appends start here)
        self._persisted_dead: Optional[List[int]] = None
        self._dir: Optional[str] = None  # where THIS collection's files live (see _collection_dir)
        if persist_directory is not None:
            # the reference tests os.path.exists(DB_PATH) before opening (src/medical_engine.py:34-37)
            os.makedirs(persist_directory, exist_ok=True)
            self._dir = self._collection_dir(persist_directory, collection_name)
            if os.path.exists(os.path.join(self._dir, _MANIFEST)):
                self._load()
            elif os.path.exists(os.path.join(persist_directory, "chroma.sqlite3")):
                # a directory written by the reference's real Chroma (sqlite + hnsw segments) holds nothing this backend
                # can open: the collection starts empty, and searching it before re-ingesting is an error, not []
                self._foreign_chroma = True
                warnings.warn(f"{persist_directory} was written by chromadb (chroma.sqlite3); mqvs cannot read it -- re-run "
                              "src/ingest_medical.py with the swapped import to rebuild the collection", RuntimeWarning)
        if self._index is None and self._dim is not None:
            self._ensure_index(self._dim)

    # ------------------------------------------------------------------------------------------- basics
    @staticmethod
    def _collection_dir(persist_directory: str, collection_name: str) -> str:
        """Chroma keeps several collections in one persist_directory; here every collection owns a sub-directory
        `mqvs.<name>/`, so opening, persisting or deleting one never touches another.  A directory written before
        collections were namespaced (files at the top level) keeps that layout for the collection that owns it."""
        legacy = os.path.join(persist_directory, _MANIFEST)
        if os.path.exists(legacy):
            try:
                with open(legacy, "r", encoding="utf-8") as f:
                    if json.load(f).get("collection_name", collection_name) == collection_name:
                        return persist_directory
            except (OSError, ValueError):
                pass
        safe = "".join(c if c.isalnum() or c in "-_." else "_" for c in collection_name) or "_"
        return os.path.join(persist_directory, "mqvs." + safe)

    @property
    def collection_path(self) -> Optional[str]:
        """Directory that holds this collection's files (rows.mqvs, docs.jsonl, dead.json, manifest.json), or None."""
        return self._dir

    @property
    def embeddings(self) -> Optional[Embeddings]:
        return self._embedding_function

    @property
    def _collection(self) -> _CollectionView:
        """What `langchain_chroma.Chroma._collection` is used for in the wild: `.count()`, `.query(...)`, `.get`, `.delete`."""
        return _CollectionView(self)

    def __len__(self) -> int:
        return int(self._live.sum())

    def _ensure_index(self, dim: int) -> None:
        if self._index is None:
            self._dim = int(dim)
            self._index = self._index_factory(self._dim, self._store_dtype, self._device, self._capacity)
            if self._scoring_shadow is not None and hasattr(self._index, "set_shadow"):
                self._index.set_shadow(self._scoring_shadow)
            if self._raw_space and self._space != "cosine" and hasattr(self._index, "set_metric"):
                self._index.set_metric(self._space)  # while empty: row norms are kept from here on
            else:
                self._raw_space = False
        elif dim != self._dim:
            raise ValueError(f"embedding dimension {dim} does not match the collection's {self._dim}")

    def _embed_query(self, query: str) -> np.ndarray:
        if self._embedding_function is None:
            raise ValueError("this store was built without an embedding_function; use the *_by_vector methods")
        return np.asarray(self._embedding_function.embed_query(query), dtype=np.float32)

    # ------------------------------------------------------------------------------------------- ingest
    def add_texts(self, texts: Iterable[str], metadatas: Optional[List[Optional[dict]]] = None,
                  ids: Optional[List[str]] = None, **kwargs: Any) -> List[str]:
        texts = list(texts)
        if self._embedding_function is None:
            raise ValueError("add_texts needs an embedding_function; use add_embeddings for pre-computed vectors")
        if not texts:
            return []
        vectors = np.asarray(self._embedding_function.embed_documents(texts), dtype=np.float32)
        return self._add(texts, vectors, metadatas, ids)

    def add_documents(self, documents: List[Document], **kwargs: Any) -> List[str]:
        ids = kwargs.pop("ids", None)
        if ids is None and documents and all(getattr(d, "id", None) for d in documents):
            ids = [d.id for d in documents]
        return self.add_texts([d.page_content for d in documents], [d.metadata for d in documents], ids=ids)

    def add_embeddings(self, text_embeddings: Any, metadatas: Optional[List[Optional[dict]]] = None,
                       ids: Optional[List[str]] = None, *, metadata_columns: Optional[Mapping[str, Any]] = None,
                       **kwargs: Any) -> List[str]:
        """Pre-computed vectors.  `text_embeddings` is an iterable of (text, vector) pairs (FAISS style) or a
        (texts, embeddings) pair where embeddings is an [n, dim] numpy array or CUDA torch tensor (zero-copy
        via data_ptr) and texts is a list or None.  `metadata_columns` gives metadata column-wise for bulk loads."""
        if isinstance(text_embeddings, tuple) and len(text_embeddings) == 2 and hasattr(text_embeddings[1], "shape"):
            texts, vectors = text_embeddings
            n = vectors.shape[0]
            texts = list(texts) if texts is not None else [None] * n
        else:
            pairs = list(text_embeddings)
            texts = [t for t, _ in pairs]
            vectors = np.asarray([v for _, v in pairs], dtype=np.float32)
        return self._add(texts, vectors, metadatas, ids, metadata_columns)

    def _add(self, texts: List[Optional[str]], vectors: Any, metadatas, ids, metadata_columns=None) -> List[str]:
        with self._lock:
            return self._add_locked(texts, vectors, metadatas, ids, metadata_columns)

    @staticmethod
    def _check_columns(n: int, metadatas, metadata_columns) -> None:
        if metadata_columns is None:
            return
        if metadatas is not None:
            raise ValueError("give metadatas or metadata_columns, not both")
        for name, col in metadata_columns.items():
            if np.asarray(col).shape[:1] != (n,):
                raise ValueError(f"column {name!r} has shape {np.asarray(col).shape} for {n} rows")

    def _add_locked(self, texts: List[Optional[str]], vectors: Any, metadatas, ids, metadata_columns=None) -> List[str]:
        n = len(texts)
        if n == 0:
            return []
        if not hasattr(vectors, "data_ptr"):
            vectors = np.ascontiguousarray(vectors, dtype=np.float32)
        if vectors.ndim != 2 or vectors.shape[0] != n:
            raise ValueError(f"expected {n} embeddings of one width, got shape {tuple(vectors.shape)}")
        if metadatas is not None and len(metadatas) != n:
            raise ValueError(f"{len(metadatas)} metadatas for {n} texts")
        if metadatas is not None and not all(md is None or isinstance(md, Mapping) for md in metadatas):
            raise TypeError("every metadata must be a mapping or None")
        if ids is None:
            ids = [str(uuid.uuid4()) for _ in range(n)]
        ids = [str(i) for i in ids]
        if len(ids) != n:
            raise ValueError(f"{len(ids)} ids for {n} texts")
        if len(set(ids)) != n:
            raise ValueError("duplicate ids within one call")
        self._check_columns(n, metadatas, metadata_columns)  # before anything reaches the device: rows and tables move together
        self._ensure_index(vectors.shape[1])
        first = self._index.add(vectors)  # raises (and appends nothing) on non-finite rows
        assert first == len(self._ids), "host document table and device rows out of step"
        # upsert semantics: a repeated id retires its previous row (Chroma's collection.upsert)
        self._live = np.concatenate([self._live, np.ones(n, dtype=bool)])
        for j, doc_id in enumerate(ids):
            old = self._id2row.get(doc_id)
            if old is not None:
                self._live[old] = False
                self._live_key = None
            self._id2row[doc_id] = first + j
        self._ids.extend(ids)
        self._texts.extend(texts)
        if metadata_columns is not None:
            self._cols.extend_columns(n, metadata_columns)
        else:
            self._cols.extend(metadatas if metadatas is not None else [None] * n)
        if self._persist_directory is not None and self._auto_persist:
            self.persist()
        return ids

    def add_synthetic(self, seed: int, n: int, first_index: Optional[int] = None,
                      metadata_columns: Optional[Mapping[str, Any]] = None, id_prefix: str = "row-") -> range:
        """Bench / test utility behind BASELINE.json's configs 2-5 (synthetic 768-d unit vectors; 100M rows cannot be staged
        through the host): appends n rows generated ON THE DEVICE by the counter-based generator of mqvs_add_synthetic
        (row r is a function of (seed, first_index + r) only).  Ids are implicit (`id_prefix + row`), texts empty, metadata
        optional and column-wise -- nothing is kept per row on the host.  Not persisted (in-memory collections only)."""
        if self._persist_directory is not None:
            raise ValueError("add_synthetic serves in-memory collections (persist_directory=None)")
        self._check_columns(int(n), None, metadata_columns)
        with self._lock:
            self._ensure_index(self._dim or _native.DIM_PAD)
            start = len(self._ids)
            first = self._index.add_synthetic(seed, start if first_index is None else first_index, int(n))
            assert first == start, "host document table and device rows out of step"
            self._live = np.concatenate([self._live, np.ones(n, dtype=bool)])
            self._ids.extend_implicit(n, id_prefix)
            self._texts.extend_implicit(n, None)
            if metadata_columns is not None:
                self._cols.extend_columns(n, metadata_columns)
            else:
                self._cols.extend_empty(n)
            return range(start, start + n)

    def _row_of(self, doc_id: str) -> Optional[int]:
        r = self._id2row.get(doc_id)
        if r is None:
            r = self._ids.implicit_row(doc_id)
            if r is not None and not self._live[r]:
                r = None
        return r

    @classmethod
    def from_texts(cls, texts: List[str], embedding: Optional[Embeddings] = None,
                   metadatas: Optional[List[dict]] = None, ids: Optional[List[str]] = None,
                   collection_name: str = "langchain", persist_directory: Optional[str] = None,
                   collection_metadata: Optional[Dict[str, Any]] = None, **kwargs: Any) -> "MediQueryVectorStore":
        store = cls(collection_name=collection_name, embedding_function=embedding,
                    persist_directory=persist_directory, collection_metadata=collection_metadata, **kwargs)
        store.add_texts(texts, metadatas=metadatas, ids=ids)
        return store

    @classmethod
    def from_documents(cls, documents: List[Document], embedding: Optional[Embeddings] = None,
                       ids: Optional[List[str]] = None, collection_name: str = "langchain",
                       persist_directory: Optional[str] = None, collection_metadata: Optional[Dict[str, Any]] = None,
                       **kwargs: Any) -> "MediQueryVectorStore":
        if ids is None and documents and all(getattr(d, "id", None) for d in documents):
            ids = [d.id for d in documents]
        return cls.from_texts([d.page_content for d in documents], embedding, [d.metadata for d in documents], ids,
                              collection_name, persist_directory, collection_metadata, **kwargs)

    # ------------------------------------------------------------------------------------------- search
    def rank_space(self) -> str:
        """The space the next search ranks in: "cosine" while every stored raw row is a unit vector (all spaces agree
        there and every kernel path is available), else the collection's hnsw:space over the raw vectors."""
        if not self._raw_space or self._index is None:
            return "cosine"
        want = self._space if self._index.norm_deviation() > _UNIT_TOL else "cosine"
        if self._index.metric != want:
            self._index.set_metric(want)
        return want

    def search_batch(self, queries: Any, k: int = DEFAULT_K, filter: Optional[Mapping[str, Any]] = None,
                     where_document: Optional[Mapping[str, Any]] = None) -> Tuple[np.ndarray, np.ndarray]:
        """Batch extension used by the benchmark configs: queries [B, dim] -> (scores [B, k'], row ids [B, k']) with
        k' = min(k, allowed rows); rows best-first, ties -> lower row id.  Scores are cosine similarities, or the
        ranking score of `rank_space()` when that is "l2" / "ip" (non-unit rows; include/mqvs.h)."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if k < 1:
            raise ValueError("k must be >= 1")
        with self._lock:  # filter_build + search share the handle's mask buffer; the tables must not move underneath
            if self._index is None or len(self._ids) == 0:
                if self._foreign_chroma:
                    raise RuntimeError(f"{self._persist_directory} holds a chromadb collection (chroma.sqlite3) that mqvs "
                                       "cannot read; re-run src/ingest_medical.py with the swapped import before searching")
                return np.zeros((q.shape[0], 0), np.float32), np.zeros((q.shape[0], 0), np.int64)
            mask, allowed = self._allow_mask(filter, where_document)
            k_eff = min(k, allowed)  # Chroma clamps n_results to the collection size
            if k_eff == 0:
                return np.zeros((q.shape[0], 0), np.float32), np.zeros((q.shape[0], 0), np.int64)
            if k_eff > _native.MAX_K:
                raise ValueError(f"k is limited to {_native.MAX_K}")
            self.rank_space()
            return self._index.search(q, k_eff, mask)

    def _allow_mask(self, filter: Optional[Mapping[str, Any]], where_document: Optional[Mapping[str, Any]]):
        """-> (mask or None, allowed rows).  Metadata filters over typed fields are evaluated ON THE DEVICE (columns stay
        resident in HBM, the bitmask never visits the host) and so is where_document ($contains over the documents' UTF-8
        text, resident as a CSR text column); mixed-type fields, bool / None constants and empty needles use the host
        evaluator.  Both produce the same bitmask (tests/test_gpu_parity.py)."""
        if not filter and not where_document and self._all_live():
            return None, len(self._ids)  # nothing to mask: no pass over the live bitmap either (100M rows = 30 ms of numpy)
        if (filter or where_document) and hasattr(self._index, "filter_build"):
            if (self._device_filter is None or self._device_filter.index is not self._index
                    or self._device_filter.cols is not self._cols):
                self._device_filter = DeviceFilter(self._index, self._cols, lambda: self._texts)
            try:
                return self._device_filter.build(filter, self._live, where_document)
            except NotOnDevice:
                pass
        return compile_mask(filter, self._cols, self._live, where_document, self._texts)

    def _all_live(self) -> bool:
        """No tombstones?  Cached per mutation of the live bitmap (identity + length change on every add; deletes reset it)."""
        key = (id(self._live), self._live.shape[0])
        if self._live_key != key:
            self._live_key, self._live_all = key, bool(self._live.all())
        return self._live_all

    def _docs(self, scores: np.ndarray, rows: np.ndarray) -> List[Tuple[Document, float]]:
        out = []
        for s, r in zip(scores.tolist(), rows.tolist()):
            if r < 0:
                continue
            out.append((Document(page_content=self._texts[r] or "", metadata=self._cols.row(r), id=self._ids[r]), s))
        return out

    def _similar(self, vector: Any, k: int, filter: Optional[Mapping[str, Any]],
                 where_document: Optional[Mapping[str, Any]] = None) -> List[Tuple[Document, float]]:
        """(Document, distance in the collection's hnsw:space) pairs, best (smallest distance) first."""
        q = np.asarray(vector, dtype=np.float32)
        with self._lock:  # row numbers -> documents under the same lock as the search: a compaction in between renumbers rows
            scores, rows = self.search_batch(q[None, :], k, filter, where_document)
            ranked = self.rank_space()
            hits = self._docs(scores[0], rows[0])
        if ranked == "cosine":
            return [(d, _distance(self._space, s)) for d, s in hits]
        q_norm = float(np.linalg.norm(q.astype(np.float64)))
        return [(d, _raw_distance(ranked, t, q_norm)) for d, t in hits]

    def similarity_search(self, query: str, k: int = DEFAULT_K, filter: Optional[Dict[str, Any]] = None,
                          **kwargs: Any) -> List[Document]:
        return [d for d, _ in self._similar(self._embed_query(query), k, filter, kwargs.get("where_document"))]

    def similarity_search_with_score(self, query: str, k: int = DEFAULT_K, filter: Optional[Dict[str, Any]] = None,
                                     where_document: Optional[Dict[str, Any]] = None, **kwargs: Any
                                     ) -> List[Tuple[Document, float]]:
        """(Document, distance) pairs, LOWER is better, in the collection's hnsw:space."""
        return self._similar(self._embed_query(query), k, filter, where_document)

    def similarity_search_by_vector(self, embedding: List[float], k: int = DEFAULT_K,
                                    filter: Optional[Dict[str, Any]] = None, **kwargs: Any) -> List[Document]:
        return [d for d, _ in self._similar(embedding, k, filter, kwargs.get("where_document"))]

    def similarity_search_by_vector_with_relevance_scores(self, embedding: List[float], k: int = DEFAULT_K,
                                                          filter: Optional[Dict[str, Any]] = None, **kwargs: Any
                                                          ) -> List[Tuple[Document, float]]:
        # langchain_chroma returns the raw distance here despite the name; kept for compatibility
        return self._similar(embedding, k, filter)

    def similarity_search_with_vectors(self, query: str, k: int = DEFAULT_K, filter: Optional[Dict[str, Any]] = None,
                                       where_document: Optional[Dict[str, Any]] = None, **kwargs: Any
                                       ) -> List[Tuple[Document, np.ndarray]]:
        """langchain_chroma's (Document, stored embedding) pairs, best first: the vectors as the collection returns them for
        `get(include=["embeddings"])` (raw vectors in an l2 / ip collection, unit rows otherwise)."""
        q = self._embed_query(query)
        with self._lock:
            hits = self._similar(q, k, filter, where_document)
            if not hits:
                return []
            emb = self.get(ids=[d.id for d, _ in hits], include=["embeddings"])
        by_id = dict(zip(emb["ids"], emb["embeddings"]))
        return [(d, np.asarray(by_id[d.id])) for d, _ in hits]

    def _select_relevance_score_fn(self) -> Callable[[float], float]:
        if self._relevance_fn is not None:
            return self._relevance_fn
        space = self._space
        if space == "cosine":
            return lambda d: 1.0 - d
        if space == "l2":
            return lambda d: 1.0 - d / (2.0 ** 0.5)
        return lambda d: 1.0 - d if d > 0 else -d

    def similarity_search_with_relevance_scores(self, query: str, k: int = DEFAULT_K, **kwargs: Any
                                                ) -> List[Tuple[Document, float]]:
        thr = kwargs.pop("score_threshold", None)
        fn = self._select_relevance_score_fn()
        out = [(d, fn(dist)) for d, dist in self.similarity_search_with_score(query, k, kwargs.get("filter"))]
        return [(d, s) for d, s in out if thr is None or s >= thr]

    _similarity_search_with_relevance_scores = similarity_search_with_relevance_scores

    def max_marginal_relevance_search_by_vector(self, embedding: List[float], k: int = DEFAULT_K, fetch_k: int = 20,
                                                lambda_mult: float = 0.5, filter: Optional[Dict[str, Any]] = None,
                                                **kwargs: Any) -> List[Document]:
        """Shortlist fetch_k rows on the GPU, then the usual greedy MMR over those (host, fetch_k is tiny)."""
        q = np.asarray(embedding, dtype=np.float32)
        with self._lock:
            return self._mmr_locked(q, k, fetch_k, lambda_mult, filter)

    def _mmr_locked(self, q: np.ndarray, k: int, fetch_k: int, lambda_mult: float, filter) -> List[Document]:
        scores, rows = self.search_batch(q[None, :], fetch_k, filter)
        rows = rows[0][rows[0] >= 0]
        if rows.shape[0] == 0:
            return []
        cand = self._index.get_rows(rows).astype(np.float64)
        qn = q.astype(np.float64) / max(float(np.linalg.norm(q)), 1e-12)
        cn = cand / np.maximum(np.linalg.norm(cand, axis=1, keepdims=True), 1e-12)
        to_query = cn @ qn
        chosen = [int(np.argmax(to_query))]
        while len(chosen) < min(k, rows.shape[0]):
            redundancy = (cn @ cn[chosen].T).max(axis=1)
            mmr = lambda_mult * to_query - (1.0 - lambda_mult) * redundancy
            mmr[chosen] = -np.inf
            chosen.append(int(np.argmax(mmr)))
        return [d for d, _ in self._docs(scores[0][chosen], rows[chosen])]

    def max_marginal_relevance_search(self, query: str, k: int = DEFAULT_K, fetch_k: int = 20,
                                      lambda_mult: float = 0.5, filter: Optional[Dict[str, Any]] = None,
                                      **kwargs: Any) -> List[Document]:
        return self.max_marginal_relevance_search_by_vector(self._embed_query(query), k, fetch_k, lambda_mult, filter)

    # ------------------------------------------------------------------------------------------- update / async
    def update_document(self, document_id: str, document: Document) -> None:
        """langchain_chroma's update_document: re-embed and replace the row behind `document_id` (upsert: the old row is
        tombstoned, the new one appended -- ids stay stable for callers)."""
        self.update_documents([document_id], [document])

    def update_documents(self, ids: List[str], documents: List[Document]) -> None:
        with self._lock:
            missing = [i for i in ids if self._row_of(str(i)) is None]
            if missing:
                raise ValueError(f"unknown document ids: {missing[:3]}")
            self.add_texts([d.page_content for d in documents], [d.metadata for d in documents], ids=[str(i) for i in ids])

    # LangChain's VectorStore exposes a-prefixed coroutines that run the blocking call in a worker thread; the handle is
    # safe to use from any thread (tests/test_gpu_parity.py::test_one_handle_from_several_host_threads...).  With
    # langchain_core installed the base class provides them; these are the stand-ins otherwise.
    if not HAVE_LANGCHAIN:
        async def asimilarity_search(self, query: str, k: int = DEFAULT_K, **kwargs: Any) -> List[Document]:
            import asyncio
            return await asyncio.to_thread(self.similarity_search, query, k, **kwargs)

        async def asimilarity_search_with_score(self, query: str, k: int = DEFAULT_K, **kwargs: Any):
            import asyncio
            return await asyncio.to_thread(self.similarity_search_with_score, query, k, **kwargs)

        async def asimilarity_search_by_vector(self, embedding: List[float], k: int = DEFAULT_K, **kwargs: Any) -> List[Document]:
            import asyncio
            return await asyncio.to_thread(self.similarity_search_by_vector, embedding, k, **kwargs)

        async def amax_marginal_relevance_search(self, query: str, k: int = DEFAULT_K, fetch_k: int = 20,
                                                 lambda_mult: float = 0.5, **kwargs: Any) -> List[Document]:
            import asyncio
            return await asyncio.to_thread(self.max_marginal_relevance_search, query, k, fetch_k, lambda_mult, **kwargs)

        async def aadd_texts(self, texts: Iterable[str], metadatas: Optional[List[Optional[dict]]] = None, **kwargs: Any) -> List[str]:
            import asyncio
            return await asyncio.to_thread(self.add_texts, list(texts), metadatas, **kwargs)

        async def aadd_documents(self, documents: List[Document], **kwargs: Any) -> List[str]:
            import asyncio
            return await asyncio.to_thread(self.add_documents, documents, **kwargs)

    def as_retriever(self, **kwargs: Any) -> Any:
        if HAVE_LANGCHAIN:  # pragma: no cover
            return super().as_retriever(**kwargs)
        return MediQueryRetriever(self, kwargs.get("search_type", "similarity"), kwargs.get("search_kwargs"))

    # ------------------------------------------------------------------------------------------- get / delete
    def get(self, ids: Optional[Sequence[str]] = None, where: Optional[Mapping[str, Any]] = None,
            limit: Optional[int] = None, offset: Optional[int] = None, where_document: Any = None,
            include: Optional[List[str]] = None) -> Dict[str, Any]:
        with self._lock:  # the tables and the row numbering must not move (add / delete / compaction) while rows are collected
            return self._get_locked(ids, where, limit, offset, where_document, include)

    def _get_locked(self, ids, where, limit, offset, where_document, include) -> Dict[str, Any]:
        include = list(include) if include is not None else ["metadatas", "documents"]
        if ids is not None:
            if isinstance(ids, str):
                ids = [ids]
            rows = [r for r in (self._row_of(str(i)) for i in ids) if r is not None and self._live[r]]
            if where or where_document:
                from .filters import evaluate, evaluate_document
                ok = evaluate(where, self._cols) if where else np.ones(len(self._ids), bool)
                if where_document:
                    ok = ok & evaluate_document(where_document, self._texts)
                rows = [r for r in rows if ok[r]]
        else:
            from .filters import evaluate, evaluate_document
            ok = self._live if not where else (evaluate(where, self._cols) & self._live)
            if where_document:  # collection.get(where_document={"$contains": ...}), as in chromadb
                ok = ok & evaluate_document(where_document, self._texts)
            rows = np.flatnonzero(ok).tolist()
        rows = rows[(offset or 0):]
        if limit is not None:
            rows = rows[:limit]
        out: Dict[str, Any] = {"ids": [self._ids[r] for r in rows], "embeddings": None, "metadatas": None,
                               "documents": None, "included": include}
        if "metadatas" in include:
            out["metadatas"] = [self._cols.row(r) for r in rows]
        if "documents" in include:
            out["documents"] = [self._texts[r] for r in rows]
        if "embeddings" in include:
            emb = (self._index.get_rows(np.asarray(rows, np.int64)) if rows
                   else np.zeros((0, self._dim or 0), np.float32))
            if rows and self._raw_space:  # l2 / ip collections hand back the raw vectors: unit row x stored norm
                lo, hi = min(rows), max(rows) + 1
                emb = emb * self._index.get_norms(lo, hi - lo)[np.asarray(rows) - lo, None]
            out["embeddings"] = emb
        return out

    def get_by_ids(self, ids: Sequence[str]) -> List[Document]:
        got = self.get(ids=list(ids))
        return [Document(page_content=t or "", metadata=m, id=i)
                for i, t, m in zip(got["ids"], got["documents"], got["metadatas"])]

    def delete(self, ids: Optional[List[str]] = None, **kwargs: Any) -> Optional[bool]:
        """Rows are tombstoned (cleared from the live bitmap that is AND-ed into every search mask).  `where=` /
        `where_document=` keyword arguments select rows the way chromadb's collection.delete does (langchain_chroma passes
        them through); with `ids` as well, only those of the ids that match go."""
        where, where_document = kwargs.get("where"), kwargs.get("where_document")
        if ids is None and not where and not where_document:
            return False
        with self._lock:
            if where or where_document:
                ids = self.get(ids=ids, where=where, where_document=where_document, include=[])["ids"]
            for i in ids:
                r = self._id2row.pop(str(i), None)
                if r is None:
                    r = self._ids.implicit_row(str(i))
                if r is not None:
                    self._live[r] = False
                    self._live_key = None
            if self._persist_directory is not None and self._auto_persist:
                self.persist()
        return True

    def delete_collection(self) -> None:
        """Drops THIS collection: its rows, tables and files (other collections of the persist_directory are untouched)."""
        with self._lock:
            if self._index is not None:
                self._index.close()
            self._index, self._ids, self._id2row, self._texts = None, _RowColumn(), {}, _RowColumn()
            self._cols, self._live = MetadataColumns(), np.zeros(0, dtype=bool)
            self._device_filter = None
            if self._dir is not None:
                self._persisted_rows, self._persisted_dead = 0, None
                for f in (_ROWS_FILE, _NORMS_FILE, _DOCS_FILE, _MANIFEST, _DEAD_FILE):
                    try:
                        os.remove(os.path.join(self._dir, f))
                    except FileNotFoundError:
                        pass
                if self._dir != self._persist_directory:
                    try:
                        os.rmdir(self._dir)
                    except OSError:
                        pass

    def reset_collection(self) -> None:
        """langchain_chroma's reset_collection: drop every row and start again with an empty collection of the same name,
        space and storage type."""
        self.delete_collection()
        self._raw_space = self._space != "cosine"
        if self._dim is not None:
            self._ensure_index(self._dim)

    # ------------------------------------------------------------------------------------------- persistence
    def persist(self) -> None:
        """rows.mqvs (raw normalised rows, mqvs_save / mqvs_save_append) + docs.jsonl (id, text, metadata per row) +
        dead.json (tombstoned rows) + manifest.json.  APPEND-ONLY for rows and documents: adding to a store of N rows
        writes only the new rows (re-running the reference's ingest on an existing ./medical_db appends,
        src/ingest_medical.py:106-110), so the cost of `auto_persist` does not grow with the store.  Order of writes:
        documents, rows (their count in the file header moves last), tombstones, manifest -- a crash in between leaves a
        directory that still opens (surplus document lines are ignored)."""
        d = self._dir
        if d is None:
            raise ValueError("no persist_directory")
        with self._lock:
            self._persist_locked(d)

    def _persist_locked(self, d: str) -> None:
        os.makedirs(d, exist_ok=True)
        n_dead = int((~self._live).sum())
        if self._compact_at > 0 and len(self._ids) >= 1024 and n_dead > self._compact_at * len(self._ids):
            self._compact_locked()  # tombstones cost scan bandwidth for ever otherwise; also resets the append position
        n, done = len(self._ids), self._persisted_rows
        rows_path, docs_path = os.path.join(d, _ROWS_FILE), os.path.join(d, _DOCS_FILE)
        append = 0 < done <= n and os.path.exists(rows_path) and os.path.exists(docs_path)

        def record(r: int) -> str:
            return json.dumps({"id": self._ids[r], "text": self._texts[r], "metadata": self._cols.row(r)}, ensure_ascii=False) + "\n"

        if append:
            if n > done:
                with open(docs_path, "a", encoding="utf-8") as f:
                    for r in range(done, n):
                        f.write(record(r))
                    f.flush()
                    os.fsync(f.fileno())
                self._index.save_append(rows_path, done)
        else:
            tmp = docs_path + ".tmp"
            with open(tmp, "w", encoding="utf-8") as f:
                for r in range(n):
                    f.write(record(r))
            os.replace(tmp, docs_path)
            if self._index is not None:
                self._index.save(rows_path)
        dead = np.flatnonzero(~self._live).tolist()
        if dead != self._persisted_dead:
            tmp = os.path.join(d, _DEAD_FILE + ".tmp")
            with open(tmp, "w", encoding="utf-8") as f:
                json.dump(dead, f)
            os.replace(tmp, os.path.join(d, _DEAD_FILE))
            self._persisted_dead = dead
        manifest = {"format": 2, "collection_name": self._collection_name, "dim": self._dim,
                    "store_dtype": self._store_dtype, "rows": n, "space": self._space,
                    "collection_metadata": self._collection_metadata}
        tmp = os.path.join(d, _MANIFEST + ".tmp")
        with open(tmp, "w", encoding="utf-8") as f:
            json.dump(manifest, f, ensure_ascii=False, indent=1)
        os.replace(tmp, os.path.join(d, _MANIFEST))
        self._persisted_rows = n

    def compact(self) -> int:
        """Drops tombstoned rows (deleted / upserted) from HBM and from the host tables: the stored rows are gathered
        bit-for-bit on the device (mqvs_compact), row ids are renumbered, string ids stay.  Runs by itself from persist()
        once more than `compact_dead_fraction` of the rows are dead.  Returns the number of rows dropped."""
        with self._lock:
            dropped = self._compact_locked()
            if dropped and self._dir is not None and self._auto_persist:
                self._persist_locked(self._dir)
            return dropped

    def _compact_locked(self) -> int:
        dead = int((~self._live).sum())
        if dead == 0 or self._index is None or not hasattr(self._index, "compact"):
            return 0
        from .filters import pack_mask
        keep = np.flatnonzero(self._live)
        new_n = self._index.compact(pack_mask(self._live))
        assert new_n == keep.shape[0], "device and host disagree on the live rows"
        ids, texts = _RowColumn(), _RowColumn()
        ids.extend([self._ids[r] for r in keep.tolist()])
        texts.extend([self._texts[r] for r in keep.tolist()])
        metas = [self._cols.row(r) for r in keep.tolist()]
        self._ids, self._texts = ids, texts
        self._cols = MetadataColumns()
        self._cols.extend(metas)
        self._live = np.ones(new_n, dtype=bool)
        self._id2row = {self._ids[j]: j for j in range(new_n)}
        self._device_filter = None
        self._persisted_rows, self._persisted_dead = 0, None  # the files are rewritten whole on the next persist
        return dead

    def _load(self) -> None:
        d = self._dir
        with open(os.path.join(d, _MANIFEST), "r", encoding="utf-8") as f:
            manifest = json.load(f)
        self._store_dtype = manifest["store_dtype"]
        if "hnsw:space" not in self._collection_metadata:
            self._space = manifest.get("space", self._space)
        if manifest.get("dim") is None:
            return
        # a directory written before row norms were kept (or in the cosine space) opens with unit-vector semantics
        self._raw_space = self._space != "cosine" and (not manifest["rows"] or os.path.exists(os.path.join(d, _NORMS_FILE)))
        self._ensure_index(manifest["dim"])
        if manifest["rows"]:
            self._index.load_file(os.path.join(d, _ROWS_FILE))
        n = len(self._index)  # the row file is the authority: its header moves only after the rows are on disk
        metas: List[dict] = []
        live: List[bool] = []
        lines = 0
        with open(os.path.join(d, _DOCS_FILE), "r", encoding="utf-8") as f:
            for line in f:
                lines += 1
                if len(self._ids) == n:
                    continue  # document lines of an append whose rows never made it to disk
                rec = json.loads(line)
                self._ids.append(rec["id"])
                self._texts.append(rec["text"])
                metas.append(rec["metadata"])
                live.append(bool(rec.get("live", True)))  # format 1 kept the tombstone in the record
        if len(self._ids) != n:
            raise RuntimeError(f"{d}: {n} stored rows but {len(self._ids)} documents")
        self._cols.extend(metas)
        self._live = np.asarray(live, dtype=bool)
        dead_path = os.path.join(d, _DEAD_FILE)
        if os.path.exists(dead_path):
            with open(dead_path, "r", encoding="utf-8") as f:
                dead = [r for r in json.load(f) if r < n]
            self._live[dead] = False
            self._live_key = None
            self._persisted_dead = dead
        self._id2row = {i: r for r, i in enumerate(self._ids) if self._live[r]}
        # appends continue from here unless the directory needs a clean rewrite (older format, surplus document lines)
        self._persisted_rows = n if manifest.get("format", 1) >= 2 and lines == n else 0
