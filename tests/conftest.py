"""Shared fixtures.  `-m "not gpu"` runs here (CPU, no GPU); `-m gpu` runs on a B200 and calls through the C ABI.

OracleIndex is a TEST DOUBLE with the `_native.Index` interface, backed by oracle/ -- it lets the host-side
logic (document table, filters, persistence, sharding protocol) be exercised without a GPU.  It lives in
tests/ on purpose: the product has no CPU path and fails loudly without libmqvs.so + a B200.
"""
import hashlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` under gpurun)")


def have_gpu() -> bool:
    try:
        from mediquery_rag_b200 import _native
        return _native.device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def gpu():
    if not have_gpu():
        pytest.fail("this test is marked gpu but no B200 / libmqvs.so is usable: the product has no CPU fallback")
    return True


class OracleIndex:
    """Test double for _native.Index computing with the CPU oracle."""

    def __init__(self, dim, store_dtype="fp32", device=0, capacity_rows=0):
        self.dim, self.store_dtype = dim, store_dtype
        self.rows = np.zeros((0, 768), np.float32 if store_dtype == "fp32" else np.uint16)
        self.metric, self.norms, self._dev = "cosine", None, 0.0

    def __len__(self):
        return self.rows.shape[0]

    def add(self, x):
        x = np.ascontiguousarray(x, np.float32)
        st, bad = O.normalize_rows(x, self.store_dtype, 768)
        if bad:
            raise RuntimeError("non-finite")
        first = len(self)
        self.rows = np.concatenate([self.rows, st])
        nr = O.row_norms(x, 768)
        self._dev = max(self._dev, float(np.abs(nr - np.float32(1.0)).max())) if len(nr) else self._dev
        if self.norms is not None:
            self.norms = np.concatenate([self.norms, nr])
        return first

    def search(self, q, k, mask=None):
        q = np.ascontiguousarray(q, np.float32)
        qn, _ = O.normalize_rows(q, "fp32", 768)
        if self.metric != "cosine":
            return O.brute_force_topk(qn, self.rows, k, mask, metric=self.metric, norms=self.norms,
                                      q_norms=O.row_norms(q, 768))
        return O.brute_force_topk(qn, self.rows, k, mask)

    def set_metric(self, space):
        if space != "cosine" and self.norms is None:
            if len(self):
                raise RuntimeError("rows were added without their norms")
            self.norms = np.zeros(0, np.float32)
        self.metric = space

    def norm_deviation(self):
        return self._dev

    def get_norms(self, first_row=0, n=None):
        n = len(self) - first_row if n is None else n
        return self.norms[first_row:first_row + n].copy()

    def add_synthetic(self, seed, first_index, n):
        return self.add(O.synth_rows(seed, first_index, n, self.dim))

    def search_device(self, q, k, out_scores, out_ids, mask=None, id_base=0, stream=0):
        """CPU-tensor version of the all-device call (gloo tests)."""
        import torch
        s, i = self.search(q.numpy(), k, mask)
        i = np.where(i >= 0, i + id_base, -1)
        out_scores.copy_(torch.from_numpy(s))
        out_ids.copy_(torch.from_numpy(i))

    def get_rows(self, ids):
        return O.stored_as_f32(self.rows)[np.asarray(ids, np.int64), : self.dim]

    def save_append(self, path, first_row):
        assert first_row <= len(self)
        self.save(path)

    def save(self, path):
        np.save(path + ".npy", self.rows)
        os.replace(path + ".npy", path)
        if self.norms is not None:
            np.save(path + ".norms.npy", self.norms)
            os.replace(path + ".norms.npy", path + ".norms")

    def load_file(self, path):
        with open(path, "rb") as f:
            self.rows = np.load(f)
        if self.norms is not None:
            with open(path + ".norms", "rb") as f:
                self.norms = np.load(f)[: len(self)]
            self._dev = float(np.abs(self.norms - np.float32(1.0)).max()) if len(self.norms) else 0.0

    def compact(self, keep_words):
        n = len(self)
        keep = ((np.asarray(keep_words, np.uint32)[np.arange(n) >> 5] >> (np.arange(n) & 31).astype(np.uint32)) & 1).astype(bool)
        self.rows = self.rows[keep]
        if self.norms is not None:
            self.norms = self.norms[keep]
        return len(self)

    def close(self):
        pass


class HashEmbedder:
    """Deterministic stand-in for OllamaEmbeddings (no daemon here): text -> 768-d UNIT vector from its sha256
    (Ollama's /api/embed returns L2-normalised rows; `unit=False` gives raw vectors of norm ~16 instead)."""

    def __init__(self, dim=768, unit=True):
        self.dim, self.unit = dim, unit

    def _one(self, text):
        seed = int.from_bytes(hashlib.sha256(text.encode("utf-8")).digest()[:8], "little")
        v = O.synth_rows(seed, 0, 1, self.dim)[0]
        return (v / np.float32(np.linalg.norm(v.astype(np.float64)))).astype(np.float32) if self.unit else v

    def embed_documents(self, texts):
        return [self._one(t).tolist() for t in texts]

    def embed_query(self, text):
        return self._one(text).tolist()


@pytest.fixture
def embedder():
    return HashEmbedder()


@pytest.fixture(scope="session")
def medical_docs():
    import gzip
    import json
    with gzip.open(os.path.join(GOLDEN, "medical_docs.json.gz"), "rt", encoding="utf-8") as f:
        return json.load(f)
