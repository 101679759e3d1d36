"""Chroma's `hnsw:space` on NON-unit vectors.  langchain_chroma hands collection_metadata to
chromadb, whose default space is "l2" (the reference passes none: /root/reference/src/medical_engine.py:52); on the unit
vectors Ollama returns all three spaces order rows identically, on raw vectors they do not.  CPU part: the oracle's
restatement of the l2 / ip ranking (oracle/canonical.c:orc_rank_score) against a float64 brute force over the raw
vectors, and the host logic of MediQueryVectorStore (space selection, distances, persistence of the row norms) through
the OracleIndex test double.  The GPU kernels are compared with the same oracle in tests/test_gpu_parity.py."""
import os

import numpy as np
import pytest

import oracle as O
from conftest import HashEmbedder, OracleIndex
from mediquery_rag_b200 import Document, MediQueryVectorStore

SEED, QSEED = 0x5EED0006, 0x0BADC0DE


def raw_vectors(n, dim=768, lo=0.2, hi=3.0, seed=3):
    x = O.synth_rows(SEED, 0, n, dim)
    scale = np.random.default_rng(seed).uniform(lo, hi, (n, 1)).astype(np.float32)
    return (x / np.linalg.norm(x, axis=1, keepdims=True) * scale).astype(np.float32)


def f64_distances(space, q, x):
    Q, X = q.astype(np.float64), x.astype(np.float64)
    if space == "l2":
        return (Q * Q).sum(1)[:, None] - 2.0 * Q @ X.T + (X * X).sum(1)[None, :]
    if space == "ip":
        return 1.0 - Q @ X.T
    return 1.0 - (Q / np.linalg.norm(Q, axis=1, keepdims=True)) @ (X / np.linalg.norm(X, axis=1, keepdims=True)).T


@pytest.mark.parametrize("store_dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("space", ["l2", "ip"])
def test_oracle_space_ranking_against_float64_over_the_raw_vectors(space, store_dtype):
    n, B, k = 6000, 9, 10
    x = raw_vectors(n)
    q = raw_vectors(B, lo=0.5, hi=2.0, seed=4) + 0.3 * x[:B]
    st, _ = O.normalize_rows(x, store_dtype, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    nr, nq = O.row_norms(x, 768), O.row_norms(q, 768)
    assert np.allclose(nr, np.linalg.norm(x.astype(np.float64), axis=1), rtol=1e-6)
    s, i = O.brute_force_topk(qn, st, k, metric=space, norms=nr, q_norms=nq)
    # referee in float64: inner products with the vectors AS STORED (unit row in the store's dtype x its fp32 norm), the
    # squared length of a row is the one measured on the raw row (a bf16 store quantises directions, not lengths)
    xs = O.stored_as_f32(st).astype(np.float64) * nr[:, None].astype(np.float64)
    dots = q.astype(np.float64) @ xs.T
    d = 1.0 - dots if space == "ip" else (nq.astype(np.float64) ** 2)[:, None] - 2.0 * dots + (nr.astype(np.float64) ** 2)[None, :]
    ref = np.argsort(d, axis=1, kind="stable")[:, :k]
    got_d = O.distances_from_rank_scores(space, s, nq)
    ref_d = np.take_along_axis(d, ref, 1)
    # tolerance of north_star (1e-5 relative, on the scale of the terms that are subtracted: ||q||^2 + ||x||^2)
    scale = (nq.astype(np.float64) ** 2)[:, None] + 9.0
    assert np.all(np.abs(got_d - ref_d) <= 1e-5 * scale)
    for b in range(B):  # ids equal except where two neighbours are closer than fp32 can tell apart
        for j in np.flatnonzero(i[b] != ref[b]):
            assert abs(d[b, i[b, j]] - d[b, ref[b, j]]) <= 2e-6 * scale[b, 0]
    # and the space matters: the cosine ranking of the same rows is different
    assert not np.array_equal(O.brute_force_topk(qn, st, k)[1], i)


def test_unit_vectors_rank_identically_in_every_space():
    n, B, k = 3000, 5, 10
    x = raw_vectors(n, lo=1.0, hi=1.0)
    q = raw_vectors(B, lo=1.0, hi=1.0, seed=9)
    st, _ = O.normalize_rows(x, "fp32", 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    nr, nq = O.row_norms(x, 768), O.row_norms(q, 768)
    assert np.abs(nr - 1).max() <= 1e-6
    ci = O.brute_force_topk(qn, st, k)[1]
    for space in ("l2", "ip"):
        assert np.array_equal(O.brute_force_topk(qn, st, k, metric=space, norms=nr, q_norms=nq)[1], ci)


def store_of(space, x, texts, **kw):
    vs = MediQueryVectorStore(collection_metadata={"hnsw:space": space} if space else None, index_factory=OracleIndex, **kw)
    vs.add_embeddings((texts, x))
    return vs


@pytest.mark.parametrize("space", ["l2", "ip", None])
def test_vectorstore_ranks_raw_vectors_in_the_collections_space(space, tmp_path):
    n, k = 500, 6
    x = raw_vectors(n)
    texts = [f"doc {r}" for r in range(n)]
    vs = store_of(space, x, texts, persist_directory=str(tmp_path / "db"))
    eff = space or "l2"  # chromadb's default space
    assert vs.rank_space() == eff
    q = (0.7 * x[17] + 0.2 * x[400]).astype(np.float32)
    d = f64_distances(eff, q[None, :], x)[0]
    want = np.argsort(d, kind="stable")[:k]
    got = vs.similarity_search_by_vector_with_relevance_scores(q.tolist(), k=k)  # (doc, distance) as in langchain_chroma
    assert [doc.page_content for doc, _ in got] == [texts[r] for r in want]
    assert np.allclose([dist for _, dist in got], d[want], rtol=1e-5, atol=1e-5 * float(q @ q + 9))
    assert [doc.page_content for doc in vs.similarity_search_by_vector(q.tolist(), k=k)] == [texts[r] for r in want]
    # get(include=embeddings) hands back the raw vectors
    emb = vs.get(ids=[vs._ids[3], vs._ids[250]], include=["embeddings"])["embeddings"]
    assert np.allclose(emb, x[[3, 250]], rtol=1e-5, atol=1e-6)
    # the norms are persisted with the rows: a reopened store ranks the same way
    again = MediQueryVectorStore(collection_metadata={"hnsw:space": space} if space else None, index_factory=OracleIndex,
                                 persist_directory=str(tmp_path / "db"))
    assert os.path.exists(os.path.join(again.collection_path, "rows.mqvs.norms")) and again.rank_space() == eff
    assert [doc.page_content for doc in again.similarity_search_by_vector(q.tolist(), k=k)] == [texts[r] for r in want]
    # appending keeps working
    again.add_embeddings((["late"], x[:1] * np.float32(0.5)))
    assert len(MediQueryVectorStore(index_factory=OracleIndex, persist_directory=str(tmp_path / "db"),
                                    collection_metadata={"hnsw:space": space} if space else None)._index.get_norms()) == n + 1


def test_unit_collections_stay_on_the_cosine_kernels_and_switch_when_a_raw_row_arrives():
    """The reference's case: Ollama returns unit vectors, so the default l2 space is served by the cosine ranking (every
    kernel path available) and distances are 2 - 2s; the first non-unit row switches the collection to raw l2."""
    emb = HashEmbedder()
    texts = [f"t{r}" for r in range(200)]
    vs = MediQueryVectorStore(embedding_function=emb, index_factory=OracleIndex)
    vs.add_texts(texts)
    assert vs.rank_space() == "cosine" and vs._index.norm_deviation() <= 1e-6
    x = np.asarray(emb.embed_documents(texts), np.float32)
    qv = np.asarray(emb.embed_query("t5 and more"), np.float32)
    d = f64_distances("l2", qv[None, :], x)[0]
    got = vs.similarity_search_with_score("t5 and more", k=4)
    assert [doc.page_content for doc, _ in got] == [texts[r] for r in np.argsort(d, kind="stable")[:4]]
    assert np.allclose([dist for _, dist in got], np.sort(d)[:4], atol=2e-6)
    vs.add_embeddings((["big"], (x[:1] * np.float32(2.0))))
    assert vs.rank_space() == "l2"
    got = vs.similarity_search_by_vector_with_relevance_scores(x[0].tolist(), k=2)
    assert [doc.page_content for doc, _ in got] == ["t0", "big"] and np.allclose([s for _, s in got], [0.0, 1.0], atol=1e-5)
    # in the cosine space the norm never matters: "big" ties with t0 (same direction), lower row first
    vc = MediQueryVectorStore(embedding_function=emb, index_factory=OracleIndex, collection_metadata={"hnsw:space": "cosine"})
    vc.add_texts(texts)
    vc.add_embeddings((["big"], (x[:1] * np.float32(2.0))))
    assert vc.rank_space() == "cosine"
    assert [doc.page_content for doc in vc.similarity_search_by_vector(x[0].tolist(), k=2)] == ["t0", "big"]


def test_directories_written_without_norms_open_with_unit_semantics(tmp_path):
    """A persist_directory of the earlier format (space l2 in the manifest, no rows.mqvs.norms) keeps its behaviour."""
    x = raw_vectors(50)
    db = str(tmp_path / "db")
    vs = MediQueryVectorStore(index_factory=OracleIndex, persist_directory=db)
    vs.add_embeddings(([f"d{r}" for r in range(50)], x))
    os.remove(os.path.join(vs.collection_path, "rows.mqvs.norms"))
    old = MediQueryVectorStore(index_factory=OracleIndex, persist_directory=db)
    assert old.rank_space() == "cosine" and len(old) == 50
    d = f64_distances("cosine", x[7:8], x)[0]
    assert [doc.page_content for doc in old.similarity_search_by_vector(x[7].tolist(), k=3)] == \
        [f"d{r}" for r in np.argsort(d, kind="stable")[:3]]


@pytest.mark.parametrize("space", ["cosine", "l2"])
def test_oracle_agrees_with_scikit_learn_brute_force_neighbours(space):
    """An independent, published implementation of exact k-nearest neighbours (scikit-learn's brute-force NearestNeighbors,
    float64) returns the same rows as the oracle: cosine on the normalised rows, (squared) euclidean on the raw vectors.
    chromadb itself is not installable here; this is the closest third-party referee in the image."""
    from sklearn.neighbors import NearestNeighbors
    n, B, k = 5000, 12, 10
    x = raw_vectors(n)
    q = (raw_vectors(B, lo=0.5, hi=2.0, seed=8) + 0.4 * x[100:100 + B]).astype(np.float32)
    st, _ = O.normalize_rows(x, "fp32", 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    if space == "cosine":
        s, i = O.brute_force_topk(qn, st, k)
        nn = NearestNeighbors(n_neighbors=k, algorithm="brute", metric="cosine").fit(x.astype(np.float64))
        dist, ref = nn.kneighbors(q.astype(np.float64))
        got = 1.0 - s.astype(np.float64)
        tol = 2e-6
    else:
        s, i = O.brute_force_topk(qn, st, k, metric="l2", norms=O.row_norms(x, 768), q_norms=O.row_norms(q, 768))
        nn = NearestNeighbors(n_neighbors=k, algorithm="brute", metric="euclidean").fit(x.astype(np.float64))
        dist, ref = nn.kneighbors(q.astype(np.float64))
        dist = dist ** 2
        got = O.distances_from_rank_scores("l2", s, O.row_norms(q, 768))
        tol = 1e-5 * float((q.astype(np.float64) ** 2).sum(1).max() + 9.0)
    assert np.all(np.abs(got - dist) <= tol)
    assert (i == ref).mean() > 0.97  # same rows, same order: positions may differ only where fp32 cannot separate two neighbours
