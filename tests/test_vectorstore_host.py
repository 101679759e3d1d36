"""Host-side logic of the LangChain surface (document table, upsert/delete, filters, persistence, retriever),
run against the OracleIndex test double so it needs no GPU.  Reads like the calls the reference makes:
Chroma.from_documents(documents=, embedding=, persist_directory=) (src/ingest_medical.py:106-110),
Chroma(persist_directory=, embedding_function=) (src/medical_engine.py:52), .similarity_search(q, k=5)
(src/agents/nodes.py:93)."""
import numpy as np
import pytest

import oracle as O
from conftest import HashEmbedder, OracleIndex
from mediquery_rag_b200 import Document, FilterError, MediQueryVectorStore
from mediquery_rag_b200.filters import MetadataColumns, compile_mask, evaluate, pack_mask


def make_docs(medical_docs, n=None):
    docs = medical_docs["docs"][:n]
    return [Document(page_content=d["page_content"], metadata=d["metadata"]) for d in docs]


def test_reference_call_patterns(tmp_path, medical_docs, embedder):
    db = str(tmp_path / "medical_db")
    docs = make_docs(medical_docs)
    vs = MediQueryVectorStore.from_documents(documents=docs, embedding=embedder, persist_directory=db,
                                             index_factory=OracleIndex)
    assert len(vs) == 154
    import os
    assert os.path.exists(db)  # medical_engine.py:34-37 exits when the directory is missing
    # reopen exactly like medical_engine.py:52
    vs2 = MediQueryVectorStore(persist_directory=db, embedding_function=embedder, index_factory=OracleIndex)
    assert len(vs2) == 154
    q = docs[40].page_content
    got = vs2.similarity_search(q, k=5)
    assert len(got) == 5 and got[0].page_content == q
    assert got[0].metadata.get("title") == docs[40].metadata["title"]
    assert [d.page_content for d in got] == [d.page_content for d in vs.similarity_search(q, k=5)]
    # default k is langchain_chroma's 4; k=3 is medical_engine.py:70's default
    assert len(vs2.similarity_search(q)) == 4 and len(vs2.similarity_search(q, k=3)) == 3
    # re-running ingest appends (fresh uuid4 ids), as with Chroma
    vs3 = MediQueryVectorStore.from_documents(documents=docs[:10], embedding=embedder, persist_directory=db,
                                              index_factory=OracleIndex)
    assert len(vs3) == 164


def test_results_equal_oracle_ranking(medical_docs, embedder):
    docs = make_docs(medical_docs)
    vs = MediQueryVectorStore.from_documents(docs, embedder, index_factory=OracleIndex)
    x = np.asarray(embedder.embed_documents([d.page_content for d in docs]), np.float32)
    st, _ = O.normalize_rows(x)
    qv = np.asarray(embedder.embed_query("高血压应该怎么办"), np.float32)
    qn, _ = O.normalize_rows(qv)
    s, i = O.brute_force_topk(qn, st, 4)
    got = vs.similarity_search_with_score("高血压应该怎么办", k=4)
    assert [d.page_content for d, _ in got] == [docs[r].page_content for r in i[0]]
    # default space is Chroma's "l2": squared L2 on unit vectors = 2 - 2 cos, lower is better
    assert np.allclose([dist for _, dist in got], 2.0 - 2.0 * s[0], atol=1e-6)
    assert got[0][1] <= got[-1][1]
    cos = MediQueryVectorStore(embedding_function=embedder, collection_metadata={"hnsw:space": "cosine"},
                               index_factory=OracleIndex)
    cos.add_documents(docs)
    got_c = cos.similarity_search_with_score("高血压应该怎么办", k=4)
    assert np.allclose([d for _, d in got_c], 1.0 - s[0], atol=1e-6)
    rel = cos.similarity_search_with_relevance_scores("高血压应该怎么办", k=4)
    assert np.allclose([r for _, r in rel], s[0], atol=1e-6)


def test_edge_conventions(embedder):
    vs = MediQueryVectorStore(embedding_function=embedder, index_factory=OracleIndex)
    assert vs.similarity_search("anything", k=5) == []  # empty store -> []
    ids = vs.add_texts(["a", "b", "c"], metadatas=[{"t": 1}, {"t": 2}, None])
    assert len(ids) == 3 and len(set(ids)) == 3
    assert len(vs.similarity_search("a", k=10)) == 3  # k > N clamps
    assert vs.similarity_search("a", k=1)[0].page_content == "a"
    # vectors only, no embedder
    raw = MediQueryVectorStore(index_factory=OracleIndex)
    raw.add_embeddings([("x", [1.0, 0.0, 0.0]), ("y", [0.0, 1.0, 0.0])], metadatas=[{"n": 1}, {"n": 2}])
    assert raw.similarity_search_by_vector([0.9, 0.1, 0.0], k=1)[0].page_content == "x"
    with pytest.raises(ValueError):
        raw.similarity_search("text query without embedder")
    with pytest.raises(ValueError):
        raw.add_embeddings([("z", [1.0, 0.0])])  # wrong width


def test_upsert_delete_get(embedder):
    vs = MediQueryVectorStore(embedding_function=embedder, index_factory=OracleIndex)
    vs.add_texts(["one", "two", "three"], metadatas=[{"k": "a"}, {"k": "b"}, {"k": "a"}], ids=["1", "2", "3"])
    vs.add_texts(["two (edited)"], metadatas=[{"k": "c"}], ids=["2"])  # upsert retires the old row
    assert len(vs) == 3
    assert vs.similarity_search("two", k=3)[0].page_content != "two" or True
    texts = {d.page_content for d in vs.similarity_search("two", k=10)}
    assert texts == {"one", "two (edited)", "three"}
    got = vs.get(where={"k": "a"})
    assert got["ids"] == ["1", "3"] and got["documents"] == ["one", "three"]
    assert vs.get(ids=["2"])["metadatas"] == [{"k": "c"}]
    emb = vs.get(ids=["1"], include=["embeddings"])["embeddings"]
    assert emb.shape == (1, 768) and abs(np.linalg.norm(emb[0]) - 1.0) < 1e-5
    vs.delete(ids=["1"])
    assert len(vs) == 2 and vs.get(ids=["1"])["ids"] == []
    assert {d.page_content for d in vs.similarity_search("one", k=10)} == {"two (edited)", "three"}
    assert [d.id for d in vs.get_by_ids(["3"])] == ["3"]


def test_filter_search_and_compiler_vs_naive_evaluator(embedder):
    rng = np.random.default_rng(0)
    n = 500
    mds = [{"source_id": int(rng.integers(0, 16)), "tags": ["x", "y", "z"][int(rng.integers(0, 3))]} if i % 7 else
           ({"source_id": 3} if i % 2 else None) for i in range(n)]
    texts = [f"doc {i}" for i in range(n)]
    vs = MediQueryVectorStore(embedding_function=embedder, index_factory=OracleIndex)
    vs.add_texts(texts, metadatas=mds)
    filters = [{"source_id": 3}, {"source_id": {"$in": [1, 2, 5]}}, {"tags": {"$ne": "x"}}, {"source_id": {"$gte": 12}},
               {"$and": [{"tags": "y"}, {"source_id": {"$lt": 8}}]}, {"$or": [{"tags": "z"}, {"source_id": {"$nin": [0, 1]}}]},
               {"missing": 1}]
    x = np.asarray(embedder.embed_documents(texts), np.float32)
    st, _ = O.normalize_rows(x)
    qn, _ = O.normalize_rows(np.asarray(embedder.embed_query("doc 77"), np.float32))
    for w in filters:
        naive = O.eval_where(w, [m or {} for m in mds])
        assert np.array_equal(evaluate(w, vs._cols), naive), w
        got = vs.similarity_search("doc 77", k=10, filter=w)
        _, want = O.brute_force_topk(qn, st, 10, naive)
        want = [r for r in want[0] if r >= 0]
        assert [d.page_content for d in got] == [texts[r] for r in want], w
    with pytest.raises(FilterError):
        vs.similarity_search("q", filter={"a": 1, "b": 2})
    with pytest.raises(FilterError):
        vs.similarity_search("q", filter={"a": {"$regex": "x"}})


def test_bulk_metadata_columns_and_mask_packing():
    cols = MetadataColumns()
    cols.extend_columns(100, {"source_id": np.arange(100) % 16})
    ok = evaluate({"source_id": {"$in": [0, 5]}}, cols)
    assert ok.sum() == 13
    mask, allowed = compile_mask({"source_id": {"$in": [0, 5]}}, cols, np.ones(100, bool))
    assert allowed == 13 and mask.dtype == np.uint32 and mask.shape == (4,)
    assert np.array_equal(mask, O.mask_from_bool(ok))
    assert compile_mask(None, cols, np.ones(100, bool)) == (None, 100)
    live = np.ones(100, bool)
    live[0] = False
    m2, a2 = compile_mask(None, cols, live)
    assert a2 == 99 and not (m2[0] & 1)
    assert pack_mask(np.zeros(0, bool)).shape == (0,)


def test_retriever_modes(embedder):
    vs = MediQueryVectorStore(embedding_function=embedder, collection_metadata={"hnsw:space": "cosine"},
                              index_factory=OracleIndex)
    vs.add_texts([f"text {i}" for i in range(30)], metadatas=[{"g": i % 3} for i in range(30)])
    r = vs.as_retriever(search_kwargs={"k": 5})
    assert [d.page_content for d in r.invoke("text 3")] == [d.page_content for d in vs.similarity_search("text 3", k=5)]
    r = vs.as_retriever(search_kwargs={"k": 5, "filter": {"g": 0}})
    assert all(d.metadata["g"] == 0 for d in r.invoke("text 3"))
    r = vs.as_retriever(search_type="similarity_score_threshold", search_kwargs={"k": 5, "score_threshold": 0.99})
    assert [d.page_content for d in r.invoke("text 3")] == ["text 3"]
    mmr = vs.as_retriever(search_type="mmr", search_kwargs={"k": 4, "fetch_k": 10}).invoke("text 3")
    assert len(mmr) == 4 and mmr[0].page_content == "text 3" and len({d.page_content for d in mmr}) == 4
    with pytest.raises(ValueError):
        vs.as_retriever(search_type="bogus")


def test_non_finite_embeddings_are_rejected_and_nothing_is_added():
    vs = MediQueryVectorStore(index_factory=OracleIndex)
    vs.add_embeddings([("ok", [1.0, 2.0])])
    with pytest.raises(RuntimeError):
        vs.add_embeddings([("bad", [float("nan"), 1.0])])
    assert len(vs) == 1


def test_where_document_contains(embedder):
    vs = MediQueryVectorStore(embedding_function=embedder, index_factory=OracleIndex)
    vs.add_texts(["运动 降低 血压", "睡眠 与 记忆", "运动 与 睡眠", "饮食"], metadatas=[{"g": 1}, {"g": 1}, {"g": 2}, {"g": 2}])
    got = vs.similarity_search("q", k=10, where_document={"$contains": "运动"})
    assert {d.page_content for d in got} == {"运动 降低 血压", "运动 与 睡眠"}
    got = vs.similarity_search_with_score("q", k=10, filter={"g": 2}, where_document={"$not_contains": "运动"})
    assert [d.page_content for d, _ in got] == ["饮食"]
    got = vs.similarity_search("q", k=10, where_document={"$or": [{"$contains": "饮食"}, {"$and": [{"$contains": "睡眠"}, {"$contains": "记忆"}]}]})
    assert {d.page_content for d in got} == {"饮食", "睡眠 与 记忆"}
    with pytest.raises(FilterError):
        vs.similarity_search("q", where_document={"$regex": "x"})


def test_update_documents_and_async_wrappers(medical_docs, embedder):
    import asyncio
    docs = make_docs(medical_docs, 30)
    vs = MediQueryVectorStore(embedding_function=embedder, index_factory=OracleIndex)
    ids = vs.add_documents(docs)
    new = Document(page_content="问题：更新后的文档\n答案：全新的内容", metadata={"title": "更新", "tags": "x", "source": "s"})
    vs.update_document(ids[3], new)
    assert len(vs) == 30                                  # the old row is tombstoned, not duplicated
    got = vs.similarity_search(new.page_content, k=1)
    assert got[0].page_content == new.page_content and got[0].id == ids[3]
    assert docs[3].page_content not in [d.page_content for d in vs.similarity_search(docs[3].page_content, k=30)]
    with pytest.raises(ValueError):
        vs.update_document("no-such-id", new)

    async def go():
        a = await vs.asimilarity_search(docs[7].page_content, k=3)
        b = await vs.asimilarity_search_with_score(docs[7].page_content, k=3)
        c = await vs.asimilarity_search_by_vector(embedder.embed_query(docs[7].page_content), k=3)
        more = await vs.aadd_documents(docs[:2])
        return a, b, c, more

    a, b, c, more = asyncio.run(go())
    sync = vs.similarity_search(docs[7].page_content, k=3)
    assert a[0].page_content == docs[7].page_content and [d.page_content for d, _ in b] == [d.page_content for d in a]
    assert [d.page_content for d in c] == [d.page_content for d in a] and len(more) == 2 and len(vs) == 32
    assert [d.page_content for d in sync][0] == docs[7].page_content


def test_persistence_is_append_only_and_survives_reopen_delete_and_old_format(tmp_path, medical_docs, embedder):
    import json, os
    db = str(tmp_path / "db")
    docs = make_docs(medical_docs, 40)
    vs = MediQueryVectorStore.from_documents(docs[:20], embedder, persist_directory=db, index_factory=OracleIndex)
    cdir = vs.collection_path  # every collection owns a sub-directory of the persist_directory
    assert os.path.dirname(cdir) == db
    size0 = os.path.getsize(os.path.join(cdir, "docs.jsonl"))
    first_line = open(os.path.join(cdir, "docs.jsonl"), encoding="utf-8").readline()
    ids2 = vs.add_documents(docs[20:30])                      # appends 10 records, rewrites nothing
    assert os.path.getsize(os.path.join(cdir, "docs.jsonl")) > size0
    assert open(os.path.join(cdir, "docs.jsonl"), encoding="utf-8").readline() == first_line
    assert sum(1 for _ in open(os.path.join(cdir, "docs.jsonl"), encoding="utf-8")) == 30
    vs.delete(ids2[:3])
    assert json.load(open(os.path.join(cdir, "dead.json"))) == [20, 21, 22]
    assert json.load(open(os.path.join(cdir, "manifest.json")))["format"] == 2
    # reopen (src/medical_engine.py:52), keep appending (src/ingest_medical.py:106 re-run), reopen again
    vs2 = MediQueryVectorStore(persist_directory=db, embedding_function=embedder, index_factory=OracleIndex)
    assert len(vs2) == 27 and vs2.similarity_search(docs[25].page_content, k=1)[0].page_content == docs[25].page_content
    assert docs[21].page_content not in [d.page_content for d in vs2.similarity_search(docs[21].page_content, k=27)]
    vs2.add_documents(docs[30:40])
    vs3 = MediQueryVectorStore(persist_directory=db, embedding_function=embedder, index_factory=OracleIndex)
    assert len(vs3) == 37 and [d.page_content for d in vs3.similarity_search(docs[35].page_content, k=3)] == \
        [d.page_content for d in vs2.similarity_search(docs[35].page_content, k=3)]
    # surplus document lines (a crash after the documents of an append were written but before its rows): ignored, and
    # the next persist rewrites the directory cleanly
    with open(os.path.join(cdir, "docs.jsonl"), "a", encoding="utf-8") as f:
        f.write(json.dumps({"id": "ghost", "text": "never stored", "metadata": {}}) + "\n")
    vs4 = MediQueryVectorStore(persist_directory=db, embedding_function=embedder, index_factory=OracleIndex)
    assert len(vs4) == 37 and "ghost" not in vs4.get()["ids"]
    vs4.add_documents(docs[:1])
    assert sum(1 for _ in open(os.path.join(cdir, "docs.jsonl"), encoding="utf-8")) == 41
    assert len(MediQueryVectorStore(persist_directory=db, embedding_function=embedder, index_factory=OracleIndex)) == 38
    # a format-1 directory (tombstones inside the records) still opens
    old = str(tmp_path / "old")
    os.makedirs(old)
    v = MediQueryVectorStore.from_documents(docs[:5], embedder, persist_directory=str(tmp_path / "scratch"), index_factory=OracleIndex)
    v._index.save(os.path.join(old, "rows.mqvs"))
    with open(os.path.join(old, "docs.jsonl"), "w", encoding="utf-8") as f:
        for r in range(5):
            f.write(json.dumps({"id": v._ids[r], "text": v._texts[r], "metadata": v._cols.row(r), "live": r != 2}, ensure_ascii=False) + "\n")
    json.dump({"format": 1, "collection_name": "langchain", "dim": v._dim, "store_dtype": "fp32", "rows": 5, "space": "l2",
               "collection_metadata": {}}, open(os.path.join(old, "manifest.json"), "w"))
    vo = MediQueryVectorStore(persist_directory=old, embedding_function=embedder, index_factory=OracleIndex)
    assert len(vo) == 4 and docs[2].page_content not in [d.page_content for d in vo.similarity_search(docs[2].page_content, k=4)]


def test_similarity_search_with_vectors_and_reset_collection(embedder, tmp_path):
    """Two more calls of langchain_chroma's surface: (Document, embedding) pairs, and reset_collection (same name, space
    and storage, no rows -- also on disk)."""
    db = str(tmp_path / "db")
    vs = MediQueryVectorStore(embedding_function=embedder, index_factory=OracleIndex, persist_directory=db)
    texts = [f"text {r}" for r in range(40)]
    vs.add_texts(texts, ids=[str(r) for r in range(40)])
    pairs = vs.similarity_search_with_vectors("text 7", k=3)
    assert [d.page_content for d, _ in pairs] == [d.page_content for d in vs.similarity_search("text 7", k=3)]
    want = np.asarray(embedder.embed_documents([d.page_content for d, _ in pairs]), np.float32)
    assert np.allclose(np.stack([v for _, v in pairs]), want, atol=1e-6)
    vs.reset_collection()
    assert len(vs) == 0 and vs.similarity_search("text 7", k=3) == [] and vs.get()["ids"] == []
    assert len(MediQueryVectorStore(embedding_function=embedder, index_factory=OracleIndex, persist_directory=db)) == 0
    vs.add_texts(["again"], ids=["0"])
    assert [d.page_content for d in vs.similarity_search("again", k=2)] == ["again"]
    assert len(MediQueryVectorStore(embedding_function=embedder, index_factory=OracleIndex, persist_directory=db)) == 1


def test_collections_sharing_a_persist_directory_do_not_touch_each_other(tmp_path, medical_docs, embedder):
    """Chroma keeps several collections in one persist_directory; opening, persisting or deleting one must leave the
    others alone (every collection owns `mqvs.<name>/`)."""
    import os
    db = str(tmp_path / "db")
    docs = make_docs(medical_docs, 12)
    a = MediQueryVectorStore.from_documents(docs[:8], embedder, collection_name="alpha", persist_directory=db, index_factory=OracleIndex)
    b = MediQueryVectorStore(collection_name="beta", persist_directory=db, embedding_function=embedder, index_factory=OracleIndex)
    assert len(b) == 0 and a.collection_path != b.collection_path
    b.add_documents(docs[8:12])
    assert len(MediQueryVectorStore(collection_name="alpha", persist_directory=db, embedding_function=embedder,
                                    index_factory=OracleIndex)) == 8
    b.delete_collection()
    assert not os.path.exists(os.path.join(b.collection_path, "manifest.json"))
    again = MediQueryVectorStore(collection_name="alpha", persist_directory=db, embedding_function=embedder, index_factory=OracleIndex)
    assert len(again) == 8 and again.similarity_search(docs[3].page_content, k=1)[0].page_content == docs[3].page_content
    # a directory of the older layout (files at the top level) belongs to the collection named in its manifest only
    old = str(tmp_path / "old")
    legacy = MediQueryVectorStore.from_documents(docs[:5], embedder, persist_directory=old, index_factory=OracleIndex)
    for f in os.listdir(legacy.collection_path):
        os.replace(os.path.join(legacy.collection_path, f), os.path.join(old, f))
    os.rmdir(legacy.collection_path)
    assert len(MediQueryVectorStore(persist_directory=old, embedding_function=embedder, index_factory=OracleIndex)) == 5
    other = MediQueryVectorStore(collection_name="other", persist_directory=old, embedding_function=embedder, index_factory=OracleIndex)
    other.add_documents(docs[5:7])
    assert len(other) == 2
    assert len(MediQueryVectorStore(persist_directory=old, embedding_function=embedder, index_factory=OracleIndex)) == 5


def test_a_directory_written_by_real_chroma_is_reported_not_silently_empty(tmp_path, medical_docs, embedder):
    import warnings
    db = tmp_path / "medical_db"
    db.mkdir()
    (db / "chroma.sqlite3").write_bytes(b"SQLite format 3\0")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        vs = MediQueryVectorStore(persist_directory=str(db), embedding_function=embedder, index_factory=OracleIndex)
    assert any("re-run src/ingest_medical.py" in str(x.message) for x in w)
    with pytest.raises(RuntimeError, match="chromadb"):
        vs.similarity_search("高血压", k=5)
    docs = make_docs(medical_docs, 6)
    vs.add_documents(docs)  # re-running the ingest (src/ingest_medical.py:106) fills the collection: searches work again
    assert vs.similarity_search(docs[2].page_content, k=1)[0].page_content == docs[2].page_content


def test_bulk_rows_with_implicit_ids_and_tombstone_compaction(tmp_path, embedder):
    vs = MediQueryVectorStore(index_factory=OracleIndex, dim=768, collection_metadata={"hnsw:space": "cosine"})
    rows = vs.add_synthetic(0x5EED0002, 5000, metadata_columns={"g": np.arange(5000) % 3})
    assert rows == range(0, 5000) and len(vs) == 5000 and len(vs._ids._runs) == 1
    q, planted = O.planted_queries(0x0BADC0DE, 2, 0x5EED0002, 5000, 768)
    doc = vs.similarity_search_by_vector(q[0], k=1)[0]
    assert doc.id == f"row-{planted[0]}" and doc.metadata == {"g": int(planted[0] % 3)} and doc.page_content == ""
    assert vs.get(ids=[f"row-{planted[0]}", "row-99999", "nope"])["ids"] == [f"row-{planted[0]}"]
    vs.delete([f"row-{planted[0]}"])
    assert vs.similarity_search_by_vector(q[0], k=1)[0].id != f"row-{planted[0]}" and len(vs) == 4999
    more = vs.add_embeddings(([f"t{i}" for i in range(3)], O.synth_rows(5, 0, 3, 768)), ids=["x", "y", "z"])
    assert more == ["x", "y", "z"] and vs.get(ids=["y"])["documents"] == ["t1"] and len(vs) == 5002
    # compaction: rows renumbered, string ids stay, results unchanged
    before = [d.id for d in vs.similarity_search_by_vector(q[1], k=5)]
    vs.delete([f"row-{r}" for r in range(0, 5000, 2)])
    kept = [d.id for d in vs.similarity_search_by_vector(q[1], k=5)]
    dead = 2500 + int(planted[0] % 2 == 1)
    assert vs.compact() == dead and len(vs._index) == len(vs) == 5003 - dead
    assert [d.id for d in vs.similarity_search_by_vector(q[1], k=5)] == kept
    assert [i for i in before if int(i.split("-")[1]) % 2 == 1] == kept[:sum(int(i.split("-")[1]) % 2 == 1 for i in before)]
    assert vs.get(ids=["z"])["documents"] == ["t2"] and vs.get(ids=["row-0"])["ids"] == []
    sc, rid = vs.search_batch(q, 3, filter={"g": 1})
    assert all(vs._cols.row(int(r))["g"] == 1 for r in rid.ravel())


def test_a_bad_metadata_column_is_refused_before_any_row_reaches_the_index():
    """Rows and host tables move together: a column of the wrong length must fail BEFORE the device add, or every later add
    would find the document table and the device rows out of step."""
    vs = MediQueryVectorStore(index_factory=OracleIndex, dim=768, collection_metadata={"hnsw:space": "cosine"})
    x = O.synth_rows(7, 0, 10, 768)
    with pytest.raises(ValueError):
        vs.add_embeddings((None, x), metadata_columns={"g": np.arange(9)})
    with pytest.raises(ValueError):
        vs.add_embeddings((None, x), metadatas=[{}] * 10, metadata_columns={"g": np.arange(10)})
    with pytest.raises(ValueError):
        vs.add_synthetic(1, 10, metadata_columns={"g": np.arange(11)})
    assert len(vs) == 0 and (vs._index is None or len(vs._index) == 0)
    ids = vs.add_embeddings((None, x), metadata_columns={"g": np.arange(10)})
    assert len(ids) == 10 and len(vs) == 10
    assert vs.similarity_search_by_vector(x[3], k=1, filter={"g": 3})[0].id == ids[3]


def test_get_honours_where_document(medical_docs, embedder):
    """collection.get(where=..., where_document={"$contains": ...}) as chromadb offers it: both conditions, live rows only."""
    docs = make_docs(medical_docs, 60)
    vs = MediQueryVectorStore.from_documents(docs, embedder, index_factory=OracleIndex)
    needle = "血压"
    want = [d.page_content for d in docs if needle in d.page_content]
    assert 0 < len(want) < 60
    got = vs.get(where_document={"$contains": needle})
    assert got["documents"] == want
    assert vs.get(where_document={"$not_contains": needle})["documents"] == [d.page_content for d in docs if needle not in d.page_content]
    title = docs[[needle in d.page_content for d in docs].index(True)].metadata["title"]
    both = vs.get(where={"title": title}, where_document={"$contains": needle})
    assert both["documents"] and all(needle in t for t in both["documents"]) and all(m["title"] == title for m in both["metadatas"])
    some_ids = vs.get()["ids"][:10]
    by_id = vs.get(ids=some_ids, where_document={"$contains": needle})
    assert by_id["documents"] == [d.page_content for d in docs[:10] if needle in d.page_content]
    vs.delete(got["ids"][:1])
    assert vs.get(where_document={"$contains": needle})["documents"] == want[1:]


def test_delete_by_where_and_where_document(medical_docs, embedder):
    docs = make_docs(medical_docs, 60)
    vs = MediQueryVectorStore.from_documents(docs, embedder, index_factory=OracleIndex)
    needle = "血压"
    n_hit = sum(needle in d.page_content for d in docs)
    assert vs.delete(where_document={"$contains": needle}) is True and len(vs) == 60 - n_hit
    assert all(needle not in d.page_content for d in vs.similarity_search(needle, k=10))
    ids = vs.get()["ids"]
    source = vs.get()["metadatas"][0]["source"]                       # every chunk carries the same source (ingest_medical.py:78)
    before = len(vs)
    assert vs.delete(ids=ids[5:8], where={"source": source}) is True  # ids AND where: only those three go
    assert len(vs) == before - 3 and vs.get(ids=ids[5:8])["ids"] == [] and vs.get(ids=ids[:1])["ids"] == ids[:1]
    assert vs.delete(ids=ids[:2], where={"source": "nothing has this"}) is True and len(vs) == before - 3
    title = vs.get()["metadatas"][0]["title"]
    vs.delete(where={"title": title})
    assert vs.get(where={"title": title})["ids"] == [] and len(vs) < before - 3
    assert vs.delete() is False


def test_collection_view_counts_and_queries_like_chromadb(medical_docs, embedder):
    docs = make_docs(medical_docs, 80)
    vs = MediQueryVectorStore.from_documents(docs, embedder, index_factory=OracleIndex)
    col = vs._collection
    assert col.count() == 80 and col.name == "langchain" and col.metadata == {}
    qs = [docs[3].page_content, docs[50].page_content]
    res = col.query(query_texts=qs, n_results=4)
    assert [len(x) for x in res["ids"]] == [4, 4] and res["documents"][0][0] == qs[0] and res["documents"][1][0] == qs[1]
    for j, qtext in enumerate(qs):  # same rows and distances as the LangChain call, query by query
        ref = vs.similarity_search_with_score(qtext, k=4)
        assert res["documents"][j] == [d.page_content for d, _ in ref]
        assert np.allclose(res["distances"][j], [dist for _, dist in ref], atol=1e-6)
        assert res["metadatas"][j] == [d.metadata for d, _ in ref]
    emb = np.asarray(embedder.embed_documents(qs), np.float32)
    by_vec = col.query(query_embeddings=emb, n_results=4, include=["distances"])
    assert by_vec["ids"] == res["ids"] and by_vec["documents"] is None
    needle = "血压"
    filt = col.query(query_embeddings=emb[0], n_results=100, where_document={"$contains": needle})
    assert len(filt["ids"]) == 1 and 0 < len(filt["ids"][0]) == sum(needle in d.page_content for d in docs)
    col.delete(ids=res["ids"][0][:1])
    assert col.count() == 79 and col.get(ids=res["ids"][0][:1])["ids"] == []


def test_a_metadata_that_is_not_a_mapping_is_refused_before_any_row_reaches_the_index(embedder):
    vs = MediQueryVectorStore(embedding_function=embedder, index_factory=OracleIndex)
    with pytest.raises(TypeError):
        vs.add_texts(["a", "b"], metadatas=[{"g": 1}, "not a dict"])
    assert len(vs) == 0 and (vs._index is None or len(vs._index) == 0)
    assert vs.add_texts(["a", "b"], metadatas=[{"g": 1}, None], ids=["x", "y"]) == ["x", "y"]
    assert vs.get(ids=["x"])["metadatas"] == [{"g": 1}]


def test_a_compaction_cannot_slip_between_a_search_and_its_documents(embedder):
    """A compaction renumbers rows; a search maps its row numbers to documents under the same lock as the scan.  Here another
    thread deletes + compacts right after the scan returned its rows: it has to wait until the documents are collected."""
    import threading
    texts = [f"chunk {i} " + "x" * (i % 7) for i in range(1500)]
    vs = MediQueryVectorStore(embedding_function=embedder, index_factory=OracleIndex, compact_dead_fraction=0.0)
    ids = vs.add_texts(texts, ids=[f"id-{i}" for i in range(len(texts))])
    real, started = vs.search_batch, []

    def search_then_compaction_elsewhere(*a, **kw):
        out = real(*a, **kw)
        if not started:
            t = threading.Thread(target=lambda: (vs.delete(ids[:1000]), vs.compact()))
            started.append(t)
            t.start()
            t.join(0.3)  # finishes here if nothing holds it back -> the rows in `out` would then name other documents
        return out

    vs.search_batch = search_then_compaction_elsewhere
    got = vs.similarity_search(texts[1200], k=1)
    started[0].join()
    assert got[0].page_content == texts[1200] and got[0].id == "id-1200"
    assert len(vs) == 500 and vs.similarity_search(texts[1200], k=1)[0].page_content == texts[1200]
    vs.search_batch, started[:] = real, []
    assert vs.get(ids=["id-1200"])["documents"] == [texts[1200]]
