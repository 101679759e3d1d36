"""GPU parity (run with `-m gpu` on a B200): the CUDA path, called through the C ABI (ctypes -> libmqvs.so),
against the CPU oracle on the same seeded inputs.  Bar: ids identical and score BITS identical to the canonical
fp32 oracle (which tests/test_oracle.py ties to the fp64 referee within north_star's 1e-5 relative)."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

SEED = 0x5EED0002
QSEED = 0x0BADC0DE


def native():
    import mediquery_rag_b200 as M
    return M


def assert_bit_equal(s, i, es, ei):
    assert np.array_equal(i, ei), f"ids differ at {np.argwhere(i != ei)[:4].tolist()}"
    assert np.array_equal(s.view(np.uint32), es.view(np.uint32)), "score bits differ"


def build(n, dim, dtype, seed=SEED):
    x = O.synth_rows(seed, 0, n, dim)
    idx = native().Index(dim, dtype)
    idx.add(x)
    st, _ = O.normalize_rows(x, dtype, 768)
    return idx, x, st


@pytest.mark.parametrize("dim", [768, 384, 100, 3])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_k1_normalise_pack_is_bit_exact(gpu, dim, dtype):
    n = 3000
    x = O.synth_rows(SEED, 0, n, dim) * np.float32(3.7)
    x[5] = 0.0  # zero row stays zero
    idx = native().Index(dim, dtype)
    assert idx.add(x) == 0 and len(idx) == n
    got = idx.get_rows(np.arange(n))
    st, bad = O.normalize_rows(x, dtype, 768)
    assert bad == 0
    assert np.array_equal(got.view(np.uint32), O.stored_as_f32(st)[:, :dim].view(np.uint32))
    assert not got[5].any()


def test_device_side_generator_matches_oracle_generator(gpu):
    idx = native().Index(768, "fp32")
    idx.add_synthetic(SEED, 1000, 500)
    st, _ = O.normalize_rows(O.synth_rows(SEED, 1000, 500, 768), "fp32", 768)
    assert np.array_equal(idx.get_rows(np.arange(500)), st)


CASES = [
    # n, dim, dtype, B, k, mask fraction
    (154, 768, "fp32", 1, 4, None),      # config 1: Chroma default k
    (154, 768, "fp32", 1, 5, None),      # k actually used at src/agents/nodes.py:93
    (154, 768, "bf16", 3, 3, None),      # k default of src/medical_engine.py:64
    (1000, 768, "bf16", 5, 10, None),
    (20000, 768, "fp32", 9, 10, None),
    (20000, 768, "bf16", 8, 10, 0.3),
    (20000, 768, "fp32", 64, 10, None),
    (5000, 384, "fp32", 2, 100, None),
    (3000, 100, "bf16", 3, 100, 0.01),
    (40000, 768, "fp32", 2, 1000, None),
    (33, 768, "fp32", 7, 32, 0.5),
    (100000, 768, "bf16", 4, 10, 0.0001),
    (70001, 768, "fp32", 4, 33, 0.9),
]


@pytest.mark.parametrize("n,dim,dtype,B,k,frac", CASES)
def test_search_matches_oracle_bitwise(gpu, n, dim, dtype, B, k, frac):
    idx, x, st = build(n, dim, dtype)
    q, planted = O.planted_queries(QSEED, B, SEED, n, dim)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    mask = None
    if frac is not None:
        mask = O.mask_from_bool(np.random.default_rng(7).random(n) < frac)
    s, i = idx.search(q, k, mask)
    es, ei = O.brute_force_topk(qn, st, k, mask)
    assert_bit_equal(s, i, es, ei)
    if mask is None and k <= n:
        assert all(i[b, 0] == planted[b] for b in range(0, B, 2))
    idx.close()


def test_edge_cases(gpu):
    M = native()
    idx = M.Index(768, "fp32")
    q = O.synth_rows(QSEED, 0, 2, 768)
    s, i = idx.search(q, 5)  # empty store
    assert (i == -1).all() and np.isneginf(s).all()
    base = O.synth_rows(SEED, 0, 64, 768)
    x = np.concatenate([base, base[:8], base[:8]])  # exact duplicates -> ties broken to the lower id
    idx.add(x)
    s, i = idx.search(base[3:4], 5)
    assert i[0, :3].tolist() == [3, 67, 75] and s[0, 0] == s[0, 1] == s[0, 2]
    s, i = idx.search(q, 200)  # k > N: padded with (-inf, -1)
    assert (i[:, :80] >= 0).all() and (i[:, 80:] == -1).all() and np.isneginf(s[:, 80:]).all()
    allowed = np.zeros(80, bool)
    s, i = idx.search(q, 4, O.mask_from_bool(allowed))  # all-zero mask
    assert (i == -1).all()
    allowed[[4, 9, 30]] = True
    s, i = idx.search(q, 6, O.mask_from_bool(allowed))  # fewer than k survivors
    st, _ = O.normalize_rows(x, "fp32", 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    es, ei = O.brute_force_topk(qn, st, 6, allowed)
    assert_bit_equal(s, i, es, ei)
    # non-finite embeddings / queries are rejected with an error code, nothing is appended
    bad = base[:2].copy()
    bad[1, 7] = np.nan
    with pytest.raises(M.MqvsError):
        idx.add(bad)
    assert len(idx) == 80
    with pytest.raises(M.MqvsError):
        idx.search(bad, 3)
    with pytest.raises(M.MqvsError):
        idx.search(q, 0)
    with pytest.raises(M.MqvsError):
        idx.search(q, 5000)


def test_orthogonal_basis_known_answers(gpu):
    idx = native().Index(768, "fp32")
    idx.add(np.eye(768, dtype=np.float32)[:200] * 2.5)
    q = np.zeros((1, 768), np.float32)
    q[0, 17], q[0, 40] = 3.0, 1.0
    s, i = idx.search(q, 3)
    assert i[0].tolist() == [17, 40, 0]
    assert abs(s[0, 0] - 3 / np.sqrt(10)) < 1e-6 and s[0, 2] == 0.0


def test_incremental_adds_and_regrowth_keep_ids_stable(gpu):
    idx = native().Index(768, "bf16")
    x = O.synth_rows(SEED, 0, 5000, 768)
    firsts = [idx.add(x[a:b]) for a, b in ((0, 10), (10, 1500), (1500, 1501), (1501, 5000))]
    assert firsts == [0, 10, 1500, 1501]
    st, _ = O.normalize_rows(x, "bf16", 768)
    q, _ = O.planted_queries(QSEED, 6, SEED, 5000, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    assert_bit_equal(*idx.search(q, 10), *O.brute_force_topk(qn, st, 10))


def test_save_load_round_trip(gpu, tmp_path):
    idx, x, st = build(2500, 768, "bf16")
    q, _ = O.planted_queries(QSEED, 3, SEED, 2500, 768)
    before = idx.search(q, 7)
    path = str(tmp_path / "rows.mqvs")
    idx.save(path)
    other = native().Index(768, "bf16")
    other.load_file(path)
    assert len(other) == 2500
    assert_bit_equal(*other.search(q, 7), *before)
    wrong = native().Index(768, "fp32")
    with pytest.raises(native().MqvsError):
        wrong.load_file(path)
    # append-only persistence: only the new rows are written; a file that does not hold exactly first_row rows is refused
    import os
    size0 = os.path.getsize(path)
    extra = O.synth_rows(SEED + 9, 0, 700, 768)
    idx.add(extra)
    idx.save_append(path, 2500)
    assert os.path.getsize(path) == size0 + 700 * 768 * 2
    again = native().Index(768, "bf16")
    again.load_file(path)
    assert len(again) == 3200
    assert_bit_equal(*again.search(q, 7), *idx.search(q, 7))
    assert np.array_equal(again.get_rows(np.arange(2490, 2510)), idx.get_rows(np.arange(2490, 2510)))
    with pytest.raises(native().MqvsError):
        idx.save_append(path, 2500)              # the file holds 3200 rows now
    # the vector store on top: re-running the reference's ingest on an existing directory appends
    from mediquery_rag_b200 import MediQueryVectorStore
    db = str(tmp_path / "medical_db")
    vs = MediQueryVectorStore(persist_directory=db)
    vs.add_embeddings(([f"a{r}" for r in range(300)], x[:300]))
    rows_size = os.path.getsize(os.path.join(vs.collection_path, "rows.mqvs"))
    vs.add_embeddings(([f"b{r}" for r in range(50)], x[300:350]))
    assert os.path.getsize(os.path.join(vs.collection_path, "rows.mqvs")) == rows_size + 50 * 768 * 4
    vs2 = MediQueryVectorStore(persist_directory=db)
    assert len(vs2) == 350 and vs2.similarity_search_by_vector(x[320], k=1)[0].page_content == "b20"


def test_cross_shard_merge_equals_unsharded(gpu):
    """W logical shards searched one after another on one GPU and merged with the cross-rank merge kernel
     == the unsharded result == the oracle; N not divisible by W, one shard smaller than k."""
    import torch
    M = native()
    n, B, k = 10007, 5, 10
    x = O.synth_rows(SEED, 0, n, 768)
    bounds = [0, 4000, 4006, 8000, n]  # shard 1 holds 6 < k rows
    st, _ = O.normalize_rows(x, "fp32", 768)
    q, _ = O.planted_queries(QSEED, B, SEED, n, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    W = len(bounds) - 1
    qd = torch.from_numpy(q).cuda()
    all_s = torch.empty((W, B, k), dtype=torch.float32, device="cuda")
    all_i = torch.empty((W, B, k), dtype=torch.int64, device="cuda")
    shards = []
    stream = torch.cuda.current_stream().cuda_stream
    for w in range(W):
        sh = M.Index(768, "fp32")
        sh.add(x[bounds[w]:bounds[w + 1]])
        sh.search_device(qd, k, all_s[w], all_i[w], id_base=bounds[w], stream=stream)
        shards.append(sh)
    out_s = torch.empty((B, k), dtype=torch.float32, device="cuda")
    out_i = torch.empty((B, k), dtype=torch.int64, device="cuda")
    M.merge_candidates(0, all_s, all_i, out_s, out_i, stream=stream)
    torch.cuda.synchronize()
    es, ei = O.brute_force_topk(qn, st, k)
    assert_bit_equal(out_s.cpu().numpy(), out_i.cpu().numpy(), es, ei)
    ms, mi = O.merge_topk(all_s.permute(1, 0, 2).reshape(B, W * k).cpu().numpy(),
                          all_i.permute(1, 0, 2).reshape(B, W * k).cpu().numpy(), k)
    assert_bit_equal(out_s.cpu().numpy(), out_i.cpu().numpy(), ms, mi)


def test_vectorstore_on_gpu_config1(gpu, medical_docs, embedder, tmp_path):
    """BASELINE config 1: the 154 chunks of the knowledge base through the LangChain surface on the GPU."""
    from mediquery_rag_b200 import Document, MediQueryVectorStore
    docs = [Document(page_content=d["page_content"], metadata=d["metadata"]) for d in medical_docs["docs"]]
    db = str(tmp_path / "medical_db")
    vs = MediQueryVectorStore.from_documents(documents=docs, embedding=embedder, persist_directory=db)
    x = np.asarray(embedder.embed_documents([d.page_content for d in docs]), np.float32)
    st, _ = O.normalize_rows(x, "fp32", 768)
    reopened = MediQueryVectorStore(persist_directory=db, embedding_function=embedder)
    for query, k in (("高血压患者应该如何运动？", 5), ("睡眠不足有什么危害", 4), (docs[100].page_content, 3)):
        qn, _ = O.normalize_rows(np.asarray(embedder.embed_query(query), np.float32), "fp32", 768)
        es, ei = O.brute_force_topk(qn, st, k)
        for store in (vs, reopened):
            got = store.similarity_search_with_score(query, k=k)
            assert [d.page_content for d, _ in got] == [docs[r].page_content for r in ei[0]]
            assert np.allclose([dist for _, dist in got], 2.0 - 2.0 * es[0], atol=1e-6)
    assert vs.similarity_search(docs[100].page_content, k=3)[0].metadata.get("title") == docs[100].metadata["title"]
    assert len(vs.similarity_search("q", k=5, filter={"source": "《超越百岁》"})) == 5
    assert vs.similarity_search("q", k=5, filter={"source": "other"}) == []


@pytest.mark.parametrize("dtype,B", [("fp32", 16), ("bf16", 5)])
def test_full_size_config2_against_oracle_and_properties(gpu, dtype, B):
    """BASELINE config 2 size (1M x 768): full oracle comparison plus size-independent properties."""
    n, k = 1_000_000, 10
    idx = native().Index(768, dtype, capacity_rows=n)
    idx.add_synthetic(SEED, 0, n)
    q, planted = O.planted_queries(QSEED, B, SEED, n, 768)
    s, i = idx.search(q, k)
    assert all(i[b, 0] == planted[b] for b in range(0, B, 2))           # planted neighbours are found first
    assert np.all(np.diff(s, axis=1) <= 0) and (i >= 0).all()            # sorted, full
    assert all(len(set(row)) == k for row in i.tolist())                 # no duplicates
    # returned scores are the canonical fp32 scores of the returned rows (rows regenerated from the hash)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    for b in range(B):
        rows = np.concatenate([O.synth_rows(SEED, int(r), 1, 768) for r in i[b]])
        st_b, _ = O.normalize_rows(rows, dtype, 768)
        assert np.array_equal(O.canonical_scores(qn[b:b + 1], st_b)[0].view(np.uint32), s[b].view(np.uint32))
    # idempotence + batch independence: one query alone gives the same answer as inside the batch
    s1, i1 = idx.search(q[3:4], k)
    assert np.array_equal(i1[0], i[3]) and np.array_equal(s1[0], s[3])
    # full oracle
    st, _ = O.normalize_rows(O.synth_rows(SEED, 0, n, 768), dtype, 768)
    es, ei = O.brute_force_topk(qn, st, k)
    assert_bit_equal(s, i, es, ei)


def test_hypothesis_random_shapes(gpu):
    from hypothesis import given, settings, strategies as hs

    @settings(max_examples=25, deadline=None, derandomize=True)
    @given(n=hs.integers(1, 5000), B=hs.integers(1, 9), k=hs.integers(1, 40), frac=hs.sampled_from([None, 0.0, 0.02, 0.5, 1.0]),
           dtype=hs.sampled_from(["fp32", "bf16"]), seed=hs.integers(0, 2 ** 31))
    def run(n, B, k, frac, dtype, seed):
        idx, x, st = build(n, 768, dtype, seed)
        q = O.synth_rows(seed + 1, 0, B, 768)
        qn, _ = O.normalize_rows(q, "fp32", 768)
        mask = None if frac is None else O.mask_from_bool(np.random.default_rng(seed).random(n) < frac)
        assert_bit_equal(*idx.search(q, k, mask), *O.brute_force_topk(qn, st, k, mask))
        idx.close()

    run()


# ------------------------------------------------------------------------------------------ tensor-core path
TC_CASES = [
    # n, dtype, B, k, mask fraction   (path forced to the tcgen05 shortlist + fp32 rescoring pipeline)
    (8192, "bf16", 16, 10, None),
    (8192, "fp32", 16, 10, None),
    (50000, "bf16", 100, 10, None),
    (50000, "fp32", 130, 10, None),
    (100003, "bf16", 300, 100, None),     # k' = 256, ragged last corpus tile, 3 query tiles
    (100003, "fp32", 64, 32, None),
    (70000, "bf16", 64, 10, 0.3),
    (70000, "fp32", 33, 64, 0.05),
    (300000, "bf16", 512, 10, None),
    (4096, "bf16", 1, 5, None),           # smallest store the path accepts, single query
    (65536, "fp32", 128, 128, 0.001),     # fewer allowed rows (about 65) than k: padded results
    (37000, "bf16", 128, 10, None),       # fewer corpus tiles (145) than CTAs: some CTAs own no tile at all
    (200001, "fp32", 7, 1, 0.02),         # k = 1, ragged tail, sparse mask
    (151552, "bf16", 96, 100, None),      # exactly 4 tiles per CTA (148 x 4 x 256 rows), k' = 256
    (100003, "bf16", 256, 100, None),     # two query tiles in the single-launch kernel, both full
    (70000, "fp32", 200, 10, 0.3),        # two query tiles, the second ragged, with a mask
    (300000, "bf16", 129, 10, None),      # one query in the second tile
    (20000, "fp32", 160, 32, None),       # fewer corpus tiles (79) than CTAs, two query tiles
]


# path 2 = tensor-core path (single cooperative launch for B <= 256, segmented pipeline above); 4 = always segmented
TC_CASES_PATHS = [c + (2,) for c in TC_CASES] + [c + (4,) for c in TC_CASES if c[2] <= 256]


@pytest.mark.parametrize("n,dtype,B,k,frac,path", TC_CASES_PATHS)
def test_tensor_core_path_matches_oracle_bitwise(gpu, n, dtype, B, k, frac, path):
    idx, x, st = build(n, 768, dtype, seed=0x5EED0003)
    idx.set_path(path)
    q, planted = O.planted_queries(QSEED, B, 0x5EED0003, n, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    mask = None if frac is None else O.mask_from_bool(np.random.default_rng(11).random(n) < frac)
    s, i = idx.search(q, k, mask)
    assert idx.stats()["last_path"] == 1
    assert_bit_equal(s, i, *O.brute_force_topk(qn, st, k, mask))
    idx.close()


@pytest.mark.parametrize("path,dups,must_repair", [(4, 300, True), (2, 300, False), (2, 1500, True)])
def test_tensor_core_path_falls_back_to_exact_scan_when_margin_cannot_be_proven(gpu, path, dups, must_repair):
    """Hundreds of exact copies of the best row.  Segmented pipeline (path 4): the k'-th approximate score ties the k-th,
    the margin check must flag the query and the exact SIMT scan must repair it.  Single-launch pipeline (path 2): every
    candidate above the learnt threshold is rescored, so 300 ties are resolved exactly without a repair; 1500 ties
    overflow its shortlist, the streaming selection takes over and the margin check flags the query.  In every case:
    same ids (lowest row ids of the tie) and score bits as the oracle."""
    n = 20000
    x = O.synth_rows(SEED, 0, n, 768)
    x[5000:5000 + dups] = x[123]
    idx = native().Index(768, "bf16")
    idx.add(x)
    idx.set_path(path)
    q = np.stack([x[123] + 0.01 * x[7], x[9000], x[123]])
    qn, _ = O.normalize_rows(q, "fp32", 768)
    st, _ = O.normalize_rows(x, "bf16", 768)
    s, i = idx.search(q, 10)
    assert_bit_equal(s, i, *O.brute_force_topk(qn, st, 10))
    stats = idx.stats()
    assert stats["last_path"] == 1
    if must_repair:
        assert stats["exact_fallbacks"] >= 2


def test_auto_path_choice_and_agreement_between_paths(gpu):
    n = 131072
    idx = native().Index(768, "bf16", capacity_rows=n)
    idx.add_synthetic(SEED, 0, n)
    q, _ = O.planted_queries(QSEED, 48, SEED, n, 768)
    idx.set_path(0)
    auto = idx.search(q, 10)
    assert idx.stats()["last_path"] == 1          # a real batch on a large store -> tensor cores
    mid = idx.search(q[:4], 10)
    assert idx.stats()["last_path"] == 1
    single = idx.search(q[:1], 10)
    assert idx.stats()["last_path"] == 0          # single query -> bandwidth-bound SIMT scan
    idx.set_path(1)
    simt = idx.search(q, 10)
    assert_bit_equal(*auto, *simt)
    assert_bit_equal(mid[0], mid[1], simt[0][:4], simt[1][:4])
    assert_bit_equal(single[0], single[1], simt[0][:1], simt[1][:1])


def test_full_size_config3_shape_properties(gpu):
    """BASELINE config 3 (10M x 768 bf16, k=100, tensor cores + fp32 rerank) at a batch the test can afford;
    size-independent checks: planted neighbours first, sorted, scores = canonical fp32 of the returned rows,
    and agreement with the exact SIMT scan on a subset of the queries."""
    n, B, k = 10_000_000, 256, 100
    idx = native().Index(768, "bf16", capacity_rows=n)
    idx.add_synthetic(0x5EED0003, 0, n)
    q, planted = O.planted_queries(QSEED, B, 0x5EED0003, n, 768)
    s, i = idx.search(q, k)
    st = idx.stats()
    assert st["last_path"] == 1 and st["exact_fallbacks"] == 0
    assert all(i[b, 0] == planted[b] for b in range(0, B, 2))
    assert np.all(np.diff(s, axis=1) <= 0) and (i >= 0).all()
    qn, _ = O.normalize_rows(q, "fp32", 768)
    for b in (0, 1, 77):
        rows = np.concatenate([O.synth_rows(0x5EED0003, int(r), 1, 768) for r in i[b]])
        st_b, _ = O.normalize_rows(rows, "bf16", 768)
        assert np.array_equal(O.canonical_scores(qn[b:b + 1], st_b)[0].view(np.uint32), s[b].view(np.uint32))
    idx.set_path(1)
    es, ei = idx.search(q[:6], k)
    assert_bit_equal(s[:6], i[:6], es, ei)


@pytest.mark.parametrize("path,dups", [(4, 300), (2, 1500)])
def test_device_side_repair_on_the_async_path_and_host_repair_beyond_its_capacity(gpu, path, dups):
    """mqvs_search_async cannot ask the host, so flagged queries are re-scanned by a device-driven pass (up to 64 per
    call); the blocking call additionally repairs any number.  100 queries all aimed at a many-fold duplicated row
    (enough copies to defeat the shortlist of the pipeline under test)."""
    import torch
    n, B, k = 20000, 100, 10
    x = O.synth_rows(SEED, 0, n, 768)
    x[5000:5000 + dups] = x[123]
    noise = O.synth_rows(77, 0, B, 768)
    q = x[123][None, :] + 0.01 * noise
    qn, _ = O.normalize_rows(q, "fp32", 768)
    st, _ = O.normalize_rows(x, "bf16", 768)
    es, ei = O.brute_force_topk(qn, st, k)
    idx = native().Index(768, "bf16")
    idx.add(x)
    idx.set_path(path)
    s, i = idx.search(q, k)                       # blocking: device repairs 64, host repairs the other 36
    assert_bit_equal(s, i, es, ei)
    assert idx.stats()["exact_fallbacks"] >= B and idx.stats()["unrepaired"] == 0
    qd = torch.from_numpy(q[:40]).cuda()          # async, within the device pass' capacity
    out_s = torch.empty((40, k), dtype=torch.float32, device="cuda")
    out_i = torch.empty((40, k), dtype=torch.int64, device="cuda")
    idx.search_device(qd, k, out_s, out_i, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert_bit_equal(out_s.cpu().numpy(), out_i.cpu().numpy(), es[:40], ei[:40])
    assert idx.stats()["unrepaired"] == 0
    # async, MORE flagged queries than one device repair round takes: the rounds are sized to the batch, so the asynchronous
    # search is exact on its own (nothing is left for a host that is not looking)
    qd = torch.from_numpy(q).cuda()
    out_s = torch.empty((B, k), dtype=torch.float32, device="cuda")
    out_i = torch.empty((B, k), dtype=torch.int64, device="cuda")
    idx.search_device(qd, k, out_s, out_i, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert_bit_equal(out_s.cpu().numpy(), out_i.cpu().numpy(), es, ei)
    assert idx.stats()["unrepaired"] == 0


def test_hypothesis_random_shapes_tensor_path(gpu):
    from hypothesis import given, settings, strategies as hs

    @settings(max_examples=12, deadline=None, derandomize=True)
    @given(n=hs.integers(4096, 40000), B=hs.integers(1, 200), k=hs.integers(1, 128),
           frac=hs.sampled_from([None, 0.0, 0.003, 0.5, 1.0]), dtype=hs.sampled_from(["fp32", "bf16"]), seed=hs.integers(0, 2 ** 31))
    def run(n, B, k, frac, dtype, seed):
        idx, x, st = build(n, 768, dtype, seed)
        idx.set_path(2 if seed % 3 else 4)
        q = O.synth_rows(seed + 1, 0, B, 768)
        qn, _ = O.normalize_rows(q, "fp32", 768)
        mask = None if frac is None else O.mask_from_bool(np.random.default_rng(seed).random(n) < frac)
        got = idx.search(q, k, mask)
        assert idx.stats()["last_path"] == 1
        assert_bit_equal(*got, *O.brute_force_topk(qn, st, k, mask))
        idx.close()

    run()


@pytest.mark.parametrize("n,dim,dtype,B,k,frac", [(154, 768, "fp32", 1, 5, None), (154, 768, "bf16", 4, 16, None),
                                                  (7, 768, "fp32", 2, 10, None), (50000, 300, "bf16", 3, 4, 0.2),
                                                  (300000, 768, "fp32", 1, 10, None), (300000, 768, "fp32", 2, 3, 0.001)])
def test_single_launch_latency_path_equals_multi_launch_path_and_oracle(gpu, n, dim, dtype, B, k, frac):
    """B <= 4, k <= 16 searches run as ONE kernel (in-kernel query normalisation, last-CTA merge); path 3 forces the
    classic normalise + scan + merge launches.  Both must equal the oracle bit for bit; non-finite queries still raise."""
    idx, x, st = build(n, dim, dtype)
    q, _ = O.planted_queries(QSEED, B, SEED, n, dim)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    mask = None if frac is None else O.mask_from_bool(np.random.default_rng(5).random(n) < frac)
    want = O.brute_force_topk(qn, st, k, mask)
    idx.set_path(1)
    launches0 = idx.stats()["kernel_launches"]
    fused = idx.search(q, k, mask)
    assert idx.stats()["kernel_launches"] - launches0 == 1       # the whole search was one launch
    fused2 = idx.search(q, k, mask)                               # the ticket resets itself
    idx.set_path(3)
    classic = idx.search(q, k, mask)
    assert_bit_equal(*fused, *want)
    assert_bit_equal(*fused2, *want)
    assert_bit_equal(*classic, *want)
    idx.set_path(1)
    bad = q.copy()
    bad[B - 1, 0] = np.inf
    with pytest.raises(native().MqvsError):
        idx.search(bad, k)
    assert_bit_equal(*idx.search(q, k, mask), *want)              # and the handle keeps working afterwards
    # a host caller's queries ride in the kernel's parameter block (scan_fused_inline_kernel); device-resident queries take
    # the same kernel reading them from global memory
    import torch
    out_s = torch.empty((B, k), dtype=torch.float32, device="cuda")
    out_i = torch.empty((B, k), dtype=torch.int64, device="cuda")
    mask_dev = None if mask is None else torch.from_numpy(mask.view(np.int32)).cuda()
    l0 = idx.stats()["kernel_launches"]
    idx.search_device(torch.from_numpy(q).cuda(), k, out_s, out_i, mask=mask_dev)
    torch.cuda.synchronize()
    assert idx.stats()["kernel_launches"] - l0 == 1 and idx.stats()["last_pipeline"] == 1
    assert_bit_equal(out_s.cpu().numpy(), out_i.cpu().numpy(), *want)


def test_tensor_path_batches_larger_than_one_launch_chunk(gpu):
    """More than 4096 queries are processed in launch chunks (candidate-list memory is bounded): blocking call."""
    n, B, k = 8192, 4096 + 700, 10
    idx, x, st = build(n, 768, "bf16")
    idx.set_path(2)
    q = O.synth_rows(QSEED, 0, B, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    s, i = idx.search(q, k)
    assert idx.stats()["last_path"] == 1
    assert_bit_equal(s, i, *O.brute_force_topk(qn, st, k))


@pytest.mark.parametrize("path", [2, 4])
@pytest.mark.parametrize("dtype,k", [("bf16", 10), ("fp32", 100)])
def test_tensor_path_adversarial_order_forces_list_compaction(gpu, dtype, k, path):
    """Rows ordered by INCREASING similarity to the queries: thresholds learnt on earlier segments never filter anything,
    every score is appended, the per-(CTA, query) lists overflow and the warp-sort compaction path runs for real."""
    n, B = 160_000, 16
    rng = np.random.default_rng(42)
    q0 = rng.standard_normal(768).astype(np.float32)
    q0 /= np.linalg.norm(q0)
    t = (np.arange(n, dtype=np.float32) / n * 0.9)[:, None]
    x = (1.0 - t) * O.synth_rows(SEED, 0, n, 768) / np.sqrt(256.0).astype(np.float32) + t * q0[None, :]
    q = q0[None, :] + 0.02 * rng.standard_normal((B, 768)).astype(np.float32)
    idx = native().Index(768, dtype)
    idx.add(x)
    idx.set_path(path)
    st, _ = O.normalize_rows(x, dtype, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    s, i = idx.search(q, k)
    assert idx.stats()["last_path"] == 1
    es, ei = O.brute_force_topk(qn, st, k)
    assert (ei[:, 0] > n - 20000).all()         # the best rows really are among the last ones
    assert_bit_equal(s, i, es, ei)


def test_single_launch_tensor_pipeline_one_launch_and_agreement_with_the_segmented_one(gpu):
    """B <= 128 on the tensor path: prep kernel + ONE cooperative persistent kernel (threshold learning through grid
    barriers, scan, selection, fp32 rescoring, margin check); the blocking call enqueues the repair launches only when a
    query was flagged, so a clean search is exactly two launches.  Must agree bit
    for bit with the segmented pipeline (path 4), the exact SIMT scan (path 1) and the oracle; repeated calls on one
    handle reuse the barrier counters; a non-finite query still raises and leaves the handle usable."""
    n, B, k = 300_000, 64, 10
    idx = native().Index(768, "fp32", capacity_rows=n)
    idx.add_synthetic(SEED, 0, n)
    q, planted = O.planted_queries(QSEED, B, SEED, n, 768)
    idx.set_path(0)
    l0 = idx.stats()["kernel_launches"]
    one = idx.search(q, k)
    st = idx.stats()
    # (an fp32 store's first tensor search after an add also converts the new rows into the bf16 scoring shadow: one launch)
    assert st["last_path"] == 1 and st["kernel_launches"] - l0 == 3 and st["exact_fallbacks"] == 0
    again = idx.search(q, k)
    assert idx.stats()["kernel_launches"] - st["kernel_launches"] == 2
    assert all(one[1][b, 0] == planted[b] for b in range(0, B, 2))
    idx.set_path(4)
    seg = idx.search(q, k)
    idx.set_path(1)
    simt = idx.search(q, k)
    assert_bit_equal(*one, *simt)
    assert_bit_equal(*again, *simt)
    assert_bit_equal(*seg, *simt)
    idx.set_path(0)
    bad = q.copy()
    bad[17, 5] = np.nan
    with pytest.raises(native().MqvsError):
        idx.search(bad, k)
    assert_bit_equal(*idx.search(q, k), *simt)
    # ragged batch sizes around the tile / warp boundaries
    for b in (3, 31, 33, 127, 128):
        got = idx.search(q[:1].repeat(b, axis=0) + 0.01 * O.synth_rows(b, 0, b, 768), k)
        idx.set_path(1)
        want = idx.search(q[:1].repeat(b, axis=0) + 0.01 * O.synth_rows(b, 0, b, 768), k)
        idx.set_path(0)
        assert_bit_equal(*got, *want)


@pytest.mark.parametrize("W,B,k", [(2, 1, 10), (4, 37, 10), (8, 100, 10), (3, 5, 100)])
def test_peer_exchange_kernel_equals_merge_of_gathered_candidates(gpu, W, B, k):
    """The peer exchange kernel (candidate exchange through peer inboxes + flags + merge) against mqvs_merge_candidates on the
    gathered candidates and the oracle merge.  W logical ranks share this one GPU, inboxes connected by pointer; all
    ranks run in ONE cooperative launch (mqvs_xchg_merge_colocated: kernels that spin on each other's flags must not be
    separate launches on one device), so the CTAs really wait for each other's flags; ties across ranks (equal scores)
    must resolve to the lower global id; repeated calls exercise the parity double-buffering."""
    import torch
    M = native()
    ex = [M._native.PeerExchange(0, r, W, max_entries=4096) for r in range(W)]
    try:
        for e in ex:
            e.connect_ptrs([x.base for x in ex])
        rng = np.random.default_rng(W * 1000 + B)
        for call in range(3):
            sc = np.sort(rng.integers(0, 50, size=(W, B, k)).astype(np.float32) / 7.0, axis=2)[:, :, ::-1].copy()  # many ties
            ids = np.stack([rng.permutation(B * k * 4)[:B * k].reshape(B, k) * W + r for r in range(W)]).astype(np.int64)
            ids[W - 1, :, k - 1] = -1                      # an empty slot (a shard with fewer than k rows)
            sc[W - 1, :, k - 1] = -np.inf
            d_s = [torch.from_numpy(sc[r]).cuda() for r in range(W)]
            d_i = [torch.from_numpy(ids[r]).cuda() for r in range(W)]
            o_s = [torch.empty((B, k), dtype=torch.float32, device="cuda") for _ in range(W)]
            o_i = [torch.empty((B, k), dtype=torch.int64, device="cuda") for _ in range(W)]
            torch.cuda.synchronize()
            M._native.PeerExchange.merge_colocated(ex, d_s, d_i, o_s, o_i, stream=torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            es, ei = O.merge_topk(sc.transpose(1, 0, 2).reshape(B, W * k), ids.transpose(1, 0, 2).reshape(B, W * k), k)
            g_s, g_i = torch.empty_like(o_s[0]), torch.empty_like(o_i[0])
            M._native.merge_candidates(0, torch.from_numpy(sc).cuda(), torch.from_numpy(ids).cuda(), g_s, g_i)
            torch.cuda.synchronize()
            for r in range(W):
                assert_bit_equal(o_s[r].cpu().numpy(), o_i[r].cpu().numpy(), es, ei)
            assert_bit_equal(g_s.cpu().numpy(), g_i.cpu().numpy(), es, ei)
    finally:
        for e in ex:
            e.close()


def test_device_side_where_evaluation_equals_host_evaluator(gpu):
    """K5 filter_eval_kernel: every filter of tests/test_device_filter_cpu.py compiled by DeviceFilter, evaluated on the
    device into the packed bitmask and consumed by a search with k = number of allowed rows: the returned id SET must be
    exactly the rows the host evaluator allows (and the count the kernel reports must match); then the same through
    MediQueryVectorStore.similarity_search_by_vector(filter=...), where the device path is taken automatically."""
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_device_filter_cpu import FILTERS, make_cols
    from mediquery_rag_b200 import MediQueryVectorStore
    from mediquery_rag_b200._native import DeviceMask
    from mediquery_rag_b200.filters import DeviceFilter, evaluate

    cols = make_cols()
    n = len(cols)
    x = O.synth_rows(SEED, 0, n, 768)
    idx = native().Index(768, "fp32")
    idx.add(x)
    live = np.random.default_rng(1).random(n) > 0.1
    dev = DeviceFilter(idx, cols)
    q = O.synth_rows(QSEED, 0, 1, 768)
    for where in FILTERS:
        want = evaluate(where, cols) & live
        mask, allowed = dev.build(where, live)
        assert isinstance(mask, DeviceMask) and allowed == int(want.sum()), where
        if allowed:
            s, i = idx.search(q, allowed, mask)
            assert set(i[0].tolist()) == set(np.flatnonzero(want).tolist()), where
    idx.close()

    store = MediQueryVectorStore(collection_metadata={"hnsw:space": "cosine"})  # x holds raw (non-unit) vectors
    store.add_embeddings(([f"doc {r}" for r in range(n)], x), metadatas=[cols.row(r) for r in range(n)])
    for where in FILTERS[::3]:
        want = evaluate(where, store._cols)
        mask, allowed = store._allow_mask(where, None)
        assert isinstance(mask, DeviceMask) and allowed == int(want.sum())
        docs = store.similarity_search_by_vector(q[0], k=7, filter=where)
        es, ei = O.brute_force_topk(O.normalize_rows(q, "fp32", 768)[0], O.normalize_rows(x, "fp32", 768)[0], min(7, allowed),
                                    O.mask_from_bool(want)) if allowed else (None, np.zeros((1, 0), np.int64))
        assert [d.page_content for d in docs] == [f"doc {r}" for r in ei[0].tolist()]
    # a where_document filter or a mixed-type field still goes through the host evaluator, same answers
    mask, allowed = store._allow_mask({"mixed": 1}, None)
    assert not isinstance(mask, DeviceMask) and allowed == int(evaluate({"mixed": 1}, store._cols).sum())


def test_one_handle_from_several_host_threads_and_two_handles_at_once(gpu):
    """LangGraph may call the retriever off the main thread (/root/reference/src/agents/graph.py:95 opens its sqlite
    checkpointer with check_same_thread=False): a handle serialises its calls, so four threads searching one store get
    the same answers as one thread.  Two different stores searched concurrently each launch the cooperative
    single-launch kernel (grid barriers): cooperative launches are scheduled whole, so they must not deadlock."""
    import threading
    n, B, k = 200_000, 48, 10
    a = native().Index(768, "bf16", capacity_rows=n)
    a.add_synthetic(SEED, 0, n)
    b = native().Index(768, "fp32", capacity_rows=n)
    b.add_synthetic(SEED + 1, 0, n)
    q, _ = O.planted_queries(QSEED, B, SEED, n, 768)
    want_a, want_b = a.search(q, k), b.search(q, k)
    assert a.stats()["last_pipeline"] == 3 and b.stats()["last_pipeline"] == 3
    errors = []

    def worker(idx, want, reps, lo):
        try:
            for r in range(reps):
                sl = slice(lo, lo + 1 + (r % 40))          # batch sizes 1..40: SIMT, single-launch SIMT and tensor paths
                s, i = idx.search(q[sl], k)
                if not (np.array_equal(i, want[1][sl]) and np.array_equal(s.view(np.uint32), want[0][sl].view(np.uint32))):
                    errors.append((lo, r))
        except Exception as exc:  # noqa: BLE001
            errors.append(repr(exc))

    threads = [threading.Thread(target=worker, args=(a, want_a, 40, t)) for t in range(4)]
    threads += [threading.Thread(target=worker, args=(b, want_b, 40, t)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in threads), "searches did not finish (deadlock?)"
    assert errors == []


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_single_launch_pipeline_on_a_corpus_made_of_ties(gpu, dtype):
    """A corpus drawn from only 40 distinct vectors (every score value occurs ~2500 times): thresholds learnt from
    chunk maxima tie with thousands of rows, candidate lists overflow and are compacted, the tail's shortlist overflows
    into the streaming selection, margins cannot be proven and the repair scan must run -- and the answer must still
    be the oracle's, i.e. the LOWEST row ids of each tie group, on the single-launch and the segmented pipeline."""
    n, B, k = 100_000, 24, 10
    rng = np.random.default_rng(7)
    base = O.synth_rows(SEED, 0, 40, 768)
    pick = rng.integers(0, 40, n)
    x = base[pick]
    q = base[rng.integers(0, 40, B)] + 0.05 * O.synth_rows(QSEED, 0, B, 768)
    st, _ = O.normalize_rows(x, dtype, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    es, ei = O.brute_force_topk(qn, st, k)
    idx = native().Index(768, dtype)
    idx.add(x)
    for path in (2, 4):
        idx.set_path(path)
        s, i = idx.search(q, k)
        assert idx.stats()["last_path"] == 1
        assert_bit_equal(s, i, es, ei)
    # a mask that leaves exactly one row of every tie group (ids then are forced) and one that leaves nothing
    keep = np.zeros(n, bool)
    keep[[int(np.flatnonzero(pick == g)[-1]) for g in range(40)]] = True
    for allowed in (keep, np.zeros(n, bool)):
        mask = O.mask_from_bool(allowed)
        want = O.brute_force_topk(qn, st, k, mask)
        for path in (2, 4):
            idx.set_path(path)
            assert_bit_equal(*idx.search(q, k, mask), *want)


def test_device_side_where_evaluation_random_trees(gpu):
    """Random `where` trees (hypothesis, derandomised) through the real filter_eval_kernel: reported count and the id set
    returned by a search with k = allowed rows must match the host evaluator."""
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from hypothesis import given, settings
    from test_device_filter_cpu import make_cols, where_trees
    from mediquery_rag_b200.filters import DeviceFilter, evaluate

    cols = make_cols(seed=3)
    n = len(cols)
    idx = native().Index(768, "bf16")
    idx.add(O.synth_rows(SEED, 0, n, 768))
    live = np.random.default_rng(2).random(n) > 0.2
    dev = DeviceFilter(idx, cols)
    q = O.synth_rows(QSEED, 0, 1, 768)

    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(where=where_trees())
    def run(where):
        want = evaluate(where, cols) & live
        mask, allowed = dev.build(where, live)
        assert allowed == int(want.sum()), where
        if allowed:
            _, i = idx.search(q, allowed, mask)
            assert set(i[0].tolist()) == set(np.flatnonzero(want).tolist()), where

    run()


def test_plain_c_host_program_through_the_c_abi(gpu, tmp_path):
    """examples/mqvs_demo.c (no Python, no torch): create, add 154 rows, search k=5 from compiled C against libmqvs.so."""
    import os, shutil, subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir, exe = os.path.join(root, "mediquery-rag_b200"), str(tmp_path / "mqvs_demo")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(root, "include"),
                    os.path.join(root, "examples", "mqvs_demo.c"), "-L" + lib_dir, "-lmqvs", "-Wl,-rpath," + lib_dir, "-lm",
                    "-o", exe], check=True)
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr
    lines = run.stdout.splitlines()
    assert "154 rows stored" in lines[1] and lines[2].split()[:2] == ["row", "42"]


def test_full_size_config4_shard_and_config5_filter_properties(gpu):
    """One rank's share of BASELINE config 4 (100M x 768 bf16 over 8 GPUs = 12.5M rows) and the config-5 filter on it,
    through size-independent properties: planted neighbours first, scores sorted and equal to the canonical fp32 score of
    the returned rows (rows regenerated on the CPU from the counter-based generator), the single-launch tensor pipeline
    (B = 1 and B = 64) bit-identical to the exact SIMT scan, and with a device-built `where` mask every returned row
    satisfies the predicate while the unfiltered winners that satisfy it are all still there."""
    from mediquery_rag_b200.filters import DeviceFilter, MetadataColumns
    n, k = 12_500_000, 10
    seed = 0x5EED0004
    idx = native().Index(768, "bf16", capacity_rows=n)
    idx.add_synthetic(seed, 0, n)
    q, planted = O.planted_queries(QSEED, 64, seed, n, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    s64, i64 = idx.search(q, k)
    assert idx.stats()["last_pipeline"] == 3 and idx.stats()["exact_fallbacks"] == 0
    s1, i1 = idx.search(q[:1], k)
    assert idx.stats()["last_pipeline"] == 3          # single queries take the tensor pipeline from 1M rows on
    assert all(i64[b, 0] == planted[b] for b in range(0, 64, 2))
    assert np.all(np.diff(s64, axis=1) <= 0) and (i64 >= 0).all()
    for b in (0, 1, 37):
        rows = np.concatenate([O.synth_rows(seed, int(r), 1, 768) for r in i64[b]])
        st_b, _ = O.normalize_rows(rows, "bf16", 768)
        assert np.array_equal(O.canonical_scores(qn[b:b + 1], st_b)[0].view(np.uint32), s64[b].view(np.uint32))
    idx.set_path(1)
    es, ei = idx.search(q[:3], k)
    assert_bit_equal(s64[:3], i64[:3], es, ei)
    assert_bit_equal(s1, i1, es[:1], ei[:1])
    idx.set_path(0)
    # config-5 style categorical filter, evaluated on the device
    cols = MetadataColumns()
    src = (np.arange(n, dtype=np.int64) * 2654435761 >> 7) % 100
    cols.extend_columns(n, {"source_id": src})
    mask, allowed = DeviceFilter(idx, cols).build({"source_id": {"$in": list(range(10))}}, None)
    assert allowed == int((src < 10).sum())
    fs, fi = idx.search(q, k, mask)
    assert (src[fi] < 10).all() and np.all(np.diff(fs, axis=1) <= 0)
    for b in range(64):
        keep = i64[b][src[i64[b]] < 10]              # unfiltered winners that pass the filter stay, in the same order
        assert np.array_equal(fi[b][:keep.shape[0]], keep)


@pytest.mark.parametrize("dtype,B", [("bf16", 64), ("fp32", 20), ("bf16", 200), ("bf16", 300)])
def test_band_pass_repairs_dense_neighbourhoods_without_exact_scans(gpu, dtype, B):
    """Clustered corpus (2000-member clusters at cosine 0.98 to their centre): the scores around rank k are closer
    together than the tensor-core error bound, so the margin check flags (nearly) every query.  The band pass -- one more
    tensor-core pass that collects and rescoring every row within eps of the k-th exact score -- must settle them all:
    bit-identical to the oracle, band_repairs > 0, and no query left for the exact scan.  B = 300 goes through the
    segmented pipeline first (3 query tiles), B <= 256 through the single-launch one (one or two query tiles)."""
    n, members, k = 200_000, 2000, 10
    rng = np.random.default_rng(0)
    sigma = np.sqrt((1.0 / 0.98 ** 2 - 1.0) / 256.0)
    centres = rng.standard_normal((n // members, 768)).astype(np.float32)
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    x = (np.repeat(centres, members, axis=0) + sigma * rng.uniform(-1, 1, (n, 768))).astype(np.float32)
    q = (x[rng.integers(0, n, B)] + 0.5 * sigma * rng.uniform(-1, 1, (B, 768))).astype(np.float32)
    st, _ = O.normalize_rows(x, dtype, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    es, ei = O.brute_force_topk(qn, st, k)
    idx = native().Index(768, dtype)
    idx.add(x)
    idx.set_path(2)
    s, i = idx.search(q, k)
    stats = idx.stats()
    assert_bit_equal(s, i, es, ei)
    assert stats["band_repairs"] >= min(B, 128) // 2 and stats["exact_fallbacks"] == 0, stats
    if B <= 256:                                   # the same through the all-device call (what the sharded path uses)
        import torch
        idx.set_path(2)
        before = idx.stats()["band_repairs"]
        out_s = torch.empty((B, k), dtype=torch.float32, device="cuda")
        out_i = torch.empty((B, k), dtype=torch.int64, device="cuda")
        idx.search_device(torch.from_numpy(q).cuda(), k, out_s, out_i, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert_bit_equal(out_s.cpu().numpy(), out_i.cpu().numpy(), es, ei)
        assert idx.stats()["band_repairs"] > before and idx.stats()["exact_fallbacks"] == 0
        with pytest.raises(TypeError):             # raw pointers cross the ABI: the binding refuses a float64 tensor
            idx.search_device(torch.from_numpy(q.astype(np.float64)).cuda(), k, out_s, out_i)
    idx.set_path(4)                                # without the band pass the same queries need exact scans
    s4, i4 = idx.search(q[:16], k)
    assert_bit_equal(s4, i4, es[:16], ei[:16])
    assert idx.stats()["exact_fallbacks"] > 0


def _raw_rows(n, seed=SEED, lo=0.25, hi=4.0):
    """non-unit vectors: synthetic rows (norm ~16) rescaled to lengths spread over [lo, hi]"""
    x = O.synth_rows(seed, 0, n, 768)
    scale = np.random.default_rng(11).uniform(lo, hi, (n, 1)).astype(np.float32) / np.float32(16.0)
    return (x * scale).astype(np.float32)


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("space", ["l2", "ip"])
def test_l2_and_ip_spaces_over_raw_vectors_are_bit_exact(gpu, space, dtype, tmp_path):
    """Chroma's hnsw:space on NON-unit vectors: K1 keeps ||x|| per row; the exact scan and the
    tensor pipelines (epilogue transform + exact rescoring; single-launch up to 256 queries, segmented above) map the cosine of the
    normalised pair to the space's ranking score (common.cuh:rank_score) before top-k.  Bit-identical on every path to
    oracle/canonical.c:orc_topk_metric, which tests/test_spaces.py ties to a float64 brute force over raw vectors."""
    import torch
    M = native()
    n = 70_001  # large enough that the cosine space would take the tensor-core path for B >= 2
    x = _raw_rows(n)
    idx = M.Index(768, dtype)
    idx.set_metric(space)
    idx.add(x[:30_000])
    idx.add(x[30_000:])  # regrowth moves the norms with the rows
    nr = O.row_norms(x, 768)
    assert np.array_equal(idx.get_norms().view(np.uint32), nr.view(np.uint32))
    assert idx.norm_deviation() == float(np.abs(nr - np.float32(1.0)).max())
    st, _ = O.normalize_rows(x, dtype, 768)
    rng = np.random.default_rng(2)
    for B, k, frac in ((1, 10, None), (3, 5, None), (9, 10, 0.3), (2, 40, None), (1, 200, 0.5), (200, 100, 0.5), (300, 10, None),
                       (64, 10, None)):
        q = (_raw_rows(B, seed=QSEED) * np.float32(3.0) + 0.5 * x[rng.integers(0, n, B)]).astype(np.float32)
        qn, _ = O.normalize_rows(q, "fp32", 768)
        mask = None if frac is None else O.mask_from_bool(rng.random(n) < frac)
        s, i = idx.search(q, k, mask)
        es, ei = O.brute_force_topk(qn, st, k, mask, metric=space, norms=nr, q_norms=O.row_norms(q, 768))
        assert_bit_equal(s, i, es, ei)
        stats = idx.stats()
        tensor = B >= 2 and k <= 128  # api.cu:use_tensor_path below 1M rows: every batch of two or more queries
        assert stats["last_path"] == int(tensor) and stats["last_pipeline"] == ((3 if B <= 256 else 2) if tensor else 0), stats
        assert stats["exact_fallbacks"] == 0
    assert not np.array_equal(O.brute_force_topk(qn, st, k)[1], ei)  # the space matters on these rows
    idx.set_path(1)                                                  # the exact scan alone gives the same answer
    assert_bit_equal(*idx.search(q, k), es, ei)
    assert idx.stats()["last_path"] == 0
    idx.set_path(4)                                                  # and so does the segmented tensor pipeline
    assert_bit_equal(*idx.search(q, k), es, ei)
    assert idx.stats()["last_pipeline"] == 2
    idx.set_path(0)
    # all-device call on a caller stream, with an id base (the sharded path)
    qd = torch.from_numpy(q).cuda()
    out_s = torch.empty((B, k), dtype=torch.float32, device="cuda")
    out_i = torch.empty((B, k), dtype=torch.int64, device="cuda")
    idx.search_device(qd, k, out_s, out_i, id_base=1000, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert_bit_equal(out_s.cpu().numpy(), out_i.cpu().numpy(), es, ei + 1000)
    assert idx.stats()["last_pipeline"] == 3
    # the metric can be switched once the norms are kept: cosine ranks the normalised rows (tensor path for B = 64)
    idx.set_metric("cosine")
    s, i = idx.search(q, k)
    assert_bit_equal(s, i, *O.brute_force_topk(qn, st, k))
    assert idx.stats()["last_path"] == 1
    idx.set_metric(space)
    # persistence: the norms travel next to the rows (rows.mqvs.norms), appends included
    path = str(tmp_path / "rows.mqvs")
    idx.save(path)
    extra = _raw_rows(500, seed=SEED + 7)
    idx.add(extra)
    idx.save_append(path, n)
    again = M.Index(768, dtype)
    again.set_metric(space)
    again.load_file(path)
    assert len(again) == n + 500 and again.norm_deviation() == idx.norm_deviation()
    assert np.array_equal(again.get_norms(), np.concatenate([nr, O.row_norms(extra, 768)]))
    assert_bit_equal(*again.search(q[:5], 10), *idx.search(q[:5], 10))
    # rows stored without their norms cannot be ranked in l2 / ip afterwards, and a row file without norms does not open
    plain = M.Index(768, dtype)
    plain.add(x[:100])
    with pytest.raises(M.MqvsError):
        plain.set_metric(space)
    plain.save(str(tmp_path / "plain.mqvs"))
    fresh = M.Index(768, dtype)
    fresh.set_metric(space)
    with pytest.raises(M.MqvsError):
        fresh.load_file(str(tmp_path / "plain.mqvs"))
    with pytest.raises(M.MqvsError):
        plain.get_norms()


def test_vectorstore_default_space_on_raw_vectors_ranks_by_squared_l2(gpu, tmp_path):
    """The reference opens Chroma without collection_metadata (src/medical_engine.py:52), i.e. in chromadb's default "l2"
    space.  With an embedder that does NOT normalise, the drop-in must rank by the squared L2 distance of the raw vectors
    and report it; with unit vectors (Ollama) it stays on the cosine kernels and reports 2 - 2s."""
    from conftest import HashEmbedder
    from mediquery_rag_b200 import MediQueryVectorStore
    texts = [f"text {r}" for r in range(3000)]
    raw, unit = HashEmbedder(unit=False), HashEmbedder()
    db = str(tmp_path / "db")
    vs = MediQueryVectorStore(embedding_function=raw, persist_directory=db)
    x = np.asarray(raw.embed_documents(texts), np.float32)
    x *= np.random.default_rng(5).uniform(0.5, 1.5, (len(texts), 1)).astype(np.float32)
    vs.add_embeddings((texts, x))
    assert vs.rank_space() == "l2"
    for store in (vs, MediQueryVectorStore(embedding_function=raw, persist_directory=db)):
        for query in ("text 5 again", "anything"):
            qv = np.asarray(raw.embed_query(query), np.float64)
            d = ((qv[None, :] - x.astype(np.float64)) ** 2).sum(1)
            got = store.similarity_search_with_score(query, k=6)
            assert [doc.page_content for doc, _ in got] == [texts[r] for r in np.argsort(d, kind="stable")[:6]]
            assert np.allclose([dist for _, dist in got], np.sort(d)[:6], rtol=1e-5)
    vu = MediQueryVectorStore(embedding_function=unit)
    vu.add_texts(texts)
    assert vu.rank_space() == "cosine" and vu._index.norm_deviation() <= 1e-6
    xu = np.asarray(unit.embed_documents(texts), np.float64)
    d = ((np.asarray(unit.embed_query("text 5 again"), np.float64)[None, :] - xu) ** 2).sum(1)
    got = vu.similarity_search_with_score("text 5 again", k=6)
    assert [doc.page_content for doc, _ in got] == [texts[r] for r in np.argsort(d, kind="stable")[:6]]
    assert np.allclose([dist for _, dist in got], np.sort(d)[:6], atol=2e-6)
    assert vu._index.stats()["last_pipeline"] == 1  # the single-launch latency kernel, as before


@pytest.mark.parametrize("space,dtype,B", [("l2", "bf16", 64), ("ip", "fp32", 20), ("l2", "fp32", 200), ("ip", "bf16", 300)])
def test_spaces_on_the_tensor_pipeline_band_pass_and_exact_repair(gpu, space, dtype, B):
    """The hard cases of the tensor path in the l2 / ip spaces: a clustered corpus of NON-unit rows (members of a cluster
    share a direction within cosine 0.98 and a length within 2 %), where ranking scores around rank k are closer together
    than amax * eps, so queries are flagged and the band pass must settle them in the transformed space; plus hundreds of
    exact duplicates of one row, which overflow the shortlist and force the exact-scan repair (scan_fix_kernel with the
    space transform).  Bit-identical to the oracle in every case."""
    n, members, k = 120_000, 1500, 10
    rng = np.random.default_rng(1)
    sigma = np.sqrt((1.0 / 0.98 ** 2 - 1.0) / 256.0)
    centres = rng.standard_normal((n // members, 768)).astype(np.float32)
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    length = np.repeat(rng.uniform(0.5, 3.0, n // members), members) * rng.uniform(0.98, 1.02, n)
    x = ((np.repeat(centres, members, axis=0) + sigma * rng.uniform(-1, 1, (n, 768))) * length[:, None]).astype(np.float32)
    x[60_000:61_500] = x[777]                        # 1500 exact copies: more than the direct shortlist holds
    picks = rng.integers(0, n, B)
    picks[0] = 777
    q = (x[picks] * rng.uniform(0.8, 1.2, (B, 1)) + 0.5 * sigma * rng.uniform(-1, 1, (B, 768))).astype(np.float32)
    q[0] = x[777]
    st, _ = O.normalize_rows(x, dtype, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    nr = O.row_norms(x, 768)
    es, ei = O.brute_force_topk(qn, st, k, metric=space, norms=nr, q_norms=O.row_norms(q, 768))
    idx = native().Index(768, dtype)
    idx.set_metric(space)
    idx.add(x)
    idx.set_path(2)
    s, i = idx.search(q, k)
    stats = idx.stats()
    assert_bit_equal(s, i, es, ei)
    assert stats["last_path"] == 1 and stats["last_pipeline"] == (3 if B <= 256 else 2), stats
    assert stats["band_repairs"] > 0, stats
    # the all-device call (band pass + device-driven exact repair enqueued blindly) gives the same answer
    import torch
    out_s = torch.empty((B, k), dtype=torch.float32, device="cuda")
    out_i = torch.empty((B, k), dtype=torch.int64, device="cuda")
    idx.search_device(torch.from_numpy(q).cuda(), k, out_s, out_i, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert idx.stats()["unrepaired"] == 0
    assert_bit_equal(out_s.cpu().numpy(), out_i.cpu().numpy(), es, ei)
    # the duplicates: query 0 is (a multiple of) row 777; in the l2 space its copies tie at the top, lowest row ids first
    if space == "l2":
        assert i[0, 0] == 777 and list(i[0, 1:4]) == [60_000, 60_001, 60_002]
    idx.set_path(4)                                  # segmented pipeline without the band pass: exact scans repair instead
    s4, i4 = idx.search(q[:12], k)
    assert_bit_equal(s4, i4, es[:12], ei[:12])
    assert idx.stats()["last_pipeline"] == 2 and idx.stats()["exact_fallbacks"] > 0


@pytest.mark.parametrize("space", ["l2", "ip"])
def test_full_size_config2_in_the_l2_and_ip_spaces_properties(gpu, space):
    """BASELINE config 2 size (1M x 768 fp32) in Chroma's l2 / ip spaces over the raw synthetic rows (norms 15.5-16.5),
    through size-independent properties: both kernel families agree bit for bit (single-launch tensor pipeline for the batch,
    exact SIMT scan for one query), results are sorted and duplicate-free, the returned scores are the oracle's ranking
    scores of the returned rows (rows regenerated from the hash), and no row of a 100k-row sample beats the k-th result."""
    n, B, k = 1_000_000, 16, 10
    idx = native().Index(768, "fp32", capacity_rows=n)
    idx.set_metric(space)
    idx.add_synthetic(SEED, 0, n)
    q = (O.synth_rows(QSEED, 0, B, 768) * np.float32(0.9) + 0.3 * O.synth_rows(SEED, 4242, B, 768)).astype(np.float32)
    s, i = idx.search(q, k)
    assert idx.stats()["last_pipeline"] == 3
    assert np.all(np.diff(s, axis=1) <= 0) and (i >= 0).all() and all(len(set(r)) == k for r in i.tolist())
    s1, i1 = idx.search(q[3:4], k)
    assert idx.stats()["last_pipeline"] == 0
    assert np.array_equal(i1[0], i[3]) and np.array_equal(s1[0].view(np.uint32), s[3].view(np.uint32))
    qn, _ = O.normalize_rows(q, "fp32", 768)
    qnorm = O.row_norms(q, 768)
    for b in range(0, B, 3):
        rows = np.concatenate([O.synth_rows(SEED, int(r), 1, 768) for r in i[b]])
        st_b, _ = O.normalize_rows(rows, "fp32", 768)
        es, ei = O.brute_force_topk(qn[b:b + 1], st_b, k, metric=space, norms=O.row_norms(rows, 768), q_norms=qnorm[b:b + 1])
        assert np.array_equal(ei[0], np.arange(k)) and np.array_equal(es[0].view(np.uint32), s[b].view(np.uint32))
    lo = 300_000
    sample = O.synth_rows(SEED, lo, 100_000, 768)
    st_s, _ = O.normalize_rows(sample, "fp32", 768)
    ss, si = O.brute_force_topk(qn, st_s, 1, metric=space, norms=O.row_norms(sample, 768), q_norms=qnorm, id_base=lo)
    for b in range(B):  # the sample's best row is either in the result or no better than its k-th entry
        assert si[b, 0] in i[b] or ss[b, 0] <= s[b, k - 1]
    assert np.array_equal(idx.get_norms(lo, 1000).view(np.uint32), O.row_norms(sample[:1000], 768).view(np.uint32))


# ---------------------------------------------------------------------------------------------- one handle, several shards
MULTI_CASES = [
    # W, n, dtype, B, k, mask fraction
    (2, 20000, "fp32", 9, 10, None),       # SIMT scans in every shard
    (3, 9000, "bf16", 3, 5, 0.3),          # latency kernel per shard, a shard with a partial block, mask
    (8, 5000, "fp32", 4, 10, None),        # more shards than blocks: empty shards
    (2, 140001, "bf16", 64, 10, None),     # single-launch tensor pipeline per shard (half-height query tiles)
    (3, 250000, "fp32", 130, 32, 0.2),     # two query tiles, mask split on block boundaries
    (4, 300000, "bf16", 300, 10, None),    # segmented pipeline per shard
    (2, 70000, "bf16", 5, 200, None),      # k > 128
]


@pytest.mark.parametrize("W,n,dtype,B,k,frac", MULTI_CASES)
def test_multi_device_handle_equals_single_shard_handle_and_oracle(gpu, W, n, dtype, B, k, frac):
    """mqvs_create(n_dev = W): one handle over W block-cyclic row shards (all on this one GPU here; one per GPU on a
    box).  Global ids, scores and tie order must be bit-identical to a single-shard handle over the same rows and to
    the oracle -- through host calls, device calls (mqvs_search_async), get_rows and a device-built filter."""
    import torch
    M = native()
    x = O.synth_rows(0x5EED0007, 0, n, 768)
    x[n - 7] = x[11]  # a tie that straddles shards: the lower GLOBAL id must win
    multi = M.Index(768, dtype, device=[0] * W)
    for a in range(0, n, 33333):  # appended in pieces that end inside blocks
        assert multi.add(x[a:a + 33333]) == a
    assert len(multi) == n
    st, _ = O.normalize_rows(x, dtype, 768)
    q, planted = O.planted_queries(QSEED, B, 0x5EED0007, n, 768)
    q[0] = x[11]
    qn, _ = O.normalize_rows(q, "fp32", 768)
    mask = None if frac is None else O.mask_from_bool(np.random.default_rng(5).random(n) < frac)
    es, ei = O.brute_force_topk(qn, st, k, mask)
    s, i = multi.search(q, k, mask)
    assert_bit_equal(s, i, es, ei)
    if frac is None:
        assert i[0, 0] == 11 and i[0, 1] == n - 7
    # device buffers on the home device, asynchronous
    qd = torch.from_numpy(q).cuda()
    out_s = torch.empty((B, k), dtype=torch.float32, device="cuda")
    out_i = torch.empty((B, k), dtype=torch.int64, device="cuda")
    if mask is None:
        multi.search_device(qd, k, out_s, out_i, id_base=1000, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        got_i = out_i.cpu().numpy()
        assert_bit_equal(out_s.cpu().numpy(), np.where(got_i >= 0, got_i - 1000, got_i), es, ei)
    rows = np.array([0, 4095, 4096, n - 1, n // 2, 4096 * W + 17 if 4096 * W + 17 < n else 1])
    assert np.array_equal(multi.get_rows(rows).view(np.uint32), O.stored_as_f32(st)[rows].view(np.uint32))
    st8 = multi.stats()
    assert st8["rows"] == n and st8["unrepaired"] == 0
    # Chroma `where` evaluated on every shard's device; the token goes back into search
    from mediquery_rag_b200.filters import DeviceFilter, MetadataColumns, evaluate
    cols = MetadataColumns()
    cols.extend_columns(n, {"source_id": (np.arange(n) * 7919) % 23, "score": np.linspace(0.0, 1.0, n)})
    live = np.ones(n, bool)
    live[[3, 4096, n - 1]] = False
    where = {"$and": [{"source_id": {"$in": [1, 5, 9, 22]}}, {"score": {"$gte": 0.25}}]}
    dmask, allowed = DeviceFilter(multi, cols).build(where, live)
    ok = evaluate(where, cols) & live
    assert allowed == int(ok.sum())
    s, i = multi.search(q, min(k, 128), dmask)
    assert_bit_equal(s, i, *O.brute_force_topk(qn, st, min(k, 128), O.mask_from_bool(ok)))
    multi.close()


def test_multi_device_handle_persistence_is_device_count_independent(gpu, tmp_path):
    """A row file written by a 3-shard handle opens in a 1-shard and a 2-shard handle (and the other way round), appends
    included; l2-space norms travel with it."""
    M = native()
    n = 30000
    x = O.synth_rows(0x5EED0008, 0, n, 768) * np.linspace(0.5, 2.0, n, dtype=np.float32)[:, None]
    q = x[[5, 9000, 29999]] * np.float32(0.9)
    path = str(tmp_path / "rows.mqvs")
    a = M.Index(768, "bf16", device=[0, 0, 0])
    a.set_metric("l2")
    a.add(x[:20000])
    a.save(path)
    a.add(x[20000:])
    a.save_append(path, 20000)
    want = a.search(q, 7)
    es, ei = O.brute_force_topk(O.normalize_rows(q, "fp32", 768)[0], O.normalize_rows(x, "bf16", 768)[0], 7, metric="l2",
                                norms=O.row_norms(x, 768), q_norms=O.row_norms(q, 768))
    assert_bit_equal(want[0], want[1], es, ei)
    for devs in (0, [0, 0]):
        b = M.Index(768, "bf16", device=devs)
        b.set_metric("l2")
        b.load_file(path)
        assert len(b) == n and abs(b.norm_deviation() - a.norm_deviation()) < 1e-6
        got = b.search(q, 7)
        assert_bit_equal(got[0], got[1], es, ei)
        assert np.array_equal(b.get_norms(4090, 20), a.get_norms(4090, 20))
        b.close()
    single = M.Index(768, "bf16")
    single.set_metric("l2")
    single.add(x)
    p2 = str(tmp_path / "single.mqvs")
    single.save(p2)
    c = M.Index(768, "bf16", device=[0, 0, 0, 0])
    c.set_metric("l2")
    c.load_file(p2)
    got = c.search(q, 7)
    assert_bit_equal(got[0], got[1], es, ei)
    bad = x[:3].copy()
    bad[1, 0] = np.inf
    with pytest.raises(M.MqvsError):
        c.add(bad)          # rejected by one shard: nothing is appended anywhere
    assert len(c) == n
    for h in (a, single, c):
        h.close()


def test_vectorstore_over_several_shards_equals_the_unsharded_store(gpu, tmp_path, embedder):
    """MediQueryVectorStore(devices=[...]): the object the reference constructs (src/medical_engine.py:52) over a
    sharded corpus -- documents, ids, filters from metadata, tombstones and persistence as on one shard."""
    from mediquery_rag_b200 import MediQueryVectorStore
    n = 12000
    x = O.synth_rows(0x5EED0009, 0, n, 768)
    texts = [f"chunk {r}" for r in range(n)]
    metas = [{"title": f"t{r}", "group": r % 7, "tags": "a，b" if r % 2 else "c"} for r in range(n)]
    stores = {}
    for name, kw in (("one", dict(device=0)), ("many", dict(devices=[0, 0, 0]))):
        vs = MediQueryVectorStore(persist_directory=str(tmp_path / name), embedding_function=embedder, **kw)
        ids = vs.add_embeddings((texts, x), metadatas=metas, ids=[f"id{r}" for r in range(n)])
        vs.delete(ids[100:140])
        vs.add_embeddings(([f"again {r}" for r in range(50)], x[200:250]), ids=ids[200:250])  # upsert: old rows die
        stores[name] = vs
    q = x[220] + 0.01 * x[5]
    for filt in (None, {"group": {"$in": [2, 3]}}, {"$and": [{"group": {"$ne": 1}}, {"title": {"$gte": "t5"}}]}):
        a = stores["one"].similarity_search_with_score("", k=6, filter=filt) if False else stores["one"]._similar(q, 6, filt)
        b = stores["many"]._similar(q, 6, filt)
        assert [(d.page_content, d.metadata, d.id) for d, _ in a] == [(d.page_content, d.metadata, d.id) for d, _ in b]
        assert [s for _, s in a] == [s for _, s in b]
    assert stores["many"]._similar(q, 1, None)[0][0].page_content == "again 20"
    again = MediQueryVectorStore(persist_directory=str(tmp_path / "many"), embedding_function=embedder, devices=[0, 0])
    assert len(again) == len(stores["one"]) == n - 40
    assert [d.id for d, _ in again._similar(q, 6, {"group": 3})] == [d.id for d, _ in stores["one"]._similar(q, 6, {"group": 3})]
    # bulk synthetic rows without a Python object per row, through the same object
    big = MediQueryVectorStore(devices=[0, 0], store_dtype="bf16", collection_metadata={"hnsw:space": "cosine"})
    big.add_synthetic(0x5EED0004, 100000, metadata_columns={"source_id": np.arange(100000) % 10})
    qs, planted = O.planted_queries(QSEED, 8, 0x5EED0004, 100000, 768)
    sc, rows = big.search_batch(qs, 10, filter={"source_id": {"$lt": 5}})
    st, _ = O.normalize_rows(O.synth_rows(0x5EED0004, 0, 100000, 768), "bf16", 768)
    es, ei = O.brute_force_topk(O.normalize_rows(qs, "fp32", 768)[0], st, 10, O.mask_from_bool(np.arange(100000) % 10 < 5))
    assert_bit_equal(sc, rows, es, ei)
    doc = big.similarity_search_by_vector(qs[0], k=1)[0]
    assert doc.id == f"row-{planted[0]}" and doc.metadata == {"source_id": int(planted[0] % 10)}


def test_filtered_searches_from_several_threads_do_not_share_a_mask(gpu, embedder):
    """The handle's device-built filter mask is one buffer: the store holds its lock across build + search, so concurrent
    filtered searches (the a* coroutines run in worker threads) each see their own filter."""
    import threading
    from mediquery_rag_b200 import MediQueryVectorStore
    n = 30000
    x = O.synth_rows(0x5EED000A, 0, n, 768)
    vs = MediQueryVectorStore()
    vs.add_embeddings(([None] * n, x), metadata_columns={"g": np.arange(n) % 5})
    q = O.synth_rows(99, 0, 4, 768)
    want = {g: vs.search_batch(q, 10, filter={"g": g})[1] for g in range(5)}
    errors = []

    def worker(g):
        for _ in range(20):
            rows = vs.search_batch(q, 10, filter={"g": g})[1]
            if not np.array_equal(rows, want[g]) or not ((rows % 5) == g).all():
                errors.append(g)

    th = [threading.Thread(target=worker, args=(g,)) for g in range(5)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors


def test_tombstone_compaction_keeps_rows_bit_for_bit(gpu, tmp_path):
    from mediquery_rag_b200 import MediQueryVectorStore
    M = native()
    n = 40000
    x = O.synth_rows(0x5EED000B, 0, n, 768) * np.float32(1.7)
    for dtype in ("fp32", "bf16"):
        idx = M.Index(768, dtype)
        idx.set_metric("l2")
        idx.add(x)
        keep = np.random.default_rng(3).random(n) < 0.4
        keep[:5000] = True          # a long prefix that stays where it is
        before = idx.get_rows(np.flatnonzero(keep))
        norms = idx.get_norms()[keep]
        assert idx.compact(O.mask_from_bool(keep)) == int(keep.sum()) == len(idx)
        assert np.array_equal(idx.get_rows(np.arange(len(idx))).view(np.uint32), before.view(np.uint32))
        assert np.array_equal(idx.get_norms(), norms)
        q = x[np.flatnonzero(keep)[[7, 9000]]]
        s, i = idx.search(q, 5)
        xs = x[keep]
        es, ei = O.brute_force_topk(O.normalize_rows(q, "fp32", 768)[0], O.normalize_rows(xs, dtype, 768)[0], 5, metric="l2",
                                    norms=O.row_norms(xs, 768), q_norms=O.row_norms(q, 768))
        assert_bit_equal(s, i, es, ei)
        idx.close()
    # the store compacts by itself once most rows are dead, and the directory reopens
    vs = MediQueryVectorStore(persist_directory=str(tmp_path / "db"), collection_metadata={"hnsw:space": "cosine"})
    ids = vs.add_embeddings(([f"c{r}" for r in range(3000)], x[:3000]))
    vs.delete(ids[:2000])
    assert len(vs._ids) == 1000 and len(vs) == 1000 and len(vs._index) == 1000
    assert vs.similarity_search_by_vector(x[2500], k=1)[0].page_content == "c2500"
    re = MediQueryVectorStore(persist_directory=str(tmp_path / "db"), collection_metadata={"hnsw:space": "cosine"})
    assert len(re) == 1000 and re.similarity_search_by_vector(x[2999], k=1)[0].id == ids[2999]


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_host_ingest_pipeline_is_bit_exact_over_many_chunks(gpu, dtype):
    """mqvs_add from host memory streams 8192-row chunks through two staging slots (pageable rows bounced through pinned
    buffers by host threads, pinned rows copied from where they lie): same stored bits as the oracle's K1 either way."""
    import torch
    n = 30000
    x = O.synth_rows(0x5EED000C, 0, n, 768) * np.float32(2.5)
    st, _ = O.normalize_rows(x, dtype, 768)
    want = O.stored_as_f32(st)
    for src in (x, torch.from_numpy(x).pin_memory().numpy()):
        idx = native().Index(768, dtype)
        idx.add(src[:100])
        idx.add(src[100:])
        probe = np.r_[0:50, 8150:8250, 16380:16400, n - 30:n]
        assert np.array_equal(idx.get_rows(probe).view(np.uint32), want[probe].view(np.uint32))
        idx.close()


WIDE_K_CASES = [
    # n, dtype, B, k, mask fraction: k > 128 on the single-launch tensor pipeline (wide shortlists, k' = 512 / 1152)
    (100003, "bf16", 64, 200, None),
    (100003, "fp32", 20, 1000, None),
    (300000, "bf16", 130, 256, 0.3),
    (70000, "fp32", 1, 129, None),
    (70000, "bf16", 300, 150, None),     # more than 256 queries: 256 at a time
    (40000, "bf16", 8, 1024, 0.02),      # about 800 allowed rows: fewer than k, padded
]


@pytest.mark.parametrize("n,dtype,B,k,frac", WIDE_K_CASES)
def test_wide_k_on_the_tensor_pipeline_matches_oracle_bitwise(gpu, n, dtype, B, k, frac):
    idx, x, st = build(n, 768, dtype, seed=0x5EED000D)
    idx.set_path(2)
    q, planted = O.planted_queries(QSEED, B, 0x5EED000D, n, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    mask = None if frac is None else O.mask_from_bool(np.random.default_rng(13).random(n) < frac)
    s, i = idx.search(q, k, mask)
    stt = idx.stats()
    assert stt["last_path"] == 1 and stt["last_pipeline"] == 3
    assert_bit_equal(s, i, *O.brute_force_topk(qn, st, k, mask))
    idx.close()


def test_wide_k_with_thousands_of_ties_falls_back_and_stays_exact(gpu):
    """6000 exact copies of the best row against k = 300: the wide per-CTA lists (2048 entries, no in-kernel compaction) fill up,
    the query is flagged and the repair chain answers it -- same ids (lowest row ids of the tie) and bits as the oracle; the
    asynchronous call does the same on its own."""
    import torch
    n, k = 60000, 300
    x = O.synth_rows(SEED, 0, n, 768)
    x[20000:26000] = x[77]
    idx = native().Index(768, "bf16")
    idx.add(x)
    idx.set_path(2)
    q = np.stack([x[77] + 0.01 * O.synth_rows(5, 0, 1, 768)[0], x[40000]])
    qn, _ = O.normalize_rows(q, "fp32", 768)
    st, _ = O.normalize_rows(x, "bf16", 768)
    es, ei = O.brute_force_topk(qn, st, k)
    s, i = idx.search(q, k)
    assert_bit_equal(s, i, es, ei)
    out_s = torch.empty((2, k), dtype=torch.float32, device="cuda")
    out_i = torch.empty((2, k), dtype=torch.int64, device="cuda")
    idx.search_device(torch.from_numpy(q).cuda(), k, out_s, out_i, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert_bit_equal(out_s.cpu().numpy(), out_i.cpu().numpy(), es, ei)
    assert idx.stats()["unrepaired"] == 0
    idx.close()


@pytest.mark.parametrize("devices", [0, [0, 0, 0]])
def test_tag_token_filter_on_the_device_equals_the_host_evaluator(gpu, medical_docs, devices):
    """f4: the comma-joined `tags` of the reference's 154 chunks as a token-set column on the device
    (mqvs_column_set_tokens + MQVS_OP_CONTAINS / NOT_CONTAINS in filter_eval_kernel), one shard and three; the searches
    through the resulting masks equal the oracle under the host evaluator's mask."""
    from mediquery_rag_b200.filters import DeviceFilter, MetadataColumns, evaluate
    metas = [d["metadata"] for d in medical_docs["docs"]] * 60   # 9240 rows: three blocks of a sharded handle
    n = len(metas)
    cols = MetadataColumns()
    cols.extend(metas)
    x = O.synth_rows(0x5EED000E, 0, n, 768)
    idx = native().Index(768, "fp32", device=devices)
    idx.add(x)
    st, _ = O.normalize_rows(x, "fp32", 768)
    q = O.synth_rows(QSEED, 0, 3, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    dev = DeviceFilter(idx, cols)
    live = np.ones(n, bool)
    live[::11] = False
    for where in ({"tags": {"$contains": "癌症"}}, {"tags": {"$not_contains": "运动"}}, {"tags": {"$contains": "no such tag"}},
                  {"$and": [{"tags": {"$contains": "代谢性疾病"}}, {"tags": {"$not_contains": "糖尿病"}}]},
                  {"$or": [{"tags": {"$contains": "睡眠"}}, {"title": {"$lt": "什"}}]}):
        ok = evaluate(where, cols) & live
        mask, allowed = dev.build(where, live)
        assert allowed == int(ok.sum())
        s, i = idx.search(q, 5, mask)
        assert_bit_equal(s, i, *O.brute_force_topk(qn, st, 5, O.mask_from_bool(ok)))
    idx.close()


@pytest.mark.parametrize("devices", [0, [0, 0, 0]])
def test_where_document_on_the_device_equals_the_host_evaluator(gpu, medical_docs, devices):
    """f4: Chroma's where_document ($contains / $not_contains over the stored documents, $and / $or) evaluated by
    filter_eval_kernel over a CSR text column (mqvs_column_set_text: the reference's real page_content strings, UTF-8), one
    shard and three (the text is split on block boundaries with the rows); masks and searches equal the host evaluator's,
    alone and AND-ed with a metadata filter and tombstones.  Then the same through the object the reference holds."""
    from mediquery_rag_b200 import MediQueryVectorStore
    from mediquery_rag_b200.filters import DeviceFilter, MetadataColumns, compile_mask, evaluate, evaluate_document
    base_t = [d["page_content"] for d in medical_docs["docs"]]
    base_m = [d["metadata"] for d in medical_docs["docs"]]
    texts = (base_t + [None, "", "血"]) * 58            # 9106 rows: three blocks of a sharded handle, rows without documents
    metas = (base_m + [None, {"tags": "x"}, None]) * 58
    n = len(texts)
    cols = MetadataColumns()
    cols.extend(metas)
    x = O.synth_rows(0x5EED000F, 0, n, 768)
    idx = native().Index(768, "fp32", device=devices)
    idx.add(x)
    st, _ = O.normalize_rows(x, "fp32", 768)
    q = O.synth_rows(QSEED, 0, 3, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    dev = DeviceFilter(idx, cols, texts)
    live = np.ones(n, bool)
    live[::11] = False
    long_needle = base_t[7][3:120]  # a needle longer than a warp's stride of start positions
    for wd in ({"$contains": "血压"}, {"$not_contains": "血压"}, {"$contains": "问题："}, {"$contains": "no such text"},
               {"$contains": long_needle}, {"$contains": base_t[3][-9:]}, {"$contains": base_t[5][:4]},
               {"$and": [{"$contains": "运动"}, {"$not_contains": "癌"}]},
               {"$or": [{"$contains": "睡眠"}, {"$contains": "饮食"}, {"$and": [{"$contains": "糖"}, {"$not_contains": "胰岛素"}]}]}):
        for where in (None, {"tags": {"$not_contains": "运动"}}):
            ok = evaluate_document(wd, texts) & live
            if where:
                ok &= evaluate(where, cols)
            mask, allowed = dev.build(where, live, wd)
            assert allowed == int(ok.sum()), (wd, where)
            s, i = idx.search(q, 5, mask)
            assert_bit_equal(s, i, *O.brute_force_topk(qn, st, 5, O.mask_from_bool(ok)))
    assert 0 < int(evaluate_document({"$contains": long_needle}, texts).sum()) <= 58
    idx.close()

    kw = dict(device=0) if devices == 0 else dict(devices=devices)
    vs = MediQueryVectorStore(collection_metadata={"hnsw:space": "cosine"}, **kw)
    ids = vs.add_embeddings((texts, x), metadatas=metas)
    vs.delete(ids[::11])
    built = []
    orig = DeviceFilter.build
    try:
        DeviceFilter.build = lambda self, *a, **k: (built.append(a), orig(self, *a, **k))[1]
        for wd, filt in (({"$contains": "血压"}, None), ({"$not_contains": "运动"}, {"tags": {"$contains": "癌症"}})):
            got = vs._similar(q[0], 6, filt, wd)
            host_mask, host_allowed = compile_mask(filt, cols, live, wd, texts)
            es, ei = O.brute_force_topk(qn[:1], st, 6, host_mask)
            assert [d.id for d, _ in got] == [ids[r] for r in ei[0]]
            assert all(("血压" in d.page_content) if "$contains" in wd else ("运动" not in d.page_content) for d, _ in got)
    finally:
        DeviceFilter.build = orig
    assert len(built) == 2  # both filters went through the device program (no host fallback)


FP32_CASES = [c for c in TC_CASES if c[1] == "fp32"] + [(300000, "fp32", 300, 10, None), (120000, "fp32", 1, 10, None)]


@pytest.mark.parametrize("n,dtype,B,k,frac", FP32_CASES)
def test_fp32_store_returns_the_same_bits_with_and_without_its_scoring_shadow(gpu, n, dtype, B, k, frac):
    """fp32 stores: the tensor pipelines stream a bf16 SHADOW of the rows by default (half the bytes, kind::f16) and take
    every exact score from the fp32 rows; with the shadow off they stream the fp32 rows themselves (kind::tf32) as in
    round 1.  Both legs, on both tensor pipelines, must be the oracle's bits -- and stats.scan_bytes says which
    representation was streamed."""
    idx, x, st = build(n, 768, "fp32", seed=0x5EED0003)
    q, _ = O.planted_queries(QSEED, B, 0x5EED0003, n, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    mask = None if frac is None else O.mask_from_bool(np.random.default_rng(11).random(n) < frac)
    want = O.brute_force_topk(qn, st, k, mask)
    for path in (2, 4):
        idx.set_path(path)
        for shadow in (True, False, True):
            idx.set_shadow(shadow)
            before = idx.stats()["scan_bytes"]
            s, i = idx.search(q, k, mask)
            assert idx.stats()["last_path"] == 1
            assert_bit_equal(s, i, *want)
            assert idx.stats()["scan_bytes"] - before == n * 768 * (2 if shadow else 4), (path, shadow)
            assert (idx.shadow_bytes() > 0) == shadow
    idx.close()


def test_scoring_shadow_on_hard_corpora_and_across_growth_compaction_and_reload(gpu, tmp_path):
    """The shadow's error bound (both operands rounded) through the repair chain -- a corpus made of ties, a clustered
    corpus of non-unit rows in the l2 space -- and its upkeep: rows added later are converted before the next tensor
    search, compaction and mqvs_load rebuild it, a sharded handle keeps one per shard."""
    M = native()
    # (1) 40 distinct vectors: thresholds tie with thousands of rows, lists overflow, margins cannot be proven
    n, B, k = 100_000, 24, 10
    rng = np.random.default_rng(7)
    base = O.synth_rows(SEED, 0, 40, 768)
    x = base[rng.integers(0, 40, n)]
    q = base[rng.integers(0, 40, B)] + 0.05 * O.synth_rows(QSEED, 0, B, 768)
    st, _ = O.normalize_rows(x, "fp32", 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    want = O.brute_force_topk(qn, st, k)
    idx = M.Index(768, "fp32")
    idx.add(x)
    for path in (2, 4):
        idx.set_path(path)
        assert_bit_equal(*idx.search(q, k), *want)
    assert idx.shadow_bytes() >= n * 1536 and idx.stats()["unrepaired"] == 0
    idx.close()
    # (2) clustered non-unit rows, l2 space: flagged queries go through the band pass over the shadow
    n, members, B = 120_000, 1500, 40
    rng = np.random.default_rng(1)
    sigma = np.sqrt((1.0 / 0.98 ** 2 - 1.0) / 256.0)
    centres = rng.standard_normal((n // members, 768)).astype(np.float32)
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    length = np.repeat(rng.uniform(0.5, 3.0, n // members), members) * rng.uniform(0.98, 1.02, n)
    x = ((np.repeat(centres, members, axis=0) + sigma * rng.uniform(-1, 1, (n, 768))) * length[:, None]).astype(np.float32)
    q = (x[rng.integers(0, n, B)] + 0.5 * sigma * rng.uniform(-1, 1, (B, 768))).astype(np.float32)
    st, _ = O.normalize_rows(x, "fp32", 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    want = O.brute_force_topk(qn, st, k, metric="l2", norms=O.row_norms(x, 768), q_norms=O.row_norms(q, 768))
    for devices in (0, [0, 0, 0]):
        idx = M.Index(768, "fp32", device=devices)
        idx.set_metric("l2")
        idx.add(x)
        idx.set_path(2)
        assert_bit_equal(*idx.search(q, k), *want)
        assert idx.stats()["band_repairs"] > 0 and idx.shadow_bytes() >= n * 1536
        idx.close()
    # (3) upkeep: incremental adds with regrowth, compaction, save + load
    n = 90_000
    x = O.synth_rows(0x5EED0010, 0, n, 768)
    st, _ = O.normalize_rows(x, "fp32", 768)
    q, _ = O.planted_queries(QSEED, 8, 0x5EED0010, n, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    idx = M.Index(768, "fp32")
    idx.set_path(2)
    for lo, hi in ((0, 30_000), (30_000, 31_000), (31_000, n)):
        idx.add(x[lo:hi])
        assert_bit_equal(*idx.search(q, 10), *O.brute_force_topk(qn, st[:hi], 10))
    keep = np.random.default_rng(5).random(n) < 0.5
    assert idx.compact(O.mask_from_bool(keep)) == int(keep.sum())
    assert_bit_equal(*idx.search(q, 10), *O.brute_force_topk(qn, st[keep], 10))
    path = str(tmp_path / "rows.mqvs")
    idx.save(path)
    idx.close()
    again = M.Index(768, "fp32")
    again.load_file(path)
    again.set_path(2)
    before = again.stats()["scan_bytes"]
    assert_bit_equal(*again.search(q, 10), *O.brute_force_topk(qn, st[keep], 10))
    assert again.stats()["scan_bytes"] - before == int(keep.sum()) * 1536
    again.close()


@pytest.mark.parametrize("dim,dtype,n,B,k,frac", [(1024, "fp32", 30000, 1, 5, None), (1024, "bf16", 70000, 64, 10, None),
                                                  (900, "fp32", 100003, 130, 10, 0.3), (1024, "bf16", 150000, 300, 100, None),
                                                  (1000, "fp32", 5000, 3, 10, None), (1024, "fp32", 131072, 2, 10, None)])
def test_rows_wider_than_768_run_on_the_1024_build(gpu, dim, dtype, n, B, k, frac):
    """Embedding models wider than dmeta-embedding-zh (bge-large / mxbai: 1024): MQVS_DIM_PAD is a build-time constant, the
    package builds libmqvs_w1024.so from the same sources and the binding picks it by dimension.  Every pipeline (SIMT,
    latency kernel, single-launch tensor, segmented tensor on CTA pairs, the fp32 store's bf16 shadow) against the oracle
    restated at that width: same bits."""
    idx = native().Index(dim, dtype)
    assert idx.width == 1024
    x = O.synth_rows(0x5EED0011, 0, n, dim)
    idx.add(x)
    st, _ = O.normalize_rows(x, dtype, 1024)
    q, planted = O.planted_queries(QSEED, B, 0x5EED0011, n, dim)
    qn, _ = O.normalize_rows(q, "fp32", 1024)
    mask = None if frac is None else O.mask_from_bool(np.random.default_rng(11).random(n) < frac)
    want = O.brute_force_topk(qn, st, k, mask)
    for path in (0, 1, 3, 2, 4):
        idx.set_path(path)
        assert_bit_equal(*idx.search(q, k, mask), *want)
        if path in (2, 4):
            assert idx.stats()["last_path"] == 1
    assert np.array_equal(idx.get_rows(np.arange(5)).view(np.uint32), O.stored_as_f32(st[:5])[:, :dim].view(np.uint32))
    idx.close()


def test_vectorstore_with_a_1024_dimensional_embedder(gpu, tmp_path):
    from mediquery_rag_b200 import MediQueryVectorStore
    n, dim = 3000, 1024
    x = O.synth_rows(0x5EED0012, 0, n, dim)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    vs = MediQueryVectorStore(persist_directory=str(tmp_path / "db"), collection_metadata={"hnsw:space": "cosine"})
    ids = vs.add_embeddings(([f"doc {r}" for r in range(n)], x), metadatas=[{"g": r % 3} for r in range(n)])
    got = vs.similarity_search_by_vector(x[1234], k=3, filter={"g": 1234 % 3})
    assert got[0].page_content == "doc 1234" and vs._index.width == 1024
    again = MediQueryVectorStore(persist_directory=str(tmp_path / "db"), collection_metadata={"hnsw:space": "cosine"})
    assert len(again) == n and again._index.width == 1024
    assert [d.id for d in again.similarity_search_by_vector(x[77], k=4)] == [d.id for d in vs.similarity_search_by_vector(x[77], k=4)]
    assert again.similarity_search_by_vector(x[77], k=1)[0].id == ids[77]


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_kernels_against_the_independent_float64_referee_at_one_million_rows(gpu, dtype):
    """The parity claim without the kernel-ordered oracle in between: the CUDA results themselves (single-launch tensor
    pipeline and segmented pipeline at B = 16, SIMT scan at B = 1, all three spaces over raw non-unit vectors; the CTA-pair
    kernel at B = 160; fp32 stores through their bf16 scoring shadow) against float64 numpy over the stored values, under north_star's stated tolerance -- ids identical apart from near ties, scores within 1e-5 relative."""
    from referee import assert_within_north_star, f64_topk
    n, B, k = 1_000_000, 16, 10
    x = O.synth_rows(SEED, 0, n, 768)
    x *= np.linspace(0.03, 0.09, n, dtype=np.float32)[:, None]
    q, _ = O.planted_queries(QSEED, B, SEED, n, 768)
    q = (q * np.float32(1.3)).astype(np.float32)
    st, _ = O.normalize_rows(x, dtype, 768)
    norms = O.row_norms(x, 768)
    idx = native().Index(768, dtype, capacity_rows=n)
    idx.set_metric("l2")
    idx.add(x)
    del x
    for space in ("cosine", "l2", "ip"):
        idx.set_metric(space)
        ref_d, ref_i = f64_topk(q, st, norms, k, space)
        s, i = idx.search(q, k)
        assert idx.stats()["last_path"] == 1
        assert_within_north_star(space, s, i, q, ref_d, ref_i)
        s1, i1 = idx.search(q[:1], k)
        assert_within_north_star(space, s1, i1, q[:1], ref_d[:1], ref_i[:1])
        idx.set_path(4)                      # the segmented tensor pipeline (cold pass, radix selection, rescore kernel)
        s4, i4 = idx.search(q, k)
        assert idx.stats()["last_pipeline"] == 2
        assert_within_north_star(space, s4, i4, q, ref_d, ref_i)
        idx.set_path(0)
    # a batch of more than one query tile: the CTA-pair (cta_group::2) scoring kernel of the segmented pipeline, cosine space
    idx.set_metric("cosine")
    qb, _ = O.planted_queries(QSEED + 1, 160, SEED, n, 768)
    ref_d, ref_i = f64_topk(qb, st, norms, k, "cosine")
    idx.set_path(4)
    sb, ib = idx.search(qb, k)
    assert idx.stats()["last_pipeline"] == 2
    assert_within_north_star("cosine", sb, ib, qb, ref_d, ref_i)
    idx.close()


@pytest.mark.parametrize("order", ["scalar", "simd16"])
def test_kernels_against_the_published_arithmetic_of_chroma_hnswlib(gpu, order):
    """The CUDA results against what the reference's own dependency computes (oracle/hnswlib_restated.c: chroma-hnswlib's
    distance functions over the RAW fp32 vectors in its sequential / 16-lane summation order, its exhaustive search): an
    fp32 collection, all three spaces, SIMT scan (B = 1), single-launch tensor pipeline over the bf16 scoring shadow
    (cosine) and over the fp32 rows (l2 / ip), segmented pipeline -- ids identical apart from near ties, distances within
    north_star's 1e-5 relative."""
    from referee import assert_within_north_star
    n, B, k = 300_000, 8, 10
    x = O.synth_rows(SEED, 0, n, 768)
    x *= np.linspace(0.03, 0.09, n, dtype=np.float32)[:, None]
    q, _ = O.planted_queries(QSEED, B, SEED, n, 768)
    q = (q * np.float32(1.3)).astype(np.float32)
    idx = native().Index(768, "fp32", capacity_rows=n)
    idx.set_metric("l2")
    idx.add(x)
    for space in ("cosine", "l2", "ip"):
        idx.set_metric(space)
        ref_d, ref_i = O.hnswlib_topk(space, x, q, k, order=order)
        ref_d = ref_d.astype(np.float64)
        for path, want in ((0, None), (2, 1), (4, 1), (1, 0)):
            idx.set_path(path)
            s, i = idx.search(q, k)
            if want is not None:
                assert idx.stats()["last_path"] == want
            assert_within_north_star(space, s, i, q, ref_d, ref_i)
        idx.set_path(0)
        s1, i1 = idx.search(q[:1], k)
        assert_within_north_star(space, s1, i1, q[:1], ref_d[:1], ref_i[:1])
    idx.close()


@pytest.mark.parametrize("path,dups,clusters", [(2, 1500, False), (4, 300, False), (2, 0, True), (2, 0, False)])
def test_repeated_async_searches_with_alternating_buffers_and_a_growing_store_stay_exact(gpu, path, dups, clusters):
    """A steady stream of identical asynchronous searches (a serving loop, bench.py's `value`) whose device-side repair
    chain runs every time: same bits as the oracle whether it has work (duplicates defeat the shortlist -> exact-scan rounds;
    dense clusters -> band pass) or not (random rows), with one or two alternating output buffers (callers that allocate
    their outputs per call), and after rows were added."""
    import torch
    n, B, k = 20000, 40, 10
    x = O.synth_rows(SEED, 0, n, 768)
    if dups:
        x[5000:5000 + dups] = x[123]
    if clusters:  # 100-member clusters at cosine ~0.98: the margin fails, the band pass repairs
        centres = x[:200].copy()
        x = (centres[np.arange(n) % 200] + 0.012 * x).astype(np.float32)
    noise = O.synth_rows(77, 0, B, 768)
    q = (x[123][None, :] + 0.01 * noise).astype(np.float32)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    st, _ = O.normalize_rows(x, "bf16", 768)
    es, ei = O.brute_force_topk(qn, st, k)
    idx = native().Index(768, "bf16")
    idx.add(x)
    idx.set_path(path)
    stream = torch.cuda.current_stream().cuda_stream
    qd = torch.from_numpy(q).cuda()
    bufs = [(torch.empty((B, k), dtype=torch.float32, device="cuda"), torch.empty((B, k), dtype=torch.int64, device="cuda"))
            for _ in range(2)]
    for rep in range(7):
        out_s, out_i = bufs[rep % 2 if rep >= 3 else 0]
        out_s.fill_(7.0)
        out_i.fill_(-7)
        idx.search_device(qd, k, out_s, out_i, stream=stream)
        torch.cuda.synchronize()
        assert_bit_equal(out_s.cpu().numpy(), out_i.cpu().numpy(), es, ei)
    stt = idx.stats()
    assert stt["unrepaired"] == 0 and stt["last_path"] == 1
    if dups:
        assert stt["exact_fallbacks"] >= 7 * B
    if clusters:
        assert stt["band_repairs"] + stt["exact_fallbacks"] >= 7
    # the store changes: the captured chains describe the old corpus and must not run
    extra = O.synth_rows(SEED + 9, 0, 3000, 768)
    extra[17] = x[123] * 1.0000001
    idx.add(extra)
    st2, _ = O.normalize_rows(np.concatenate([x, extra]), "bf16", 768)
    es2, ei2 = O.brute_force_topk(qn, st2, k)
    for rep in range(3):
        out_s, out_i = bufs[0]
        idx.search_device(qd, k, out_s, out_i, stream=stream)
        torch.cuda.synchronize()
        assert_bit_equal(out_s.cpu().numpy(), out_i.cpu().numpy(), es2, ei2)
    assert idx.stats()["unrepaired"] == 0
    idx.close()


def test_timing_levels_of_the_search_events(gpu, monkeypatch):
    """stats.scan_ms (what bench.py's roofline divides by) comes from two timed CUDA events around the scan launches;
    MQVS_TIMING=1 adds the first-kernel-to-last bracket, MQVS_TIMING=0 records nothing.  Read when a handle is created;
    results do not depend on it."""
    n, B, k = 70_000, 8, 10
    q, _ = O.planted_queries(QSEED, B, SEED, n, 768)
    seen = {}
    for level in (None, "1", "0"):
        if level is None:
            monkeypatch.delenv("MQVS_TIMING", raising=False)
        else:
            monkeypatch.setenv("MQVS_TIMING", level)
        idx = native().Index(768, "bf16", capacity_rows=n)
        idx.add_synthetic(SEED, 0, n)
        for B_ in (B, 1):  # tensor pipeline, latency kernel
            s, i = idx.search(q[:B_], k)
            st = idx.stats()
            seen[(level, B_)] = (s.copy(), i.copy())
            if level == "0":
                assert st["scan_ms"] == 0.0 and st["last_search_ms"] == 0.0
            elif level == "1":
                assert st["last_search_ms"] >= st["last_scan_ms"] > 0.0
            else:
                assert st["last_search_ms"] == st["last_scan_ms"] > 0.0
        idx.close()
    for B_ in (B, 1):
        for level in ("1", "0"):
            assert_bit_equal(*seen[(level, B_)], *seen[(None, B_)])


@pytest.mark.parametrize("W,dtype", [(3, "fp32"), (4, "bf16"), (2, "fp32")])
def test_tombstone_compaction_of_a_multi_device_handle_equals_the_single_device_one(gpu, W, dtype, tmp_path):
    """mqvs_compact on a handle over W row shards: survivors keep their GLOBAL order and therefore change shard (block-cyclic
    placement of the new row numbers).  Rows, stored norms and search results must be the same bits as after compacting a
    single-device handle over the same rows -- prefix kept in place, deletions inside and across 4096-row blocks, a second
    compaction, appends afterwards (the parts must still be balanced and contiguous)."""
    from mediquery_rag_b200 import MediQueryVectorStore
    M = native()
    n = 50_000 + 17
    x = O.synth_rows(0x5EED000C, 0, n, 768) * np.float32(1.7)
    multi, single = M.Index(768, dtype, device=[0] * W), M.Index(768, dtype)
    for idx in (multi, single):
        idx.set_metric("l2")
        idx.add(x)
    rng = np.random.default_rng(5)
    keep = rng.random(n) < 0.45
    keep[:9000] = True                 # a prefix of more than two blocks that stays where it is
    keep[20_000:29_000] = False        # whole blocks disappear
    live = np.flatnonzero(keep)
    for idx in (multi, single):
        assert idx.compact(O.mask_from_bool(keep)) == live.size == len(idx)
    every = np.arange(live.size)
    assert np.array_equal(multi.get_rows(every).view(np.uint32), single.get_rows(every).view(np.uint32))
    assert np.array_equal(multi.get_norms(), single.get_norms())
    xs = x[keep]
    q = xs[[7, 9000, live.size - 1]] + np.float32(0.01) * O.synth_rows(3, 0, 3, 768)
    for B_ in (3, 1):
        s, i = multi.search(q[:B_], 5)
        s1, i1 = single.search(q[:B_], 5)
        assert_bit_equal(s, i, s1, i1)
        es, ei = O.brute_force_topk(O.normalize_rows(q[:B_], "fp32", 768)[0], O.normalize_rows(xs, dtype, 768)[0], 5, metric="l2",
                                    norms=O.row_norms(xs, 768), q_norms=O.row_norms(q[:B_], 768))
        assert_bit_equal(s, i, es, ei)
    # again (now nothing but a tail goes), then appends land behind the survivors on every part
    keep2 = np.ones(live.size, bool)
    keep2[-3000:] = rng.random(3000) < 0.5
    extra = O.synth_rows(0x5EED000D, 0, 6000, 768)
    for idx in (multi, single):
        assert idx.compact(O.mask_from_bool(keep2)) == int(keep2.sum())
        assert idx.add(extra) == int(keep2.sum())
    every = np.arange(len(single))
    assert len(multi) == len(single)
    assert np.array_equal(multi.get_rows(every).view(np.uint32), single.get_rows(every).view(np.uint32))
    qq = np.concatenate([extra[[5, 5999]], xs[keep2][[100]]])
    assert_bit_equal(*multi.search(qq, 4), *single.search(qq, 4))
    multi.close()
    single.close()
    # through the store: a collection over several shards compacts by itself once most rows are dead, persists, reopens on one device
    vs = MediQueryVectorStore(persist_directory=str(tmp_path / "db"), collection_metadata={"hnsw:space": "cosine"},
                              devices=[0] * W, store_dtype=dtype)
    ids = vs.add_embeddings(([f"c{r}" for r in range(12_000)], x[:12_000]), metadatas=[{"g": r % 5} for r in range(12_000)])
    vs.delete(ids[:9000])
    assert len(vs._ids) == 3000 and len(vs) == 3000 and len(vs._index) == 3000
    hit = vs.similarity_search_by_vector(x[11_000], k=1, filter={"g": 11_000 % 5})[0]
    assert hit.page_content == "c11000"
    re = MediQueryVectorStore(persist_directory=str(tmp_path / "db"), collection_metadata={"hnsw:space": "cosine"}, store_dtype=dtype)
    assert len(re) == 3000 and re.similarity_search_by_vector(x[11_999], k=1)[0].id == ids[11_999]


def test_collection_query_batches_through_the_cuda_path(gpu, medical_docs, embedder):
    """`vectorstore._collection.query(query_embeddings=[...])` -- chromadb's batch call, which langchain_chroma itself uses --
    over the reference's 154 chunks: the CUDA path and the oracle double return the same ids, documents and distances for
    all 154 queries at once, with and without a where / where_document filter."""
    from conftest import OracleIndex
    from mediquery_rag_b200 import Document, MediQueryVectorStore
    docs = [Document(page_content=d["page_content"], metadata=d["metadata"]) for d in medical_docs["docs"]]
    ids = [f"chunk-{j}" for j in range(len(docs))]
    emb = np.asarray(embedder.embed_documents([d.page_content for d in docs]), np.float32)
    stores = [MediQueryVectorStore(embedding_function=embedder, **kw) for kw in ({}, {"index_factory": OracleIndex})]
    for vs in stores:
        vs.add_documents(docs, ids=ids)
    for kw in ({}, {"where_document": {"$contains": "血压"}}, {"where": {"tags": {"$contains": docs[0].metadata["tags"].replace("，", ",").split(",")[0].strip()}}}):
        got, want = (vs._collection.query(query_embeddings=emb, n_results=5, **kw) for vs in stores)
        assert got["ids"] == want["ids"] and got["documents"] == want["documents"] and got["metadatas"] == want["metadatas"]
        assert np.array_equal(np.asarray(got["distances"], np.float64), np.asarray(want["distances"], np.float64))
    top1 = stores[0]._collection.query(query_embeddings=emb, n_results=1)
    # every chunk finds itself (or an identical chunk with a lower row number: the knowledge base repeats a few)
    assert [t[0] for t in top1["documents"]] == [d.page_content for d in docs]
    assert all(ids.index(r[0]) <= j for j, r in enumerate(top1["ids"]))
    assert stores[0]._collection.count() == 154


def test_very_large_async_batch_with_every_query_flagged_is_repaired_in_rounds_of_1024(gpu):
    """B > 4096 (k <= 32): the blind asynchronous search enqueues ceil(B / 1024) exact-scan rounds.  5000 queries all aimed at
    a row with 300 exact copies defeat the segmented pipeline's shortlist, so every round has a full list to work through."""
    import torch
    n, B, k = 20000, 5000, 10
    x = O.synth_rows(SEED, 0, n, 768)
    x[5000:5300] = x[123]
    q = (x[123][None, :] + 0.01 * O.synth_rows(78, 0, B, 768)).astype(np.float32)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    st, _ = O.normalize_rows(x, "bf16", 768)
    es, ei = O.brute_force_topk(qn, st, k)
    idx = native().Index(768, "bf16")
    idx.add(x)
    idx.set_path(4)
    qd = torch.from_numpy(q).cuda()
    out_s = torch.empty((B, k), dtype=torch.float32, device="cuda")
    out_i = torch.empty((B, k), dtype=torch.int64, device="cuda")
    before = idx.stats()["exact_fallbacks"]
    idx.search_device(qd, k, out_s, out_i, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert_bit_equal(out_s.cpu().numpy(), out_i.cpu().numpy(), es, ei)
    stt = idx.stats()
    assert stt["unrepaired"] == 0 and stt["exact_fallbacks"] - before >= B // 2
    s, i = idx.search(q, k)  # the blocking call over the same batch
    assert_bit_equal(s, i, es, ei)
    idx.close()
