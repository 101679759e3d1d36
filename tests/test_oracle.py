"""Oracle self-consistency: canonical fp32 (C, fmaf) vs float64 referee vs numpy sgemm; generator; tie rules."""
import numpy as np
import pytest

import oracle as O

SEED = 0x5EED0002


def test_generator_numpy_and_c_agree_bitwise():
    a = O.synth_rows(SEED, 12345, 300, 768)
    b = O.synth_rows(SEED, 12345, 300, 768, use_numpy=True)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert a.min() >= -1.0 and a.max() < 1.0 and abs(a.mean()) < 0.01
    # counter-based: a window equals the same rows of a longer run
    assert np.array_equal(O.synth_rows(SEED, 100, 50, 768), O.synth_rows(SEED, 0, 150, 768)[100:])


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_normalised_rows_are_unit(dtype):
    x = O.synth_rows(SEED, 0, 500, 768) * 7.5
    st, bad = O.normalize_rows(x, dtype)
    assert bad == 0
    norms = np.linalg.norm(O.stored_as_f32(st).astype(np.float64), axis=1)
    assert np.allclose(norms, 1.0, atol=2e-3 if dtype == "bf16" else 1e-6)


def test_bf16_rounding_is_nearest_even():
    x = np.zeros((1, 128), np.float32)
    # values straddling bf16 ties: 1 + 2^-8 is exactly halfway between 1 and 1+2^-7 -> even (1.0)
    x[0, :4] = [1.0, 1.0 + 2.0 ** -8, 1.0 + 2.0 ** -7 + 2.0 ** -8, 1.0 + 2.0 ** -8 + 2.0 ** -20]
    st, _ = O.normalize_rows(x, "fp32")
    inv = 1.0 / np.linalg.norm(x[0].astype(np.float64))
    stb, _ = O.normalize_rows(x, "bf16")
    import torch
    want = torch.from_numpy(st).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(stb, want) and inv > 0


def test_non_finite_rows_are_counted_and_zeroed():
    x = O.synth_rows(SEED, 0, 4, 768)
    x[1, 5] = np.nan
    x[3, 0] = np.inf
    st, bad = O.normalize_rows(x, "fp32")
    assert bad == 2 and not st[1].any() and not st[3].any() and st[0].any()


def test_zero_row_stays_zero_and_scores_zero():
    x = O.synth_rows(SEED, 0, 3, 768)
    x[1] = 0
    st, bad = O.normalize_rows(x, "fp32")
    assert bad == 0 and not st[1].any()
    qn, _ = O.normalize_rows(O.synth_rows(1, 0, 1, 768))
    assert O.canonical_scores(qn, st)[0, 1] == 0.0


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_canonical_fp32_vs_f64_referee_and_sgemm(dtype):
    n, B, k = 30000, 16, 10
    st, _ = O.normalize_rows(O.synth_rows(SEED, 0, n, 768), dtype)
    q, planted = O.planted_queries(0x0BADC0DE, B, SEED, n, 768)
    qn, _ = O.normalize_rows(q)
    s, i = O.brute_force_topk(qn, st, k)
    rs, ri = O.referee_topk_f64(qn, st, k)
    gs, gi = O.sgemm_topk(qn, O.stored_as_f32(st), k)
    # tolerance stated by north_star: scores within 1e-5 relative of the exact result
    assert np.max(np.abs(s - rs) / np.abs(rs)) < 1e-5
    assert np.max(np.abs(gs - rs) / np.abs(rs)) < 1e-5
    # ids identical unless two neighbours are closer than fp32 can order (documented near-tie rule)
    for a, b in ((i, ri), (gi, ri)):
        for row in np.argwhere(a != b)[:, 0]:
            assert set(a[row]) == set(b[row]) or np.min(np.abs(np.diff(rs[row]))) < 2e-7
    assert all(i[b, 0] == planted[b] for b in range(0, B, 2))
    assert np.all(np.diff(s, axis=1) <= 0)


def test_exact_ties_break_to_lower_id_and_mask_applies_first():
    base = O.synth_rows(SEED, 0, 64, 768)
    x = np.concatenate([base, base[:8], base[:8]])  # rows 64..71 and 72..79 duplicate rows 0..7
    st, _ = O.normalize_rows(x)
    qn, _ = O.normalize_rows(base[3:4])
    s, i = O.brute_force_topk(qn, st, 5)
    assert i[0, :3].tolist() == [3, 67, 75] and s[0, 0] == s[0, 1] == s[0, 2]
    allowed = np.ones(80, bool)
    allowed[3] = False
    s2, i2 = O.brute_force_topk(qn, st, 5, allowed)
    assert i2[0, :2].tolist() == [67, 75]
    rs, ri = O.referee_topk_f64(qn, st, 5, allowed)
    assert np.array_equal(ri, i2)


def test_k_larger_than_allowed_rows_is_padded():
    st, _ = O.normalize_rows(O.synth_rows(SEED, 0, 40, 768))
    qn, _ = O.normalize_rows(O.synth_rows(9, 0, 2, 768))
    allowed = np.zeros(40, bool)
    allowed[[4, 9, 30]] = True
    s, i = O.brute_force_topk(qn, st, 6, allowed)
    assert set(i[0, :3]) == {4, 9, 30} and (i[:, 3:] == -1).all() and np.isinf(s[:, 3:]).all()
    s0, i0 = O.brute_force_topk(qn, st[:0], 3)
    assert (i0 == -1).all()


def test_orthogonal_basis_scores_are_exact():
    x = np.eye(768, dtype=np.float32)[:100] * 3.0
    st, _ = O.normalize_rows(x)
    s, i = O.brute_force_topk(st[17:18], st, 3)
    assert i[0, 0] == 17 and s[0, 0] == 1.0 and s[0, 1] == 0.0 and i[0, 1] == 0


def test_merge_topk_matches_global_search():
    st, _ = O.normalize_rows(O.synth_rows(SEED, 0, 5000, 768))
    qn, _ = O.normalize_rows(O.synth_rows(3, 0, 4, 768))
    s, i = O.brute_force_topk(qn, st, 10)
    parts = [O.brute_force_topk(qn, st[a:b], 10, id_base=a) for a, b in ((0, 1700), (1700, 1705), (1705, 5000))]
    ms, mi = O.merge_topk(np.concatenate([p[0] for p in parts], 1), np.concatenate([p[1] for p in parts], 1), 10)
    assert np.array_equal(mi, i) and np.array_equal(ms, s)


def test_where_evaluator_grammar():
    mds = [{"title": "a", "n": 1}, {"title": "b", "n": 2}, {}, {"title": "a", "n": 3}]
    ev = lambda w: O.eval_where(w, mds).astype(int).tolist()
    assert ev({"title": "a"}) == [1, 0, 0, 1]
    assert ev({"n": {"$gte": 2}}) == [0, 1, 0, 1]
    assert ev({"title": {"$ne": "a"}}) == [0, 1, 1, 0]
    assert ev({"$and": [{"title": "a"}, {"n": {"$lt": 3}}]}) == [1, 0, 0, 0]
    assert ev({"$or": [{"title": "b"}, {"n": {"$in": [3]}}]}) == [0, 1, 0, 1]
    assert ev(None) == [1, 1, 1, 1]


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_oracle_against_an_independent_float64_referee_at_one_million_rows(dtype):
    """The canonical-order oracle is what the kernels are compared with bit for bit; this ties IT to arithmetic that shares
    nothing with the kernels (float64 numpy over the stored values, chromadb's distance definitions) at config 2's full
    size, for both storage types and all three spaces, within north_star's tolerance."""
    from referee import assert_within_north_star, f64_topk
    n, B, k = 1_000_000, 6, 10
    x = O.synth_rows(SEED, 0, n, 768)
    x *= np.linspace(0.03, 0.09, n, dtype=np.float32)[:, None]       # raw norms 0.5 .. 1.5: the spaces really differ
    q, _ = O.planted_queries(0x0BADC0DE, B, SEED, n, 768)
    q = (q * np.float32(1.3)).astype(np.float32)
    st, _ = O.normalize_rows(x, dtype, 768)
    norms = O.row_norms(x, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    del x
    for space in ("cosine", "l2", "ip"):
        kw = {} if space == "cosine" else dict(metric=space, norms=norms, q_norms=O.row_norms(q, 768))
        s, i = O.brute_force_topk(qn, st, k, **kw)
        ref_d, ref_i = f64_topk(q, st, norms, k, space)
        assert_within_north_star(space, s, i, q, ref_d, ref_i)


@pytest.mark.parametrize("order", ["scalar", "simd16"])
def test_oracle_against_the_published_arithmetic_of_chroma_hnswlib(order):
    """What the reference actually runs is chroma-hnswlib (un-vendored; oracle/hnswlib_restated.c restates its distance
    functions -- sequential and 16-lane SIMD summation order, `1 - ip` / squared L2 over the RAW vectors, normalize_vector
    for the cosine space -- and its exhaustive search).  The canonical oracle must agree with it under north_star's
    tolerance in all three spaces: ids identical apart from near ties, distances within 1e-5 relative."""
    from referee import assert_within_north_star
    n, B, k = 200_000, 6, 10
    x = O.synth_rows(SEED, 0, n, 768)
    x *= np.linspace(0.03, 0.09, n, dtype=np.float32)[:, None]       # raw norms 0.5 .. 1.5
    q, _ = O.planted_queries(0x0BADC0DE, B, SEED, n, 768)
    q = (q * np.float32(1.3)).astype(np.float32)
    st, _ = O.normalize_rows(x, "fp32", 768)
    norms = O.row_norms(x, 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    for space in ("cosine", "l2", "ip"):
        kw = {} if space == "cosine" else dict(metric=space, norms=norms, q_norms=O.row_norms(q, 768))
        s, i = O.brute_force_topk(qn, st, k, **kw)
        ref_d, ref_i = O.hnswlib_topk(space, x, q, k, order=order)
        assert_within_north_star(space, s, i, q, ref_d.astype(np.float64), ref_i)


def test_hnswlib_restatement_known_answers():
    """The restated exhaustive search on hand-made distances: ties keep the LOWER label in the result (the heap's worst is
    the larger label), a filter removes rows before selection, short stores are padded."""
    L = O.lib()
    import ctypes
    d = np.array([0.5, 0.1, 0.5, 0.3, 0.1, 0.9], np.float32)
    out_d, out_i = np.empty(3, np.float32), np.empty(3, np.int64)
    L.hnsw_bruteforce_knn(ctypes.c_void_p(d.ctypes.data), 6, None, 3, ctypes.c_void_p(out_d.ctypes.data), ctypes.c_void_p(out_i.ctypes.data))
    assert out_i.tolist() == [1, 4, 3] and out_d.tolist() == [np.float32(0.1), np.float32(0.1), np.float32(0.3)]
    allow = O.mask_from_bool(np.array([1, 0, 1, 0, 0, 1], bool))
    L.hnsw_bruteforce_knn(ctypes.c_void_p(d.ctypes.data), 6, ctypes.c_void_p(allow.ctypes.data), 3, ctypes.c_void_p(out_d.ctypes.data), ctypes.c_void_p(out_i.ctypes.data))
    assert out_i.tolist() == [0, 2, 5]
    L.hnsw_bruteforce_knn(ctypes.c_void_p(d.ctypes.data), 2, None, 3, ctypes.c_void_p(out_d.ctypes.data), ctypes.c_void_p(out_i.ctypes.data))
    assert out_i.tolist() == [1, 0, -1] and np.isinf(out_d[2])
    # unit basis vectors: ip distance exactly 0 / 1, squared l2 exactly 0 / 2, in either summation order
    e = np.eye(4, 768, dtype=np.float32)
    for order in ("scalar", "simd16"):
        dd, ii = O.hnswlib_topk("ip", e, e[2:3], 4, order=order)
        assert ii[0, 0] == 2 and dd[0].tolist() == [0.0, 1.0, 1.0, 1.0]
        dd, ii = O.hnswlib_topk("l2", e, e[2:3], 4, order=order)
        assert ii[0].tolist() == [2, 0, 1, 3] and dd[0].tolist() == [0.0, 2.0, 2.0, 2.0]


def test_hnswlib_restatement_with_a_filter_agrees_with_the_masked_oracle():
    """hnswlib's exhaustive search takes a filter functor (chromadb hands it the ids a `where` allows): same rows as the
    canonical oracle under the same allow-bitmask, k clamped to the allowed rows."""
    from referee import assert_within_north_star
    n, B, k = 30_000, 4, 10
    x = O.synth_rows(SEED, 0, n, 768)
    q, _ = O.planted_queries(0x0BADC0DE, B, SEED, n, 768)
    allowed = (np.arange(n) % 7) == 3
    st, _ = O.normalize_rows(x, "fp32", 768)
    qn, _ = O.normalize_rows(q, "fp32", 768)
    s, i = O.brute_force_topk(qn, st, k, O.mask_from_bool(allowed))
    ref_d, ref_i = O.hnswlib_topk("cosine", x, q, k, allowed=allowed)
    assert allowed[ref_i].all()
    assert_within_north_star("cosine", s, i, q, ref_d.astype(np.float64), ref_i)
    few = np.zeros(n, bool)
    few[[5, 77, 4000]] = True
    d3, i3 = O.hnswlib_topk("l2", x, q, k, allowed=few)
    assert (np.sort(i3[:, :3], axis=1) == np.array([5, 77, 4000])).all() and (i3[:, 3:] == -1).all() and np.isinf(d3[:, 3:]).all()
