"""Host half of the device-side `where` evaluation (mediquery_rag_b200.filters.DeviceFilter): the Chroma filter ->
postfix mqvs_filter_op program compiler, the typed / dictionary-coded columns it uploads, and its refusal (NotOnDevice)
of the cases whose semantics only the host evaluator reproduces.  The index is a numpy interpreter of the program that
mirrors kernels_filter.cu line by line, so compiler + program semantics are pinned against filters.evaluate on CPU;
tests/test_gpu_parity.py runs the same filters through the real kernel."""
import numpy as np
import pytest

from mediquery_rag_b200 import _native
from mediquery_rag_b200.filters import (DeviceFilter, MetadataColumns, NotOnDevice, evaluate, evaluate_document, pack_mask,
                                        texts_csr)


class ProgramInterpreter:
    """Stands in for _native.Index: column_set / filter_build with the kernel's semantics, in numpy."""

    def __init__(self, n):
        self.n, self.cols = n, {}

    def __len__(self):
        return self.n

    def column_set(self, col_id, values, present=None):
        assert values.dtype in (np.int64, np.float64)
        self.cols[col_id] = (values.copy(), None if present is None else np.asarray(present, bool).copy())

    def column_set_tokens(self, col_id, offsets, tokens):
        assert offsets.dtype == np.uint64 and tokens.dtype == np.int64 and offsets[0] == 0 and offsets[-1] == tokens.shape[0]
        self.cols[col_id] = ("tokens", offsets.copy(), tokens.copy())

    def column_set_text(self, col_id, offsets, data, present=None):
        assert offsets.dtype == np.uint64 and data.dtype == np.uint8 and offsets[0] == 0 and offsets[-1] == data.shape[0]
        self.cols[col_id] = ("text", offsets.copy(), data.tobytes(), None if present is None else np.asarray(present, bool).copy())

    def filter_build(self, ops, pool=None, live_words=None):
        n, stack = self.n, []
        for o in ops:
            if o.op in (_native.OP["$text_contains"], _native.OP["$text_not_contains"]):  # byte-wise search, CSR rows
                _, off, data, present = self.cols[o.col]
                assert 1 <= o.set_len <= _native.MAX_NEEDLE and o.set_off * 8 + o.set_len <= pool.shape[0] * 8
                needle = pool.tobytes()[o.set_off * 8:o.set_off * 8 + o.set_len]
                found = np.zeros(n, bool)
                for r in range(min(n, off.shape[0] - 1)):
                    found[r] = (present is None or present[r]) and needle in data[int(off[r]):int(off[r + 1])]
                stack.append(found != (o.op == _native.OP["$text_not_contains"]))
                continue
            if o.op in (_native.OP["$and"], _native.OP["$or"]):
                b2, b1 = stack.pop(), stack.pop()
                stack.append(b1 & b2 if o.op == _native.OP["$and"] else b1 | b2)
                continue
            if o.op in (_native.OP["$contains"], _native.OP["$not_contains"]):  # token-set membership, CSR rows
                _, off, tok = self.cols[o.col]
                found = np.zeros(n, bool)
                for r in range(min(n, off.shape[0] - 1)):
                    found[r] = (tok[int(off[r]):int(off[r + 1])] == o.ival).any()
                stack.append(found != (o.op == _native.OP["$not_contains"]))
                continue
            vals, present = self.cols[o.col]
            m = vals.shape[0]
            v = np.zeros(n, vals.dtype)
            v[:min(m, n)] = vals[:n]
            pres = np.zeros(n, bool)
            pres[:min(m, n)] = True if present is None else present[:n]
            name = {c: k for k, c in _native.OP.items()}[o.op]
            negated = name in ("$ne", "$nin")
            if name in ("$in", "$nin"):
                consts = pool[o.set_off:o.set_off + o.set_len].view(vals.dtype)
                hit = np.isin(v, consts) != negated
            else:
                c = o.ival if vals.dtype == np.int64 else o.fval
                hit = {"$eq": v == c, "$ne": v != c, "$gt": v > c, "$gte": v >= c, "$lt": v < c, "$lte": v <= c}[name]
            stack.append(np.where(pres, hit, negated))
        allowed = stack.pop() if ops else np.ones(n, bool)
        assert not stack
        words = pack_mask(allowed)
        if live_words is not None:
            words &= live_words
        return words, int(sum(bin(int(w)).count("1") for w in words))


def make_cols(n=300, seed=0):
    rng = np.random.default_rng(seed)
    cols = MetadataColumns()
    rows = []
    for r in range(n):
        md = {"year": int(rng.integers(1990, 2025)), "title": f"t{int(rng.integers(0, 12))}"}
        if rng.random() < 0.8:
            md["score"] = float(np.round(rng.random() * 10, 1))
        if rng.random() < 0.5:
            md["tags"] = str(rng.choice(["sleep,diet", "diet", "exercise,sleep", ""]))
        if rng.random() < 0.3:
            md["mixed"] = rng.choice([1, "one", 2.5], p=[0.4, 0.3, 0.3]).item() if False else [1, "one", 2.5][int(rng.integers(0, 3))]
        rows.append(md)
    cols.extend(rows)
    cols.extend_columns(50, {"source_id": rng.integers(0, 16, 50), "year": rng.integers(1990, 2025, 50)})
    return cols


FILTERS = [
    {"year": 2001}, {"year": {"$ne": 2001}}, {"year": {"$gt": 2010}}, {"year": {"$lte": 1995}}, {"year": {"$gte": 2010.0}},
    {"year": {"$in": [1999, 2003, 2020]}}, {"year": {"$nin": [1999, 2003, 2020]}}, {"year": {"$in": []}},
    {"score": {"$gt": 5}}, {"score": {"$lt": 2.5}}, {"score": {"$eq": 3.3}}, {"score": {"$ne": 3.3}}, {"score": {"$in": [1.0, 2.5, 7]}},
    {"title": "t3"}, {"title": {"$ne": "t3"}}, {"title": {"$in": ["t1", "t9", "nope"]}}, {"title": {"$nin": ["t1"]}},
    {"title": {"$gt": "t5"}}, {"title": {"$gte": "t5"}}, {"title": {"$lt": "t10"}}, {"title": {"$lte": "t10"}}, {"title": {"$eq": "absent"}},
    {"tags": "diet"}, {"tags": {"$ne": ""}}, {"source_id": {"$in": [0, 1, 2]}}, {"source_id": {"$lt": 8}}, {"source_id": {"$ne": 3}},
    {"nosuchfield": 1}, {"nosuchfield": {"$ne": 1}},
    {"$and": [{"year": {"$gte": 2000}}, {"score": {"$lt": 5}}]},
    {"$or": [{"title": "t1"}, {"title": "t2"}, {"year": {"$lt": 1992}}]},
    {"$and": [{"$or": [{"year": 1999}, {"score": {"$gte": 9}}]}, {"title": {"$nin": ["t0"]}}, {"source_id": {"$ne": 5}}]},
]


@pytest.mark.parametrize("where", FILTERS, ids=[str(f)[:50] for f in FILTERS])
def test_compiled_program_equals_host_evaluator(where):
    cols = make_cols()
    idx = ProgramInterpreter(len(cols))
    live = np.random.default_rng(1).random(len(cols)) > 0.1
    words, allowed = DeviceFilter(idx, cols).build(where, live)
    want = evaluate(where, cols) & live
    assert np.array_equal(words, pack_mask(want)) and allowed == int(want.sum())


@pytest.mark.parametrize("where", [{"mixed": 1}, {"year": True}, {"year": {"$gt": 2000.5}}, {"year": "1999"}, {"title": 3},
                                   {"title": {"$in": ["t1", 2]}}, {"score": None}])
def test_cases_only_the_host_evaluator_reproduces_are_refused(where):
    cols = make_cols()
    with pytest.raises(NotOnDevice):
        DeviceFilter(ProgramInterpreter(len(cols)), cols).build(where, None)


def test_columns_are_reuploaded_only_when_the_metadata_changed():
    cols = make_cols()
    idx = ProgramInterpreter(len(cols))
    uploads = []
    orig = idx.column_set
    idx.column_set = lambda *a, **k: (uploads.append(a[0]), orig(*a, **k))[1]
    f = DeviceFilter(idx, cols)
    f.build({"year": 2001}, None)
    f.build({"year": {"$gt": 2001}}, None)
    assert uploads == [0]
    cols.extend([{"year": 2001}])
    idx.n = len(cols)
    words, allowed = f.build({"year": 2001}, None)
    assert uploads == [0, 0] and allowed == int(evaluate({"year": 2001}, cols).sum())


def where_trees():
    """hypothesis strategy: random Chroma `where` trees over the typed fields of make_cols()."""
    from hypothesis import strategies as hs
    ints = hs.integers(1985, 2030)
    floats = hs.sampled_from([0.0, 1.0, 2.5, 3.3, 5.0, 7.7, 9.9, 10.0, -1.0])
    titles = hs.sampled_from([f"t{i}" for i in range(14)] + ["", "t", "zz"])
    cmp_ops = hs.sampled_from(["$eq", "$ne", "$gt", "$gte", "$lt", "$lte"])
    set_ops = hs.sampled_from(["$in", "$nin"])

    def leaf(field, consts):
        return hs.one_of(hs.tuples(cmp_ops, consts).map(lambda t: {field: {t[0]: t[1]}}),
                         hs.tuples(set_ops, hs.lists(consts, min_size=0, max_size=5)).map(lambda t: {field: {t[0]: t[1]}}),
                         consts.map(lambda c: {field: c}))

    leaves = hs.one_of(leaf("year", ints), leaf("score", floats), leaf("title", titles),
                       leaf("source_id", hs.integers(-1, 17)), leaf("tags", hs.sampled_from(["diet", "sleep,diet", "", "x"])),
                       leaf("nosuchfield", ints))
    return hs.recursive(leaves, lambda kids: hs.tuples(hs.sampled_from(["$and", "$or"]), hs.lists(kids, min_size=1, max_size=3))
                        .map(lambda t: {t[0]: t[1]}), max_leaves=8)


def test_random_where_trees_compile_to_programs_that_equal_the_host_evaluator():
    from hypothesis import given, settings

    cols = make_cols(seed=3)
    live = np.random.default_rng(2).random(len(cols)) > 0.2

    @settings(max_examples=200, deadline=None, derandomize=True)
    @given(where=where_trees())
    def run(where):
        idx = ProgramInterpreter(len(cols))
        words, allowed = DeviceFilter(idx, cols).build(where, live)
        want = evaluate(where, cols) & live
        assert np.array_equal(words, pack_mask(want)) and allowed == int(want.sum())

    run()


def test_tag_token_filters_on_the_reference_knowledge_base(medical_docs):
    """`tags` is ONE comma-joined string per chunk (src/ingest_medical.py:66-67,75-79): $contains / $not_contains on a
    metadata field is token membership.  Host evaluator and the compiled program (token dictionary + CSR lists) agree with
    a plain Python restatement on the 154 real chunks, for every tag that occurs and for combinations."""
    import re
    metas = [d["metadata"] for d in medical_docs["docs"]]
    cols = MetadataColumns()
    cols.extend(metas)
    n = len(metas)
    split = [[t.strip() for t in re.split("[,\uff0c]", m["tags"]) if t.strip()] for m in metas]
    every = sorted({t for ts in split for t in ts})
    assert len(every) > 100
    dev = DeviceFilter(ProgramInterpreter(n), cols)
    for tag in every[::7] + ["癌症", "运动", "no such tag", " 癌症 "]:
        want = np.array([tag.strip() in ts for ts in split])
        for where, expect in (({"tags": {"$contains": tag}}, want), ({"tags": {"$not_contains": tag}}, ~want)):
            assert np.array_equal(evaluate(where, cols), expect)
            words, allowed = dev.build(where, None)
            assert np.array_equal(words, pack_mask(expect)) and allowed == int(expect.sum())
    both = {"$and": [{"tags": {"$contains": "癌症"}}, {"$or": [{"tags": {"$contains": "癌症检测"}}, {"title": {"$gte": "癌"}}]}]}
    words, allowed = dev.build(both, None)
    assert np.array_equal(words, pack_mask(evaluate(both, cols))) and 0 < allowed < n
    assert len(dev.index.cols) == 2  # one token column for `tags`, one dictionary-coded column for `title`
    # a substring of a tag is not a tag; a whole-string equality still goes through $eq
    assert not evaluate({"tags": {"$contains": "癌"}}, cols).any()
    assert evaluate({"tags": metas[0]["tags"]}, cols)[0]
    with pytest.raises(Exception):
        evaluate({"tags": {"$contains": 5}}, cols)


DOCUMENT_FILTERS = [
    {"$contains": "血压"}, {"$not_contains": "血压"}, {"$contains": "问题："}, {"$contains": "no such text"}, {"$contains": "\n答案："},
    {"$and": [{"$contains": "运动"}, {"$not_contains": "癌"}]}, {"$or": [{"$contains": "睡眠"}, {"$contains": "饮食"}, {"$contains": "VO2"}]},
    {"$and": [{"$or": [{"$contains": "糖"}, {"$contains": "脂"}]}, {"$not_contains": "胰岛素"}]},
]


def test_where_document_programs_on_the_reference_knowledge_base(medical_docs):
    """where_document over the reference's real page_content ("问题：{title}\\n答案：{content}", src/ingest_medical.py:71): the
    compiled program (text column = the documents' UTF-8 bytes in CSR form, needle in the constant pool, byte-wise search as
    in kernels_filter.cu) agrees with the host evaluator (`needle in text` on str), alone and AND-ed with a metadata filter."""
    texts = [d["page_content"] for d in medical_docs["docs"]] + [None, "", "血"]
    metas = [d["metadata"] for d in medical_docs["docs"]] + [None, {"tags": "x"}, None]
    cols = MetadataColumns()
    cols.extend(metas)
    n = len(texts)
    dev = DeviceFilter(ProgramInterpreter(n), cols, texts)
    live = np.ones(n, bool)
    live[::9] = False
    for wd in DOCUMENT_FILTERS:
        want = evaluate_document(wd, texts)
        words, allowed = dev.build(None, None, wd)
        assert np.array_equal(words, pack_mask(want)) and allowed == int(want.sum()), wd
        where = {"tags": {"$not_contains": "运动"}}
        both = want & evaluate(where, cols) & live
        words, allowed = dev.build(where, live, wd)
        assert np.array_equal(words, pack_mask(both)) and allowed == int(both.sum()), wd
    assert 0 < int(evaluate_document({"$contains": "血压"}, texts).sum()) < n
    # one text column and one token column, uploaded once
    assert sorted(type(c[0]) is str and c[0] for c in dev.index.cols.values()) == ["text", "tokens"]
    # what only the host evaluator reproduces: the empty needle (matches every row that has a document), oversized needles
    for wd in ({"$contains": ""}, {"$contains": "x" * 2000}):
        with pytest.raises(NotOnDevice):
            dev.build(None, None, wd)
    with pytest.raises(Exception):
        dev.build(None, None, {"$regex": "a"})


def test_text_columns_follow_the_documents():
    """The text column is rebuilt when rows are added; bulk rows without documents cost no bytes; a needle that straddles two
    documents is not found (rows are searched one by one)."""
    from mediquery_rag_b200.vectorstore import _RowColumn
    texts = _RowColumn()
    texts.extend(["alpha beta", "gamma"])
    texts.extend_implicit(1000, None)
    texts.extend(["beta\u00e9", None, "\U0001f600 beta"])
    cols = MetadataColumns()
    cols.extend_empty(len(texts))
    off, data, present = texts_csr(texts)
    assert off.shape[0] == len(texts) + 1 and int(off[-1]) == data.shape[0] == len("alpha betagammabeta\u00e9\U0001f600 beta".encode())
    assert present.sum() == 4 and not present[2:1002].any()
    plain = [texts[r] for r in range(len(texts))]
    off2, data2, present2 = texts_csr(plain)
    assert np.array_equal(off, off2) and np.array_equal(data, data2) and np.array_equal(present, present2)
    idx = ProgramInterpreter(len(texts))
    dev = DeviceFilter(idx, cols, lambda: texts)
    for needle in ("beta", "a", "betagamma", "\u00e9", "\U0001f600", "gamma"):
        wd = {"$contains": needle}
        words, allowed = dev.build(None, None, wd)
        assert np.array_equal(words, pack_mask(evaluate_document(wd, plain))), needle
    texts.extend(["new beta"])
    plain.append("new beta")
    cols.extend_empty(1)
    idx.n = len(texts)
    words, allowed = dev.build(None, None, {"$contains": "beta"})
    assert allowed == 4 and np.array_equal(words, pack_mask(evaluate_document({"$contains": "beta"}, plain)))
