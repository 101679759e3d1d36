"""Product-side record parser (mediquery_rag_b200.ingest) against the golden outputs of the reference's own
parse_custom_format and against the oracle restatement; ingest_file end to end on the test double."""
import gzip
import json
import os

import pytest

import oracle as O
from conftest import GOLDEN, OracleIndex
from mediquery_rag_b200 import ingest

REF_DATA = "/root/reference/data/medical_data.txt"


def test_edge_cases_match_reference_outputs():
    want = json.load(open(os.path.join(GOLDEN, "splitter_cases.json"), encoding="utf-8"))
    got = ingest.parse_file(os.path.join(GOLDEN, "splitter_cases.txt"))
    assert [(d.page_content, d.metadata) for d in got] == [(w["page_content"], w["metadata"]) for w in want]
    assert ingest.parse_file(os.path.join(GOLDEN, "missing.txt")) == []


def test_round_trip_of_the_golden_knowledge_base(medical_docs, tmp_path):
    """Re-serialise the 154 golden docs into the record format and parse them back: identical documents."""
    lines = []
    for n, d in enumerate(medical_docs["docs"]):
        title = d["metadata"]["title"]
        body = d["page_content"].split("\n答案：", 1)[1]
        lines.append(f"chunk_id:\t{n:03d}\ntitle:\t{title}\ncontent:\t{body}\nsource:\t书\ntags:\t{d['metadata']['tags']}\nreviewed_at:\t1\n\n")
    path = tmp_path / "kb.txt"
    path.write_text("".join(lines), encoding="utf-8")
    got = ingest.parse_file(str(path))
    assert len(got) == 154
    assert [(d.page_content, d.metadata) for d in got] == [(w["page_content"], w["metadata"]) for w in medical_docs["docs"]]
    assert [(t, m) for t, m in O.parse_custom_format(str(path))] == [(d.page_content, d.metadata) for d in got]


@pytest.mark.skipif(not os.path.exists(REF_DATA), reason="reference data file only exists in the build container")
def test_shipped_knowledge_base_matches_reference_outputs(medical_docs):
    got = ingest.parse_file(REF_DATA)
    assert [(d.page_content, d.metadata) for d in got] == [(w["page_content"], w["metadata"]) for w in medical_docs["docs"]]


def test_ingest_file_builds_a_searchable_store(tmp_path, embedder):
    db = str(tmp_path / "db")
    store = ingest.ingest_file(os.path.join(GOLDEN, "splitter_cases.txt"), embedder, db, batch_size=4,
                               index_factory=OracleIndex)
    assert len(store) == 10 and os.path.exists(os.path.join(store.collection_path, "manifest.json"))
    hit = store.similarity_search("问题：plain record\n答案：first answer line\nsecond line", k=1)[0]
    assert hit.metadata["title"] == "plain record" and hit.metadata["tags"] == "alpha，beta"
