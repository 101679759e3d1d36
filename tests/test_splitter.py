"""Oracle splitter vs golden outputs of the reference's own parse_custom_format (oracle/gen_golden.py)."""
import hashlib
import json
import os

import pytest

import oracle as O
from conftest import GOLDEN

REF_DATA = "/root/reference/data/medical_data.txt"


def test_edge_cases_match_reference_outputs():
    want = json.load(open(os.path.join(GOLDEN, "splitter_cases.json"), encoding="utf-8"))
    got = O.parse_custom_format(os.path.join(GOLDEN, "splitter_cases.txt"))
    assert len(got) == len(want) == 10
    for (text, md), w in zip(got, want):
        assert text == w["page_content"]
        assert md == w["metadata"]


def test_missing_file_gives_empty_list():
    assert O.parse_custom_format(os.path.join(GOLDEN, "nope.txt")) == []


def test_medical_golden_shape(medical_docs):
    # 154 docs, page_content 24-398 chars, metadata keys {title,tags,source}
    docs = medical_docs["docs"]
    assert medical_docs["count"] == len(docs) == 154
    lens = [len(d["page_content"]) for d in docs]
    assert min(lens) == 24 and max(lens) == 398
    assert all(set(d["metadata"]) == {"title", "tags", "source"} for d in docs)
    assert all(d["page_content"].startswith("问题：") for d in docs)


@pytest.mark.skipif(not os.path.exists(REF_DATA), reason="reference data file only exists in the build container")
def test_medical_data_matches_reference_outputs(medical_docs):
    assert hashlib.sha256(open(REF_DATA, "rb").read()).hexdigest() == medical_docs["source_sha256"]
    got = O.parse_custom_format(REF_DATA)
    assert len(got) == 154
    for (text, md), w in zip(got, medical_docs["docs"]):
        assert text == w["page_content"] and md == w["metadata"]
