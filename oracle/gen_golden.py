#!/usr/bin/env python
"""oracle/gen_golden.py -- TEST INFRASTRUCTURE ONLY.  Regenerates tests/golden/ from the real reference.

Run in the build container (needs /root/reference; the GPU box does not have it, which is why the outputs
are committed):   python oracle/gen_golden.py

It imports the reference's OWN src/ingest_medical.py -- with empty stand-ins for its three third-party
imports, which are not installed here (langchain_ollama, langchain_chroma, langchain_core.documents) --
and records what its parse_custom_format (src/ingest_medical.py:11-87) returns for
  (1) tests/golden/splitter_cases.txt   hand-written edge-case records          -> splitter_cases.json
  (2) /root/reference/data/medical_data.txt  the shipped knowledge base (config #1) -> medical_docs.json.gz
No reference source is copied; only the function's outputs are stored.
"""
from __future__ import annotations

import gzip
import hashlib
import importlib.util
import json
import os
import sys
import types

REF = os.environ.get("MQVS_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


class _Doc:
    def __init__(self, page_content, metadata=None, **kw):
        self.page_content = page_content
        self.metadata = metadata or {}


def _import_reference_ingest():
    for name, attrs in (("langchain_ollama", {"OllamaEmbeddings": object}),
                        ("langchain_chroma", {"Chroma": object}),
                        ("langchain_core", {}),
                        ("langchain_core.documents", {"Document": _Doc})):
        mod = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(mod, k, v)
        sys.modules[name] = mod
    spec = importlib.util.spec_from_file_location("ref_ingest_medical", os.path.join(REF, "src", "ingest_medical.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)  # module body under __main__ guard does not run
    return mod


def _dump(docs):
    return [{"page_content": d.page_content, "metadata": d.metadata} for d in docs]


def main() -> None:
    ref = _import_reference_ingest()
    cases = _dump(ref.parse_custom_format(os.path.join(GOLDEN, "splitter_cases.txt")))
    with open(os.path.join(GOLDEN, "splitter_cases.json"), "w", encoding="utf-8") as f:
        json.dump(cases, f, ensure_ascii=False, indent=1)
    data_path = os.path.join(REF, "data", "medical_data.txt")
    docs = _dump(ref.parse_custom_format(data_path))
    with open(data_path, "rb") as f:
        src_sha = hashlib.sha256(f.read()).hexdigest()
    payload = {"source_sha256": src_sha, "count": len(docs), "docs": docs}
    raw = json.dumps(payload, ensure_ascii=False).encode("utf-8")
    with open(os.path.join(GOLDEN, "medical_docs.json.gz"), "wb") as f:
        with gzip.GzipFile(fileobj=f, mode="wb", mtime=0) as g:
            g.write(raw)
    assert ref.parse_custom_format(os.path.join(GOLDEN, "does_not_exist.txt")) == []
    print(f"splitter_cases: {len(cases)} docs; medical_docs: {len(docs)} docs (source sha256 {src_sha[:12]})")


if __name__ == "__main__":
    main()
