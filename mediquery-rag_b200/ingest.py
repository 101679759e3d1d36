"""Ingest front half for the knowledge-base file format MediQuery-RAG ships (data/medical_data.txt).

The reference keeps its own parser (`parse_custom_format`, /root/reference/src/ingest_medical.py:11-87) and only
needs its Chroma import swapped.  This module offers the same record semantics as a library
function plus batched, GPU-side ingestion, for callers that want one call from text file to a searchable store:

    python -m mediquery_rag_b200.ingest data/medical_data.txt ./medical_db --embedder module:factory

Record format: records start at the literal `chunk_id:`; fields `title:`, `content:`, `source:`, `tags:` in that
order, one per line except `content`, which runs until the next `source:` (or, failing that, `tags:`) label.
"""
from __future__ import annotations

import importlib
import os
import re
import sys
from typing import Any, Dict, Iterator, List, Optional, Tuple

from .documents import Document

SOURCE_NAME = "《超越百岁》"       # constant the reference stamps on every chunk (src/ingest_medical.py:78)
UNTITLED = "未命名"               # src/ingest_medical.py:35


# The one-line fields follow the reference's pattern `label:\s*(.*?)\n` (src/ingest_medical.py:34, :66) to the letter,
# backtracking included: `\s*` may swallow line breaks (an empty field takes the next line), but when nothing that ends
# in a line break follows, the regex engine backs off and the field is EMPTY rather than absent; and it retries at later
# occurrences of the label.  A hand-rolled scan differs on exactly those inputs (a differential fuzz against the
# reference found 1.5 % of structured records), so the fields are read with the same pattern here.
_FIELD = {label: re.compile(label + r"\s*(.*?)\n") for label in ("title:", "tags:")}
_CONTENT = re.compile(r"content:\s*")


def _line_value(record: str, label: str) -> Optional[str]:
    """Stripped value of a one-line field, or None when the pattern matches nowhere in the record."""
    m = _FIELD[label].search(record)
    return m.group(1).strip() if m else None


def _body(record: str) -> str:
    """Text after the first `content:` up to the first `source:` (failing that, `tags:`) after it; a label that slipped
    inside that span ends it (src/ingest_medical.py:39-63)."""
    m = _CONTENT.search(record)
    if m is None:
        return ""
    begin = m.end()
    stop = record.find("source:", begin)
    if stop < 0:
        stop = record.find("tags:", begin)
    if stop < 0:
        return record[begin:].strip()
    text = record[begin:stop]
    for label in ("source:", "tags:"):
        cut = text.find(label)
        if cut >= 0:
            text = text[:cut]
    return text.strip()


def iter_records(text: str) -> Iterator[Tuple[str, Dict[str, str]]]:
    """(page_content, metadata) per record, with the reference's conventions."""
    for record in text.split("chunk_id:"):
        if not record.strip():
            continue
        title = _line_value(record, "title:")
        title = UNTITLED if title is None else title
        tags = _line_value(record, "tags:") or ""
        body = _body(record)
        if title or body:
            yield f"问题：{title}\n答案：{body}", {"title": title, "tags": tags, "source": SOURCE_NAME}


def parse_file(path: str) -> List[Document]:
    if not os.path.exists(path):
        return []
    with open(path, "r", encoding="utf-8") as f:
        return [Document(page_content=t, metadata=m) for t, m in iter_records(f.read())]


def ingest_file(path: str, embedding: Any, persist_directory: Optional[str] = None, batch_size: int = 4096,
                **store_kwargs: Any):
    """Parse `path`, embed in batches, add to a (new or reopened) store; returns the store."""
    from .vectorstore import MediQueryVectorStore

    docs = parse_file(path)
    store = MediQueryVectorStore(embedding_function=embedding, persist_directory=persist_directory,
                                 auto_persist=False, **store_kwargs)
    for a in range(0, len(docs), batch_size):
        store.add_documents(docs[a:a + batch_size])
    if persist_directory is not None:
        store.persist()
    return store


def main(argv: Optional[List[str]] = None) -> int:
    import argparse

    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("text_file")
    ap.add_argument("persist_directory")
    ap.add_argument("--embedder", required=True, help="module:callable returning an Embeddings object, e.g. "
                                                      "langchain_ollama:OllamaEmbeddings")
    ap.add_argument("--embedder-arg", action="append", default=[], help="key=value passed to the embedder factory")
    ap.add_argument("--store-dtype", default="fp32", choices=["fp32", "bf16"])
    args = ap.parse_args(argv)
    mod, _, attr = args.embedder.partition(":")
    factory = getattr(importlib.import_module(mod), attr)
    kwargs = dict(kv.split("=", 1) for kv in args.embedder_arg)
    store = ingest_file(args.text_file, factory(**kwargs), args.persist_directory, store_dtype=args.store_dtype)
    print(f"ingested {len(store)} chunks into {args.persist_directory}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
