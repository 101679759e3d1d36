"""oracle/splitter.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's record splitter, `parse_custom_format`
(/root/reference/src/ingest_medical.py:11-87), returning plain (page_content, metadata) pairs.
PINNED: tests/golden/splitter_*.json(.gz) were produced by importing the reference's own function
(oracle/gen_golden.py) and tests/test_splitter.py checks this restatement against them.

Behaviour followed, with the reference line it comes from:
  :15-17  missing file -> []                      :24     records are the pieces between literal "chunk_id:"
  :30     whitespace-only pieces are skipped      :34-35  title = first "title:\\s*(.*?)\\n", stripped, else "未命名"
  :39-63  content starts after the first "content:\\s*"; it ends at the first "source:" after it, else the first
          "tags:" after it, else the end of the piece; inside that span anything from an embedded "source:" or
          "tags:" on is dropped (only reachable for "tags:" before "source:"); stripped
  :66-67  tags = first "tags:\\s*(.*?)\\n", stripped, else ""
  :70-80  kept when title or content is non-empty; page_content = "问题：{title}\\n答案：{content}";
          metadata = {"title", "tags", "source": "《超越百岁》"}
"""
from __future__ import annotations

import os
import re
from typing import Dict, List, Tuple

_TITLE = re.compile(r"title:\s*(.*?)\n")
_TAGS = re.compile(r"tags:\s*(.*?)\n")
_CONTENT = re.compile(r"content:\s*")
SOURCE_NAME = "《超越百岁》"
DEFAULT_TITLE = "未命名"


def _first_field(rx: "re.Pattern[str]", piece: str, default: str) -> str:
    m = rx.search(piece)
    return m.group(1).strip() if m else default


def _content_of(piece: str) -> str:
    m = _CONTENT.search(piece)
    if m is None:
        return ""
    begin = m.end()
    stop = piece.find("source:", begin)
    if stop < 0:
        stop = piece.find("tags:", begin)
    if stop < 0:
        return piece[begin:].strip()
    span = piece[begin:stop]
    for label in ("source:", "tags:"):
        cut = span.find(label)
        if cut >= 0:
            span = span[:cut]
    return span.strip()


def parse_text(text: str) -> List[Tuple[str, Dict[str, str]]]:
    out: List[Tuple[str, Dict[str, str]]] = []
    for piece in text.split("chunk_id:"):
        if not piece.strip():
            continue
        title = _first_field(_TITLE, piece, DEFAULT_TITLE)
        content = _content_of(piece)
        tags = _first_field(_TAGS, piece, "")
        if title or content:
            out.append((f"问题：{title}\n答案：{content}",
                        {"title": title, "tags": tags, "source": SOURCE_NAME}))
    return out


def parse_custom_format(file_path: str) -> List[Tuple[str, Dict[str, str]]]:
    if not os.path.exists(file_path):
        return []
    with open(file_path, "r", encoding="utf-8") as f:
        return parse_text(f.read())
