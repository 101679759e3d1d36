"""Value types the reference passes across the vector-store boundary.

`Document` is what /root/reference/src/ingest_medical.py:73-80 builds and what src/agents/nodes.py:94 and
src/medical_engine.py:75-78 read (`.page_content`, `.metadata.get('title')`); `Embeddings` is the protocol of
OllamaEmbeddings at src/medical_engine.py:43.  When langchain_core is installed its own classes are used, so
objects flow between the reference and this backend unchanged; otherwise minimal stand-ins with the same
attributes are provided (this image has no langchain_core).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

try:  # pragma: no cover - langchain_core is absent in the build image
    from langchain_core.documents import Document  # type: ignore
    from langchain_core.embeddings import Embeddings  # type: ignore

    HAVE_LANGCHAIN = True
except Exception:  # ImportError or a broken install
    HAVE_LANGCHAIN = False

    class Document:  # type: ignore[no-redef]
        """page_content + metadata (+ optional id), keyword-constructible like langchain_core's."""

        __slots__ = ("page_content", "metadata", "id")

        def __init__(self, page_content: str, metadata: Optional[Dict[str, Any]] = None, id: Optional[str] = None,
                     **_: Any):
            self.page_content = page_content
            self.metadata = dict(metadata) if metadata else {}
            self.id = id

        def __repr__(self) -> str:
            return f"Document(id={self.id!r}, page_content={self.page_content!r}, metadata={self.metadata!r})"

        def __eq__(self, other: object) -> bool:
            return (isinstance(other, Document) and self.page_content == other.page_content
                    and self.metadata == other.metadata and self.id == other.id)

    class Embeddings:  # type: ignore[no-redef]
        """embed_documents(list[str]) -> list[list[float]]; embed_query(str) -> list[float]."""

        def embed_documents(self, texts: List[str]) -> List[List[float]]:
            raise NotImplementedError

        def embed_query(self, text: str) -> List[float]:
            raise NotImplementedError
