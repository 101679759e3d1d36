"""mediquery-rag_b200 -- B200-native exact vector search behind MediQuery-RAG's `vectorstore`.

The directory name carries a hyphen (it is the name the build contract gives the package); import it as
`mediquery_rag_b200` (the shim package next to it extends its __path__ to this directory).

    from mediquery_rag_b200 import MediQueryVectorStore as Chroma      # the one-line swap in the reference
"""
from . import _native  # noqa: F401
from ._native import Index, MqvsError, merge_candidates  # noqa: F401
from .documents import Document, Embeddings  # noqa: F401
from .filters import FilterError  # noqa: F401
from .vectorstore import MediQueryRetriever, MediQueryVectorStore  # noqa: F401

__all__ = ["Index", "MqvsError", "merge_candidates", "Document", "Embeddings", "FilterError",
           "MediQueryVectorStore", "MediQueryRetriever"]
