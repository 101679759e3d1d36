"""oracle/ -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference's retrieval path).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package, and only as the checker.  The product (mediquery-rag_b200/) never does: it fails loudly when
its CUDA library is missing instead of falling back to anything here.

PARITY UNPINNED for the vector arithmetic: the reference delegates it to un-vendored, un-pinned
langchain-chroma / chromadb / chroma-hnswlib (README.md:630-639) and has no tests or golden vectors, so
the referee is the exact brute-force definition BASELINE.json's north_star prescribes (see canonical.c);
hnswlib_restated.c restates the published arithmetic of that dependency (chroma-hnswlib's distance functions
and exhaustive search, from upstream knowledge) as a second, independently ordered checker.
The text splitter (src/ingest_medical.py:11-87) IS pinned: tests/golden/ holds outputs produced by
importing the reference's own parse_custom_format (oracle/gen_golden.py).
"""
from .oracle import (  # noqa: F401
    DIM_ALIGN,
    brute_force_topk,
    canonical_scores,
    distances_from_rank_scores,
    eval_where,
    f64_scores,
    hnswlib_topk,
    lib,
    mask_from_bool,
    merge_topk,
    normalize_rows,
    pad_dim,
    planted_queries,
    referee_topk_f64,
    row_norms,
    sgemm_topk,
    stored_as_f32,
    synth_rows,
)
from .splitter import parse_custom_format  # noqa: F401
