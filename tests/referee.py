"""An arithmetic referee that owes nothing to the kernels' accumulation order: float64 numpy over the STORED values, chunked so
that a million rows fit in memory.  Distances follow chromadb's published definitions (cosine: 1 - <q^, x^>; l2: squared
euclidean ||q - x||^2; ip: 1 - <q, x>) for the vectors the backend represents (x = ||x|| * stored unit row)."""
import numpy as np

import oracle as O


def f64_topk(q_raw, stored, norms, k, space, chunk=50_000):
    """-> (distance [B, k] float64 ascending, ids [B, k]); ties -> lower id.  q_raw: raw fp32 queries [B, 768]."""
    q = q_raw.astype(np.float64)
    qn = np.linalg.norm(q, axis=1)
    qh = q / np.maximum(qn, 1e-12)[:, None]
    n, B = stored.shape[0], q.shape[0]
    best_d = np.full((B, 0), np.inf)
    best_i = np.zeros((B, 0), np.int64)
    for a in range(0, n, chunk):
        xs = O.stored_as_f32(stored[a:a + chunk]).astype(np.float64)
        cos = qh @ xs.T                                     # [B, m]
        if space == "cosine":
            d = 1.0 - cos
        else:
            an = norms[a:a + chunk].astype(np.float64)[None, :]
            dot = cos * an * qn[:, None]                    # <q, x>
            d = 1.0 - dot if space == "ip" else qn[:, None] ** 2 + an ** 2 - 2.0 * dot
        ids = np.broadcast_to(np.arange(a, a + xs.shape[0])[None, :], d.shape)
        d = np.concatenate([best_d, d], axis=1)
        ids = np.concatenate([best_i, ids], axis=1)
        order = np.lexsort((ids, d), axis=1)[:, :k]
        best_d, best_i = np.take_along_axis(d, order, 1), np.take_along_axis(ids, order, 1)
    return best_d, best_i


def assert_within_north_star(space, got_scores, got_ids, q_raw, ref_d, ref_i, rel=1e-5):
    """north_star's bar: ids identical apart from documented tie-breaking (two neighbours the fp32 result cannot separate),
    scores within 1e-5 relative.  got_scores are the backend's ranking scores (cosine, or common.cuh:rank_score)."""
    qn = O.row_norms(q_raw, 768)
    got_d = 1.0 - got_scores.astype(np.float64) if space == "cosine" else O.distances_from_rank_scores(space, got_scores, qn)
    # the distance is a difference of terms of this size: the relative tolerance applies to them (l2: ||q||^2 + ||x||^2)
    scale = 1.0 if space == "cosine" else (1.0 + np.abs(ref_d).max(axis=1, keepdims=True) + (qn.astype(np.float64) ** 2)[:, None])
    err = np.abs(got_d - ref_d) / scale
    assert err.max() <= rel, f"{space}: distance differs from the float64 referee by {err.max():.2e} (relative)"
    for row in np.argwhere(got_ids != ref_i)[:, 0]:
        a, b = got_ids[row], ref_i[row]
        gaps = np.abs(np.diff(ref_d[row]))
        same_set = set(a.tolist()) == set(b.tolist())
        # an id may differ only where the referee's own neighbouring distances are closer than fp32 resolves
        assert (same_set and gaps.min() <= 4e-7 * float(scale[row if scale.shape[0] > 1 else 0].max())) or \
            abs(ref_d[row][-1] - got_d[row][-1]) <= 4e-7 * float(np.max(scale)), f"{space}: ids differ beyond a near tie in row {row}"
