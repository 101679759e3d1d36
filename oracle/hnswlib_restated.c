/* oracle/hnswlib_restated.c -- TEST INFRASTRUCTURE ONLY (nothing under mediquery-rag_b200/ may link, load or call this).
 *
 * The reference's retrieval arithmetic is not in /root/reference: src/medical_engine.py:52 / src/ingest_medical.py:106-110 /
 * src/agents/nodes.py:93 hand it to langchain-chroma -> chromadb -> chroma-hnswlib (README.md:630-639 floors chromadb >= 0.4.0;
 * chromadb 0.4.15 - 0.5.x pin chroma-hnswlib == 0.7.3; neither package is vendored or installable here).  This file restates
 * the PUBLISHED arithmetic of that dependency from upstream knowledge -- not its graph search, which north_star excludes
 * (exact results required), but what every distance it ever evaluates is, and what its exhaustive index returns:
 *
 *   hnswlib/space_ip.h   InnerProduct():          res = 0; for i: res += a[i] * b[i]           (fp32, sequential)
 *                        InnerProductDistance():  1.0f - InnerProduct()
 *                        InnerProductSIMD16Ext*(): 16 fp32 lanes accumulate a[16j + l] * b[16j + l] (mul, then add: no FMA),
 *                                                  lanes summed l = 0 .. 15 in order            (dim % 16 == 0: 768 is)
 *   hnswlib/space_l2.h   L2Sqr():                 res = 0; for i: t = a[i] - b[i]; res += t * t
 *                        L2SqrSIMD16Ext*():       the same lane structure
 *   python_bindings      Index(space="cosine"):   InnerProductSpace + normalize_vector() on every add and query:
 *                                                  norm = 0; for i: norm += x[i] * x[i]; norm = 1 / (sqrtf(norm) + 1e-30f);
 *                                                  y[i] = x[i] * norm
 *   hnswlib/bruteforce.h BruteforceSearch::searchKnn(): max-heap of (distance, label); the first k rows enter, a later
 *                                                  row enters when dist <= the current worst and the worst leaves.
 *
 * The SSE / AVX / AVX-512 variants differ in lane count (4 / 8 / 16 partial sums); the 16-lane form is restated because it is
 * what a host with AVX-512 runs and it is the farthest from sequential order.  tests/test_oracle.py holds the canonical oracle
 * (oracle/canonical.c) and tests/test_gpu_parity.py the CUDA path against BOTH orders under north_star's tolerance (ids
 * identical outside near ties, distances within 1e-5 relative): different summation orders of 768 fp32 products agree to ~1e-7.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

float hnsw_ip_scalar(const float* a, const float* b, int64_t dim) {
    float res = 0.0f;
    for (int64_t i = 0; i < dim; ++i) res += a[i] * b[i];
    return res;
}

float hnsw_ip_simd16(const float* a, const float* b, int64_t dim) {
    float lane[16] = {0};
    const int64_t d16 = dim / 16 * 16;
    for (int64_t j = 0; j < d16; j += 16)
        for (int l = 0; l < 16; ++l) lane[l] = lane[l] + a[j + l] * b[j + l];
    float res = 0.0f;
    for (int l = 0; l < 16; ++l) res += lane[l];
    for (int64_t i = d16; i < dim; ++i) res += a[i] * b[i];  /* ...Residuals: scalar tail */
    return res;
}

float hnsw_l2_scalar(const float* a, const float* b, int64_t dim) {
    float res = 0.0f;
    for (int64_t i = 0; i < dim; ++i) {
        const float t = a[i] - b[i];
        res += t * t;
    }
    return res;
}

float hnsw_l2_simd16(const float* a, const float* b, int64_t dim) {
    float lane[16] = {0};
    const int64_t d16 = dim / 16 * 16;
    for (int64_t j = 0; j < d16; j += 16)
        for (int l = 0; l < 16; ++l) {
            const float t = a[j + l] - b[j + l];
            lane[l] = lane[l] + t * t;
        }
    float res = 0.0f;
    for (int l = 0; l < 16; ++l) res += lane[l];
    for (int64_t i = d16; i < dim; ++i) {
        const float t = a[i] - b[i];
        res += t * t;
    }
    return res;
}

/* python_bindings/bindings.cpp normalize_vector() */
void hnsw_normalize(const float* x, float* y, int64_t dim) {
    float norm = 0.0f;
    for (int64_t i = 0; i < dim; ++i) norm += x[i] * x[i];
    norm = 1.0f / (sqrtf(norm) + 1e-30f);
    for (int64_t i = 0; i < dim; ++i) y[i] = x[i] * norm;
}

/* space: 0 = "l2", 1 = "ip", 2 = "cosine"; order: 0 = scalar, 1 = 16-lane.  rows [n, dim] RAW fp32 vectors as handed to
 * collection.add(); q [dim] a raw query.  -> distance of every row (what hnswlib's dist_func_ returns, lower = better). */
void hnsw_distances(int space, int order, const float* rows, int64_t n, int64_t dim, const float* q, float* out) {
    float* qn = (float*)malloc(sizeof(float) * (size_t)dim * 2);
    float* xn = qn + dim;
    if (space == 2) hnsw_normalize(q, qn, dim);
    for (int64_t r = 0; r < n; ++r) {
        const float* x = rows + r * dim;
        const float* qq = q;
        if (space == 2) {
            hnsw_normalize(x, xn, dim);
            x = xn;
            qq = qn;
        }
        if (space == 0) out[r] = order ? hnsw_l2_simd16(x, qq, dim) : hnsw_l2_scalar(x, qq, dim);
        else out[r] = 1.0f - (order ? hnsw_ip_simd16(x, qq, dim) : hnsw_ip_scalar(x, qq, dim));
    }
    free(qn);
}

/* bruteforce.h searchKnn over precomputed distances: binary max-heap keyed by (distance, label) like
 * std::priority_queue<std::pair<float, labeltype>>.  allow: optional bitmask (bit r & 31 of word r >> 5), the filter functor.
 * -> labels / distances best first (the python binding reverses the heap's pop order), -1 / +inf padded. */
typedef struct { float d; int64_t id; } hn_t;
static int hn_less(hn_t a, hn_t b) { return a.d < b.d || (a.d == b.d && a.id < b.id); }
static void sift_up(hn_t* h, int i) {
    while (i > 0) {
        const int p = (i - 1) / 2;
        if (!hn_less(h[p], h[i])) break;
        const hn_t t = h[p]; h[p] = h[i]; h[i] = t;
        i = p;
    }
}
static void sift_down(hn_t* h, int n, int i) {
    for (;;) {
        int m = i;
        const int l = 2 * i + 1, r = l + 1;
        if (l < n && hn_less(h[m], h[l])) m = l;
        if (r < n && hn_less(h[m], h[r])) m = r;
        if (m == i) return;
        const hn_t t = h[m]; h[m] = h[i]; h[i] = t;
        i = m;
    }
}
void hnsw_bruteforce_knn(const float* dist, int64_t n, const uint32_t* allow, int k, float* out_d, int64_t* out_id) {
    hn_t* h = (hn_t*)malloc(sizeof(hn_t) * (size_t)(k + 1));
    int sz = 0;
    float lastdist = INFINITY;
    for (int64_t r = 0; r < n; ++r) {
        if (allow && !((allow[r >> 5] >> (r & 31)) & 1u)) continue;
        if (sz < k) {
            h[sz].d = dist[r]; h[sz].id = r; sift_up(h, sz); ++sz;
            if (sz == k) lastdist = h[0].d;
        } else if (dist[r] <= lastdist) {
            h[sz].d = dist[r]; h[sz].id = r; sift_up(h, sz); ++sz;
            h[0] = h[--sz]; sift_down(h, sz, 0);  /* pop the worst */
            lastdist = h[0].d;
        }
    }
    for (int i = 0; i < k; ++i) { out_d[i] = INFINITY; out_id[i] = -1; }
    for (int i = sz - 1; i >= 0; --i) {  /* heap pops worst first */
        out_d[i] = h[0].d; out_id[i] = h[0].id;
        h[0] = h[--sz]; sift_down(h, sz, 0);
    }
    free(h);
}
