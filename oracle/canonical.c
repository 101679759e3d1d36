/*
 * oracle/canonical.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the arithmetic on MediQuery-RAG's retrieval path, used only by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the checker.
 * Nothing under mediquery-rag_b200/ may import, link or execute this file.
 *
 * PARITY UNPINNED: the reference (/root/reference) does none of this arithmetic itself.  Its three call
 * sites -- Chroma.from_documents (src/ingest_medical.py:106-110), Chroma(persist_directory=...)
 * (src/medical_engine.py:52) and vectorstore.similarity_search(q, k=5) (src/agents/nodes.py:93;
 * k=3 at src/medical_engine.py:70) -- hand the work to langchain-chroma -> chromadb -> chroma-hnswlib,
 * none of which is vendored, pinned (README.md:630-639 gives only ">=" floors) or installable here, and
 * the reference has no tests or golden vectors.  What is restated is therefore the published definition
 * those packages implement on unit vectors, and what BASELINE.json's north_star prescribes as referee:
 *
 *     stored row   x^ = x / max(||x||, eps)          (Ollama's /api/embed returns L2-normalised rows)
 *     score(q, x)  = <q^, x^>  in fp32               (chroma: distance = 1 - <q,x> for "cosine",
 *                                                      ||q-x||^2 = 2 - 2<q,x> for the default "l2")
 *     result       = k best rows, score descending, ties -> lower row id, filter mask applied first
 *
 * fp32 addition is not associative, so "exact fp32" needs an order.  The canonical order below is the
 * one the CUDA kernels use (one warp per row, lane l owns elements j with (j mod 128) div 4 == l,
 * each lane accumulates its elements in ascending j with fused multiply-add, then a 5-level xor
 * butterfly).  Restated here with fmaf() it is bit-identical to the GPU, which is what lets the
 * parity tests demand identical ids and identical score bits; tests/test_oracle.py separately checks this
 * canonical fp32 against an fp64 referee and numpy's sgemm to <= 1e-5 relative.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_LANES 32
#define ORC_EPS 1e-12f

/* round-to-nearest-even fp32 -> bf16 bits (finite inputs; NaN kept quiet) */
static inline uint16_t f32_to_bf16_rne(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x0040u);
    uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb;
    return (uint16_t)(u >> 16);
}
static inline float bf16_to_f32(uint16_t h) {
    uint32_t u = ((uint32_t)h) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

/* canonical fp32 dot over dim_pad (multiple of 128) elements */
static inline float canonical_dot_f32(const float* a, const float* b, int dim_pad) {
    float p[ORC_LANES];
    for (int l = 0; l < ORC_LANES; ++l) {
        float acc = 0.0f;
        for (int c = 0; c < dim_pad; c += 128)
            for (int e = 0; e < 4; ++e) {
                int j = c + l * 4 + e;
                acc = fmaf(a[j], b[j], acc);
            }
        p[l] = acc;
    }
    for (int m = 16; m >= 1; m >>= 1) {
        float t[ORC_LANES];
        for (int l = 0; l < ORC_LANES; ++l) t[l] = p[l] + p[l ^ m];
        memcpy(p, t, sizeof(p));
    }
    return p[0];
}

static inline float canonical_dot_f32_bf16(const float* q, const uint16_t* x, int dim_pad) {
    float p[ORC_LANES];
    for (int l = 0; l < ORC_LANES; ++l) {
        float acc = 0.0f;
        for (int c = 0; c < dim_pad; c += 128)
            for (int e = 0; e < 4; ++e) {
                int j = c + l * 4 + e;
                acc = fmaf(q[j], bf16_to_f32(x[j]), acc);
            }
        p[l] = acc;
    }
    for (int m = 16; m >= 1; m >>= 1) {
        float t[ORC_LANES];
        for (int l = 0; l < ORC_LANES; ++l) t[l] = p[l] + p[l ^ m];
        memcpy(p, t, sizeof(p));
    }
    return p[0];
}

/* order-preserving key: larger key == better (score desc, then id asc); 0 == empty slot */
static inline uint64_t pack_key(float s, uint32_t id) {
    if (s == 0.0f) s = 0.0f; /* -0 -> +0 */
    uint32_t u;
    memcpy(&u, &s, 4);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((uint64_t)u << 32) | (uint64_t)(0xffffffffu - id);
}

/* ------------------------------------------------------------------------------------------------ */
/* K1 restated: x[n, dim] fp32 -> rows normalised, zero-padded to dim_pad; returns # non-finite rows   */
int64_t orc_normalize_f32(const float* x, int64_t n, int dim, int dim_pad, float* out) {
    int64_t bad = 0;
    for (int64_t r = 0; r < n; ++r) {
        float* o = out + r * (int64_t)dim_pad;
        memcpy(o, x + r * (int64_t)dim, sizeof(float) * (size_t)dim);
        for (int j = dim; j < dim_pad; ++j) o[j] = 0.0f;
        float ss = canonical_dot_f32(o, o, dim_pad);
        if (!isfinite(ss)) {
            bad += 1;
            for (int j = 0; j < dim_pad; ++j) o[j] = 0.0f;
            continue;
        }
        float inv = 1.0f / fmaxf(sqrtf(ss), ORC_EPS);
        for (int j = 0; j < dim; ++j) o[j] = o[j] * inv;
    }
    return bad;
}

int64_t orc_normalize_bf16(const float* x, int64_t n, int dim, int dim_pad, uint16_t* out) {
    int64_t bad = 0;
    for (int64_t r = 0; r < n; ++r) {
        float tmp[4096];
        uint16_t* o = out + r * (int64_t)dim_pad;
        memcpy(tmp, x + r * (int64_t)dim, sizeof(float) * (size_t)dim);
        for (int j = dim; j < dim_pad; ++j) tmp[j] = 0.0f;
        float ss = canonical_dot_f32(tmp, tmp, dim_pad);
        if (!isfinite(ss)) {
            bad += 1;
            for (int j = 0; j < dim_pad; ++j) o[j] = 0;
            continue;
        }
        float inv = 1.0f / fmaxf(sqrtf(ss), ORC_EPS);
        for (int j = 0; j < dim_pad; ++j) o[j] = f32_to_bf16_rne(tmp[j] * inv);
    }
    return bad;
}

/* ||x|| of the raw rows as K1 measures it: sqrtf of the canonical sum of squares (0 for a non-finite row).
 * Needed by Chroma's "l2" / "ip" spaces when the raw vectors are not unit vectors (langchain_chroma passes
 * collection_metadata={"hnsw:space": ...} through; the reference passes none -> "l2", src/medical_engine.py:52). */
void orc_norms_f32(const float* x, int64_t n, int dim, int dim_pad, float* norms) {
    float tmp[4096];
    for (int64_t r = 0; r < n; ++r) {
        memcpy(tmp, x + r * (int64_t)dim, sizeof(float) * (size_t)dim);
        for (int j = dim; j < dim_pad; ++j) tmp[j] = 0.0f;
        float ss = canonical_dot_f32(tmp, tmp, dim_pad);
        norms[r] = isfinite(ss) ? sqrtf(ss) : 0.0f;
    }
}

/* Ranking score of a row in the collection's space (metric 0 = cosine, 1 = l2, 2 = ip), larger = better:
 * s = canonical cosine of the normalised pair, a = ||x|| (raw row), h = 0.5 / max(||q||, eps):
 *   cosine s;   ip a*s = <q,x>/||q||;   l2 a*s - (a*a)*h = (||q||^2 - ||q-x||^2) / (2||q||)   -- chromadb's
 * "l2" is the squared distance, "ip" is 1 - <q,x>; both are monotone in these scores for a fixed query.
 * Every product is rounded to fp32 on its own and the last step is one fused multiply-add (the kernels' order). */
float orc_rank_score(int metric, float s, float a, float h) {
    if (metric == 2) return a * s;
    if (metric == 1) {
        float aa = a * a;
        float w = aa * h;
        return fmaf(a, s, -w);
    }
    return s;
}

/* scores[b, r] = canonical <q_b, x_r>; store_dtype 0 = fp32 rows, 1 = bf16 rows */
void orc_scores(const float* q, int B, const void* rows, int store_dtype, int64_t n, int dim_pad,
                float* scores) {
    for (int64_t r = 0; r < n; ++r) {
        for (int b = 0; b < B; ++b) {
            const float* qb = q + (int64_t)b * dim_pad;
            scores[(int64_t)b * n + r] =
                store_dtype == 0
                    ? canonical_dot_f32(qb, (const float*)rows + r * (int64_t)dim_pad, dim_pad)
                    : canonical_dot_f32_bf16(qb, (const uint16_t*)rows + r * (int64_t)dim_pad, dim_pad);
        }
    }
}

static int cmp_key_desc(const void* a, const void* b) {
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? 1 : (x > y ? -1 : 0);
}

/*
 * Exact brute-force top-k: out_scores/out_ids [B, k], best first, ties -> lower id, rows whose mask bit
 * is 0 are skipped (mask == NULL: all allowed), missing slots = (-inf, -1).  id_base is added to ids.
 */
void orc_topk_metric(const float* q, int B, const void* rows, int store_dtype, int64_t n, int dim_pad,
                     const uint32_t* mask, int k, int64_t id_base, int metric, const float* row_norms,
                     const float* q_norms, float* out_scores, int64_t* out_ids);

void orc_topk(const float* q, int B, const void* rows, int store_dtype, int64_t n, int dim_pad,
              const uint32_t* mask, int k, int64_t id_base, float* out_scores, int64_t* out_ids) {
    orc_topk_metric(q, B, rows, store_dtype, n, dim_pad, mask, k, id_base, 0, NULL, NULL, out_scores, out_ids);
}

/* The same in the collection's space: q = NORMALISED queries, q_norms [B] = ||q|| of the raw queries, row_norms [n] =
 * ||x|| of the raw rows (both NULL for cosine); scores out are orc_rank_score. */
void orc_topk_metric(const float* q, int B, const void* rows, int store_dtype, int64_t n, int dim_pad,
                     const uint32_t* mask, int k, int64_t id_base, int metric, const float* row_norms,
                     const float* q_norms, float* out_scores, int64_t* out_ids) {
    for (int b = 0; b < B; ++b) {
        const float* qb = q + (int64_t)b * dim_pad;
        const float h = metric ? 0.5f / fmaxf(q_norms[b], ORC_EPS) : 0.0f;
        uint64_t* keys = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1));
        int64_t m = 0;
        for (int64_t r = 0; r < n; ++r) {
            if (mask && !((mask[r >> 5] >> (r & 31)) & 1u)) continue;
            float s = store_dtype == 0
                          ? canonical_dot_f32(qb, (const float*)rows + r * (int64_t)dim_pad, dim_pad)
                          : canonical_dot_f32_bf16(qb, (const uint16_t*)rows + r * (int64_t)dim_pad, dim_pad);
            if (metric) s = orc_rank_score(metric, s, row_norms[r], h);
            keys[m++] = pack_key(s, (uint32_t)r);
        }
        qsort(keys, (size_t)m, sizeof(uint64_t), cmp_key_desc);
        for (int j = 0; j < k; ++j) {
            if (j < m) {
                uint32_t u = (uint32_t)(keys[j] >> 32);
                u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
                float s;
                memcpy(&s, &u, 4);
                out_scores[(int64_t)b * k + j] = s;
                out_ids[(int64_t)b * k + j] = id_base + (int64_t)(0xffffffffu - (uint32_t)keys[j]);
            } else {
                out_scores[(int64_t)b * k + j] = -INFINITY;
                out_ids[(int64_t)b * k + j] = -1;
            }
        }
        free(keys);
    }
}

/* counter-based synthetic rows: v[r, c] = uniform[-1, 1) from splitmix64(seed, r * dim + c) */
static inline uint64_t splitmix64(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
void orc_synth_rows(uint64_t seed, int64_t first_row, int64_t n, int dim, float* out) {
    for (int64_t r = 0; r < n; ++r)
        for (int c = 0; c < dim; ++c) {
            uint64_t z = splitmix64(seed, (uint64_t)(first_row + r) * (uint64_t)dim + (uint64_t)c);
            out[r * (int64_t)dim + c] = (float)(z >> 40) * (1.0f / 8388608.0f) - 1.0f;
        }
}
