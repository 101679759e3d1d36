// tc.h -- host interface of the tensor-core search path (kernels_tc.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mqvs {
namespace tc {

struct Plan {
    int kprime;  // shortlist size per query (k' > k): rescored exactly
    int cap;     // per-(CTA, query) candidate list capacity
    int b_pad;   // queries padded to a multiple of 128
    int grid;    // persistent CTAs (one per SM)
};

struct Buffers {       // all device memory, owned by the store handle
    void* q_op;        // [b_pad, 768] query operand matrix (bf16 or fp32)
    float* eps;        // [B]
    uint32_t* tau;     // [b_pad]
    uint32_t* cnt;     // [grid, b_pad]
    uint64_t* lists;   // [grid, b_pad, cap]
    uint64_t* short_keys;  // [b_pad, kprime]
    uint64_t* exact_keys;  // [b_pad, kprime]
    float* tau_tilde;  // [b_pad]
    uint32_t* flags;   // [B] 1 = exactness margin not proven, needs the exact scan
    uint32_t* flag_count;
    uint32_t* cmax;    // single-launch pipeline: [128][grid * 8] chunk maxima of every CTA's first tile
    unsigned int* gbar;  // single-launch pipeline: [3] grid barrier counters
    uint32_t* band;       // [B] band-pass threshold of every flagged query: ord(k-th exact score - eps)
    uint32_t* band_list;  // [128] flagged queries, compacted
    uint32_t* band_n;     // [1]
    float* hq;            // l2 / ip spaces: [B padded to 128] 0.5 / max(||q||, eps) per query (0 in the ip space)
    float* h_op;          // [256] the same per operand row of the band pass (compacted queries)
};

// Chroma's l2 / ip spaces over raw (non-unit) vectors on the tensor pipelines (common.cuh:rank_score; mqvs_set_metric)
struct Space {
    int metric = 0;                // MQVS_METRIC_*; 0 = cosine: the other fields are unused
    const float* norms = nullptr;  // [n_rows] ||x|| of the raw rows (K1)
    const float* norm_bounds = nullptr;  // [ceil(n_rows / 32)][2] largest and smallest norm of every 32-row chunk
    float amax = 1.0f;             // largest ||x||: the tensor-core error bound of a ranking score is amax * eps
    float* qnorm = nullptr;        // [B] out: ||q|| of the raw queries (the exact-scan repair ranks with them)
    // bf16 scoring shadow of an fp32 store: the `rows` / `dtype` arguments of the searches below are then the SHADOW (what
    // the tensor cores stream), and everything exact -- K4 rescoring, the tail of the single-launch kernel -- reads these:
    const void* exact_rows = nullptr;         // fp32 rows [n_rows, 768], or nullptr = `rows` are the stored rows themselves
    const unsigned int* shadow_err = nullptr; // device: bits of the largest ||x - bf16(x)|| over the rows (added to eps)
};

Plan plan(int k, int B, int sm_count);
size_t op_bytes(const Plan& p, int dtype);
size_t lists_bytes(const Plan& p);
size_t cnt_bytes(const Plan& p);

// Enqueue the whole tensor-core search on `st`.  Returns nullptr or an error message.
const char* search(const Buffers& b, const Plan& p, const void* rows, int dtype, uint32_t n_rows, const float* qn, int B,
                   int k, const uint32_t* mask, int64_t id_base, float* out_scores, int64_t* out_ids, cudaStream_t st,
                   const Space& space = Space());

// Single-launch pipeline for B <= 256 (one or two query tiles): prep kernel + one cooperative persistent kernel that learns the
// thresholds, scans, selects, rescores and finalizes.  Takes the RAW queries (normalises them into qn itself).
// *coop_refused: the cooperative launch was rejected (not enough co-resident CTAs): use search() instead.
bool fused_ok(int B, int k, int sm_count);
size_t cmax_bytes(const Plan& p);
const char* search_fused(const Buffers& b, const Plan& p, const void* rows, int dtype, uint32_t n_rows, const float* q_raw,
                         int dim, float* qn, int B, int k, const uint32_t* mask, int64_t id_base, float* out_scores,
                         int64_t* out_ids, unsigned int* bad_epoch, unsigned int epoch, cudaStream_t st, bool* coop_refused,
                         const Space& space = Space());

// Repair of flagged queries by a second single-launch pass with given thresholds (kernels_tc_fused.cu); enqueue after
// search() / search_fused() and before the exact-scan repair.  counters[2] counts the queries it repaired.
const char* search_band(const Buffers& b, const Plan& p, const void* rows, int dtype, uint32_t n_rows, const float* qn, int B,
                        int k, const uint32_t* mask, int64_t id_base, float* out_scores, int64_t* out_ids,
                        unsigned long long* counters, cudaStream_t st, const Space& space = Space());

void trace_dump();  // MQVS_TRACE=1 developer aid: per-launch timings of the last searches to stderr
int launches_per_search(const Plan& p, uint32_t n_rows, int B);  // prep + (score, select) per segment + rescore + finalize

}  // namespace tc
}  // namespace mqvs
