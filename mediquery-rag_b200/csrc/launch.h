// launch.h -- host-callable launchers of the libmqvs kernels (defined in the kernels_*.cu files).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mqvs {

// sets the thread-local message mqvs_last_error() returns; returns `code` (api.cu)
int mqvs_fail(int code, const char* fmt, ...);

// K1: L2-normalise + pack.  x [n, dim] fp32 (device) -> out [n, 768] fp32 or bf16, zero-padded; counts
// rows whose squared norm is not finite into *bad (and stores them as zeros).  norms (optional, [n]) receives ||x|| of
// every raw row (what the l2 / ip spaces rank with); max_dev (optional, two words) is raised to the bit patterns of the
// largest | ||x|| - 1 | seen ([0], atomicMax: tells the host whether the collection consists of unit vectors) and of the
// largest ||x|| ([1]: scales the tensor-core error bound in those spaces).
void launch_normalize(const float* x, uint64_t n, int dim, int out_dtype, void* out, unsigned long long* bad,
                      cudaStream_t st, float* norms = nullptr, unsigned int* max_dev = nullptr);
// Same K1 fed by the counter-based generator instead of memory.
void launch_synth_normalize(uint64_t seed, uint64_t first_index, uint64_t n, int dim, int out_dtype, void* out,
                            unsigned long long* bad, cudaStream_t st, float* norms = nullptr,
                            unsigned int* max_dev = nullptr);

// (largest, smallest) ||x|| of the 32-row chunks [first_chunk, ceil(n_rows / 32)): what lets the tensor epilogue of the
// l2 / ip spaces reject a chunk without transforming its scores (tc_device.cuh:chunk_score_bound).
void launch_norm_bounds(const float* norms, uint64_t n_rows, uint64_t first_chunk, float* bounds, cudaStream_t st);

// bf16 scoring shadow of an fp32 store: rows [first, first + n) of `rows` ([*, 768] fp32, normalised) -> the same rows of
// `shadow` ([*, 768] bf16, round to nearest); *err_bits is raised (atomicMax on the bit pattern) to the largest
// ||x - bf16(x)|| among them, rounded up.
void launch_shadow_build(const float* rows, uint64_t first, uint64_t n, void* shadow, unsigned int* err_bits, cudaStream_t st);

// K2 (SIMT) + K3: exact fp32 scan of `n_rows` stored rows against nq (<= scan_max_queries_per_pass) normalised
// queries, fused per-warp top-k with optional allow-mask; writes per-CTA sorted candidate keys
// cand[q][cta][k].  Returns the grid size used (number of CTAs) so the merge knows the layout.
// scan_plan: how many of the `nq_remaining` queries the next pass takes and the grid it runs with;
// scan_max_grid: upper bound of that grid over all passes (the candidate buffer's per-query stride).
void scan_plan(int dtype, int nq_remaining, int k, int sm_count, int* nq_pass, int* grid);
int scan_max_grid(int sm_count);
void launch_scan_simt(const void* rows, int dtype, uint32_t n_rows, const float* qn, int nq, const uint32_t* mask,
                      int k, uint64_t* cand, int cand_stride, int grid, cudaStream_t st);
// The same scan ranking by the l2 / ip space of the RAW vectors (mqvs_set_metric): row_norms [n_rows] as K1 stored them,
// q_norms [nq] = ||q|| of the raw queries (K1 again); scores are common.cuh:rank_score.
void launch_scan_simt_metric(const void* rows, int dtype, uint32_t n_rows, const float* qn, int nq, const uint32_t* mask,
                             int k, uint64_t* cand, int cand_stride, int grid, int metric, const float* row_norms,
                             const float* q_norms, cudaStream_t st);

// Device-driven repair of tensor-path queries flagged "not proven exact": scan_fix compacts the flags into a list itself
// (first max_fix <= 256 flagged queries, published in fix_list / fix_n; counters[0] += listed, [1] += left over when
// count_leftover is set -- the last round of a search), re-scans
// the corpus exactly for the listed queries (cand [max_fix][scan_fix_grid][k]); merge_fix merges into the outputs and
// clears their flags.
int scan_fix_grid(int sm_count);  // CTAs (= candidate stride per listed query) of the repair scan
void launch_scan_fix(const void* rows, int dtype, uint32_t n_rows, const float* qn_all, const uint32_t* flags, int B, int max_fix,
                     uint32_t* fix_list, uint32_t* fix_n, unsigned long long* counters, const uint32_t* mask, int k, uint64_t* cand,
                     int sm_count, cudaStream_t st, int metric, const float* row_norms, const float* q_norms_all,
                     bool count_leftover, const uint32_t* flag_count);
void launch_merge_fix(const uint64_t* keys_in, int max_fix, int m, int k, int64_t id_base, const uint32_t* fix_list,
                      const uint32_t* fix_n, uint32_t* flags, float* out_scores, int64_t* out_ids, cudaStream_t st,
                      uint32_t* flag_count);

// Latency path: the whole search (query normalisation, scan, top-k, final merge) in ONE launch, for B <= 4, k <= 16.
// cand: scratch of B * scan_fused_max_grid * k keys (no memset needed); ticket: device uint32, zero between calls;
// bad_epoch: device uint32 raised to `epoch` when a query is non-finite.  out_scores / out_ids may be MAPPED HOST memory
// (the last CTA writes them with plain stores); host_status (mapped host, optional) then receives {epoch, non-finite flag}.
// q_host (optional): the raw queries in HOST memory; they are then carried inside the kernel's parameter block instead of
// being read from q_raw (no host-to-device copy of the queries at all).
bool scan_fused_ok(int B, int k);
int scan_fused_max_grid(int sm_count);
void launch_scan_fused(const void* rows, int dtype, uint32_t n_rows, const float* q_raw, int dim, int nq,
                       const uint32_t* mask, int k, uint64_t* cand, float* out_scores, int64_t* out_ids, int64_t id_base,
                       unsigned int* ticket, unsigned int* bad_epoch, unsigned int epoch, int sm_count, cudaStream_t st,
                       unsigned int* host_status = nullptr, const float* q_host = nullptr);

// Merge: keys_in [B][m] (any order, 0 = empty) -> best k per query as (score, id + id_base); (-inf, -1) fill.
void launch_merge_keys(const uint64_t* keys_in, int B, int m, int k, int64_t id_base, float* out_scores,
                       int64_t* out_ids, cudaStream_t st);
// Cross-shard merge of already-decoded candidates [W, B, k] -> [B, k].  shard_shift >= 0: the ids of slice w are LOCAL rows
// of part w of a block-cyclic multi-device handle (store.h:ShardMap, blocks of 1 << shard_shift rows) and are mapped to
// global rows before the merge; id_base is added to every id written.
void launch_merge_candidates(const float* scores, const int64_t* ids, int W, int B, int k, float* out_scores,
                             int64_t* out_ids, cudaStream_t st, int shard_shift = -1, int64_t id_base = 0);
// Gather stored rows bit-for-bit: out[i] = rows[ids[i]] (row_bytes each, 16-byte vectors).  Tombstone compaction.
void launch_gather_raw_rows(const void* rows, size_t row_bytes, const uint32_t* ids, uint64_t n, void* out, cudaStream_t st);
// Gather stored rows as fp32 [n, dim].
void launch_gather_rows(const void* rows, int dtype, const int64_t* ids, uint64_t n, int dim, float* out,
                        cudaStream_t st);

}  // namespace mqvs
