// store.h -- the handle behind mqvs_t (include/mqvs.h) and the internal interface between api.cu (one row shard on
// one device) and api_multi.cu (n_dev > 1: block-cyclic row shards on several devices driven by one host process).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/mqvs.h"
#include "filter.h"
#include "launch.h"
#include "tc.h"

namespace mqvs {

// One NVTX range per API call (visible in nsys timelines and to `ncu --nvtx --nvtx-include`): header-only NVTX3, a no-op
// (one load + branch) while no tool is attached to the process.
struct ApiRange {
    explicit ApiRange(const char* name) { nvtxRangePushA(name); }
    ~ApiRange() { nvtxRangePop(); }
    ApiRange(const ApiRange&) = delete;
    ApiRange& operator=(const ApiRange&) = delete;
};
#define MQVS_RANGE(name_) ::mqvs::ApiRange mqvs_api_range_(name_)

// thread-local error message + code (api.cu); every failure path goes through this
int mqvs_fail(int code, const char* fmt, ...);

#define MQVS_CU(expr)                                                                                          \
    do {                                                                                                       \
        cudaError_t e_ = (expr);                                                                               \
        if (e_ != cudaSuccess) {                                                                               \
            int code_ = (e_ == cudaErrorMemoryAllocation) ? MQVS_E_OOM : MQVS_E_CUDA;                          \
            cudaGetLastError();                                                                                \
            return ::mqvs::mqvs_fail(code_, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                                      \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return MQVS_OK;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        size_t want = std::max(need, (size_t)4096);
        MQVS_CU(cudaMalloc(&p, want));
        bytes = want;
        return MQVS_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
};

// pinned (page-locked, mapped) host block that grows on demand
struct PinBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return MQVS_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
        size_t want = std::max(need, (size_t)4096);
        MQVS_CU(cudaHostAlloc(&p, want, cudaHostAllocMapped | cudaHostAllocPortable));
        bytes = want;
        return MQVS_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
    }
};

inline size_t elem_size(int dtype) { return dtype == MQVS_F32 ? 4 : 2; }

// Block-cyclic row placement of a multi-device handle: global row r lives in block r >> kShardShift, block b on part
// b % W at local block b / W.  4096 rows = 16 tiles of 256 = 128 mask words, so packed masks, metadata columns and
// corpus tiles all split on block boundaries; parts fill in global order and differ by at most one block.
constexpr int kShardShift = 12;
constexpr uint64_t kShardBlock = 1ull << kShardShift;
static_assert(kShardBlock == MQVS_SHARD_BLOCK, "include/mqvs.h documents the block size");

struct ShardMap {
    int W = 1;
    uint64_t part_of(uint64_t r) const { return (r >> kShardShift) % (uint64_t)W; }
    uint64_t local_of(uint64_t r) const { return (((r >> kShardShift) / (uint64_t)W) << kShardShift) | (r & (kShardBlock - 1)); }
    uint64_t global_of(int part, uint64_t local) const {
        return ((((local >> kShardShift) * (uint64_t)W + (uint64_t)part)) << kShardShift) | (local & (kShardBlock - 1));
    }
    // rows of part p among the first n global rows
    uint64_t count(int part, uint64_t n) const {
        const uint64_t nb = n >> kShardShift, rem = n & (kShardBlock - 1);
        const uint64_t full = nb > (uint64_t)part ? (nb - (uint64_t)part - 1) / (uint64_t)W + 1 : 0;
        return full * kShardBlock + (nb % (uint64_t)W == (uint64_t)part ? rem : 0);
    }
};

struct Multi;  // api_multi.cu

}  // namespace mqvs

struct mqvs_store {
    std::mutex mu;
    int dev = 0;
    int dim = 0;
    int dtype = MQVS_F32;
    int sm_count = 148;
    bool coop_ok = false;  // cooperative launches available (the single-launch tensor pipeline needs grid barriers)
    int path = 0;
    void* rows = nullptr;
    uint64_t n = 0, cap = 0;
    // l2 / ip spaces over raw vectors (mqvs_set_metric): ||x|| per row, kept from the first non-cosine metric on
    int metric = MQVS_METRIC_COSINE;
    bool keep_norms = false;
    float* norms = nullptr;   // [cap] when keep_norms
    float norm_dev = 0.0f;    // largest | ||x|| - 1 | over the rows accepted so far
    float norm_max = 0.0f;    // largest ||x|| over the rows accepted so far
    mqvs::DevBuf norm_bounds; // [ceil(n / 32)][2] (max, min) norm per 32-row chunk, brought up to date before a tensor search
    uint64_t bounds_rows = 0; // rows the bounds cover
    // bf16 scoring SHADOW of an fp32 store (mqvs_set_shadow; on by default): the tensor pipelines stream this copy -- half
    // the bytes per row, kind::f16 instead of kind::tf32 -- and take every exact score (shortlist rescoring, band pass,
    // exact-scan repair) from the fp32 rows, so results are the same bits with or without it.  Built lazily: rows
    // [shadow_rows, n) are converted before the next tensor search; anything that rewrites rows in place resets shadow_rows.
    bool shadow_on = true;
    void* shadow = nullptr;              // [shadow_cap, 768] bf16
    uint64_t shadow_cap = 0;             // rows allocated
    uint64_t shadow_rows = 0;            // rows converted so far
    unsigned int* shadow_err = nullptr;  // device word: bits of the largest ||x - bf16(x)|| over the converted rows
    const void* last_score_rows = nullptr;  // what the last tensor search streamed (the band pass streams the same)
    int last_score_dtype = MQVS_F32;
    cudaStream_t st = nullptr;
    // ring of event quads so that back-to-back async searches never wait on each other's timing
    static constexpr int kEvRing = 32;
    struct EvSet {
        cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
        bool pending = false;
    } evr[kEvRing];
    int ev_cursor = 0;
    // CUDA events per search: 2 = the two that bracket the scan launches (default: stats.scan_ms, what bench.py's roofline
    // divides by), 1 = four (also first kernel to last: stats.last_search_ms), 0 = none.  A timed event costs ~3 us of stream
    // time: four of them were 11 us of a 282 us search at 1M rows (measured, round 2), two are 6.
    int timing_level = 2;
    bool timing_from_env = false;  // MQVS_TIMING was set: every path records its CUDA events, the latency path included
    bool timing = true;  // record the CUDA events behind stats.scan_ms / last_search_ms (off for the shards of a multi-device handle)
    mqvs::DevBuf qraw, qn, qnorm, cand, out_s, out_i, maskbuf, idbuf, gatherbuf;
    mqvs::DevBuf tc_qop, tc_small, tc_cnt, tc_lists, tc_keys, tc_fixcand, tc_cmax;  // tensor-core path scratch
    // host ingest pipeline (mqvs_add from host memory): two device staging buffers fed by a copy stream, two pinned
    // bounce buffers for pageable sources; K1 of chunk i overlaps the H2D copy of chunk i + 1
    mqvs::DevBuf stage[2];
    mqvs::PinBuf bounce[2];
    cudaStream_t copy_st = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_k1[2] = {nullptr, nullptr};
    int max_device_fix = 64;                             // flagged queries one exact-scan repair round takes on the device
    uint32_t* fix_dev = nullptr;                         // [kMaxFixRound] list + [1] count
    unsigned long long* fix_counters = nullptr;          // device [4]: repaired by the exact scan on device, left to the host, repaired by the band pass
    uint32_t* last_flags_dev = nullptr;                  // device [1 + B] of the last tensor-core search
    mqvs::tc::Buffers last_b{};                          // scratch layout + plan of the last tensor-core search: the
    mqvs::tc::Plan last_plan{};                          // blocking call enqueues its repair only after it has seen flags
    mqvs::tc::Space last_space{};                        // and the space it ranked in
    unsigned long long* bad_dev = nullptr;   // [0] = non-finite rows counted by K1, [1] = epoch of the last fused
    unsigned long long* bad_host = nullptr;  // pinned mirror     search that saw a non-finite query (low 32 bits),
                                             // [2] = bits of the largest | ||x|| - 1 | of the current mqvs_add (low 32 bits)
    unsigned int* ticket_dev = nullptr;      // CTA ticket of the fused latency kernel (zero between launches)
    unsigned int epoch = 0;                  // id of the last fused search
    // Host-output searches: the kernels write results and status straight into this mapped pinned block (no
    // device-to-host copies): [B*k int64 ids | B*k floats scores | status words]; the latency path uses the first
    // 64 slots of each.  status: [0] epoch reached, [1] non-finite flag (latency path)
    mqvs::PinBuf host_out;
    // mapped pinned words the tensor pipelines report through in a host-output search: [0] flag count, [1..B] flags
    mqvs::PinBuf flags_map;
    bool last_fused = false;                 // the last search reported non-finite queries through the epoch word
    mqvs_stats stats{};
    uint64_t host_fallbacks = 0;
    // device-side `where` evaluation
    struct Column {
        mqvs::DevBuf values, present, offsets;  // offsets: MQVS_COL_TOKENS only
        uint64_t n = 0;
        int type = 0;
        bool used = false, has_present = false;
    } columns[mqvs::kFilterMaxCols];
    mqvs::DevBuf f_table, f_prog, f_pool, f_live, f_mask;
    unsigned long long* f_count_dev = nullptr;
    unsigned long long* f_count_host = nullptr;  // pinned
    bool f_table_dirty = true;

    mqvs::Multi* multi = nullptr;  // n_dev > 1: this handle only fans out to multi->parts (api_multi.cu)

    size_t row_bytes() const { return (size_t)MQVS_DIM_PAD * mqvs::elem_size(dtype); }
};

namespace mqvs {
namespace api {

// ---- one shard on one device (api.cu); callers hold s->mu and have made s->dev current
int create_single(int dev, int dim, int store_dtype, uint64_t capacity_rows, mqvs_store** out);
void destroy_single(mqvs_store* s);
int grow(mqvs_store* s, uint64_t need_rows);

// Ingest in two steps so that several shards can accept or reject one mqvs_add together: add_stage runs K1 for the
// segments into rows [s->n, s->n + total) WITHOUT committing them and reports what it saw; add_commit publishes them.
struct AddSeg {
    const float* src;     // host or device fp32 [n, dim]; nullptr for synthetic rows
    uint64_t n;
    uint64_t first_index; // synthetic rows: index of the segment's first row in the counter-based generator
};
struct AddReport {
    uint64_t rows = 0, bad = 0, launches = 0;
    float norm_dev = 0.0f, norm_max = 0.0f;
};
int add_stage(mqvs_store* s, const AddSeg* segs, size_t n_segs, int x_is_device, bool synth, uint64_t seed, AddReport* rep);
void add_commit(mqvs_store* s, const AddReport& rep);

int enqueue_search(mqvs_store* s, const float* q_raw_dev, int B, int k, const uint32_t* mask_dev, int64_t id_base,
                   float* out_scores_dev, int64_t* out_ids_dev, cudaStream_t st, unsigned int* host_status = nullptr,
                   bool defer_repair = false, bool report_to_host = false, const float* q_raw_host = nullptr);
// after the stream has been synchronised: host-driven repair of whatever the device-side repair left flagged (reads the
// flags the last tensor search reported; *changed says whether outputs were rewritten and the stream synchronised again)
int finish_search(mqvs_store* s, int B, int k, const uint32_t* mask_dev, int64_t id_base, float* out_scores_dev,
                  int64_t* out_ids_dev, cudaStream_t st, bool* changed);
int column_set_single(mqvs_store* s, int col_id, int col_type, const void* values_host, const uint8_t* present_host, uint64_t n);
int column_set_tokens_single(mqvs_store* s, int col_id, const uint64_t* offsets_host, const int64_t* tokens_host, uint64_t n);
int column_set_text_single(mqvs_store* s, int col_id, const uint64_t* offsets_host, const uint8_t* bytes_host,
                           const uint8_t* present_host, uint64_t n);
int filter_build_single(mqvs_store* s, const mqvs_filter_op* prog, int n_ops, const void* pool, uint64_t pool_len,
                        const uint32_t* live_words_host, const uint32_t** mask_dev_out, uint64_t* allowed_out);
int get_rows_single(mqvs_store* s, const int64_t* row_ids, uint64_t n, float* out_host);
bool takes_latency_kernel(const mqvs_store* s, int B, int k);
void settle_events(mqvs_store* s);
void fill_stats(mqvs_store* s);  // settles events + repair counters into s->stats

// on-disk row file (one format for every device count): 40-byte header + rows in GLOBAL order
struct FileHeader {
    char magic[8];
    uint32_t version, dim, dtype, dim_pad;
    uint64_t rows;
};
struct NormsHeader {
    char magic[8];
    uint64_t rows;
};

// ---- n_dev > 1 (api_multi.cu): every entry point of include/mqvs.h that takes a handle has a multi_* twin
int multi_create(int n_dev, const int* dev_ids, int dim, int store_dtype, uint64_t capacity_rows, mqvs_store** out);
void multi_destroy(mqvs_store* s);
int multi_count(mqvs_store* s, uint64_t* out);
int multi_reserve(mqvs_store* s, uint64_t capacity_rows);
int multi_add(mqvs_store* s, const void* x, int x_is_device, bool synth, uint64_t seed, uint64_t first_index, uint64_t n,
              uint64_t* first_row_out);
int multi_search(mqvs_store* s, const float* q, int q_is_device, int B, int k, const uint32_t* mask, int mask_is_device,
                 int64_t id_base, float* out_scores, int64_t* out_ids, int out_is_device, cudaStream_t caller_stream,
                 bool blocking);
int multi_get_rows(mqvs_store* s, const int64_t* row_ids, uint64_t n, float* out_host);
int multi_compact(mqvs_store* s, const uint32_t* keep_words_host, uint64_t* new_count_out);
int multi_save(mqvs_store* s, const char* path, uint64_t first_row);
int multi_load(mqvs_store* s, const char* path);
int multi_column_set(mqvs_store* s, int col_id, int col_type, const void* values_host, const uint8_t* present_host, uint64_t n);
int multi_column_set_tokens(mqvs_store* s, int col_id, const uint64_t* offsets_host, const int64_t* tokens_host, uint64_t n);
int multi_column_set_text(mqvs_store* s, int col_id, const uint64_t* offsets_host, const uint8_t* bytes_host,
                          const uint8_t* present_host, uint64_t n);
int multi_filter_build(mqvs_store* s, const mqvs_filter_op* prog, int n_ops, const void* pool, uint64_t pool_len,
                       const uint32_t* live_words_host, const uint32_t** mask_dev_out, uint64_t* allowed_out);
int multi_get_stats(mqvs_store* s, mqvs_stats* out);
int multi_set_metric(mqvs_store* s, int metric);
int multi_norm_deviation(mqvs_store* s, float* out);
int multi_get_norms(mqvs_store* s, uint64_t first_row, uint64_t n, float* out_host);
int multi_set_path(mqvs_store* s, int path);
int multi_set_shadow(mqvs_store* s, int enable);
int multi_shadow_bytes(mqvs_store* s, uint64_t* bytes_out);

}  // namespace api
}  // namespace mqvs
