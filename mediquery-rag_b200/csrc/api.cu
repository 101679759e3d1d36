// api.cu -- the C ABI of libmqvs.so (include/mqvs.h): store lifetime, ingest, search orchestration, persistence.
// No torch types, no exceptions across the boundary; CUDA errors become MQVS_E_CUDA with a thread-local message.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
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

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(expr)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            int code_ = (e_ == cudaErrorMemoryAllocation) ? MQVS_E_OOM : MQVS_E_CUDA;                    \
            cudaGetLastError();                                                                          \
            return fail(code_, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                                \
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
        CU(cudaMalloc(&p, want));
        bytes = want;
        return MQVS_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
};

size_t elem_size(int dtype) { return dtype == MQVS_F32 ? 4 : 2; }

}  // namespace

namespace mqvs {
int mqvs_fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
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
    DevBuf norm_bounds;       // [ceil(n / 32)][2] (max, min) norm per 32-row chunk, brought up to date before a tensor search
    uint64_t bounds_rows = 0; // rows the bounds cover
    cudaStream_t st = nullptr;
    // ring of event quads so that back-to-back async searches never wait on each other's timing
    static constexpr int kEvRing = 32;
    struct EvSet {
        cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
        bool pending = false;
    } evr[kEvRing];
    int ev_cursor = 0;
    DevBuf stage, qraw, qn, qnorm, cand, out_s, out_i, maskbuf, idbuf, gatherbuf;
    DevBuf tc_qop, tc_small, tc_cnt, tc_lists, tc_keys, tc_fixcand, tc_cmax;  // tensor-core path scratch
    static constexpr int kMaxDeviceFix = 64;             // flagged queries repaired on the device per search
    uint32_t* fix_dev = nullptr;                         // [kMaxDeviceFix] list + [1] count
    unsigned long long* fix_counters = nullptr;          // device [4]: repaired by the exact scan on device, left to the host, repaired by the band pass
    uint32_t* flags_host = nullptr;                      // pinned [1 + max B]: flag_count, flags[B]
    size_t flags_host_n = 0;
    uint32_t* last_flags_dev = nullptr;                  // device [1 + B] of the last tensor-core search
    mqvs::tc::Buffers last_b{};                          // scratch layout + plan of the last tensor-core search: the
    mqvs::tc::Plan last_plan{};                          // blocking call enqueues its repair only after it has seen flags
    mqvs::tc::Space last_space{};                        // and the space it ranked in
    unsigned long long* bad_dev = nullptr;   // [0] = non-finite rows counted by K1, [1] = epoch of the last fused
    unsigned long long* bad_host = nullptr;  // pinned mirror     search that saw a non-finite query (low 32 bits),
                                             // [2] = bits of the largest | ||x|| - 1 | of the current mqvs_add (low 32 bits)
    unsigned int* ticket_dev = nullptr;      // CTA ticket of the fused latency kernel (zero between launches)
    unsigned int epoch = 0;                  // id of the last fused search
    // latency path with host outputs: the kernel writes results + status straight into this mapped pinned block
    // (no device-to-host copies): [64 floats scores | 64 int64 ids | 2 uint32 status]
    char* host_out = nullptr;
    // pinned staging for host outputs of every other path: D2H copies into pageable user memory are staged and
    // synchronous inside the driver (~10 us each), copies into pinned memory are not
    char* out_pinned = nullptr;
    size_t out_pinned_bytes = 0;
    bool last_fused = false;                 // the last search reported non-finite queries through the epoch word
    mqvs_stats stats{};
    uint64_t host_fallbacks = 0;
    // device-side `where` evaluation
    struct Column {
        DevBuf values, present;
        uint64_t n = 0;
        int type = 0;
        bool used = false, has_present = false;
    } columns[mqvs::kFilterMaxCols];
    DevBuf f_table, f_prog, f_pool, f_live, f_mask;
    unsigned long long* f_count_dev = nullptr;
    unsigned long long* f_count_host = nullptr;  // pinned
    bool f_table_dirty = true;

    size_t row_bytes() const { return (size_t)MQVS_DIM_PAD * elem_size(dtype); }
};

namespace {

int grow(mqvs_store* s, uint64_t need_rows) {
    if (need_rows <= s->cap) return MQVS_OK;
    if (need_rows > 0xfffffff0ull) return fail(MQVS_E_INVALID, "a shard holds at most 2^32-16 rows");
    uint64_t want = std::max<uint64_t>(need_rows, std::max<uint64_t>(1024, s->cap + s->cap / 2));
    void* fresh = nullptr;
    cudaError_t e = cudaMalloc(&fresh, want * s->row_bytes());
    if (e != cudaSuccess && want > need_rows) {  // retry with the exact size before giving up
        cudaGetLastError();
        want = need_rows;
        e = cudaMalloc(&fresh, want * s->row_bytes());
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(MQVS_E_OOM, "cudaMalloc of %llu rows (%.2f GB) failed: %s", (unsigned long long)want,
                    want * s->row_bytes() / 1e9, cudaGetErrorString(e));
    }
    float* fresh_norms = nullptr;
    if (s->keep_norms && cudaMalloc(&fresh_norms, want * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(fresh);
        return fail(MQVS_E_OOM, "cudaMalloc of %llu row norms failed", (unsigned long long)want);
    }
    cudaError_t ce = cudaSuccess;
    if (s->n) ce = cudaMemcpyAsync(fresh, s->rows, s->n * s->row_bytes(), cudaMemcpyDeviceToDevice, s->st);
    if (ce == cudaSuccess && s->n && fresh_norms)
        ce = cudaMemcpyAsync(fresh_norms, s->norms, s->n * sizeof(float), cudaMemcpyDeviceToDevice, s->st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(s->st);
    if (ce != cudaSuccess) {
        cudaGetLastError();
        cudaFree(fresh);
        if (fresh_norms) cudaFree(fresh_norms);
        return fail(MQVS_E_CUDA, "moving %llu rows to the larger allocation failed: %s", (unsigned long long)s->n, cudaGetErrorString(ce));
    }
    if (s->rows) cudaFree(s->rows);
    s->rows = fresh;
    if (fresh_norms) {
        if (s->norms) cudaFree(s->norms);
        s->norms = fresh_norms;
    }
    s->cap = want;
    return MQVS_OK;
}

// fold one finished search's CUDA-event timings into the stats (waits for that search if still running)
void settle_slot(mqvs_store* s, int slot) {
    auto& ev = s->evr[slot];
    if (!ev.pending) return;
    if (cudaEventSynchronize(ev.e[3]) == cudaSuccess) {
        float total = 0.f, scan = 0.f;
        cudaEventElapsedTime(&total, ev.e[0], ev.e[3]);
        cudaEventElapsedTime(&scan, ev.e[1], ev.e[2]);
        s->stats.last_search_ms = total;
        s->stats.last_scan_ms = scan;
        s->stats.scan_ms += scan;
    } else {
        cudaGetLastError();
    }
    ev.pending = false;
}
void settle_events(mqvs_store* s) {
    for (int i = 0; i < mqvs_store::kEvRing; ++i) settle_slot(s, (s->ev_cursor + i) % mqvs_store::kEvRing);
}

// Tensor-core path is worth it once the batch amortises its fixed launches and the corpus fills the machine.
bool use_tensor_path(const mqvs_store* s, int B, int k) {
    if (s->path == 1 || s->path == 3) return false;
    if (k > 128 || s->n < 4096) return false;
    if (s->path == 2 || s->path == 4) return true;
    // measured on B200 (profiles/r01_*): the register-resident SIMT scan holds the HBM roofline for 1 query (bf16 rows)
    // or 1-2 queries (fp32 rows); beyond that its multi-query variants fall to 2-3.4 TB/s while the tensor path streams
    // the corpus once for up to 128 queries at 6.3-7.1 TB/s.
    // From 1M rows on even single queries take the (single-launch) tensor pipeline: same bandwidth as the SIMT scan when
    // the SM clock is at boost (0.471 vs 0.487 ms at 1M x 768 fp32, 2.67 vs 2.68 ms at 12.5M bf16), and it keeps it when
    // the clock drops under a power cap -- on 8 power-capped GPUs the SIMT scan alternated between 2.72 and 3.17 ms per
    // 12.5M-row shard while the TMA-fed pipeline stayed at 2.71-2.75 (profiles/r01_b1_paths_w8.json).  Below 1M rows the
    // SIMT path's single launch wins on latency.
    const int simt_max = s->dtype == MQVS_F32 ? 2 : 1;
    // (l2 / ip spaces: the transform makes the tensor epilogue ~4 % slower than the SIMT scan for the 1-2 queries that scan
    // holds the roofline for -- 0.508 vs 0.489 ms at 1M x 768 fp32, B=1 -- so those stay on the SIMT scan at every size)
    if (s->n >= 1000000ull && (s->metric == MQVS_METRIC_COSINE || B > simt_max)) return true;
    return B > simt_max && s->n >= 65536;
}

// Exact SIMT scan of queries [b0, b0 + nb) of s->qn into out rows [b0, ...).  Returns passes launched.
int enqueue_simt(mqvs_store* s, int b0, int nb, int k, const uint32_t* mask_dev, int64_t id_base, float* out_scores_dev,
                 int64_t* out_ids_dev, cudaStream_t st, uint32_t* passes_out) {
    using namespace mqvs;
    const int gmax = scan_max_grid(s->sm_count);
    int rc;
    if ((rc = s->cand.ensure((size_t)nb * gmax * k * sizeof(uint64_t))) != MQVS_OK) return rc;
    CU(cudaMemsetAsync(s->cand.p, 0, (size_t)nb * gmax * k * sizeof(uint64_t), st));
    uint32_t passes = 0;
    for (int b = 0; b < nb;) {
        int nq = 0, grid = 0;
        scan_plan(s->dtype, nb - b, k, s->sm_count, &nq, &grid);
        if (s->metric != MQVS_METRIC_COSINE)
            launch_scan_simt_metric(s->rows, s->dtype, (uint32_t)s->n, (const float*)s->qn.p + (size_t)(b0 + b) * MQVS_DIM_PAD, nq,
                                    mask_dev, k, (uint64_t*)s->cand.p + (size_t)b * gmax * k, gmax, grid, s->metric, s->norms,
                                    (const float*)s->qnorm.p + (b0 + b), st);
        else
            launch_scan_simt(s->rows, s->dtype, (uint32_t)s->n, (const float*)s->qn.p + (size_t)(b0 + b) * MQVS_DIM_PAD, nq,
                             mask_dev, k, (uint64_t*)s->cand.p + (size_t)b * gmax * k, gmax, grid, st);
        b += nq;
        ++passes;
    }
    launch_merge_keys((const uint64_t*)s->cand.p, nb, gmax * k, k, id_base, out_scores_dev + (size_t)b0 * k,
                      out_ids_dev + (size_t)b0 * k, st);
    CU(cudaGetLastError());
    *passes_out = passes;
    return MQVS_OK;
}

// Repair of tensor-path queries whose shortlist was not proven exact: first the band pass (all flagged queries together
// through the tensor pipeline once more, with given thresholds), then whatever it could not settle is re-scanned exactly.
// Entirely device-driven (no host sync), so mqvs_search_async enqueues it unconditionally after every tensor search;
// the blocking mqvs_search looks at the flag count first and skips it when nothing is flagged (the common case).
void enqueue_tensor_repair(mqvs_store* s, int B, int k, const uint32_t* mask_dev, int64_t id_base, float* out_scores_dev,
                           int64_t* out_ids_dev, cudaStream_t st, int band_rounds) {
    using namespace mqvs;
    const tc::Buffers& b = s->last_b;
    static const bool no_band = getenv("MQVS_NO_BAND") != nullptr;  // developer knob: measure what the band pass costs
    if (!no_band && s->coop_ok && s->path != 4 && tc::fused_ok(1, k, s->sm_count)) {
        for (int r = 0; r < band_rounds; ++r) {  // 128 flagged queries per round
            const char* err = tc::search_band(b, s->last_plan, s->rows, s->dtype, (uint32_t)s->n, (const float*)s->qn.p, B, k,
                                              mask_dev, id_base, out_scores_dev, out_ids_dev, s->fix_counters, st, s->last_space);
            if (err) {
                cudaGetLastError();  // a refused cooperative launch: the exact scan below still repairs everything
                break;
            }
            s->stats.kernel_launches += 2;
        }
    }
    uint32_t* fix_list = s->fix_dev;
    uint32_t* fix_n = s->fix_dev + mqvs_store::kMaxDeviceFix;
    launch_scan_fix(s->rows, s->dtype, (uint32_t)s->n, (const float*)s->qn.p, b.flags, B, mqvs_store::kMaxDeviceFix, fix_list, fix_n,
                    s->fix_counters, mask_dev, k, (uint64_t*)s->tc_fixcand.p, s->sm_count, st, s->last_space.metric,
                    s->last_space.norms, s->last_space.qnorm);
    launch_merge_fix((const uint64_t*)s->tc_fixcand.p, mqvs_store::kMaxDeviceFix, scan_fix_grid(s->sm_count) * k, k, id_base,
                     fix_list, fix_n, b.flags, out_scores_dev, out_ids_dev, st);
    s->stats.kernel_launches += 2;
}

// Enqueue one search on `st`; every pointer is device memory.  q_raw [B, dim] is normalised into s->qn first.
int enqueue_search(mqvs_store* s, const float* q_raw_dev, int B, int k, const uint32_t* mask_dev, int64_t id_base,
                   float* out_scores_dev, int64_t* out_ids_dev, cudaStream_t st, unsigned int* host_status = nullptr,
                   bool defer_repair = false) {
    using namespace mqvs;
    const int slot = s->ev_cursor;
    s->ev_cursor = (s->ev_cursor + 1) % mqvs_store::kEvRing;
    settle_slot(s, slot);  // oldest entry of the ring: long finished unless 32 searches are in flight
    cudaEvent_t* ev = s->evr[slot].e;
    const bool tensor = use_tensor_path(s, B, k);
    const bool fused = !tensor && s->path != 3 && s->metric == MQVS_METRIC_COSINE && scan_fused_ok(B, k);
    s->last_fused = fused;
    tc::Plan plan{};
    int rc;
    if (fused) {
        // latency path: normalisation, scan, top-k and the final merge in one launch
        if ((rc = s->cand.ensure((size_t)B * scan_fused_max_grid(s->sm_count) * k * sizeof(uint64_t))) != MQVS_OK) return rc;
        CU(cudaEventRecord(ev[0], st));
        CU(cudaEventRecord(ev[1], st));
        launch_scan_fused(s->rows, s->dtype, (uint32_t)s->n, q_raw_dev, s->dim, B, mask_dev, k, (uint64_t*)s->cand.p,
                          out_scores_dev, out_ids_dev, id_base, s->ticket_dev, (unsigned int*)(s->bad_dev + 1), ++s->epoch,
                          s->sm_count, st, host_status);
        CU(cudaEventRecord(ev[2], st));
        CU(cudaEventRecord(ev[3], st));
        CU(cudaGetLastError());
        s->evr[slot].pending = true;
        s->stats.kernel_launches += 1;
        s->stats.scan_launches += 1;
        s->stats.scan_bytes += s->n * s->row_bytes();
        s->stats.last_scan_passes = 1;
        s->stats.last_path = 0;
        s->stats.last_pipeline = 1;
        s->stats.searches += 1;
        s->stats.queries += (uint64_t)B;
        return MQVS_OK;
    }
    // single-launch tensor pipeline for one or two query tiles (B <= 256); path 4 = keep the segmented pipeline (tests compare)
    bool tc_one = tensor && s->path != 4 && s->coop_ok && tc::fused_ok(B, k, s->sm_count);
    if (tensor) {
        plan = tc::plan(k, B, s->sm_count);
        const size_t small = ((size_t)plan.b_pad * 2 + (size_t)B * 4 + 128 + 256 + 32 + 144) * sizeof(uint32_t);
        if ((rc = s->tc_qop.ensure(tc::op_bytes(plan, s->dtype))) != MQVS_OK) return rc;
        if ((rc = s->tc_small.ensure(small)) != MQVS_OK) return rc;
        if ((rc = s->tc_cnt.ensure(tc::cnt_bytes(plan))) != MQVS_OK) return rc;
        if ((rc = s->tc_lists.ensure(tc::lists_bytes(plan))) != MQVS_OK) return rc;
        if ((rc = s->tc_keys.ensure((size_t)plan.b_pad * plan.kprime * sizeof(uint64_t) * 2)) != MQVS_OK) return rc;
        if ((rc = s->tc_fixcand.ensure((size_t)mqvs_store::kMaxDeviceFix * scan_fix_grid(s->sm_count) * k * sizeof(uint64_t))) != MQVS_OK) return rc;
        if ((rc = s->tc_cmax.ensure(tc::cmax_bytes(plan))) != MQVS_OK) return rc;
    }
    if ((rc = s->qn.ensure((size_t)B * MQVS_DIM_PAD * sizeof(float))) != MQVS_OK) return rc;
    if (s->metric != MQVS_METRIC_COSINE && (rc = s->qnorm.ensure((size_t)B * sizeof(float))) != MQVS_OK) return rc;

    tc::Buffers b{};
    if (tensor) {
        char* sm = (char*)s->tc_small.p;
        b.q_op = s->tc_qop.p;
        b.tau_tilde = (float*)sm;
        b.tau = (uint32_t*)(b.tau_tilde + plan.b_pad);
        b.eps = (float*)(b.tau + plan.b_pad);    // [B]
        b.flag_count = (uint32_t*)(b.eps + B);   // flag_count immediately followed by flags[B]: one D2H copy
        b.flags = b.flag_count + 1;
        b.gbar = (unsigned int*)(b.flags + B);  // [4]
        b.band = (uint32_t*)(b.gbar + 8);        // [B]
        b.band_list = b.band + B;               // [128]
        b.band_n = b.band_list + 128;           // [1]
        b.h_op = (float*)(b.band_n + 8);        // [256]
        b.hq = b.h_op + 256;                    // [B rounded up to 128]: every chunk of the segmented pipeline keeps its slice
        b.cnt = (uint32_t*)s->tc_cnt.p;
        b.lists = (uint64_t*)s->tc_lists.p;
        b.short_keys = (uint64_t*)s->tc_keys.p;
        b.exact_keys = b.short_keys + (size_t)plan.b_pad * plan.kprime;
        b.cmax = (uint32_t*)s->tc_cmax.p;
        s->last_flags_dev = b.flag_count;
    }
    tc::Space space{};
    if (tensor) {
        if (s->metric != MQVS_METRIC_COSINE) {
            space.metric = s->metric;
            space.norms = s->norms;
            space.amax = s->norm_max;
            space.qnorm = (float*)s->qnorm.p;
            if (s->bounds_rows != s->n) {  // rows were added (or loaded): redo the chunks they touched
                const size_t had = s->norm_bounds.bytes;
                const size_t need = (size_t)((s->n + 31) / 32) * 2 * sizeof(float);
                uint64_t first_chunk = s->bounds_rows / 32;
                if (need > had) {  // grows: keep a 1.5x margin, recompute everything (4 bytes per row read, once per growth)
                    if ((rc = s->norm_bounds.ensure(need + need / 2)) != MQVS_OK) return rc;
                    first_chunk = 0;
                }
                launch_norm_bounds(s->norms, s->n, first_chunk, (float*)s->norm_bounds.p, st);
                s->stats.kernel_launches += 1;
                s->bounds_rows = s->n;
            }
            space.norm_bounds = (const float*)s->norm_bounds.p;
        }
        s->last_b = b;
        s->last_plan = plan;
        s->last_space = space;
    }
    auto enqueue_repair = [&]() {
        // nobody can look at the flags here: one band round per 128 queries of the batch, at most two
        if (!defer_repair) enqueue_tensor_repair(s, B, k, mask_dev, id_base, out_scores_dev, out_ids_dev, st, B > 128 ? 2 : 1);
    };
    if (tc_one) {
        s->last_fused = true;  // non-finite queries are reported through the epoch word, as on the latency path
        CU(cudaEventRecord(ev[0], st));
        CU(cudaEventRecord(ev[1], st));
        bool coop_refused = false;
        const char* err = tc::search_fused(b, plan, s->rows, s->dtype, (uint32_t)s->n, q_raw_dev, s->dim, (float*)s->qn.p, B, k,
                                           mask_dev, id_base, out_scores_dev, out_ids_dev, (unsigned int*)(s->bad_dev + 1),
                                           ++s->epoch, st, &coop_refused, space);
        if (err && coop_refused) {
            // the device cannot hold one CTA per SM right now (MPS with an SM limit, a partitioned GPU): this handle
            // keeps to the segmented pipeline from here on
            s->coop_ok = false;
            tc_one = false;
            s->last_fused = false;
        } else if (err) {
            return fail(MQVS_E_CUDA, "tensor-core search: %s", err);
        }
    }
    if (tc_one) {
        CU(cudaEventRecord(ev[2], st));  // brackets prep + the single scan/select/rescore launch
        enqueue_repair();
        CU(cudaEventRecord(ev[3], st));
        CU(cudaGetLastError());
        s->evr[slot].pending = true;
        s->stats.kernel_launches += 2;
        s->stats.scan_launches += 1;
        s->stats.scan_bytes += s->n * s->row_bytes();
        s->stats.last_scan_passes = 1;
        s->stats.last_path = 1;
        s->stats.last_pipeline = 3;
        s->stats.searches += 1;
        s->stats.queries += (uint64_t)B;
        return MQVS_OK;
    }

    CU(cudaEventRecord(ev[0], st));
    CU(cudaMemsetAsync(s->bad_dev, 0, sizeof(unsigned long long), st));
    launch_normalize(q_raw_dev, (uint64_t)B, s->dim, MQVS_F32, s->qn.p, s->bad_dev, st,
                     s->metric != MQVS_METRIC_COSINE ? (float*)s->qnorm.p : nullptr);
    s->stats.kernel_launches += 1;
    uint32_t passes = 0;
    if (tensor) {
        CU(cudaEventRecord(ev[1], st));
        const char* err = tc::search(b, plan, s->rows, s->dtype, (uint32_t)s->n, (const float*)s->qn.p, B, k, mask_dev,
                                     id_base, out_scores_dev, out_ids_dev, st, space);
        if (err) return fail(MQVS_E_CUDA, "tensor-core search: %s", err);
        CU(cudaEventRecord(ev[2], st));  // brackets prep + score + select + rescore + finalize
        enqueue_repair();
        passes = 1;
        s->stats.kernel_launches += tc::launches_per_search(plan, (uint32_t)s->n, B);
        s->stats.last_path = 1;
        s->stats.last_pipeline = 2;
    } else {
        CU(cudaEventRecord(ev[1], st));
        // scan launches are bracketed together with the (tiny) candidate memset
        if ((rc = enqueue_simt(s, 0, B, k, mask_dev, id_base, out_scores_dev, out_ids_dev, st, &passes)) != MQVS_OK) return rc;
        CU(cudaEventRecord(ev[2], st));
        s->stats.kernel_launches += passes + 1;
        s->stats.last_path = 0;
        s->stats.last_pipeline = 0;
    }
    CU(cudaEventRecord(ev[3], st));
    CU(cudaGetLastError());

    s->evr[slot].pending = true;
    s->stats.scan_launches += passes;
    s->stats.scan_bytes += (uint64_t)passes * s->n * s->row_bytes();
    s->stats.last_scan_passes = passes;
    s->stats.searches += 1;
    s->stats.queries += (uint64_t)B;
    return MQVS_OK;
}

}  // namespace

extern "C" {

MQVS_API int mqvs_version(void) { return MQVS_VERSION; }

MQVS_API const char* mqvs_last_error(void) { return g_err.c_str(); }

MQVS_API int mqvs_device_count(int* out) {
    if (!out) return fail(MQVS_E_INVALID, "out is NULL");
    int n = 0;
    CU(cudaGetDeviceCount(&n));
    *out = n;
    return MQVS_OK;
}

MQVS_API int mqvs_create(int n_dev, const int* dev_ids, int dim, int store_dtype, uint64_t capacity_rows, mqvs_t** out) {
    if (!out) return fail(MQVS_E_INVALID, "out is NULL");
    *out = nullptr;
    if (n_dev != 1)
        return fail(MQVS_E_INVALID, "one device per handle: shard rows across ranks (one process per GPU) and merge "
                                    "with mqvs_merge_candidates; got n_dev=%d", n_dev);
    if (dim < 1 || dim > MQVS_DIM_PAD) return fail(MQVS_E_INVALID, "dim must be in [1, %d], got %d", MQVS_DIM_PAD, dim);
    if (store_dtype != MQVS_F32 && store_dtype != MQVS_BF16) return fail(MQVS_E_INVALID, "unknown store_dtype %d", store_dtype);
    int dev = dev_ids ? dev_ids[0] : 0;
    int count = 0;
    CU(cudaGetDeviceCount(&count));
    if (dev < 0 || dev >= count) return fail(MQVS_E_INVALID, "device %d not present (%d visible)", dev, count);
    CU(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10)
        return fail(MQVS_E_CUDA, "libmqvs is built for sm_100a only; device %d is sm_%d%d", dev, prop.major, prop.minor);
    mqvs_store* s = new (std::nothrow) mqvs_store();
    if (!s) return fail(MQVS_E_OOM, "host allocation failed");
    s->dev = dev;
    s->dim = dim;
    s->dtype = store_dtype;
    s->sm_count = prop.multiProcessorCount;
    s->coop_ok = prop.cooperativeLaunch != 0;
    auto bail = [&](int rc) {
        mqvs_destroy(s);
        return rc;
    };
    if (cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(MQVS_E_CUDA, "stream creation failed"));
    for (auto& set : s->evr)
        for (auto& e : set.e)
            if (cudaEventCreate(&e) != cudaSuccess) return bail(fail(MQVS_E_CUDA, "event creation failed"));
    if (cudaMalloc(&s->bad_dev, 3 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(s->bad_dev, 0, 3 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc(&s->ticket_dev, sizeof(unsigned int)) != cudaSuccess ||
        cudaMemset(s->ticket_dev, 0, sizeof(unsigned int)) != cudaSuccess)
        return bail(fail(MQVS_E_OOM, "cudaMalloc failed"));
    if (cudaMallocHost(&s->bad_host, 3 * sizeof(unsigned long long)) != cudaSuccess) return bail(fail(MQVS_E_OOM, "cudaMallocHost failed"));
    if (cudaHostAlloc((void**)&s->host_out, 64 * 4 + 64 * 8 + 16, cudaHostAllocMapped) != cudaSuccess) {
        cudaGetLastError();
        s->host_out = nullptr;  // not fatal: the latency path copies its results back instead
    } else {
        memset(s->host_out, 0, 64 * 4 + 64 * 8 + 16);
    }
    if (cudaMalloc(&s->fix_dev, (mqvs_store::kMaxDeviceFix + 1) * sizeof(uint32_t)) != cudaSuccess ||
        cudaMalloc(&s->fix_counters, 4 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(s->fix_counters, 0, 4 * sizeof(unsigned long long)) != cudaSuccess)
        return bail(fail(MQVS_E_OOM, "cudaMalloc failed"));
    if (capacity_rows) {
        int rc = grow(s, capacity_rows);
        if (rc != MQVS_OK) return bail(rc);
    }
    *out = s;
    return MQVS_OK;
}

MQVS_API void mqvs_destroy(mqvs_t* s) {
    if (!s) return;
    cudaSetDevice(s->dev);
    if (s->st) cudaStreamSynchronize(s->st);
    for (auto& c : s->columns) {
        c.values.release();
        c.present.release();
    }
    for (DevBuf* b : {&s->f_table, &s->f_prog, &s->f_pool, &s->f_live, &s->f_mask}) b->release();
    if (s->f_count_dev) cudaFree(s->f_count_dev);
    if (s->f_count_host) cudaFreeHost(s->f_count_host);
    for (DevBuf* b : {&s->norm_bounds, &s->stage, &s->qraw, &s->qn, &s->qnorm, &s->cand, &s->out_s, &s->out_i, &s->maskbuf, &s->idbuf, &s->gatherbuf,
                      &s->tc_qop, &s->tc_small, &s->tc_cnt, &s->tc_lists, &s->tc_keys, &s->tc_fixcand, &s->tc_cmax})
        b->release();
    if (s->fix_dev) cudaFree(s->fix_dev);
    if (s->fix_counters) cudaFree(s->fix_counters);
    if (s->flags_host) cudaFreeHost(s->flags_host);
    if (s->rows) cudaFree(s->rows);
    if (s->norms) cudaFree(s->norms);
    if (s->bad_dev) cudaFree(s->bad_dev);
    if (s->ticket_dev) cudaFree(s->ticket_dev);
    if (s->bad_host) cudaFreeHost(s->bad_host);
    if (s->host_out) cudaFreeHost(s->host_out);
    if (s->out_pinned) cudaFreeHost(s->out_pinned);
    for (auto& set : s->evr)
        for (auto& e : set.e)
            if (e) cudaEventDestroy(e);
    if (s->st) cudaStreamDestroy(s->st);
    cudaGetLastError();
    delete s;
}

MQVS_API int mqvs_count(mqvs_t* s, uint64_t* out) {
    if (!s || !out) return fail(MQVS_E_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(s->mu);
    *out = s->n;
    return MQVS_OK;
}

MQVS_API int mqvs_reserve(mqvs_t* s, uint64_t capacity_rows) {
    if (!s) return fail(MQVS_E_INVALID, "NULL handle");
    std::lock_guard<std::mutex> lk(s->mu);
    CU(cudaSetDevice(s->dev));
    return grow(s, capacity_rows);
}

static int add_common(mqvs_t* s, const void* x, int x_is_device, bool synth, uint64_t seed, uint64_t first_index,
                      uint64_t n, uint64_t* first_row_out) {
    using namespace mqvs;
    std::lock_guard<std::mutex> lk(s->mu);
    CU(cudaSetDevice(s->dev));
    if (first_row_out) *first_row_out = s->n;
    if (n == 0) return MQVS_OK;
    int rc = grow(s, s->n + n);
    if (rc != MQVS_OK) return rc;
    char* dst = (char*)s->rows + s->n * s->row_bytes();
    float* norms_dst = s->keep_norms ? s->norms + s->n : nullptr;
    unsigned int* dev_word = (unsigned int*)(s->bad_dev + 2);
    CU(cudaMemsetAsync(s->bad_dev, 0, sizeof(unsigned long long), s->st));
    CU(cudaMemsetAsync(s->bad_dev + 2, 0, sizeof(unsigned long long), s->st));
    uint64_t launches = 0;
    if (synth) {
        launch_synth_normalize(seed, first_index, n, s->dim, s->dtype, dst, s->bad_dev, s->st, norms_dst, dev_word);
        launches = 1;
    } else if (x_is_device) {
        launch_normalize((const float*)x, n, s->dim, s->dtype, dst, s->bad_dev, s->st, norms_dst, dev_word);
        launches = 1;
    } else {
        const uint64_t chunk = 65536;
        if ((rc = s->stage.ensure((size_t)std::min(n, chunk) * s->dim * sizeof(float))) != MQVS_OK) return rc;
        for (uint64_t r = 0; r < n; r += chunk) {
            uint64_t m = std::min(chunk, n - r);
            CU(cudaMemcpyAsync(s->stage.p, (const float*)x + r * s->dim, (size_t)m * s->dim * sizeof(float),
                               cudaMemcpyHostToDevice, s->st));
            launch_normalize((const float*)s->stage.p, m, s->dim, s->dtype, dst + r * s->row_bytes(), s->bad_dev, s->st,
                             norms_dst ? norms_dst + r : nullptr, dev_word);
            CU(cudaStreamSynchronize(s->st));  // staging buffer is reused
            ++launches;
        }
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(s->bad_host, s->bad_dev, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    s->stats.kernel_launches += launches;
    if (*s->bad_host)
        return fail(MQVS_E_INVALID, "%llu of %llu embeddings have a non-finite norm; nothing was added",
                    *s->bad_host, (unsigned long long)n);
    {
        const uint32_t bits = (uint32_t)s->bad_host[2], bits_max = (uint32_t)(s->bad_host[2] >> 32);
        float dev, mx;
        memcpy(&dev, &bits, 4);
        memcpy(&mx, &bits_max, 4);
        s->norm_dev = std::max(s->norm_dev, dev);
        s->norm_max = std::max(s->norm_max, mx);
    }
    s->n += n;
    s->stats.rows = s->n;
    return MQVS_OK;
}

MQVS_API int mqvs_add(mqvs_t* s, const void* x, int x_dtype, int x_is_device, uint64_t n, uint64_t* first_row_out) {
    if (!s) return fail(MQVS_E_INVALID, "NULL handle");
    if (n && !x) return fail(MQVS_E_INVALID, "x is NULL");
    if (x_dtype != MQVS_F32) return fail(MQVS_E_INVALID, "embeddings must be fp32 (x_dtype=%d)", x_dtype);
    return add_common(s, x, x_is_device, false, 0, 0, n, first_row_out);
}

MQVS_API int mqvs_add_synthetic(mqvs_t* s, uint64_t seed, uint64_t first_index, uint64_t n, uint64_t* first_row_out) {
    if (!s) return fail(MQVS_E_INVALID, "NULL handle");
    return add_common(s, nullptr, 1, true, seed, first_index, n, first_row_out);
}

static int check_search_args(mqvs_t* s, const float* q, int B, int k, const void* out_scores, const void* out_ids) {
    if (!s) return fail(MQVS_E_INVALID, "NULL handle");
    if (B < 0) return fail(MQVS_E_INVALID, "B must be >= 0, got %d", B);
    if (k < 1 || k > MQVS_MAX_K) return fail(MQVS_E_INVALID, "k must be in [1, %d], got %d", MQVS_MAX_K, k);
    if (B && (!q || !out_scores || !out_ids)) return fail(MQVS_E_INVALID, "NULL query or output buffer");
    return MQVS_OK;
}

static int search_locked(mqvs_t* s, const float* q, int q_is_device, int B, int k, const uint32_t* mask, int mask_is_device,
                         float* out_scores, int64_t* out_ids, int out_is_device);

MQVS_API int mqvs_search(mqvs_t* s, const float* q, int q_is_device, int B, int k, const uint32_t* mask, int mask_is_device,
                float* out_scores, int64_t* out_ids, int out_is_device) {
    int rc = check_search_args(s, q, B, k, out_scores, out_ids);
    if (rc != MQVS_OK) return rc;
    if (B == 0) return MQVS_OK;
    std::lock_guard<std::mutex> lk(s->mu);
    CU(cudaSetDevice(s->dev));
    return search_locked(s, q, q_is_device, B, k, mask, mask_is_device, out_scores, out_ids, out_is_device);
}

static int search_locked(mqvs_t* s, const float* q, int q_is_device, int B, int k, const uint32_t* mask, int mask_is_device,
                         float* out_scores, int64_t* out_ids, int out_is_device) {
    int rc;
    const float* q_dev = q;
    if (!q_is_device) {
        if ((rc = s->qraw.ensure((size_t)B * s->dim * sizeof(float))) != MQVS_OK) return rc;
        CU(cudaMemcpyAsync(s->qraw.p, q, (size_t)B * s->dim * sizeof(float), cudaMemcpyHostToDevice, s->st));
        q_dev = (const float*)s->qraw.p;
    }
    const uint32_t* mask_dev = mask;
    if (mask && !mask_is_device) {
        size_t bytes = (size_t)((s->n + 31) / 32) * sizeof(uint32_t);
        if ((rc = s->maskbuf.ensure(bytes)) != MQVS_OK) return rc;
        CU(cudaMemcpyAsync(s->maskbuf.p, mask, bytes, cudaMemcpyHostToDevice, s->st));
        mask_dev = (const uint32_t*)s->maskbuf.p;
    }
    float* os = out_scores;
    int64_t* oi = out_ids;
    // latency path (B <= 4, k <= 16, the reference's own regime) with host outputs: results and status are written by the
    // kernel into mapped pinned memory -- the call is one H2D copy of the queries, one launch, one stream sync
    const bool zero_copy = !out_is_device && s->host_out != nullptr && !use_tensor_path(s, B, k) && s->path != 3 &&
                           s->metric == MQVS_METRIC_COSINE && mqvs::scan_fused_ok(B, k);
    if (zero_copy) {
        os = (float*)s->host_out;
        oi = (int64_t*)(s->host_out + 64 * 4);
        unsigned int* status = (unsigned int*)(s->host_out + 64 * 4 + 64 * 8);
        if ((rc = enqueue_search(s, q_dev, B, k, mask_dev, 0, os, oi, s->st, status)) != MQVS_OK) return rc;
        CU(cudaStreamSynchronize(s->st));
        if (status[0] != s->epoch) return fail(MQVS_E_CUDA, "latency path: the kernel did not report completion");
        memcpy(out_scores, os, (size_t)B * k * sizeof(float));
        memcpy(out_ids, oi, (size_t)B * k * sizeof(int64_t));
        if (status[1]) return fail(MQVS_E_INVALID, "queries with a non-finite norm in a batch of %d", B);
        return MQVS_OK;
    }
    const size_t n_out = (size_t)B * k;
    float* host_s = nullptr;
    int64_t* host_i = nullptr;
    if (!out_is_device) {
        if ((rc = s->out_s.ensure(n_out * sizeof(float))) != MQVS_OK) return rc;
        if ((rc = s->out_i.ensure(n_out * sizeof(int64_t))) != MQVS_OK) return rc;
        os = (float*)s->out_s.p;
        oi = (int64_t*)s->out_i.p;
        if (s->out_pinned_bytes < n_out * 12) {
            if (s->out_pinned) cudaFreeHost(s->out_pinned);
            s->out_pinned = nullptr;
            s->out_pinned_bytes = 0;
            CU(cudaMallocHost((void**)&s->out_pinned, n_out * 12));
            s->out_pinned_bytes = n_out * 12;
        }
        host_i = (int64_t*)s->out_pinned;
        host_s = (float*)(s->out_pinned + n_out * sizeof(int64_t));
    }
    if ((rc = enqueue_search(s, q_dev, B, k, mask_dev, 0, os, oi, s->st, nullptr, /*defer_repair=*/true)) != MQVS_OK) return rc;
    if (!out_is_device) {
        CU(cudaMemcpyAsync(host_s, os, n_out * sizeof(float), cudaMemcpyDeviceToHost, s->st));
        CU(cudaMemcpyAsync(host_i, oi, n_out * sizeof(int64_t), cudaMemcpyDeviceToHost, s->st));
    }
    CU(cudaMemcpyAsync(s->bad_host, s->bad_dev, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->st));
    const bool tensor = s->stats.last_path == 1;
    if (tensor) {
        if (s->flags_host_n < (size_t)B + 1) {
            if (s->flags_host) cudaFreeHost(s->flags_host);
            s->flags_host = nullptr;
            s->flags_host_n = 0;
            CU(cudaMallocHost(&s->flags_host, ((size_t)B + 1) * sizeof(uint32_t)));
            s->flags_host_n = (size_t)B + 1;
        }
        CU(cudaMemcpyAsync(s->flags_host, s->last_flags_dev, ((size_t)B + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->st));
    }
    CU(cudaStreamSynchronize(s->st));
    if (s->last_fused ? (unsigned int)s->bad_host[1] == s->epoch : s->bad_host[0] != 0)
        return fail(MQVS_E_INVALID, "queries with a non-finite norm in a batch of %d", B);
    if (tensor && s->flags_host[0] > 0) {
        // some shortlists were not proven exact: band pass + exact scan of up to 64 queries on the device, then look again
        enqueue_tensor_repair(s, B, k, mask_dev, 0, os, oi, s->st, (int)((s->flags_host[0] + 127) / 128));
        CU(cudaMemcpyAsync(s->flags_host, s->last_flags_dev, ((size_t)B + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->st));
        CU(cudaStreamSynchronize(s->st));
        // more flagged queries than the device repair takes: re-run the rest on the exact scan from here
        for (int b = 0; b < B; ++b) {
            if (!s->flags_host[1 + b]) continue;  // flags of device-repaired queries were cleared
            uint32_t passes = 0;
            if ((rc = enqueue_simt(s, b, 1, k, mask_dev, 0, os, oi, s->st, &passes)) != MQVS_OK) return rc;
            s->stats.kernel_launches += passes + 1;
            s->host_fallbacks += 1;
        }
        if (!out_is_device) {
            CU(cudaMemcpyAsync(host_s, os, n_out * sizeof(float), cudaMemcpyDeviceToHost, s->st));
            CU(cudaMemcpyAsync(host_i, oi, n_out * sizeof(int64_t), cudaMemcpyDeviceToHost, s->st));
        }
        CU(cudaStreamSynchronize(s->st));
    }
    if (!out_is_device) {
        memcpy(out_scores, host_s, n_out * sizeof(float));
        memcpy(out_ids, host_i, n_out * sizeof(int64_t));
    }
    settle_events(s);
    mqvs::tc::trace_dump();
    return MQVS_OK;
}

MQVS_API int mqvs_search_async(mqvs_t* s, const float* q_dev, int B, int k, const uint32_t* mask_dev, int64_t id_base,
                      float* out_scores_dev, int64_t* out_ids_dev, void* cuda_stream) {
    int rc = check_search_args(s, q_dev, B, k, out_scores_dev, out_ids_dev);
    if (rc != MQVS_OK) return rc;
    if (B == 0) return MQVS_OK;
    std::lock_guard<std::mutex> lk(s->mu);
    CU(cudaSetDevice(s->dev));
    return enqueue_search(s, q_dev, B, k, mask_dev, id_base, out_scores_dev, out_ids_dev, (cudaStream_t)cuda_stream);
}

MQVS_API int mqvs_merge_candidates(int device, const float* scores_dev, const int64_t* ids_dev, int W, int B, int k,
                          float* out_scores_dev, int64_t* out_ids_dev, void* cuda_stream) {
    if (W < 1 || B < 0 || k < 1 || k > MQVS_MAX_K || (long long)W * k > 16384)
        return fail(MQVS_E_INVALID, "bad merge shape W=%d B=%d k=%d (need W*k <= 16384)", W, B, k);
    if (B == 0) return MQVS_OK;
    if (!scores_dev || !ids_dev || !out_scores_dev || !out_ids_dev) return fail(MQVS_E_INVALID, "NULL buffer");
    CU(cudaSetDevice(device));
    mqvs::launch_merge_candidates(scores_dev, ids_dev, W, B, k, out_scores_dev, out_ids_dev, (cudaStream_t)cuda_stream);
    CU(cudaGetLastError());
    return MQVS_OK;
}

MQVS_API int mqvs_get_rows(mqvs_t* s, const int64_t* row_ids, uint64_t n, float* out_host) {
    if (!s) return fail(MQVS_E_INVALID, "NULL handle");
    if (n == 0) return MQVS_OK;
    if (!row_ids || !out_host) return fail(MQVS_E_INVALID, "NULL buffer");
    std::lock_guard<std::mutex> lk(s->mu);
    for (uint64_t i = 0; i < n; ++i)
        if (row_ids[i] < 0 || (uint64_t)row_ids[i] >= s->n)
            return fail(MQVS_E_INVALID, "row id %lld out of range [0, %llu)", (long long)row_ids[i], (unsigned long long)s->n);
    CU(cudaSetDevice(s->dev));
    int rc;
    if ((rc = s->idbuf.ensure(n * sizeof(int64_t))) != MQVS_OK) return rc;
    if ((rc = s->gatherbuf.ensure(n * s->dim * sizeof(float))) != MQVS_OK) return rc;
    CU(cudaMemcpyAsync(s->idbuf.p, row_ids, n * sizeof(int64_t), cudaMemcpyHostToDevice, s->st));
    mqvs::launch_gather_rows(s->rows, s->dtype, (const int64_t*)s->idbuf.p, n, s->dim, (float*)s->gatherbuf.p, s->st);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out_host, s->gatherbuf.p, n * s->dim * sizeof(float), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    s->stats.kernel_launches += 1;
    return MQVS_OK;
}

namespace {
struct FileHeader {
    char magic[8];
    uint32_t version, dim, dtype, dim_pad;
    uint64_t rows;
};
struct NormsHeader {
    char magic[8];
    uint64_t rows;
};

// `<path>.norms`: ||x|| of every raw row, rewritten whole (4 bytes per row) through a temporary + rename
int save_norms(mqvs_store* s, const char* path) {
    if (!s->keep_norms) return MQVS_OK;
    std::vector<float> host((size_t)std::max<uint64_t>(s->n, 1));
    if (s->n) CU(cudaMemcpy(host.data(), s->norms, s->n * sizeof(float), cudaMemcpyDeviceToHost));
    const std::string dst = std::string(path) + ".norms", tmp = dst + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return fail(MQVS_E_IO, "cannot open %s for writing", tmp.c_str());
    NormsHeader h{};
    memcpy(h.magic, "MQVSNRM1", 8);
    h.rows = s->n;
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1 && fwrite(host.data(), sizeof(float), (size_t)s->n, f) == (size_t)s->n;
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), dst.c_str()) != 0) {
        remove(tmp.c_str());
        return fail(MQVS_E_IO, "writing %s failed", dst.c_str());
    }
    return MQVS_OK;
}

// after the rows: norms of rows [0, rows) back into HBM, deviation recomputed
int load_norms(mqvs_store* s, const char* path, uint64_t rows) {
    if (!s->keep_norms) return MQVS_OK;
    const std::string src = std::string(path) + ".norms";
    FILE* f = fopen(src.c_str(), "rb");
    if (!f)
        return fail(MQVS_E_IO, "%s is missing: the collection was saved without row norms (cosine space) and cannot be "
                               "opened in the l2 / ip space", src.c_str());
    NormsHeader h{};
    std::vector<float> host((size_t)std::max<uint64_t>(rows, 1));
    const bool ok = fread(&h, sizeof(h), 1, f) == 1 && memcmp(h.magic, "MQVSNRM1", 8) == 0 && h.rows >= rows &&
                    fread(host.data(), sizeof(float), (size_t)rows, f) == (size_t)rows;
    fclose(f);
    if (!ok) return fail(MQVS_E_IO, "%s does not hold the norms of %llu rows", src.c_str(), (unsigned long long)rows);
    if (rows) CU(cudaMemcpy(s->norms, host.data(), rows * sizeof(float), cudaMemcpyHostToDevice));
    float dev = 0.0f, mx = 0.0f;
    for (uint64_t r = 0; r < rows; ++r) {
        dev = std::max(dev, fabsf(host[r] - 1.0f));
        mx = std::max(mx, host[r]);
    }
    s->norm_dev = dev;
    s->norm_max = mx;
    s->bounds_rows = 0;  // other rows than the chunk bounds were computed from
    return MQVS_OK;
}
}  // namespace

MQVS_API int mqvs_save(mqvs_t* s, const char* path) {
    if (!s || !path) return fail(MQVS_E_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(s->mu);
    CU(cudaSetDevice(s->dev));
    std::string tmp = std::string(path) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return fail(MQVS_E_IO, "cannot open %s for writing", tmp.c_str());
    FileHeader h{};
    memcpy(h.magic, "MQVSROW1", 8);
    h.version = MQVS_VERSION;
    h.dim = (uint32_t)s->dim;
    h.dtype = (uint32_t)s->dtype;
    h.dim_pad = MQVS_DIM_PAD;
    h.rows = s->n;
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
    const uint64_t chunk = 65536;
    std::vector<char> host((size_t)std::min<uint64_t>(chunk, std::max<uint64_t>(s->n, 1)) * s->row_bytes());
    for (uint64_t r = 0; ok && r < s->n; r += chunk) {
        uint64_t m = std::min(chunk, s->n - r);
        cudaError_t e = cudaMemcpy(host.data(), (char*)s->rows + r * s->row_bytes(), m * s->row_bytes(), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            fclose(f);
            remove(tmp.c_str());
            return fail(MQVS_E_CUDA, "device-to-host copy failed: %s", cudaGetErrorString(e));
        }
        ok = fwrite(host.data(), s->row_bytes(), m, f) == m;
    }
    ok = (fclose(f) == 0) && ok;
    if (ok) {  // norms first: a row file is never newer than its norms
        int rc = save_norms(s, path);
        if (rc != MQVS_OK) {
            remove(tmp.c_str());
            return rc;
        }
    }
    if (!ok || rename(tmp.c_str(), path) != 0) {
        remove(tmp.c_str());
        return fail(MQVS_E_IO, "writing %s failed", path);
    }
    return MQVS_OK;
}

MQVS_API int mqvs_save_append(mqvs_t* s, const char* path, uint64_t first_row) {
    if (!s || !path) return fail(MQVS_E_INVALID, "NULL argument");
    if (first_row == 0) return mqvs_save(s, path);
    std::lock_guard<std::mutex> lk(s->mu);
    CU(cudaSetDevice(s->dev));
    if (first_row > s->n) return fail(MQVS_E_INVALID, "first_row %llu beyond the %llu stored rows", (unsigned long long)first_row, (unsigned long long)s->n);
    FILE* f = fopen(path, "r+b");
    if (!f) return fail(MQVS_E_IO, "cannot open %s for appending", path);
    FileHeader h{};
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "MQVSROW1", 8) != 0) {
        fclose(f);
        return fail(MQVS_E_IO, "%s is not an mqvs row file", path);
    }
    if ((int)h.dim != s->dim || (int)h.dtype != s->dtype || h.dim_pad != MQVS_DIM_PAD || h.rows != first_row) {
        fclose(f);
        return fail(MQVS_E_INVALID, "%s holds %llu rows (dim=%u dtype=%u); cannot append from row %llu", path,
                    (unsigned long long)h.rows, h.dim, h.dtype, (unsigned long long)first_row);
    }
    bool ok = fseek(f, (long)(sizeof(h) + first_row * s->row_bytes()), SEEK_SET) == 0;
    const uint64_t chunk = 65536;
    std::vector<char> host((size_t)std::min<uint64_t>(chunk, std::max<uint64_t>(s->n - first_row, 1)) * s->row_bytes());
    for (uint64_t r = first_row; ok && r < s->n; r += chunk) {
        uint64_t m = std::min(chunk, s->n - r);
        cudaError_t e = cudaMemcpy(host.data(), (char*)s->rows + r * s->row_bytes(), m * s->row_bytes(), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            fclose(f);
            return fail(MQVS_E_CUDA, "device-to-host copy failed: %s", cudaGetErrorString(e));
        }
        ok = fwrite(host.data(), s->row_bytes(), m, f) == m;
    }
    if (ok) {  // holds the norms of MORE rows than the header claims until the count below moves: load accepts that
        int rc = save_norms(s, path);
        if (rc != MQVS_OK) {
            fclose(f);
            return rc;
        }
    }
    // the row count in the header moves last: a crash before this leaves the old, consistent file (plus ignored bytes)
    h.rows = s->n;
    ok = ok && fflush(f) == 0 && fseek(f, 0, SEEK_SET) == 0 && fwrite(&h, sizeof(h), 1, f) == 1;
    ok = (fclose(f) == 0) && ok;
    if (!ok) return fail(MQVS_E_IO, "appending to %s failed", path);
    return MQVS_OK;
}

MQVS_API int mqvs_load(mqvs_t* s, const char* path) {
    if (!s || !path) return fail(MQVS_E_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(s->mu);
    CU(cudaSetDevice(s->dev));
    FILE* f = fopen(path, "rb");
    if (!f) return fail(MQVS_E_IO, "cannot open %s", path);
    FileHeader h{};
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "MQVSROW1", 8) != 0) {
        fclose(f);
        return fail(MQVS_E_IO, "%s is not an mqvs row file", path);
    }
    if ((int)h.dim != s->dim || (int)h.dtype != s->dtype || h.dim_pad != MQVS_DIM_PAD) {
        fclose(f);
        return fail(MQVS_E_INVALID, "%s holds dim=%u dtype=%u, store is dim=%d dtype=%d", path, h.dim, h.dtype, s->dim, s->dtype);
    }
    s->n = 0;
    int rc = grow(s, h.rows);
    if (rc != MQVS_OK) {
        fclose(f);
        return rc;
    }
    const uint64_t chunk = 65536;
    std::vector<char> host((size_t)std::min<uint64_t>(chunk, std::max<uint64_t>(h.rows, 1)) * s->row_bytes());
    for (uint64_t r = 0; r < h.rows; r += chunk) {
        uint64_t m = std::min(chunk, h.rows - r);
        if (fread(host.data(), s->row_bytes(), m, f) != m) {
            fclose(f);
            return fail(MQVS_E_IO, "%s is truncated", path);
        }
        cudaError_t e = cudaMemcpy((char*)s->rows + r * s->row_bytes(), host.data(), m * s->row_bytes(), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            fclose(f);
            return fail(MQVS_E_CUDA, "host-to-device copy failed: %s", cudaGetErrorString(e));
        }
    }
    fclose(f);
    if ((rc = load_norms(s, path, h.rows)) != MQVS_OK) return rc;
    s->n = h.rows;
    s->stats.rows = s->n;
    return MQVS_OK;
}

MQVS_API int mqvs_column_set(mqvs_t* s, int col_id, int col_type, const void* values_host, const uint8_t* present_host,
                             uint64_t n) {
    if (!s) return fail(MQVS_E_INVALID, "NULL handle");
    if (col_id < 0 || col_id >= mqvs::kFilterMaxCols) return fail(MQVS_E_INVALID, "column id must be in [0, %d)", mqvs::kFilterMaxCols);
    if (col_type != MQVS_COL_I64 && col_type != MQVS_COL_F64) return fail(MQVS_E_INVALID, "unknown column type %d", col_type);
    if (n && !values_host) return fail(MQVS_E_INVALID, "values is NULL");
    std::lock_guard<std::mutex> lk(s->mu);
    CU(cudaSetDevice(s->dev));
    auto& c = s->columns[col_id];
    int rc;
    if ((rc = c.values.ensure(std::max<uint64_t>(n, 1) * 8)) != MQVS_OK) return rc;
    if (n) CU(cudaMemcpyAsync(c.values.p, values_host, n * 8, cudaMemcpyHostToDevice, s->st));
    c.has_present = present_host != nullptr;
    if (present_host) {
        const uint64_t words = (n + 31) / 32;
        std::vector<uint32_t> bits(std::max<uint64_t>(words, 1), 0u);
        for (uint64_t r = 0; r < n; ++r)
            if (present_host[r]) bits[r >> 5] |= 1u << (r & 31);
        if ((rc = c.present.ensure(bits.size() * 4)) != MQVS_OK) return rc;
        CU(cudaMemcpyAsync(c.present.p, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice, s->st));
        CU(cudaStreamSynchronize(s->st));  // `bits` goes out of scope
    }
    CU(cudaStreamSynchronize(s->st));
    c.n = n;
    c.type = col_type;
    c.used = true;
    s->f_table_dirty = true;
    return MQVS_OK;
}

MQVS_API int mqvs_filter_build(mqvs_t* s, const mqvs_filter_op* prog, int n_ops, const void* pool, uint64_t pool_len,
                               const uint32_t* live_words_host, const uint32_t** mask_dev_out, uint64_t* allowed_out) {
    using namespace mqvs;
    if (!s || !mask_dev_out || !allowed_out) return fail(MQVS_E_INVALID, "NULL argument");
    if (n_ops < 0 || n_ops > kFilterMaxOps) return fail(MQVS_E_INVALID, "a filter program holds at most %d operations", kFilterMaxOps);
    if (n_ops && !prog) return fail(MQVS_E_INVALID, "prog is NULL");
    std::lock_guard<std::mutex> lk(s->mu);
    CU(cudaSetDevice(s->dev));
    // validate: known columns, sets inside the pool, a well-formed postfix expression
    int depth = 0;
    for (int i = 0; i < n_ops; ++i) {
        const mqvs_filter_op& o = prog[i];
        if (o.op == MQVS_OP_AND || o.op == MQVS_OP_OR) {
            if (depth < 2) return fail(MQVS_E_INVALID, "filter program: operator %d at %d lacks operands", o.op, i);
            --depth;
        } else if (o.op >= MQVS_OP_EQ && o.op <= MQVS_OP_NIN) {
            if (o.col < 0 || o.col >= kFilterMaxCols || !s->columns[o.col].used)
                return fail(MQVS_E_INVALID, "filter program: column %d was never set", o.col);
            if ((o.op == MQVS_OP_IN || o.op == MQVS_OP_NIN) && (uint64_t)o.set_off + o.set_len > pool_len)
                return fail(MQVS_E_INVALID, "filter program: set of operation %d leaves the constant pool", i);
            if (++depth > 32) return fail(MQVS_E_INVALID, "filter program nests deeper than 32");
        } else {
            return fail(MQVS_E_INVALID, "filter program: unknown operation %d", o.op);
        }
    }
    if (n_ops && depth != 1) return fail(MQVS_E_INVALID, "filter program leaves %d values on the stack", depth);
    const uint64_t words = (s->n + 31) / 32;
    int rc;
    if (!s->f_count_dev) {
        CU(cudaMalloc(&s->f_count_dev, sizeof(unsigned long long)));
        CU(cudaMallocHost(&s->f_count_host, sizeof(unsigned long long)));
    }
    if (s->f_table_dirty) {
        FilterColumn table[kFilterMaxCols];
        for (int i = 0; i < kFilterMaxCols; ++i) {
            const auto& c = s->columns[i];
            table[i] = FilterColumn{c.values.p, c.has_present ? (const uint32_t*)c.present.p : nullptr, c.used ? c.n : 0, c.type};
        }
        if ((rc = s->f_table.ensure(sizeof(table))) != MQVS_OK) return rc;
        CU(cudaMemcpyAsync(s->f_table.p, table, sizeof(table), cudaMemcpyHostToDevice, s->st));
        CU(cudaStreamSynchronize(s->st));  // `table` is a stack array
        s->f_table_dirty = false;
    }
    if ((rc = s->f_prog.ensure(std::max(n_ops, 1) * sizeof(mqvs_filter_op))) != MQVS_OK) return rc;
    if ((rc = s->f_pool.ensure(std::max<uint64_t>(pool_len, 1) * 8)) != MQVS_OK) return rc;
    if ((rc = s->f_mask.ensure(std::max<uint64_t>(words, 1) * 4)) != MQVS_OK) return rc;
    if (n_ops) CU(cudaMemcpyAsync(s->f_prog.p, prog, n_ops * sizeof(mqvs_filter_op), cudaMemcpyHostToDevice, s->st));
    if (pool_len) CU(cudaMemcpyAsync(s->f_pool.p, pool, pool_len * 8, cudaMemcpyHostToDevice, s->st));
    if (live_words_host) {
        if ((rc = s->f_live.ensure(std::max<uint64_t>(words, 1) * 4)) != MQVS_OK) return rc;
        if (words) CU(cudaMemcpyAsync(s->f_live.p, live_words_host, words * 4, cudaMemcpyHostToDevice, s->st));
    }
    CU(cudaMemsetAsync(s->f_count_dev, 0, sizeof(unsigned long long), s->st));
    FilterArgs a{};
    a.prog = (const mqvs_filter_op*)s->f_prog.p;
    a.n_ops = n_ops;
    a.cols = (const FilterColumn*)s->f_table.p;
    a.pool = s->f_pool.p;
    a.live = live_words_host ? (const uint32_t*)s->f_live.p : nullptr;
    a.n_rows = s->n;
    a.n_words = words;
    a.mask = (uint32_t*)s->f_mask.p;
    a.allowed = s->f_count_dev;
    launch_filter_eval(a, s->st);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(s->f_count_host, s->f_count_dev, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    s->stats.kernel_launches += words ? 1 : 0;
    *mask_dev_out = (const uint32_t*)s->f_mask.p;
    *allowed_out = *s->f_count_host;
    return MQVS_OK;
}

MQVS_API int mqvs_get_stats(mqvs_t* s, mqvs_stats* out) {
    if (!s || !out) return fail(MQVS_E_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(s->mu);
    cudaSetDevice(s->dev);
    settle_events(s);
    s->stats.rows = s->n;
    unsigned long long c[4] = {0, 0, 0, 0};
    if (s->fix_counters && cudaMemcpy(c, s->fix_counters, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) cudaGetLastError();
    s->stats.band_repairs = c[2];
    s->stats.exact_fallbacks = c[0] + s->host_fallbacks;  // async callers: c[1] - host_fallbacks queries stayed unrepaired
    s->stats.unrepaired = c[1] > s->host_fallbacks ? c[1] - s->host_fallbacks : 0;
    *out = s->stats;
    return MQVS_OK;
}

MQVS_API int mqvs_set_metric(mqvs_t* s, int metric) {
    if (!s) return fail(MQVS_E_INVALID, "NULL handle");
    if (metric != MQVS_METRIC_COSINE && metric != MQVS_METRIC_L2 && metric != MQVS_METRIC_IP)
        return fail(MQVS_E_INVALID, "unknown metric %d", metric);
    std::lock_guard<std::mutex> lk(s->mu);
    if (metric != MQVS_METRIC_COSINE && !s->keep_norms) {
        if (s->n)
            return fail(MQVS_E_INVALID, "the %llu rows already stored were added without their norms: choose the l2 / ip "
                                        "metric while the store is empty", (unsigned long long)s->n);
        CU(cudaSetDevice(s->dev));
        if (s->cap && cudaMalloc(&s->norms, s->cap * sizeof(float)) != cudaSuccess) {
            cudaGetLastError();
            s->norms = nullptr;
            return fail(MQVS_E_OOM, "cudaMalloc of %llu row norms failed", (unsigned long long)s->cap);
        }
        s->keep_norms = true;
    }
    s->metric = metric;
    return MQVS_OK;
}

MQVS_API int mqvs_norm_deviation(mqvs_t* s, float* out) {
    if (!s || !out) return fail(MQVS_E_INVALID, "NULL argument");
    std::lock_guard<std::mutex> lk(s->mu);
    *out = s->norm_dev;
    return MQVS_OK;
}

MQVS_API int mqvs_get_norms(mqvs_t* s, uint64_t first_row, uint64_t n, float* out_host) {
    if (!s) return fail(MQVS_E_INVALID, "NULL handle");
    if (n == 0) return MQVS_OK;
    if (!out_host) return fail(MQVS_E_INVALID, "NULL buffer");
    std::lock_guard<std::mutex> lk(s->mu);
    if (!s->keep_norms) return fail(MQVS_E_INVALID, "norms are kept only for the l2 / ip metrics (mqvs_set_metric)");
    if (first_row > s->n || n > s->n - first_row)
        return fail(MQVS_E_INVALID, "rows [%llu, %llu) out of range [0, %llu)", (unsigned long long)first_row,
                    (unsigned long long)(first_row + n), (unsigned long long)s->n);
    CU(cudaSetDevice(s->dev));
    CU(cudaStreamSynchronize(s->st));
    CU(cudaMemcpy(out_host, s->norms + first_row, n * sizeof(float), cudaMemcpyDeviceToHost));
    return MQVS_OK;
}

MQVS_API int mqvs_set_path(mqvs_t* s, int path) {
    if (!s) return fail(MQVS_E_INVALID, "NULL handle");
    if (path < 0 || path > 4) return fail(MQVS_E_INVALID, "path must be 0..4");
    std::lock_guard<std::mutex> lk(s->mu);
    s->path = path;
    return MQVS_OK;
}

}  // extern "C"
