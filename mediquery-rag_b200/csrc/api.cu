// api.cu -- the C ABI of libmqvs.so (include/mqvs.h): store lifetime, ingest, search orchestration, persistence.
// No torch types, no exceptions across the boundary; CUDA errors become MQVS_E_CUDA with a thread-local message.
#include <fcntl.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <thread>

#include "store.h"

namespace {

thread_local std::string g_err;

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

#define CU MQVS_CU
#define fail ::mqvs::mqvs_fail
using mqvs::DevBuf;
using mqvs::elem_size;

namespace mqvs {
namespace api {

constexpr int kMaxFixRound = 1024;  // flagged queries one exact-scan repair round can take (scan_fix_kernel's list)
constexpr int kWideChunk = 256;    // queries per pass of a k > 128 search (the single-launch tensor pipeline's two query tiles)  // capacity of the device-side repair list (flagged queries per exact-scan round)

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
static void settle_slot(mqvs_store* s, int slot) {
    auto& ev = s->evr[slot];
    if (!ev.pending) return;
    const bool scan_only = s->timing_level == 2;
    if (cudaEventSynchronize(ev.e[scan_only ? 2 : 3]) == cudaSuccess) {
        float total = 0.f, scan = 0.f;
        cudaEventElapsedTime(&scan, ev.e[1], ev.e[2]);
        if (scan_only) total = scan;
        else cudaEventElapsedTime(&total, ev.e[0], ev.e[3]);
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
static bool use_tensor_path(const mqvs_store* s, int B, int k) {
    if (s->path == 1 || s->path == 3) return false;
    if (s->n < 4096) return false;
    // k > 128: wide shortlists exist in the single-launch pipeline only (larger batches run it 256 queries at a time)
    if (k > 128 && (s->path == 4 || !s->coop_ok || !mqvs::tc::fused_ok(1, k, s->sm_count))) return false;
    if (s->path == 2 || s->path == 4) return true;
    // measured on B200 (profiles/r01_*): the register-resident SIMT scan holds the HBM roofline for 1 query (bf16 rows)
    // or 1-2 queries (fp32 rows); beyond that its multi-query variants fall to 2-3.4 TB/s while the tensor path streams
    // the corpus once for up to 128 queries at 6.3-7.1 TB/s.
    // From 1M rows on even single queries take the (single-launch) tensor pipeline: same bandwidth as the SIMT scan when
    // the SM clock is at boost (0.471 vs 0.487 ms at 1M x 768 fp32, 2.67 vs 2.68 ms at 12.5M bf16), and it keeps it when
    // the clock drops under a power cap -- on 8 power-capped GPUs the SIMT scan alternated between 2.72 and 3.17 ms per
    // 12.5M-row shard while the TMA-fed pipeline stayed at 2.71-2.75 (profiles/r01_b1_paths_w8.json).  Below 1M rows the
    // SIMT path's single launch wins on latency.
    int simt_max = s->dtype == MQVS_F32 ? 2 : 1;
    // fp32 store with its bf16 scoring shadow (half the bytes per scan): the tensor pipelines overtake the fp32 SIMT scan for
    // single queries at ~100k rows and for 2 queries at every size measured (profiles/r02_shadow_crossover.json: B=1 at 131k /
    // 262k / 524k rows: 101 / 127 / 183 us vs 109 / 168 / 299 us; B=2 at 65k rows: 85 vs 103 us)
    const bool shadow = s->dtype == MQVS_F32 && s->shadow_on && s->metric == MQVS_METRIC_COSINE;
    if (shadow) simt_max = 1;
    if (shadow && s->n >= 131072ull) return true;
    // (l2 / ip spaces: the transform makes the tensor epilogue ~4 % slower than the SIMT scan for the 1-2 queries that scan
    // holds the roofline for -- 0.508 vs 0.489 ms at 1M x 768 fp32, B=1 -- so those stay on the SIMT scan at every size)
    if (s->n >= 1000000ull && (s->metric == MQVS_METRIC_COSINE || B > simt_max)) return true;
    if (B > simt_max && s->n >= 65536) return true;
    // Everything else from 4096 rows up (round 2, profiles/r02_small_store_crossover.json: blocking host-buffer calls on one
    // B200, every space, both row types): the single-launch tensor pipeline beats both SIMT forms for EVERY batch of two or
    // more queries -- 8192 bf16 rows: B = 2 64 vs 78 us, B = 64 83 vs 504 us, B = 1024 0.40 vs 7.1 ms (the SIMT scan takes at
    // most four queries per pass); fp32 pairs below 1M rows, l2 space, 131072 rows: 100 vs 132 us.  Without cooperative
    // launches the segmented pipeline's fixed launches pay from 16 queries on.  Single queries stay on the latency kernel
    // (52.9 vs 61.1 us at 4096 rows, a tie from 8192).
    const bool one_launch = s->coop_ok && B <= 256 && mqvs::tc::fused_ok(B, k, s->sm_count);
    return B >= 2 && k <= 128 && (one_launch || B >= 16);
}

// Brings the bf16 scoring shadow of an fp32 store up to date (store.h).  -> true when the tensor pipelines may stream it.
// Allocation follows the store's capacity; when HBM cannot hold the extra 50 % the store goes on scoring its fp32 rows.
static bool shadow_sync(mqvs_store* s) {
    if (!s->shadow_on || s->dtype != MQVS_F32 || s->n == 0) return false;
    if (s->shadow_cap < s->n) {
        if (s->shadow) cudaFree(s->shadow);  // (cudaFree waits for work that may still read it)
        s->shadow = nullptr;
        s->shadow_cap = s->shadow_rows = 0;
        if (cudaMalloc(&s->shadow, s->cap * (size_t)MQVS_DIM_PAD * 2) != cudaSuccess) {
            cudaGetLastError();
            s->shadow = nullptr;
            s->shadow_on = false;
            return false;
        }
        s->shadow_cap = s->cap;
    }
    if (!s->shadow_err && cudaMalloc(&s->shadow_err, sizeof(unsigned int)) != cudaSuccess) {
        cudaGetLastError();
        s->shadow_err = nullptr;
        return false;
    }
    if (s->shadow_rows < s->n) {
        if (s->shadow_rows == 0 && cudaMemsetAsync(s->shadow_err, 0, sizeof(unsigned int), s->st) != cudaSuccess) return false;
        mqvs::launch_shadow_build((const float*)s->rows, s->shadow_rows, s->n - s->shadow_rows, s->shadow, s->shadow_err, s->st);
        // searches may run on a caller's stream: the conversion (rare: once per add) is finished before any of them starts
        if (cudaStreamSynchronize(s->st) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        s->stats.kernel_launches += 1;
        s->shadow_rows = s->n;
    }
    return true;
}

// Will this search run the single-launch latency kernel (every CTA reads the raw queries itself)?
bool takes_latency_kernel(const mqvs_store* s, int B, int k) {
    return !use_tensor_path(s, B, k) && s->path != 3 && s->metric == MQVS_METRIC_COSINE && mqvs::scan_fused_ok(B, k);
}

// Exact SIMT scan of queries [b0, b0 + nb) of s->qn into out rows [b0, ...).  Returns passes launched.
static int enqueue_simt(mqvs_store* s, int b0, int nb, int k, const uint32_t* mask_dev, int64_t id_base, float* out_scores_dev,
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
// Exact-scan rounds take s->max_device_fix flagged queries each; a caller that cannot look at the flags (mqvs_search_async)
// asks for ceil(B / max_device_fix) rounds so that EVERY flagged query is repaired whatever happens (idle rounds exit at once).
static void enqueue_tensor_repair(mqvs_store* s, int B, int k, const uint32_t* mask_dev, int64_t id_base, float* out_scores_dev,
                                  int64_t* out_ids_dev, cudaStream_t st, int band_rounds, int fix_rounds) {
    using namespace mqvs;
    const tc::Buffers& b = s->last_b;
    static const bool no_band = getenv("MQVS_NO_BAND") != nullptr;  // developer knob: measure what the band pass costs
    if (!no_band && s->coop_ok && s->path != 4 && tc::fused_ok(1, k, s->sm_count)) {
        for (int r = 0; r < band_rounds; ++r) {  // 128 flagged queries per round
            const char* err = tc::search_band(b, s->last_plan, s->last_score_rows, s->last_score_dtype, (uint32_t)s->n, (const float*)s->qn.p, B, k,
                                              mask_dev, id_base, out_scores_dev, out_ids_dev, s->fix_counters, st, s->last_space);
            if (err) {
                cudaGetLastError();  // a refused cooperative launch: the exact scan below still repairs everything
                break;
            }
            s->stats.kernel_launches += 2;
        }
    }
    uint32_t* fix_list = s->fix_dev;
    uint32_t* fix_n = s->fix_dev + kMaxFixRound;
    for (int r = 0; r < fix_rounds; ++r) {
        launch_scan_fix(s->rows, s->dtype, (uint32_t)s->n, (const float*)s->qn.p, b.flags, B, s->max_device_fix, fix_list, fix_n,
                        s->fix_counters, mask_dev, k, (uint64_t*)s->tc_fixcand.p, s->sm_count, st, s->last_space.metric,
                        s->last_space.norms, s->last_space.qnorm, /*count_leftover=*/r + 1 == fix_rounds, b.flag_count);
        launch_merge_fix((const uint64_t*)s->tc_fixcand.p, s->max_device_fix, scan_fix_grid(s->sm_count) * k, k, id_base,
                         fix_list, fix_n, b.flags, out_scores_dev, out_ids_dev, st, b.flag_count);
        s->stats.kernel_launches += 2;
    }
}

// Enqueue one search on `st`; every pointer is device memory.  q_raw [B, dim] is normalised into s->qn first.
// report_to_host: the caller will synchronise and read the outcome on the host -- the tensor pipelines then keep their
// flag words in mapped pinned memory (s->flags_map) and the fused kernels report non-finite queries there too
// (flags_map word [1 + B]), so the blocking call needs no device-to-host copy at all.
int enqueue_search(mqvs_store* s, const float* q_raw_dev, int B, int k, const uint32_t* mask_dev, int64_t id_base,
                   float* out_scores_dev, int64_t* out_ids_dev, cudaStream_t st, unsigned int* host_status,
                   bool defer_repair, bool report_to_host, const float* q_raw_host) {
    using namespace mqvs;
    const int slot = s->ev_cursor;
    s->ev_cursor = (s->ev_cursor + 1) % mqvs_store::kEvRing;
    settle_slot(s, slot);  // oldest entry of the ring: long finished unless 32 searches are in flight
    cudaEvent_t* ev = s->evr[slot].e;
// (the shards of a multi-device handle skip the four timing events per search: at eight devices driven by one host thread
// every API call is on the critical path of the last device's start)
#define MQVS_EV(i_) do { if (s->timing && (s->timing_level != 2 || (i_) == 1 || (i_) == 2)) CU(cudaEventRecord(ev[i_], st)); } while (0)
    const bool tensor = use_tensor_path(s, B, k);
    const bool fused = !tensor && s->path != 3 && s->metric == MQVS_METRIC_COSINE && scan_fused_ok(B, k);
    s->last_fused = fused;
    tc::Plan plan{};
    int rc;
    if (fused) {
        // latency path: normalisation, scan, top-k and the final merge in one launch
        if ((rc = s->cand.ensure((size_t)B * scan_fused_max_grid(s->sm_count) * k * sizeof(uint64_t))) != MQVS_OK) return rc;
        // A blocking host-buffer call (host_status: the caller watches the kernel's completion word) is timed by that caller on
        // the host clock: two timed events cost ~6 us of stream time around a kernel that takes 17 us on the reference's 154
        // chunks (profiles/r02_latency_ab.log: 38.6 -> 32.6 us per call).  MQVS_TIMING=1 / 2 brings the events back.
        const bool events = s->timing && (host_status == nullptr || s->timing_from_env);
        if (events) {
            MQVS_EV(0);
            MQVS_EV(1);
        }
        // (q_raw_host: a host caller's queries ride in the kernel's parameter block -- no copy of their own)
        launch_scan_fused(s->rows, s->dtype, (uint32_t)s->n, q_raw_dev, s->dim, B, mask_dev, k, (uint64_t*)s->cand.p,
                          out_scores_dev, out_ids_dev, id_base, s->ticket_dev, (unsigned int*)(s->bad_dev + 1), ++s->epoch,
                          s->sm_count, st, host_status, q_raw_host);
        if (events) {
            MQVS_EV(2);
            MQVS_EV(3);
        }
        CU(cudaGetLastError());
        s->evr[slot].pending = events;
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
    // Two query tiles over a long corpus: the segmented pipeline's CTA-pair kernel (cta_group::2: each SM stages half a corpus
    // tile) scans at the tensor roof where the single-launch kernel's two-tile ring is bound by the SM's L2 port -- 3.95 vs
    // 4.3-4.5 ms at 12.5M bf16 rows, B = 256 (profiles/r02_*); its extra launches (~0.15 ms) win from ~3M rows on.
    if (tc_one && s->path == 0 && B > 128 && k <= 128 && s->n >= 3000000ull) tc_one = false;
    // fp32 stores: the tensor pipelines stream the bf16 shadow (exact scores still come from the fp32 rows)
    const void* score_rows = s->rows;
    int score_dtype = s->dtype;
    const bool shadowed = tensor && shadow_sync(s);
    if (shadowed) {
        score_rows = s->shadow;
        score_dtype = MQVS_BF16;
    }
    if (tensor) {
        plan = tc::plan(k, B, s->sm_count);
        const size_t small = ((size_t)plan.b_pad * 2 + (size_t)B * 4 + 128 + 256 + 32 + 144) * sizeof(uint32_t);
        if ((rc = s->tc_qop.ensure(tc::op_bytes(plan, s->dtype))) != MQVS_OK) return rc;
        if ((rc = s->tc_small.ensure(small)) != MQVS_OK) return rc;
        if ((rc = s->tc_cnt.ensure(tc::cnt_bytes(plan))) != MQVS_OK) return rc;
        if ((rc = s->tc_lists.ensure(tc::lists_bytes(plan))) != MQVS_OK) return rc;
        if ((rc = s->tc_keys.ensure((size_t)plan.b_pad * plan.kprime * sizeof(uint64_t) * 2)) != MQVS_OK) return rc;
        // exact-scan repair rounds: 64 flagged queries each for small batches, up to 256 for large ones (fewer idle rounds)
        // (very large batches with small k: 1024 per round -- a blind asynchronous search enqueues ceil(B / m) rounds, and even an
        // idle round is two launches: 128 of them per 16384-query batch were 0.45 ms of config 5's 24 ms step)
        s->max_device_fix = B <= 64 ? 64 : (B <= 1024 ? 128 : (B <= 4096 || k > 32 ? 256 : kMaxFixRound));
        if ((rc = s->tc_fixcand.ensure((size_t)s->max_device_fix * scan_fix_grid(s->sm_count) * k * sizeof(uint64_t))) != MQVS_OK) return rc;
        if (report_to_host && ((size_t)B + 4) * sizeof(uint32_t) > s->flags_map.bytes) {
            if ((rc = s->flags_map.ensure(((size_t)B + 4) * sizeof(uint32_t) * 2)) != MQVS_OK) return rc;
            memset(s->flags_map.p, 0, s->flags_map.bytes);
        }
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
        b.flag_count = (uint32_t*)(b.eps + B);   // flag_count immediately followed by flags[B]
        b.flags = b.flag_count + 1;
        b.gbar = (unsigned int*)(b.flags + B);  // [4]
        if (report_to_host) {  // the host reads them after its stream sync: mapped pinned words instead (written rarely)
            b.flag_count = (uint32_t*)s->flags_map.p;  // [0] flag count, [1] non-finite-query epoch, [2 ..] flags[B]
            b.flags = b.flag_count + 2;
        }
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
        if (shadowed) {
            space.exact_rows = s->rows;
            space.shadow_err = s->shadow_err;
        }
        s->last_score_rows = score_rows;
        s->last_score_dtype = score_dtype;
        s->last_b = b;
        s->last_plan = plan;
        s->last_space = space;
    }
    auto enqueue_repair = [&]() {
        // nobody can look at the flags here: one band round per 128 queries of the batch, at most two, then as many exact-scan
        // rounds as it takes to cover all B queries -- the asynchronous search is exact on its own, whatever gets flagged
        // (measured and rejected in round 2: the chain as ONE conditional-graph launch -- a gate kernel reading the flag count
        // + an IF body -- for searches that repeat: 281.4 vs 282.6 us per step at 1M rows, B = 64, and 10 us SLOWER at B = 1;
        // idle launches pipeline behind the running kernel and cost next to nothing, a graph launch does not)
        if (!defer_repair)
            enqueue_tensor_repair(s, B, k, mask_dev, id_base, out_scores_dev, out_ids_dev, st, B > 128 ? 2 : 1,
                                  (B + s->max_device_fix - 1) / s->max_device_fix);
    };
    // where the fused kernels raise the epoch when a query is not finite
    unsigned int* bad_epoch = (unsigned int*)(s->bad_dev + 1);
    if (report_to_host && tc_one) bad_epoch = (unsigned int*)s->flags_map.p + 1;
    if (tc_one) {
        s->last_fused = true;  // non-finite queries are reported through the epoch word, as on the latency path
        MQVS_EV(0);
        MQVS_EV(1);
        bool coop_refused = false;
        const char* err = tc::search_fused(b, plan, score_rows, score_dtype, (uint32_t)s->n, q_raw_dev, s->dim, (float*)s->qn.p, B, k,
                                           mask_dev, id_base, out_scores_dev, out_ids_dev, bad_epoch,
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
        MQVS_EV(2);  // brackets prep + the single scan/select/rescore launch
        enqueue_repair();
        MQVS_EV(3);
        CU(cudaGetLastError());
        s->evr[slot].pending = s->timing;
        s->stats.kernel_launches += 2;
        s->stats.scan_launches += 1;
        s->stats.scan_bytes += s->n * (uint64_t)MQVS_DIM_PAD * elem_size(score_dtype);  // what the scan streamed (the shadow when there is one)
        s->stats.last_scan_passes = 1;
        s->stats.last_path = 1;
        s->stats.last_pipeline = 3;
        s->stats.searches += 1;
        s->stats.queries += (uint64_t)B;
        return MQVS_OK;
    }

    MQVS_EV(0);
    CU(cudaMemsetAsync(s->bad_dev, 0, sizeof(unsigned long long), st));
    launch_normalize(q_raw_dev, (uint64_t)B, s->dim, MQVS_F32, s->qn.p, s->bad_dev, st,
                     s->metric != MQVS_METRIC_COSINE ? (float*)s->qnorm.p : nullptr);
    s->stats.kernel_launches += 1;
    uint32_t passes = 0;
    if (tensor) {
        MQVS_EV(1);
        const char* err = tc::search(b, plan, score_rows, score_dtype, (uint32_t)s->n, (const float*)s->qn.p, B, k, mask_dev,
                                     id_base, out_scores_dev, out_ids_dev, st, space);
        if (err) return fail(MQVS_E_CUDA, "tensor-core search: %s", err);
        MQVS_EV(2);  // brackets prep + score + select + rescore + finalize
        enqueue_repair();
        passes = 1;
        s->stats.kernel_launches += tc::launches_per_search(plan, (uint32_t)s->n, B);
        s->stats.last_path = 1;
        s->stats.last_pipeline = 2;
    } else {
        MQVS_EV(1);
        // scan launches are bracketed together with the (tiny) candidate memset
        if ((rc = enqueue_simt(s, 0, B, k, mask_dev, id_base, out_scores_dev, out_ids_dev, st, &passes)) != MQVS_OK) return rc;
        MQVS_EV(2);
        s->stats.kernel_launches += passes + 1;
        s->stats.last_path = 0;
        s->stats.last_pipeline = 0;
    }
    MQVS_EV(3);
    CU(cudaGetLastError());

    s->evr[slot].pending = s->timing;
    s->stats.scan_launches += passes;
    s->stats.scan_bytes += (uint64_t)passes * s->n * (uint64_t)MQVS_DIM_PAD * elem_size(score_dtype);
    s->stats.last_scan_passes = passes;
    s->stats.searches += 1;
    s->stats.queries += (uint64_t)B;
    return MQVS_OK;
}

}  // namespace api
}  // namespace mqvs

using namespace mqvs::api;

extern "C" {

MQVS_API int mqvs_version(void) { return MQVS_VERSION; }
MQVS_API int mqvs_dim_pad(void) { return MQVS_DIM_PAD; }

MQVS_API const char* mqvs_last_error(void) { return g_err.c_str(); }

MQVS_API int mqvs_device_count(int* out) {
    if (!out) return fail(MQVS_E_INVALID, "out is NULL");
    int n = 0;
    CU(cudaGetDeviceCount(&n));
    *out = n;
    return MQVS_OK;
}

MQVS_API int mqvs_create(int n_dev, const int* dev_ids, int dim, int store_dtype, uint64_t capacity_rows, mqvs_t** out) {
    MQVS_RANGE("mqvs_create");
    if (!out) return fail(MQVS_E_INVALID, "out is NULL");
    *out = nullptr;
    if (n_dev < 1 || n_dev > MQVS_MAX_DEVICES)
        return fail(MQVS_E_INVALID, "n_dev must be in [1, %d], got %d", MQVS_MAX_DEVICES, n_dev);
    if (n_dev > 1 && !dev_ids) return fail(MQVS_E_INVALID, "dev_ids is NULL with n_dev=%d", n_dev);
    if (dim < 1 || dim > MQVS_DIM_PAD)
        return fail(MQVS_E_INVALID, "dim must be in [1, %d] for this build of libmqvs (MQVS_DIM_PAD; wider rows: make WIDTH=<multiple "
                                    "of 128 >= dim>), got %d", MQVS_DIM_PAD, dim);
    if (store_dtype != MQVS_F32 && store_dtype != MQVS_BF16) return fail(MQVS_E_INVALID, "unknown store_dtype %d", store_dtype);
    if (n_dev > 1) return multi_create(n_dev, dev_ids, dim, store_dtype, capacity_rows, out);
    return create_single(dev_ids ? dev_ids[0] : 0, dim, store_dtype, capacity_rows, out);
}

}  // extern "C"

namespace mqvs {
namespace api {

int create_single(int dev, int dim, int store_dtype, uint64_t capacity_rows, mqvs_store** out) {
    *out = nullptr;
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
    s->shadow_on = store_dtype == MQVS_F32 && getenv("MQVS_NO_SHADOW") == nullptr;
    s->sm_count = prop.multiProcessorCount;
    s->coop_ok = prop.cooperativeLaunch != 0;
    // MQVS_TIMING=1: four CUDA events per search instead of the two around the scan (stats.last_search_ms then covers the
    // first kernel to the last); MQVS_TIMING=0: none (stats.scan_ms / last_search_ms stay 0)
    if (const char* t = getenv("MQVS_TIMING")) {
        s->timing_level = atoi(t) == 1 ? 1 : (atoi(t) == 0 ? 0 : 2);
        s->timing_from_env = true;
    }
    s->timing = s->timing_level != 0;
    auto bail = [&](int rc) {
        destroy_single(s);
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
    if (s->host_out.ensure(4096) != MQVS_OK) return bail(MQVS_E_OOM);
    memset(s->host_out.p, 0, s->host_out.bytes);
    if (cudaMalloc(&s->fix_dev, (kMaxFixRound + 1) * sizeof(uint32_t)) != cudaSuccess ||
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

void destroy_single(mqvs_store* s) {
    if (!s) return;
    cudaSetDevice(s->dev);
    if (s->st) cudaStreamSynchronize(s->st);
    if (s->copy_st) cudaStreamSynchronize(s->copy_st);
    for (auto& c : s->columns) {
        c.values.release();
        c.present.release();
        c.offsets.release();
    }
    for (DevBuf* b : {&s->f_table, &s->f_prog, &s->f_pool, &s->f_live, &s->f_mask}) b->release();
    if (s->f_count_dev) cudaFree(s->f_count_dev);
    if (s->f_count_host) cudaFreeHost(s->f_count_host);
    for (DevBuf* b : {&s->norm_bounds, &s->stage[0], &s->stage[1], &s->qraw, &s->qn, &s->qnorm, &s->cand, &s->out_s, &s->out_i,
                      &s->maskbuf, &s->idbuf, &s->gatherbuf, &s->tc_qop, &s->tc_small, &s->tc_cnt, &s->tc_lists, &s->tc_keys,
                      &s->tc_fixcand, &s->tc_cmax})
        b->release();
    for (PinBuf* b : {&s->bounce[0], &s->bounce[1], &s->host_out, &s->flags_map}) b->release();
    if (s->fix_dev) cudaFree(s->fix_dev);
    if (s->fix_counters) cudaFree(s->fix_counters);
    if (s->rows) cudaFree(s->rows);
    if (s->shadow) cudaFree(s->shadow);
    if (s->shadow_err) cudaFree(s->shadow_err);
    if (s->norms) cudaFree(s->norms);
    if (s->bad_dev) cudaFree(s->bad_dev);
    if (s->ticket_dev) cudaFree(s->ticket_dev);
    if (s->bad_host) cudaFreeHost(s->bad_host);
    for (auto& set : s->evr)
        for (auto& e : set.e)
            if (e) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
        if (s->ev_h2d[i]) cudaEventDestroy(s->ev_h2d[i]);
        if (s->ev_k1[i]) cudaEventDestroy(s->ev_k1[i]);
    }
    if (s->copy_st) cudaStreamDestroy(s->copy_st);
    if (s->st) cudaStreamDestroy(s->st);
    cudaGetLastError();
    delete s;
}

// ---- ingest
namespace {

// pageable host memory -> pinned bounce buffer with a few host threads (one core copies ~10 GB/s, PCIe Gen5 takes 50+)
void parallel_memcpy(void* dst, const void* src, size_t bytes) {
    const size_t kPerThread = 4u << 20;
    int nt = (int)std::min<size_t>(4, bytes / kPerThread);
    if (nt <= 1) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    const size_t part = (bytes / nt + 63) & ~(size_t)63;
    for (int t = 1; t < nt; ++t) {
        const size_t off = (size_t)t * part, len = off >= bytes ? 0 : std::min(part, bytes - off);
        if (len) th.emplace_back([=] { memcpy((char*)dst + off, (const char*)src + off, len); });
    }
    memcpy(dst, src, std::min(part, bytes));
    for (auto& t : th) t.join();
}

bool is_pinned_host(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

}  // namespace

// Host rows -> K1 through a two-slot pipeline: the H2D copy of chunk i + 1 (copy stream) overlaps K1 of chunk i (the
// store's stream); pageable sources are bounced through pinned buffers by a few host threads while the previous chunk is
// in flight, pinned sources (cudaHostAlloc / cudaHostRegister by the caller) are copied from where they lie.
static int ingest_host(mqvs_store* s, const AddSeg* segs, size_t n_segs, char* dst, float* norms_dst, unsigned int* dev_word,
                       uint64_t* launches) {
    const uint64_t chunk = 8192;  // rows per slot: 25 MB at 768 columns
    const size_t slot_bytes = (size_t)chunk * s->dim * sizeof(float);
    int rc;
    if (!s->copy_st) {
        CU(cudaStreamCreateWithFlags(&s->copy_st, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CU(cudaEventCreateWithFlags(&s->ev_h2d[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&s->ev_k1[i], cudaEventDisableTiming));
        }
    }
    uint64_t total = 0;
    for (size_t i = 0; i < n_segs; ++i) total += segs[i].n;
    const size_t need = std::min<size_t>(slot_bytes, (size_t)total * s->dim * sizeof(float));
    for (int i = 0; i < 2; ++i)
        if ((rc = s->stage[i].ensure(need)) != MQVS_OK) return rc;
    bool used[2] = {false, false};
    uint64_t done = 0, c = 0;
    for (size_t g = 0; g < n_segs; ++g) {
        const bool pinned = is_pinned_host(segs[g].src);
        for (uint64_t r = 0; r < segs[g].n; r += chunk, ++c) {
            const int slot = (int)(c & 1);
            const uint64_t m = std::min(chunk, segs[g].n - r);
            const size_t bytes = (size_t)m * s->dim * sizeof(float);
            if (used[slot]) CU(cudaEventSynchronize(s->ev_k1[slot]));  // K1 has consumed this slot (so its H2D is long done)
            const void* src = segs[g].src + r * s->dim;
            if (!pinned) {
                if ((rc = s->bounce[slot].ensure(need)) != MQVS_OK) return rc;
                parallel_memcpy(s->bounce[slot].p, src, bytes);
                src = s->bounce[slot].p;
            }
            CU(cudaMemcpyAsync(s->stage[slot].p, src, bytes, cudaMemcpyHostToDevice, s->copy_st));
            CU(cudaEventRecord(s->ev_h2d[slot], s->copy_st));
            CU(cudaStreamWaitEvent(s->st, s->ev_h2d[slot], 0));
            launch_normalize((const float*)s->stage[slot].p, m, s->dim, s->dtype, dst + done * s->row_bytes(), s->bad_dev, s->st,
                             norms_dst ? norms_dst + done : nullptr, dev_word);
            CU(cudaEventRecord(s->ev_k1[slot], s->st));
            used[slot] = true;
            done += m;
            ++*launches;
        }
    }
    return MQVS_OK;
}

int add_stage(mqvs_store* s, const AddSeg* segs, size_t n_segs, int x_is_device, bool synth, uint64_t seed, AddReport* rep) {
    *rep = AddReport();
    uint64_t total = 0;
    for (size_t i = 0; i < n_segs; ++i) total += segs[i].n;
    rep->rows = total;
    if (total == 0) return MQVS_OK;
    int rc = grow(s, s->n + total);
    if (rc != MQVS_OK) return rc;
    char* dst = (char*)s->rows + s->n * s->row_bytes();
    float* norms_dst = s->keep_norms ? s->norms + s->n : nullptr;
    unsigned int* dev_word = (unsigned int*)(s->bad_dev + 2);
    CU(cudaMemsetAsync(s->bad_dev, 0, sizeof(unsigned long long), s->st));
    CU(cudaMemsetAsync(s->bad_dev + 2, 0, sizeof(unsigned long long), s->st));
    if (synth || x_is_device) {
        uint64_t done = 0;
        for (size_t i = 0; i < n_segs; ++i) {
            if (segs[i].n == 0) continue;
            if (synth)
                launch_synth_normalize(seed, segs[i].first_index, segs[i].n, s->dim, s->dtype, dst + done * s->row_bytes(), s->bad_dev,
                                       s->st, norms_dst ? norms_dst + done : nullptr, dev_word);
            else
                launch_normalize(segs[i].src, segs[i].n, s->dim, s->dtype, dst + done * s->row_bytes(), s->bad_dev, s->st,
                                 norms_dst ? norms_dst + done : nullptr, dev_word);
            done += segs[i].n;
            ++rep->launches;
        }
    } else if ((rc = ingest_host(s, segs, n_segs, dst, norms_dst, dev_word, &rep->launches)) != MQVS_OK) {
        cudaStreamSynchronize(s->st);
        return rc;
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(s->bad_host, s->bad_dev, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    s->stats.kernel_launches += rep->launches;
    rep->bad = *s->bad_host;
    const uint32_t bits = (uint32_t)s->bad_host[2], bits_max = (uint32_t)(s->bad_host[2] >> 32);
    memcpy(&rep->norm_dev, &bits, 4);
    memcpy(&rep->norm_max, &bits_max, 4);
    return MQVS_OK;
}

void add_commit(mqvs_store* s, const AddReport& rep) {
    s->norm_dev = std::max(s->norm_dev, rep.norm_dev);
    s->norm_max = std::max(s->norm_max, rep.norm_max);
    s->n += rep.rows;
    s->stats.rows = s->n;
}

}  // namespace api
}  // namespace mqvs

extern "C" {

MQVS_API void mqvs_destroy(mqvs_t* s) {
    if (!s) return;
    if (s->multi) return multi_destroy(s);
    destroy_single(s);
}

MQVS_API int mqvs_count(mqvs_t* s, uint64_t* out) {
    if (!s || !out) return fail(MQVS_E_INVALID, "NULL argument");
    if (s->multi) return multi_count(s, out);
    std::lock_guard<std::mutex> lk(s->mu);
    *out = s->n;
    return MQVS_OK;
}

MQVS_API int mqvs_reserve(mqvs_t* s, uint64_t capacity_rows) {
    if (!s) return fail(MQVS_E_INVALID, "NULL handle");
    if (s->multi) return multi_reserve(s, capacity_rows);
    std::lock_guard<std::mutex> lk(s->mu);
    CU(cudaSetDevice(s->dev));
    return grow(s, capacity_rows);
}

static int add_common(mqvs_t* s, const void* x, int x_is_device, bool synth, uint64_t seed, uint64_t first_index,
                      uint64_t n, uint64_t* first_row_out) {
    if (s->multi) return multi_add(s, x, x_is_device, synth, seed, first_index, n, first_row_out);
    std::lock_guard<std::mutex> lk(s->mu);
    CU(cudaSetDevice(s->dev));
    if (first_row_out) *first_row_out = s->n;
    if (n == 0) return MQVS_OK;
    AddSeg seg{(const float*)x, n, first_index};
    AddReport rep;
    int rc = add_stage(s, &seg, 1, x_is_device, synth, seed, &rep);
    if (rc != MQVS_OK) return rc;
    if (rep.bad)
        return fail(MQVS_E_INVALID, "%llu of %llu embeddings have a non-finite norm; nothing was added",
                    (unsigned long long)rep.bad, (unsigned long long)n);
    add_commit(s, rep);
    return MQVS_OK;
}

MQVS_API int mqvs_add(mqvs_t* s, const void* x, int x_dtype, int x_is_device, uint64_t n, uint64_t* first_row_out) {
    MQVS_RANGE("mqvs_add");
    if (!s) return fail(MQVS_E_INVALID, "NULL handle");
    if (n && !x) return fail(MQVS_E_INVALID, "x is NULL");
    if (x_dtype != MQVS_F32) return fail(MQVS_E_INVALID, "embeddings must be fp32 (x_dtype=%d)", x_dtype);
    return add_common(s, x, x_is_device, false, 0, 0, n, first_row_out);
}

MQVS_API int mqvs_add_synthetic(mqvs_t* s, uint64_t seed, uint64_t first_index, uint64_t n, uint64_t* first_row_out) {
    MQVS_RANGE("mqvs_add_synthetic");
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
    MQVS_RANGE("mqvs_search");
    int rc = check_search_args(s, q, B, k, out_scores, out_ids);
    if (rc != MQVS_OK) return rc;
    if (B == 0) return MQVS_OK;
    if (k > 128 && B > kWideChunk) {
        // wide shortlists (k > 128) live in the single-launch pipeline, which takes up to 256 queries: larger batches go
        // through it 256 at a time (one corpus pass per 256 queries; the SIMT scan this replaces took 1-4 per pass)
        for (int b0 = 0; b0 < B; b0 += kWideChunk) {
            rc = mqvs_search(s, q + (size_t)b0 * s->dim, q_is_device, std::min(kWideChunk, B - b0), k, mask, mask_is_device,
                             out_scores + (size_t)b0 * k, out_ids + (size_t)b0 * k, out_is_device);
            if (rc != MQVS_OK) return rc;
        }
        return MQVS_OK;
    }
    if (s->multi)
        return multi_search(s, q, q_is_device, B, k, mask, mask_is_device, 0, out_scores, out_ids, out_is_device, nullptr, true);
    std::lock_guard<std::mutex> lk(s->mu);
    CU(cudaSetDevice(s->dev));
    return search_locked(s, q, q_is_device, B, k, mask, mask_is_device, out_scores, out_ids, out_is_device);
}

}  // extern "C"

namespace mqvs {
namespace api {

// After the caller has synchronised the stream of a report_to_host search: repair on the device what the tensor pipeline
// could not prove exact (band pass + exact-scan rounds sized by the flag count), then whatever even that left flagged
// (band overflow and more flagged queries than the rounds took) on the exact scan, one query at a time.
int finish_search(mqvs_store* s, int B, int k, const uint32_t* mask_dev, int64_t id_base, float* os, int64_t* oi,
                  cudaStream_t st, bool* changed) {
    *changed = false;
    if (s->stats.last_path != 1) return MQVS_OK;
    volatile uint32_t* fl = (volatile uint32_t*)s->flags_map.p;
    if (fl == nullptr || fl[0] == 0) return MQVS_OK;
    const int flagged = (int)fl[0];
    enqueue_tensor_repair(s, B, k, mask_dev, id_base, os, oi, st, (flagged + 127) / 128,
                          (flagged + s->max_device_fix - 1) / s->max_device_fix);
    CU(cudaStreamSynchronize(st));
    int rc;
    for (int b = 0; b < B; ++b) {
        if (!fl[2 + b]) continue;  // flags of device-repaired queries were cleared
        uint32_t passes = 0;
        if ((rc = enqueue_simt(s, b, 1, k, mask_dev, id_base, os, oi, st, &passes)) != MQVS_OK) return rc;
        s->stats.kernel_launches += passes + 1;
        s->host_fallbacks += 1;
        fl[2 + b] = 0;
    }
    fl[0] = 0;
    CU(cudaStreamSynchronize(st));
    *changed = true;
    return MQVS_OK;
}

}  // namespace api
}  // namespace mqvs

static int search_locked(mqvs_t* s, const float* q, int q_is_device, int B, int k, const uint32_t* mask, int mask_is_device,
                         float* out_scores, int64_t* out_ids, int out_is_device) {
    int rc;
    const float* q_dev = q;
    // latency path (B <= 4, k <= 16, the reference's own regime): one launch that also reports completion + non-finite status;
    // a host caller's queries travel inside that launch's parameter block, so the call is ONE launch and one stream sync
    const bool zero_copy = !out_is_device && !use_tensor_path(s, B, k) && s->path != 3 && s->metric == MQVS_METRIC_COSINE &&
                           mqvs::scan_fused_ok(B, k);
    static const bool no_inline = getenv("MQVS_NO_INLINE_QUERIES") != nullptr;  // developer knob: measure what it buys
    const bool inline_q = zero_copy && !q_is_device && !no_inline;
    if (inline_q) {
        q_dev = nullptr;
    } else if (!q_is_device) {
        // Pinned caller memory, up to 256 queries (768 KB): the first kernel of the search (query normalisation) reads the raw
        // queries straight from the host over PCIe -- no copy-engine launch in front of it.  Larger batches and pageable
        // memory go through the copy engine.
        static const bool no_zero_copy = getenv("MQVS_NO_ZERO_COPY_QUERIES") != nullptr;  // developer knob: measure what it buys
        void* mapped = nullptr;
        if (!no_zero_copy && B <= 256 && is_pinned_host(q) && cudaHostGetDevicePointer(&mapped, (void*)q, 0) == cudaSuccess && mapped) {
            q_dev = (const float*)mapped;
        } else {
            cudaGetLastError();
            if ((rc = s->qraw.ensure((size_t)B * s->dim * sizeof(float))) != MQVS_OK) return rc;
            CU(cudaMemcpyAsync(s->qraw.p, q, (size_t)B * s->dim * sizeof(float), cudaMemcpyHostToDevice, s->st));
            q_dev = (const float*)s->qraw.p;
        }
    }
    const uint32_t* mask_dev = mask;
    if (mask && !mask_is_device) {
        size_t bytes = (size_t)((s->n + 31) / 32) * sizeof(uint32_t);
        if ((rc = s->maskbuf.ensure(bytes)) != MQVS_OK) return rc;
        CU(cudaMemcpyAsync(s->maskbuf.p, mask, bytes, cudaMemcpyHostToDevice, s->st));
        mask_dev = (const uint32_t*)s->maskbuf.p;
    }
    const size_t n_out = (size_t)B * k;
    float* os = out_scores;
    int64_t* oi = out_ids;
    unsigned int* status = nullptr;
    // Host outputs: every path writes results (and the latency kernel its status) straight into mapped pinned memory --
    // the call is one H2D copy of the queries, the launches, one stream sync; nothing is copied back.
    if (!out_is_device) {
        const size_t need = n_out * 12 + 64;
        if (need > s->host_out.bytes) {
            if ((rc = s->host_out.ensure(need * 2)) != MQVS_OK) return rc;
            memset(s->host_out.p, 0, s->host_out.bytes);
        }
        oi = (int64_t*)s->host_out.p;
        os = (float*)((char*)s->host_out.p + n_out * sizeof(int64_t));
        status = (unsigned int*)((char*)s->host_out.p + n_out * 12 + ((8 - (n_out * 12) % 8) % 8));
        status[0] = 0;
    }
    if (zero_copy) {
        const bool host_timed = s->timing && !s->timing_from_env;  // (enqueue_search records no events for this call then)
        const auto t_launch = std::chrono::steady_clock::now();
        if ((rc = enqueue_search(s, q_dev, B, k, mask_dev, 0, os, oi, s->st, status, false, false, inline_q ? q : nullptr)) != MQVS_OK)
            return rc;
        // The kernel's last store is the completion word in mapped host memory (after a system-wide fence behind the results):
        // watching that word returns as soon as it lands, without the driver's stream-sync round trip behind the kernel and
        // the timing event that follows it.  Bounded: a scan that takes longer (or a failed launch) falls through to the sync.
        static const bool no_spin = getenv("MQVS_NO_SPIN_WAIT") != nullptr;  // developer knob: measure what it buys
        if (!no_spin) {
            volatile unsigned int* done = status;
            const auto t0 = std::chrono::steady_clock::now();
            for (unsigned spins = 1; done[0] != s->epoch; ++spins) {
                if ((spins & 1023u) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(2)) break;
#if defined(__x86_64__) || defined(__i386__)
                __builtin_ia32_pause();
#endif
            }
            std::atomic_thread_fence(std::memory_order_acquire);
        }
        if (((volatile unsigned int*)status)[0] != s->epoch) CU(cudaStreamSynchronize(s->st));
        if (((volatile unsigned int*)status)[0] != s->epoch) return fail(MQVS_E_CUDA, "latency path: the kernel did not report completion");
        if (host_timed) {  // launch call -> completion word seen: what the caller waited for the device
            const float ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_launch).count();
            s->stats.last_search_ms = s->stats.last_scan_ms = ms;
            s->stats.scan_ms += ms;
        }
        memcpy(out_scores, os, n_out * sizeof(float));
        memcpy(out_ids, oi, n_out * sizeof(int64_t));
        if (status[1]) return fail(MQVS_E_INVALID, "queries with a non-finite norm in a batch of %d", B);
        return MQVS_OK;
    }
    if ((rc = enqueue_search(s, q_dev, B, k, mask_dev, 0, os, oi, s->st, nullptr, /*defer_repair=*/true, /*report_to_host=*/true)) != MQVS_OK)
        return rc;
    const bool tensor = s->stats.last_path == 1;
    const bool epoch_word = tensor && s->last_fused;  // single-launch tensor pipeline: non-finite epoch in flags_map[1]
    if (!epoch_word) CU(cudaMemcpyAsync(s->bad_host, s->bad_dev, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->st));
    CU(cudaStreamSynchronize(s->st));
    const bool bad = epoch_word ? ((volatile uint32_t*)s->flags_map.p)[1] == s->epoch
                                : (s->last_fused ? (unsigned int)s->bad_host[1] == s->epoch : s->bad_host[0] != 0);
    if (bad) return fail(MQVS_E_INVALID, "queries with a non-finite norm in a batch of %d", B);
    bool changed = false;
    if ((rc = finish_search(s, B, k, mask_dev, 0, os, oi, s->st, &changed)) != MQVS_OK) return rc;
    if (!out_is_device) {
        memcpy(out_scores, os, n_out * sizeof(float));
        memcpy(out_ids, oi, n_out * sizeof(int64_t));
    }
    settle_events(s);
    mqvs::tc::trace_dump();
    return MQVS_OK;
}

extern "C" {

MQVS_API int mqvs_search_async(mqvs_t* s, const float* q_dev, int B, int k, const uint32_t* mask_dev, int64_t id_base,
                      float* out_scores_dev, int64_t* out_ids_dev, void* cuda_stream) {
    MQVS_RANGE("mqvs_search_async");
    int rc = check_search_args(s, q_dev, B, k, out_scores_dev, out_ids_dev);
    if (rc != MQVS_OK) return rc;
    if (B == 0) return MQVS_OK;
    if (k > 128 && B > kWideChunk) {
        for (int b0 = 0; b0 < B; b0 += kWideChunk) {
            rc = mqvs_search_async(s, q_dev + (size_t)b0 * s->dim, std::min(kWideChunk, B - b0), k, mask_dev, id_base,
                                   out_scores_dev + (size_t)b0 * k, out_ids_dev + (size_t)b0 * k, cuda_stream);
            if (rc != MQVS_OK) return rc;
        }
        return MQVS_OK;
    }
    if (s->multi)
        return multi_search(s, q_dev, 1, B, k, mask_dev, 1, id_base, out_scores_dev, out_ids_dev, 1, (cudaStream_t)cuda_stream, false);
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
    MQVS_RANGE("mqvs_get_rows");
    if (!s) return fail(MQVS_E_INVALID, "NULL handle");
    if (n == 0) return MQVS_OK;
    if (!row_ids || !out_host) return fail(MQVS_E_INVALID, "NULL buffer");
    if (s->multi) return multi_get_rows(s, row_ids, n, out_host);
    std::lock_guard<std::mutex> lk(s->mu);
    return get_rows_single(s, row_ids, n, out_host);
}

}  // extern "C"

int mqvs::api::get_rows_single(mqvs_store* s, const int64_t* row_ids, uint64_t n, float* out_host) {
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
    ok = ok && fflush(f) == 0 && fsync(fileno(f)) == 0;
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
// flush a stdio stream all the way to the device: the orderings the row file promises must survive a power loss too
bool flush_to_disk(FILE* f) { return fflush(f) == 0 && fsync(fileno(f)) == 0; }
}  // namespace

extern "C" {

MQVS_API int mqvs_save(mqvs_t* s, const char* path) {
    MQVS_RANGE("mqvs_save");
    if (!s || !path) return fail(MQVS_E_INVALID, "NULL argument");
    if (s->multi) return multi_save(s, path, 0);
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
    ok = ok && flush_to_disk(f);
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
    MQVS_RANGE("mqvs_save_append");
    if (!s || !path) return fail(MQVS_E_INVALID, "NULL argument");
    if (first_row == 0) return mqvs_save(s, path);
    if (s->multi) return multi_save(s, path, first_row);
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
    // the row count in the header moves last, after the rows are ON DISK (fsync): a crash or a power loss before this
    // leaves the old, consistent file (plus ignored bytes)
    h.rows = s->n;
    ok = ok && flush_to_disk(f) && fseek(f, 0, SEEK_SET) == 0 && fwrite(&h, sizeof(h), 1, f) == 1 && flush_to_disk(f);
    ok = (fclose(f) == 0) && ok;
    if (!ok) return fail(MQVS_E_IO, "appending to %s failed", path);
    return MQVS_OK;
}

MQVS_API int mqvs_load(mqvs_t* s, const char* path) {
    MQVS_RANGE("mqvs_load");
    if (!s || !path) return fail(MQVS_E_INVALID, "NULL argument");
    if (s->multi) return multi_load(s, path);
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
    s->shadow_rows = 0;  // rows were rewritten in place
    s->stats.rows = s->n;
    return MQVS_OK;
}

MQVS_API int mqvs_column_set_tokens(mqvs_t* s, int col_id, const uint64_t* offsets_host, const int64_t* tokens_host, uint64_t n) {
    MQVS_RANGE("mqvs_column_set_tokens");
    if (!s) return fail(MQVS_E_INVALID, "NULL handle");
    if (col_id < 0 || col_id >= mqvs::kFilterMaxCols) return fail(MQVS_E_INVALID, "column id must be in [0, %d)", mqvs::kFilterMaxCols);
    if (!offsets_host) return fail(MQVS_E_INVALID, "offsets is NULL");
    if (offsets_host[0] != 0) return fail(MQVS_E_INVALID, "offsets must start at 0");
    for (uint64_t r = 0; r < n; ++r)
        if (offsets_host[r + 1] < offsets_host[r]) return fail(MQVS_E_INVALID, "offsets must not decrease (row %llu)", (unsigned long long)r);
    if (offsets_host[n] && !tokens_host) return fail(MQVS_E_INVALID, "tokens is NULL");
    if (s->multi) return multi_column_set_tokens(s, col_id, offsets_host, tokens_host, n);
    std::lock_guard<std::mutex> lk(s->mu);
    return column_set_tokens_single(s, col_id, offsets_host, tokens_host, n);
}

MQVS_API int mqvs_column_set_text(mqvs_t* s, int col_id, const uint64_t* offsets_host, const uint8_t* bytes_host,
                                  const uint8_t* present_host, uint64_t n) {
    MQVS_RANGE("mqvs_column_set_text");
    if (!s) return fail(MQVS_E_INVALID, "NULL handle");
    if (col_id < 0 || col_id >= mqvs::kFilterMaxCols) return fail(MQVS_E_INVALID, "column id must be in [0, %d)", mqvs::kFilterMaxCols);
    if (!offsets_host) return fail(MQVS_E_INVALID, "offsets is NULL");
    if (offsets_host[0] != 0) return fail(MQVS_E_INVALID, "offsets must start at 0");
    for (uint64_t r = 0; r < n; ++r)
        if (offsets_host[r + 1] < offsets_host[r]) return fail(MQVS_E_INVALID, "offsets must not decrease (row %llu)", (unsigned long long)r);
    if (offsets_host[n] && !bytes_host) return fail(MQVS_E_INVALID, "bytes is NULL");
    if (s->multi) return multi_column_set_text(s, col_id, offsets_host, bytes_host, present_host, n);
    std::lock_guard<std::mutex> lk(s->mu);
    return column_set_text_single(s, col_id, offsets_host, bytes_host, present_host, n);
}

MQVS_API int mqvs_compact(mqvs_t* s, const uint32_t* keep_words_host, uint64_t* new_count_out) {
    MQVS_RANGE("mqvs_compact");
    if (!s || !keep_words_host || !new_count_out) return fail(MQVS_E_INVALID, "NULL argument");
    if (s->multi) return multi_compact(s, keep_words_host, new_count_out);
    std::lock_guard<std::mutex> lk(s->mu);
    CU(cudaSetDevice(s->dev));
    CU(cudaStreamSynchronize(s->st));
    std::vector<uint32_t> keep;
    keep.reserve((size_t)s->n);
    for (uint64_t r = 0; r < s->n; ++r)
        if ((keep_words_host[r >> 5] >> (r & 31)) & 1u) keep.push_back((uint32_t)r);
    const uint64_t m = keep.size();
    if (m < s->n) {
        // survivors move down in ascending order, chunk by chunk through a bounce buffer: the destination of a chunk never
        // reaches beyond its first source row, and every earlier source row has already been consumed
        const uint64_t chunk = 16384;
        DevBuf ids, tmp;
        int rc;
        if ((rc = ids.ensure(std::min<uint64_t>(chunk, std::max<uint64_t>(m, 1)) * sizeof(uint32_t))) != MQVS_OK) return rc;
        if ((rc = tmp.ensure(std::min<uint64_t>(chunk, std::max<uint64_t>(m, 1)) * s->row_bytes())) != MQVS_OK) {
            ids.release();
            return rc;
        }
        cudaError_t e = cudaSuccess;
        for (uint64_t done = 0; done < m && e == cudaSuccess; done += chunk) {
            const uint64_t c = std::min(chunk, m - done);
            if (keep[done + c - 1] == done + c - 1) continue;  // nothing dead up to here: rows are already in place
            e = cudaMemcpyAsync(ids.p, keep.data() + done, c * sizeof(uint32_t), cudaMemcpyHostToDevice, s->st);
            if (e != cudaSuccess) break;
            mqvs::launch_gather_raw_rows(s->rows, s->row_bytes(), (const uint32_t*)ids.p, c, tmp.p, s->st);
            e = cudaMemcpyAsync((char*)s->rows + done * s->row_bytes(), tmp.p, c * s->row_bytes(), cudaMemcpyDeviceToDevice, s->st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s->st);  // `ids` is reused
        }
        if (e == cudaSuccess && s->keep_norms && m) {
            std::vector<float> nh((size_t)s->n);
            e = cudaMemcpy(nh.data(), s->norms, s->n * sizeof(float), cudaMemcpyDeviceToHost);
            for (uint64_t i = 0; i < m; ++i) nh[i] = nh[keep[i]];
            if (e == cudaSuccess) e = cudaMemcpy(s->norms, nh.data(), m * sizeof(float), cudaMemcpyHostToDevice);
        }
        ids.release();
        tmp.release();
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(MQVS_E_CUDA, "compaction failed: %s (the store is in an undefined state)", cudaGetErrorString(e));
        }
        s->stats.kernel_launches += (m + chunk - 1) / chunk;
        s->n = m;
        s->stats.rows = m;
        s->bounds_rows = 0;
        s->shadow_rows = 0;  // rows moved: the scoring shadow is rebuilt before the next tensor search
        for (auto& c : s->columns) c.used = false;  // row numbers changed under them
        s->f_table_dirty = true;
    }
    *new_count_out = s->n;
    return MQVS_OK;
}

MQVS_API int mqvs_column_set(mqvs_t* s, int col_id, int col_type, const void* values_host, const uint8_t* present_host,
                             uint64_t n) {
    MQVS_RANGE("mqvs_column_set");
    if (!s) return fail(MQVS_E_INVALID, "NULL handle");
    if (col_id < 0 || col_id >= mqvs::kFilterMaxCols) return fail(MQVS_E_INVALID, "column id must be in [0, %d)", mqvs::kFilterMaxCols);
    if (col_type != MQVS_COL_I64 && col_type != MQVS_COL_F64) return fail(MQVS_E_INVALID, "unknown column type %d", col_type);
    if (n && !values_host) return fail(MQVS_E_INVALID, "values is NULL");
    if (s->multi) return multi_column_set(s, col_id, col_type, values_host, present_host, n);
    std::lock_guard<std::mutex> lk(s->mu);
    return column_set_single(s, col_id, col_type, values_host, present_host, n);
}

}  // extern "C"

int mqvs::api::column_set_tokens_single(mqvs_store* s, int col_id, const uint64_t* offsets_host, const int64_t* tokens_host, uint64_t n) {
    CU(cudaSetDevice(s->dev));
    auto& c = s->columns[col_id];
    int rc;
    const uint64_t nnz = offsets_host[n];
    if ((rc = c.offsets.ensure((n + 1) * 8)) != MQVS_OK) return rc;
    if ((rc = c.values.ensure(std::max<uint64_t>(nnz, 1) * 8)) != MQVS_OK) return rc;
    CU(cudaMemcpyAsync(c.offsets.p, offsets_host, (n + 1) * 8, cudaMemcpyHostToDevice, s->st));
    if (nnz) CU(cudaMemcpyAsync(c.values.p, tokens_host, nnz * 8, cudaMemcpyHostToDevice, s->st));
    CU(cudaStreamSynchronize(s->st));
    c.n = n;
    c.type = MQVS_COL_TOKENS;
    c.used = true;
    c.has_present = false;
    s->f_table_dirty = true;
    return MQVS_OK;
}

int mqvs::api::column_set_text_single(mqvs_store* s, int col_id, const uint64_t* offsets_host, const uint8_t* bytes_host,
                                      const uint8_t* present_host, uint64_t n) {
    CU(cudaSetDevice(s->dev));
    auto& c = s->columns[col_id];
    int rc;
    const uint64_t nbytes = offsets_host[n];
    if ((rc = c.offsets.ensure((n + 1) * 8)) != MQVS_OK) return rc;
    if ((rc = c.values.ensure(std::max<uint64_t>(nbytes, 8))) != MQVS_OK) return rc;
    CU(cudaMemcpyAsync(c.offsets.p, offsets_host, (n + 1) * 8, cudaMemcpyHostToDevice, s->st));
    if (nbytes) CU(cudaMemcpyAsync(c.values.p, bytes_host, nbytes, cudaMemcpyHostToDevice, s->st));
    c.has_present = present_host != nullptr;
    std::vector<uint32_t> bits;
    if (present_host) {
        bits.assign(std::max<uint64_t>((n + 31) / 32, 1), 0u);
        for (uint64_t r = 0; r < n; ++r)
            if (present_host[r]) bits[r >> 5] |= 1u << (r & 31);
        if ((rc = c.present.ensure(bits.size() * 4)) != MQVS_OK) return rc;
        CU(cudaMemcpyAsync(c.present.p, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice, s->st));
    }
    CU(cudaStreamSynchronize(s->st));
    c.n = n;
    c.type = MQVS_COL_TEXT;
    c.used = true;
    s->f_table_dirty = true;
    return MQVS_OK;
}

int mqvs::api::column_set_single(mqvs_store* s, int col_id, int col_type, const void* values_host, const uint8_t* present_host,
                                 uint64_t n) {
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

extern "C" {

MQVS_API int mqvs_filter_build(mqvs_t* s, const mqvs_filter_op* prog, int n_ops, const void* pool, uint64_t pool_len,
                               const uint32_t* live_words_host, const uint32_t** mask_dev_out, uint64_t* allowed_out) {
    MQVS_RANGE("mqvs_filter_build");
    using namespace mqvs;
    if (!s || !mask_dev_out || !allowed_out) return fail(MQVS_E_INVALID, "NULL argument");
    if (n_ops < 0 || n_ops > kFilterMaxOps) return fail(MQVS_E_INVALID, "a filter program holds at most %d operations", kFilterMaxOps);
    if (n_ops && !prog) return fail(MQVS_E_INVALID, "prog is NULL");
    if (s->multi) return multi_filter_build(s, prog, n_ops, pool, pool_len, live_words_host, mask_dev_out, allowed_out);
    std::lock_guard<std::mutex> lk(s->mu);
    return filter_build_single(s, prog, n_ops, pool, pool_len, live_words_host, mask_dev_out, allowed_out);
}

}  // extern "C"

int mqvs::api::filter_build_single(mqvs_store* s, const mqvs_filter_op* prog, int n_ops, const void* pool, uint64_t pool_len,
                                   const uint32_t* live_words_host, const uint32_t** mask_dev_out, uint64_t* allowed_out) {
    using namespace mqvs;
    CU(cudaSetDevice(s->dev));
    // validate: known columns, sets inside the pool, a well-formed postfix expression
    int depth = 0;
    for (int i = 0; i < n_ops; ++i) {
        const mqvs_filter_op& o = prog[i];
        if (o.op == MQVS_OP_AND || o.op == MQVS_OP_OR) {
            if (depth < 2) return fail(MQVS_E_INVALID, "filter program: operator %d at %d lacks operands", o.op, i);
            --depth;
        } else if (o.op >= MQVS_OP_EQ && o.op <= MQVS_OP_TEXT_NOT_CONTAINS) {
            if (o.col < 0 || o.col >= kFilterMaxCols || !s->columns[o.col].used)
                return fail(MQVS_E_INVALID, "filter program: column %d was never set", o.col);
            const bool token_op = o.op == MQVS_OP_CONTAINS || o.op == MQVS_OP_NOT_CONTAINS;
            const bool text_op = o.op == MQVS_OP_TEXT_CONTAINS || o.op == MQVS_OP_TEXT_NOT_CONTAINS;
            if (token_op != (s->columns[o.col].type == MQVS_COL_TOKENS) || text_op != (s->columns[o.col].type == MQVS_COL_TEXT))
                return fail(MQVS_E_INVALID, "filter program: operation %d does not fit the type of column %d", o.op, o.col);
            if (text_op && (o.set_len < 1 || o.set_len > MQVS_MAX_NEEDLE || (uint64_t)o.set_off * 8 + o.set_len > pool_len * 8))
                return fail(MQVS_E_INVALID, "filter program: needle of operation %d is empty, longer than %d bytes or leaves the "
                                            "constant pool", i, MQVS_MAX_NEEDLE);
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
            table[i] = FilterColumn{c.values.p, c.has_present ? (const uint32_t*)c.present.p : nullptr, c.used ? c.n : 0, c.type,
                                    (const uint64_t*)c.offsets.p};
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

void mqvs::api::fill_stats(mqvs_store* s) {
    cudaSetDevice(s->dev);
    settle_events(s);
    s->stats.rows = s->n;
    unsigned long long c[4] = {0, 0, 0, 0};
    if (s->fix_counters && cudaMemcpy(c, s->fix_counters, sizeof(c), cudaMemcpyDeviceToHost) != cudaSuccess) cudaGetLastError();
    s->stats.band_repairs = c[2];
    s->stats.exact_fallbacks = c[0] + s->host_fallbacks;  // c[1]: queries the last device repair round of a search left flagged
    s->stats.unrepaired = c[1] > s->host_fallbacks ? c[1] - s->host_fallbacks : 0;
}

extern "C" {

MQVS_API int mqvs_get_stats(mqvs_t* s, mqvs_stats* out) {
    if (!s || !out) return fail(MQVS_E_INVALID, "NULL argument");
    if (s->multi) return multi_get_stats(s, out);
    std::lock_guard<std::mutex> lk(s->mu);
    fill_stats(s);
    *out = s->stats;
    return MQVS_OK;
}

MQVS_API int mqvs_set_metric(mqvs_t* s, int metric) {
    if (!s) return fail(MQVS_E_INVALID, "NULL handle");
    if (metric != MQVS_METRIC_COSINE && metric != MQVS_METRIC_L2 && metric != MQVS_METRIC_IP)
        return fail(MQVS_E_INVALID, "unknown metric %d", metric);
    if (s->multi) return multi_set_metric(s, metric);
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
    if (s->multi) return multi_norm_deviation(s, out);
    std::lock_guard<std::mutex> lk(s->mu);
    *out = s->norm_dev;
    return MQVS_OK;
}

MQVS_API int mqvs_get_norms(mqvs_t* s, uint64_t first_row, uint64_t n, float* out_host) {
    if (!s) return fail(MQVS_E_INVALID, "NULL handle");
    if (n == 0) return MQVS_OK;
    if (!out_host) return fail(MQVS_E_INVALID, "NULL buffer");
    if (s->multi) return multi_get_norms(s, first_row, n, out_host);
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

MQVS_API int mqvs_set_shadow(mqvs_t* s, int enable) {
    if (!s) return fail(MQVS_E_INVALID, "NULL handle");
    if (s->multi) return multi_set_shadow(s, enable);
    std::lock_guard<std::mutex> lk(s->mu);
    CU(cudaSetDevice(s->dev));
    s->shadow_on = enable != 0 && s->dtype == MQVS_F32;
    if (!s->shadow_on && s->shadow) {
        CU(cudaStreamSynchronize(s->st));
        cudaFree(s->shadow);
        s->shadow = nullptr;
        s->shadow_cap = s->shadow_rows = 0;
    }
    return MQVS_OK;
}

MQVS_API int mqvs_shadow_bytes(mqvs_t* s, uint64_t* bytes_out) {
    if (!s || !bytes_out) return fail(MQVS_E_INVALID, "NULL argument");
    if (s->multi) return multi_shadow_bytes(s, bytes_out);
    std::lock_guard<std::mutex> lk(s->mu);
    *bytes_out = s->shadow ? s->shadow_cap * (uint64_t)MQVS_DIM_PAD * 2 : 0;
    return MQVS_OK;
}

MQVS_API int mqvs_set_path(mqvs_t* s, int path) {
    if (!s) return fail(MQVS_E_INVALID, "NULL handle");
    if (path < 0 || path > 4) return fail(MQVS_E_INVALID, "path must be 0..4");
    if (s->multi) return multi_set_path(s, path);
    std::lock_guard<std::mutex> lk(s->mu);
    s->path = path;
    return MQVS_OK;
}

}  // extern "C"
