// api_multi.cu -- mqvs_create(n_dev > 1, ...): ONE handle in ONE host process over several row shards, one per device.
//
// Why: the reference holds its vector store as one module-global object in one process
// (/root/reference/src/medical_engine.py:52, injected at /root/reference/main.py:42-48, searched at
// /root/reference/src/agents/nodes.py:93), while BASELINE config 4 (100M x 768 bf16 = 153.6 GB) wants the HBM and the
// bandwidth of a whole 8-GPU box.  This layer gives that object: every entry point of include/mqvs.h works on a
// multi-device handle with GLOBAL row ids.
//
// Placement is block-cyclic (store.h:ShardMap): global block r / 4096 lives on part block % W, so appends keep the parts
// balanced (they differ by at most one block), packed masks / metadata columns split on block boundaries (128 mask words),
// and a part's local row order is its global row order, which keeps the tie rule (lower global id first) intact.
//
// A search = per part, on that part's own stream and device: queries in (H2D from one pinned copy, or a peer copy from the
// home device), the part's ordinary single-device search (api.cu:enqueue_search -- every pipeline, filter mask, space),
// its exact top-k written STRAIGHT INTO THE HOME DEVICE'S GATHER BUFFER over NVLink peer access (the rescoring tail's
// stores are the transfer; a peer copy where peer access is unavailable), an event.  Then on the home device: wait for
// the W events, one merge_candidates launch that maps local rows to global rows and keeps the best k of W*k.  No kernel
// waits for another kernel (nothing spins on a peer's flag), so several parts may share one GPU.
#include <thread>

#include "store.h"

namespace mqvs {

struct Multi {
    int W = 0;
    ShardMap map;
    std::vector<mqvs_store*> parts;
    std::vector<int> devs;
    int home = 0;                     // device of parts[0]: merges run there, device pointers of the ABI live there
    std::vector<char> direct;         // part p can store into home memory (same device, or peer access enabled)
    std::vector<cudaEvent_t> done;    // per part, on its device
    cudaEvent_t q_ev = nullptr;       // home: queries of an asynchronous search are ready
    DevBuf g_s, g_i;                  // home: gathered candidates [W][B * k]
    std::vector<DevBuf> loc_s, loc_i; // per part (its device): local outputs when it cannot store into home memory
    PinBuf q_pin, mask_pin, res_pin;  // host staging: queries, per-part mask words, final results (mapped)
    uint64_t n = 0;                   // rows stored (all parts)
    uint64_t searches = 0, queries = 0, merge_launches = 0;
    bool has_token = false;           // mqvs_filter_build ran: every part's f_mask holds the current filter
};

namespace api {

namespace {

struct DeviceGuard {  // restores the caller's current device
    int prev = 0;
    DeviceGuard() { cudaGetDevice(&prev); }
    ~DeviceGuard() { cudaSetDevice(prev); }
};

// global rows [first, first + n) as runs that stay inside one block: (part, local row, count, offset from `first`)
template <typename F>
void for_block_runs(const ShardMap& map, uint64_t first, uint64_t n, F&& f) {
    for (uint64_t r = first; r < first + n;) {
        const uint64_t end = std::min<uint64_t>(((r >> kShardShift) + 1) << kShardShift, first + n);
        f((int)map.part_of(r), map.local_of(r), end - r, r - first);
        r = end;
    }
}

// packed bit words of global rows [0, n) -> part p's local words (blocks are 128 words: plain copies)
void split_words(const ShardMap& map, const uint32_t* words, uint64_t n, int p, uint32_t* out) {
    const uint64_t total_words = (n + 31) / 32, wpb = kShardBlock / 32;
    for (uint64_t b = (uint64_t)p; b * wpb < total_words; b += (uint64_t)map.W) {
        const uint64_t cnt = std::min<uint64_t>(wpb, total_words - b * wpb);
        memcpy(out + (b / map.W) * wpb, words + b * wpb, cnt * 4);
    }
}

}  // namespace

int multi_create(int n_dev, const int* dev_ids, int dim, int store_dtype, uint64_t capacity_rows, mqvs_store** out) {
    *out = nullptr;
    DeviceGuard guard;
    mqvs_store* P = new (std::nothrow) mqvs_store();
    Multi* M = new (std::nothrow) Multi();
    if (!P || !M) {
        delete P;
        delete M;
        return mqvs_fail(MQVS_E_OOM, "host allocation failed");
    }
    P->multi = M;
    P->dim = dim;
    P->dtype = store_dtype;
    P->dev = dev_ids[0];
    M->W = n_dev;
    M->map.W = n_dev;
    M->home = dev_ids[0];
    M->devs.assign(dev_ids, dev_ids + n_dev);
    M->direct.assign(n_dev, 0);
    M->done.assign(n_dev, nullptr);
    M->loc_s.resize(n_dev);
    M->loc_i.resize(n_dev);
    auto bail = [&](int rc) {
        std::string msg = mqvs_last_error();
        multi_destroy(P);
        return mqvs_fail(rc, "%s", msg.c_str());
    };
    for (int p = 0; p < n_dev; ++p) {
        mqvs_store* c = nullptr;
        int rc = create_single(dev_ids[p], dim, store_dtype, M->map.count(p, capacity_rows), &c);
        if (rc != MQVS_OK) return bail(rc);
        M->parts.push_back(c);
        c->timing = false;
        P->sm_count = c->sm_count;
        if (cudaEventCreateWithFlags(&M->done[p], cudaEventDisableTiming) != cudaSuccess) return bail(mqvs_fail(MQVS_E_CUDA, "event creation failed"));
        if (dev_ids[p] == M->home) {
            M->direct[p] = 1;
        } else {  // this part's kernels store their results into home memory when the link allows it
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, dev_ids[p], M->home) == cudaSuccess && can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(M->home, 0);
                if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) M->direct[p] = 1;
            }
            cudaGetLastError();
        }
    }
    cudaSetDevice(M->home);
    if (cudaEventCreateWithFlags(&M->q_ev, cudaEventDisableTiming) != cudaSuccess) return bail(mqvs_fail(MQVS_E_CUDA, "event creation failed"));
    *out = P;
    return MQVS_OK;
}

void multi_destroy(mqvs_store* P) {
    if (!P) return;
    DeviceGuard guard;
    Multi* M = P->multi;
    if (M) {
        for (size_t p = 0; p < M->parts.size(); ++p) {
            cudaSetDevice(M->devs[p]);
            if (M->parts[p]->st) cudaStreamSynchronize(M->parts[p]->st);
            M->loc_s[p].release();
            M->loc_i[p].release();
            if (M->done[p]) cudaEventDestroy(M->done[p]);
        }
        cudaSetDevice(M->home);
        cudaDeviceSynchronize();
        M->g_s.release();
        M->g_i.release();
        M->q_pin.release();
        M->mask_pin.release();
        M->res_pin.release();
        if (M->q_ev) cudaEventDestroy(M->q_ev);
        for (mqvs_store* c : M->parts) destroy_single(c);
        cudaGetLastError();
        delete M;
    }
    delete P;
}

int multi_count(mqvs_store* P, uint64_t* out) {
    std::lock_guard<std::mutex> lk(P->mu);
    *out = P->multi->n;
    return MQVS_OK;
}

int multi_reserve(mqvs_store* P, uint64_t capacity_rows) {
    std::lock_guard<std::mutex> lk(P->mu);
    DeviceGuard guard;
    Multi* M = P->multi;
    for (int p = 0; p < M->W; ++p) {
        MQVS_CU(cudaSetDevice(M->devs[p]));
        int rc = grow(M->parts[p], M->map.count(p, capacity_rows));
        if (rc != MQVS_OK) return rc;
    }
    return MQVS_OK;
}

int multi_add(mqvs_store* P, const void* x, int x_is_device, bool synth, uint64_t seed, uint64_t first_index, uint64_t n,
              uint64_t* first_row_out) {
    std::lock_guard<std::mutex> lk(P->mu);
    DeviceGuard guard;
    Multi* M = P->multi;
    if (first_row_out) *first_row_out = M->n;
    if (n == 0) return MQVS_OK;
    const int W = M->W;
    std::vector<std::vector<AddSeg>> segs(W);
    for_block_runs(M->map, M->n, n, [&](int p, uint64_t /*local*/, uint64_t cnt, uint64_t off) {
        segs[p].push_back(AddSeg{x ? (const float*)x + off * P->dim : nullptr, cnt, first_index + off});
    });
    // (parts fill in global order, so part p's new rows go to its rows [count(p, n0), count(p, n0 + n)), in run order)
    std::vector<AddReport> rep(W);
    std::vector<int> rcs(W, MQVS_OK);
    std::vector<std::string> errs(W);
    auto stage_part = [&](int p) {
        if (segs[p].empty()) return;
        cudaSetDevice(M->devs[p]);
        mqvs_store* c = M->parts[p];
        if (x_is_device && !synth && !M->direct[p]) {
            // device rows on the home device, no peer access from this part: peer-copy each run into the part's staging buffer
            uint64_t total = 0;
            for (auto& g : segs[p]) total += g.n;
            DevBuf tmp;
            rcs[p] = tmp.ensure((size_t)total * P->dim * sizeof(float));
            if (rcs[p] == MQVS_OK) {
                std::vector<AddSeg> moved;
                uint64_t off = 0;
                for (auto& g : segs[p]) {
                    cudaMemcpyPeerAsync((float*)tmp.p + off * P->dim, M->devs[p], g.src, M->home, (size_t)g.n * P->dim * sizeof(float), c->st);
                    moved.push_back(AddSeg{(const float*)tmp.p + off * P->dim, g.n, 0});
                    off += g.n;
                }
                rcs[p] = add_stage(c, moved.data(), moved.size(), 1, false, 0, &rep[p]);
            }
            tmp.release();
        } else {
            rcs[p] = add_stage(c, segs[p].data(), segs[p].size(), x_is_device, synth, seed, &rep[p]);
        }
        if (rcs[p] != MQVS_OK) errs[p] = mqvs_last_error();
    };
    if (!x_is_device && !synth && W > 1) {  // host rows: one PCIe link per device, so one host thread per part
        std::vector<std::thread> th;
        for (int p = 0; p < W; ++p) th.emplace_back(stage_part, p);
        for (auto& t : th) t.join();
    } else {
        if (x_is_device && !synth) {  // rows on the home device: whatever produced them on the home's default stream is done
            cudaSetDevice(M->home);
            cudaDeviceSynchronize();
        }
        for (int p = 0; p < W; ++p) stage_part(p);
    }
    uint64_t bad = 0;
    for (int p = 0; p < W; ++p) {
        if (rcs[p] != MQVS_OK) return mqvs_fail(rcs[p], "%s", errs[p].c_str());
        bad += rep[p].bad;
    }
    if (bad)
        return mqvs_fail(MQVS_E_INVALID, "%llu of %llu embeddings have a non-finite norm; nothing was added",
                         (unsigned long long)bad, (unsigned long long)n);
    for (int p = 0; p < W; ++p) add_commit(M->parts[p], rep[p]);
    M->n += n;
    return MQVS_OK;
}

int multi_search(mqvs_store* P, const float* q, int q_is_device, int B, int k, const uint32_t* mask, int mask_is_device,
                 int64_t id_base, float* out_scores, int64_t* out_ids, int out_is_device, cudaStream_t caller_stream,
                 bool blocking) {
    std::lock_guard<std::mutex> lk(P->mu);
    DeviceGuard guard;
    Multi* M = P->multi;
    const int W = M->W;
    const size_t n_out = (size_t)B * k, q_bytes = (size_t)B * P->dim * sizeof(float);
    int rc;
    if (mask && mask_is_device && (mask != reinterpret_cast<const uint32_t*>(M) || !M->has_token))
        return mqvs_fail(MQVS_E_INVALID, "a device mask for a multi-device handle must be the pointer mqvs_filter_build returned");
    MQVS_CU(cudaSetDevice(M->home));
    mqvs_store* H = M->parts[0];
    cudaStream_t ms = blocking ? H->st : caller_stream;  // the merge runs here
    if ((rc = M->g_s.ensure(W * n_out * sizeof(float))) != MQVS_OK) return rc;
    if ((rc = M->g_i.ensure(W * n_out * sizeof(int64_t))) != MQVS_OK) return rc;
    if (!q_is_device) {
        if ((rc = M->q_pin.ensure(q_bytes)) != MQVS_OK) return rc;
        memcpy(M->q_pin.p, q, q_bytes);
    } else if (!blocking) {
        MQVS_CU(cudaEventRecord(M->q_ev, caller_stream));
    }
    // host mask words -> every part's local words, assembled in one pinned block
    std::vector<size_t> mask_off(W, 0);
    if (mask && !mask_is_device) {
        size_t total = 0;
        for (int p = 0; p < W; ++p) {
            mask_off[p] = total;
            total += (size_t)(M->parts[p]->n + 31) / 32 + 1;
        }
        if ((rc = M->mask_pin.ensure(total * 4)) != MQVS_OK) return rc;
        for (int p = 0; p < W; ++p) split_words(M->map, mask, M->n, p, (uint32_t*)M->mask_pin.p + mask_off[p]);
    }
    std::vector<const uint32_t*> mask_p(W, nullptr);
    std::vector<float*> os(W, nullptr);
    std::vector<int64_t*> oi(W, nullptr);
    std::vector<char> searched(W, 0);
    for (int p = 0; p < W; ++p) {
        mqvs_store* c = M->parts[p];
        float* gs = (float*)M->g_s.p + p * n_out;
        int64_t* gi = (int64_t*)M->g_i.p + p * n_out;
        if (c->n == 0 && p != 0) continue;  // (part 0 always runs: it validates the queries even when the store is empty)
        MQVS_CU(cudaSetDevice(M->devs[p]));
        const float* q_dev = nullptr;
        if (!q_is_device && !takes_latency_kernel(c, B, k)) {
            // zero-copy: the first kernel of the search (query normalisation / operand build) reads the raw queries once,
            // straight from the mapped pinned copy over this device's own PCIe link -- no copy to enqueue per device
            q_dev = (const float*)M->q_pin.p;
        } else if (!q_is_device) {
            if ((rc = c->qraw.ensure(q_bytes)) != MQVS_OK) return rc;
            MQVS_CU(cudaMemcpyAsync(c->qraw.p, M->q_pin.p, q_bytes, cudaMemcpyHostToDevice, c->st));
            q_dev = (const float*)c->qraw.p;
        } else {
            if (!blocking) MQVS_CU(cudaStreamWaitEvent(c->st, M->q_ev, 0));
            if (M->devs[p] == M->home) {
                q_dev = q;
            } else {
                if ((rc = c->qraw.ensure(q_bytes)) != MQVS_OK) return rc;
                MQVS_CU(cudaMemcpyPeerAsync(c->qraw.p, M->devs[p], q, M->home, q_bytes, c->st));
                q_dev = (const float*)c->qraw.p;
            }
        }
        if (mask && mask_is_device) {
            mask_p[p] = (const uint32_t*)c->f_mask.p;
        } else if (mask) {
            const size_t bytes = (size_t)((c->n + 31) / 32) * 4;
            if ((rc = c->maskbuf.ensure(std::max<size_t>(bytes, 4))) != MQVS_OK) return rc;
            if (bytes) MQVS_CU(cudaMemcpyAsync(c->maskbuf.p, (uint32_t*)M->mask_pin.p + mask_off[p], bytes, cudaMemcpyHostToDevice, c->st));
            mask_p[p] = (const uint32_t*)c->maskbuf.p;
        }
        if (M->direct[p]) {
            os[p] = gs;
            oi[p] = gi;
        } else {
            if ((rc = M->loc_s[p].ensure(n_out * sizeof(float))) != MQVS_OK) return rc;
            if ((rc = M->loc_i[p].ensure(n_out * sizeof(int64_t))) != MQVS_OK) return rc;
            os[p] = (float*)M->loc_s[p].p;
            oi[p] = (int64_t*)M->loc_i[p].p;
        }
        // blocking: the host looks at the flags after the merge and repairs only then (the common case has nothing to repair);
        // asynchronous: the device-side repair rounds run unconditionally, as on a single-device handle
        if ((rc = enqueue_search(c, q_dev, B, k, mask_p[p], 0, os[p], oi[p], c->st, nullptr, /*defer_repair=*/blocking,
                                 /*report_to_host=*/blocking)) != MQVS_OK)
            return rc;
        searched[p] = 1;
        if (!M->direct[p]) {
            MQVS_CU(cudaMemcpyPeerAsync(gs, M->home, os[p], M->devs[p], n_out * sizeof(float), c->st));
            MQVS_CU(cudaMemcpyPeerAsync(gi, M->home, oi[p], M->devs[p], n_out * sizeof(int64_t), c->st));
        }
        if (blocking && p == 0 && !(c->stats.last_path == 1 && c->last_fused))
            MQVS_CU(cudaMemcpyAsync(c->bad_host, c->bad_dev, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->st));
        MQVS_CU(cudaEventRecord(M->done[p], c->st));
    }
    MQVS_CU(cudaSetDevice(M->home));
    float* fs = out_scores;
    int64_t* fi = out_ids;
    if (!out_is_device) {
        if ((rc = M->res_pin.ensure(n_out * 12)) != MQVS_OK) return rc;
        fi = (int64_t*)M->res_pin.p;
        fs = (float*)((char*)M->res_pin.p + n_out * sizeof(int64_t));
    }
    auto merge = [&]() -> int {
        for (int p = 0; p < W; ++p) {
            if (searched[p]) {
                if (!(p == 0 && ms == H->st)) MQVS_CU(cudaStreamWaitEvent(ms, M->done[p], 0));
            } else {
                MQVS_CU(cudaMemsetAsync((int64_t*)M->g_i.p + p * n_out, 0xFF, n_out * sizeof(int64_t), ms));  // ids -1: no candidates
            }
        }
        launch_merge_candidates((const float*)M->g_s.p, (const int64_t*)M->g_i.p, W, B, k, fs, fi, ms, kShardShift, id_base);
        MQVS_CU(cudaGetLastError());
        ++M->merge_launches;
        return MQVS_OK;
    };
    if ((rc = merge()) != MQVS_OK) return rc;
    M->searches += 1;
    M->queries += (uint64_t)B;
    if (!blocking) return MQVS_OK;
    MQVS_CU(cudaStreamSynchronize(ms));
    {
        mqvs_store* c = M->parts[0];
        const bool epoch_word = c->stats.last_path == 1 && c->last_fused;
        const bool bad = epoch_word ? ((volatile uint32_t*)c->flags_map.p)[1] == c->epoch
                                    : (c->last_fused ? (unsigned int)c->bad_host[1] == c->epoch : c->bad_host[0] != 0);
        if (bad) return mqvs_fail(MQVS_E_INVALID, "queries with a non-finite norm in a batch of %d", B);
    }
    bool any = false;
    for (int p = 0; p < W; ++p) {
        if (!searched[p]) continue;
        mqvs_store* c = M->parts[p];
        bool changed = false;
        MQVS_CU(cudaSetDevice(M->devs[p]));
        if ((rc = finish_search(c, B, k, mask_p[p], 0, os[p], oi[p], c->st, &changed)) != MQVS_OK) return rc;
        if (changed) {
            any = true;
            if (!M->direct[p]) {
                MQVS_CU(cudaMemcpyPeerAsync((float*)M->g_s.p + p * n_out, M->home, os[p], M->devs[p], n_out * sizeof(float), c->st));
                MQVS_CU(cudaMemcpyPeerAsync((int64_t*)M->g_i.p + p * n_out, M->home, oi[p], M->devs[p], n_out * sizeof(int64_t), c->st));
            }
            MQVS_CU(cudaEventRecord(M->done[p], c->st));
        }
        settle_events(c);
    }
    if (any) {
        MQVS_CU(cudaSetDevice(M->home));
        if ((rc = merge()) != MQVS_OK) return rc;
        MQVS_CU(cudaStreamSynchronize(ms));
    }
    if (!out_is_device) {
        memcpy(out_scores, fs, n_out * sizeof(float));
        memcpy(out_ids, fi, n_out * sizeof(int64_t));
    }
    tc::trace_dump();
    return MQVS_OK;
}

int multi_get_rows(mqvs_store* P, const int64_t* row_ids, uint64_t n, float* out_host) {
    std::lock_guard<std::mutex> lk(P->mu);
    DeviceGuard guard;
    Multi* M = P->multi;
    std::vector<std::vector<int64_t>> ids(M->W);
    std::vector<std::vector<uint64_t>> pos(M->W);
    for (uint64_t i = 0; i < n; ++i) {
        if (row_ids[i] < 0 || (uint64_t)row_ids[i] >= M->n)
            return mqvs_fail(MQVS_E_INVALID, "row id %lld out of range [0, %llu)", (long long)row_ids[i], (unsigned long long)M->n);
        const int p = (int)M->map.part_of((uint64_t)row_ids[i]);
        ids[p].push_back((int64_t)M->map.local_of((uint64_t)row_ids[i]));
        pos[p].push_back(i);
    }
    std::vector<float> tmp;
    for (int p = 0; p < M->W; ++p) {
        if (ids[p].empty()) continue;
        tmp.resize(ids[p].size() * (size_t)P->dim);
        int rc = get_rows_single(M->parts[p], ids[p].data(), ids[p].size(), tmp.data());
        if (rc != MQVS_OK) return rc;
        for (size_t j = 0; j < ids[p].size(); ++j)
            memcpy(out_host + pos[p][j] * (size_t)P->dim, tmp.data() + j * (size_t)P->dim, (size_t)P->dim * sizeof(float));
    }
    return MQVS_OK;
}

// ---- tombstone compaction over several devices.  The survivors keep their global order, so surviving row r becomes global
// row r' = popcount(keep[0 .. r)) <= r and, the placement being block-cyclic in the global row, may change device.  Chunk by
// chunk in ascending r': every part gathers the chunk's source rows it holds (gather_raw_rows_kernel, bit for bit) and hands
// them to a pinned bounce buffer, the host puts them in chunk order, and the chunk is written to the (part, local row)
// coordinates of the new rows [a, a + c).  Those coordinates are the old rows [a, a + c): every live one among them has a new
// index below a + c and was consumed by this chunk or an earlier one, and no later source lies below a + c -- the same
// argument as the single-device compaction's.  Rows cross PCIe twice; compaction is rare and runs at a few GB/s.
int multi_compact(mqvs_store* P, const uint32_t* keep_words_host, uint64_t* new_count_out) {
    std::lock_guard<std::mutex> lk(P->mu);
    DeviceGuard guard;
    Multi* M = P->multi;
    const int W = M->W;
    for (int p = 0; p < W; ++p) {
        MQVS_CU(cudaSetDevice(M->devs[p]));
        MQVS_CU(cudaStreamSynchronize(M->parts[p]->st));
    }
    std::vector<uint64_t> keep;
    keep.reserve((size_t)M->n);
    for (uint64_t r = 0; r < M->n; ++r)
        if ((keep_words_host[r >> 5] >> (r & 31)) & 1u) keep.push_back(r);
    const uint64_t m = keep.size(), n_old = M->n;
    if (m == n_old) {
        *new_count_out = m;
        return MQVS_OK;
    }
    const size_t rb = P->row_bytes();
    const uint64_t chunk = 8192;
    PinBuf bounce;
    std::vector<PinBuf> part_host(W);
    std::vector<DevBuf> ids(W), tmp(W);
    auto release = [&]() {
        bounce.release();
        for (int p = 0; p < W; ++p) {
            cudaSetDevice(M->devs[p]);
            part_host[p].release();
            ids[p].release();
            tmp[p].release();
        }
    };
    auto broken = [&](cudaError_t e) {
        release();
        cudaGetLastError();
        return mqvs_fail(MQVS_E_CUDA, "compaction failed: %s (the store is in an undefined state)", cudaGetErrorString(e));
    };
    int rc = bounce.ensure(std::min(chunk, m ? m : 1) * rb);
    // stored norms follow their rows: small enough (4 bytes per row) to do on the host in one go, before any row moves
    const bool norms = M->parts[0]->keep_norms;
    std::vector<std::vector<float>> old_norms(W), new_norms(W);
    if (rc == MQVS_OK && norms) {
        for (int p = 0; p < W; ++p) {
            const uint64_t cnt = M->map.count(p, n_old);
            old_norms[p].resize((size_t)cnt);
            new_norms[p].resize((size_t)M->map.count(p, m));
            cudaSetDevice(M->devs[p]);
            if (cnt) {
                cudaError_t e = cudaMemcpy(old_norms[p].data(), M->parts[p]->norms, cnt * sizeof(float), cudaMemcpyDeviceToHost);
                if (e != cudaSuccess) return broken(e);
            }
        }
        for (uint64_t r = 0; r < m; ++r)
            new_norms[M->map.part_of(r)][M->map.local_of(r)] = old_norms[M->map.part_of(keep[r])][M->map.local_of(keep[r])];
    }
    if (rc != MQVS_OK) {
        release();
        return rc;
    }
    uint64_t launches = 0;
    std::vector<std::vector<uint32_t>> loc(W), pos(W);
    for (uint64_t a = 0; a < m; a += chunk) {
        const uint64_t c = std::min(chunk, m - a);
        if (keep[a + c - 1] == a + c - 1) continue;  // nothing dead up to here: these rows are already in place
        // 1. every part gathers the chunk's source rows it holds
        for (int p = 0; p < W; ++p) {
            loc[p].clear();
            pos[p].clear();
        }
        for (uint64_t j = 0; j < c; ++j) {
            const int p = (int)M->map.part_of(keep[a + j]);
            loc[p].push_back((uint32_t)M->map.local_of(keep[a + j]));
            pos[p].push_back((uint32_t)j);
        }
        for (int p = 0; p < W; ++p) {
            if (loc[p].empty()) continue;
            mqvs_store* part = M->parts[p];
            cudaSetDevice(M->devs[p]);
            const size_t np = loc[p].size();
            if ((rc = ids[p].ensure(chunk * sizeof(uint32_t))) != MQVS_OK || (rc = tmp[p].ensure(chunk * rb)) != MQVS_OK ||
                (rc = part_host[p].ensure(chunk * rb)) != MQVS_OK) {
                release();
                return rc;
            }
            cudaError_t e = cudaMemcpyAsync(ids[p].p, loc[p].data(), np * sizeof(uint32_t), cudaMemcpyHostToDevice, part->st);
            if (e != cudaSuccess) return broken(e);
            launch_gather_raw_rows(part->rows, rb, (const uint32_t*)ids[p].p, np, tmp[p].p, part->st);
            e = cudaMemcpyAsync(part_host[p].p, tmp[p].p, np * rb, cudaMemcpyDeviceToHost, part->st);
            if (e != cudaSuccess) return broken(e);
            ++launches;
        }
        for (int p = 0; p < W; ++p) {
            if (loc[p].empty()) continue;
            cudaSetDevice(M->devs[p]);
            cudaError_t e = cudaStreamSynchronize(M->parts[p]->st);
            if (e != cudaSuccess) return broken(e);
            for (size_t i = 0; i < pos[p].size(); ++i)
                memcpy((char*)bounce.p + (size_t)pos[p][i] * rb, (const char*)part_host[p].p + i * rb, rb);
        }
        // 2. the chunk, in order, to the coordinates of the new rows [a, a + c)
        cudaError_t werr = cudaSuccess;
        for_block_runs(M->map, a, c, [&](int p, uint64_t local, uint64_t cnt, uint64_t off) {
            if (werr != cudaSuccess) return;
            cudaSetDevice(M->devs[p]);
            werr = cudaMemcpyAsync((char*)M->parts[p]->rows + local * rb, (const char*)bounce.p + off * rb, cnt * rb,
                                   cudaMemcpyHostToDevice, M->parts[p]->st);
        });
        if (werr != cudaSuccess) return broken(werr);
        for (int p = 0; p < W; ++p) {  // the bounce buffer is reused by the next chunk
            cudaSetDevice(M->devs[p]);
            cudaError_t e = cudaStreamSynchronize(M->parts[p]->st);
            if (e != cudaSuccess) return broken(e);
        }
    }
    for (int p = 0; p < W; ++p) {
        mqvs_store* part = M->parts[p];
        cudaSetDevice(M->devs[p]);
        const uint64_t cnt = M->map.count(p, m);
        if (norms && cnt) {
            cudaError_t e = cudaMemcpy(part->norms, new_norms[p].data(), cnt * sizeof(float), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) return broken(e);
        }
        part->n = cnt;
        part->stats.rows = cnt;
        part->bounds_rows = 0;
        part->shadow_rows = 0;  // rows moved: the scoring shadow is rebuilt before the next tensor search
        for (auto& col : part->columns) col.used = false;  // row numbers changed under them
        part->f_table_dirty = true;
    }
    M->parts[0]->stats.kernel_launches += launches;
    M->n = m;
    M->has_token = false;
    release();
    *new_count_out = m;
    return MQVS_OK;
}

// ---- persistence: the same file as a single-device handle writes (rows in GLOBAL order), so a directory written on eight
// GPUs opens on one and the other way round
namespace {

bool flush_to_disk(FILE* f) { return fflush(f) == 0 && fsync(fileno(f)) == 0; }

int write_rows(mqvs_store* P, FILE* f, uint64_t first_row) {
    Multi* M = P->multi;
    const size_t rb = P->row_bytes();
    std::vector<char> host((size_t)kShardBlock * rb);
    int rc = MQVS_OK;
    for_block_runs(M->map, first_row, M->n - first_row, [&](int p, uint64_t local, uint64_t cnt, uint64_t) {
        if (rc != MQVS_OK) return;
        cudaSetDevice(M->devs[p]);
        cudaError_t e = cudaMemcpy(host.data(), (char*)M->parts[p]->rows + local * rb, cnt * rb, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            cudaGetLastError();
            rc = mqvs_fail(MQVS_E_CUDA, "device-to-host copy failed: %s", cudaGetErrorString(e));
        } else if (fwrite(host.data(), rb, cnt, f) != cnt) {
            rc = mqvs_fail(MQVS_E_IO, "short write");
        }
    });
    return rc;
}

int save_norms_multi(mqvs_store* P, const char* path) {
    Multi* M = P->multi;
    if (!M->parts[0]->keep_norms) return MQVS_OK;
    std::vector<float> host((size_t)std::max<uint64_t>(M->n, 1));
    int rc = MQVS_OK;
    for_block_runs(M->map, 0, M->n, [&](int p, uint64_t local, uint64_t cnt, uint64_t off) {
        if (rc != MQVS_OK) return;
        cudaSetDevice(M->devs[p]);
        if (cudaMemcpy(host.data() + off, M->parts[p]->norms + local, cnt * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) {
            cudaGetLastError();
            rc = mqvs_fail(MQVS_E_CUDA, "device-to-host copy of row norms failed");
        }
    });
    if (rc != MQVS_OK) return rc;
    const std::string dst = std::string(path) + ".norms", tmp = dst + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return mqvs_fail(MQVS_E_IO, "cannot open %s for writing", tmp.c_str());
    NormsHeader h{};
    memcpy(h.magic, "MQVSNRM1", 8);
    h.rows = M->n;
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1 && fwrite(host.data(), sizeof(float), (size_t)M->n, f) == (size_t)M->n;
    ok = ok && flush_to_disk(f);
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), dst.c_str()) != 0) {
        remove(tmp.c_str());
        return mqvs_fail(MQVS_E_IO, "writing %s failed", dst.c_str());
    }
    return MQVS_OK;
}

}  // namespace

int multi_save(mqvs_store* P, const char* path, uint64_t first_row) {
    std::lock_guard<std::mutex> lk(P->mu);
    DeviceGuard guard;
    Multi* M = P->multi;
    FileHeader h{};
    memcpy(h.magic, "MQVSROW1", 8);
    h.version = MQVS_VERSION;
    h.dim = (uint32_t)P->dim;
    h.dtype = (uint32_t)P->dtype;
    h.dim_pad = MQVS_DIM_PAD;
    h.rows = M->n;
    if (first_row == 0) {
        const std::string tmp = std::string(path) + ".tmp";
        FILE* f = fopen(tmp.c_str(), "wb");
        if (!f) return mqvs_fail(MQVS_E_IO, "cannot open %s for writing", tmp.c_str());
        int rc = fwrite(&h, sizeof(h), 1, f) == 1 ? write_rows(P, f, 0) : mqvs_fail(MQVS_E_IO, "short write");
        bool ok = rc == MQVS_OK && flush_to_disk(f);
        ok = (fclose(f) == 0) && ok;
        if (ok && (rc = save_norms_multi(P, path)) != MQVS_OK) ok = false;
        if (!ok || rename(tmp.c_str(), path) != 0) {
            remove(tmp.c_str());
            return rc != MQVS_OK ? rc : mqvs_fail(MQVS_E_IO, "writing %s failed", path);
        }
        return MQVS_OK;
    }
    if (first_row > M->n)
        return mqvs_fail(MQVS_E_INVALID, "first_row %llu beyond the %llu stored rows", (unsigned long long)first_row, (unsigned long long)M->n);
    FILE* f = fopen(path, "r+b");
    if (!f) return mqvs_fail(MQVS_E_IO, "cannot open %s for appending", path);
    FileHeader old{};
    if (fread(&old, sizeof(old), 1, f) != 1 || memcmp(old.magic, "MQVSROW1", 8) != 0) {
        fclose(f);
        return mqvs_fail(MQVS_E_IO, "%s is not an mqvs row file", path);
    }
    if ((int)old.dim != P->dim || (int)old.dtype != P->dtype || old.dim_pad != MQVS_DIM_PAD || old.rows != first_row) {
        fclose(f);
        return mqvs_fail(MQVS_E_INVALID, "%s holds %llu rows (dim=%u dtype=%u); cannot append from row %llu", path,
                         (unsigned long long)old.rows, old.dim, old.dtype, (unsigned long long)first_row);
    }
    int rc = fseek(f, (long)(sizeof(h) + first_row * P->row_bytes()), SEEK_SET) == 0 ? write_rows(P, f, first_row)
                                                                                      : mqvs_fail(MQVS_E_IO, "seek failed");
    if (rc == MQVS_OK) rc = save_norms_multi(P, path);
    if (rc != MQVS_OK) {
        fclose(f);
        return rc;
    }
    // the row count moves last, after the rows are on disk
    bool ok = flush_to_disk(f) && fseek(f, 0, SEEK_SET) == 0 && fwrite(&h, sizeof(h), 1, f) == 1 && flush_to_disk(f);
    ok = (fclose(f) == 0) && ok;
    if (!ok) return mqvs_fail(MQVS_E_IO, "appending to %s failed", path);
    return MQVS_OK;
}

int multi_load(mqvs_store* P, const char* path) {
    std::lock_guard<std::mutex> lk(P->mu);
    DeviceGuard guard;
    Multi* M = P->multi;
    FILE* f = fopen(path, "rb");
    if (!f) return mqvs_fail(MQVS_E_IO, "cannot open %s", path);
    FileHeader h{};
    if (fread(&h, sizeof(h), 1, f) != 1 || memcmp(h.magic, "MQVSROW1", 8) != 0) {
        fclose(f);
        return mqvs_fail(MQVS_E_IO, "%s is not an mqvs row file", path);
    }
    if ((int)h.dim != P->dim || (int)h.dtype != P->dtype || h.dim_pad != MQVS_DIM_PAD) {
        fclose(f);
        return mqvs_fail(MQVS_E_INVALID, "%s holds dim=%u dtype=%u, store is dim=%d dtype=%d", path, h.dim, h.dtype, P->dim, P->dtype);
    }
    int rc = MQVS_OK;
    for (int p = 0; p < M->W && rc == MQVS_OK; ++p) {
        cudaSetDevice(M->devs[p]);
        M->parts[p]->n = 0;
        M->parts[p]->shadow_rows = 0;  // rows are rewritten in place below
        rc = grow(M->parts[p], M->map.count(p, h.rows));
    }
    M->n = 0;
    const size_t rb = P->row_bytes();
    std::vector<char> host((size_t)kShardBlock * rb);
    if (rc == MQVS_OK)
        for_block_runs(M->map, 0, h.rows, [&](int p, uint64_t local, uint64_t cnt, uint64_t) {
            if (rc != MQVS_OK) return;
            if (fread(host.data(), rb, cnt, f) != cnt) {
                rc = mqvs_fail(MQVS_E_IO, "%s is truncated", path);
                return;
            }
            cudaSetDevice(M->devs[p]);
            if (cudaMemcpy((char*)M->parts[p]->rows + local * rb, host.data(), cnt * rb, cudaMemcpyHostToDevice) != cudaSuccess) {
                cudaGetLastError();
                rc = mqvs_fail(MQVS_E_CUDA, "host-to-device copy failed");
            }
        });
    fclose(f);
    if (rc != MQVS_OK) return rc;
    if (M->parts[0]->keep_norms) {
        const std::string src = std::string(path) + ".norms";
        FILE* g = fopen(src.c_str(), "rb");
        if (!g)
            return mqvs_fail(MQVS_E_IO, "%s is missing: the collection was saved without row norms (cosine space) and cannot be "
                                        "opened in the l2 / ip space", src.c_str());
        NormsHeader nh{};
        std::vector<float> nh_host((size_t)std::max<uint64_t>(h.rows, 1));
        const bool ok = fread(&nh, sizeof(nh), 1, g) == 1 && memcmp(nh.magic, "MQVSNRM1", 8) == 0 && nh.rows >= h.rows &&
                        fread(nh_host.data(), sizeof(float), (size_t)h.rows, g) == (size_t)h.rows;
        fclose(g);
        if (!ok) return mqvs_fail(MQVS_E_IO, "%s does not hold the norms of %llu rows", src.c_str(), (unsigned long long)h.rows);
        for (mqvs_store* c : M->parts) {
            c->norm_dev = 0.0f;
            c->norm_max = 0.0f;
            c->bounds_rows = 0;
        }
        for_block_runs(M->map, 0, h.rows, [&](int p, uint64_t local, uint64_t cnt, uint64_t off) {
            if (rc != MQVS_OK) return;
            mqvs_store* c = M->parts[p];
            cudaSetDevice(M->devs[p]);
            if (cudaMemcpy(c->norms + local, nh_host.data() + off, cnt * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
                cudaGetLastError();
                rc = mqvs_fail(MQVS_E_CUDA, "host-to-device copy of row norms failed");
                return;
            }
            for (uint64_t r = 0; r < cnt; ++r) {
                c->norm_dev = std::max(c->norm_dev, fabsf(nh_host[off + r] - 1.0f));
                c->norm_max = std::max(c->norm_max, nh_host[off + r]);
            }
        });
        if (rc != MQVS_OK) return rc;
    }
    for (int p = 0; p < M->W; ++p) {
        M->parts[p]->n = M->map.count(p, h.rows);
        M->parts[p]->stats.rows = M->parts[p]->n;
    }
    M->n = h.rows;
    M->has_token = false;
    return MQVS_OK;
}

int multi_column_set(mqvs_store* P, int col_id, int col_type, const void* values_host, const uint8_t* present_host, uint64_t n) {
    std::lock_guard<std::mutex> lk(P->mu);
    DeviceGuard guard;
    Multi* M = P->multi;
    std::vector<std::vector<uint64_t>> vals(M->W);
    std::vector<std::vector<uint8_t>> pres(M->W);
    for (int p = 0; p < M->W; ++p) {
        vals[p].reserve(M->map.count(p, n));
        if (present_host) pres[p].reserve(M->map.count(p, n));
    }
    const uint64_t* v = (const uint64_t*)values_host;
    for_block_runs(M->map, 0, n, [&](int p, uint64_t, uint64_t cnt, uint64_t off) {
        vals[p].insert(vals[p].end(), v + off, v + off + cnt);
        if (present_host) pres[p].insert(pres[p].end(), present_host + off, present_host + off + cnt);
    });
    for (int p = 0; p < M->W; ++p) {
        int rc = column_set_single(M->parts[p], col_id, col_type, vals[p].data(), present_host ? pres[p].data() : nullptr, vals[p].size());
        if (rc != MQVS_OK) return rc;
    }
    return MQVS_OK;
}

int multi_column_set_tokens(mqvs_store* P, int col_id, const uint64_t* offsets_host, const int64_t* tokens_host, uint64_t n) {
    std::lock_guard<std::mutex> lk(P->mu);
    DeviceGuard guard;
    Multi* M = P->multi;
    std::vector<std::vector<uint64_t>> off(M->W);
    std::vector<std::vector<int64_t>> tok(M->W);
    for (int p = 0; p < M->W; ++p) off[p].push_back(0);
    for_block_runs(M->map, 0, n, [&](int p, uint64_t, uint64_t cnt, uint64_t first) {
        for (uint64_t r = first; r < first + cnt; ++r) {
            tok[p].insert(tok[p].end(), tokens_host + offsets_host[r], tokens_host + offsets_host[r + 1]);
            off[p].push_back(tok[p].size());
        }
    });
    for (int p = 0; p < M->W; ++p) {
        int rc = column_set_tokens_single(M->parts[p], col_id, off[p].data(), tok[p].data(), off[p].size() - 1);
        if (rc != MQVS_OK) return rc;
    }
    return MQVS_OK;
}

int multi_column_set_text(mqvs_store* P, int col_id, const uint64_t* offsets_host, const uint8_t* bytes_host,
                          const uint8_t* present_host, uint64_t n) {
    std::lock_guard<std::mutex> lk(P->mu);
    DeviceGuard guard;
    Multi* M = P->multi;
    std::vector<std::vector<uint64_t>> off(M->W);
    std::vector<std::vector<uint8_t>> txt(M->W), pres(M->W);
    for (int p = 0; p < M->W; ++p) off[p].push_back(0);
    for_block_runs(M->map, 0, n, [&](int p, uint64_t, uint64_t cnt, uint64_t first) {
        // a run of rows is a contiguous byte range of the CSR text
        txt[p].insert(txt[p].end(), bytes_host + offsets_host[first], bytes_host + offsets_host[first + cnt]);
        const uint64_t shift = off[p].back() - offsets_host[first];
        for (uint64_t r = first; r < first + cnt; ++r) off[p].push_back(offsets_host[r + 1] + shift);
        if (present_host) pres[p].insert(pres[p].end(), present_host + first, present_host + first + cnt);
    });
    for (int p = 0; p < M->W; ++p) {
        int rc = column_set_text_single(M->parts[p], col_id, off[p].data(), txt[p].data(), present_host ? pres[p].data() : nullptr,
                                        off[p].size() - 1);
        if (rc != MQVS_OK) return rc;
    }
    return MQVS_OK;
}

int multi_filter_build(mqvs_store* P, const mqvs_filter_op* prog, int n_ops, const void* pool, uint64_t pool_len,
                       const uint32_t* live_words_host, const uint32_t** mask_dev_out, uint64_t* allowed_out) {
    std::lock_guard<std::mutex> lk(P->mu);
    DeviceGuard guard;
    Multi* M = P->multi;
    uint64_t allowed = 0;
    std::vector<uint32_t> live;
    for (int p = 0; p < M->W; ++p) {
        mqvs_store* c = M->parts[p];
        const uint32_t* live_p = nullptr;
        if (live_words_host) {
            live.assign((size_t)(c->n + 31) / 32 + 1, 0u);
            split_words(M->map, live_words_host, M->n, p, live.data());
            live_p = live.data();
        }
        const uint32_t* m = nullptr;
        uint64_t a = 0;
        int rc = filter_build_single(c, prog, n_ops, pool, pool_len, live_p, &m, &a);
        if (rc != MQVS_OK) return rc;
        allowed += a;
    }
    M->has_token = true;
    *mask_dev_out = reinterpret_cast<const uint32_t*>(M);  // opaque token: every part's own f_mask is the mask
    *allowed_out = allowed;
    return MQVS_OK;
}

int multi_get_stats(mqvs_store* P, mqvs_stats* out) {
    std::lock_guard<std::mutex> lk(P->mu);
    DeviceGuard guard;
    Multi* M = P->multi;
    mqvs_stats t{};
    for (int p = 0; p < M->W; ++p) {
        mqvs_store* c = M->parts[p];
        fill_stats(c);
        const mqvs_stats& s = c->stats;
        t.kernel_launches += s.kernel_launches;
        t.scan_launches += s.scan_launches;
        t.scan_bytes += s.scan_bytes;
        t.scan_ms = std::max(t.scan_ms, s.scan_ms);  // the parts scan concurrently: device time of the slowest
        t.last_search_ms = std::max(t.last_search_ms, s.last_search_ms);
        t.last_scan_ms = std::max(t.last_scan_ms, s.last_scan_ms);
        t.last_scan_passes = std::max(t.last_scan_passes, s.last_scan_passes);
        t.exact_fallbacks += s.exact_fallbacks;
        t.unrepaired += s.unrepaired;
        t.band_repairs += s.band_repairs;
        if (c->n) {
            t.last_path = s.last_path;
            t.last_pipeline = s.last_pipeline;
        }
    }
    t.rows = M->n;
    t.searches = M->searches;
    t.queries = M->queries;
    t.kernel_launches += M->merge_launches;
    *out = t;
    return MQVS_OK;
}

int multi_set_metric(mqvs_store* P, int metric) {
    std::lock_guard<std::mutex> lk(P->mu);
    DeviceGuard guard;
    for (mqvs_store* c : P->multi->parts) {
        int rc = mqvs_set_metric(c, metric);
        if (rc != MQVS_OK) return rc;
    }
    P->metric = metric;
    return MQVS_OK;
}

int multi_norm_deviation(mqvs_store* P, float* out) {
    std::lock_guard<std::mutex> lk(P->mu);
    float d = 0.0f;
    for (mqvs_store* c : P->multi->parts) d = std::max(d, c->norm_dev);
    *out = d;
    return MQVS_OK;
}

int multi_get_norms(mqvs_store* P, uint64_t first_row, uint64_t n, float* out_host) {
    std::lock_guard<std::mutex> lk(P->mu);
    DeviceGuard guard;
    Multi* M = P->multi;
    if (!M->parts[0]->keep_norms) return mqvs_fail(MQVS_E_INVALID, "norms are kept only for the l2 / ip metrics (mqvs_set_metric)");
    if (first_row > M->n || n > M->n - first_row)
        return mqvs_fail(MQVS_E_INVALID, "rows [%llu, %llu) out of range [0, %llu)", (unsigned long long)first_row,
                         (unsigned long long)(first_row + n), (unsigned long long)M->n);
    int rc = MQVS_OK;
    for_block_runs(M->map, first_row, n, [&](int p, uint64_t local, uint64_t cnt, uint64_t off) {
        if (rc != MQVS_OK) return;
        cudaSetDevice(M->devs[p]);
        cudaStreamSynchronize(M->parts[p]->st);
        if (cudaMemcpy(out_host + off, M->parts[p]->norms + local, cnt * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) {
            cudaGetLastError();
            rc = mqvs_fail(MQVS_E_CUDA, "device-to-host copy of row norms failed");
        }
    });
    return rc;
}

int multi_set_path(mqvs_store* P, int path) {
    std::lock_guard<std::mutex> lk(P->mu);
    for (mqvs_store* c : P->multi->parts) {
        std::lock_guard<std::mutex> lc(c->mu);
        c->path = path;
    }
    P->path = path;
    return MQVS_OK;
}

int multi_set_shadow(mqvs_store* P, int enable) {
    std::lock_guard<std::mutex> lk(P->mu);
    DeviceGuard guard;
    for (mqvs_store* c : P->multi->parts) {
        std::lock_guard<std::mutex> lc(c->mu);
        cudaSetDevice(c->dev);
        c->shadow_on = enable != 0 && c->dtype == MQVS_F32;
        if (!c->shadow_on && c->shadow) {
            cudaStreamSynchronize(c->st);
            cudaFree(c->shadow);
            c->shadow = nullptr;
            c->shadow_cap = c->shadow_rows = 0;
        }
    }
    return MQVS_OK;
}

int multi_shadow_bytes(mqvs_store* P, uint64_t* bytes_out) {
    std::lock_guard<std::mutex> lk(P->mu);
    uint64_t total = 0;
    for (mqvs_store* c : P->multi->parts) total += c->shadow ? c->shadow_cap * (uint64_t)MQVS_DIM_PAD * 2 : 0;
    *bytes_out = total;
    return MQVS_OK;
}

}  // namespace api
}  // namespace mqvs

extern "C" MQVS_API int mqvs_shard_of_row(int n_dev, uint64_t row, int* shard_out, uint64_t* local_row_out) {
    if (n_dev < 1 || n_dev > MQVS_MAX_DEVICES || !shard_out || !local_row_out)
        return mqvs::mqvs_fail(MQVS_E_INVALID, "n_dev must be in [1, %d] and the outputs non-NULL", MQVS_MAX_DEVICES);
    mqvs::ShardMap m;
    m.W = n_dev;
    *shard_out = (int)m.part_of(row);
    *local_row_out = m.local_of(row);
    return MQVS_OK;
}
