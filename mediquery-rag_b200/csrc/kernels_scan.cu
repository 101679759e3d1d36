// kernels_scan.cu -- K2 (SIMT) + K3: bandwidth-bound exact fp32 scan with fused per-warp top-k.
//
// Does the arithmetic of vectorstore.similarity_search(query, k) (/root/reference/src/agents/nodes.py:93,
// src/medical_engine.py:70) for small query batches: every stored row is streamed from HBM exactly once per
// pass (128-bit / 64-bit coalesced ld.global.nc, L1 no-allocate), scored against QB register-resident
// queries in the canonical fp32 order (common.cuh), and offered to a warp-private top-k list that lives in
// registers.  Roofline: HBM; algorithmic bytes per pass = n_rows * 768 * sizeof(elem).
//
// Top-k (K3): per (warp, query) an unsorted list of 64-bit keys spread over the lanes (slot j on lane j%32),
// the smallest key cached as a score threshold.  A row is inserted only when it beats the threshold, which
// after warm-up happens with probability ~ k / rows_seen.  At the end the warps' lists are sorted together in
// shared memory and the CTA writes its best k keys, sorted, to cand[q][cta][k].
// The optional allow-mask (Chroma `where`, compiled on the host) is tested BEFORE the row is loaded, so
// filtered-out rows cost no HBM traffic.
#include "common.cuh"
#include "launch.h"

namespace mqvs {

namespace {

constexpr int kScanWarps = 8;
constexpr int kScanThreads = kScanWarps * 32;

template <int QB, int KPL>
struct TopK {
    uint64_t keys[QB][KPL];
    uint64_t minkey[QB];
    float thr[QB];

    __device__ __forceinline__ void init(int k, int lane) {
#pragma unroll
        for (int b = 0; b < QB; ++b) {
#pragma unroll
            for (int j = 0; j < KPL; ++j) keys[b][j] = (j * 32 + lane) < k ? 0ull : ~0ull;
            minkey[b] = 0ull;
            thr[b] = -INFINITY;
        }
    }
    // warp-uniform: replace the smallest key of list b with `key` (caller checked key > minkey[b])
    __device__ __forceinline__ void insert(int b, uint64_t key, int lane) {
        bool done = false;
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
            unsigned hit = __ballot_sync(0xffffffffu, !done && keys[b][j] == minkey[b]);
            if (hit) {
                if (lane == __ffs(hit) - 1) keys[b][j] = key;
                done = true;
            }
        }
        uint64_t m = keys[b][0];
#pragma unroll
        for (int j = 1; j < KPL; ++j) m = keys[b][j] < m ? keys[b][j] : m;
        m = warp_min_u64(m);
        minkey[b] = m;
        thr[b] = m ? key_score(m) : -INFINITY;
    }
};

// The row walk shared by every SIMT kernel: warp `gw` of `warps_total` visits groups of U consecutive rows in
// ascending order, scores them against the QB register-resident queries and offers them to the warp's top-k lists.
// l2 / ip spaces over the raw vectors (mqvs_set_metric): the cosine of the normalised pair is mapped to the space's
// ranking score with the row's stored norm (common.cuh:rank_score) before it meets the top-k list.
template <int QB>
struct SpaceXf {
    int metric;
    const float* __restrict__ norms;  // [n_rows] ||x|| of the raw rows (K1)
    float h[QB];                      // 0.5 / max(||q||, eps) per register-resident query
};

template <typename T, int QB, int KPL, int U, bool XF = false>
__device__ __forceinline__ void scan_rows(const T* __restrict__ rows, uint32_t n_rows, const float4 (&q)[QB][kChunks],
                                          const uint32_t* __restrict__ mask, TopK<QB, KPL>& top, int lane,
                                          const SpaceXf<QB>* xf = nullptr) {
    const uint32_t warps_total = gridDim.x * kScanWarps;
    const uint32_t gw = blockIdx.x * kScanWarps + (threadIdx.x >> 5);
    for (uint64_t r0 = (uint64_t)gw * U; r0 < n_rows; r0 += (uint64_t)warps_total * U) {
        RowFrag<T> x[U];
        bool live[U];
        uint32_t mw = 0xffffffffu;
        if (mask != nullptr) mw = __ldg(mask + (r0 >> 5)) >> (r0 & 31);  // U divides 32: one word covers the group
#pragma unroll
        for (int u = 0; u < U; ++u) {
            live[u] = (r0 + u < n_rows) && ((mw >> u) & 1u);
            if (live[u]) x[u].load(rows + (r0 + u) * (uint64_t)kDim, lane);
        }
        float nrm[U];
        if constexpr (XF) {
#pragma unroll
            for (int u = 0; u < U; ++u) nrm[u] = live[u] ? __ldg(xf->norms + r0 + u) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!live[u]) continue;
            float4 xv[kChunks];
#pragma unroll
            for (int c = 0; c < kChunks; ++c) xv[c] = x[u].get(c);
#pragma unroll
            for (int b = 0; b < QB; ++b) {
                float s = warp_butterfly_sum(lane_dot(q[b], xv));
                if constexpr (XF) s = rank_score(xf->metric, s, nrm[u], xf->h[b]);
                if (s >= top.thr[b]) {
                    uint64_t key = pack_key(s, (uint32_t)(r0 + u));
                    if (key > top.minkey[b]) top.insert(b, key, lane);
                }
            }
        }
    }
}

// CTA-level reduction: all warps' lists of query b -> sort -> the CTA's best k, written to cand[b][blockIdx.x][k].
template <int QB, int KPL>
__device__ __forceinline__ void cta_reduce(const TopK<QB, KPL>& top, int nq, int k, uint64_t* __restrict__ cand,
                                           int cand_stride, uint64_t* smem_keys, int lane, int warp) {
    const int sort_n = next_pow2(kScanWarps * k);
    for (int i = threadIdx.x; i < QB * sort_n; i += kScanThreads) smem_keys[i] = 0ull;
    __syncthreads();
#pragma unroll
    for (int b = 0; b < QB; ++b)
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
            int slot = j * 32 + lane;
            if (slot < k) smem_keys[b * sort_n + warp * k + slot] = top.keys[b][j];
        }
    __syncthreads();
    for (int b = 0; b < nq; ++b) {
        uint64_t* s = smem_keys + b * sort_n;
        block_bitonic_sort_desc(s, sort_n);
        uint64_t* out = cand + ((size_t)b * cand_stride + blockIdx.x) * k;
        for (int j = threadIdx.x; j < k; j += kScanThreads) out[j] = s[j];
    }
}

template <typename T, int QB, int KPL, int U, bool XF = false>
__device__ __forceinline__ void scan_body(const T* __restrict__ rows, uint32_t n_rows, const float* __restrict__ qn, int nq,
                                          const uint32_t* __restrict__ mask, int k, uint64_t* __restrict__ cand,
                                          int cand_stride, uint64_t* smem_keys, const SpaceXf<QB>* xf = nullptr) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // queries -> registers (lane's canonical share); missing queries are zero and never written out
    float4 q[QB][kChunks];
#pragma unroll
    for (int b = 0; b < QB; ++b)
#pragma unroll
        for (int c = 0; c < kChunks; ++c)
            q[b][c] = b < nq ? *reinterpret_cast<const float4*>(qn + (size_t)b * kDim + c * 128 + lane * 4)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
    TopK<QB, KPL> top;
    top.init(k, lane);
    scan_rows<T, QB, KPL, U, XF>(rows, n_rows, q, mask, top, lane, xf);
    cta_reduce<QB, KPL>(top, nq, k, cand, cand_stride, smem_keys, lane, warp);
}

// ------------------------------------------------------------------------------------------------ latency path
// One launch for the whole search when B <= 4 and k <= 16 (the reference's own regime: one query, k = 3..5,
// /root/reference/src/agents/nodes.py:93): each warp normalises the raw queries itself (K1's arithmetic, so the
// result is bit-identical), the scan runs as above, and the LAST CTA to finish (atomic ticket) merges the per-CTA
// candidates and writes the final scores/ids -- no separate normalise, memset or merge launches.
struct FusedArgs {
    const float* q_raw;      // [nq, dim] fp32, un-normalised
    int dim;
    int vec4;
    float* out_scores;       // [nq, k]
    int64_t* out_ids;
    int64_t id_base;
    unsigned int* ticket;    // zero on entry; the last CTA leaves it zero again
    unsigned int* bad_epoch; // raised to `epoch` (atomicMax) when a query has a non-finite norm
    unsigned int epoch;
    unsigned int* host_status;  // optional mapped host words: [0] <- epoch when the results are written, [1] <- 1 if a query was non-finite
};

template <typename T, int QB, int U>
__device__ __forceinline__ void scan_fused_body(const T* __restrict__ rows, uint32_t n_rows, int nq, const uint32_t* __restrict__ mask,
                                                int k, uint64_t* __restrict__ cand, const FusedArgs& a) {
    __shared__ __align__(16) uint64_t smem_keys[QB * 128];  // next_pow2(8 warps * k), k <= 16
    __shared__ uint64_t win[8 * 32];
    __shared__ unsigned int s_ticket;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 q[QB][kChunks];
#pragma unroll
    for (int b = 0; b < QB; ++b) {
        if (b < nq) {
            load_lane_share(a.q_raw + (size_t)b * a.dim, a.dim, a.vec4 != 0, lane, q[b]);
            const bool ok = normalize_lane_share(q[b]);
            if (!ok && blockIdx.x == 0 && warp == 0 && lane == 0) atomicMax(a.bad_epoch, a.epoch);
        } else {
#pragma unroll
            for (int c = 0; c < kChunks; ++c) q[b][c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    TopK<QB, 1> top;
    top.init(k, lane);
    scan_rows<T, QB, 1, U>(rows, n_rows, q, mask, top, lane);
    cta_reduce<QB, 1>(top, nq, k, cand, (int)gridDim.x, smem_keys, lane, warp);
    // publish this CTA's candidates, take a ticket; the last CTA merges
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(a.ticket, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    __threadfence();
    for (int b = 0; b < nq; ++b)
        block_select_small<24>(cand + (size_t)b * gridDim.x * k, (int)gridDim.x * k, k, a.id_base,
                               a.out_scores + (size_t)b * k, a.out_ids + (size_t)b * k, win);
    if (threadIdx.x == 0) {
        *a.ticket = 0u;
        if (a.host_status != nullptr) {  // results went straight to mapped host memory: report status the same way
            a.host_status[1] = (*reinterpret_cast<volatile unsigned int*>(a.bad_epoch) == a.epoch) ? 1u : 0u;
            __threadfence_system();
            a.host_status[0] = a.epoch;
        }
    }
}

template <typename T, int QB, int U>
__global__ void __launch_bounds__(kScanThreads)
scan_fused_kernel(const T* __restrict__ rows, uint32_t n_rows, int nq, const uint32_t* __restrict__ mask, int k,
                  uint64_t* __restrict__ cand, FusedArgs a) {
    scan_fused_body<T, QB, U>(rows, n_rows, nq, mask, k, cand, a);
}

// The same search with the raw queries INSIDE the kernel's parameter block (up to 4 x MQVS_DIM_PAD floats: 12 KB at 768 --
// kernel parameters may be ~32 KB from CUDA 12.1 on): a host caller's queries then need no host-to-device copy of their own,
// the launch itself carries them (the reference's regime -- one query, 154 chunks -- is launch-latency-bound, and the
// separate copy was a quarter of the call).  __grid_constant__: the address of the parameter may be taken, every warp
// reads its lane shares from it as from any other read-only memory.
template <int QB>
struct InlineQueries {
    float v[QB * kDim];
};
template <typename T, int QB, int U>
__global__ void __launch_bounds__(kScanThreads)
scan_fused_inline_kernel(const T* __restrict__ rows, uint32_t n_rows, int nq, const uint32_t* __restrict__ mask, int k,
                         uint64_t* __restrict__ cand, FusedArgs a, const __grid_constant__ InlineQueries<QB> q) {
    a.q_raw = q.v;
    scan_fused_body<T, QB, U>(rows, n_rows, nq, mask, k, cand, a);
}

template <typename T, int QB, int KPL, int U>
__global__ void __launch_bounds__(kScanThreads)
scan_simt_kernel(const T* __restrict__ rows, uint32_t n_rows, const float* __restrict__ qn, int nq,
                 const uint32_t* __restrict__ mask, int k, uint64_t* __restrict__ cand, int cand_stride) {
    extern __shared__ __align__(16) uint64_t smem_keys[];  // [QB][sort_n]
    scan_body<T, QB, KPL, U>(rows, n_rows, qn, nq, mask, k, cand, cand_stride, smem_keys);
}

template <typename T, int QB, int KPL, int U>
__global__ void __launch_bounds__(kScanThreads)
scan_simt_metric_kernel(const T* __restrict__ rows, uint32_t n_rows, const float* __restrict__ qn, int nq,
                        const uint32_t* __restrict__ mask, int k, uint64_t* __restrict__ cand, int cand_stride, int metric,
                        const float* __restrict__ row_norms, const float* __restrict__ q_norms) {
    extern __shared__ __align__(16) uint64_t smem_keys[];  // [QB][sort_n]
    SpaceXf<QB> xf;
    xf.metric = metric;
    xf.norms = row_norms;
#pragma unroll
    for (int b = 0; b < QB; ++b) xf.h[b] = b < nq ? __fdiv_rn(0.5f, fmaxf(__ldg(q_norms + b), kEps)) : 0.0f;
    scan_body<T, QB, KPL, U, true>(rows, n_rows, qn, nq, mask, k, cand, cand_stride, smem_keys, &xf);
}

// Device-driven repair pass of the tensor-core path: re-scans the corpus exactly for the (few) queries whose
// shortlist could not be proven exact.  Every CTA derives the list of flagged queries itself (one warp, ballot +
// popcount over flags[B], first max_fix in index order), so no separate compaction launch is needed and no host round
// trip; block 0 publishes the list for merge_fix_kernel and updates the repair counters.  With no flag set every CTA
// returns at once.
template <typename T, int KPL, int U, bool XF>
__global__ void __launch_bounds__(kScanThreads)
scan_fix_kernel(const T* __restrict__ rows, uint32_t n_rows, const float* __restrict__ qn_all,
                const uint32_t* __restrict__ flags, int B, int max_fix, uint32_t* __restrict__ fix_list,
                uint32_t* __restrict__ fix_n, unsigned long long* __restrict__ counters,
                const uint32_t* __restrict__ mask, int k, uint64_t* __restrict__ cand, int cand_stride, int metric,
                const float* __restrict__ row_norms, const float* __restrict__ q_norms_all, int count_leftover,
                const uint32_t* __restrict__ flag_count) {
    extern __shared__ __align__(16) uint64_t smem_keys[];
    __shared__ uint32_t list_s[1024];
    __shared__ uint32_t n_s, total_s;
    const int lane = threadIdx.x & 31;
    // Nothing flagged (the common case of the rounds an asynchronous search enqueues blind): one word says so.  Walking the
    // B flags instead cost every CTA of every idle round B / 32 dependent loads -- 104 us per round at B = 16384, 64 rounds
    // per search: a quarter of config 5's step (profiles/r02_ncu_launches_c5shard_*.csv).  flag_count >= the flags that are
    // set (every setter adds, the repairs subtract), so zero is safe to trust; any other value only costs the walk.
    if (__ldcg(flag_count) == 0u) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *fix_n = 0u;
        return;
    }
    if (threadIdx.x < 32) {
        uint32_t base = 0;
        for (int c0 = 0; c0 < B; c0 += 256) {  // eight independent loads in flight per lane
            uint32_t fl[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) fl[u] = c0 + u * 32 + lane < B ? __ldcg(flags + c0 + u * 32 + lane) : 0u;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int c = c0 + u * 32;
                const bool f = fl[u] != 0u;
                const unsigned m = __ballot_sync(0xffffffffu, f);
                const uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
                if (f && pos < (uint32_t)max_fix) list_s[pos] = (uint32_t)(c + lane);
                base += __popc(m);
            }
        }
        if (lane == 0) {
            total_s = base;
            n_s = base < (uint32_t)max_fix ? base : (uint32_t)max_fix;
        }
    }
    __syncthreads();
    const uint32_t n = n_s;
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            *fix_n = n;
            if (total_s) {
                atomicAdd(counters + 0, (unsigned long long)n);              // repaired on the device
                // left flagged after the LAST round of a search (earlier rounds leave them to the next one)
                if (count_leftover) atomicAdd(counters + 1, (unsigned long long)(total_s - n));
            }
        }
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) fix_list[i] = list_s[i];
    }
    for (uint32_t f = 0; f < n; ++f) {
        if constexpr (XF) {  // l2 / ip spaces: the same ranking score as scan_simt_metric_kernel
            SpaceXf<1> xf;
            xf.metric = metric;
            xf.norms = row_norms;
            xf.h[0] = __fdiv_rn(0.5f, fmaxf(__ldg(q_norms_all + list_s[f]), kEps));
            scan_body<T, 1, KPL, U, true>(rows, n_rows, qn_all + (size_t)list_s[f] * kDim, 1, mask, k,
                                          cand + (size_t)f * cand_stride * k, cand_stride, smem_keys, &xf);
        } else {
            scan_body<T, 1, KPL, U>(rows, n_rows, qn_all + (size_t)list_s[f] * kDim, 1, mask, k,
                                    cand + (size_t)f * cand_stride * k, cand_stride, smem_keys);
        }
        __syncthreads();
    }
}

struct ScanCfg {
    int qb, kpl;
};

ScanCfg pick_cfg(int nq, int k) {
    ScanCfg c;
    c.kpl = k <= 32 ? 1 : (k <= 128 ? 4 : 32);
    int qmax = c.kpl == 1 ? 4 : (c.kpl == 4 ? 2 : 1);
    c.qb = nq >= 4 && qmax >= 4 ? 4 : (nq >= 2 && qmax >= 2 ? 2 : 1);
    return c;
}

using ScanKernel = void (*)(const void*, uint32_t, const float*, int, const uint32_t*, int, uint64_t*, int);

template <typename T, int QB, int KPL, int U>
const void* kernel_ptr() {
    return reinterpret_cast<const void*>(&scan_simt_kernel<T, QB, KPL, U>);
}

template <typename T, int QB, int KPL, int U>
const void* metric_kernel_ptr() {
    return reinterpret_cast<const void*>(&scan_simt_metric_kernel<T, QB, KPL, U>);
}

const void* find_kernel(int dtype, const ScanCfg& c, bool metric = false) {
#define MQVS_CASE(QB_, KPL_)                                                                                  \
    if (c.qb == QB_ && c.kpl == KPL_) {                                                                       \
        if (metric)                                                                                           \
            return dtype == MQVS_F32 ? metric_kernel_ptr<float, QB_, KPL_, 2>()                               \
                                     : metric_kernel_ptr<__nv_bfloat16, QB_, KPL_, 4>();                      \
        return dtype == MQVS_F32 ? kernel_ptr<float, QB_, KPL_, 2>() : kernel_ptr<__nv_bfloat16, QB_, KPL_, 4>(); \
    }
    MQVS_CASE(1, 1)
    MQVS_CASE(2, 1)
    MQVS_CASE(4, 1)
    MQVS_CASE(1, 4)
    MQVS_CASE(2, 4)
    MQVS_CASE(1, 32)
#undef MQVS_CASE
    return nullptr;
}

size_t scan_smem(const ScanCfg& c, int k) { return (size_t)c.qb * next_pow2(kScanWarps * k) * sizeof(uint64_t); }

constexpr int kMaxCtasPerSm = 3;

}  // namespace

bool scan_fused_ok(int B, int k) { return B >= 1 && B <= 4 && k <= 16; }

template <typename T, int QB, int U>
int launch_fused_t(const void* rows, uint32_t n_rows, int nq, const uint32_t* mask, int k, uint64_t* cand, int sm_count,
                   const FusedArgs& a, cudaStream_t st, const float* q_host) {
    auto kern = scan_fused_kernel<T, QB, U>;
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)(q_host ? (const void*)scan_fused_inline_kernel<T, QB, U> : (const void*)kern),
                                                      kScanThreads, 0) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        occ = 1;
    }
    int grid = sm_count * (occ < 2 ? occ : 2);  // <= 296 CTAs: grid * k <= 4736 keys for the in-kernel merge
    // tiny stores (the reference ships 154 chunks): no more CTAs than there are row groups, so the last CTA merges
    // a handful of candidate lists instead of 296 mostly empty ones
    const uint32_t groups = (n_rows + kScanWarps * U - 1) / (kScanWarps * U);
    if ((uint32_t)grid > groups) grid = groups > 0 ? (int)groups : 1;
    if (q_host != nullptr) {
        // the queries ride in the parameter block: rows of a.dim floats each, packed (the kernel reads them with stride a.dim)
        InlineQueries<QB> q;
        memcpy(q.v, q_host, (size_t)nq * a.dim * sizeof(float));
        scan_fused_inline_kernel<T, QB, U><<<grid, kScanThreads, 0, st>>>((const T*)rows, n_rows, nq, mask, k, cand, a, q);
        return grid;
    }
    kern<<<grid, kScanThreads, 0, st>>>((const T*)rows, n_rows, nq, mask, k, cand, a);
    return grid;
}

int scan_fused_max_grid(int sm_count) { return sm_count * 2; }

void launch_scan_fused(const void* rows, int dtype, uint32_t n_rows, const float* q_raw, int dim, int nq,
                       const uint32_t* mask, int k, uint64_t* cand, float* out_scores, int64_t* out_ids, int64_t id_base,
                       unsigned int* ticket, unsigned int* bad_epoch, unsigned int epoch, int sm_count, cudaStream_t st,
                       unsigned int* host_status, const float* q_host) {
    // q_host: the raw queries in HOST memory -- they travel inside the kernel's parameter block (q_raw is then unused)
    // (parameter space is not global memory: the streaming 128-bit ld.global path is for q_raw only)
    FusedArgs a{q_raw, dim, q_host ? 0 : ((dim % 4 == 0 && (reinterpret_cast<uintptr_t>(q_raw) & 15) == 0) ? 1 : 0),
                out_scores, out_ids, id_base, ticket, bad_epoch, epoch, host_status};
    const int qb = nq >= 3 ? 4 : nq;  // 1, 2 or 4 register-resident queries
#define MQVS_FUSED(T_, U_)                                                                                     \
    (qb == 1   ? launch_fused_t<T_, 1, U_>(rows, n_rows, nq, mask, k, cand, sm_count, a, st, q_host)             \
     : qb == 2 ? launch_fused_t<T_, 2, U_>(rows, n_rows, nq, mask, k, cand, sm_count, a, st, q_host)             \
               : launch_fused_t<T_, 4, U_>(rows, n_rows, nq, mask, k, cand, sm_count, a, st, q_host))
    if (dtype == MQVS_F32)
        MQVS_FUSED(float, 2);
    else
        MQVS_FUSED(__nv_bfloat16, 4);
#undef MQVS_FUSED
}

int scan_fix_grid(int sm_count) { return sm_count * 2; }

void launch_scan_fix(const void* rows, int dtype, uint32_t n_rows, const float* qn_all, const uint32_t* flags, int B, int max_fix,
                     uint32_t* fix_list, uint32_t* fix_n, unsigned long long* counters, const uint32_t* mask, int k, uint64_t* cand,
                     int sm_count, cudaStream_t st, int metric, const float* row_norms, const float* q_norms_all,
                     bool count_leftover, const uint32_t* flag_count) {
    const int kpl = k <= 32 ? 1 : (k <= 128 ? 4 : 32);
    const size_t smem = (size_t)next_pow2(kScanWarps * k) * sizeof(uint64_t);
    const int grid = scan_fix_grid(sm_count), stride = grid;  // every slot of a listed query is written: no memset
    if (max_fix > 1024) max_fix = 1024;  // list_s
    const int leftover = count_leftover ? 1 : 0;
#define MQVS_FIX(T_, KPL_, U_)                                                                                                     \
    do {                                                                                                                           \
        if (smem > 48 * 1024) {                                                                                                    \
            cudaFuncSetAttribute(scan_fix_kernel<T_, KPL_, U_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);     \
            cudaFuncSetAttribute(scan_fix_kernel<T_, KPL_, U_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        }                                                                                                                          \
        if (metric != MQVS_METRIC_COSINE)                                                                                          \
            scan_fix_kernel<T_, KPL_, U_, true><<<grid, kScanThreads, smem, st>>>(                                                 \
                (const T_*)rows, n_rows, qn_all, flags, B, max_fix, fix_list, fix_n, counters, mask, k, cand, stride, metric,      \
                row_norms, q_norms_all, leftover, flag_count);                                                                     \
        else                                                                                                                       \
            scan_fix_kernel<T_, KPL_, U_, false><<<grid, kScanThreads, smem, st>>>(                                                \
                (const T_*)rows, n_rows, qn_all, flags, B, max_fix, fix_list, fix_n, counters, mask, k, cand, stride, 0, nullptr,  \
                nullptr, leftover, flag_count);                                                                                    \
    } while (0)
    if (dtype == MQVS_F32) {
        if (kpl == 1) MQVS_FIX(float, 1, 2); else if (kpl == 4) MQVS_FIX(float, 4, 2); else MQVS_FIX(float, 32, 2);
    } else {
        if (kpl == 1) MQVS_FIX(__nv_bfloat16, 1, 4); else if (kpl == 4) MQVS_FIX(__nv_bfloat16, 4, 4); else MQVS_FIX(__nv_bfloat16, 32, 4);
    }
#undef MQVS_FIX
}

void scan_plan(int dtype, int nq_remaining, int k, int sm_count, int* nq_pass, int* grid) {
    ScanCfg c = pick_cfg(nq_remaining, k);
    const void* kern = find_kernel(dtype, c);
    size_t smem = scan_smem(c, k);
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kScanThreads, smem) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        occ = 1;
    }
    *nq_pass = nq_remaining < c.qb ? nq_remaining : c.qb;
    *grid = sm_count * (occ < kMaxCtasPerSm ? occ : kMaxCtasPerSm);
}

int scan_max_grid(int sm_count) { return sm_count * kMaxCtasPerSm; }

void launch_scan_simt_metric(const void* rows, int dtype, uint32_t n_rows, const float* qn, int nq, const uint32_t* mask,
                             int k, uint64_t* cand, int cand_stride, int grid, int metric, const float* row_norms,
                             const float* q_norms, cudaStream_t st) {
    ScanCfg c = pick_cfg(nq, k);
    const void* kern = find_kernel(dtype, c, true);
    const size_t smem = scan_smem(c, k);
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    void* args[] = {(void*)&rows, (void*)&n_rows, (void*)&qn, (void*)&nq, (void*)&mask, (void*)&k, (void*)&cand,
                    (void*)&cand_stride, (void*)&metric, (void*)&row_norms, (void*)&q_norms};
    cudaLaunchKernel(kern, dim3(grid), dim3(kScanThreads), args, smem, st);
}

void launch_scan_simt(const void* rows, int dtype, uint32_t n_rows, const float* qn, int nq, const uint32_t* mask,
                      int k, uint64_t* cand, int cand_stride, int grid, cudaStream_t st) {
    ScanCfg c = pick_cfg(nq, k);
    const void* kern = find_kernel(dtype, c);
    void* args[] = {(void*)&rows, (void*)&n_rows, (void*)&qn, (void*)&nq, (void*)&mask, (void*)&k, (void*)&cand,
                    (void*)&cand_stride};
    cudaLaunchKernel(kern, dim3(grid), dim3(kScanThreads), args, scan_smem(c, k), st);
}

}  // namespace mqvs
