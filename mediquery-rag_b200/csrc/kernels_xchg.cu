// kernels_xchg.cu -- C1 fused: cross-shard candidate exchange over NVLink peer memory + the W*k -> k merge, ONE kernel.
//
// after the local search every rank holds its exact top-k per query ([B, k] scores + global
// ids); the global top-k is the best k of the W*k candidates.  The NCCL form of this step is two all-gathers plus
// mqvs_merge_candidates (three launches, ~50-90 us of latency for 80 B .. 2 MB messages).  Here each rank owns an
// "inbox" allocation that every peer maps through CUDA IPC; xchg_merge_kernel
//   1. stores this rank's candidates of the CTA's queries straight into every peer's inbox (P2P stores over NVLink),
//   2. raises a per-(sender, CTA) flag in each peer's inbox (fence.sys + st.release.sys),
//   3. waits for the W flags of its own CTA index in the local inbox (ld.acquire.sys),
//   4. merges the W*k candidates of its queries (score descending, ties -> lower global id) and writes [B, k].
// CTA i only ever depends on CTA i of the peers, so no grid-wide synchronisation is needed on either side.  Inboxes are
// double-buffered by call parity: a peer can be at most one call ahead (it needs our flag of call n+1 to finish n+1).
#include <cuda_runtime.h>

#include <cstring>
#include <new>

#include "common.cuh"
#include "launch.h"

namespace mqvs {
namespace {

constexpr int kMaxWorld = 8;
constexpr int kXThreads = 128;
constexpr int kMaxGrid = 2048;   // flag slots per sender
constexpr int kHandleBytes = 64;

struct XchgArgs {
    const float* loc_s;      // [B, k] this rank's candidates
    const int64_t* loc_i;
    float* peer_s[kMaxWorld];      // inbox of rank r (this call's parity): scores [W][cap]
    int64_t* peer_i[kMaxWorld];    //                                      ids    [W][cap]
    uint32_t* peer_flag[kMaxWorld];  //                                    flags  [W][kMaxGrid]
    int B, k, W, rank, cap, n_pow2;
    uint32_t epoch;
    float* out_s;            // [B, k]
    int64_t* out_i;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_cg_f32(const float* p) { return __ldcg(p); }

__device__ __forceinline__ bool before(uint32_t ao, int64_t ai, uint32_t bo, int64_t bi) {  // a ranks strictly before b
    return ao > bo || (ao == bo && ai < bi);
}

// The work of CTA `cta` of `n_cta` of the rank described by `a`.
__device__ __forceinline__ void xchg_body(const XchgArgs& a, const int cta, const int n_cta) {
    extern __shared__ __align__(16) unsigned char raw[];
    int64_t* sid = reinterpret_cast<int64_t*>(raw);
    uint32_t* sord = reinterpret_cast<uint32_t*>(sid + a.n_pow2);
    const int tid = threadIdx.x;
    const int per = a.W * a.k;

    // 1. this CTA's queries -> every inbox (own included), slot [rank]
    for (int b = cta; b < a.B; b += n_cta)
        for (int idx = tid; idx < per; idx += kXThreads) {
            const int r = idx / a.k, j = idx - r * a.k;
            const size_t src = (size_t)b * a.k + j, dst = (size_t)a.rank * a.cap + src;
            a.peer_s[r][dst] = a.loc_s[src];
            a.peer_i[r][dst] = a.loc_i[src];
        }
    __syncthreads();
    // 2. signal every peer, 3. wait for every peer
    if (tid < a.W) {
        __threadfence_system();
        st_release_sys(a.peer_flag[tid] + (size_t)a.rank * kMaxGrid + cta, a.epoch);
        const uint32_t* mine = a.peer_flag[a.rank] + (size_t)tid * kMaxGrid + cta;
        const long long t0 = clock64();
        while (ld_acquire_sys(mine) != a.epoch) {
            __nanosleep(100);
            if (clock64() - t0 > 60000000000ll) __trap();  // ~30 s: a peer that never arrives must not hang the device
        }
        __threadfence_system();
    }
    __syncthreads();
    // 4. merge W*k -> k per query
    const float* in_s = a.peer_s[a.rank];
    const int64_t* in_i = a.peer_i[a.rank];
    for (int b = cta; b < a.B; b += n_cta) {
        for (int i = tid; i < a.n_pow2; i += kXThreads) {
            uint32_t ord = 0;
            int64_t id = INT64_MAX;
            if (i < per) {
                const int r = i / a.k, j = i - r * a.k;
                const size_t src = (size_t)r * a.cap + (size_t)b * a.k + j;
                const int64_t v = __ldcg(in_i + src);
                if (v >= 0) {
                    id = v;
                    ord = (uint32_t)(pack_key(ld_cg_f32(in_s + src), 0) >> 32);
                }
            }
            sord[i] = ord;
            sid[i] = id;
        }
        for (int size = 2; size <= a.n_pow2; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                __syncthreads();
                for (int t = tid; t < (a.n_pow2 >> 1); t += kXThreads) {
                    const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                    const bool desc = ((lo & size) == 0);
                    const uint32_t ao = sord[lo], co = sord[hi];
                    const int64_t ai = sid[lo], ci = sid[hi];
                    const bool swap = desc ? before(co, ci, ao, ai) : before(ao, ai, co, ci);
                    if (swap) {
                        sord[lo] = co; sid[lo] = ci;
                        sord[hi] = ao; sid[hi] = ai;
                    }
                }
            }
        __syncthreads();
        for (int j = tid; j < a.k; j += kXThreads) {
            const size_t o = (size_t)b * a.k + j;
            const bool ok = j < a.n_pow2 && sord[j] != 0;
            a.out_s[o] = ok ? key_score((uint64_t)sord[j] << 32) : -INFINITY;
            a.out_i[o] = ok ? sid[j] : -1;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kXThreads) xchg_merge_kernel(const XchgArgs a) { xchg_body(a, (int)blockIdx.x, (int)gridDim.x); }

// All W ranks of an exchange in ONE cooperative launch (every CTA co-resident): CTAs [r * per_rank, (r + 1) * per_rank) do
// rank r's work.  Ranks that share a GPU must not run as separate launches -- a kernel spinning on a flag that another
// launch has yet to write is not guaranteed to ever see that launch scheduled -- so this is how W logical ranks on one
// device are exercised (tests) or served (several shards of one exchange placed on one GPU).
__global__ void __launch_bounds__(kXThreads) xchg_merge_all_kernel(const XchgArgs* __restrict__ all, int per_rank) {
    const int rank = (int)blockIdx.x / per_rank;
    xchg_body(all[rank], (int)blockIdx.x % per_rank, per_rank);
}

}  // namespace
}  // namespace mqvs

struct mqvs_xchg {
    int device = 0, rank = 0, world = 1;
    uint64_t cap = 0;  // entries (B * k) per sender per call
    size_t bytes = 0;
    char* own = nullptr;
    char* peer[mqvs::kMaxWorld] = {};
    bool connected = false;
    bool borrowed = false;  // peers were given as raw pointers (same process): nothing to close
    uint32_t epoch = 0;
    int sm_count = 148;

    size_t off_scores(int parity) const { return (size_t)parity * world * cap * sizeof(float); }
    size_t off_ids(int parity) const { return 2 * (size_t)world * cap * sizeof(float) + (size_t)parity * world * cap * sizeof(int64_t); }
    size_t off_flags(int parity) const {
        return 2 * (size_t)world * cap * (sizeof(float) + sizeof(int64_t)) + (size_t)parity * world * mqvs::kMaxGrid * sizeof(uint32_t);
    }
    size_t total_bytes() const { return off_flags(2); }
};

extern "C" {

MQVS_API int mqvs_xchg_create(int device, int rank, int world, uint64_t max_entries, mqvs_xchg_t** out) {
    using namespace mqvs;
    if (!out) return mqvs_fail(MQVS_E_INVALID, "out is NULL");
    *out = nullptr;
    if (world < 2 || world > kMaxWorld || rank < 0 || rank >= world)
        return mqvs_fail(MQVS_E_INVALID, "peer exchange serves 2..%d ranks of one box (rank %d of %d)", kMaxWorld, rank, world);
    if (max_entries == 0 || max_entries > (1ull << 26)) return mqvs_fail(MQVS_E_INVALID, "max_entries out of range");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return mqvs_fail(MQVS_E_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    mqvs_xchg* x = new (std::nothrow) mqvs_xchg();
    if (!x) return mqvs_fail(MQVS_E_OOM, "host allocation failed");
    x->device = device;
    x->rank = rank;
    x->world = world;
    x->cap = (max_entries + 3) / 4 * 4;  // keeps every sender's slice 16-byte aligned
    x->bytes = x->total_bytes();
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) x->sm_count = prop.multiProcessorCount;
    e = cudaMalloc(&x->own, x->bytes);
    if (e == cudaSuccess) e = cudaMemset(x->own, 0, x->bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (x->own) cudaFree(x->own);
        delete x;
        return mqvs_fail(MQVS_E_OOM, "inbox allocation of %.1f MB failed: %s", x->bytes / 1e6, cudaGetErrorString(e));
    }
    x->peer[rank] = x->own;
    *out = x;
    return MQVS_OK;
}

MQVS_API int mqvs_xchg_get_handle(mqvs_xchg_t* x, void* handle_out, int handle_bytes) {
    using namespace mqvs;
    if (!x || !handle_out || handle_bytes < kHandleBytes) return mqvs_fail(MQVS_E_INVALID, "need a %d-byte handle buffer", kHandleBytes);
    static_assert(sizeof(cudaIpcMemHandle_t) == kHandleBytes, "CUDA IPC handle size");
    cudaSetDevice(x->device);
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, x->own);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return mqvs_fail(MQVS_E_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    }
    memcpy(handle_out, &h, kHandleBytes);
    return MQVS_OK;
}

MQVS_API int mqvs_xchg_connect(mqvs_xchg_t* x, const void* handles) {
    using namespace mqvs;
    if (!x || !handles) return mqvs_fail(MQVS_E_INVALID, "NULL argument");
    cudaSetDevice(x->device);
    for (int r = 0; r < x->world; ++r) {
        if (r == x->rank || x->peer[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + (size_t)r * kHandleBytes, kHandleBytes);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return mqvs_fail(MQVS_E_CUDA, "cudaIpcOpenMemHandle for rank %d: %s", r, cudaGetErrorString(e));
        }
        x->peer[r] = (char*)p;
    }
    x->connected = true;
    return MQVS_OK;
}

MQVS_API int mqvs_xchg_base(mqvs_xchg_t* x, void** base_out) {
    if (!x || !base_out) return mqvs::mqvs_fail(MQVS_E_INVALID, "NULL argument");
    *base_out = x->own;
    return MQVS_OK;
}

MQVS_API int mqvs_xchg_connect_ptrs(mqvs_xchg_t* x, void* const* bases) {
    using namespace mqvs;
    if (!x || !bases) return mqvs_fail(MQVS_E_INVALID, "NULL argument");
    cudaSetDevice(x->device);
    for (int r = 0; r < x->world; ++r) {
        if (r == x->rank) continue;
        if (!bases[r]) return mqvs_fail(MQVS_E_INVALID, "no inbox pointer for rank %d", r);
        int peer_dev = -1;
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, bases[r]) == cudaSuccess) peer_dev = attr.device;
        cudaGetLastError();
        if (peer_dev >= 0 && peer_dev != x->device) {
            cudaError_t e = cudaDeviceEnablePeerAccess(peer_dev, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();
                return mqvs_fail(MQVS_E_CUDA, "no peer access from device %d to %d: %s", x->device, peer_dev, cudaGetErrorString(e));
            }
            cudaGetLastError();
        }
        x->peer[r] = (char*)bases[r];
    }
    x->borrowed = true;
    x->connected = true;
    return MQVS_OK;
}

MQVS_API int mqvs_xchg_max_entries(mqvs_xchg_t* x, uint64_t* out) {
    if (!x || !out) return mqvs::mqvs_fail(MQVS_E_INVALID, "NULL argument");
    *out = x->cap;
    return MQVS_OK;
}

}  // extern "C"

static int xchg_check(mqvs_xchg_t* x, const float* scores_dev, const int64_t* ids_dev, int B, int k, float* out_scores_dev,
                      int64_t* out_ids_dev) {
    using namespace mqvs;
    if (!x || !x->connected) return mqvs_fail(MQVS_E_INVALID, "exchange is not connected");
    if (B < 0 || k < 1 || (uint64_t)B * k > x->cap || (long long)x->world * k > 16384)
        return mqvs_fail(MQVS_E_INVALID, "exchange shape B=%d k=%d exceeds the inbox (%llu entries)", B, k, (unsigned long long)x->cap);
    if (B && (!scores_dev || !ids_dev || !out_scores_dev || !out_ids_dev)) return mqvs_fail(MQVS_E_INVALID, "NULL buffer");
    return MQVS_OK;
}

// argument block of this rank's next call (advances its epoch)
static mqvs::XchgArgs xchg_args(mqvs_xchg_t* x, const float* scores_dev, const int64_t* ids_dev, int B, int k, float* out_scores_dev,
                                int64_t* out_ids_dev) {
    using namespace mqvs;
    const uint32_t epoch = ++x->epoch;
    const int parity = (int)(epoch & 1u);
    XchgArgs a{};
    a.loc_s = scores_dev;
    a.loc_i = ids_dev;
    for (int r = 0; r < x->world; ++r) {
        a.peer_s[r] = (float*)(x->peer[r] + x->off_scores(parity));
        a.peer_i[r] = (int64_t*)(x->peer[r] + x->off_ids(parity));
        a.peer_flag[r] = (uint32_t*)(x->peer[r] + x->off_flags(parity));
    }
    a.B = B;
    a.k = k;
    a.W = x->world;
    a.rank = x->rank;
    a.cap = (int)x->cap;
    a.n_pow2 = next_pow2(x->world * k);
    a.epoch = epoch;
    a.out_s = out_scores_dev;
    a.out_i = out_ids_dev;
    return a;
}

extern "C" {

MQVS_API int mqvs_xchg_merge(mqvs_xchg_t* x, const float* scores_dev, const int64_t* ids_dev, int B, int k,
                             float* out_scores_dev, int64_t* out_ids_dev, void* cuda_stream) {
    using namespace mqvs;
    int rc = xchg_check(x, scores_dev, ids_dev, B, k, out_scores_dev, out_ids_dev);
    if (rc != MQVS_OK) return rc;
    if (B == 0) return MQVS_OK;
    cudaError_t e = cudaSetDevice(x->device);
    if (e != cudaSuccess) return mqvs_fail(MQVS_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    const XchgArgs a = xchg_args(x, scores_dev, ids_dev, B, k, out_scores_dev, out_ids_dev);
    // every CTA must be able to run without waiting for another CTA of this grid to retire (CTA i waits for the peers'
    // CTA i): keep the grid within what is resident at once
    int grid = B < x->sm_count * 8 ? B : x->sm_count * 8;
    if (grid > kMaxGrid) grid = kMaxGrid;
    const size_t smem = (size_t)a.n_pow2 * (sizeof(int64_t) + sizeof(uint32_t));
    if (smem > 48 * 1024) cudaFuncSetAttribute(xchg_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    xchg_merge_kernel<<<grid, kXThreads, smem, (cudaStream_t)cuda_stream>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) return mqvs_fail(MQVS_E_CUDA, "xchg_merge_kernel launch: %s", cudaGetErrorString(e));
    return MQVS_OK;
}

MQVS_API int mqvs_xchg_merge_colocated(mqvs_xchg_t* const* ranks, int W, const float* const* scores_dev,
                                       const int64_t* const* ids_dev, int B, int k, float* const* out_scores_dev,
                                       int64_t* const* out_ids_dev, void* cuda_stream) {
    using namespace mqvs;
    if (!ranks || W < 2 || W > kMaxWorld || !scores_dev || !ids_dev || !out_scores_dev || !out_ids_dev)
        return mqvs_fail(MQVS_E_INVALID, "need the 2..%d ranks of one exchange and their buffers", kMaxWorld);
    for (int r = 0; r < W; ++r) {
        if (!ranks[r] || ranks[r]->world != W || ranks[r]->rank != r || ranks[r]->device != ranks[0]->device)
            return mqvs_fail(MQVS_E_INVALID, "rank %d: the exchanges must be ranks 0..W-1 of one world on one device", r);
        int rc = xchg_check(ranks[r], scores_dev[r], ids_dev[r], B, k, out_scores_dev[r], out_ids_dev[r]);
        if (rc != MQVS_OK) return rc;
    }
    if (B == 0) return MQVS_OK;
    cudaError_t e = cudaSetDevice(ranks[0]->device);
    if (e != cudaSuccess) return mqvs_fail(MQVS_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    XchgArgs host[kMaxWorld];
    for (int r = 0; r < W; ++r) host[r] = xchg_args(ranks[r], scores_dev[r], ids_dev[r], B, k, out_scores_dev[r], out_ids_dev[r]);
    XchgArgs* dev = nullptr;
    e = cudaMalloc(&dev, sizeof(XchgArgs) * W);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dev, host, sizeof(XchgArgs) * W, cudaMemcpyHostToDevice, (cudaStream_t)cuda_stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (dev) cudaFree(dev);
        return mqvs_fail(MQVS_E_CUDA, "argument upload failed: %s", cudaGetErrorString(e));
    }
    const size_t smem = (size_t)host[0].n_pow2 * (sizeof(int64_t) + sizeof(uint32_t));
    if (smem > 48 * 1024) cudaFuncSetAttribute(xchg_merge_all_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, xchg_merge_all_kernel, kXThreads, smem) != cudaSuccess || occ < 1) occ = 1;
    int per_rank = ranks[0]->sm_count * occ / W;  // the whole grid must be resident at once (cooperative launch)
    if (per_rank > B) per_rank = B;
    if (per_rank > kMaxGrid) per_rank = kMaxGrid;
    if (per_rank < 1) {
        cudaFree(dev);
        return mqvs_fail(MQVS_E_INVALID, "%d ranks do not fit on the device at once", W);
    }
    void* args[] = {(void*)&dev, (void*)&per_rank};
    e = cudaLaunchCooperativeKernel((const void*)xchg_merge_all_kernel, dim3(per_rank * W), dim3(kXThreads), args, smem, (cudaStream_t)cuda_stream);
    cudaError_t e2 = cudaStreamSynchronize((cudaStream_t)cuda_stream);  // `dev` is freed below; this entry point is not a hot path
    cudaFree(dev);
    if (e != cudaSuccess || e2 != cudaSuccess) {
        cudaGetLastError();
        return mqvs_fail(MQVS_E_CUDA, "xchg_merge_all_kernel: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    }
    return MQVS_OK;
}

MQVS_API void mqvs_xchg_destroy(mqvs_xchg_t* x) {
    if (!x) return;
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < x->world; ++r)
        if (r != x->rank && x->peer[r] && !x->borrowed) cudaIpcCloseMemHandle(x->peer[r]);
    if (x->own) cudaFree(x->own);
    cudaGetLastError();
    delete x;
}

}  // extern "C"
