// kernels_select.cu -- final selection: per-CTA candidate keys -> best k per query, and the cross-shard merge
// that follows the all-gather of per-rank candidates.  Ordering rule everywhere:
// score descending, ties -> lower id; empty slots come out as (-inf, -1).
#include "common.cuh"
#include "launch.h"

namespace mqvs {

namespace {

constexpr int kSelThreads = 256;

// One CTA per query.  Streams the m candidate keys through a shared-memory buffer of `buf_n` keys whose first
// `kp` (= next_pow2(k)) slots always hold the best keys seen so far, sorted.
__global__ void __launch_bounds__(kSelThreads)
merge_keys_kernel(const uint64_t* __restrict__ keys_in, int m, int k, int kp, int buf_n, int64_t id_base,
                  float* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
    extern __shared__ __align__(16) uint64_t buf[];
    const uint64_t* in = keys_in + (size_t)blockIdx.x * m;
    const int chunk = buf_n - kp;
    for (int i = threadIdx.x; i < kp; i += kSelThreads) buf[i] = 0ull;
    for (int base = 0; base < m; base += chunk) {
        for (int i = threadIdx.x; i < chunk; i += kSelThreads) buf[kp + i] = (base + i < m) ? in[base + i] : 0ull;
        block_bitonic_sort_desc(buf, buf_n);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += kSelThreads) {
        uint64_t key = buf[j];
        size_t o = (size_t)blockIdx.x * k + j;
        out_scores[o] = key ? key_score(key) : -INFINITY;
        out_ids[o] = key ? (int64_t)key_id(key) + id_base : -1;
    }
}

// Small-k variant (k <= 32): no sort (block_select_small, common.cuh).  For the batch-1 latency path this replaces
// three 2048-key bitonic sorts (~25 us) by ~3 us.
template <int kPerLane>
__global__ void __launch_bounds__(kSelThreads)
merge_keys_small_kernel(const uint64_t* __restrict__ keys_in, int m, int k, int64_t id_base,
                        float* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
    __shared__ uint64_t win[8 * 32];
    block_select_small<kPerLane>(keys_in + (size_t)blockIdx.x * m, m, k, id_base, out_scores + (size_t)blockIdx.x * k,
                                 out_ids + (size_t)blockIdx.x * k, win);
}

// Repair pass of the tensor-core path: block f merges the exact-scan candidates of flagged query fix_list[f] into
// that query's output row and clears its flag.  Same body as merge_keys_kernel, indirect through the device list.
__global__ void __launch_bounds__(kSelThreads)
merge_fix_kernel(const uint64_t* __restrict__ keys_in, int m, int k, int kp, int buf_n, int64_t id_base,
                 const uint32_t* __restrict__ fix_list, const uint32_t* __restrict__ fix_n, uint32_t* __restrict__ flags,
                 float* __restrict__ out_scores, int64_t* __restrict__ out_ids, uint32_t* __restrict__ flag_count) {
    extern __shared__ __align__(16) uint64_t buf[];
    if (blockIdx.x >= *fix_n) return;
    const uint32_t q = fix_list[blockIdx.x];
    const uint64_t* in = keys_in + (size_t)blockIdx.x * m;
    const int chunk = buf_n - kp;
    for (int i = threadIdx.x; i < kp; i += kSelThreads) buf[i] = 0ull;
    for (int base = 0; base < m; base += chunk) {
        for (int i = threadIdx.x; i < chunk; i += kSelThreads) buf[kp + i] = (base + i < m) ? in[base + i] : 0ull;
        block_bitonic_sort_desc(buf, buf_n);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += kSelThreads) {
        uint64_t key = buf[j];
        size_t o = (size_t)q * k + j;
        out_scores[o] = key ? key_score(key) : -INFINITY;
        out_ids[o] = key ? (int64_t)key_id(key) + id_base : -1;
    }
    if (threadIdx.x == 0) {
        flags[q] = 0u;
        atomicSub(flag_count, 1u);  // (keeps the count exact, so the rounds after the last repair exit on one word)
    }
}

// Cross-shard merge with 64-bit ids: entries are (order-preserving score bits, id); sorted by score desc, id asc.
struct Cand {
    uint32_t ord;  // 0 == empty
    int64_t id;
};
__device__ __forceinline__ bool cand_before(const Cand& a, const Cand& b) {  // a ranks strictly before b
    return a.ord > b.ord || (a.ord == b.ord && a.id < b.id);
}

__global__ void __launch_bounds__(kSelThreads)
merge_candidates_kernel(const float* __restrict__ scores, const int64_t* __restrict__ ids, int W, int B, int k,
                        int n_pow2, float* __restrict__ out_scores, int64_t* __restrict__ out_ids, int shard_shift,
                        int64_t id_base) {
    extern __shared__ __align__(16) unsigned char raw[];
    int64_t* sid = reinterpret_cast<int64_t*>(raw);
    uint32_t* sord = reinterpret_cast<uint32_t*>(sid + n_pow2);
    const int b = blockIdx.x;
    const int total = W * k;
    for (int i = threadIdx.x; i < n_pow2; i += kSelThreads) {
        uint32_t ord = 0;
        int64_t id = INT64_MAX;
        if (i < total) {
            int w = i / k, j = i - w * k;
            size_t src = ((size_t)w * B + b) * k + j;
            int64_t v = ids[src];
            if (v >= 0) {
                if (shard_shift >= 0)  // local row of part w -> global row (block-cyclic, store.h:ShardMap)
                    v = ((((v >> shard_shift) * W + w)) << shard_shift) | (v & ((1ll << shard_shift) - 1));
                id = v;
                ord = (uint32_t)(pack_key(scores[src], 0) >> 32);
            }
        }
        sord[i] = ord;
        sid[i] = id;
    }
    for (int size = 2; size <= n_pow2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (n_pow2 >> 1); t += kSelThreads) {
                int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                bool desc = ((lo & size) == 0);
                Cand a{sord[lo], sid[lo]}, c{sord[hi], sid[hi]};
                bool swap = desc ? cand_before(c, a) : cand_before(a, c);
                if (swap) {
                    sord[lo] = c.ord; sid[lo] = c.id;
                    sord[hi] = a.ord; sid[hi] = a.id;
                }
            }
        }
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += kSelThreads) {
        size_t o = (size_t)b * k + j;
        bool ok = j < n_pow2 && sord[j] != 0;
        out_scores[o] = ok ? key_score((uint64_t)sord[j] << 32) : -INFINITY;
        out_ids[o] = ok ? sid[j] + id_base : -1;
    }
}

}  // namespace

void launch_merge_keys(const uint64_t* keys_in, int B, int m, int k, int64_t id_base, float* out_scores,
                       int64_t* out_ids, cudaStream_t st) {
    if (B == 0) return;
    if (k <= 32 && m <= 24 * kSelThreads) {  // register-resident selection: up to 24 keys per thread
        if (m <= 8 * kSelThreads)
            merge_keys_small_kernel<8><<<B, kSelThreads, 0, st>>>(keys_in, m, k, id_base, out_scores, out_ids);
        else
            merge_keys_small_kernel<24><<<B, kSelThreads, 0, st>>>(keys_in, m, k, id_base, out_scores, out_ids);
        return;
    }
    int kp = next_pow2(k);
    int buf_n = kp * 2 > 2048 ? kp * 2 : 2048;
    if (next_pow2(kp + m) < buf_n) buf_n = next_pow2(kp + m) > 2 * kp ? next_pow2(kp + m) : 2 * kp;
    size_t smem = (size_t)buf_n * sizeof(uint64_t);
    merge_keys_kernel<<<B, kSelThreads, smem, st>>>(keys_in, m, k, kp, buf_n, id_base, out_scores, out_ids);
}

void launch_merge_fix(const uint64_t* keys_in, int max_fix, int m, int k, int64_t id_base, const uint32_t* fix_list,
                      const uint32_t* fix_n, uint32_t* flags, float* out_scores, int64_t* out_ids, cudaStream_t st,
                      uint32_t* flag_count) {
    int kp = next_pow2(k);
    int buf_n = kp * 2 > 2048 ? kp * 2 : 2048;
    size_t smem = (size_t)buf_n * sizeof(uint64_t);
    merge_fix_kernel<<<max_fix, kSelThreads, smem, st>>>(keys_in, m, k, kp, buf_n, id_base, fix_list, fix_n, flags,
                                                         out_scores, out_ids, flag_count);
}

void launch_merge_candidates(const float* scores, const int64_t* ids, int W, int B, int k, float* out_scores,
                             int64_t* out_ids, cudaStream_t st, int shard_shift, int64_t id_base) {
    if (B == 0) return;
    int n_pow2 = next_pow2(W * k);
    size_t smem = (size_t)n_pow2 * (sizeof(int64_t) + sizeof(uint32_t));
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(merge_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    merge_candidates_kernel<<<B, kSelThreads, smem, st>>>(scores, ids, W, B, k, n_pow2, out_scores, out_ids, shard_shift, id_base);
}

}  // namespace mqvs
