// common.cuh -- shared device helpers for libmqvs (sm_100a).
//
// Canonical fp32 order (the definition of "exact fp32 score" in this backend; oracle/canonical.c restates it):
// a row of MQVS_DIM_PAD = 768 elements is owned by one warp; lane l owns elements j with (j mod 128) div 4 == l
// (one 16-byte fp32 vector per 512-byte chunk), accumulates them in ascending j with fused multiply-add, and
// the 32 partial sums are combined by a 5-level xor butterfly (16, 8, 4, 2, 1).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mqvs.h"

namespace mqvs {

constexpr int kDim = MQVS_DIM_PAD;   // padded row width (elements)
constexpr int kChunks = kDim / 128;  // 16-byte fp32 vectors per lane per row (6)
constexpr int kLaneElems = kChunks * 4;
constexpr float kEps = 1e-12f;

// ---------------------------------------------------------------------------------------------- keys
// Order-preserving 64-bit key: larger == better (score descending, then LOWER id).  0 == empty slot.
__host__ __device__ __forceinline__ uint64_t pack_key(float s, uint32_t id) {
    if (s == 0.0f) s = 0.0f;  // -0 -> +0
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(s);
#else
    uint32_t u;
    memcpy(&u, &s, 4);
#endif
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((uint64_t)u << 32) | (uint64_t)(0xffffffffu - id);
}
__device__ __forceinline__ float key_score(uint64_t key) {
    uint32_t u = (uint32_t)(key >> 32);
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    return __uint_as_float(u);
}
__device__ __forceinline__ uint32_t key_id(uint64_t key) { return 0xffffffffu - (uint32_t)key; }

// ---------------------------------------------------------------------------------------------- warp
__device__ __forceinline__ float warp_butterfly_sum(float v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, m));
    return v;
}
__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        uint64_t o = __shfl_xor_sync(0xffffffffu, v, m);
        v = o < v ? o : v;
    }
    return v;
}

// ---------------------------------------------------------------------------------------------- loads
// Streaming 128-bit / 64-bit loads that bypass L1 allocation (each corpus byte is used once per pass).
__device__ __forceinline__ float4 ld_stream_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_stream_u2(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
// 4 consecutive bf16 (as two 32-bit words) -> 4 floats, exact
__device__ __forceinline__ float4 bf16x4_to_f4(uint2 w) {
    float4 r;
    r.x = __uint_as_float(w.x << 16);
    r.y = __uint_as_float(w.x & 0xffff0000u);
    r.z = __uint_as_float(w.y << 16);
    r.w = __uint_as_float(w.y & 0xffff0000u);
    return r;
}

// One lane's share of a stored row, as fp32 (kChunks float4s), for either storage type.
template <typename T>
struct RowFrag;
template <>
struct RowFrag<float> {
    float4 v[kChunks];
    __device__ __forceinline__ void load(const float* row, int lane) {
#pragma unroll
        for (int c = 0; c < kChunks; ++c) v[c] = ld_stream_f4(row + c * 128 + lane * 4);
    }
    __device__ __forceinline__ float4 get(int c) const { return v[c]; }
};
template <>
struct RowFrag<__nv_bfloat16> {
    uint2 v[kChunks];
    __device__ __forceinline__ void load(const __nv_bfloat16* row, int lane) {
#pragma unroll
        for (int c = 0; c < kChunks; ++c) v[c] = ld_stream_u2(row + c * 128 + lane * 4);
    }
    __device__ __forceinline__ float4 get(int c) const { return bf16x4_to_f4(v[c]); }
};

// lane-partial of the canonical dot: acc over (c, e) ascending with FMA
__device__ __forceinline__ float lane_dot(const float4 (&q)[kChunks], const float4 (&x)[kChunks]) {
    float acc = 0.0f;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        acc = __fmaf_rn(q[c].x, x[c].x, acc);
        acc = __fmaf_rn(q[c].y, x[c].y, acc);
        acc = __fmaf_rn(q[c].z, x[c].z, acc);
        acc = __fmaf_rn(q[c].w, x[c].w, acc);
    }
    return acc;
}

// One lane's canonical share of an fp32 vector of `dim` <= kDim elements (zeros beyond dim).  vec4: dim % 4 == 0 and
// the vector is 16-byte aligned, so the share can be fetched as 128-bit loads.
__device__ __forceinline__ void load_lane_share(const float* p, int dim, bool vec4, int lane, float4 (&v)[kChunks]) {
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        const int j = c * 128 + lane * 4;
        if (vec4 && j + 3 < dim) {
            v[c] = ld_stream_f4(p + j);
        } else {
            v[c].x = j + 0 < dim ? p[j + 0] : 0.0f;
            v[c].y = j + 1 < dim ? p[j + 1] : 0.0f;
            v[c].z = j + 2 < dim ? p[j + 2] : 0.0f;
            v[c].w = j + 3 < dim ? p[j + 3] : 0.0f;
        }
    }
}
// K1's arithmetic on a lane share (whole warp participates): v <- v / max(||v||, eps) in the canonical order.
// Returns false (and zeroes v) when the squared norm is not finite.
// *norm_out (optional) receives ||v|| as K1 computed it (0 for a non-finite row).
__device__ __forceinline__ bool normalize_lane_share(float4 (&v)[kChunks], float* norm_out = nullptr) {
    const float ss = warp_butterfly_sum(lane_dot(v, v));
    const bool ok = isfinite(ss);
    const float nrm = ok ? __fsqrt_rn(ss) : 0.0f;
    if (norm_out != nullptr) *norm_out = nrm;
    const float inv = ok ? __fdiv_rn(1.0f, fmaxf(nrm, kEps)) : 0.0f;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        v[c].x = ok ? __fmul_rn(v[c].x, inv) : 0.0f;
        v[c].y = ok ? __fmul_rn(v[c].y, inv) : 0.0f;
        v[c].z = ok ? __fmul_rn(v[c].z, inv) : 0.0f;
        v[c].w = ok ? __fmul_rn(v[c].w, inv) : 0.0f;
    }
    return ok;
}

// ---------------------------------------------------------------------------------------------- metrics
// Ranking score of a stored row under the collection's space (mqvs_set_metric; oracle/canonical.c:orc_rank_score
// restates it operation by operation).  s = canonical cosine of the normalised query and the normalised row,
// a = ||x|| of the raw row as K1 measured it, h = 0.5 / max(||q||, eps):
//   cosine: s          ip: a*s  (= <q, x> / ||q||)          l2: a*s - a*a*h  (= (||q||^2 - ||q - x||^2) / (2 ||q||))
// Larger is better in every space, so the top-k machinery is the same.
__device__ __forceinline__ float rank_score(int metric, float s, float a, float h) {
    if (metric == MQVS_METRIC_IP) return __fmul_rn(a, s);
    if (metric == MQVS_METRIC_L2) return __fmaf_rn(a, s, -__fmul_rn(__fmul_rn(a, a), h));
    return s;
}

// ---------------------------------------------------------------------------------------------- sort
// Block-wide bitonic sort of n (power of two) 64-bit keys in shared memory, DESCENDING.
__device__ __forceinline__ void block_bitonic_sort_desc(uint64_t* s, int n) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);
                uint64_t a = s[lo], b = s[hi];
                if ((a < b) == desc) {
                    s[lo] = b;
                    s[hi] = a;
                }
            }
        }
    }
    __syncthreads();
}

// Sort-free top-k for small k (<= 32) by a 256-thread block: each of the 8 warps repeatedly extracts the maximum of its
// slice of the m <= kPerLane * 256 candidate keys (register-resident, warp-shuffle max), then warp 0 does the same over
// the 8 x k winners.  `win` is 8 * 32 keys of shared memory.  Keys are unique (they embed the row id) or 0 (= empty).
template <int kPerLane>
__device__ __forceinline__ uint64_t warp_pop_max(uint64_t (&ks)[kPerLane]) {
    uint64_t best = ks[0];
#pragma unroll
    for (int r = 1; r < kPerLane; ++r) best = ks[r] > best ? ks[r] : best;
    uint64_t top = best;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const uint64_t other = __shfl_xor_sync(0xffffffffu, top, o);
        top = other > top ? other : top;
    }
    if (top != 0ull && best == top) {  // exactly one lane owns it
        bool done = false;
#pragma unroll
        for (int r = 0; r < kPerLane; ++r)
            if (!done && ks[r] == top) { ks[r] = 0ull; done = true; }
    }
    return top;
}
template <int kPerLane>
__device__ __forceinline__ void block_select_small(const uint64_t* __restrict__ in, int m, int k, int64_t id_base,
                                                   float* __restrict__ out_scores, int64_t* __restrict__ out_ids,
                                                   uint64_t* win) {
    static_assert(kPerLane >= 8, "the final round holds 8 x k winners");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t key[kPerLane];
#pragma unroll
    for (int r = 0; r < kPerLane; ++r) {
        const int i = (r * 8 + warp) * 32 + lane;  // coalesced, warps interleaved
        key[r] = i < m ? __ldcg(in + i) : 0ull;
    }
    for (int j = 0; j < k; ++j) {
        const uint64_t top = warp_pop_max<kPerLane>(key);
        if (lane == 0) win[warp * 32 + j] = top;
    }
    __syncthreads();
    if (warp == 0) {
        uint64_t fin[kPerLane];
#pragma unroll
        for (int r = 0; r < kPerLane; ++r) fin[r] = (r < 8 && lane < k) ? win[r * 32 + lane] : 0ull;
        for (int j = 0; j < k; ++j) {
            const uint64_t top = warp_pop_max<kPerLane>(fin);
            if (lane == 0) {
                out_scores[j] = top ? key_score(top) : -INFINITY;
                out_ids[j] = top ? (int64_t)key_id(top) + id_base : -1;
            }
        }
    }
    __syncthreads();
}

__host__ __device__ __forceinline__ int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace mqvs
