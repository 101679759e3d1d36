// kernels_ingest.cu -- K1: fused L2-normalise + fp32/bf16 pack (one warp per row), and the row gather.
//
// Replaces the work hidden behind collection.upsert(embeddings=...) in Chroma.from_documents
// (/root/reference/src/ingest_medical.py:106-110): rows arrive as fp32 [n, dim], leave as unit rows of
// 768 elements (zero-padded) in the store's dtype.  HBM-bound: reads 4*dim, writes 768*sizeof(elem) per row.
#include "common.cuh"
#include "launch.h"

namespace mqvs {

namespace {

constexpr int kIngestWarps = 8;

struct MemSource {
    const float* x;
    int dim;
    bool vec4;  // dim % 4 == 0 and base 16-byte aligned
    __device__ __forceinline__ void load(uint64_t row, int lane, float4 (&v)[kChunks]) const {
        load_lane_share(x + row * (uint64_t)dim, dim, vec4, lane, v);
    }
};

__device__ __forceinline__ uint64_t splitmix64(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float synth_value(uint64_t seed, uint64_t idx) {
    return __fsub_rn(__fmul_rn((float)(splitmix64(seed, idx) >> 40), 1.0f / 8388608.0f), 1.0f);
}

struct SynthSource {
    uint64_t seed, first_index;
    int dim;
    __device__ __forceinline__ void load(uint64_t row, int lane, float4 (&v)[kChunks]) const {
        uint64_t base = (first_index + row) * (uint64_t)dim;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            int j = c * 128 + lane * 4;
            v[c].x = j + 0 < dim ? synth_value(seed, base + j + 0) : 0.0f;
            v[c].y = j + 1 < dim ? synth_value(seed, base + j + 1) : 0.0f;
            v[c].z = j + 2 < dim ? synth_value(seed, base + j + 2) : 0.0f;
            v[c].w = j + 3 < dim ? synth_value(seed, base + j + 3) : 0.0f;
        }
    }
};

template <typename Source, typename OutT>
__global__ void __launch_bounds__(kIngestWarps * 32)
normalize_pack_kernel(Source src, uint64_t n, OutT* __restrict__ out, unsigned long long* __restrict__ bad,
                      float* __restrict__ norms, unsigned int* __restrict__ max_dev) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps_total = (uint64_t)gridDim.x * kIngestWarps;
    float dev_max = 0.0f, nrm_max = 0.0f;  // largest | ||x|| - 1 | and largest ||x|| this warp has seen
    for (uint64_t row = (uint64_t)blockIdx.x * kIngestWarps + (threadIdx.x >> 5); row < n; row += warps_total) {
        float4 v[kChunks];
        src.load(row, lane, v);
        float nrm;
        const bool ok = normalize_lane_share(v, &nrm);
        if (!ok && lane == 0) atomicAdd(bad, 1ull);
        if (norms != nullptr && lane == 0) norms[row] = nrm;
        dev_max = fmaxf(dev_max, fabsf(nrm - 1.0f));
        nrm_max = fmaxf(nrm_max, nrm);
        OutT* o = out + row * (uint64_t)kDim;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            const float4 y = v[c];
            if constexpr (sizeof(OutT) == 4) {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(o) + c * 128 + lane * 4) = y;
            } else {
                __nv_bfloat162 a = __floats2bfloat162_rn(y.x, y.y);
                __nv_bfloat162 b = __floats2bfloat162_rn(y.z, y.w);
                uint2 w;
                w.x = *reinterpret_cast<uint32_t*>(&a);
                w.y = *reinterpret_cast<uint32_t*>(&b);
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(o) + c * 128 + lane * 4) = w;
            }
        }
    }
    // non-negative floats order like their bit patterns: one atomicMax per warp per launch
    if (max_dev != nullptr && lane == 0) {
        if (dev_max > 0.0f) atomicMax(max_dev, __float_as_uint(dev_max));
        if (nrm_max > 0.0f) atomicMax(max_dev + 1, __float_as_uint(nrm_max));  // [1]: largest ||x||
    }
}

template <typename Source>
void launch_k1(Source src, uint64_t n, int out_dtype, void* out, unsigned long long* bad, float* norms,
               unsigned int* max_dev, cudaStream_t st) {
    if (n == 0) return;
    uint64_t blocks = (n + kIngestWarps - 1) / kIngestWarps;
    int grid = (int)(blocks < 148ull * 16 ? blocks : 148ull * 16);
    if (out_dtype == MQVS_F32)
        normalize_pack_kernel<Source, float><<<grid, kIngestWarps * 32, 0, st>>>(src, n, (float*)out, bad, norms, max_dev);
    else
        normalize_pack_kernel<Source, __nv_bfloat16><<<grid, kIngestWarps * 32, 0, st>>>(src, n, (__nv_bfloat16*)out, bad,
                                                                                         norms, max_dev);
}

// one warp per 32-row chunk: (max, min) of its rows' norms
__global__ void norm_bounds_kernel(const float* __restrict__ norms, uint64_t n_rows, uint64_t first_chunk, uint64_t n_chunks,
                                   float2* __restrict__ bounds) {
    const int lane = threadIdx.x & 31;
    const uint64_t chunk = first_chunk + (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (chunk >= n_chunks) return;
    const uint64_t r = chunk * 32 + lane;
    const bool in = r < n_rows;
    float hi = in ? norms[r] : 0.0f, lo = in ? norms[r] : INFINITY;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, m));
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, m));
    }
    if (lane == 0) bounds[chunk] = make_float2(hi, lo);
}

// bf16 scoring shadow of an fp32 store (store.h: mqvs_store::shadow): one warp per row, 16-byte loads / 8-byte stores; the
// largest ||x - bf16(x)|| over the rows converted (what the tensor path's error bound adds for the rounded corpus operand)
// is raised into *err_bits -- non-negative floats order like their bit patterns.
__global__ void __launch_bounds__(kIngestWarps * 32)
shadow_build_kernel(const float* __restrict__ rows, uint64_t first, uint64_t n, __nv_bfloat16* __restrict__ shadow,
                    unsigned int* __restrict__ err_bits) {
    const int lane = threadIdx.x & 31;
    const uint64_t warps_total = (uint64_t)gridDim.x * kIngestWarps;
    float err_max = 0.0f;
    for (uint64_t i = (uint64_t)blockIdx.x * kIngestWarps + (threadIdx.x >> 5); i < n; i += warps_total) {
        const float* src = rows + (first + i) * (uint64_t)kDim;
        __nv_bfloat16* dst = shadow + (first + i) * (uint64_t)kDim;
        float err2 = 0.0f;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
            const float4 y = ld_stream_f4(src + c * 128 + lane * 4);
            const __nv_bfloat162 a = __floats2bfloat162_rn(y.x, y.y), b = __floats2bfloat162_rn(y.z, y.w);
            const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
            const float d0 = y.x - fa.x, d1 = y.y - fa.y, d2 = y.z - fb.x, d3 = y.w - fb.y;
            err2 += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
            uint2 w;
            w.x = *reinterpret_cast<const uint32_t*>(&a);
            w.y = *reinterpret_cast<const uint32_t*>(&b);
            *reinterpret_cast<uint2*>(dst + c * 128 + lane * 4) = w;
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) err2 += __shfl_xor_sync(0xffffffffu, err2, m);
        err_max = fmaxf(err_max, err2);
    }
    // sqrt rounded up a little: the bound must not be smaller than the true norm of the rounding error
    if (lane == 0 && err_max > 0.0f) atomicMax(err_bits, __float_as_uint(sqrtf(err_max) * 1.0001f));
}

template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ rows, const int64_t* __restrict__ ids, uint64_t n, int dim,
                                   float* __restrict__ out) {
    uint64_t i = blockIdx.x;
    if (i >= n) return;
    const T* r = rows + (uint64_t)ids[i] * kDim;
    for (int j = threadIdx.x; j < dim; j += blockDim.x) out[i * (uint64_t)dim + j] = (float)r[j];
}

// one warp per row: 16-byte vectors, streaming loads
__global__ void __launch_bounds__(256)
gather_raw_rows_kernel(const uint4* __restrict__ rows, uint32_t vec_per_row, const uint32_t* __restrict__ ids, uint64_t n,
                       uint4* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= n) return;
    const uint4* src = rows + (uint64_t)ids[i] * vec_per_row;
    uint4* dst = out + i * vec_per_row;
    for (uint32_t v = threadIdx.x & 31; v < vec_per_row; v += 32) dst[v] = __ldcs(src + v);
}

}  // namespace

void launch_gather_raw_rows(const void* rows, size_t row_bytes, const uint32_t* ids, uint64_t n, void* out, cudaStream_t st) {
    if (n == 0) return;
    gather_raw_rows_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>((const uint4*)rows, (uint32_t)(row_bytes / 16), ids, n, (uint4*)out);
}

void launch_normalize(const float* x, uint64_t n, int dim, int out_dtype, void* out, unsigned long long* bad,
                      cudaStream_t st, float* norms, unsigned int* max_dev) {
    MemSource src{x, dim, (dim % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0)};
    launch_k1(src, n, out_dtype, out, bad, norms, max_dev, st);
}

void launch_synth_normalize(uint64_t seed, uint64_t first_index, uint64_t n, int dim, int out_dtype, void* out,
                            unsigned long long* bad, cudaStream_t st, float* norms, unsigned int* max_dev) {
    SynthSource src{seed, first_index, dim};
    launch_k1(src, n, out_dtype, out, bad, norms, max_dev, st);
}

void launch_norm_bounds(const float* norms, uint64_t n_rows, uint64_t first_chunk, float* bounds, cudaStream_t st) {
    const uint64_t n_chunks = (n_rows + 31) / 32;
    if (first_chunk >= n_chunks) return;
    const uint64_t blocks = (n_chunks - first_chunk + 7) / 8;
    norm_bounds_kernel<<<(unsigned)blocks, 256, 0, st>>>(norms, n_rows, first_chunk, n_chunks, reinterpret_cast<float2*>(bounds));
}

void launch_shadow_build(const float* rows, uint64_t first, uint64_t n, void* shadow, unsigned int* err_bits, cudaStream_t st) {
    if (n == 0) return;
    const uint64_t blocks = (n + kIngestWarps - 1) / kIngestWarps;
    const int grid = (int)(blocks < 148ull * 16 ? blocks : 148ull * 16);
    shadow_build_kernel<<<grid, kIngestWarps * 32, 0, st>>>(rows, first, n, (__nv_bfloat16*)shadow, err_bits);
}

void launch_gather_rows(const void* rows, int dtype, const int64_t* ids, uint64_t n, int dim, float* out,
                        cudaStream_t st) {
    if (n == 0) return;
    if (dtype == MQVS_F32)
        gather_rows_kernel<float><<<(unsigned)n, 128, 0, st>>>((const float*)rows, ids, n, dim, out);
    else
        gather_rows_kernel<__nv_bfloat16><<<(unsigned)n, 128, 0, st>>>((const __nv_bfloat16*)rows, ids, n, dim, out);
}

}  // namespace mqvs
