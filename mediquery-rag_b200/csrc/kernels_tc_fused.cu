// kernels_tc_fused.cu -- the tensor-core search as ONE cooperative launch, for batches of up to 256 queries (one or two query tiles):
// in-kernel threshold learning through grid barriers, the corpus scan, selection, exact fp32 rescoring and the margin check.
// This file: the prep kernels, the host side (search_fused, search_band) and the cosine instantiations; the kernel itself and
// its tail live in tc_fused_device.cuh (also instantiated for the l2 / ip spaces by kernels_tc_fused_space.cu), the device
// building blocks (tile epilogue, chunk maxima, PTX wrappers) shared with kernels_tc.cu in tc_device.cuh.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "tc_fused_device.cuh"

namespace mqvs {
namespace tc {

// the l2 / ip instantiations live in kernels_tc_fused_space.cu
extern template cudaError_t launch_fused_any<true>(int, int, int, const CUtensorMap&, const CUtensorMap&, const FusedArgs&, int,
                                                   cudaStream_t);

namespace {
Trace& g_trace = trace();
}

// Raw queries -> everything tc_fused_kernel needs, in one launch: K1's normalisation (bit-identical: same lane shares,
// same order), the fp32 copy the exact rescoring reads, the tensor-core operand matrix (bf16-rounded or fp32) padded to
// 128 rows, eps[q], tau = -inf, and the zeroed grid-barrier counters / flag count.
template <typename T>
__global__ void __launch_bounds__(128)
prep_fused_kernel(const float* __restrict__ q_raw, int dim, int vec4, int B, float* __restrict__ qn, T* __restrict__ q_op,
                  float* __restrict__ eps, uint32_t* __restrict__ tau, unsigned int* __restrict__ gbar,
                  uint32_t* __restrict__ flag_count, unsigned int* __restrict__ bad_epoch, unsigned int epoch,
                  int metric, float amax, float* __restrict__ hq, float* __restrict__ qnorm,
                  const unsigned int* __restrict__ shadow_err) {
    const int lane = threadIdx.x & 31, q = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        gbar[0] = gbar[1] = gbar[2] = 0u;
        *flag_count = 0u;
    }
    float4 v[kChunks];
    float nrm = 0.0f;
    if (q < B) {
        load_lane_share(q_raw + (size_t)q * dim, dim, vec4 != 0, lane, v);
        if (!normalize_lane_share(v, &nrm) && lane == 0) atomicMax(bad_epoch, epoch);
    } else {
#pragma unroll
        for (int c = 0; c < kChunks; ++c) v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float err2 = 0.f;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        const size_t o = (size_t)q * kDim + c * 128 + lane * 4;
        if (q < B) *reinterpret_cast<float4*>(qn + o) = v[c];
        if constexpr (sizeof(T) == 2) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(v[c].x), h1 = __float2bfloat16_rn(v[c].y);
            const __nv_bfloat16 h2 = __float2bfloat16_rn(v[c].z), h3 = __float2bfloat16_rn(v[c].w);
            const float d0 = v[c].x - __bfloat162float(h0), d1 = v[c].y - __bfloat162float(h1);
            const float d2 = v[c].z - __bfloat162float(h2), d3 = v[c].w - __bfloat162float(h3);
            err2 += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
            uint2 pk;
            pk.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            pk.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
            *reinterpret_cast<uint2*>(q_op + o) = pk;
        } else {
            *reinterpret_cast<float4*>(q_op + o) = v[c];
        }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) err2 += __shfl_xor_sync(0xffffffffu, err2, m);
    if (lane == 0) {
        // same bounds as prep_queries_kernel
        float e = sizeof(T) == 2 ? sqrtf(err2) * 1.004f * 1.0001f + 1e-4f : 2.0f * 0.0009765625f * 1.001f + 1e-4f;
        if (sizeof(T) == 2 && shadow_err != nullptr) e += __uint_as_float(__ldg(shadow_err)) * 1.0002f;  // rounded corpus operand
        if (metric != MQVS_METRIC_COSINE) {
            // ranking score t = a s - a^2 h: an error of e in the cosine is at most amax e in t, plus the fp32 rounding of
            // the transform itself (three roundings of terms no larger than amax + amax^2 h)
            const float h = metric == MQVS_METRIC_L2 ? __fdiv_rn(0.5f, fmaxf(nrm, kEps)) : 0.0f;
            hq[q] = q < B ? h : 0.0f;
            if (q < B) qnorm[q] = nrm;
            e = amax * e * 1.0001f + 1e-6f * (amax + amax * amax * h);
        }
        if (q < B) eps[q] = e;
        tau[q] = kOrdNegInf;
    }
}

// ---- band pass, launch 1 of 2: compact the flagged queries (up to 128 per search, in index order; the exact scan takes
// what is left) and build their operand matrix: row j = (already normalised) query band_list[j], bf16-rounded or fp32,
// with its threshold tau[j] = band[query]; unused rows are zero with a threshold nothing reaches.  Every block derives
// the same list on its own (one warp, ballot + popcount), so no second launch or grid barrier is needed; block 0
// publishes it for tc_fused_kernel.
template <typename T>
__global__ void __launch_bounds__(128)
prep_band_kernel(const float* __restrict__ qn, const uint32_t* __restrict__ flags, int B, uint32_t* __restrict__ band_list,
                 uint32_t* __restrict__ band_n, const uint32_t* __restrict__ band, T* __restrict__ q_op,
                 uint32_t* __restrict__ tau, unsigned int* __restrict__ gbar, const float* __restrict__ hq,
                 float* __restrict__ h_op, const uint32_t* __restrict__ flag_count) {
    __shared__ uint32_t list_s[BM];
    __shared__ uint32_t n_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, j = blockIdx.x * 4 + warp;
    // nothing flagged: one word says so (walking B flags with one warp is 0.2 ms at B = 16384; see scan_fix_kernel)
    if (__ldcg(flag_count) == 0u) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            gbar[0] = gbar[1] = gbar[2] = 0u;
            *band_n = 0u;
        }
        return;
    }
    if (warp == 0) {
        uint32_t base = 0;
        for (int c = 0; c < B && base < (uint32_t)BM; c += 32) {
            const bool f = c + lane < B && flags[c + lane] != 0u;
            const unsigned m = __ballot_sync(0xffffffffu, f);
            const uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
            if (f && pos < (uint32_t)BM) list_s[pos] = (uint32_t)(c + lane);
            base += __popc(m);
        }
        if (lane == 0) n_s = base < (uint32_t)BM ? base : (uint32_t)BM;
    }
    __syncthreads();
    const uint32_t n = n_s;
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            gbar[0] = gbar[1] = gbar[2] = 0u;
            *band_n = n;
        }
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) band_list[i] = list_s[i];
    }
    if (n == 0) return;
    const bool live = (uint32_t)j < n;
    const uint32_t qi = live ? list_s[j] : 0u;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        const float4 v = live ? *reinterpret_cast<const float4*>(qn + (size_t)qi * kDim + c * 128 + lane * 4)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
        const size_t o = (size_t)j * kDim + c * 128 + lane * 4;
        if constexpr (sizeof(T) == 2) {
            uint2 pk;
            pk.x = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v.x)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v.y)) << 16);
            pk.y = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v.z)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v.w)) << 16);
            *reinterpret_cast<uint2*>(q_op + o) = pk;
        } else {
            *reinterpret_cast<float4*>(q_op + o) = v;
        }
    }
    if (lane == 0) {
        tau[j] = live ? band[qi] : 0xFF800000u;  // ord(+inf): no score reaches it
        if (hq != nullptr) h_op[j] = live ? hq[qi] : 0.0f;
    }
}

bool fused_ok(int B, int k, int sm_count) {
    // one or two query tiles (the two TMEM accumulators); one list / 8 chunk maxima per CTA must fit
    // select_lists_body<192> and kth_largest_u32 (1280 values); CTA c finishes queries c and c + G, so 2 G >= B
    // (k' = 1152 must not exceed the sm_count * 8 chunk maxima a threshold is the k'-th largest of)
    return B >= 1 && B <= 2 * BM && k <= MQVS_MAX_K && sm_count >= BM && sm_count <= 160 && (k <= 256 || sm_count * 8 >= 1152);
}
size_t cmax_bytes(const Plan& p) { return (size_t)p.b_pad * p.grid * 8 * sizeof(uint32_t); }

namespace {
cudaError_t launch_fused_space(int dtype, int kprime, int query_tiles, const CUtensorMap& mq, const CUtensorMap& mx,
                               const FusedArgs& a, int grid, cudaStream_t st) {
    return a.norms != nullptr ? launch_fused_any<true>(dtype, kprime, query_tiles, mq, mx, a, grid, st)
                              : launch_fused_any<false>(dtype, kprime, query_tiles, mq, mx, a, grid, st);
}
}  // namespace

const char* search_fused(const Buffers& b, const Plan& p, const void* rows, int dtype, uint32_t n_rows, const float* q_raw,
                         int dim, float* qn, int B, int k, const uint32_t* mask, int64_t id_base, float* out_scores,
                         int64_t* out_ids, unsigned int* bad_epoch, unsigned int epoch, cudaStream_t st, bool* coop_refused,
                         const Space& space) {
    CUtensorMap mq, mx;
    *coop_refused = false;
    const int rows_q = p.b_pad;  // 128 or 256 operand rows
    // up to 64 queries: half-height operand tiles (M = 64) -- half the query-operand traffic per corpus chunk
    static const bool no_half = getenv("MQVS_NO_HALF_TILE") != nullptr;  // developer knob: measure what it buys
    const bool half = B <= 64 && !no_half;
    if (!make_map(&mq, b.q_op, (uint64_t)rows_q, dtype, half ? 64 : BM)) return "cuTensorMapEncodeTiled failed for the query matrix";
    if (!make_map(&mx, rows, (uint64_t)n_rows, dtype, BN)) return "cuTensorMapEncodeTiled failed for the corpus";
    g_trace.mark("begin", st);
    const int vec4 = (dim % 4 == 0 && (reinterpret_cast<uintptr_t>(q_raw) & 15) == 0) ? 1 : 0;
    if (dtype == MQVS_F32)
        prep_fused_kernel<float><<<rows_q / 4, 128, 0, st>>>(q_raw, dim, vec4, B, qn, (float*)b.q_op, b.eps, b.tau, b.gbar,
                                                         b.flag_count, bad_epoch, epoch, space.metric, space.amax, b.hq,
                                                         space.qnorm, nullptr);
    else
        prep_fused_kernel<__nv_bfloat16><<<rows_q / 4, 128, 0, st>>>(q_raw, dim, vec4, B, qn, (__nv_bfloat16*)b.q_op, b.eps, b.tau,
                                                                 b.gbar, b.flag_count, bad_epoch, epoch, space.metric,
                                                                 space.amax, b.hq, space.qnorm, space.exact_rows ? space.shadow_err : nullptr);
    g_trace.mark("prep", st);
    FusedArgs a{};
    a.n_rows = n_rows;
    a.n_queries = (uint32_t)B;
    a.n_tiles = (n_rows + BN - 1) / BN;
    a.ncols = (uint32_t)p.grid * 8;
    a.mask = mask;
    a.tau = b.tau;
    a.cmax = b.cmax;
    a.cnt = b.cnt;
    a.lists = b.lists;
    a.gbar = b.gbar;
    a.rows = rows;
    a.rows_f32 = dtype == MQVS_BF16 ? space.exact_rows : nullptr;
    a.qn = qn;
    a.eps = b.eps;
    a.k = k;
    a.id_base = id_base;
    a.out_scores = out_scores;
    a.out_ids = out_ids;
    a.flags = b.flags;
    a.flag_count = b.flag_count;
    a.band = b.band;
    a.band_list = nullptr;
    a.band_n = nullptr;
    a.counters = nullptr;
    a.metric = space.metric;
    a.norms = space.metric != MQVS_METRIC_COSINE ? space.norms : nullptr;
    a.nbounds = reinterpret_cast<const float2*>(space.norm_bounds);
    a.hq = b.hq;
    a.h_op = b.hq;  // operand row q is query q
    a.dbg = g_trace.on ? g_trace.fused_dbg(p.grid) : nullptr;
    if (a.dbg) cudaMemsetAsync(a.dbg, 0, 256 * 16 * sizeof(unsigned long long), st);
    cudaError_t e = launch_fused_space(dtype, p.kprime, half ? 0 : (rows_q <= BM ? 1 : 2), mq, mx, a, p.grid, st);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) {
        cudaGetLastError();
        *coop_refused = true;
    }
    if (e != cudaSuccess) return cudaGetErrorString(e);
    g_trace.mark("fused", st);
    e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

// Band pass = the cheap repair.  A query is flagged when its k-th exact score does not clear tau + eps, i.e. when rows
// outside the rescored shortlist could still beat it (dense neighbourhoods: many scores within eps of each other -- the
// normal case for a corpus of near-duplicate chunks).  Instead of one exact scan PER flagged query, all flagged queries
// (up to 128) go through the single-launch tensor pipeline once more with GIVEN thresholds: band[q] = (k-th exact score
// of the first pass) - eps.  Every row that can reach the k-th score has an approximate score above that, so collecting
// and rescoring ALL rows above it gives the exact answer; no threshold is learnt, nothing is compared with a margin.
// Only a query whose band holds more than 1024 rows (or overflows a per-CTA list) stays flagged for the exact scan.
// Measured on 1M clustered bf16 rows (2000-member clusters at cosine 0.98, 64 of 64 queries flagged): see DESIGN.md.
const char* search_band(const Buffers& b, const Plan& p, const void* rows, int dtype, uint32_t n_rows, const float* qn, int B,
                        int k, const uint32_t* mask, int64_t id_base, float* out_scores, int64_t* out_ids,
                        unsigned long long* counters, cudaStream_t st, const Space& space) {
    CUtensorMap mq, mx;
    const float* hq = space.metric != MQVS_METRIC_COSINE ? b.hq : nullptr;
    if (!make_map(&mq, b.q_op, (uint64_t)BM, dtype, BM)) return "cuTensorMapEncodeTiled failed for the query matrix";
    if (!make_map(&mx, rows, (uint64_t)n_rows, dtype, BN)) return "cuTensorMapEncodeTiled failed for the corpus";
    if (dtype == MQVS_F32)
        prep_band_kernel<float><<<BM / 4, 128, 0, st>>>(qn, b.flags, B, b.band_list, b.band_n, b.band, (float*)b.q_op, b.tau, b.gbar,
                                                        hq, b.h_op, b.flag_count);
    else
        prep_band_kernel<__nv_bfloat16><<<BM / 4, 128, 0, st>>>(qn, b.flags, B, b.band_list, b.band_n, b.band, (__nv_bfloat16*)b.q_op,
                                                                 b.tau, b.gbar, hq, b.h_op, b.flag_count);
    FusedArgs a{};
    a.n_rows = n_rows;
    a.n_queries = 0;  // read from band_n on the device
    a.n_tiles = (n_rows + BN - 1) / BN;
    a.ncols = (uint32_t)p.grid * 8;
    a.mask = mask;
    a.tau = b.tau;
    a.cmax = b.cmax;
    a.cnt = b.cnt;
    a.lists = b.lists;
    a.gbar = b.gbar;
    a.rows = rows;
    a.rows_f32 = dtype == MQVS_BF16 ? space.exact_rows : nullptr;
    a.qn = qn;
    a.eps = b.eps;
    a.k = k;
    a.id_base = id_base;
    a.out_scores = out_scores;
    a.out_ids = out_ids;
    a.flags = b.flags;
    a.flag_count = b.flag_count;
    a.dbg = nullptr;
    a.band = b.band;
    a.band_list = b.band_list;
    a.band_n = b.band_n;
    a.counters = counters;
    a.metric = space.metric;
    a.norms = space.metric != MQVS_METRIC_COSINE ? space.norms : nullptr;
    a.nbounds = reinterpret_cast<const float2*>(space.norm_bounds);
    a.hq = b.hq;
    a.h_op = b.h_op;
    cudaError_t e = launch_fused_space(dtype, p.kprime, 1, mq, mx, a, p.grid, st);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return cudaGetErrorString(e);
    }
    g_trace.mark("band pass", st);
    return nullptr;
}

}  // namespace tc
}  // namespace mqvs
