// tc_fused_device.cuh -- device side of the single-launch tensor pipeline (tc_fused_kernel + its tail) and the dispatch over
// its instantiations.  Included by two translation units so that they build in parallel: kernels_tc_fused.cu (cosine, XF =
// false) and kernels_tc_fused_space.cu (the l2 / ip spaces over raw vectors, XF = true).
#pragma once

#include <type_traits>

#include "launch.h"
#include "tc_device.cuh"

namespace mqvs {
namespace tc {

// ------------------------------------------------------------------------------------------ single-launch pipeline
struct FusedArgs;
template <typename T, int CAP, int KPRIME, bool XF>
__device__ __forceinline__ void fused_tail(const FusedArgs& a, int G, uint8_t* smem_al, int slot, int b_pad);

// Batches of one or two query tiles (up to 256 queries) are HBM-bound: the corpus is streamed once and the tensor pipe idles
// most of the time, so what the segmented pipeline above pays for -- two short cold segments and three selection
// launches, ~200 us next to a 425 us scan at 1M x 768 fp32 -- is pure latency.  tc_fused_kernel does the whole search
// in ONE cooperative launch of one persistent CTA per SM:
//   1. every CTA scores its first tile; the epilogue reads the accumulator once and publishes, per query, the maximum
//      of each 32-row chunk (8 per CTA, masked and out-of-range rows excluded);
//   2. grid barrier; CTA c takes query c and finds the k'-th largest of the G*8 chunk maxima.  The chunks are disjoint,
//      so at least k' allowed rows score >= that value: a valid threshold tau, within a few ranks of the exact k'-th
//      best of the first G*256 rows.  Second grid barrier publishes tau[0..B);
//   3. the first tile -- still in TMEM -- is read again against tau, then the CTA streams its remaining tiles exactly as
//      tc_score_kernel does (the TMA and MMA warps never stop: the second accumulator covers the barriers);
//   4. third grid barrier; CTA c merges the G per-CTA lists of query c, rescoring its k' best rows in the canonical fp32
//      order, sorts, writes the top-k and checks the exactness margin (select_lists + rescore + finalize above).
struct FusedArgs {
    uint32_t n_rows, n_queries, n_tiles, ncols;  // ncols = gridDim.x * 8 chunk maxima per query
    const uint32_t* mask;
    uint32_t* tau;         // [128]: -inf on entry (prep)
    uint32_t* cmax;        // [128][ncols] order-preserving chunk maxima, 0 = no row
    uint32_t* cnt;         // [grid][128]
    uint64_t* lists;       // [grid][128][CAP]
    unsigned int* gbar;    // [3] grid barrier counters, zero on entry (prep)
    const void* rows;      // corpus as the tensor cores stream it (and, unless rows_f32 is set, as the exact rescoring reads it)
    const void* rows_f32;  // bf16 instantiations over the scoring SHADOW of an fp32 store: the stored fp32 rows the exact
                           // rescoring reads instead; nullptr otherwise
    const float* qn;       // [B][768] normalised fp32 queries
    const float* eps;      // [B]
    int k;
    long long id_base;
    float* out_scores;
    int64_t* out_ids;
    uint32_t* flags;
    uint32_t* flag_count;
    unsigned long long* dbg;  // MQVS_TRACE only: [grid][16] globaltimer stamps of the phases, else nullptr
    uint32_t* band;             // [B] out: threshold of the band pass for every query this launch flags
    // band pass only (nullptr / unused in the first launch): the flagged queries, compacted
    const uint32_t* band_list;  // [<= 128] query index of operand row / TMEM lane j
    const uint32_t* band_n;     // device: how many
    unsigned long long* counters;  // [2] += queries the band pass repaired
    // l2 / ip spaces over raw vectors (tc.h:Space); norms == nullptr: cosine
    const float* norms;    // [n_rows] ||x||
    const float2* nbounds; // [ceil(n_rows / 32)] (max, min) norm per 32-row chunk
    const float* hq;       // [B] by query: 0.5 / max(||q||, eps) (l2), 0 (ip)
    const float* h_op;     // the same by operand row (= hq in the first launch, the compacted queries in the band pass)
    int metric;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define MQVS_STAMP(i_) do { if (a.dbg) a.dbg[blockIdx.x * 16 + (i_)] = globaltimer_ns(); } while (0)

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// One thread per CTA: publish this CTA's writes, arrive, wait for all `expected` CTAs (cooperative launch: co-resident).
__device__ __forceinline__ void grid_arrive_and_wait(unsigned int* ctr, unsigned int expected) {
    __threadfence();
    atomicAdd(ctr, 1u);
    const long long t0 = clock64();
    while (ld_acquire_u32(ctr) < expected) {
        __nanosleep(40);
        if (clock64() - t0 > 6000000000ll) __trap();  // seconds: a lost CTA must fail loudly, never hang the device
    }
    __threadfence();
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }  // the 4 epilogue warps

// k-th largest of n <= 1280 values (global, written by other CTAs before a grid barrier), by the 128 epilogue threads:
// binary search over the 32 key bits, one count + named barrier per bit.  Every thread returns the same value.
__device__ __forceinline__ uint32_t kth_largest_u32(const uint32_t* __restrict__ vals, int n, int kth, int et, int* s_cnt) {
    uint32_t v[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) v[i] = i * 128 + et < n ? __ldcg(vals + i * 128 + et) : 0u;
    uint32_t lo = 0;
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand = lo | (1u << bit);
        int c = 0;
#pragma unroll
        for (int i = 0; i < 10; ++i) c += v[i] >= cand ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        int* sc = s_cnt + ((bit & 1) << 2);
        if ((et & 31) == 0) sc[et >> 5] = c;
        epi_bar_sync();
        if (sc[0] + sc[1] + sc[2] + sc[3] >= kth) lo = cand;
    }
    return lo;
}

// MT = query tiles per launch: 1 (B <= 128) or 2 (B <= 256; the two TMEM accumulators then hold the first corpus tile of both
// query tiles through the cold phase, and alternate between the query tiles afterwards).
// XF = the l2 / ip spaces over raw vectors (FusedArgs::norms): a separate instantiation, so the cosine kernels are the
// same machine code with or without that feature.
// BMQ = rows of the query operand tile (the MMA's M): 128, or 64 for batches of up to 64 queries.  Every K-chunk stage
// re-fetches the query operand from L2 beside the corpus chunk (16 KB beside 32 KB at M = 128); the L2 -> SM fabric, not HBM,
// is what saturates first (~10 TB/s chip-wide, profiles/), so in the HBM-bound regime a half-height operand tile (8 KB)
// hands the difference to the corpus stream.  With M = 64 the accumulator occupies TMEM lanes 0-15 of every 32-lane
// quadrant (row r -> lane (r / 16) * 32 + r % 16), i.e. epilogue thread (quad, lane < 16) owns query quad * 16 + lane.
template <typename T, int CAP, int KPRIME, int MT, bool XF, int BMQ = BM>
__global__ void __launch_bounds__(kThreads, 1)
tc_fused_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, const FusedArgs a) {
    static_assert(MT == 1 || MT == 2, "two accumulators: at most two query tiles");
    static_assert(BMQ == BM || (BMQ == 64 && MT == 1), "half-height query tiles serve a single query tile");
    constexpr int kQ = MT * BM;  // query slots per CTA (stride of the per-CTA list / count arrays)
    constexpr int kABytesQ = BMQ * kRowBytes;               // query operand bytes per K-chunk
    constexpr int kStageBytesQ = kABytesQ + kBBytes;        // one stage of the [A | B] ring
    constexpr int kLanesPerQuad = BMQ / 4;                  // TMEM lanes of a quadrant that hold accumulator rows
    // Two query tiles over fp32 rows share every staged corpus chunk ([A0 | A1 | B] stages): the tf32 chunks are twice as
    // many bytes per flop, staging them once instead of twice is worth more than overlapping one query tile's epilogue
    // with the other's MMAs (measured at 1M x 768, B = 256: 0.95 -> 0.70 ms).  Over bf16 rows it is the other way round
    // (0.38 -> 0.43 ms), so there each query tile streams the chunk itself and the accumulators alternate.
    constexpr bool kShareB = MT == 2 && sizeof(T) == 4;
    constexpr int kStage2Bytes = 2 * kABytes + kBBytes, kStages2 = 3;
    static_assert(kStages2 * kStage2Bytes <= kStages * kStageBytes, "the two-tile ring must fit the one-tile ring's shared memory");
    using K = Kind<T>;
    constexpr int kChunks768 = kDim / K::kElemsPerChunk;
    constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)K::kFmt << 7) | ((uint32_t)K::kFmt << 10) |
                                ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BMQ >> 4) << 24);

    // Band pass (second launch, see search_band): thresholds are GIVEN (k-th exact score of the first launch minus eps),
    // every row above them is collected and rescored, nothing is learnt.  No flagged query -> nothing to do.
    const bool band = a.band_list != nullptr;
    const uint32_t n_queries = band ? min(__ldcg(a.band_n), (uint32_t)kQ) : a.n_queries;
    if (band && n_queries == 0) return;  // uniform over the grid, before any barrier or TMEM allocation

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + (size_t)kStages * kStageBytes);
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kStages;
    const uint32_t bar_tfull = bar_empty + 8 * kStages, bar_tempty = bar_tfull + 16;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
    int* s_cnt = reinterpret_cast<int*>(smem_al + (size_t)kStages * kStageBytes + 128);  // [2][4]
    uint32_t* hit_stage = reinterpret_cast<uint32_t*>(smem_al + (size_t)kStages * kStageBytes + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t G = gridDim.x;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const uint32_t first_tile = blockIdx.x;
    const uint32_t n_my = first_tile < a.n_tiles ? (a.n_tiles - first_tile + G - 1) / G : 0;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            if constexpr (!kShareB) {
                for (uint32_t nt = first_tile; nt < a.n_tiles; nt += G)
                    for (int mt = 0; mt < MT; ++mt)
                        for (int kc = 0; kc < kChunks768; ++kc) {
                            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                            const uint32_t sa = smem_base + stage * kStageBytesQ, sb = sa + kABytesQ;
                            mbar_expect_tx(bar_full + 8 * stage, kStageBytesQ);
                            tma_load_2d(sa, &map_q, bar_full + 8 * stage, kc * K::kElemsPerChunk, mt * BM);
                            tma_load_2d(sb, &map_x, bar_full + 8 * stage, kc * K::kElemsPerChunk, (int)(nt * BN));
                            if (++stage == kStages) { stage = 0; phase ^= 1; }
                        }
            } else {
                // two query tiles share every corpus chunk: a stage is [A0 | A1 | B] (64 KB, three stages in the same
                // 192 KB), so the corpus is staged once per chunk, not once per query tile
                for (uint32_t nt = first_tile; nt < a.n_tiles; nt += G)
                    for (int kc = 0; kc < kChunks768; ++kc) {
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                        const uint32_t sa = smem_base + stage * kStage2Bytes, sb = sa + 2 * kABytes;
                        mbar_expect_tx(bar_full + 8 * stage, kStage2Bytes);
                        tma_load_2d(sa, &map_q, bar_full + 8 * stage, kc * K::kElemsPerChunk, 0);
                        tma_load_2d(sa + kABytes, &map_q, bar_full + 8 * stage, kc * K::kElemsPerChunk, BM);
                        tma_load_2d(sb, &map_x, bar_full + 8 * stage, kc * K::kElemsPerChunk, (int)(nt * BN));
                        if (++stage == kStages2) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            if constexpr (!kShareB) {
                for (uint32_t w = 0; w < n_my * MT; ++w) {  // w = tile * MT + query tile
                    const uint32_t acc = w & 1, acc_phase = (w >> 1) & 1;
                    mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
                    tcgen05_fence_after();
                    const uint32_t tmem_d = tmem_base + acc * BN;
                    for (int kc = 0; kc < kChunks768; ++kc) {
                        mbar_wait(bar_full + 8 * stage, phase);
                        tcgen05_fence_after();
                        const uint32_t sa = smem_base + stage * kStageBytesQ, sb = sa + kABytesQ;
                        const uint64_t adesc = smem_desc(sa), bdesc = smem_desc(sb);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            tcgen05_mma<K::kTf32>(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, kIdesc, (kc | kk) != 0);
                        tcgen05_commit(bar_empty + 8 * stage);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                    tcgen05_commit(bar_tfull + 8 * acc);
                }
            } else {
                // accumulator mt belongs to query tile mt; both are filled from the same staged corpus chunk and complete
                // together, once per corpus tile (the epilogue's w = 2 t + mt sees accumulator mt in phase t & 1)
                for (uint32_t t = 0; t < n_my; ++t) {
                    mbar_wait(bar_tempty + 0, (t & 1) ^ 1);
                    mbar_wait(bar_tempty + 8, (t & 1) ^ 1);
                    tcgen05_fence_after();
                    for (int kc = 0; kc < kChunks768; ++kc) {
                        mbar_wait(bar_full + 8 * stage, phase);
                        tcgen05_fence_after();
                        const uint32_t sa = smem_base + stage * kStage2Bytes;
                        const uint64_t a0 = smem_desc(sa), a1 = smem_desc(sa + kABytes), bdesc = smem_desc(sa + 2 * kABytes);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            tcgen05_mma<K::kTf32>(tmem_base, a0 + 2 * kk, bdesc + 2 * kk, kIdesc, (kc | kk) != 0);
                            tcgen05_mma<K::kTf32>(tmem_base + BN, a1 + 2 * kk, bdesc + 2 * kk, kIdesc, (kc | kk) != 0);
                        }
                        tcgen05_commit(bar_empty + 8 * stage);
                        if (++stage == kStages2) { stage = 0; phase ^= 1; }
                    }
                    tcgen05_commit(bar_tfull + 0);
                    tcgen05_commit(bar_tfull + 8);
                }
            }
        }
    } else {
        // ===================== epilogue =====================
        const int quad = warp & 3;
        const int et = (int)threadIdx.x - 64;  // 0..127
        const uint32_t lane_taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
        // this thread's query of every query tile: TMEM lane quad*32 + lane of accumulator mt
        uint32_t q[MT], cnt[MT];
        bool q_ok[MT];
        uint64_t* lst[MT];
        float tau[MT];
        ChunkTop trk[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            q[mt] = mt * BM + quad * kLanesPerQuad + lane;
            q_ok[mt] = lane < kLanesPerQuad && q[mt] < n_queries;
            lst[mt] = a.lists + ((size_t)blockIdx.x * kQ + q[mt]) * CAP;
            cnt[mt] = 0;
            trk[mt].out = a.cmax + (size_t)(q_ok[mt] ? q[mt] : 0) * a.ncols + blockIdx.x * 8;
            trk[mt].on = false;
#pragma unroll
            for (int i = 0; i < 8; ++i) trk[mt].v[i] = 0u;
        }
        TileSpace sp[MT];
        if constexpr (XF) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                sp[mt].norms = a.norms;
                sp[mt].bounds = a.nbounds;
                sp[mt].h = q_ok[mt] ? __ldg(a.h_op + q[mt]) : 0.0f;
            }
        }

        if (!band) {
            // ---- 1. chunk maxima of the first tile (of every query tile: accumulator mt holds it)
            if (et == 0) MQVS_STAMP(0);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                if (n_my > 0) {
                    mbar_wait(bar_tfull + 8 * mt, 0);
                    tcgen05_fence_after();
                    if (et == 0 && mt == 0) MQVS_STAMP(1);
                    tile_chunk_maxima(lane_taddr + mt * BN, first_tile * BN, a.n_rows, a.mask, trk[mt].v, sp[mt], lane);
                }
                if (q_ok[mt]) {
#pragma unroll
                    for (int ch = 0; ch < BN / 32; ++ch) trk[mt].out[ch] = trk[mt].v[ch];  // zeros when this CTA has no tile
                }
            }
            __threadfence();
            epi_bar_sync();
            if (et == 0) { MQVS_STAMP(2); grid_arrive_and_wait(a.gbar + 0, G); MQVS_STAMP(3); }
            epi_bar_sync();

            // ---- 2. CTA c: tau = k'-th largest chunk maximum, for its queries c (and c + G)
#pragma unroll
            for (int j = 0; j < MT; ++j) {
                const uint32_t qo = blockIdx.x + j * G;
                if (qo < n_queries) {
                    const uint32_t t = kth_largest_u32(a.cmax + (size_t)qo * a.ncols, (int)a.ncols, KPRIME, et, s_cnt);
                    if (et == 0) a.tau[qo] = t > kOrdNegInf ? t : kOrdNegInf;
                }
            }
            __threadfence();
            epi_bar_sync();
            if (et == 0) { MQVS_STAMP(4); grid_arrive_and_wait(a.gbar + 1, G); MQVS_STAMP(5); }
            epi_bar_sync();
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            tau[mt] = q_ok[mt] ? ord_to_float(__ldcg(a.tau + q[mt])) : INFINITY;
            trk[mt].mn = trk[mt].v[0];
#pragma unroll
            for (int i = 1; i < 8; ++i) trk[mt].mn = min(trk[mt].mn, trk[mt].v[i]);
        }

        // ---- 3. first tile again, now against tau, then the rest of this CTA's tiles
        uint32_t w = 0, t_idx = 0;
        for (uint32_t nt = first_tile; nt < a.n_tiles; nt += G, ++t_idx) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt, ++w) {
                const uint32_t acc = w & 1, acc_phase = (w >> 1) & 1;
                if (t_idx > 0 || band) {  // (the first launch already waited for its first tile in step 1)
                    if (q_ok[mt]) tau[mt] = fmaxf(tau[mt], ord_to_float(__ldcg(a.tau + q[mt])));  // raised by refreshes / compactions
                    mbar_wait(bar_tfull + 8 * acc, acc_phase);
                    tcgen05_fence_after();
                }
                trk[mt].on = !band && t_idx > 0 && q_ok[mt];  // the first tile's chunks are already in trk
                epilogue_tile<CAP, KPRIME, true>(lane_taddr + acc * BN, nt * BN, a.n_rows, a.mask, hit_stage + (warp - 2) * 1024,
                                                 a.tau + q[mt], lane, tau[mt], cnt[mt], lst[mt], &trk[mt], sp[mt]);
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
            }
            // Threshold refresh after 2, 4, 8, ... tiles and after the last one: every (CTA, query) mirrors its 8 largest chunk
            // maxima so far in cmax, so the k'-th largest of a query's G*8 slots is a valid threshold that tightens as the scan
            // proceeds (a query then collects ~k' candidates per doubling of the rows scanned instead of k' N / (256 G)).
            // Slots are replaced one word at a time and a chunk is only ever written to one slot, so whatever mix of old and
            // new values the reader sees, distinct slots are distinct chunks, i.e. distinct rows scoring at least that much.
            if (!band && t_idx >= 1 && (((t_idx + 1) & t_idx) == 0 || t_idx + 1 == n_my)) {
#pragma unroll
                for (int j = 0; j < MT; ++j) {
                    const uint32_t qo = blockIdx.x + j * G;
                    if (qo < n_queries) {
                        const uint32_t t = kth_largest_u32(a.cmax + (size_t)qo * a.ncols, (int)a.ncols, KPRIME, et, s_cnt);
                        if (et == 0 && t > kOrdNegInf) atomicMax(a.tau + qo, t);
                    }
                }
            }
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
            if (q_ok[mt]) a.cnt[(size_t)blockIdx.x * kQ + q[mt]] = cnt[mt];
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }

    // ---- 4. all lists are complete once every CTA is here; CTA c finishes query c
    __threadfence();
    if (threadIdx.x == 0) { MQVS_STAMP(6); grid_arrive_and_wait(a.gbar + 2, G); MQVS_STAMP(7); }
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < MT; ++j) {
        const uint32_t slot = blockIdx.x + j * G;
        if (slot < n_queries) fused_tail<T, CAP, KPRIME, XF>(a, (int)G, smem_al, (int)slot, kQ);
        __syncthreads();
    }
}

// ---- tail of tc_fused_kernel: CTA c turns the G per-CTA lists of query c into the final top-k.
// No sort of approximate keys: after the last threshold refresh tau[q] is within a few ranks of the k'-th best score, so
// the candidates that still clear it (~k', counted) ARE the shortlist.  All of them are rescored exactly, and the output
// order comes from ranking each exact key against the others (keys are unique: they embed the row id).  A sorting
// network on one warp costs 10+ us here (single-warp issue latency); this tail is ~6 us at k = 10.
template <typename T, int CAP, int KPRIME, bool XF>
__device__ __forceinline__ void fused_tail(const FusedArgs& a, int G, uint8_t* smem_al, const int slot, const int b_pad) {
    constexpr int NW = kThreads / 32;
    constexpr int kCcap = KPRIME > 256 ? 4096 : 1024;  // shortlist capacity of the direct path (>= 2 k' with room for ties)
    constexpr int kCbuf = kCcap > 2048 ? kCcap : 2048;
    // the operand ring is idle now (every TMA load was consumed by an MMA that has retired): reuse it
    uint64_t* C = reinterpret_cast<uint64_t*>(smem_al);   // kCbuf keys (direct path uses kCcap; the fallback 2048)
    uint64_t* ex = C + kCbuf;                             // kCcap exact keys
    int* pre_s = reinterpret_cast<int*>(ex + kCcap);      // kThreads + 1
    int* warp_tot = pre_s + kThreads + 1;                 // NW
    int* fill_s = warp_tot + NW;
    int* valid_s = fill_s + 1;
    const bool band = a.band_list != nullptr;
    // slot = operand row / (query tile, TMEM lane) of the query this call finishes; b_pad = slots per CTA in lists / cnt
    const int qf = band ? (int)__ldcg(a.band_list + slot) : slot;   // the query it is
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // list lengths -> exclusive prefix sum (one list per thread)
    const int mine = tid < G ? min((int)__ldcg(a.cnt + (size_t)tid * b_pad + slot), CAP) : 0;
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    if (tid == 0) { *fill_s = 0; *valid_s = 0; }
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    pre_s[tid + 1] = base + incl;
    if (tid == 0) pre_s[0] = 0;
    __syncthreads();
    const int total = pre_s[G];

    // every row that is not a candidate scored < tau[q] (filtered) or <= tau[q] (dropped by a list compaction, which
    // raised tau[q] to its k'-th key), so tau[q] bounds the approximate score of everything outside the shortlist
    uint32_t tq = __ldcg(a.tau + slot);
    // band pass: a list compaction raised the threshold, i.e. dropped rows inside the band (hundreds of ties in one CTA's
    // share): the query stays flagged for the exact scan
    if (band && tq != __ldcg(a.band + qf)) return;
    for (int idx = tid; idx < total; idx += kThreads) {
        int lo = 0, hi = G;  // list g with pre_s[g] <= idx < pre_s[g + 1]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (pre_s[mid] <= idx) lo = mid; else hi = mid;
        }
        const uint64_t key = __ldcg(a.lists + ((size_t)lo * b_pad + slot) * CAP + (idx - pre_s[lo]));
        if ((uint32_t)(key >> 32) >= tq) {
            const int slot = atomicAdd(fill_s, 1);
            if (slot < kCcap) C[slot] = key;
        }
    }
    __syncthreads();
    int nC = *fill_s;
    if (band && nC > kCcap) return;  // too many rows inside the band to rescore here: stays flagged
    if (nC > kCcap) {
        if constexpr (KPRIME > 256) {
            // wide shortlists: the streaming selection's buffer cannot hold 2 k' keys -- flag the query for the repair chain
            if (tid == 0) {
                a.flags[qf] = 1u;
                atomicAdd(a.flag_count, 1u);
                a.band[qf] = tq;  // (everything at or above the learnt threshold: the band pass retries with its own capacity)
            }
            return;
        } else {
            // far more survivors than k' (hundreds of exact ties, adversarial orderings): the general streaming selection
            // picks the k' best approximate keys; its k'-th key then bounds everything it left out
            __syncthreads();
            select_lists_body<kThreads>(a.lists, a.cnt, G, b_pad, CAP, KPRIME, nullptr, slot, C, pre_s, warp_tot, fill_s);
            if (C[KPRIME - 1]) tq = max(tq, (uint32_t)(C[KPRIME - 1] >> 32));
            nC = KPRIME;
        }
    }
    const float tt = ord_to_float(tq);
    if (tid == 0) MQVS_STAMP(8);

    // K4: canonical fp32 rescoring, one warp per row, four rows in flight per warp
    float4 qv[kChunks];
#pragma unroll
    for (int c = 0; c < kChunks; ++c) qv[c] = *reinterpret_cast<const float4*>(a.qn + (size_t)qf * kDim + c * 128 + lane * 4);
    float hq = 0.0f;
    if constexpr (XF) hq = __ldg(a.hq + qf);
    int n_valid = 0;
    auto rescore = [&](auto* rows) {  // rows: the stored rows, in their own type
        using TR = std::remove_cv_t<std::remove_pointer_t<decltype(rows)>>;
        constexpr int kInFlight = 4;
        for (int i0 = warp * kInFlight; i0 < nC; i0 += NW * kInFlight) {
            uint64_t key[kInFlight];
            RowFrag<TR> x[kInFlight];
#pragma unroll
            for (int u = 0; u < kInFlight; ++u) {
                key[u] = i0 + u < nC ? C[i0 + u] : 0ull;
                if (key[u] != 0ull) x[u].load(rows + (size_t)key_id(key[u]) * kDim, lane);  // warp-uniform
            }
#pragma unroll
            for (int u = 0; u < kInFlight; ++u) {
                uint64_t out = 0ull;
                if (key[u] != 0ull) {
                    float4 xv[kChunks];
#pragma unroll
                    for (int c = 0; c < kChunks; ++c) xv[c] = x[u].get(c);
                    float sc = warp_butterfly_sum(lane_dot(qv, xv));
                    if constexpr (XF) sc = rank_score(a.metric, sc, __ldg(a.norms + key_id(key[u])), hq);  // exact, as the SIMT scan
                    out = pack_key(sc, key_id(key[u]));
                    ++n_valid;
                }
                if (lane == 0 && i0 + u < nC) ex[i0 + u] = out;
            }
        }
    };
    if constexpr (sizeof(T) == 2) {
        // scored over the bf16 shadow of an fp32 store: the exact scores come from the stored fp32 rows (block-uniform branch)
        if (a.rows_f32 != nullptr) rescore(reinterpret_cast<const float*>(a.rows_f32));
        else rescore(reinterpret_cast<const T*>(a.rows));
    } else {
        rescore(reinterpret_cast<const T*>(a.rows));
    }
    if (lane == 0 && n_valid) atomicAdd(valid_s, n_valid);
    __syncthreads();
    if (tid == 0) MQVS_STAMP(9);

    // rank of each exact key among the others = its output position
    const int n_valid_all = *valid_s;
    for (int i = tid; i < nC; i += kThreads) {
        const uint64_t e = ex[i];
        if (e == 0ull) continue;
        int rank = 0;
        for (int j = 0; j < nC; ++j) rank += ex[j] > e ? 1 : 0;
        if (rank < a.k) {
            a.out_scores[(size_t)qf * a.k + rank] = key_score(e);
            a.out_ids[(size_t)qf * a.k + rank] = (int64_t)key_id(e) + a.id_base;
        }
        if (rank == a.k - 1) {  // the k-th exact score must clear the bound of everything outside the shortlist
            // band pass: every row that could reach the first launch's k-th exact score was inside the band and has been
            // rescored, so the result is exact by construction
            const bool ok = band || tt == -INFINITY || key_score(e) > tt + a.eps[qf];
            if (band) {
                a.flags[qf] = 0u;
                atomicSub(a.flag_count, 1u);
                atomicAdd(a.counters + 2, 1ull);
            } else {
                a.flags[qf] = ok ? 0u : 1u;
                if (!ok) {
                    atomicAdd(a.flag_count, 1u);
                    a.band[qf] = ord_u32(__fsub_rd(key_score(e), a.eps[qf]));  // anything that can beat e scores at least this
                }
            }
        }
    }
    for (int j = n_valid_all + tid; j < a.k; j += kThreads) {  // fewer candidates than k: pad
        a.out_scores[(size_t)qf * a.k + j] = -INFINITY;
        a.out_ids[(size_t)qf * a.k + j] = -1;
    }
    if (tid == 0 && n_valid_all < a.k) {
        if (band) {  // everything above the band threshold was rescored and there are fewer than k such rows
            a.flags[qf] = 0u;
            atomicSub(a.flag_count, 1u);
            atomicAdd(a.counters + 2, 1ull);
        } else {
            const bool ok = tt == -INFINITY;  // nothing was ever discarded: the candidates are all the allowed rows
            a.flags[qf] = ok ? 0u : 1u;
            if (!ok) {
                atomicAdd(a.flag_count, 1u);
                a.band[qf] = kOrdNegInf;
            }
        }
    }
    if (tid == 0) MQVS_STAMP(10);
}

// One dispatch over (row type, k', query tiles) per value of XF.
template <typename T, int KP, int MT, bool XF, int BMQ = BM>
cudaError_t launch_fused(const CUtensorMap& mq, const CUtensorMap& mx, const FusedArgs& a, int grid, cudaStream_t st) {
    const void* kern = (const void*)tc_fused_kernel<T, (KP > 256 ? 2048 : 512), KP, MT, XF, BMQ>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
    void* args[] = {(void*)&mq, (void*)&mx, (void*)&a};
    // cooperative: the grid barriers need every CTA resident (one per SM), whatever else shares the device
    return cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kThreads), args, kSmemBytes, st);
}

// query_tiles: 2, 1, or 0 = one HALF-HEIGHT tile (M = 64; the map of the query matrix then has a 64-row box)
template <bool XF>
cudaError_t launch_fused_any(int dtype, int kprime, int query_tiles, const CUtensorMap& mq, const CUtensorMap& mx,
                             const FusedArgs& a, int grid, cudaStream_t st) {
#define MQVS_TCF(KP_)                                                                                     \
    if (kprime == KP_) {                                                                                  \
        if (query_tiles == 0)                                                                             \
            return dtype == MQVS_F32 ? launch_fused<float, KP_, 1, XF, 64>(mq, mx, a, grid, st)           \
                                     : launch_fused<__nv_bfloat16, KP_, 1, XF, 64>(mq, mx, a, grid, st);  \
        if (query_tiles == 1)                                                                             \
            return dtype == MQVS_F32 ? launch_fused<float, KP_, 1, XF>(mq, mx, a, grid, st)               \
                                     : launch_fused<__nv_bfloat16, KP_, 1, XF>(mq, mx, a, grid, st);      \
        return dtype == MQVS_F32 ? launch_fused<float, KP_, 2, XF>(mq, mx, a, grid, st)                   \
                                 : launch_fused<__nv_bfloat16, KP_, 2, XF>(mq, mx, a, grid, st);          \
    }
    MQVS_TCF(32)
    MQVS_TCF(64)
    MQVS_TCF(128)
    MQVS_TCF(256)
    MQVS_TCF(512)
    MQVS_TCF(1152)
#undef MQVS_TCF
    return cudaErrorInvalidValue;
}

}  // namespace tc
}  // namespace mqvs
