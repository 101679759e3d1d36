// kernels_tc.cu -- K2' (tensor-core scoring) + K3 (fused shortlist epilogue) + K4 (exact fp32 rescoring): the SEGMENTED
// pipeline that serves batches of more than 256 queries (cold pass, corpus segments of growing size with a selection
// after each, rescoring, margin check).  Batches of up to 256 queries run the same building blocks as one cooperative
// launch: kernels_tc_fused.cu.  Shared device code: tc_device.cuh.
//
// For query batches where the contraction [B, 768] x [768, N] is real GEMM work, scoring runs on the 5th-gen
// tensor cores: tcgen05.mma (kind::f16 for bf16 rows with bf16-rounded queries, kind::tf32 for fp32 rows),
// operands staged by TMA into 128B-swizzled shared memory through a 4-stage mbarrier ring, fp32 accumulators in
// TMEM, double-buffered so the epilogue of tile t overlaps the MMAs of tile t+1.
//
// Orientation: QUERIES on the MMA M axis (TMEM lanes), CORPUS rows on N (TMEM columns).  After tcgen05.ld each
// epilogue thread owns one query and walks that query's scores for 256 corpus rows in registers: one compare per
// score against a per-query threshold held in a register.  Survivors are appended to a thread-private list
// in global memory ([cta][query][CAP] 64-bit keys); when a list passes CAP-32 entries the warp sorts it in
// registers (bitonic, CAP/32 keys per lane), keeps the best k' and raises the threshold, which is shared across
// CTAs through an atomicMax on tau[query] -- every CTA filters with the tightest bound any CTA has proven.
// The score matrix (B x N; 164 GB at 4096 x 10M) is never materialised.
//
// The tensor-core scores are approximate (bf16-rounded queries / tf32 operands), so they only SHORTLIST.
// K4 recomputes the k' best candidates per query in the canonical fp32 order (common.cuh) and the final top-k
// is taken from those exact scores.  Exactness is checked, not assumed: every discarded row had an approximate
// score <= tau~, hence an exact score <= tau~ + eps (eps = rigorous bound on |approx - exact|); if the k-th exact
// score does not exceed tau~ + eps the query is flagged and re-run on the exact SIMT scan (kernels_scan.cu).
//
// Roofline: HBM while B is small (corpus streamed once per batch), tensor pipe for large B:
// 2 * N * 768 * B flop per batch.

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "launch.h"

#include "tc_device.cuh"

namespace mqvs {
namespace tc {

// ------------------------------------------------------------------------------------------ main kernel
struct ScoreArgs {
    uint32_t n_rows;        // valid corpus rows
    uint32_t n_queries;     // valid queries (B); lanes beyond are padding
    uint32_t m_tiles;       // ceil(B / 128)
    uint32_t tile_begin;    // this launch scores corpus tiles [tile_begin, tile_end) (256 rows each)
    uint32_t tile_end;
    const uint32_t* mask;   // optional allow-mask over rows (device), or nullptr
    uint32_t* tau;          // [B_pad] shared per-query threshold, order-preserving uint, atomicMax
    uint32_t* cnt;          // [grid][B_pad] list lengths
    uint64_t* lists;        // [grid][B_pad][CAP]
    uint32_t b_pad;
    uint32_t* cmax;         // cold pass only (else nullptr): [B_pad][ncols] chunk maxima, no list appends
    uint32_t ncols;         // 8 * CTAs of the cold pass
    // l2 / ip spaces over raw vectors (XF instantiations only; tc.h:Space)
    const float* norms;     // [n_rows] ||x||
    const float2* nbounds;  // [ceil(n_rows / 32)] (max, min) norm per 32-row chunk
    const float* hq;        // [B_pad] 0.5 / max(||q||, eps) (l2), 0 (ip)
};

// XF = the l2 / ip spaces: the epilogue maps each chunk's cosines to ranking scores (tc_device.cuh:apply_space) before the
// thresholds see them; a separate instantiation, so the cosine kernels are unchanged.
template <typename T, int CAP, int KPRIME, bool XF>
__global__ void __launch_bounds__(kThreads, 1)
tc_score_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, ScoreArgs a) {
    using K = Kind<T>;
    constexpr int kChunks768 = kDim / K::kElemsPerChunk;  // 12 (bf16) or 24 (tf32) K-chunks per tile
    constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)K::kFmt << 7) | ((uint32_t)K::kFmt << 10) |
                                ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + (size_t)kStages * kStageBytes);
    // bars[0..S): full, [S..2S): empty, [2S..2S+2): tmem_full, [2S+2..2S+4): tmem_empty
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kStages;
    const uint32_t bar_tfull = bar_empty + 8 * kStages, bar_tempty = bar_tfull + 16;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
    uint32_t* hit_stage = reinterpret_cast<uint32_t*>(smem_al + (size_t)kStages * kStageBytes + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 4);  // one arrive per epilogue warp
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

    const uint32_t first_tile = a.tile_begin + blockIdx.x;
    const uint32_t total_work = a.tile_end > first_tile ? ((a.tile_end - first_tile + gridDim.x - 1) / gridDim.x) * a.m_tiles : 0;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t nt = first_tile; nt < a.tile_end; nt += gridDim.x)
                for (uint32_t mt = 0; mt < a.m_tiles; ++mt)
                    for (int kc = 0; kc < kChunks768; ++kc) {
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                        const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
                        mbar_expect_tx(bar_full + 8 * stage, kStageBytes);
                        tma_load_2d(sa, &map_q, bar_full + 8 * stage, kc * K::kElemsPerChunk, (int)(mt * BM));
                        tma_load_2d(sb, &map_x, bar_full + 8 * stage, kc * K::kElemsPerChunk, (int)(nt * BN));
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t w = 0; w < total_work; ++w) {
                const uint32_t acc = w & 1, acc_phase = (w >> 1) & 1;
                mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);  // epilogue has drained this accumulator
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kc = 0; kc < kChunks768; ++kc) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
                    const uint64_t adesc = smem_desc(sa), bdesc = smem_desc(sb);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)  // 4 x 32-byte K-slices inside the 128-byte swizzle row
                        tcgen05_mma<K::kTf32>(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, kIdesc, (kc | kk) != 0);
                    tcgen05_commit(bar_empty + 8 * stage);  // frees the smem slot when these MMAs retire
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                tcgen05_commit(bar_tfull + 8 * acc);  // accumulator complete -> epilogue
            }
        }
    } else {
        // ===================== epilogue: threshold + shortlist =====================
        const int quad = warp & 3;  // TMEM lanes 32*quad .. 32*quad+31 are this warp's
        uint32_t w = 0;
        for (uint32_t nt = first_tile; nt < a.tile_end; nt += gridDim.x)
            for (uint32_t mt = 0; mt < a.m_tiles; ++mt, ++w) {
                const uint32_t acc = w & 1, acc_phase = (w >> 1) & 1;
                const uint32_t q = mt * BM + quad * 32 + lane;
                const bool q_ok = q < a.n_queries;
                const size_t slot = (size_t)blockIdx.x * a.b_pad + q;
                uint64_t* lst = a.lists + slot * CAP;
                uint32_t cnt = q_ok ? a.cnt[slot] : 0;
                float tau = q_ok ? ord_to_float(__ldcg(a.tau + q)) : INFINITY;
                TileSpace sp;
                if constexpr (XF) {
                    sp.norms = a.norms;
                    sp.bounds = a.nbounds;
                    sp.h = q_ok ? __ldg(a.hq + q) : 0.0f;
                }

                mbar_wait(bar_tfull + 8 * acc, acc_phase);
                tcgen05_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;
                if (a.cmax != nullptr) {  // cold pass (uniform): only the chunk maxima that seed the thresholds
                    uint32_t cm[BN / 32];
                    tile_chunk_maxima(taddr, nt * BN, a.n_rows, a.mask, cm, sp, lane);
                    if (q_ok) {
                        uint4* dst = reinterpret_cast<uint4*>(a.cmax + (size_t)q * a.ncols + (size_t)(nt - a.tile_begin) * 8);
                        dst[0] = make_uint4(cm[0], cm[1], cm[2], cm[3]);
                        dst[1] = make_uint4(cm[4], cm[5], cm[6], cm[7]);
                    }
                } else
                epilogue_tile<CAP, KPRIME>(taddr, nt * BN, a.n_rows, a.mask, hit_stage + (warp - 2) * 1024, a.tau + q, lane, tau, cnt, lst,
                                           nullptr, sp);
                // accumulator fully read -> hand TMEM back to the MMA warp
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
                if (q_ok) a.cnt[slot] = cnt;
            }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ CTA-pair kernel
// The same scoring + shortlist kernel on CTA PAIRS (cta_group::2), for batches of more than one query tile.
//
// Why: with single-CTA 128 x 256 tiles every K-chunk stage moves 16 KB of query operand + 32 KB of corpus through the SM's
// L2 port for 2 * 128 * 256 * 64 flop.  That is 590 KB per tile step against ~4.6 us of MMA time: 128 GB/s per SM, more than
// one SM's L2 port delivers (64 B/clk: 75-95 GB/s at the 1.2-1.5 GHz these parts sustain under the power cap), so large
// batches ran port-bound at ~1.22 PFLOP/s whatever the tensor pipe could do.  A pair issues ONE 256 (queries) x 256 (rows)
// MMA per K-slice: each CTA stages only ITS 128 queries and ITS 128 of the tile's 256 corpus rows (16 + 16 KB per stage),
// the tensor cores of both SMs read both halves.  Per SM and unit of MMA work the staged bytes drop from 48 to 32 KB and,
// more to the point, each SM's port carries half a corpus tile instead of a whole one per query tile.
//
// Layout of the work: pair p (cluster p, CTAs 2p and 2p + 1) takes corpus tiles tile_begin + p, + n_pairs, ...; for every
// corpus tile it walks the query-tile PAIRS (2 mp, 2 mp + 1): CTA rank r owns query tile 2 mp + r -- its TMEM lanes are
// those 128 queries, so the epilogue, the per-(CTA, query) lists and everything downstream are exactly the single-CTA
// kernel's.  An odd last query tile leaves rank 1 idle for that step (its operand box is out of range: zero-filled).
// Synchronisation: TMA of both CTAs counts on the LEADER's full barrier (expect_tx of both halves there); the leader's MMA
// warp issues, commits with multicast to both CTAs' empty and accumulator-full barriers; both CTAs' epilogue warps
// arrive on the leader's accumulator-empty barrier (count 8).
constexpr int kPairStages = 6;
constexpr int kPairBBytes = (BN / 2) * kRowBytes;           // this CTA's half of the corpus tile: 16 KB
constexpr int kPairStageBytes = kABytes + kPairBBytes;      // 32 KB
static_assert((size_t)kPairStages * kPairStageBytes <= (size_t)kStages * kStageBytes, "the pair ring must fit the shared memory of the single-CTA ring");

template <typename T, int CAP, int KPRIME, bool XF>
__global__ void __launch_bounds__(kThreads, 1)
tc_score_pair_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_xh, ScoreArgs a) {
    using K = Kind<T>;
    constexpr int kChunks768 = kDim / K::kElemsPerChunk;
    constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)K::kFmt << 7) | ((uint32_t)K::kFmt << 10) |
                                ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);  // M = 256 across the pair

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + (size_t)kStages * kStageBytes);
    // bars[0..S): full, [S..2S): empty, [2S..2S+2): tmem_full, [2S+2..2S+4): tmem_empty   (S = kPairStages)
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kPairStages;
    const uint32_t bar_tfull = bar_empty + 8 * kPairStages, bar_tempty = bar_tfull + 16;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kPairStages + 4);
    uint32_t* hit_stage = reinterpret_cast<uint32_t*>(smem_al + (size_t)kStages * kStageBytes + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();          // 0 = leader
    const uint32_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_xh) : "memory");
        for (int s = 0; s < kPairStages; ++s) {
            mbar_init(bar_full + 8 * s, 2);   // the two producers (the leader's arrive carries the expect_tx of both halves)
            mbar_init(bar_empty + 8 * s, 1);  // one multicast commit from the leader's MMA warp
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 8);  // four epilogue warps of each CTA
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's barriers are initialised before anything of ours can reach them
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const uint32_t first_tile = a.tile_begin + pair;
    const uint32_t mp_count = (a.m_tiles + 1) >> 1;
    const uint32_t n_my = a.tile_end > first_tile ? (a.tile_end - first_tile + n_pairs - 1) / n_pairs : 0;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs: own halves, counted on the leader's barrier) =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            const uint32_t full0 = mapa_u32(bar_full, 0);  // the leader's full barriers, cluster address space
            for (uint32_t nt = first_tile; nt < a.tile_end; nt += n_pairs)
                for (uint32_t mp = 0; mp < mp_count; ++mp)
                    for (int kc = 0; kc < kChunks768; ++kc) {
                        mbar_wait_bounded(bar_empty + 8 * stage, phase ^ 1);
                        const uint32_t sa = smem_base + stage * kPairStageBytes, sb = sa + kABytes;
                        if (rank == 0) mbar_expect_tx(bar_full + 8 * stage, 2 * kPairStageBytes);
                        tma_load_2d_pair(sa, &map_q, full0 + 8 * stage, kc * K::kElemsPerChunk, (int)((2 * mp + rank) * BM));
                        tma_load_2d_pair(sb, &map_xh, full0 + 8 * stage, kc * K::kElemsPerChunk, (int)(nt * BN + rank * (BN / 2)));
                        if (rank != 0) mbar_arrive_cluster(full0 + 8 * stage);
                        if (++stage == kPairStages) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader only) =====================
        if (lane == 0 && rank == 0) {
            uint32_t stage = 0, phase = 0;
            const uint32_t total_work = n_my * mp_count;
            for (uint32_t w = 0; w < total_work; ++w) {
                const uint32_t acc = w & 1, acc_phase = (w >> 1) & 1;
                mbar_wait_bounded(bar_tempty + 8 * acc, acc_phase ^ 1);  // both CTAs' epilogues have drained this accumulator
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kc = 0; kc < kChunks768; ++kc) {
                    mbar_wait_bounded(bar_full + 8 * stage, phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_base + stage * kPairStageBytes, sb = sa + kABytes;
                    const uint64_t adesc = smem_desc(sa), bdesc = smem_desc(sb);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        tcgen05_mma_pair<K::kTf32>(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, kIdesc, (kc | kk) != 0);
                    tcgen05_commit_pair(bar_empty + 8 * stage);  // frees the slot in BOTH CTAs when these MMAs retire
                    if (++stage == kPairStages) { stage = 0; phase ^= 1; }
                }
                tcgen05_commit_pair(bar_tfull + 8 * acc);  // accumulator complete -> both epilogues
            }
        }
    } else {
        // ===================== epilogue: this CTA's 128 queries =====================
        const int quad = warp & 3;
        const uint32_t tempty0 = mapa_u32(bar_tempty, 0);
        uint32_t w = 0;
        for (uint32_t nt = first_tile; nt < a.tile_end; nt += n_pairs)
            for (uint32_t mp = 0; mp < mp_count; ++mp, ++w) {
                const uint32_t acc = w & 1, acc_phase = (w >> 1) & 1;
                const uint32_t mt = 2 * mp + rank;
                const uint32_t q = mt * BM + quad * 32 + lane;
                const bool q_ok = mt < a.m_tiles && q < a.n_queries;
                const size_t slot = (size_t)blockIdx.x * a.b_pad + (q_ok ? q : 0);
                uint64_t* lst = a.lists + slot * CAP;
                uint32_t cnt = q_ok ? a.cnt[slot] : 0;
                float tau = q_ok ? ord_to_float(__ldcg(a.tau + q)) : INFINITY;
                TileSpace sp;
                if constexpr (XF) {
                    sp.norms = a.norms;
                    sp.bounds = a.nbounds;
                    sp.h = q_ok ? __ldg(a.hq + q) : 0.0f;
                }
                mbar_wait_bounded(bar_tfull + 8 * acc, acc_phase);
                tcgen05_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;
                if (mt < a.m_tiles) {  // (uniform over the CTA; an idle half only keeps the barrier protocol going)
                    if (a.cmax != nullptr) {  // cold pass: only the chunk maxima that seed the thresholds
                        uint32_t cm[BN / 32];
                        tile_chunk_maxima(taddr, nt * BN, a.n_rows, a.mask, cm, sp, lane);
                        if (q_ok) {
                            uint4* dst = reinterpret_cast<uint4*>(a.cmax + (size_t)q * a.ncols + (size_t)(nt - a.tile_begin) * 8);
                            dst[0] = make_uint4(cm[0], cm[1], cm[2], cm[3]);
                            dst[1] = make_uint4(cm[4], cm[5], cm[6], cm[7]);
                        }
                    } else {
                        epilogue_tile<CAP, KPRIME>(taddr, nt * BN, a.n_rows, a.mask, hit_stage + (warp - 2) * 1024, a.tau + (q_ok ? q : 0), lane,
                                                   tau, cnt, lst, nullptr, sp);
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tempty0 + 8 * acc);  // hand the accumulator back to the leader's MMA warp
                if (q_ok) a.cnt[slot] = cnt;
            }
    }

    tcgen05_fence_before();
    __syncthreads();
    cluster_sync_all();  // nobody leaves (or frees tensor memory) while the partner may still use this CTA's half
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ prep
// Normalised fp32 queries -> operand matrix [b_pad, 768] (bf16-rounded or fp32 copy), zero rows for padding,
// and eps[q], the rigorous bound on |tensor-core score - exact fp32 score| for unit-norm rows.
template <typename T>
__global__ void prep_queries_kernel(const float* __restrict__ qn, int B, int b_pad, T* __restrict__ q_op,
                                    float* __restrict__ eps, uint32_t* __restrict__ tau, int metric, float amax,
                                    const float* __restrict__ qnorm, float* __restrict__ hq,
                                    const unsigned int* __restrict__ shadow_err) {
    const int q = blockIdx.x, lane = threadIdx.x;  // one warp per query row
    float err2 = 0.f;
    for (int j = lane; j < kDim; j += 32) {
        float v = q < B ? qn[(size_t)q * kDim + j] : 0.f;
        if constexpr (sizeof(T) == 2) {
            __nv_bfloat16 h = __float2bfloat16_rn(v);
            q_op[(size_t)q * kDim + j] = h;
            float d = v - __bfloat162float(h);
            err2 += d * d;
        } else {
            q_op[(size_t)q * kDim + j] = v;
        }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) err2 += __shfl_xor_sync(0xffffffffu, err2, m);
    if (lane == 0) {
        // bf16 rows: stored rows are exact operands, only the query is rounded: |err| <= ||dq|| * ||x||, with
        // ||x|| <= 1.004 after bf16 rounding of a unit row.  tf32: both operands lose 13 mantissa bits (relative
        // error < 2^-10 each): |err| <= 2^-9 * sum|q_i x_i| <= 2^-9 * 1.001.  Both get 1e-4 for the tensor core's fp32
        // accumulation (768 additions, each off by at most 2^-23 of a partial sum bounded by ||q|| ||x||).
        float e = sizeof(T) == 2 ? sqrtf(err2) * 1.004f * 1.0001f + 1e-4f : 2.0f * 0.0009765625f * 1.001f + 1e-4f;
        // bf16 SHADOW of an fp32 store: the corpus operand is rounded too: + ||q|| ||x - bf16(x)||, the largest over the rows
        if (sizeof(T) == 2 && shadow_err != nullptr) e += __uint_as_float(__ldg(shadow_err)) * 1.0002f;
        if (metric != MQVS_METRIC_COSINE) {  // l2 / ip ranking scores t = a s - a^2 h: same bound as prep_fused_kernel
            const float h = metric == MQVS_METRIC_L2 && q < B ? __fdiv_rn(0.5f, fmaxf(qnorm[q], kEps)) : 0.0f;
            if (q < B) hq[q] = h;  // (padding rows never read theirs)
            e = amax * e * 1.0001f + 1e-6f * (amax + amax * amax * h);
        }
        if (q < B) eps[q] = e;
        tau[q] = kOrdNegInf;
    }
}

// ------------------------------------------------------------------------------------------ select + rescore
constexpr int kSelThreads = 256;

// One CTA per query, run after every corpus segment: (running shortlist of earlier segments) U (this segment's per-CTA
// lists) -> new running shortlist of the kp (= k') best approximate keys, and tau[q] raised to the k'-th best score of
// everything seen so far -- the threshold the next segment's epilogue filters with.
// Nobody needs the shortlist SORTED (the next call needs the set and its minimum, K4 rescoring the set), so this is a
// selection: all candidates that beat the previous k'-th key go to shared memory and an MSB-first radix select (8-bit
// digits, per-warp histograms, bytes every key shares skipped) finds the k'-th largest key in a few passes; the keys
// above it are compacted into slots [0, k'-1) and the k'-th key itself is stored LAST, where the next call looks for the
// floor.  At 4096 queries x k' = 256 this replaced bitonic sorts of 2048 keys: 650-820 us -> see DESIGN.md per call.
// More candidates than fit in shared memory (adversarial orderings) take the general streaming selection instead.
constexpr int kSelMaxKeys = 8192;  // 64 KB of keys

__global__ void __launch_bounds__(kSelThreads)
select_lists_kernel(const uint64_t* __restrict__ lists, const uint32_t* __restrict__ cnt, uint32_t* __restrict__ tau,
                    int grid_ctas, int b_pad, int cap, int kp, int first_segment, uint64_t* __restrict__ short_keys,
                    float* __restrict__ tau_tilde) {
    extern __shared__ __align__(16) uint64_t keys[];  // kSelMaxKeys
    __shared__ int pre_s[kSelThreads + 1];
    __shared__ int warp_tot[kSelThreads / 32];
    __shared__ int fill_s;
    __shared__ unsigned hist[kSelThreads / 32][256];
    __shared__ unsigned long long red_s[kSelThreads / 32];
    __shared__ unsigned sel_digit, sel_need, sel_count;
    __shared__ unsigned long long kstar_s;
    const int q = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // list lengths -> exclusive prefix sum (one list per thread)
    const int mine = tid < grid_ctas ? min((int)cnt[(size_t)tid * b_pad + q], cap) : 0;
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    if (tid == 0) fill_s = 0;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    pre_s[tid + 1] = base + incl;
    if (tid == 0) pre_s[0] = 0;
    __syncthreads();
    const int total = pre_s[grid_ctas];
    const int n_prev = first_segment ? 0 : kp;
    uint64_t kstar = 0ull;

    if (n_prev + total > kSelMaxKeys) {
        __syncthreads();
        select_lists_body<kSelThreads>(lists, cnt, grid_ctas, b_pad, cap, kp, first_segment ? nullptr : short_keys, q, keys,
                                       pre_s, warp_tot, &fill_s);
        for (int i = tid; i < kp; i += kSelThreads) short_keys[(size_t)q * kp + i] = keys[i];  // sorted: the k'-th is last
        kstar = keys[kp - 1];
    } else {
        // gather: previous shortlist + every candidate that beats its k'-th key (0 while the shortlist is not full)
        const uint64_t floor_key = first_segment ? 0ull : short_keys[(size_t)q * kp + kp - 1];
        for (int i = tid; i < n_prev; i += kSelThreads) {
            const uint64_t key = short_keys[(size_t)q * kp + i];
            if (key != 0ull) keys[atomicAdd(&fill_s, 1)] = key;
        }
        for (int idx = tid; idx < total; idx += kSelThreads) {
            int lo = 0, hi = grid_ctas;  // list g with pre_s[g] <= idx < pre_s[g + 1]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (pre_s[mid] <= idx) lo = mid; else hi = mid;
            }
            const uint64_t key = __ldcg(lists + ((size_t)lo * b_pad + q) * cap + (idx - pre_s[lo]));
            if (key > floor_key) keys[atomicAdd(&fill_s, 1)] = key;
        }
        __syncthreads();
        const int n = fill_s;
        if (n < kp) {  // shortlist not full yet: keep everything, no floor
            for (int i = tid; i < kp; i += kSelThreads) short_keys[(size_t)q * kp + i] = i < n ? keys[i] : 0ull;
        } else {
            // bytes shared by every key carry no information: start below the highest differing bit
            uint64_t diff = 0ull;
            const uint64_t k0 = keys[0];
            for (int i = tid; i < n; i += kSelThreads) diff |= keys[i] ^ k0;
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) diff |= __shfl_xor_sync(0xffffffffu, diff, m);
            if (lane == 0) red_s[warp] = diff;
            __syncthreads();
            diff = 0ull;
#pragma unroll
            for (int w = 0; w < kSelThreads / 32; ++w) diff |= red_s[w];
            int shift = diff ? ((63 - __clzll((long long)diff)) & ~7) : 0;
            uint64_t prefix = (shift + 8 < 64) ? (k0 >> (shift + 8)) << (shift + 8) : 0ull;  // bits above the first digit
            unsigned need = (unsigned)kp;
            bool found = n == 1;
            if (n == 1) kstar = k0;
            while (!found) {
                for (int i = tid; i < (kSelThreads / 32) * 256; i += kSelThreads) (&hist[0][0])[i] = 0u;
                __syncthreads();
                for (int i = tid; i < n; i += kSelThreads) {
                    const uint64_t key = keys[i];
                    const bool in = shift + 8 >= 64 || ((key ^ prefix) >> (shift + 8)) == 0ull;
                    if (in) atomicAdd(&hist[warp][(unsigned)(key >> shift) & 255u], 1u);
                }
                __syncthreads();
                if (warp == 0) {  // digits in descending order: lane l owns 255 - 8l .. 248 - 8l
                    unsigned c[8], s = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int d = 255 - (8 * lane + j);
                        unsigned v = 0;
#pragma unroll
                        for (int w = 0; w < kSelThreads / 32; ++w) v += hist[w][d];
                        c[j] = v;
                        s += v;
                    }
                    unsigned inc = s;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const unsigned t = __shfl_up_sync(0xffffffffu, inc, off);
                        if (lane >= off) inc += t;
                    }
                    unsigned run = inc - s;
                    if (run < need && need <= inc) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (run < need && need <= run + c[j]) {
                                sel_digit = (unsigned)(255 - (8 * lane + j));
                                sel_need = need - run;
                                sel_count = c[j];
                            }
                            run += c[j];
                        }
                    }
                }
                __syncthreads();
                prefix |= (uint64_t)sel_digit << shift;
                need = sel_need;
                const unsigned count = sel_count;
                if (shift == 0) {
                    kstar = prefix;  // all 64 bits fixed (keys are unique, so count == 1 here)
                    found = true;
                } else if (count == 1) {  // one key left in the bin: it is the k'-th
                    for (int i = tid; i < n; i += kSelThreads)
                        if (((keys[i] ^ prefix) >> shift) == 0ull) kstar_s = keys[i];
                    __syncthreads();
                    kstar = kstar_s;
                    found = true;
                } else {
                    shift -= 8;
                }
                __syncthreads();
            }
            // keys above the k'-th -> slots [0, k'-1) in any order; the k'-th itself last
            if (tid == 0) fill_s = 0;
            __syncthreads();
            for (int i = tid; i < n; i += kSelThreads) {
                const uint64_t key = keys[i];
                if (key > kstar) short_keys[(size_t)q * kp + atomicAdd(&fill_s, 1)] = key;
            }
            if (tid == 0) short_keys[(size_t)q * kp + kp - 1] = kstar;
        }
    }
    if (tid == 0) {
        // every row discarded so far (epilogue filter, list compaction, this selection) scored <= the k'-th best
        uint32_t t = tau[q];
        if (kstar) t = max(t, (uint32_t)(kstar >> 32));
        tau[q] = t;
        tau_tilde[q] = ord_to_float(t);
    }
}

// After the cold pass (tc_score_kernel with ScoreArgs::cmax set): tau[q] = k'-th largest chunk maximum of query q.
__global__ void __launch_bounds__(128)
tau_from_chunk_maxima_kernel(const uint32_t* __restrict__ cmax, int ncols, int kp, uint32_t* __restrict__ tau) {
    __shared__ int s_cnt[8];
    const uint32_t* vals = cmax + (size_t)blockIdx.x * ncols;
    const int et = threadIdx.x;
    uint32_t v[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) v[i] = i * 128 + et < ncols ? vals[i * 128 + et] : 0u;
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
        __syncthreads();
        if (sc[0] + sc[1] + sc[2] + sc[3] >= kp) lo = cand;
    }
    if (et == 0) tau[blockIdx.x] = lo > kOrdNegInf ? lo : kOrdNegInf;
}

// K4: one warp per (query, shortlisted row): canonical fp32 dot -> exact key.
template <typename T>
__global__ void __launch_bounds__(256)
rescore_kernel(const T* __restrict__ rows, const float* __restrict__ qn, const uint64_t* __restrict__ short_keys,
               int B, int kp, uint64_t* __restrict__ exact_keys, int metric, const float* __restrict__ norms,
               const float* __restrict__ hq) {
    const int lane = threadIdx.x & 31;
    const size_t item = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (item >= (size_t)B * kp) return;
    const int q = (int)(item / kp);
    const uint64_t key = short_keys[item];
    if (key == 0ull) {
        if (lane == 0) exact_keys[item] = 0ull;
        return;
    }
    const uint32_t id = key_id(key);
    RowFrag<T> x;
    x.load(rows + (size_t)id * kDim, lane);
    float4 qv[kChunks], xv[kChunks];
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        qv[c] = *reinterpret_cast<const float4*>(qn + (size_t)q * kDim + c * 128 + lane * 4);
        xv[c] = x.get(c);
    }
    float s = warp_butterfly_sum(lane_dot(qv, xv));
    if (metric != MQVS_METRIC_COSINE) s = rank_score(metric, s, __ldg(norms + id), __ldg(hq + q));  // exact, as the SIMT scan
    if (lane == 0) exact_keys[item] = pack_key(s, id);
}

// Per query: sort the kp exact keys, emit top-k, verify the exactness margin.
__global__ void __launch_bounds__(kSelThreads)
finalize_kernel(const uint64_t* __restrict__ exact_keys, const float* __restrict__ tau_tilde, const float* __restrict__ eps,
                int kp, int k, int64_t id_base, float* __restrict__ out_scores, int64_t* __restrict__ out_ids,
                uint32_t* __restrict__ flags, uint32_t* __restrict__ flag_count, uint32_t* __restrict__ band) {
    __shared__ uint64_t buf[256];
    const int q = blockIdx.x;
    for (int i = threadIdx.x; i < kp; i += kSelThreads) buf[i] = exact_keys[(size_t)q * kp + i];
    block_bitonic_sort_desc(buf, kp);
    for (int j = threadIdx.x; j < k; j += kSelThreads) {
        const uint64_t key = j < kp ? buf[j] : 0ull;
        out_scores[(size_t)q * k + j] = key ? key_score(key) : -INFINITY;
        out_ids[(size_t)q * k + j] = key ? (int64_t)key_id(key) + id_base : -1;
    }
    if (threadIdx.x == 0) {
        const float tt = tau_tilde[q];
        bool ok = true;
        const uint64_t kth = (k <= kp) ? buf[k - 1] : 0ull;
        if (tt != -INFINITY)  // something was discarded: the k-th exact score must clear every discarded row's bound
            ok = kth != 0ull && key_score(kth) > tt + eps[q];
        flags[q] = ok ? 0u : 1u;
        if (!ok) {
            atomicAdd(flag_count, 1u);
            // threshold of the band pass (tc.h search_band): whatever can beat the k-th exact score scores at least this
            band[q] = kth ? ord_u32(__fsub_rd(key_score(kth), eps[q])) : kOrdNegInf;
        }
    }
}

// ------------------------------------------------------------------------------------------ host side
namespace {

PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

}  // namespace

// [rows, 768] row-major matrix, box = one 128-byte K-chunk x box_rows rows, 128B swizzle.
bool make_map(CUtensorMap* map, const void* base, uint64_t rows, int dtype, int box_rows) {
    auto enc = get_encode();
    if (!enc || rows == 0) return false;
    const bool f32 = dtype == MQVS_F32;
    cuuint64_t dims[2] = {(cuuint64_t)kDim, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kDim * (f32 ? 4 : 2)};
    cuuint32_t box[2] = {(cuuint32_t)(f32 ? 32 : 64), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// MQVS_TRACE=1: CUDA events after every launch of a tensor-core search, printed by trace_dump() (developer aid)
Trace::Trace() : on(getenv("MQVS_TRACE") != nullptr) {}
unsigned long long* Trace::fused_dbg(int grid) {
    if (!dbg && cudaMalloc(&dbg, 256 * 16 * sizeof(unsigned long long)) != cudaSuccess) dbg = nullptr;
    dbg_grid = grid;
    return dbg;
}
void Trace::mark(const char* what, cudaStream_t st) {
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    marks.emplace_back(what, e);
}
Trace& trace() {
    static Trace t;
    return t;
}

namespace {
Trace& g_trace = trace();

// CTA pairs: clusters of two CTAs, an even grid; `mxh` is the corpus map with 128-row boxes (each CTA stages half a tile)
template <typename T, int CAP, int KPRIME>
cudaError_t launch_score_pair(const CUtensorMap& mq, const CUtensorMap& mxh, const ScoreArgs& a, int grid, cudaStream_t st) {
    auto kern = a.norms != nullptr ? tc_score_pair_kernel<T, CAP, KPRIME, true> : tc_score_pair_kernel<T, CAP, KPRIME, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, mq, mxh, a);
}

template <typename T, int CAP, int KPRIME>
cudaError_t launch_score(const CUtensorMap& mq, const CUtensorMap& mx, const ScoreArgs& a, int grid, cudaStream_t st) {
    auto kern = a.norms != nullptr ? tc_score_kernel<T, CAP, KPRIME, true> : tc_score_kernel<T, CAP, KPRIME, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
    kern<<<grid, kThreads, kSmemBytes, st>>>(mq, mx, a);
    return cudaGetLastError();
}

}  // namespace

Plan plan(int k, int B, int sm_count) {
    Plan p{};
    // k' = 512 / 1152 (k <= 256 / 1024) exist in the single-launch pipeline only (wide shortlists: lists of 2048, no
    // in-kernel compaction; 1152 <= the 8 chunk maxima x 148 CTAs the thresholds are learnt from)
    p.kprime = k <= 10 ? 32 : (k <= 32 ? 64 : (k <= 64 ? 128 : (k <= 128 ? 256 : (k <= 256 ? 512 : 1152))));
    static const int kp_min = getenv("MQVS_KP_MIN") ? atoi(getenv("MQVS_KP_MIN")) : 0;  // developer knob: wider shortlists
    if ((kp_min == 64 || kp_min == 128 || kp_min == 256) && p.kprime <= 256) p.kprime = std::max(p.kprime, kp_min);
    p.cap = p.kprime > 256 ? 2048 : 512;  // per-(CTA, query) list capacity; narrow lists are compacted to their k' best near cap
    const int chunk = B < kMaxQueriesPerLaunch ? B : kMaxQueriesPerLaunch;
    p.b_pad = (chunk + BM - 1) / BM * BM;
    p.grid = sm_count;
    return p;
}

size_t op_bytes(const Plan& p, int dtype) { return (size_t)p.b_pad * kDim * (dtype == MQVS_F32 ? 4 : 2); }
size_t lists_bytes(const Plan& p) { return (size_t)p.grid * p.b_pad * p.cap * sizeof(uint64_t); }
size_t cnt_bytes(const Plan& p) { return (size_t)p.grid * p.b_pad * sizeof(uint32_t); }

int launches_per_search(const Plan& p, uint32_t n_rows, int B) {
    uint32_t tiles = (n_rows + BN - 1) / BN, segs = 0;
    const uint32_t m_tiles = (uint32_t)((std::min(B, kMaxQueriesPerLaunch) + BM - 1) / BM);
    for (uint64_t end = first_segment_end(m_tiles, p.grid); ; end = next_segment_end(end, m_tiles)) { ++segs; if (clamp_segment_end(end, tiles) >= tiles) break; }
    const int chunks = (B + kMaxQueriesPerLaunch - 1) / kMaxQueriesPerLaunch;
    return chunks * (int)(1 + 2 /* cold pass + thresholds */ + 2 * segs + 2);
}

const char* search(const Buffers& b, const Plan& p, const void* rows, int dtype, uint32_t n_rows, const float* qn_all, int B_all,
                   int k, const uint32_t* mask, int64_t id_base, float* out_scores_all, int64_t* out_ids_all, cudaStream_t st,
                   const Space& space) {
    CUtensorMap mq, mx, mxh;
    const bool xf = space.metric != MQVS_METRIC_COSINE;
    if (!make_map(&mq, b.q_op, (uint64_t)p.b_pad, dtype, BM)) return "cuTensorMapEncodeTiled failed for the query matrix";
    if (!make_map(&mx, rows, (uint64_t)n_rows, dtype, BN)) return "cuTensorMapEncodeTiled failed for the corpus";
    if (!make_map(&mxh, rows, (uint64_t)n_rows, dtype, BN / 2)) return "cuTensorMapEncodeTiled failed for the corpus (half tiles)";
    // CTA pairs (cta_group::2) once there is more than one query tile and the grid can be split into pairs
    static const bool no_pair = getenv("MQVS_NO_PAIR") != nullptr;  // developer knob: the single-CTA kernel, for comparison
    const bool pairs = !no_pair && B_all > BM && p.grid >= 2;
    const uint32_t n_tiles = (n_rows + BN - 1) / BN;
    if (cudaMemsetAsync(b.flag_count, 0, sizeof(uint32_t), st) != cudaSuccess) return "cudaMemsetAsync failed";

    g_trace.mark("begin", st);
    // queries are processed kMaxQueriesPerLaunch at a time (bounds the candidate-list memory)
    for (int q0 = 0; q0 < B_all; q0 += kMaxQueriesPerLaunch) {
        const int B = std::min(kMaxQueriesPerLaunch, B_all - q0);
        const int b_pad = (B + BM - 1) / BM * BM;
        const float* qn = qn_all + (size_t)q0 * kDim;
        if (dtype == MQVS_F32)
            prep_queries_kernel<float><<<p.b_pad, 32, 0, st>>>(qn, B, p.b_pad, (float*)b.q_op, b.eps + q0, b.tau, space.metric,
                                                               space.amax, xf ? space.qnorm + q0 : nullptr, xf ? b.hq + q0 : nullptr,
                                                               nullptr);
        else
            prep_queries_kernel<__nv_bfloat16><<<p.b_pad, 32, 0, st>>>(qn, B, p.b_pad, (__nv_bfloat16*)b.q_op, b.eps + q0, b.tau,
                                                                       space.metric, space.amax, xf ? space.qnorm + q0 : nullptr,
                                                                       xf ? b.hq + q0 : nullptr, space.exact_rows ? space.shadow_err : nullptr);

        g_trace.mark("prep", st);
        ScoreArgs a{};
        a.n_rows = n_rows;
        a.n_queries = (uint32_t)B;
        a.m_tiles = (uint32_t)(b_pad / BM);
        a.mask = mask;
        a.tau = b.tau;
        a.cnt = b.cnt;
        a.lists = b.lists;
        a.b_pad = (uint32_t)p.b_pad;
        a.cmax = nullptr;
        a.ncols = 0;
        a.norms = xf ? space.norms : nullptr;
        a.nbounds = xf ? reinterpret_cast<const float2*>(space.norm_bounds) : nullptr;
        a.hq = xf ? b.hq + q0 : nullptr;
        cudaError_t e = cudaSuccess;
#define MQVS_TC(KP_)                                                                                     \
    if (p.kprime == KP_) {                                                                               \
        if (pairs && grid >= 2)                                                                          \
            e = dtype == MQVS_F32 ? launch_score_pair<float, 512, KP_>(mq, mxh, a, grid & ~1, st)        \
                                  : launch_score_pair<__nv_bfloat16, 512, KP_>(mq, mxh, a, grid & ~1, st); \
        else                                                                                             \
            e = dtype == MQVS_F32 ? launch_score<float, 512, KP_>(mq, mx, a, grid, st)                   \
                                  : launch_score<__nv_bfloat16, 512, KP_>(mq, mx, a, grid, st);          \
    }
        // Cold pass: one tile per CTA, no list appends -- the epilogue only records the maximum of every 32-row chunk,
        // and the k'-th largest of a query's G*8 chunk maxima becomes its first threshold (>= k' distinct rows score that
        // much).  Costs one tile per CTA per query tile at tensor speed; the appending cold segment it replaces (4096 rows
        // on 16 CTAs, every score written out, then a selection over 4096 keys per query) took 2.0 ms at config 3.
        // The tiles of the cold pass are scored again by the first segment, now filtered.
        {
            const int grid = (int)std::min<uint32_t>((uint32_t)p.grid, n_tiles);
            a.tile_begin = 0;
            a.tile_end = (uint32_t)grid;
            a.cmax = b.cmax;
            a.ncols = (uint32_t)grid * 8;
            if (cudaMemsetAsync(b.cmax, 0, (size_t)p.b_pad * a.ncols * sizeof(uint32_t), st) != cudaSuccess) return "cudaMemsetAsync failed";
            MQVS_TC(32)
            MQVS_TC(64)
            MQVS_TC(128)
            MQVS_TC(256)
            if (e != cudaSuccess) return cudaGetErrorString(e);
            g_trace.mark("cold score", st);
            tau_from_chunk_maxima_kernel<<<B, 128, 0, st>>>(b.cmax, (int)a.ncols, p.kprime, b.tau);
            g_trace.mark("thresholds", st);
            a.cmax = nullptr;
            a.ncols = 0;
        }
        // Corpus segments of geometrically growing size.  After each one the per-query threshold becomes the k'-th
        // best score over EVERYTHING scored so far (select_lists_kernel), so a segment g times larger than all before
        // it appends only ~g k' candidates per query instead of re-learning its threshold in every CTA.
        // (function attributes are per DEVICE: a handle on another GPU of this process needs its own -- set it every time)
        cudaFuncSetAttribute(select_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSelMaxKeys * (int)sizeof(uint64_t));
        uint32_t begin = 0;
        for (uint64_t end = first_segment_end(a.m_tiles, p.grid); begin < n_tiles; end = next_segment_end(end, a.m_tiles)) {
            a.tile_begin = begin;
            a.tile_end = (uint32_t)clamp_segment_end(end, n_tiles);
            const int grid = (int)std::min<uint32_t>((uint32_t)p.grid, a.tile_end - a.tile_begin);
            if (cudaMemsetAsync(b.cnt, 0, (size_t)grid * p.b_pad * sizeof(uint32_t), st) != cudaSuccess) return "cudaMemsetAsync failed";
            MQVS_TC(32)
            MQVS_TC(64)
            MQVS_TC(128)
            MQVS_TC(256)
            if (e != cudaSuccess) return cudaGetErrorString(e);
            g_trace.mark("score", st);
            select_lists_kernel<<<B, kSelThreads, kSelMaxKeys * sizeof(uint64_t), st>>>(b.lists, b.cnt, b.tau, grid, p.b_pad, p.cap,
                                                                                         p.kprime, begin == 0, b.short_keys, b.tau_tilde);
            g_trace.mark("select", st);
            begin = a.tile_end;
        }
#undef MQVS_TC

        const size_t items = (size_t)B * p.kprime;
        const unsigned rs_grid = (unsigned)((items + 7) / 8);
        if (dtype == MQVS_F32 || space.exact_rows != nullptr)  // (scored over the bf16 shadow: the exact scores come from the fp32 rows)
            rescore_kernel<float><<<rs_grid, 256, 0, st>>>((const float*)(space.exact_rows ? space.exact_rows : rows), qn, b.short_keys, B,
                                                           p.kprime, b.exact_keys, space.metric, space.norms, b.hq + q0);
        else
            rescore_kernel<__nv_bfloat16><<<rs_grid, 256, 0, st>>>((const __nv_bfloat16*)rows, qn, b.short_keys, B, p.kprime,
                                                                   b.exact_keys, space.metric, space.norms, b.hq + q0);
        finalize_kernel<<<B, kSelThreads, 0, st>>>(b.exact_keys, b.tau_tilde, b.eps + q0, p.kprime, k, id_base,
                                                   out_scores_all + (size_t)q0 * k, out_ids_all + (size_t)q0 * k,
                                                   b.flags + q0, b.flag_count, b.band + q0);
        g_trace.mark("rescore+finalize", st);
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

void trace_dump() {
    if (!g_trace.on || g_trace.marks.empty()) return;
    cudaEventSynchronize(g_trace.marks.back().second);
    for (size_t i = 1; i < g_trace.marks.size(); ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, g_trace.marks[i - 1].second, g_trace.marks[i].second);
        fprintf(stderr, "[mqvs trace] %-18s %8.1f us\n", g_trace.marks[i].first, ms * 1e3f);
    }
    for (auto& m : g_trace.marks) cudaEventDestroy(m.second);
    g_trace.marks.clear();
    if (g_trace.dbg && g_trace.dbg_grid > 0) {
        // phase stamps of the last single-launch search, relative to the earliest CTA start: min / median / max over CTAs
        const int G = g_trace.dbg_grid;
        std::vector<unsigned long long> h((size_t)G * 16);
        if (cudaMemcpy(h.data(), g_trace.dbg, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
            unsigned long long t0 = ~0ull;
            for (int c = 0; c < G; ++c) t0 = std::min(t0, h[(size_t)c * 16]);
            static const char* names[11] = {"epilogue start", "first tile scored", "chunk maxima out", "barrier 1 passed",
                                            "tau selected", "barrier 2 passed", "scan done", "barrier 3 passed",
                                            "shortlist selected", "rescored", "results written"};
            for (int i = 0; i < 11; ++i) {
                std::vector<double> v;
                for (int c = 0; c < G; ++c)
                    if (h[(size_t)c * 16 + i] >= t0) v.push_back((double)(h[(size_t)c * 16 + i] - t0) * 1e-3);
                if (v.empty()) continue;
                std::sort(v.begin(), v.end());
                fprintf(stderr, "[mqvs trace]   %-18s min %8.1f  med %8.1f  max %8.1f us\n", names[i], v.front(), v[v.size() / 2], v.back());
            }
        }
        g_trace.dbg_grid = 0;
    }
}

}  // namespace tc
}  // namespace mqvs
